#!/bin/bash
# round-2 GPU call L: final validation of the committed state (all GPU tests incl. the multicast subprocess test, smoke, the
# bench line and the reference arm exactly as the driver runs them)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2l_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2l_gputest.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2l_smoke.log
/usr/bin/time -v timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2l_bench_driver_style.json 2> gpurun_out/r2l_bench_driver_style.err; echo "bench rc=$?"; grep "Elapsed (wall" gpurun_out/r2l_bench_driver_style.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench_driver_style.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3),'whole',round(d['roofline']['whole_step_frac'],3),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
for k in ('secondary_16x512','latency_stage'): print(k, d[k]['ms_per_pass'], d[k]['whole_pass_frac'])
print('voc', d['vocoder_stage']['ms'], 'ft', d['finetune_stage']['ms_per_iter'], 'build', d['build'])
PY
/usr/bin/time -v timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2l_bench_ref_driver_style.json 2> gpurun_out/r2l_bench_ref_driver_style.err; echo "ref rc=$?"; grep "Elapsed (wall" gpurun_out/r2l_bench_ref_driver_style.err; cut -c1-200 gpurun_out/r2l_bench_ref_driver_style.json
du -sh gpurun_out
