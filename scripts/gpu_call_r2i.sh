#!/bin/bash
# round-2 GPU call I: final validation (all GPU tests, smoke, bench) after the final_kernel / Snake changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2i_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2i_gputest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2i_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3),'whole',round(d['roofline']['whole_step_frac'],3),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
for k in ('secondary_16x512','latency_stage'): print(k, d[k]['ms_per_pass'], d[k]['whole_pass_frac'])
print('voc', d['vocoder_stage']['ms'], 'ft', d['finetune_stage']['ms_per_iter'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err; echo "ref rc=$?"
du -sh gpurun_out
