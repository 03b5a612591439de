#!/bin/bash
# round-2 GPU call M: the bench line and the reference arm exactly as the driver runs them (wall time recorded)
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench_driver_style.json 2> gpurun_out/r2m_bench_driver_style.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench_driver_style.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3),'whole',round(d['roofline']['whole_step_frac'],3),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
for k in ('secondary_16x512','latency_stage'): print(k, d[k]['ms_per_pass'], d[k]['whole_pass_frac'])
print('voc', d['vocoder_stage']['ms'], 'ft', d['finetune_stage']['ms_per_iter'], 'build', d['build'])
PY
t0=$(date +%s)
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench_ref_driver_style.json 2> gpurun_out/r2m_bench_ref_driver_style.err; echo "ref rc=$? wall=$(( $(date +%s) - t0 )) s"; cut -c1-200 gpurun_out/r2m_bench_ref_driver_style.json
