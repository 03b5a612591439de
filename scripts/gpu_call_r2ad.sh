#!/bin/bash
# round-2 GPU call AD: the default bench exactly as the driver runs it, on the final tree
mkdir -p gpurun_out
t0=$(date +%s); timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ad_bench_driver_style.json 2> gpurun_out/r2ad_bench_driver_style.err; rc=$?; t1=$(date +%s); echo "bench rc=$rc wall=$((t1-t0)) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ad_bench_driver_style.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3),'whole',round(d['roofline']['whole_step_frac'],3),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
for k in ('secondary_16x512','latency_stage'): print(k, d[k]['ms_per_pass'], d[k]['whole_pass_frac'])
print('voc', d['vocoder_stage']['ms'], d['vocoder_stage']['conv']['frac'], d['vocoder_stage']['snake_act']['frac'], 'ft', d['finetune_stage']['ms_per_iter'])
print('cpu', d['cpu_baseline']['value'], 'eager', d.get('gpu_eager_baseline'))
PY
