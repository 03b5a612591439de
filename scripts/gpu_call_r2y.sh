#!/bin/bash
# round-2 GPU call Y: evidence from the final tree: driver-style bench line, ncu launch lists (32 x 1000 sampler step, vocoder
# forward), ncu --set full of two conv1d_halo launches, one-utterance latency
mkdir -p gpurun_out
t0=$(date +%s); timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_bench_driver_style.json 2> gpurun_out/r2y_bench_driver_style.err; rc=$?; t1=$(date +%s); echo "bench rc=$rc wall=$((t1-t0)) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2y_bench_driver_style.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3),'whole',round(d['roofline']['whole_step_frac'],3),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
for k in ('secondary_16x512','latency_stage'): print(k, d[k]['ms_per_pass'], d[k]['whole_pass_frac'])
print('voc', d['vocoder_stage']['ms'], d['vocoder_stage']['conv']['frac'], d['vocoder_stage']['snake_act']['frac'], 'ft', d['finetune_stage']['ms_per_iter'])
PY
P32="python scripts/profile_pass.py --batch 32 --frames 1000 --steps 2"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
$P32 > gpurun_out/r2y_plain_32x1000.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -s 244 -c 262 --csv --log-file gpurun_out/r2y_launches_32x1000.csv $P32 > gpurun_out/r2y_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 300 python scripts/vocoder_time.py > gpurun_out/r2y_vocoder_time.json 2> gpurun_out/r2y_vocoder_time.err; cat gpurun_out/r2y_vocoder_time.json
timeout 900 ncu --metrics $M --clock-control none --launch-skip 476 -c 250 --csv --log-file gpurun_out/r2y_vocoder_launches.csv python scripts/vocoder_time.py --iters 1 > gpurun_out/r2y_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:conv1d_halo --launch-skip 192 -c 1 --csv --page raw --log-file gpurun_out/r2y_h1d_s3k11_full.csv python scripts/vocoder_time.py --iters 1 > gpurun_out/r2y_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:conv1d_halo --launch-skip 237 -c 1 --csv --page raw --log-file gpurun_out/r2y_h1d_s6k11_full.csv python scripts/vocoder_time.py --iters 1 > gpurun_out/r2y_ncu4.log 2>&1; echo "ncu4 rc=$?"
timeout 300 python scripts/latency_probe.py > gpurun_out/r2y_latency.jsonl 2> gpurun_out/r2y_latency.err; cat gpurun_out/r2y_latency.jsonl
du -sh gpurun_out
