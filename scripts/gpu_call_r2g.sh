#!/bin/bash
# round-2 GPU call G: validation of the final conv epilogues + bench + evidence
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2g_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2g_gputest.log | tail -8
SWEEP=epi timeout 300 python scripts/conv_sweep.py up_l1_T1000 up_l2_T1000 down_l0_T1000 down_l1_T1000 l0_128_T1000 > gpurun_out/r2g_sweep_epi.log 2>&1; cat gpurun_out/r2g_sweep_epi.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3),'whole',round(d['roofline']['whole_step_frac'],3),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
for k in ('secondary_16x512','latency_stage'): print(k, d[k]['ms_per_pass'], d[k]['whole_pass_frac'])
print('voc', d['vocoder_stage']['ms'], 'ft', d['finetune_stage']['ms_per_iter'])
PY
P32="python scripts/profile_pass.py --batch 32 --frames 1000 --steps 2"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
$P32 > gpurun_out/r2g_plain_32x1000.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -s 244 -c 262 --csv --log-file gpurun_out/r2g_launches_32x1000.csv $P32 > gpurun_out/r2g_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 300 python scripts/latency_probe.py > gpurun_out/r2g_latency_1x256.jsonl 2> gpurun_out/r2g_latency.err; cat gpurun_out/r2g_latency_1x256.jsonl
du -sh gpurun_out
