#!/bin/bash
# round-2 GPU call E: validate PDL / wide stores / packed Snake; A/B runs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2e_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2e_gputest.log | tail -8
timeout 300 python scripts/latency_probe.py > gpurun_out/r2e_latency_1x256_pdl.jsonl 2> gpurun_out/r2e_latency.err; cat gpurun_out/r2e_latency_1x256_pdl.jsonl
USB_NO_PDL=1 timeout 300 python scripts/latency_probe.py > gpurun_out/r2e_latency_1x256_nopdl.jsonl 2>> gpurun_out/r2e_latency.err; cat gpurun_out/r2e_latency_1x256_nopdl.jsonl
SWEEP=epi timeout 300 python scripts/conv_sweep.py up_l1_T1000 up_l2_T1000 up_l3_T1000 down_l0_T1000 down_l1_T1000 l0_128_T1000 > gpurun_out/r2e_sweep_epi_wide.log 2>&1; cat gpurun_out/r2e_sweep_epi_wide.log
timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2e_bench_head.json 2> gpurun_out/r2e_bench_head.err; echo "bench rc=$?"
USB_NARROW_STORE=1 timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2e_bench_head_narrow.json 2> gpurun_out/r2e_bench_head_narrow.err; echo "bench narrow rc=$?"
USB_GN_U8=1 timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2e_bench_head_gnu8.json 2> gpurun_out/r2e_bench_head_gnu8.err; echo "bench gnu8 rc=$?"
timeout 300 python scripts/vocoder_time.py > gpurun_out/r2e_vocoder.json 2> gpurun_out/r2e_vocoder.err; cat gpurun_out/r2e_vocoder.json
timeout 300 python scripts/finetune_time.py 50 > gpurun_out/r2e_finetune.json 2> gpurun_out/r2e_finetune.err; cat gpurun_out/r2e_finetune.json
for f in gpurun_out/r2e_bench_head*.json; do python - $f <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1],'value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
PY
done
du -sh gpurun_out
