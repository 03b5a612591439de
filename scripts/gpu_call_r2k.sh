#!/bin/bash
# round-2 GPU call K: cluster-of-2 TMA-multicast swapped conv kernel: op tests first, then everything, then A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/r2k_ops.log 2>&1; rc=$?; echo "ops rc=$rc"; tail -5 gpurun_out/r2k_ops.log
if [ $rc -ne 0 ]; then echo "op tests failed: stopping"; exit 0; fi
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_ops.py > gpurun_out/r2k_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2k_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2k_gputest.log | tail -6
timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2k_bench_head_mc.json 2> gpurun_out/r2k_bench_head_mc.err; echo "bench mc rc=$?"
USB_NO_MC=1 timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2k_bench_head_nomc.json 2> gpurun_out/r2k_bench_head_nomc.err; echo "bench nomc rc=$?"
for f in gpurun_out/r2k_bench_head_mc.json gpurun_out/r2k_bench_head_nomc.json; do python - $f <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1],'value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
PY
done
timeout 300 python scripts/latency_probe.py > gpurun_out/r2k_latency_mc.jsonl 2> gpurun_out/r2k_latency.err; tail -2 gpurun_out/r2k_latency_mc.jsonl
USB_NO_MC=1 timeout 300 python scripts/latency_probe.py > gpurun_out/r2k_latency_nomc.jsonl 2>> gpurun_out/r2k_latency.err; tail -2 gpurun_out/r2k_latency_nomc.jsonl
timeout 300 python scripts/finetune_time.py 50 > gpurun_out/r2k_finetune.json 2> gpurun_out/r2k_finetune.err; cat gpurun_out/r2k_finetune.json
timeout 300 python scripts/vocoder_time.py > gpurun_out/r2k_vocoder.json 2> gpurun_out/r2k_vocoder.err; tail -1 gpurun_out/r2k_vocoder.json
du -sh gpurun_out
