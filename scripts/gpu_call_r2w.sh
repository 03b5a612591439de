#!/bin/bash
# round-2 GPU call W: barrier addresses pinned in registers (no S2UR in the producer / MMA-issue loops): timing, suite, bench
mkdir -p gpurun_out
timeout 300 python scripts/vocoder_time.py > gpurun_out/r2w_vocoder.json 2> gpurun_out/r2w_vocoder.err; echo "rc=$?"; cat gpurun_out/r2w_vocoder.json
timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2w_bench_head.json 2> gpurun_out/r2w_bench_head.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2w_bench_head.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
PY
timeout 300 python scripts/latency_probe.py > gpurun_out/r2w_latency.jsonl 2> gpurun_out/r2w_latency.err; tail -2 gpurun_out/r2w_latency.jsonl
timeout 300 python scripts/finetune_time.py 50 > gpurun_out/r2w_finetune.json 2> gpurun_out/r2w_finetune.err; cat gpurun_out/r2w_finetune.json
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2w_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2w_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2w_gputest.log | tail -6
du -sh gpurun_out
