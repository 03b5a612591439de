#!/bin/bash
# round-2 GPU call U: 1-D residual layers on the swapped-operand kernel: parity, A/B timing, full suite, headline bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r2u_voc_tests.log 2>&1; rc=$?; echo "voc tests rc=$rc"; tail -3 gpurun_out/r2u_voc_tests.log
for i in 1 2; do
  timeout 300 python scripts/vocoder_time.py > gpurun_out/r2u_vocoder_swapres$i.json 2> gpurun_out/r2u_vocoder_swapres$i.err; echo "swap-res rc=$?"; cat gpurun_out/r2u_vocoder_swapres$i.json | cut -c1-130
  USB_SWAP_RES=0 timeout 300 python scripts/vocoder_time.py > gpurun_out/r2u_vocoder_h1dres$i.json 2> gpurun_out/r2u_vocoder_h1dres$i.err; echo "h1d-res rc=$?"; cat gpurun_out/r2u_vocoder_h1dres$i.json | cut -c1-130
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2u_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2u_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2u_gputest.log | tail -6
timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2u_bench_head.json 2> gpurun_out/r2u_bench_head.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u_bench_head.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
PY
du -sh gpurun_out
