#!/bin/bash
# round-2 GPU call AF: gn_apply with a 32-bit mask index (addressing only, same arithmetic): op + decoder parity, headline bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decoder.py -m gpu -q -x > gpurun_out/r2af_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2af_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2af_bench_head.json 2> gpurun_out/r2af_bench_head.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2af_bench_head.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
PY
