#!/bin/bash
# round-2 GPU call AB: gn_apply with the exponent scale folded into the affine, unconditional mask multiply, additive
# addressing (16.6 -> ~13 instructions per element): headline bench first, then the full GPU suite
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2ab_bench_head.json 2> gpurun_out/r2ab_bench_head.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ab_bench_head.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
PY
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2ab_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2ab_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2ab_gputest.log | tail -6
