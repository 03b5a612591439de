#!/bin/bash
# round-2 GPU call R: epilogue changes of the pixels-on-M kernels (groups alternate on BN = 64 tiles, residual prefetched):
# full GPU suite, vocoder timing + launch list, headline bench
mkdir -p gpurun_out
timeout 300 python scripts/vocoder_time.py > gpurun_out/r2r_vocoder.json 2> gpurun_out/r2r_vocoder.err; echo "rc=$?"; cat gpurun_out/r2r_vocoder.json
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2r_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2r_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2r_gputest.log | tail -6
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --launch-skip 488 -c 260 --csv --log-file gpurun_out/r2r_vocoder_launches.csv python scripts/vocoder_time.py --iters 1 > gpurun_out/r2r_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2r_bench_head.json 2> gpurun_out/r2r_bench_head.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench_head.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'conv',round(d['roofline']['achieved']),'gnGB',round(d['roofline_hbm']['achieved']), {k:round(v) for k,v in d['breakdown_ms_per_pass'].items() if v})
PY
du -sh gpurun_out
