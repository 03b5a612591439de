#!/bin/bash
# Tries the four halo-kernel variants (shared-memory column pitch 10/16 rows x base-offset field on/off) on the conv
# operator tests, then A/Bs the bench for the variants that pass.  Usage (GPU box): bash scripts/halo_variants.sh
mkdir -p gpurun_out
pass=""
for hy in 10 16; do for bo in 1 0; do
  USB_HALO=1 USB_HALO_HY=$hy USB_HALO_BOFF=$bo timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "conv" > gpurun_out/halo_${hy}_${bo}.log 2>&1
  rc=$?
  echo "variant hy=$hy boff=$bo rc=$rc: $(tail -1 gpurun_out/halo_${hy}_${bo}.log)"
  if [ $rc -eq 0 ]; then pass="$pass $hy:$bo"; fi
done; done
echo "passing:$pass"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-vocoder > gpurun_out/bench_base.json 2>gpurun_out/bench_base.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_base.json')); print('base', d['value'], d['roofline']['achieved'], d['breakdown_ms_per_pass']['conv_igemm_ms'])
PY
for v in $pass; do
  hy=${v%%:*}; bo=${v##*:}
  USB_HALO=1 USB_HALO_HY=$hy USB_HALO_BOFF=$bo timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-vocoder > gpurun_out/bench_halo_${hy}_${bo}.json 2>gpurun_out/bench_halo_${hy}_${bo}.err
  python - "$hy" "$bo" <<'PY'
import json,sys
hy,bo=sys.argv[1:3]
try:
    d=json.load(open(f'gpurun_out/bench_halo_{hy}_{bo}.json')); print('halo',hy,bo, d['value'], d['roofline']['achieved'], d['breakdown_ms_per_pass']['conv_igemm_ms'])
except Exception as e: print('halo',hy,bo,'failed',e)
PY
done
