#!/bin/bash
# round-2 GPU call AE: conv1d_halo_kernel with two taps per weight-ring stage: parity, then A/B (one tap / two taps where the
# ring holds >= 6 tiles / two taps where it holds >= 4)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r2ae_voc_tests.log 2>&1; rc=$?; echo "voc tests rc=$rc"; tail -3 gpurun_out/r2ae_voc_tests.log
USB_H1D_WT2_MIN=4 timeout 600 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x -k "conv1d or golden" > gpurun_out/r2ae_voc_tests4.log 2>&1; echo "voc tests (min 4) rc=$?"; tail -2 gpurun_out/r2ae_voc_tests4.log
for v in "USB_H1D_WT1=1" "USB_H1D_WT2_MIN=6" "USB_H1D_WT2_MIN=4" "USB_H1D_WT1=1" "USB_H1D_WT2_MIN=6" "USB_H1D_WT2_MIN=4"; do
  env $v timeout 300 python scripts/vocoder_time.py 2> /dev/null | head -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'conv_ms', round(d['conv_ms'],2), 'act_ms', round(d['act_ms'],2))"
done
