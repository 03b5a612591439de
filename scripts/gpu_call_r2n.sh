#!/bin/bash
# round-2 GPU call N: 1-D halo conv kernel for the vocoder (conv1d_halo_kernel): parity first, then A/B timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/r2n_voc_tests.log 2>&1; rc=$?; echo "voc+ops tests rc=$rc"; tail -8 gpurun_out/r2n_voc_tests.log
timeout 600 python -m pytest tests/test_gpu_long.py -m gpu -q -x -k vocoder > gpurun_out/r2n_voc_long.log 2>&1; echo "voc long rc=$?"; tail -4 gpurun_out/r2n_voc_long.log
for mode in 0 1 2; do
  USB_H1D=$mode timeout 300 python scripts/vocoder_time.py > gpurun_out/r2n_vocoder_h1d$mode.json 2> gpurun_out/r2n_vocoder_h1d$mode.err; echo "USB_H1D=$mode rc=$?"; cat gpurun_out/r2n_vocoder_h1d$mode.json
done
USB_H1D_NO_WRES=1 timeout 300 python scripts/vocoder_time.py > gpurun_out/r2n_vocoder_nowres.json 2> gpurun_out/r2n_vocoder_nowres.err; echo "no-wres rc=$?"; cat gpurun_out/r2n_vocoder_nowres.json
tail -3 gpurun_out/r2n_vocoder_h1d1.err
du -sh gpurun_out
