#!/bin/bash
# round-2 GPU call S: Snake kernel with six symmetric taps (112 registers); experiment: three blocks per SM; ncu source view
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r2s_voc_tests.log 2>&1; rc=$?; echo "voc tests rc=$rc"; tail -3 gpurun_out/r2s_voc_tests.log
timeout 300 python scripts/vocoder_time.py > gpurun_out/r2s_vocoder.json 2> gpurun_out/r2s_vocoder.err; echo "rc=$?"; cat gpurun_out/r2s_vocoder.json
USB_SNAKE_OCC3=1 timeout 300 python scripts/vocoder_time.py > gpurun_out/r2s_vocoder_occ3.json 2> gpurun_out/r2s_vocoder_occ3.err; echo "occ3 rc=$?"; cat gpurun_out/r2s_vocoder_occ3.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:snake_act2 --launch-skip 270 -c 1 -o gpurun_out/r2s_snake2_s3 python scripts/vocoder_time.py --iters 1 > gpurun_out/r2s_ncu1.log 2>&1; echo "ncu1 rc=$?"
ls -la gpurun_out/r2s_*; du -sh gpurun_out
