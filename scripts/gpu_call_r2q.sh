#!/bin/bash
# round-2 GPU call Q: conv1d_halo_kernel with the unrolled MMA issue loop: parity, timing, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r2q_voc_tests.log 2>&1; rc=$?; echo "voc tests rc=$rc"; tail -3 gpurun_out/r2q_voc_tests.log
timeout 600 python -m pytest tests/test_gpu_long.py -m gpu -q -x -k vocoder > gpurun_out/r2q_voc_long.log 2>&1; echo "voc long rc=$?"; tail -2 gpurun_out/r2q_voc_long.log
timeout 300 python scripts/vocoder_time.py > gpurun_out/r2q_vocoder.json 2> gpurun_out/r2q_vocoder.err; echo "rc=$?"; cat gpurun_out/r2q_vocoder.json
USB_H1D=2 timeout 300 python scripts/vocoder_time.py > gpurun_out/r2q_vocoder_h1d2.json 2> gpurun_out/r2q_vocoder_h1d2.err; echo "H1D=2 rc=$?"; cat gpurun_out/r2q_vocoder_h1d2.json
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --launch-skip 488 -c 260 --csv --log-file gpurun_out/r2q_vocoder_launches.csv python scripts/vocoder_time.py --iters 1 > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc=$?"
du -sh gpurun_out
