#!/bin/bash
# round-2 GPU call V: one-pass stage mean (parity + timing); experiment: conv1d_halo streaming layers without weight traffic
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r2v_voc_tests.log 2>&1; rc=$?; echo "voc tests rc=$rc"; tail -3 gpurun_out/r2v_voc_tests.log
timeout 600 python -m pytest tests/test_gpu_long.py -m gpu -q -x -k vocoder > gpurun_out/r2v_voc_long.log 2>&1; echo "voc long rc=$?"; tail -2 gpurun_out/r2v_voc_long.log
timeout 300 python scripts/vocoder_time.py > gpurun_out/r2v_vocoder.json 2> gpurun_out/r2v_vocoder.err; echo "rc=$?"; cat gpurun_out/r2v_vocoder.json
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --launch-skip 476 -c 250 --csv --log-file gpurun_out/r2v_vocoder_launches.csv python scripts/vocoder_time.py --iters 1 > gpurun_out/r2v_ncu.log 2>&1; echo "ncu rc=$?"
USB_DBG_FLAGS=1 timeout 900 ncu --metrics $M --clock-control none --launch-skip 476 -c 250 --csv --log-file gpurun_out/r2v_vocoder_launches_noW.csv python scripts/vocoder_time.py --iters 1 > gpurun_out/r2v_ncu2.log 2>&1; echo "ncu2 rc=$?"
du -sh gpurun_out
