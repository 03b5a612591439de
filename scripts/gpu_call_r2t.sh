#!/bin/bash
# round-2 GPU call T: Snake kernel A/B on one box: six symmetric taps + branch-free interior groups vs the previous kernel
mkdir -p gpurun_out
L=unitspeech_b200/lib
timeout 600 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r2t_voc_tests.log 2>&1; rc=$?; echo "voc tests rc=$rc"; tail -3 gpurun_out/r2t_voc_tests.log
for i in 1 2; do
  timeout 300 python scripts/vocoder_time.py > gpurun_out/r2t_vocoder_new$i.json 2> gpurun_out/r2t_vocoder_new$i.err; echo "new rc=$?"; head -1 gpurun_out/r2t_vocoder_new$i.json | cut -c1-200
  USB_SNAKE_OCC3=1 timeout 300 python scripts/vocoder_time.py > gpurun_out/r2t_vocoder_occ3_$i.json 2> gpurun_out/r2t_vocoder_occ3_$i.err; echo "occ3 rc=$?"; head -1 gpurun_out/r2t_vocoder_occ3_$i.json | cut -c1-200
  cp $L/libunitspeech_b200.so /tmp/cur.so; cp $L/libunitspeech_b200_prevsnake.so $L/libunitspeech_b200.so
  timeout 300 python scripts/vocoder_time.py > gpurun_out/r2t_vocoder_prev$i.json 2> gpurun_out/r2t_vocoder_prev$i.err; echo "prev rc=$?"; head -1 gpurun_out/r2t_vocoder_prev$i.json | cut -c1-200
  cp /tmp/cur.so $L/libunitspeech_b200.so
done
du -sh gpurun_out
