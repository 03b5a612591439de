#!/bin/bash
# round-2 GPU call X: Snake kernel with 25 outputs per run: parity + timing (compare act_ms with 13.1-13.3 ms at 19)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r2x_voc_tests.log 2>&1; rc=$?; echo "voc tests rc=$rc"; tail -3 gpurun_out/r2x_voc_tests.log
for i in 1 2; do timeout 300 python scripts/vocoder_time.py > gpurun_out/r2x_vocoder$i.json 2> gpurun_out/r2x_vocoder$i.err; echo "rc=$?"; cat gpurun_out/r2x_vocoder$i.json; done
timeout 300 python scripts/finetune_time.py 50 > gpurun_out/r2x_finetune.json 2> gpurun_out/r2x_finetune.err; cat gpurun_out/r2x_finetune.json
