#!/bin/bash
# round-2 GPU call AA: op-level parity of the generator's Conv1d layer (usb_op_conv1d) + the vocoder tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r2aa_voc_tests.log 2>&1; rc=$?; echo "voc tests rc=$rc"; tail -15 gpurun_out/r2aa_voc_tests.log
