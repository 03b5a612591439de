#!/bin/bash
# round-2 GPU call O: per-launch times of one vocoder forward with the 1-D halo kernel (ncu launch list)
mkdir -p gpurun_out
timeout 300 python scripts/vocoder_time.py --iters 1 > gpurun_out/r2o_vocoder.json 2> gpurun_out/r2o_vocoder.err; echo "plain run rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 488 -c 260 --csv --log-file gpurun_out/r2o_vocoder_launches.csv python scripts/vocoder_time.py --iters 1 > gpurun_out/r2o_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2o_ncu.log
du -sh gpurun_out
