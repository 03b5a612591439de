#!/bin/bash
# round-2 GPU call P: ncu --set full of three conv1d_halo launches of a vocoder forward (stage 6 k=11, stage 6 k=3, stage 3 k=11)
mkdir -p gpurun_out
timeout 300 python scripts/vocoder_time.py --iters 1 > gpurun_out/r2p_vocoder.json 2> gpurun_out/r2p_vocoder.err; echo "plain run rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv1d_halo --launch-skip 237 -c 1 -o gpurun_out/r2p_h1d_s6k11 python scripts/vocoder_time.py --iters 1 > gpurun_out/r2p_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:conv1d_halo --launch-skip 225 -c 1 --csv --page raw --log-file gpurun_out/r2p_h1d_s6k3_full.csv python scripts/vocoder_time.py --iters 1 > gpurun_out/r2p_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:conv1d_halo --launch-skip 192 -c 1 --csv --page raw --log-file gpurun_out/r2p_h1d_s3k11_full.csv python scripts/vocoder_time.py --iters 1 > gpurun_out/r2p_ncu3.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out/r2p_*; du -sh gpurun_out
