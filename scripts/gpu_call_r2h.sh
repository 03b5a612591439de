#!/bin/bash
# round-2 GPU call H: validate the double-buffered Snake kernel; vocoder timing; swapped-conv ncu --set full at T = 1000
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_long.py -m gpu -q -x -k "vocoder or snake" > gpurun_out/r2h_gputest_voc.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2h_gputest_voc.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2h_gputest_voc.log | tail -5
timeout 300 python scripts/vocoder_time.py > gpurun_out/r2h_vocoder.json 2> gpurun_out/r2h_vocoder.err; cat gpurun_out/r2h_vocoder.json
python scripts/act_shapes.py > gpurun_out/r2h_act_plain.log 2>&1; cat gpurun_out/r2h_act_plain.log | tail -3
timeout 600 ncu --set full --clock-control none -k regex:snake_act --launch-skip 3 --launch-count 3 --csv --page raw --log-file gpurun_out/r2h_snake2_full.csv python scripts/act_shapes.py > gpurun_out/r2h_ncu1.log 2>&1; echo "ncu1 rc=$?"
P32="python scripts/profile_pass.py --batch 32 --frames 1000 --steps 2"
$P32 > gpurun_out/r2h_plain_32x1000.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:conv_igemm_swapped -s 52 -c 26 --csv --page raw --log-file gpurun_out/r2h_swapped_T1000_full.csv $P32 > gpurun_out/r2h_ncu3.log 2>&1; echo "ncu3 rc=$?"
du -sh gpurun_out
