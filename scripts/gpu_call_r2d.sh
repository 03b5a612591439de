#!/bin/bash
# round-2 GPU call D: validate the packed-fp32x2 elementwise kernels, A/B the first conv, conv sweeps for the strided /
# transposed convs (swapped vs plain kernel, with / without epilogue), ncu of the new Snake kernel and of gn_apply
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2d_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2d_gputest.log | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2d_bench_head.json 2> gpurun_out/r2d_bench_head.err; echo "bench rc=$?"
USB_FIRST_MINB=1 timeout 600 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/r2d_bench_head_minb1.json 2> gpurun_out/r2d_bench_head_minb1.err; echo "bench minb1 rc=$?"
SWEEP=swap timeout 300 python scripts/conv_sweep.py up_l1_T1000 up_l2_T1000 up_l3_T1000 down_l0_T1000 down_l1_T1000 > gpurun_out/r2d_sweep_swap.log 2>&1; cat gpurun_out/r2d_sweep_swap.log
SWEEP=epi timeout 300 python scripts/conv_sweep.py up_l1_T1000 up_l2_T1000 up_l3_T1000 down_l0_T1000 down_l1_T1000 l0_128_T1000 > gpurun_out/r2d_sweep_epi.log 2>&1; cat gpurun_out/r2d_sweep_epi.log
python scripts/act_shapes.py > gpurun_out/r2d_act_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none -k regex:snake_act --launch-skip 3 --launch-count 3 --csv --page raw --log-file gpurun_out/r2d_snake2_full.csv python scripts/act_shapes.py > gpurun_out/r2d_ncu1.log 2>&1; echo "ncu1 rc=$?"
P32="python scripts/profile_pass.py --batch 32 --frames 1000 --steps 2"
$P32 > gpurun_out/r2d_plain_32x1000.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k "regex:gn_apply|final_kernel|first_conv" -s 32 -c 12 --csv --page raw --log-file gpurun_out/r2d_ew_T1000_full.csv $P32 > gpurun_out/r2d_ncu2.log 2>&1; echo "ncu2 rc=$?"
du -sh gpurun_out
