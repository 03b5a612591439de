#!/bin/bash
# round-2 GPU call B: GPU test suite, bench + reference arm, launch lists, compact ncu captures (outputs kept < 40 MB)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_gputest.log
tail -4 gpurun_out/r2b_gputest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
P32="python scripts/profile_pass.py --batch 32 --frames 1000 --steps 2"
$P32 > gpurun_out/r2b_plain_32x1000.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -s 244 -c 262 --csv --log-file gpurun_out/r2b_launches_32x1000.csv $P32 > gpurun_out/r2b_ncu1.log 2>&1; echo "ncu1 rc=$?"
P1="python scripts/profile_pass.py --batch 1 --frames 256 --steps 2"
$P1 > gpurun_out/r2b_plain_1x256.log 2>&1 &&
timeout 600 ncu --metrics $M --clock-control none -s 244 -c 262 --csv --log-file gpurun_out/r2b_launches_1x256.csv $P1 > gpurun_out/r2b_ncu2.log 2>&1; echo "ncu2 rc=$?"
# --set full as CSV (raw page) for one whole step of every kernel class at 32 x 1000; .ncu-rep (with source) only for 3 + 3 launches
timeout 900 ncu --set full --clock-control none -k regex:conv_igemm_swapped -s 52 -c 26 --csv --page raw --log-file gpurun_out/r2b_swapped_T1000_full.csv $P32 > gpurun_out/r2b_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:conv_igemm_halo -s 24 -c 12 --csv --page raw --log-file gpurun_out/r2b_halo_T1000_full.csv $P32 > gpurun_out/r2b_ncu4.log 2>&1; echo "ncu4 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:gn_apply -s 64 -c 32 --csv --page raw --log-file gpurun_out/r2b_gn_apply_T1000_full.csv $P32 > gpurun_out/r2b_ncu5.log 2>&1; echo "ncu5 rc=$?"
timeout 900 ncu --set full --clock-control none -k "regex:final_kernel|first_conv|attn_partial|attn_fold|attn_merge|conv_igemm_kernel" -s 56 -c 48 --csv --page raw --log-file gpurun_out/r2b_other_T1000_full.csv $P32 > gpurun_out/r2b_ncu6.log 2>&1; echo "ncu6 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_halo -s 24 -c 2 -o gpurun_out/r2b_halo_T1000 $P32 > gpurun_out/r2b_ncu7.log 2>&1; echo "ncu7 rc=$?"
du -sh gpurun_out; ls -la gpurun_out | tail -30
