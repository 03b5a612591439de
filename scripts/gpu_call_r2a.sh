#!/bin/bash
# round-2 GPU call A: full GPU test suite (incl. the long-config parity tests), the new bench line, the reference arm,
# launch lists (32 x 1000 and 1 x 256) and --set full captures of the conv kernels at T = 1000
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/r2a_clocks.csv &
SMI=$!
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_gputest.log
tail -5 gpurun_out/r2a_gputest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"
kill $SMI
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
python scripts/profile_pass.py --batch 32 --frames 1000 --steps 2 > gpurun_out/r2a_plain_32x1000.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -s 244 -c 262 --csv --log-file gpurun_out/r2a_launches_32x1000.csv python scripts/profile_pass.py --batch 32 --frames 1000 --steps 2 > gpurun_out/r2a_ncu1.log 2>&1; echo "ncu1 rc=$?"
python scripts/profile_pass.py --batch 1 --frames 256 --steps 2 > gpurun_out/r2a_plain_1x256.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 244 -c 262 --csv --log-file gpurun_out/r2a_launches_1x256.csv python scripts/profile_pass.py --batch 1 --frames 256 --steps 2 > gpurun_out/r2a_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_halo -s 24 -c 6 -o gpurun_out/r2a_halo_T1000 python scripts/profile_pass.py --batch 32 --frames 1000 --steps 2 > gpurun_out/r2a_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_swapped -s 52 -c 26 -o gpurun_out/r2a_swapped_T1000 python scripts/profile_pass.py --batch 32 --frames 1000 --steps 2 > gpurun_out/r2a_ncu4.log 2>&1; echo "ncu4 rc=$?"
ls -la gpurun_out | tail -20
