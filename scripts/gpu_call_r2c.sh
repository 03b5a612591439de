#!/bin/bash
# round-2 GPU call C: validate split-K fix / snake v2 / gn_apply rewrite / attention fold variants (full GPU suite), then
# bench + A/B timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2c_gputest.log | tail -15
timeout 300 python scripts/latency_probe.py > gpurun_out/r2c_latency_1x256.jsonl 2> gpurun_out/r2c_latency.err; echo "latency rc=$?"; cat gpurun_out/r2c_latency_1x256.jsonl
timeout 300 python scripts/latency_probe.py --frames 1000 --passes 3 > gpurun_out/r2c_latency_1x1000.jsonl 2>> gpurun_out/r2c_latency.err; cat gpurun_out/r2c_latency_1x1000.jsonl
timeout 300 python scripts/vocoder_time.py > gpurun_out/r2c_vocoder_v2.json 2> gpurun_out/r2c_vocoder.err; echo "voc v2 rc=$?"; cat gpurun_out/r2c_vocoder_v2.json
USB_SNAKE_V1=1 timeout 300 python scripts/vocoder_time.py > gpurun_out/r2c_vocoder_v1.json 2>> gpurun_out/r2c_vocoder.err; echo "voc v1 rc=$?"; cat gpurun_out/r2c_vocoder_v1.json
timeout 300 python scripts/finetune_time.py 50 > gpurun_out/r2c_finetune_split.json 2> gpurun_out/r2c_finetune.err; echo "ft rc=$?"; cat gpurun_out/r2c_finetune_split.json
USB_FT_NO_SPLITK=1 timeout 300 python scripts/finetune_time.py 50 > gpurun_out/r2c_finetune_nosplit.json 2>> gpurun_out/r2c_finetune.err; cat gpurun_out/r2c_finetune_nosplit.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
P1="python scripts/profile_pass.py --batch 1 --frames 256 --steps 2"
$P1 > gpurun_out/r2c_plain_1x256.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 244 -c 262 --csv --log-file gpurun_out/r2c_launches_1x256.csv $P1 > gpurun_out/r2c_ncu2.log 2>&1; echo "ncu2 rc=$?"
du -sh gpurun_out
