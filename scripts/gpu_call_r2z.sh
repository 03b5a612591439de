#!/bin/bash
# round-2 GPU call Z (8 GPUs): the default bench through torch.distributed.run = BASELINE.json configs[2] (256 utt x 1000 frames)
mkdir -p gpurun_out
nvidia-smi -L | head -8
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 3 --warmup 3 --headline-only > gpurun_out/r2z_bench_8gpu.json 2> gpurun_out/r2z_bench_8gpu.err; rc=$?
t1=$(date +%s); echo "bench8 rc=$rc wall=$((t1-t0)) s"
tail -3 gpurun_out/r2z_bench_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench_8gpu.json'))
print('n_gpus',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms_per_step',round(d['ms_per_step'],1),'clk',d['clocks'],'launches',d['gpu_launches'], d['config']['job_utterances'])
PY
