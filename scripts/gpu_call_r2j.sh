#!/bin/bash
# round-2 GPU call J (2 GPUs): the sharded bench path end to end (sample_sharded + NCCL gather, host-tensor e2e) and the
# reference arm under torchrun
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 800 $TR bench.py --gpus 2 --steps 3 --warmup 3 --headline-only > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r2j_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_2gpu.json'))
print('n_gpus',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),d['e2e'],'launches',d['gpu_launches'],d['config']['job_utterances'])
PY
timeout 400 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --no-config1 > gpurun_out/r2j_ref_2gpu.json 2> gpurun_out/r2j_ref_2gpu.err; echo "ref2 rc=$?"; cut -c1-300 gpurun_out/r2j_ref_2gpu.json
