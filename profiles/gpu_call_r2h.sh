#!/bin/bash
# round 2, call H (2 GPUs): smoke(), the 2-rank NCCL gradient test, and the bench line at N=2 (NUMA binding path)
mkdir -p gpurun_out
nvidia-smi -L
true
true
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo rc=$?; python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/bench_r2_n2.json').read().strip().splitlines()[-1])
    print({k:b[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', {k:b['e2e'][k] for k in ('value','ms_per_step','numa','host_cast')}, 'train', b['train'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_r2_n2.err').read()[-1500:])
PY
