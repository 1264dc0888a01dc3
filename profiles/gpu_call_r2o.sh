#!/bin/bash
# round 2, call O (2 GPUs): rank -> GPU spread over NUMA nodes, NCCL gradient equivalence, 2-rank bench line
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
python - <<'PY'
import torch
from mac_network_b200.serving import gpu_numa_nodes, device_for_rank
print("visible", torch.cuda.device_count(), "numa", gpu_numa_nodes(), "map", [device_for_rank(r, 2) for r in range(2)])
PY
timeout 600 python -m pytest tests/test_gpu_nccl_dp.py -q -m gpu 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo rc=$?
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r2_n2.json").read().strip().splitlines()[-1])
print(j["value"], json.dumps(j["e2e"])[:700]); print(json.dumps(j.get("train"))[:500])
PY
tail -3 gpurun_out/bench_r2_n2.err
