#!/bin/bash
# round 2, call Z (4 GPUs): the 4-rank bench line (rank -> GPU map, NUMA policy, NCCL training arm)
mkdir -p gpurun_out
python - <<'PY'
import torch
from mac_network_b200.serving import gpu_numa_nodes, device_for_rank
print("visible", torch.cuda.device_count(), "numa", gpu_numa_nodes(), "map4", [device_for_rank(r, 4) for r in range(4)], "map2", [device_for_rank(r, 2) for r in range(2)])
PY
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 40 --warmup 5 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err; echo rc=$?
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r2_n4.json").read().strip().splitlines()[-1])
print(round(j["value"]), round(j["e2e"]["value"]), j["e2e"]["h2d_bytes_per_step"], json.dumps(j["e2e"]["numa"])); print(json.dumps(j.get("train"))[:300])
PY
tail -2 gpurun_out/bench_r2_n4.err | cut -c1-300
