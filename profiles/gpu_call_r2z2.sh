#!/bin/bash
# round 2, call Z2 (4 GPUs): host cast forced on (staging ring of 2) with three ranks on one socket
mkdir -p gpurun_out
MAC_FORCE_HOST_CAST=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 40 --warmup 5 --skip-cpu --skip-train > gpurun_out/bench_z2.json 2> gpurun_out/bench_z2.err; echo rc=$?
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_z2.json").read().strip().splitlines()[-1])
print(round(j["value"]), round(j["e2e"]["value"]), j["e2e"]["h2d_bytes_per_step"], json.dumps(j["e2e"]["numa"]), j["e2e"]["host_cast"][:80])
PY
