#!/bin/bash
# round 2, call U (2 GPUs): host bf16 cast through a small LLC-resident staging ring vs no cast, two ranks on one socket
mkdir -p gpurun_out
run() {
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 40 --warmup 5 --skip-cpu --skip-train > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err
  python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_u.json").read().strip().splitlines()[-1])
print(round(j["value"]), round(j["e2e"]["value"]), j["e2e"]["h2d_bytes_per_step"], j["e2e"]["host_cast"][:90])
PY
}
echo "== default (no cast with 2 ranks on a socket)"; run 29521
echo "== forced cast, ring 3"; MAC_FORCE_HOST_CAST=1 run 29522
echo "== forced cast, ring 2"; MAC_FORCE_HOST_CAST=1 MAC_HOST_STAGE_RING=2 run 29523
echo "== forced cast, ring 12"; MAC_FORCE_HOST_CAST=1 MAC_HOST_STAGE_RING=12 run 29524
