#!/bin/bash
# round 2, call W (2 GPUs): chunked cast / copy ring -- pipeline test, then the 2-rank bench with 4 and 2 pieces
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k host_pipeline 2>&1 | tail -3
run() {
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 40 --warmup 5 --skip-cpu --skip-train > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err
  python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_u.json").read().strip().splitlines()[-1])
print(round(j["value"]), round(j["e2e"]["value"]), j["e2e"]["h2d_bytes_per_step"], j["e2e"]["host_cast"][:70])
PY
}
echo "== 2 ranks, chunks 4 ring 8 (default)"; run 29531
echo "== chunks 4 ring 4"; MAC_HOST_STAGE_RING=4 run 29532
echo "== chunks 8 ring 8"; MAC_HOST_CAST_CHUNKS=8 MAC_HOST_STAGE_RING=8 run 29533
echo "== 1 rank"; timeout 600 python bench.py --skip-cpu --skip-train 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['e2e']['value']), j['e2e']['numa'].get('h2d_gbs_alone'))"
