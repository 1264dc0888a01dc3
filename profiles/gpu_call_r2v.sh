#!/bin/bash
# round 2, call V (1 GPU): staging-ring size of the host bf16 cast, single rank
mkdir -p gpurun_out
for r in 2 3 4 12; do
  MAC_HOST_STAGE_RING=$r timeout 600 python bench.py --skip-cpu --skip-train > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/bench_v.json").read().strip().splitlines()[-1])
print("ring $r", round(j["value"]), round(j["e2e"]["value"]), j["e2e"]["numa"].get("h2d_gbs_alone"), j["timed_blocks"]["e2e"]["block_ms_min_median_max"])
PY
done
