#!/bin/bash
# round 2, call X (1 GPU): pipeline tests + default bench (one-piece staging ring of 2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k host_pipeline 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo bench rc=$?
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r2.json").read().strip().splitlines()[-1])
print(round(j["value"]), round(j["e2e"]["value"]), j["e2e"]["numa"].get("h2d_gbs_alone"), j["timed_blocks"]["e2e"]["block_ms_min_median_max"], j["train"]["ms_per_step"], j["roofline"]["frac"])
PY
