#!/bin/bash
# round 2, call G: tc32 (split-bf16) parity + throughput; training step breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullshape.py -q -m gpu -s -k "tc32 or backward" > gpurun_out/pytest_tc32.log 2>&1; echo rc=$? >> gpurun_out/pytest_tc32.log; grep -E "worst|passed|failed|rc=|Error|assert " gpurun_out/pytest_tc32.log | cut -c1-400
timeout 200 python bench.py --mode quick --prec tc32 --streams 4 --steps 8 --warmup 3 --min-time 0.3 2>&1 | tail -1 | cut -c1-330
timeout 200 python bench.py --mode quick --prec fp32 --streams 4 --steps 8 --warmup 3 --min-time 0.3 2>&1 | tail -1 | cut -c1-330
timeout 300 python profiles/train_breakdown.py 2>&1 | tail -5
