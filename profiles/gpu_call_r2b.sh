#!/bin/bash
# round 2, call B: whole-step kernel -- parity tests at the headline shape, throughput sweep
mkdir -p gpurun_out
timeout 200 python profiles/check_read_fused.py > gpurun_out/fused_check5.log 2>&1; echo rc=$? >> gpurun_out/fused_check5.log; tail -2 gpurun_out/fused_check5.log
timeout 900 python -m pytest tests/test_gpu_fullshape.py -x -q -m gpu -s > gpurun_out/pytest_fullshape2.log 2>&1; echo rc=$? >> gpurun_out/pytest_fullshape2.log; grep -E "worst|errors|passed|failed|launches|whole-step|Error|assert" gpurun_out/pytest_fullshape2.log | cut -c1-300
for cfg in "1 0" "2 0" "4 0" "6 0" "8 0"; do set -- $cfg; timeout 120 python bench.py --mode quick --streams $1 --fold-y $2 --steps 24 --warmup 6 --min-time 0.3 2>&1 | tail -1; done > gpurun_out/quick_sweep2.jsonl 2>&1
cut -c1-330 gpurun_out/quick_sweep2.jsonl
MAC_STEP_FUSED=0 timeout 120 python bench.py --mode quick --streams 6 --fold-y 0 --steps 24 --warmup 6 --min-time 0.3 2>&1 | tail -1 | cut -c1-330
