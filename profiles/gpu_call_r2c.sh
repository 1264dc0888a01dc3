#!/bin/bash
# round 2, call C: batch-sized projections on tensor cores (skinny_tc) in the bf16 cell -- check, parity, throughput
mkdir -p gpurun_out
timeout 120 python profiles/check_skinny_tc.py > gpurun_out/skinny_tc_check2.log 2>&1; echo rc=$? >> gpurun_out/skinny_tc_check2.log; cut -c1-220 gpurun_out/skinny_tc_check2.log
timeout 900 python -m pytest tests/test_gpu_fullshape.py tests/test_gpu_parity.py -x -q -m gpu -s -k "bf16 or whole_step or fused or host_pipeline" > gpurun_out/pytest_bf16.log 2>&1; echo rc=$? >> gpurun_out/pytest_bf16.log; grep -E "worst|errors|passed|failed|launches|whole-step|Error|assert|rc=" gpurun_out/pytest_bf16.log | cut -c1-300
for cfg in "1 1" "2 1" "4 1" "6 1" "6 0" "8 1"; do set -- $cfg; timeout 120 python bench.py --mode quick --streams $1 --fold-y $2 --steps 24 --warmup 6 --min-time 0.3 2>&1 | tail -1; done > gpurun_out/quick_sweep3.jsonl 2>&1
cut -c1-330 gpurun_out/quick_sweep3.jsonl
MAC_SMALL_TC=0 timeout 120 python bench.py --mode quick --streams 6 --fold-y 0 --steps 24 --warmup 6 --min-time 0.3 2>&1 | tail -1 | cut -c1-330
timeout 120 python bench.py --mode quick --workload gqa --streams 6 --steps 24 --warmup 6 --min-time 0.3 2>&1 | tail -1 | cut -c1-330
