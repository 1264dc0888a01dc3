#!/bin/bash
# round 2, call A: fused kernel check + phases, full-shape parity tests, quick throughput sweep
mkdir -p gpurun_out
timeout 200 python profiles/check_read_fused.py > gpurun_out/fused_check3.log 2>&1; echo rc=$? >> gpurun_out/fused_check3.log; tail -4 gpurun_out/fused_check3.log
timeout 120 python profiles/fused_phases.py > gpurun_out/fused_phases3.log 2>&1; head -2 gpurun_out/fused_phases3.log
timeout 600 python -m pytest tests/test_gpu_fullshape.py -x -q -m gpu -s > gpurun_out/pytest_fullshape.log 2>&1; echo rc=$? >> gpurun_out/pytest_fullshape.log; tail -15 gpurun_out/pytest_fullshape.log
for cfg in "1 1" "2 1" "4 0" "6 0" "6 1" "8 0"; do set -- $cfg; timeout 120 python bench.py --mode quick --streams $1 --fold-y $2 --steps 24 --warmup 6 --min-time 0.3 2>&1 | tail -1; done > gpurun_out/quick_sweep.jsonl 2>&1
cat gpurun_out/quick_sweep.jsonl | cut -c1-400
MAC_READ_FUSED=0 timeout 120 python bench.py --mode quick --streams 6 --fold-y 0 --steps 24 --warmup 6 --min-time 0.3 2>&1 | tail -1 | cut -c1-300
