#!/bin/bash
# round 2, call E: whole -m gpu suite (after the tensor-map cache change), training arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_r2.log; tail -4 gpurun_out/pytest_gpu_r2.log
timeout 200 python bench.py --mode train --train-prec bf16 --bwd-tc 1 --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-400
timeout 200 python bench.py --mode train --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-300
