#!/bin/bash
# round 2, call F: backward tests after the cancellation-free softmax backward; launch list of one tensor-core training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullshape.py tests/test_gpu_backward.py -q -m gpu -s -k "backward" > gpurun_out/pytest_bwd.log 2>&1; echo rc=$? >> gpurun_out/pytest_bwd.log; grep -E "five worst|passed|failed|rc=" gpurun_out/pytest_bwd.log | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 8100 -c 2700 --csv --log-file gpurun_out/launches_train_tc.csv python bench.py --mode train --train-prec bf16 --bwd-tc 1 --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1; tail -2 gpurun_out/ncu_train.log | cut -c1-300
python profiles/launch_summary.py gpurun_out/launches_train_tc.csv 2>&1 | head -30
