#!/bin/bash
# round 2, call L: memoryBN + tape tests; launch list (with grid sizes) of one tensor-core training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tape_backward.py tests/test_zzz_general_path_training.py tests/test_gpu_parity.py tests/test_output_unit.py -q -m gpu -k "tape or general_path or batch_norm or memory_bn or out_of_range" > gpurun_out/pytest_bn.log 2>&1; echo rc=$? >> gpurun_out/pytest_bn.log; tail -25 gpurun_out/pytest_bn.log | cut -c1-400
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/launches_train_tc2.csv python profiles/one_train_step.py > gpurun_out/ncu_train2.log 2>&1; tail -2 gpurun_out/ncu_train2.log | cut -c1-200
python profiles/launch_summary_grid.py gpurun_out/launches_train_tc2.csv 2>&1 | head -52
