#!/bin/bash
# round 2, call M: tensor-core backward with the fused transposing casts; memoryBN finite differences; training breakdown + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zzz_tensor_core_training.py tests/test_gpu_backward.py tests/test_gpu_fullshape.py tests/test_gpu_tape_backward.py -q -m gpu -k "tensor_core or backward or dp_ or memory_bn" > gpurun_out/pytest_train.log 2>&1; echo rc=$? >> gpurun_out/pytest_train.log; tail -12 gpurun_out/pytest_train.log | cut -c1-300
timeout 300 python profiles/train_breakdown.py 2>&1 | head -2
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/launches_train_tc3.csv python profiles/one_train_step.py > gpurun_out/ncu_train3.log 2>&1; tail -2 gpurun_out/ncu_train3.log | cut -c1-200
python profiles/launch_summary_grid.py gpurun_out/launches_train_tc3.csv 2>&1 | head -30
