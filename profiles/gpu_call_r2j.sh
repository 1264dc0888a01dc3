#!/bin/bash
# round 2, call J: training after split-K wgrad + the new logits-backward kernel; host-pipeline test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zzz_tensor_core_training.py tests/test_gpu_backward.py tests/test_gpu_fullshape.py tests/test_gpu_parity.py -q -m gpu -x -k "tensor_core or backward or host_pipeline or dp_" > gpurun_out/pytest_train.log 2>&1; echo rc=$? >> gpurun_out/pytest_train.log; tail -4 gpurun_out/pytest_train.log
timeout 300 python profiles/train_breakdown.py 2>&1 | head -2
