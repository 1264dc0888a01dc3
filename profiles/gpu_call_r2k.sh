#!/bin/bash
# round 2, call K: tape backward (P2 flags) vs finite differences / autograd; backward kernels after the colsum / dP / sgemm-tile
# changes; training breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tape_backward.py -q -m gpu > gpurun_out/pytest_tape.log 2>&1; echo rc=$? >> gpurun_out/pytest_tape.log; tail -30 gpurun_out/pytest_tape.log | cut -c1-600
timeout 900 python -m pytest tests/test_zzz_tensor_core_training.py tests/test_gpu_backward.py tests/test_gpu_fullshape.py tests/test_output_unit.py tests/test_stem.py tests/test_encoder.py -q -m gpu -x -k "tensor_core or backward or dp_ or output or stem or encoder" > gpurun_out/pytest_train.log 2>&1; echo rc=$? >> gpurun_out/pytest_train.log; tail -4 gpurun_out/pytest_train.log
timeout 300 python profiles/train_breakdown.py 2>&1 | head -2
