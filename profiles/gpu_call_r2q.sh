#!/bin/bash
# round 2, call Q: compute-sanitizer memcheck over the kernels added in the second half of round 2
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tape_backward.py tests/test_zzz_general_path_training.py tests/test_output_unit.py -q -m gpu -x -k "p2_read_bl or p2_control or p2_write_sum or memory_bn or out_of_range or batch_norm or args1" > gpurun_out/sanitizer_tape.log 2>&1; echo rc=$? >> gpurun_out/sanitizer_tape.log; tail -6 gpurun_out/sanitizer_tape.log | cut -c1-300
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zzz_tensor_core_training.py -q -m gpu -x -k "gradients and gqa" > gpurun_out/sanitizer_tc.log 2>&1; echo rc=$? >> gpurun_out/sanitizer_tc.log; tail -6 gpurun_out/sanitizer_tc.log | cut -c1-300
