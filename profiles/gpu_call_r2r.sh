#!/bin/bash
# round 2, call R: packed tiles in the fused read-step kernel (98 CTAs instead of 128 at the headline shape)
mkdir -p gpurun_out
timeout 300 python profiles/check_read_fused.py > gpurun_out/check_read_fused.log 2>&1; echo rc=$? >> gpurun_out/check_read_fused.log; tail -16 gpurun_out/check_read_fused.log | cut -c1-330
timeout 200 python profiles/fused_phases.py 2>&1 | head -2 | cut -c1-700
timeout 900 python -m pytest tests/test_gpu_fullshape.py tests/test_gpu_parity.py -q -m gpu -x -k "not backward" > gpurun_out/pytest_packed.log 2>&1; echo rc=$? >> gpurun_out/pytest_packed.log; tail -4 gpurun_out/pytest_packed.log | cut -c1-300
for st in 1 4 8; do timeout 300 python bench.py --mode quick --prec bf16 --streams $st --steps 40 --warmup 5 2>/dev/null | tail -1 | cut -c1-200; done
MAC_READ_PACKED=0 timeout 300 python bench.py --mode quick --prec bf16 --streams 8 --steps 40 --warmup 5 2>/dev/null | tail -1 | cut -c1-200
