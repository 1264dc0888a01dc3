#!/bin/bash
# round 2, call Y: 128-column tiles in the batch-sized tensor-core projection kernel (write + projY, N = 1024)
mkdir -p gpurun_out
timeout 300 python profiles/check_skinny_tc.py 2>&1 | tail -12 | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_fullshape.py tests/test_gpu_parity.py -q -m gpu -x -k "small_tc or host_pipeline or throughput" 2>&1 | tail -2
for bn in 128 64; do for st in 1 12; do echo "BN=$bn streams=$st"; MAC_SKINNY_BN=$bn timeout 300 python bench.py --mode quick --prec bf16 --streams $st --steps 48 --warmup 5 2>/dev/null | tail -1 | cut -c1-110; done; done
