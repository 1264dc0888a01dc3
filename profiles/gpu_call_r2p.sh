#!/bin/bash
# round 2, call P: tile choice of the persistent tcgen05 GEMM at K = 1536 (tc32) and in the training forward; streams sweep point
mkdir -p gpurun_out
for t in 0 128128 128256; do
  echo "== MAC_TC_TILE=$t tc32"
  if [ $t = 0 ]; then unset MAC_TC_TILE; else export MAC_TC_TILE=$t; fi
  timeout 300 python bench.py --mode quick --prec tc32 --streams 4 --steps 20 --warmup 3 2>/dev/null | tail -1 | cut -c1-260
done
for t in 0 128128; do
  echo "== MAC_TC_TILE=$t train breakdown"
  if [ $t = 0 ]; then unset MAC_TC_TILE; else export MAC_TC_TILE=$t; fi
  timeout 300 python profiles/train_breakdown.py 2>&1 | sed -n 2p | cut -c1-300
done
unset MAC_TC_TILE
echo "== streams 8 bf16"
timeout 300 python bench.py --mode quick --prec bf16 --streams 8 --steps 40 --warmup 5 2>/dev/null | tail -1 | cut -c1-260
