#!/bin/bash
# round 2, call S: packed read step with the in-kernel last-tile merge (no combine launch)
mkdir -p gpurun_out
timeout 300 python profiles/check_read_fused.py > gpurun_out/check_read_fused.log 2>&1; echo rc=$? >> gpurun_out/check_read_fused.log; tail -4 gpurun_out/check_read_fused.log | cut -c1-330
timeout 200 python profiles/fused_phases.py 2>&1 | head -2 | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_fullshape.py tests/test_gpu_parity.py -q -m gpu -x -k "not backward" > gpurun_out/pytest_packed.log 2>&1; echo rc=$? >> gpurun_out/pytest_packed.log; tail -3 gpurun_out/pytest_packed.log | cut -c1-300
for st in 1 12; do timeout 300 python bench.py --mode quick --prec bf16 --streams $st --steps 48 --warmup 5 2>/dev/null | tail -1 | cut -c1-120; done
timeout 600 python bench.py --rooflines-only 2>/dev/null | tail -1 | python -c "
import sys, json
j=json.loads(sys.stdin.read()); r=j.get('roofline') or j.get('rooflines_all',{}).get('read_step_fused'); print(json.dumps(r)[:600])"
