#!/bin/bash
# round 2, call N: the whole -m gpu suite, smoke(), and the default bench line
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_r2.log; tail -5 gpurun_out/pytest_gpu_r2.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo bench rc=$?; cut -c1-1500 gpurun_out/bench_r2.json
