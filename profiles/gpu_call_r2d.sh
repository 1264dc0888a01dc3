#!/bin/bash
# round 2, call D: the whole -m gpu suite, the default bench line, ncu launch list + full capture of the fused kernel
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_r2.log; tail -6 gpurun_out/pytest_gpu_r2.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo bench rc=$?; cut -c1-1500 gpurun_out/bench_r2.json; tail -3 gpurun_out/bench_r2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv --log-file gpurun_out/launches_r2.csv python bench.py --mode quick --streams 1 --steps 3 --warmup 3 --min-time 0 > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:read_step2 -s 4 -c 2 -o gpurun_out/read_step2_r2 python profiles/fused_phases.py > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
