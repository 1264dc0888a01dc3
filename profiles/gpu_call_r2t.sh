#!/bin/bash
# round 2, call T: final evidence -- whole -m gpu suite, smoke(), default bench line, launch list + ncu --set full of the packed read step
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_r2.log; tail -3 gpurun_out/pytest_gpu_r2.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo bench rc=$?
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/launches_pass_r2.csv python profiles/one_pass.py > gpurun_out/ncu_pass.log 2>&1; tail -1 gpurun_out/ncu_pass.log | cut -c1-200
python profiles/launch_summary_grid.py gpurun_out/launches_pass_r2.csv | head -16
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:read_step2 -c 1 -f -o gpurun_out/read_step2_packed python profiles/one_pass.py > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/read_step2_packed.ncu-rep 2>&1 | tail -2 | cut -c1-900
