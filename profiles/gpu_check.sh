#!/bin/bash
# One gpurun call: GPU parity tests (the newest rows first), smoke(), the default bench line, a launch list of one forward.
# Everything lands in gpurun_out/ (merged back by gpurun); summaries worth keeping are copied into profiles/ by hand.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 5 ${PYTEST_LIMIT:-320} python -m pytest tests/test_encoder.py tests/test_stem.py tests/test_full_model.py tests -m gpu -q \
    --timeout 100 --durations=12 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -k 5 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -k 5 ${BENCH_LIMIT:-240} python bench.py > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
echo "bench exit $?" >> gpurun_out/bench_bf16.err
timeout -k 5 ${NCU_LIMIT:-150} ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_bf16.csv python profiles/one_forward.py bf16 > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu.log
tail -6 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; head -c 400 gpurun_out/bench_bf16.json; echo; tail -2 gpurun_out/bench_bf16.err
