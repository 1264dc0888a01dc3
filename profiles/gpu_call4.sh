#!/bin/bash
# last call of the round: the default bench line after the final bench.py edits, and the launch list of one forward in the
# bench's throughput configuration (write unit not folded with the next projY)
mkdir -p gpurun_out
timeout -k 5 150 python bench.py > gpurun_out/bench_bf16_final.json 2> gpurun_out/bench_bf16_final.err
echo "bench exit $?" >> gpurun_out/bench_bf16_final.err
MAC_NO_FOLD_Y=1 timeout -k 5 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_bf16_nofold.csv python profiles/one_forward.py bf16 > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu.log
head -c 300 gpurun_out/bench_bf16_final.json; echo; tail -2 gpurun_out/bench_bf16_final.err; tail -1 gpurun_out/ncu.log
