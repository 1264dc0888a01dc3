#!/bin/bash
# second call of the round: knob sweep, LSTM forms, ncu --set full captures of the per-step GEMMs and K3 (hoisted inference form)
mkdir -p gpurun_out
timeout -k 5 150 python profiles/sweep.py 40 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
echo "sweep exit $?" >> gpurun_out/sweep.err
timeout -k 5 60 python profiles/lstm_bench.py > gpurun_out/lstm_bench.jsonl 2> gpurun_out/lstm_bench.err
echo "lstm exit $?" >> gpurun_out/lstm_bench.err
# third forward of profiles/one_forward.py: 26 tc_gemm launches per forward (2 invariant + 12 x 2), 12 kb_attend
timeout -k 5 120 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --launch-skip 54 --launch-count 2 \
    -o gpurun_out/tcgemm_step_r1 -f python profiles/one_forward.py bf16 > gpurun_out/ncu_full_gemm.log 2>&1
echo "ncu gemm exit $?" >> gpurun_out/ncu_full_gemm.log
timeout -k 5 90 ncu --set full --clock-control none --import-source on -k regex:"kb_attend_kernel|scale_rows|skinny_gemm" --launch-skip 74 --launch-count 3 \
    -o gpurun_out/k3_scale_skinny_r1 -f python profiles/one_forward.py bf16 > gpurun_out/ncu_full_k3.log 2>&1
echo "ncu k3 exit $?" >> gpurun_out/ncu_full_k3.log
cat gpurun_out/sweep.jsonl; cat gpurun_out/lstm_bench.jsonl; tail -2 gpurun_out/sweep.err gpurun_out/lstm_bench.err; ls -la gpurun_out
