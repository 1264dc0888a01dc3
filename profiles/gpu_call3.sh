#!/bin/bash
# final call of the round: full GPU suite on the final defaults, the default bench line, the previous configuration beside it,
# and the fp32 parity-path line
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests -m gpu -q --timeout 100 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -k 5 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -k 5 200 python bench.py > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
echo "bench exit $?" >> gpurun_out/bench_bf16.err
timeout -k 5 120 python bench.py --streams 4 --fold-y 1 --skip-cpu --skip-train > gpurun_out/bench_bf16_s4_fold.json 2> gpurun_out/bench_bf16_s4_fold.err
timeout -k 5 150 python bench.py --prec fp32 --steps 10 --skip-cpu --skip-train > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log
for f in bench_bf16 bench_bf16_s4_fold bench_fp32; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/%s.json" % sys.argv[1]))
    print(sys.argv[1], "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), "train", (d.get("train") or {}).get("ms_per_step"),
          "train_full", (d.get("train_full") or {}).get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
