#!/bin/bash
# round 2, call I: launch list of ONE tensor-core training step of the cell (after 2 warm-up steps); full -m gpu suite
mkdir -p gpurun_out
cat > /tmp/one_train_step.py <<'PY'
import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
from mac_network_b200.config import MACConfig
from mac_network_b200.dp import DPTrainer
from mac_network_b200.params import init_params, perturb_biases
from mac_network_b200.synthetic import SHAPES, make_inputs
B, S, N, d, L = SHAPES["headline"]
cfg = MACConfig.args("args", netLength=L)
pv = perturb_biases(init_params(cfg, L, seed=100), seed=101)
tr = DPTrainer(cfg, L, param_values=pv, seed=7, prec="bf16", bwd_tc=True, classifier=(28, [512]))
batch = {k: torch.from_numpy(v).cuda() for k, v in make_inputs(B, S, N, d, seed=1).items()}
ans = torch.randint(0, 28, (B,), dtype=torch.int32, device="cuda")
for i in range(3):
    if i == 2:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    tr.train_step_answers(0, batch, ans, B)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_tc.csv python /tmp/one_train_step.py > gpurun_out/ncu_train.log 2>&1; tail -2 gpurun_out/ncu_train.log | cut -c1-200
python profiles/launch_summary.py gpurun_out/launches_train_tc.csv 2>&1 | head -32
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_r2.log; tail -4 gpurun_out/pytest_gpu_r2.log
