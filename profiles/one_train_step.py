"""ONE tensor-core training step of the cell + output unit at the headline shape between cudaProfilerStart/Stop (after two
warm-up steps): run under `ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size` for a launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_network_b200.config import MACConfig                      # noqa: E402
from mac_network_b200.dp import DPTrainer                          # noqa: E402
from mac_network_b200.params import init_params, perturb_biases    # noqa: E402
from mac_network_b200.synthetic import SHAPES, make_inputs         # noqa: E402

B, S, N, d, L = SHAPES["headline"]
cfg = MACConfig.args("args", netLength=L)
pv = perturb_biases(init_params(cfg, L, seed=100), seed=101)
tr = DPTrainer(cfg, L, param_values=pv, seed=7, prec="bf16", bwd_tc=True, classifier=(28, [512]))
batch = {k: torch.from_numpy(v).cuda() for k, v in make_inputs(B, S, N, d, seed=1).items()}
ans = torch.randint(0, 28, (B,), dtype=torch.int32, device="cuda")
for i in range(3):
    if i == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    tr.train_step_answers(0, batch, ans, B)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
