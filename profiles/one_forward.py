"""One headline-shaped forward (B=64, S=40, N=196, d=512, netLength=12) without CUDA graphs -- the command the ncu
launch lists in profiles/ are taken from:
    ncu --metrics gpu__time_duration.sum --clock-control none -s <warm-up launches> -c <N> --csv --log-file ... python profiles/one_forward.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mac_network_b200.config import MACConfig
from mac_network_b200.mac_cell import MACCell, MACParams, mac_network
from mac_network_b200.params import init_params, perturb_biases
from mac_network_b200.synthetic import SHAPES, make_inputs

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
variant = sys.argv[2] if len(sys.argv) > 2 else "args"
shape = SHAPES["gqa" if variant == "gqa" else "headline"]
B, S, N, d, L = shape
cfg = MACConfig.args(variant, netLength=L)
params = MACParams(cfg, L, values=perturb_biases(init_params(cfg, L, seed=100), seed=101))
x = {k: torch.from_numpy(v).cuda() for k, v in make_inputs(B, S, N, d, seed=1234).items()}
cell = MACCell(x["vecQuestions"], x["questionWords"], x["questionCntxWords"], x["questionLengths"], x["knowledgeBase"],
               1.0, 1.0, 1.0, B, False, config=cfg, params=params, prec=prec)
for _ in range(3):
    mac_network(cell, L)
torch.cuda.synchronize()
