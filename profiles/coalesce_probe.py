"""Probe (not a bench value): how fast is the same inference when k independent B=64 batches are stacked into one
B = 64k pass (one launch sequence) instead of running on k streams?  Samples are independent in eval mode, so the
per-sample results are the same function; this only measures what launch/tile quantisation costs at B=64.
Usage: python profiles/coalesce_probe.py [k ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_network_b200.config import MACConfig
from mac_network_b200.mac_cell import MACCell, MACParams, mac_network
from mac_network_b200.synthetic import SHAPES, make_inputs

B0, S, N, d, L = SHAPES["headline"]
cfg = MACConfig.args("args", netLength=L)
params = MACParams(cfg, L, seed=1)
for k in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    B = B0 * k
    nsets = max(2, 6 // k)
    graphs = []
    for s in range(nsets):
        x = {kk: torch.from_numpy(v).cuda() for kk, v in make_inputs(B, S, N, d, seed=10 + s).items()}
        cell = MACCell(x["vecQuestions"], x["questionWords"], x["questionCntxWords"], x["questionLengths"],
                       x["knowledgeBase"], 1.0, 1.0, 1.0, B, False, config=cfg, params=params, prec="bf16")
        mac_network(cell, L)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = mac_network(cell, L)
        graphs.append((g, cell, out))
    for g, _, _ in graphs:
        g.replay()
    torch.cuda.synchronize()
    n = 24
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        graphs[i % nsets][0].replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / n
    print("stacked batches per pass: %d (B=%d)  %.3f ms/pass  -> %.0f reasoning-steps/s in units of B=64 batches"
          % (k, B, t * 1e3, L * k / t))
    del graphs
    torch.cuda.empty_cache()
