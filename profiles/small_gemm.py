"""Micro-benchmark of the M=64 projections (skinny cluster kernel vs the generic split-K kernel)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mac_network_b200 import _lib as L
lib = L.load()
t = torch.zeros(1024, device="cuda")
g = torch.cuda.CUDAGraph()
t.add_(1); torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(50): t.add_(1)
g.replay(); torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); g.replay(); e.record(); torch.cuda.synchronize()
print("trivial elementwise kernel: %.2f us per dependent launch in a graph" % (a.elapsed_time(e) / 50 * 1e3))
for (M, K, N) in [(64, 512, 512), (64, 1024, 512), (64, 1536, 512), (64, 512, 6144)]:
    x = torch.randn(M, K, device="cuda"); W = torch.randn(K, N, device="cuda"); b = torch.zeros(N, device="cuda")
    y = torch.empty(M, N, device="cuda")
    wsb = int(lib.mac_linear_workspace_bytes(M, K, N)); ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    arr_p = (ctypes.c_void_p * 1)(x.data_ptr()); arr_k = (ctypes.c_int * 1)(K)
    for mode in ("0", "3"):
        os.environ["MAC_SK_DEBUG"] = mode
        def f():
            L.check(lib.mac_linear_fwd(arr_p, arr_k, arr_k, 1, L.ptr(W), L.ptr(b), 0.0, 0, L.ptr(y), N, M, N, L.ptr(ws), wsb, L.stream_ptr()))
        for _ in range(5): f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(50): f()
        g.replay(); torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); e.record(); torch.cuda.synchronize()
        print("M=%d K=%d N=%d dbg=%-3s (1 no load, 2 no fma, 4 no dsmem, 8 no cluster.sync) %.2f us" % (M, K, N, mode, a.elapsed_time(e) / 50 * 1e3))
