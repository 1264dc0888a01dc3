"""Profiling experiment: time the tcgen05 GEMM with pieces switched off (MAC_TC_DEBUG) and both tile widths."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mac_network_b200 import _lib as L
lib = L.load()
M, K, N = 12544, int(os.environ.get("K", "1024")), 512
xs = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(6)]
W = torch.randn(K, N, device="cuda") / K ** 0.5
Wt = torch.empty(N, K, dtype=torch.bfloat16, device="cuda")
L.check(lib.mac_pack_weight_bf16(L.ptr(W), L.ptr(Wt), K, N, L.stream_ptr()))
bias = torch.zeros(N, device="cuda")
y = torch.empty(M, N, device="cuda")
for bn in ("256256", "pair"):
    os.environ["MAC_TC_PAIR"] = "1" if bn == "pair" else "0"
    for dbg in ("0", "1"):
        os.environ["MAC_TC_TILE"] = "256256"
        os.environ["MAC_TC_DEBUG"] = dbg
        for x in xs:
            L.check(lib.mac_linear_tc_fwd(L.ptr(x), L.ptr(Wt), L.ptr(bias), 3, L.ptr(y), 0, M, K, N, L.stream_ptr()))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(30):
            L.check(lib.mac_linear_tc_fwd(L.ptr(xs[i % 6]), L.ptr(Wt), L.ptr(bias), 3, L.ptr(y), 0, M, K, N, L.stream_ptr()))
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) / 30 * 1e3
        print("tile=%s debug=%s (1=no epi 2=no MMA 4=no TMA 8=no LDTM 16=no bias 32=no store): %.1f us  %.0f TFLOP/s-equivalent" % (bn, dbg, t, 2.0 * M * K * N / t / 1e6))
