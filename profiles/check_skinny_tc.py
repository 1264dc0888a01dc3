"""Hardware check of mac_linear_tc_small_fwd (csrc/skinny_tc.cuh): the batch-sized projections as three-pass split-bf16
tcgen05 products, against fp64 torch on the same fp32 inputs; and its time beside the fp32 cluster/DSMEM kernel
(mac_linear_fwd) it replaces in the bf16 configuration.  Run under `timeout`."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mac_network_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    lib = L.load()
    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    bad = 0
    for (M, ks, N, mode) in ((64, [512], 512, "plain"), (64, [512, 512], 512, "plain"), (64, [512, 512, 512], 512, "plain"),
                             (64, [512, 512], 1024, "split_out"), (64, [512], 512, "gate"), (64, [512], 6144, "tanh"),
                             (17, [128, 64], 96, "plain"), (128, [512], 512, "plain"), (64, [512, 512], 512, "single_pass")):
        K = sum(ks)
        xs = [torch.randn(M, k, device="cuda", generator=g).contiguous() for k in ks]
        W = (torch.randn(K, N, device="cuda", generator=g) * K ** -0.5).contiguous()
        b = torch.randn(N, device="cuda", generator=g) * 0.1
        hi = torch.empty(N, K, dtype=torch.bfloat16, device="cuda")
        lo = torch.empty(N, K, dtype=torch.bfloat16, device="cuda")
        L.check(lib.mac_pack_weight_bf16_split(L.ptr(W), L.ptr(hi), L.ptr(lo), K, N, L.stream_ptr()), "pack")
        n = len(xs)
        arr_p = (ctypes.c_void_p * n)(*[x.data_ptr() for x in xs])
        arr_k = (ctypes.c_int * n)(*ks)
        arr_ld = (ctypes.c_int * n)(*ks)
        ref = torch.cat(xs, 1).double() @ W.double() + b.double()
        y = torch.full((M, N), float("nan"), device="cuda")
        y2 = gn = go = gz = None
        ldy, nsplit, act, bc = N, 0, 0, 0.0
        if mode == "split_out":
            y = torch.full((M, N // 2), float("nan"), device="cuda")
            y2 = torch.full((M, N // 2), float("nan"), device="cuda")
            ldy, nsplit = N // 2, N // 2
        if mode == "gate":
            gn, go = torch.randn(M, N, device="cuda", generator=g), torch.randn(M, N, device="cuda", generator=g)
            gz = torch.empty(M, N, device="cuda")
            bc = 1.0
            z = torch.sigmoid(ref + 1.0)
            ref_gate_z = z
            ref = gn.double() * z + go.double() * (1 - z)
        if mode == "tanh":
            act = 1
            ref = torch.tanh(ref)
        lo_ptr = None if mode == "single_pass" else L.ptr(lo)

        def run():
            L.check(lib.mac_linear_tc_small_fwd(arr_p, arr_k, arr_ld, n, L.ptr(hi), lo_ptr, L.ptr(b), bc, act, L.ptr(y), ldy,
                                                L.ptr(y2), nsplit, L.ptr(gn), L.ptr(go), L.ptr(gz), M, N, L.stream_ptr()),
                    "mac_linear_tc_small_fwd")
        run()
        torch.cuda.synchronize()
        got = torch.cat([y, y2], 1) if y2 is not None else y
        err = float((got.double() - ref).abs().max() / ref.abs().max())
        rec = {"M": M, "K": ks, "N": N, "mode": mode, "max_rel": err}
        if mode == "gate":
            rec["gate_z_err"] = float((gz.double() - ref_gate_z).abs().max())
        tol = 2e-2 if mode == "single_pass" else 3e-5
        if not (err < tol):
            bad += 1
        rec["us_tc"] = timeit(run)
        if mode in ("plain", "tanh") and M <= 64:
            ws_b = lib.mac_linear_workspace_bytes(M, K, N)
            ws = torch.zeros(ws_b, dtype=torch.uint8, device="cuda")
            y32 = torch.empty(M, N, device="cuda")

            def run32():
                L.check(lib.mac_linear_fwd(arr_p, arr_k, arr_ld, n, L.ptr(W), L.ptr(b), 0.0, act, L.ptr(y32), N, M, N, L.ptr(ws),
                                           ws_b, L.stream_ptr()))
            rec["us_fp32_cluster_kernel"] = timeit(run32)
        print(json.dumps(rec), flush=True)
    print("SKINNY_TC_CHECK", "FAIL" if bad else "OK")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
