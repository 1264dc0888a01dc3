"""Hardware check of the fused read-step kernel (csrc/read_step.cuh) against the four-launch chain it replaces
(scale_rows + tc_gemm<ADDACT> + tc_gemm<LOGITS> + kb_attend, validated against the oracle by tests/test_gpu_parity.py),
through the C ABI: same inv = [P | Q], same y / control -> att and info must agree to fp32 summation order.
Also times both forms back to back on rotating inputs.  Run under `timeout` (a barrier bug would hang)."""
import ctypes
import json
import os
import sys

os.environ["MAC_READ_FUSED"] = "0"            # mac_read_fwd_inv below = the unfused reference form
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mac_network_b200 import _lib as L  # noqa: E402


def make_weights(lib, d, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    def rn(*s, scale=1.0):
        return (torch.randn(*s, device="cuda", generator=g) * scale).contiguous()
    t = {"Wx": rn(d, d, scale=d ** -0.5), "bx": rn(d, scale=0.1), "Wy": rn(d, d, scale=d ** -0.5), "by": rn(d, scale=0.1),
         "Wm": rn(2 * d, d, scale=(2 * d) ** -0.5), "bm": rn(d, scale=0.1), "Wm2": rn(d, d, scale=d ** -0.5),
         "bm2": rn(d, scale=0.1), "wr": rn(d, scale=d ** -0.5 * 4)}
    def pack(w):
        o = torch.empty((w.shape[1], w.shape[0]), dtype=torch.bfloat16, device="cuda")
        L.check(lib.mac_pack_weight_bf16(L.ptr(w), L.ptr(o), w.shape[0], w.shape[1], L.stream_ptr()))
        return o
    t["Wx16"], t["Wm16"], t["Wm216"] = pack(t["Wx"]), pack(t["Wm"]), pack(t["Wm2"])
    rw = L.ReadWeights(t["Wx"].data_ptr(), t["bx"].data_ptr(), t["Wy"].data_ptr(), t["by"].data_ptr(), t["Wm"].data_ptr(),
                       t["bm"].data_ptr(), t["Wm2"].data_ptr(), t["bm2"].data_ptr(), t["wr"].data_ptr(), 0.25,
                       t["Wx16"].data_ptr(), t["Wm16"].data_ptr(), t["Wm216"].data_ptr())
    return t, rw


def case(lib, B, N, d, seed, time_it=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t, rw = make_weights(lib, d, seed + 1)
    nsets = 8 if time_it else 1
    sets = []
    for _ in range(nsets):
        kb = torch.nn.functional.elu(torch.randn(B, N, d, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
        y = torch.randn(B, d, device="cuda", generator=g).contiguous()
        c = torch.randn(B, d, device="cuda", generator=g).contiguous()
        mem = torch.randn(B, d, device="cuda", generator=g).contiguous()
        nb = lib.mac_read_invariant_bytes(B, N, d, 1)
        inv = torch.empty(nb, dtype=torch.uint8, device="cuda")
        L.check(lib.mac_read_invariant(None, L.ptr(kb), ctypes.byref(rw), 1, L.ptr(inv), nb, B, N, d, L.stream_ptr()), "inv")
        sets.append((kb, y, c, mem, inv))
    wsb = lib.mac_read_workspace_bytes(B, N, d, 1)
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    info0, att0 = torch.empty(B, d, device="cuda"), torch.empty(B, N, device="cuda")
    info1, att1 = torch.full((B, d), float("nan"), device="cuda"), torch.full((B, N), float("nan"), device="cuda")

    def unfused(s):
        kb, y, c, mem, inv = s
        L.check(lib.mac_read_fwd_inv(None, L.ptr(kb), L.ptr(inv), L.ptr(y), L.ptr(mem), L.ptr(c), ctypes.byref(rw), 1,
                                     L.ptr(info0), L.ptr(att0), L.ptr(ws), wsb, B, N, d, L.stream_ptr()), "read_fwd_inv")

    def fused(s):
        kb, y, c, mem, inv = s
        L.check(lib.mac_read_step_fused(L.ptr(inv), L.ptr(kb), L.ptr(y), L.ptr(c), ctypes.byref(rw), L.ptr(info1),
                                        L.ptr(att1), B, N, d, L.stream_ptr()), "read_step_fused")
    unfused(sets[0])
    fused(sets[0])
    torch.cuda.synchronize()
    out = {"B": B, "N": N,
           "att_maxabs": float((att0 - att1).abs().max()), "att_ref_max": float(att0.max()),
           "info_maxabs": float((info0 - info1).abs().max()), "info_ref_maxabs": float(info0.abs().max()),
           "att_rowsum_err": float((att1.sum(1) - 1).abs().max()),
           "nan": bool(torch.isnan(att1).any() or torch.isnan(info1).any())}
    if time_it:
        for name, fn in (("unfused_us", unfused), ("fused_us", fused)):
            for s in sets:
                fn(s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 64
            e0.record()
            for i in range(iters):
                fn(sets[i % nsets])
            e1.record()
            torch.cuda.synchronize()
            out[name] = e0.elapsed_time(e1) * 1e3 / iters
    return out


def main():
    lib = L.load()
    torch.cuda.set_device(0)
    d = 512
    res = []
    for (B, N, tm) in ((2, 196, False), (3, 49, False), (5, 130, False), (2, 256, False), (7, 128, False), (3, 100, False),
                       (4, 17, False), (9, 200, False), (1, 129, False), (3, 255, False), (11, 131, False),
                       (64, 196, True), (64, 49, True), (384, 196, True)):
        r = case(lib, B, N, d, 100 + B + N, time_it=tm)
        print(json.dumps(r), flush=True)
        res.append(r)
    bad = [r for r in res if r["nan"] or r["att_maxabs"] > 2e-5 + 1e-3 * r["att_ref_max"] or
           r["info_maxabs"] > 1e-3 * max(1.0, r["info_ref_maxabs"])]
    print("FUSED_CHECK", "FAIL" if bad else "OK", len(res), "cases")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
