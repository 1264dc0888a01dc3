"""Phase breakdown of the fused read-step kernel from in-kernel SM-clock stamps (profiling hook, see read_step.cuh):
0 start | 1 last P k-block scaled | 2 GEMM 1 complete | 3 H written | 4 GEMM 2 complete | 5 logits | 6 attention | 7 end."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mac_network_b200 import _lib as L  # noqa: E402
from profiles.check_read_fused import make_weights  # noqa: E402


def main():
    lib = L.load()
    raw = ctypes.CDLL(L.LIB_PATH)
    torch.cuda.set_device(0)
    d = 512
    flags = int(os.environ.get("RS_DBG_FLAGS", "0"))
    raw.mac_dbg_read_step_flags(flags)
    for (B, N) in ((64, 196),) if flags else ((64, 196), (64, 49)):
        t, rw = make_weights(lib, d, 7)
        g = torch.Generator(device="cuda").manual_seed(3)
        kb = torch.nn.functional.elu(torch.randn(B, N, d, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
        y = torch.randn(B, d, device="cuda", generator=g)
        c = torch.randn(B, d, device="cuda", generator=g)
        nb = lib.mac_read_invariant_bytes(B, N, d, 1)
        inv = torch.empty(nb, dtype=torch.uint8, device="cuda")
        L.check(lib.mac_read_invariant(None, L.ptr(kb), ctypes.byref(rw), 1, L.ptr(inv), nb, B, N, d, L.stream_ptr()))
        info, att = torch.empty(B, d, device="cuda"), torch.empty(B, N, device="cuda")
        grid = 2 * B if N > 128 else (B + min(128 // N, 2) - 1) // min(128 // N, 2)
        dbg = torch.zeros(grid, 64, dtype=torch.int64, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

        def run():
            L.check(lib.mac_read_step_fused(L.ptr(inv), L.ptr(kb), L.ptr(y), L.ptr(c), ctypes.byref(rw), L.ptr(info),
                                            L.ptr(att), B, N, d, L.stream_ptr()))
        for _ in range(3):
            run()
        for cold in (False, True):
            if cold:
                flush.zero_()
            raw.mac_dbg_read_step_timestamps(ctypes.c_void_p(dbg.data_ptr()))
            run()
            torch.cuda.synchronize()
            raw.mac_dbg_read_step_timestamps(None)
            full = dbg.cpu().numpy().astype(np.float64)
            lead_all = full[0::2]
            full = full[full[:, 7] != 0]               # the packed form launches fewer CTAs than the buffer has rows
            s = full[:, :8]
            dl = np.diff(s, axis=1)
            names = ["scale(GEMM1 feed)", "GEMM1 tail", "epilogue1", "GEMM2", "epilogue2", "softmax", "weighted sum"]
            out = {"B": B, "N": N, "cold_L2": cold, "grid": int(full.shape[0]), "dbg_flags": flags,
                   "total_clk_median": float(np.median(s[:, 7] - s[:, 0])), "total_clk_max": float(np.max(s[:, 7] - s[:, 0])),
                   "phases_median_clk": {n: float(np.median(dl[:, i])) for i, n in enumerate(names)},
                   "phases_max_clk": {n: float(np.max(dl[:, i])) for i, n in enumerate(names)}}
            if N > 128 and full[:, 8:].any():
                lead = lead_all[lead_all[:, 7] != 0]   # leader CTAs (rank 0 of each pair)
                rel = lambda a, b: [float(np.median(lead[:, a + i] - lead[:, 0])) if lead[:, a + i].any() else None for i in range(b)]
                out["pair_pipeline_clk_from_start_leader_median"] = {
                    "tma_issue_kb": rel(16, 8), "a_landed_kb": rel(24, 8), "a_scaled_both_kb": rel(48, 8),
                    "mma_issue_kb": rel(8, 8), "g1_done_seen_by_tma": rel(56, 1), "gemm2_mma_issue_i": rel(32, 16),
                    "stamps": [float(np.median(lead[:, i] - lead[:, 0])) for i in range(8)]}
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
