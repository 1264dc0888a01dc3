"""Knob sweep of the resident-input inference arm inside ONE process (env knobs are read at launch time, CUDA graphs are
re-captured per configuration): tile shape of the tcgen05 GEMM (MAC_TC_TILE), cta_group::2 pair kernel (MAC_TC_PAIR),
number of independent passes in flight.  Prints one JSON line per configuration.
    python profiles/sweep.py [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from mac_network_b200.config import MACConfig  # noqa: E402
from mac_network_b200.mac_cell import MACParams  # noqa: E402
from mac_network_b200.params import init_params, perturb_biases  # noqa: E402
from mac_network_b200.synthetic import SHAPES  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
torch.cuda.set_device(0)
shape = SHAPES["headline"]
B, S, N, d, L = shape
cfg = MACConfig.args("args", netLength=L)
params = MACParams(cfg, L, values=perturb_biases(init_params(cfg, L, seed=100), seed=101))
main = torch.cuda.current_stream()


def measure(nstreams, nslots=8):
    slots = [bench.Slot(cfg, params, shape, 1234 + s, "bf16", True) for s in range(nslots)]
    side = [torch.cuda.Stream() for _ in range(nstreams - 1)]

    def run(n):
        fork = torch.cuda.Event()
        fork.record(main)
        for st in side:
            st.wait_event(fork)
        for k in range(n):
            j = k % nstreams
            if j == 0:
                slots[k % nslots].run()
            else:
                with torch.cuda.stream(side[j - 1]):
                    slots[k % nslots].run()
        for st in side:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
    run(8)
    torch.cuda.synchronize()
    best = None
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        best = t if best is None else min(best, t)
    del slots
    torch.cuda.empty_cache()
    return steps * L / best


configs = [({}, "default")]
for tile in ("128128", "128256", "256256"):
    configs.append(({"MAC_TC_TILE": tile}, "tile=" + tile))
configs.append(({"MAC_TC_PAIR": "1"}, "pair=1"))
configs.append(({"MAC_NO_FOLD_Y": "1"}, "no_fold_y"))
for env, name in configs:
    for k in ("MAC_TC_TILE", "MAC_TC_PAIR", "MAC_NO_FOLD_Y", "MAC_READ_QHOIST"):
        os.environ.pop(k, None)
    os.environ.update(env)
    params.touch()
    for ns in ((1, 2, 3, 4, 6, 8) if name == "default" else (2, 4, 6)):
        try:
            v = measure(ns)
            print(json.dumps({"config": name, "streams": ns, "reasoning_steps_per_s": round(v, 1)}), flush=True)
        except Exception as exc:
            print(json.dumps({"config": name, "streams": ns, "error": repr(exc)[:200]}), flush=True)
