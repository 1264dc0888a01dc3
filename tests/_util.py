"""Shared helpers for the tests: golden fixture loading and the case -> (cfg, inputs, params) rebuild."""
import glob
import json
import os

import numpy as np

from mac_network_b200.config import MACConfig
from mac_network_b200.params import init_params, perturb_biases
from mac_network_b200.synthetic import make_inputs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    """Cell fixtures (the output-unit fixtures `output_*.npz` have their own tests)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith(("output_", "stem_", "encoder_"))]


def load_golden(case):
    z = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    arrays = {k: z[k] for k in z.files if k != "meta_json"}
    return meta, arrays


def rebuild(meta, dtype=np.float64):
    cfg = MACConfig(**meta["cell_flags"]).validate()
    sh = meta["shape"]
    inputs = make_inputs(sh["B"], sh["S"], sh["N"], sh["d"], seed=meta["input_seed"], dtype=dtype)
    params = perturb_biases(init_params(cfg, sh["L"], seed=meta["param_seed"], dtype=dtype), seed=meta["bias_seed"])
    return cfg, inputs, params


def uniforms_of(meta, arrays, eval_skip=True):
    """Uniform draws in the reference's call order.  At eval the only draw the reference makes is the
    (all-ones at keep=1) variational mask, which the oracle and the product do not draw."""
    us = [arrays["uniform_%03d" % i] for i in range(meta["n_uniform"])]
    if not meta["train"]:
        return []
    return us


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / (np.maximum(np.abs(b), floor) if floor else np.abs(b).max() + 1e-300)))


def max_rel(a, b):
    """max |a-b| / max|b|  (tensor-level relative error, the 1e-4 bar of BASELINE.json)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))
