"""The oracle restatement against fixtures produced by the UNMODIFIED reference cell (run on the TF1 shim).

CPU-only.  This is what pins `oracle/mac_oracle.py`: every per-step control / memory / info /
contControl and every attention map of every fixture must match to fp64 round-off.
"""
import numpy as np
import pytest

from oracle.mac_oracle import MACOracle
from mac_network_b200.params import param_specs
from tests._util import golden_cases, load_golden, rebuild, uniforms_of


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference_fixture(case):
    meta, gold = load_golden(case)
    cfg, inputs, params = rebuild(meta, np.float64)
    sh = meta["shape"]
    orc = MACOracle(cfg, params, dtype=np.float64)
    orc.train = meta["train"]
    dp = meta["dropouts"]
    orc.run(sh["L"], inputs["vecQuestions"], inputs["questionWords"], inputs["questionCntxWords"],
            inputs["questionLengths"], inputs["knowledgeBase"], memoryDropout=dp["memory"],
            readDropout=dp["read"], writeDropout=dp["write"], uniforms=uniforms_of(meta, gold))
    out = orc.outputs()
    stored32 = gold["control"].dtype == np.float32
    tol = 2e-6 if stored32 else 1e-12
    keys = [k for k in gold if not k.startswith("uniform_") and not k.startswith("final_")]
    assert set(keys) == set(out.keys()), (sorted(keys), sorted(out.keys()))
    for k in keys:
        g = gold[k].astype(np.float64)
        err = np.max(np.abs(out[k] - g)) / (np.max(np.abs(g)) + 1e-300)
        assert err < tol, (case, k, err)
    assert np.allclose(orc.trace[-1]["memory"], gold["final_memory"], rtol=0, atol=tol * 10)
    for k in gold:              # memoryBN: the stored statistics after the forward (moved only by a training forward)
        if k.startswith("final_bn_"):
            name = [n for n in orc.p if n.endswith("/BatchNorm/" + k[len("final_bn_"):])]
            assert len(name) == 1 and np.allclose(orc.p[name[0]], gold[k], rtol=0, atol=1e-12), (case, k)
            moved = not np.array_equal(orc.p[name[0]], params[name[0]])
            assert moved == bool(meta["train"]), (case, k, "statistics move in training only")
    # all draws consumed in train mode: the oracle makes the same dropout calls in the same order
    assert next(orc.uniforms, None) is None


@pytest.mark.parametrize("case", golden_cases())
def test_variable_names_match_reference(case):
    """the names/shapes the reference's variable scopes create are the checkpoint contract."""
    meta, _ = load_golden(case)
    cfg, _, _ = rebuild(meta)
    specs = param_specs(cfg, meta["shape"]["L"])
    assert {k: list(v[0]) for k, v in specs.items()} == meta["variables"]


def test_attention_properties():
    meta, gold = load_golden("args_small")
    lengths = rebuild(meta)[1]["questionLengths"]
    qa = gold["att_question"]
    assert np.allclose(qa.sum(-1), 1.0, atol=1e-12)
    for b, n in enumerate(lengths):
        assert np.all(qa[:, b, n:] == 0.0)          # -1e30 mask underflows to exactly 0 (ops.py:10, 245)
    assert np.allclose(gold["att_kb"].sum(-1), 1.0, atol=1e-12)
    g = load_golden("args4_small")[1]["att_gate"]
    assert np.all((g > 0) & (g < 1))


@pytest.mark.parametrize("variant", ["args", "gqa"])
def test_torch_cpu_port_matches_oracle(variant):
    """The timed CPU baseline (oracle/mac_torch_cpu.py, fp32) against the fp64 numpy oracle."""
    import torch
    from oracle.mac_torch_cpu import TorchCPUCell
    from mac_network_b200.config import MACConfig
    from mac_network_b200.params import init_params, perturb_biases
    from mac_network_b200.synthetic import make_inputs
    B, S, N, d, L = 4, 7, 12, 64, 3
    cfg = MACConfig.args(variant, netLength=L, memDim=d, ctrlDim=d, attDim=d)
    inp = make_inputs(B, S, N, d, seed=1, dtype=np.float64)
    params = perturb_biases(init_params(cfg, L, seed=2, dtype=np.float64), seed=3)
    ref = MACOracle(cfg, params).run(L, inp["vecQuestions"], inp["questionWords"], inp["questionCntxWords"],
                                     inp["questionLengths"], inp["knowledgeBase"])
    cell = TorchCPUCell(cfg, params, L)
    c, m, _ = cell.forward(torch.from_numpy(inp["vecQuestions"]).float(),
                           torch.from_numpy(inp["questionCntxWords"]).float(),
                           torch.from_numpy(inp["questionLengths"]).long(), torch.from_numpy(inp["knowledgeBase"]).float())
    assert np.max(np.abs(c.numpy() - ref.control)) / np.max(np.abs(ref.control)) < 1e-5
    assert np.max(np.abs(m.numpy() - ref.memory)) / np.max(np.abs(ref.memory)) < 1e-5


@pytest.mark.parametrize("case", ["args_train_small", "gqa_train_small", "args4_small", "args1_train_small"])
def test_torch_autograd_restatement_matches_reference_fixture(case):
    """The differentiable fp64 torch restatement (gradient oracle) reproduces the reference fixtures' forward."""
    from oracle import mac_torch_autograd as TA
    meta, gold = load_golden(case)
    cfg, inputs, params = rebuild(meta, np.float64)
    dp = meta["dropouts"]
    c, m, _ = TA.run(cfg, params, inputs, meta["shape"]["L"], (dp["memory"], dp["read"], dp["write"]),
                     uniforms_of(meta, gold))
    assert np.max(np.abs(m - gold["final_memory"])) < 1e-11
    assert np.max(np.abs(c - gold["final_control"])) < 1e-11
