"""Question input unit: oracle vs fixtures from the reference's own MACnet.qEmbeddingsOp +
MACnet.encoder on the TF1 shim (CPU); product (embedding/dropout kernel, hoisted input GEMM, per-step bi-LSTM kernel,
BPTT) vs the oracle and torch.autograd on the GPU."""
import json
import os

import numpy as np
import pytest

from oracle.encoder_oracle import encoder_forward
from oracle import encoder_torch_autograd
from mac_network_b200.encoder import encoder_specs, init_encoder_params
from tests._util import GOLDEN_DIR, max_rel

CASES = ["encoder_eval", "encoder_train", "encoder_proj", "encoder_uni"]


def _load(case):
    z = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    return meta, {k: z[k] for k in z.files if k != "meta_json"}


def _rebuild(meta):
    sh = meta["shape"]
    specs = encoder_specs(sh["V"], sh["E"], sh["encDim"], ctrl_dim=sh["ctrlDim"], bi=meta.get("bi", True), proj=meta["proj"])
    return specs, init_encoder_params(specs, seed=meta["param_seed"], dtype=np.float64)


@pytest.mark.parametrize("case", CASES)
def test_encoder_oracle_matches_reference_fixture(case):
    meta, g = _load(case)
    specs, params = _rebuild(meta)
    assert {k: list(v[0]) for k, v in specs.items()} == meta["variables"]      # names/shapes the reference created
    us = [g["uniform_%03d" % i] for i in range(meta["n_uniform"])]
    out = encoder_forward(params, g["qIndices"], g["questionLengths"], keep_input=meta["keep_input"],
                          keep_question=meta["keep_question"], uniforms=us, proj=meta["proj"])
    for k in ("questionWords", "questionCntxWords", "vecQuestions"):
        assert np.max(np.abs(out[k] - g[k])) < 1e-12, k
    # dynamic_rnn semantics the cell relies on: outputs past the question end are exactly zero (before any projection)
    if not meta["proj"]:
        S = g["qIndices"].shape[1]
        pad = np.arange(S)[None, :] >= g["questionLengths"][:, None]
        assert np.all(out["questionCntxWords"][pad] == 0)


@pytest.mark.parametrize("case", CASES)
def test_encoder_torch_restatement_matches_oracle(case):
    meta, g = _load(case)
    _, params = _rebuild(meta)
    us = [g["uniform_%03d" % i] for i in range(meta["n_uniform"])]
    cntx, vecq, _ = encoder_torch_autograd.run(params, g["qIndices"], g["questionLengths"], meta["keep_input"],
                                               meta["keep_question"], us)
    assert np.max(np.abs(cntx - g["questionCntxWords"])) < 1e-12
    assert np.max(np.abs(vecq - g["vecQuestions"])) < 1e-12


def _device_params(pv):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in pv.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_encoder_gpu_matches_reference_fixture(case):
    """Eval fixtures directly; the training fixture through the oracle fed with the kernels' own Philox draws."""
    import torch
    from mac_network_b200.encoder import QuestionEncoder
    meta, g = _load(case)
    _, pv = _rebuild(meta)
    enc = QuestionEncoder(_device_params(pv), keep_input=meta["keep_input"], keep_question=meta["keep_question"], seed=77)
    q = torch.from_numpy(g["qIndices"]).cuda()
    lens = torch.from_numpy(g["questionLengths"]).cuda()
    words, cntx, vecq = enc.forward(q, lens, step=3)
    torch.cuda.synchronize()
    if meta["train"]:
        B, S = g["qIndices"].shape
        ref = encoder_forward(pv, g["qIndices"], g["questionLengths"], meta["keep_input"], meta["keep_question"],
                              uniforms=enc.dropout_uniforms(B, S, step=3), proj=meta["proj"])
    else:
        ref = {k: g[k] for k in ("questionWords", "questionCntxWords", "vecQuestions")}
    assert max_rel(words.cpu().numpy(), ref["questionWords"]) < 1e-6
    for got, key in ((cntx, "questionCntxWords"), (vecq, "vecQuestions")):
        err = max_rel(got.cpu().numpy(), ref[key])
        print("%s %s max-rel %.2e" % (case, key, err))
        assert err < 1e-4, (key, err)


def _random_batch(B, S, V, seed):
    rng = np.random.RandomState(seed)
    lengths = rng.randint(max(1, S // 2), S + 1, size=(B,)).astype(np.int32)
    lengths[0] = S
    if B > 1:
        lengths[1] = 1
    q = rng.randint(1, V + 1, size=(B, S)).astype(np.int32)
    q[np.arange(S)[None, :] >= lengths[:, None]] = 0
    return q, lengths


@pytest.mark.gpu
@pytest.mark.parametrize("shape,keeps,proj", [
    (dict(B=64, S=40, V=90, E=300, encDim=512), (1.0, 1.0), False),        # BASELINE configs[2] question shapes
    (dict(B=64, S=40, V=90, E=300, encDim=512), (0.85, 0.92), False),      # config.py:202, 206 training dropouts
    (dict(B=70, S=9, V=20, E=20, encDim=48), (0.85, 0.92), True),          # two row chunks, ragged tail, projections
])
def test_encoder_gpu_forward_backward(shape, keeps, proj):
    """Full-size forward against the fp64 oracle and BPTT against torch.autograd on the fp64 restatement."""
    import torch
    from mac_network_b200.encoder import QuestionEncoder
    B, S, V, E, D = (shape[k] for k in ("B", "S", "V", "E", "encDim"))
    ctrl = D + 16 if proj else D
    specs = encoder_specs(V, E, D, ctrl_dim=ctrl, bi=True, proj=proj)
    pv = init_encoder_params(specs, seed=41, dtype=np.float64)
    q, lengths = _random_batch(B, S, V, seed=42)
    dev = _device_params(pv)
    enc = QuestionEncoder(dev, keep_input=keeps[0], keep_question=keeps[1], seed=5)
    qd, ld = torch.from_numpy(q).cuda(), torch.from_numpy(lengths).cuda()
    words, cntx, vecq = enc.forward(qd, ld, step=1, save_for_backward=True)
    rng = np.random.RandomState(43)
    d_cntx = rng.standard_normal(cntx.shape) / np.sqrt(S)
    d_vecq = rng.standard_normal(vecq.shape)
    grads = {k: torch.zeros_like(v) for k, v in dev.items()}
    enc.backward(torch.from_numpy(d_cntx.astype(np.float32)).cuda(), torch.from_numpy(d_vecq.astype(np.float32)).cuda(), grads)
    torch.cuda.synchronize()
    us = enc.dropout_uniforms(B, S, step=1)
    ref = encoder_forward(pv, q, lengths, keeps[0], keeps[1], uniforms=us, proj=proj)
    for got, key in ((cntx, "questionCntxWords"), (vecq, "vecQuestions")):
        err = max_rel(got.cpu().numpy(), ref[key])
        print("fwd %s max-rel %.2e" % (key, err))
        assert err < 1e-4, (key, err)
    pad = np.arange(S)[None, :] >= lengths[:, None]
    if not proj:
        assert np.all(cntx.cpu().numpy()[pad] == 0)
    _, _, gref = encoder_torch_autograd.run(pv, q, lengths, keeps[0], keeps[1], us, d_cntx=d_cntx, d_vecq=d_vecq)
    for k, gr in gref.items():
        err = max_rel(grads[k].cpu().numpy(), gr)
        print("grad %-70s max-rel %.2e" % (k, err))
        assert err < 2e-4, (k, err)


@pytest.mark.gpu
def test_encoder_feeds_cell():
    """encoder -> MACCell: the cell accepts the encoder's outputs as its question inputs (model.py:787-802)."""
    import torch
    from mac_network_b200.config import MACConfig
    from mac_network_b200.encoder import QuestionEncoder
    from mac_network_b200.mac_cell import MACCell, MACParams, mac_network
    from mac_network_b200.params import init_params, perturb_biases
    from oracle.mac_oracle import MACOracle
    B, S, V, E, d, N, L = 6, 8, 15, 12, 64, 10, 2
    specs = encoder_specs(V, E, d)
    pv = init_encoder_params(specs, seed=1, dtype=np.float64)
    q, lengths = _random_batch(B, S, V, seed=2)
    enc = QuestionEncoder(_device_params(pv))
    words, cntx, vecq = enc.forward(torch.from_numpy(q).cuda(), torch.from_numpy(lengths).cuda())
    cfg = MACConfig.args("args", netLength=L, memDim=d, ctrlDim=d, attDim=d)
    cv = perturb_biases(init_params(cfg, L, seed=3), seed=4)
    kb = np.random.RandomState(5).standard_normal((B, N, d)).astype(np.float32)
    cell = MACCell(vecq, words, cntx, torch.from_numpy(lengths).cuda(), torch.from_numpy(kb).cuda(), 1.0, 1.0, 1.0, B, False,
                   config=cfg, params=MACParams(cfg, L, values=cv))
    control, memory = mac_network(cell, L)
    torch.cuda.synchronize()
    eo = encoder_forward(pv, q, lengths)
    ref = MACOracle(cfg, cv, dtype=np.float64).run(L, eo["vecQuestions"], eo["questionWords"], eo["questionCntxWords"],
                                                    lengths, kb.astype(np.float64))
    assert max_rel(memory.cpu().numpy(), ref.memory) < 1e-4
    assert max_rel(control.cpu().numpy(), ref.control) < 1e-4
