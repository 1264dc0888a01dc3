"""Size-independent properties of the oracle (the checker must itself behave like the reference's math): batch rows are
independent, the read unit is invariant to a permutation of the knowledge-base cells, padded question words and padded
embedding ids cannot influence anything, attention rows are probability distributions.  Seeds drawn by hypothesis."""
import numpy as np
from hypothesis import given, settings, strategies as st

from mac_network_b200.config import MACConfig
from mac_network_b200.encoder import encoder_specs, init_encoder_params
from mac_network_b200.params import init_params, perturb_biases
from mac_network_b200.synthetic import make_inputs
from oracle.encoder_oracle import encoder_forward
from oracle.mac_oracle import MACOracle

B, S, N, d, L = 4, 6, 7, 16, 3


def _run(cfg, pv, inp):
    orc = MACOracle(cfg, pv, dtype=np.float64)
    orc.run(L, inp["vecQuestions"], inp["questionWords"], inp["questionCntxWords"], inp["questionLengths"], inp["knowledgeBase"])
    return orc.outputs()


@settings(max_examples=8, deadline=None)
@given(seed=st.integers(0, 10_000), variant=st.sampled_from(["args", "args1", "gqa"]))
def test_cell_oracle_invariances(seed, variant):
    cfg = MACConfig.args(variant, netLength=L, memDim=d, ctrlDim=d, attDim=d)
    pv = perturb_biases(init_params(cfg, L, seed=seed, dtype=np.float64), seed=seed + 1)
    inp = make_inputs(B, S, N, d, seed=seed + 2, dtype=np.float64)
    base = _run(cfg, pv, inp)
    rng = np.random.RandomState(seed + 3)
    # (1) attention rows are distributions; masked question positions carry exactly zero weight
    assert np.allclose(base["att_kb"].sum(-1), 1.0, atol=1e-12) and np.allclose(base["att_question"].sum(-1), 1.0, atol=1e-12)
    pad = np.arange(S)[None, :] >= inp["questionLengths"][:, None]
    assert np.all(base["att_question"][:, pad] == 0.0)
    # (2) batch rows are independent: permuting the batch permutes every output
    perm = rng.permutation(B)
    out = _run(cfg, pv, {k: v[perm] for k, v in inp.items()})
    for k in ("control", "memory", "att_kb"):
        assert np.allclose(out[k], base[k][:, perm], atol=1e-12), k
    # (3) the knowledge base is a SET of cells: permuting them permutes the KB attention and changes nothing else
    pn = rng.permutation(N)
    inp2 = dict(inp)
    inp2["knowledgeBase"] = inp["knowledgeBase"][:, pn]
    out = _run(cfg, pv, inp2)
    assert np.allclose(out["memory"], base["memory"], atol=1e-10) and np.allclose(out["att_kb"], base["att_kb"][..., pn], atol=1e-12)
    # (4) question words past a question's length cannot matter
    inp3 = dict(inp)
    junk = inp["questionCntxWords"].copy()
    junk[pad] = rng.standard_normal((int(pad.sum()), d))
    inp3["questionCntxWords"] = junk
    out = _run(cfg, pv, inp3)
    assert np.array_equal(out["memory"], base["memory"]) and np.array_equal(out["control"], base["control"])


@settings(max_examples=8, deadline=None)
@given(seed=st.integers(0, 10_000))
def test_encoder_oracle_invariances(seed):
    V, E, D = 9, 8, 12
    pv = init_encoder_params(encoder_specs(V, E, D), seed=seed, dtype=np.float64)
    rng = np.random.RandomState(seed + 1)
    lengths = rng.randint(1, S + 1, size=(B,)).astype(np.int32)
    q = rng.randint(1, V + 1, size=(B, S)).astype(np.int32)
    pad = np.arange(S)[None, :] >= lengths[:, None]
    base = encoder_forward(pv, np.where(pad, 0, q), lengths)
    # ids in the padded positions are never read (dynamic_rnn stops at the length; the backward cell starts at len-1)
    other = encoder_forward(pv, q, lengths)
    assert np.array_equal(other["questionCntxWords"], base["questionCntxWords"])
    assert np.array_equal(other["vecQuestions"], base["vecQuestions"])
    assert np.all(base["questionCntxWords"][pad] == 0.0)
    # a question's encoding does not depend on how far the batch is padded (the reference trims batches, model.py:681-687)
    wide = encoder_forward(pv, np.concatenate([np.where(pad, 0, q), np.zeros((B, 3), np.int32)], axis=1), lengths)
    assert np.array_equal(wide["questionCntxWords"][:, :S], base["questionCntxWords"])
    assert np.array_equal(wide["vecQuestions"], base["vecQuestions"])
    # the final state is the forward output at len-1 and the backward output at 0 (ops.py:893-898)
    h = D // 2
    rows = np.arange(B)
    assert np.array_equal(base["vecQuestions"][:, :h], base["questionCntxWords"][rows, lengths - 1, :h])
    assert np.array_equal(base["vecQuestions"][:, h:], base["questionCntxWords"][rows, 0, h:])
