"""Persistent form of the LSTM recurrence (csrc/encoder.cu, `lstm_seq_kernel`): one thread-block cluster of 8 CTAs per
(direction, 8 batch rows) keeps the recurrent weights in shared memory for all S steps and exchanges h over DSMEM.
Opt-in with MAC_LSTM_PERSIST=1 (h == 256 only); checked against the fp64 oracle AND against the per-step kernels, forward
and (through the tensors it saves) backward.  Runs last: it is the newest kernel in the library."""
import os

import numpy as np
import pytest

from oracle.encoder_oracle import encoder_forward
from oracle import encoder_torch_autograd
from mac_network_b200.encoder import encoder_specs, init_encoder_params
from tests._util import max_rel


@pytest.mark.gpu
@pytest.mark.parametrize("B,S,keeps", [(64, 40, (1.0, 1.0)), (13, 11, (0.85, 0.92))])
def test_persistent_lstm_matches_oracle_and_step_kernels(B, S, keeps, monkeypatch):
    import torch
    from mac_network_b200.encoder import QuestionEncoder
    V, E, D = 90, 300, 512
    pv = init_encoder_params(encoder_specs(V, E, D), seed=51, dtype=np.float64)
    rng = np.random.RandomState(52)
    lengths = rng.randint(max(1, S // 2), S + 1, size=(B,)).astype(np.int32)
    lengths[0], lengths[1] = S, 1
    q = rng.randint(1, V + 1, size=(B, S)).astype(np.int32)
    q[np.arange(S)[None, :] >= lengths[:, None]] = 0
    dev = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in pv.items()}
    qd, ld = torch.from_numpy(q).cuda(), torch.from_numpy(lengths).cuda()
    d_cntx = rng.standard_normal((B, S, D)) / np.sqrt(S)
    d_vecq = rng.standard_normal((B, D))
    results = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MAC_LSTM_PERSIST", mode)
        enc = QuestionEncoder(dev, keep_input=keeps[0], keep_question=keeps[1], seed=9)
        _, cntx, vecq = enc.forward(qd, ld, step=2, save_for_backward=True)
        grads = {k: torch.zeros_like(v) for k, v in dev.items()}
        enc.backward(torch.from_numpy(d_cntx.astype(np.float32)).cuda(), torch.from_numpy(d_vecq.astype(np.float32)).cuda(), grads)
        torch.cuda.synchronize()
        results[mode] = (cntx.cpu().numpy(), vecq.cpu().numpy(), {k: g.cpu().numpy() for k, g in grads.items()})
        us = enc.dropout_uniforms(B, S, step=2)
    ref = encoder_forward(pv, q, lengths, keeps[0], keeps[1], uniforms=us)
    _, _, gref = encoder_torch_autograd.run(pv, q, lengths, keeps[0], keeps[1], us, d_cntx=d_cntx, d_vecq=d_vecq)
    for mode in ("0", "1"):
        cntx, vecq, grads = results[mode]
        assert max_rel(cntx, ref["questionCntxWords"]) < 1e-4, mode
        assert max_rel(vecq, ref["vecQuestions"]) < 1e-4, mode
        for k, gr in gref.items():
            assert max_rel(grads[k], gr) < 2e-4, (mode, k)
    # same arithmetic order per (row, unit) in both forms up to the k-loop's association: agreement far inside the tolerance
    assert max_rel(results["1"][0], results["0"][0]) < 1e-5
