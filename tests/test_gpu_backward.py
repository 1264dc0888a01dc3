"""Hand-written backward kernels against torch.autograd on the fp64 restatement (the reference uses TF autodiff)."""
import numpy as np
import pytest
import torch

from mac_network_b200.config import MACConfig
from mac_network_b200.params import init_params, perturb_biases
from mac_network_b200.synthetic import make_inputs
from tests._util import max_rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant,shape,dp", [
    ("args", (6, 9, 50, 64, 3), (1.0, 1.0, 1.0)),
    ("args", (8, 12, 196, 512, 2), (0.85, 0.85, 1.0)),
    ("gqa", (5, 7, 49, 128, 4), (0.9, 0.8, 0.9)),
    ("args4", (4, 6, 20, 64, 3), (1.0, 0.85, 1.0)),
    ("args3", (4, 6, 20, 64, 3), (0.85, 1.0, 1.0)),
    ("args1", (5, 7, 20, 64, 4), (0.85, 0.85, 1.0)),      # recurrent control chain (controlFeedPrev)
])
def test_backward_matches_autograd(variant, shape, dp):
    from mac_network_b200.autograd import mac_backward
    from mac_network_b200.mac_cell import MACCell, MACParams, mac_network
    from oracle import mac_torch_autograd as TA
    B, S, N, d, L = shape
    over = dict(netLength=L, memDim=d, ctrlDim=d, attDim=d)
    if dp[2] < 1.0:
        over["writeDropout"] = dp[2]
    cfg = MACConfig.args(variant, **over)
    inputs = make_inputs(B, S, N, d, seed=51, dtype=np.float64)
    pv = perturb_biases(init_params(cfg, L, seed=52, dtype=np.float64), seed=53)
    rng = np.random.RandomState(54)
    gc, gm = rng.standard_normal((B, d)), rng.standard_normal((B, d))
    params = MACParams(cfg, L, values={k: v.astype(np.float32) for k, v in pv.items()})
    x = {k: torch.from_numpy(np.ascontiguousarray(v if v.dtype == np.int32 else v.astype(np.float32))).cuda()
         for k, v in inputs.items()}
    cell = MACCell(x["vecQuestions"], x["questionWords"], x["questionCntxWords"], x["questionLengths"],
                   x["knowledgeBase"], dp[0], dp[1], dp[2], B, True, config=cfg, params=params, seed=4242,
                   save_for_backward=True)
    control, memory = mac_network(cell, L)
    grads = mac_backward(cell, torch.from_numpy(gc.astype(np.float32)).cuda(), torch.from_numpy(gm.astype(np.float32)).cuda())
    torch.cuda.synchronize()
    rc, rm, rg = TA.run(cfg, pv, inputs, L, dp, cell.dropout_uniforms(), gc, gm)
    assert max_rel(memory.cpu().numpy(), rm) < 1e-4 and max_rel(control.cpu().numpy(), rc) < 1e-4
    worst = {}
    for k, ref in rg.items():
        got = grads[k].cpu().numpy()
        assert got.size == ref.size, k
        got = got.reshape(ref.shape)
        scale = np.max(np.abs(ref))
        if scale < 1e-12:
            # e.g. the KB-softmax logit bias: its true gradient is exactly 0 (softmax is shift invariant); fp32
            # round-off of the explicit sum is all that is left
            assert np.max(np.abs(got)) < 1e-4, k
            continue
        worst[k] = float(np.max(np.abs(got - ref)) / scale)
    bad = {k: v for k, v in worst.items() if v > 2e-4}
    assert not bad, (bad, {k: round(v, 7) for k, v in worst.items()})


def test_dp_half_batches_sum_to_full_batch_and_optimizer_step():
    """Single-GPU check of the DP arithmetic: gradients of the two half-batch shards (each scaled by 1/B_global) add up
    to the full-batch gradient (eval-mode dropouts so the shards see the same function), and the fused
    clip+Adam+EMA kernel matches the numpy restatement."""
    from mac_network_b200.dp import DPTrainer, adam_reference, shard_rows
    B, S, N, d, L = 8, 6, 20, 64, 2
    cfg = MACConfig.args("gqa", netLength=L, memDim=d, ctrlDim=d, attDim=d)
    inputs = make_inputs(B, S, N, d, seed=71)
    pv = perturb_biases(init_params(cfg, L, seed=72), seed=73)
    rng = np.random.RandomState(74)
    tc = torch.from_numpy(rng.standard_normal((B, d)).astype(np.float32)).cuda()
    tm = torch.from_numpy(rng.standard_normal((B, d)).astype(np.float32)).cuda()
    full = {k: torch.from_numpy(v).cuda() for k, v in inputs.items()}
    tr = DPTrainer(cfg, L, param_values=pv, dropouts=(1.0, 1.0, 1.0))
    tr.grads("full", full, tc, tm, B)
    g_full = tr.bucket.clone()
    acc = torch.zeros_like(g_full)
    for r in range(2):
        rows = shard_rows(B, r, 2)
        half = {k: v[rows].contiguous() for k, v in full.items()}
        tr.grads("half%d" % r, half, tc[rows].contiguous(), tm[rows].contiguous(), B)
        acc += tr.bucket
    torch.cuda.synchronize()
    assert max_rel(acc.cpu().numpy(), g_full.cpu().numpy()) < 1e-5
    # optimizer step
    tr.bucket.copy_(g_full)
    p0, m0, v0, e0 = (t.cpu().numpy().copy() for t in (tr.params.flat, tr.adam_m, tr.adam_v, tr.ema))
    tr.apply()
    torch.cuda.synchronize()
    p1, m1, v1, e1, norm = adam_reference(p0, g_full.cpu().numpy(), m0, v0, e0, step=1, clip=8.0)
    assert abs(float(tr.norm[0]) - norm) < 1e-4 * norm
    assert np.max(np.abs(tr.params.flat.cpu().numpy() - p1)) < 1e-6
    assert np.max(np.abs(tr.ema.cpu().numpy() - e1)) < 1e-6
    assert max_rel(tr.adam_v.cpu().numpy(), v1) < 2e-4      # (1 - beta2) is rounded in fp32, as in TF


def test_training_on_the_reference_loss_decreases():
    """DPTrainer with the output unit: a few steps of clip/Adam on the mean softmax-CE must reduce the loss on a fixed
    batch, and the step must leave split-K counters / buckets in a reusable state."""
    from mac_network_b200.dp import DPTrainer
    B, S, N, d, L = 16, 8, 49, 128, 3
    cfg = MACConfig.args("gqa", netLength=L, memDim=d, ctrlDim=d, attDim=d)
    inputs = make_inputs(B, S, N, d, seed=91)
    batch = {k: torch.from_numpy(v).cuda() for k, v in inputs.items()}
    answers = torch.from_numpy(np.random.RandomState(92).randint(0, 12, size=(B,)).astype(np.int32)).cuda()
    tr = DPTrainer(cfg, L, seed=5, lr=3e-3, classifier=(12, [64]), dropouts=(1.0, 1.0, 1.0), output_dropout=1.0)
    losses = []
    for _ in range(12):
        _, ls = tr.train_step_answers(0, batch, answers, B)
        losses.append(float(ls.mean()))
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.7 * losses[0], losses
