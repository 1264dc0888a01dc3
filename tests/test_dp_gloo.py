"""Data-parallel host logic on CPU: 2 ranks over gloo.  The gradient function is the fp64 torch-autograd oracle (the
CUDA kernels need a GPU); what is under test is the sharding, the flat bucket layout, the all-reduce and the loss
normalisation: the 2-rank all-reduced gradient must equal the 1-process gradient on the concatenated batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat(grads, specs, offsets):
    buf = np.zeros(offsets["__total__"], dtype=np.float64)
    for name, (shape, _) in specs.items():
        g = np.asarray(grads[name]).reshape(-1)
        buf[offsets[name]:offsets[name] + g.size] = g
    return buf


def _problem():
    from mac_network_b200.config import MACConfig
    from mac_network_b200.params import init_params, perturb_biases, param_specs
    from mac_network_b200.synthetic import make_inputs
    B, S, N, d, L = 4, 5, 6, 16, 2
    cfg = MACConfig.args("gqa", netLength=L, memDim=d, ctrlDim=d, attDim=d)
    inputs = make_inputs(B, S, N, d, seed=61, dtype=np.float64)
    params = perturb_biases(init_params(cfg, L, seed=62, dtype=np.float64), seed=63)
    rng = np.random.RandomState(64)
    return cfg, inputs, params, rng.standard_normal((B, d)), rng.standard_normal((B, d)), L, param_specs(cfg, L)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mac_network_b200.dp import shard_rows, allreduce_sum_
    from mac_network_b200.mac_cell import flat_layout
    from oracle import mac_torch_autograd as TA
    cfg, inputs, params, tc, tm, L, specs = _problem()
    B = inputs["knowledgeBase"].shape[0]
    rows = shard_rows(B, rank, world)
    local = {k: v[rows] for k, v in inputs.items()}
    _, _, g = TA.run(cfg, params, local, L, d_control=tc[rows] / B, d_memory=tm[rows] / B)
    offsets = flat_layout(specs)
    bucket = torch.from_numpy(_flat(g, specs, offsets))
    allreduce_sum_(bucket)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), bucket.numpy())
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process(tmp_path):
    from mac_network_b200.mac_cell import flat_layout
    from oracle import mac_torch_autograd as TA
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    cfg, inputs, params, tc, tm, L, specs = _problem()
    B = inputs["knowledgeBase"].shape[0]
    _, _, g = TA.run(cfg, params, inputs, L, d_control=tc / B, d_memory=tm / B)
    full = _flat(g, specs, flat_layout(specs))
    r0 = np.load(os.path.join(str(tmp_path), "rank0.npy"))
    r1 = np.load(os.path.join(str(tmp_path), "rank1.npy"))
    assert np.array_equal(r0, r1)                                    # replicas stay identical
    assert np.max(np.abs(r0 - full)) < 1e-12 * max(1.0, np.max(np.abs(full)))


def test_shard_rows():
    from mac_network_b200.dp import shard_rows
    assert [shard_rows(512, r, 8) for r in (0, 7)] == [slice(0, 64), slice(448, 512)]
    with pytest.raises(ValueError):
        shard_rows(10, 0, 4)


def test_adam_reference_matches_textbook():
    from mac_network_b200.dp import adam_reference
    p, m, v, ema, norm = adam_reference(np.ones(4), np.full(4, 10.0), np.zeros(4), np.zeros(4), np.ones(4), step=1)
    assert np.isclose(norm, 20.0)                      # clipped to 8: g = 4
    assert np.allclose(p, 1.0 - 1e-4, atol=1e-9)       # first Adam step moves by ~lr


# ---- the question input unit's variables in the same flat bucket (SURVEY section 8(f) rank 3 under section 8(e)'s sharding)
def _enc_problem():
    from mac_network_b200.encoder import encoder_specs, init_encoder_params
    B, S, V, E, D = 6, 5, 9, 8, 16
    specs = encoder_specs(V, E, D)
    params = init_encoder_params(specs, seed=71, dtype=np.float64)
    rng = np.random.RandomState(72)
    lengths = rng.randint(1, S + 1, size=(B,)).astype(np.int32)
    lengths[0] = S
    q = rng.randint(1, V + 1, size=(B, S)).astype(np.int32)
    q[np.arange(S)[None, :] >= lengths[:, None]] = 0
    return specs, params, q, lengths, rng.standard_normal((B, S, D)), rng.standard_normal((B, D))


def _enc_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mac_network_b200.dp import shard_rows, allreduce_sum_
    from mac_network_b200.mac_cell import flat_layout
    from oracle import encoder_torch_autograd as EA
    specs, params, q, lengths, dc, dq = _enc_problem()
    B = q.shape[0]
    rows = shard_rows(B, rank, world)
    _, _, g = EA.run(params, q[rows], lengths[rows], d_cntx=dc[rows] / B, d_vecq=dq[rows] / B)
    bucket = torch.from_numpy(_flat(g, specs, flat_layout(specs)))
    allreduce_sum_(bucket)
    np.save(os.path.join(out_dir, "enc_rank%d.npy" % rank), bucket.numpy())
    dist.destroy_process_group()


def test_two_rank_encoder_gradient_equals_single_process(tmp_path):
    """Rows of different lengths land on different ranks; the summed bucket (embedding rows included: a word used on both
    ranks gets both contributions) equals the gradient of the global-mean loss on the whole batch."""
    from mac_network_b200.mac_cell import flat_layout
    from oracle import encoder_torch_autograd as EA
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_enc_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    specs, params, q, lengths, dc, dq = _enc_problem()
    B = q.shape[0]
    _, _, g = EA.run(params, q, lengths, d_cntx=dc / B, d_vecq=dq / B)
    full = _flat(g, specs, flat_layout(specs))
    r0 = np.load(os.path.join(str(tmp_path), "enc_rank0.npy"))
    r1 = np.load(os.path.join(str(tmp_path), "enc_rank1.npy"))
    assert np.array_equal(r0, r1)
    assert np.max(np.abs(r0 - full)) < 1e-12 * max(1.0, np.max(np.abs(full)))
