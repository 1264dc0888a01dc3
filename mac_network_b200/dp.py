"""Data-parallel training of the MAC cell: one process per GPU, the batch sharded over ranks, ONE collective per step.

The reference's multi-GPU path is a stub (towers >= 1 are ignored, `model.py:671-679`), so there is nothing to match
except single-process semantics: the all-reduced gradient of the global-mean loss must equal the 1-process gradient on
the concatenated batch.  Per step:
    forward (train-mode dropouts, per-rank Philox streams) -> hand-written backward into the flat gradient bucket ->
    `all_reduce(SUM)` of the bucket over NCCL/NVLink -> global-norm clip, Adam, EMA fused in one kernel (replicated).
The loss of this cell-level harness is a linear probe of the final state, sum(memory_L * t_m + control_L * t_c) / B_global,
standing in for the out-of-scope output unit / classifier (it supplies dL/dmemory and dL/dcontrol exactly as they would).
"""

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def shard_rows(global_batch, rank, world):
    """Rank r takes samples [r*B/world, (r+1)*B/world) of the global batch (mirrors initTowerBatch, model.py:139-149)."""
    if global_batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


def allreduce_sum_(bucket, group=None):
    """The path's only exchange step.  Returns the bucket (summed over ranks in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    return bucket


class DPTrainer(object):
    def __init__(self, cfg, netLength, param_values=None, seed=0, rank=0, world=1, lr=1e-4, clip=8.0, ema_decay=0.999,
                 beta1=0.9, beta2=0.999, eps=1e-8, dropouts=None, device="cuda", classifier=None, output_dropout=0.85,
                 encoder=None, stem=None, enc_dropouts=(0.85, 0.92), stem_dropout=0.82, prec="fp32", bwd_tc=False):
        """`classifier=(answerWordsNum, outClassifierDims)` adds the reference's output unit + answer loss
        (model.py:512-528, 547-576, 593-596); `encoder=(vocabulary rows, wrdEmbDim)` the question input unit
        (model.py:208-220, 279-307) and `stem=(imageInDim, stemNumLayers)` the image stem (model.py:165-204), with the
        reference's training dropouts (config.py:202-206).  All variables join the same flat buckets, so the one
        all-reduce and the one fused optimizer pass cover the whole model (`train_step_full`).
        `prec="bf16"` runs the read unit's forward projections on tensor cores in training too (activations saved in bf16,
        widened for the backward); `bwd_tc=True` runs its six backward products on tensor cores (`mac_read_bwd_tc`).
        Both are mixed precision: bf16 operands, fp32 accumulation, fp32 master weights / gradients / optimizer state."""
        from .mac_cell import MACParams, views_of
        from .params import init_params
        self.cfg, self.L, self.rank, self.world = cfg, netLength, rank, world
        self.prec, self.bwd_tc = prec, bool(bwd_tc)
        self.lib = _lib.load()
        extra_specs = extra_values = None
        if classifier is not None:
            from .output_unit import output_specs, init_output_params
            extra_specs = output_specs(cfg.ctrlDim, cfg.memDim, list(classifier[1]), classifier[0])
            extra_values = init_output_params(extra_specs, seed=seed + 17, bias_scale=0.0)
        self._enc_specs = self._stem_specs = None
        if encoder is not None or stem is not None:
            import collections
            if classifier is None or encoder is None or stem is None:
                raise ValueError("the full model needs classifier=, encoder= and stem= together")
            from .encoder import encoder_specs, init_encoder_params
            from .stem import stem_specs, init_stem_params
            self._enc_specs = encoder_specs(encoder[0], encoder[1], cfg.ctrlDim, ctrl_dim=cfg.ctrlDim, bi=True)
            self._stem_specs = stem_specs(stem[0], cfg.memDim, num_layers=stem[1])
            extra_specs = collections.OrderedDict(list(extra_specs.items()) + list(self._enc_specs.items())
                                                  + list(self._stem_specs.items()))
            extra_values = dict(extra_values)
            extra_values.update(init_encoder_params(self._enc_specs, seed=seed + 19, bias_scale=0.0))   # TF: zero biases
            extra_values.update(init_stem_params(self._stem_specs, seed=seed + 23, bias_scale=0.0))
        if extra_specs is not None and param_values is None:
            param_values = init_params(cfg, netLength, seed=seed)
        self.params = MACParams(cfg, netLength, values=param_values, seed=seed, device=device, extra_specs=extra_specs,
                                extra_values=extra_values)   # replicated
        self.out = None
        if classifier is not None:
            from .output_unit import OutputUnit
            self.out = OutputUnit({k: self.params.t[k] for k in extra_specs
                                   if k.startswith(("outputUnit/", "classifier/"))}, relu=cfg.relu, keep=output_dropout,
                                  seed=seed, version=lambda: self.params.version)
            self._views_of = views_of
        self.enc = self.stem = None
        if self._enc_specs is not None:
            from .encoder import QuestionEncoder
            from .stem import Stem
            self.enc = QuestionEncoder({k: self.params.t[k] for k in self._enc_specs}, keep_input=enc_dropouts[0],
                                       keep_question=enc_dropouts[1], seed=seed)
            self.stem = Stem({k: self.params.t[k] for k in self._stem_specs}, relu=cfg.relu, prec="fp32", seed=seed,
                             version=lambda: self.params.version)
            self.stem_dropout = float(stem_dropout)
            self._full_bufs = {}
        n = self.params.numel
        z = lambda: torch.zeros(n, dtype=torch.float32, device=self.params.device)
        self.bucket, self.adam_m, self.adam_v = z(), z(), z()
        self.ema = self.params.flat.clone()
        self.norm = torch.zeros(2, dtype=torch.float32, device=self.params.device)
        self.ows_bytes = int(self.lib.mac_optimizer_workspace_bytes())
        self.ows = torch.zeros(self.ows_bytes, dtype=torch.uint8, device=self.params.device)
        self.hp = dict(lr=lr, clip=clip, ema=ema_decay, b1=beta1, b2=beta2, eps=eps)
        self.dropouts = dropouts or (cfg.memoryDropout, cfg.readDropout, cfg.writeDropout)
        self.step_id = 0
        self.base_seed = seed
        self._cells = {}

    MAX_CACHED_CELLS = 4      # a save-for-backward cell holds L x (3 B N d + B d) floats (~1.2 GB at B=64, N=196, d=512, L=16)

    def cell_for(self, key, batch):
        """The training cell of shape-key `key`, fed with `batch`.

        The reference cell captures its input tensors at construction (mac_cell.py:59-79), and so does `MACCell`; the
        trainer therefore owns PERSISTENT input buffers per key and copies the caller's batch into them on every call, so
        a cached cell can never run on the tensors of an earlier batch.  A batch that already lives in these
        buffers (`full_forward_backward` writes into them directly) is not copied again.  At most MAX_CACHED_CELLS keys
        are kept (least recently used first out): callers with many distinct shapes -- e.g. one per trimmed question
        length -- should bucket / pad instead (attention masks the padding)."""
        from .mac_cell import MACCell
        names = ("vecQuestions", "questionWords", "questionCntxWords", "questionLengths", "knowledgeBase")
        ent = self._cells.pop(key, None)
        if ent is None:
            bufs = {}
            for n in names:
                src = batch[n]
                bufs[n] = src.to(torch.int32).clone().contiguous() if n == "questionLengths" else src.clone().contiguous()
            dm, dr, dw = self.dropouts
            cell = MACCell(bufs["vecQuestions"], bufs["questionWords"], bufs["questionCntxWords"], bufs["questionLengths"],
                           bufs["knowledgeBase"], dm, dr, dw, bufs["knowledgeBase"].shape[0], True, config=self.cfg,
                           params=self.params, prec=self.prec, save_for_backward=True)
            ent = (cell, bufs)
            while len(self._cells) >= self.MAX_CACHED_CELLS:
                old_key = next(iter(self._cells))
                del self._cells[old_key]
                if hasattr(self, "_full_bufs"):
                    self._full_bufs.pop(old_key, None)
        else:
            cell, bufs = ent
            for n in names:
                src = batch[n]
                if src.data_ptr() != bufs[n].data_ptr():
                    if tuple(src.shape) != tuple(bufs[n].shape):
                        raise ValueError("batch tensor %s has shape %s but key %r was built for %s"
                                         % (n, tuple(src.shape), key, tuple(bufs[n].shape)))
                    bufs[n].copy_(src)
        self._cells[key] = ent                 # (re-)insert as most recently used
        return ent[0]

    def grads(self, key, batch, t_control, t_memory, global_batch):
        """Forward + backward of the local shard into the flat bucket (not yet reduced)."""
        from .autograd import mac_backward
        from .mac_cell import mac_network
        cell = self.cell_for(key, batch)
        cell._rw.clear()
        # per-(step, rank) dropout stream: masks differ across ranks and steps, reproducibly
        cell.seed = (self.base_seed * 1000003 + self.step_id * 7919 + self.rank * 104729 + 1) & 0x7FFFFFFFFFFFFFFF
        control, memory = mac_network(cell, self.L)
        scale = 1.0 / float(global_batch)
        g = mac_backward(cell, None if t_control is None else t_control * scale,
                         None if t_memory is None else t_memory * scale, bucket=self.bucket, tc=self.bwd_tc)
        return control, memory, g

    def apply(self):
        """all-reduce the bucket, then clip + Adam + EMA (model.py:645-667) in one fused pass; refresh derived weights."""
        allreduce_sum_(self.bucket)
        self.step_id += 1
        h = self.hp
        check(self.lib.mac_clip_adam_ema_step(ptr(self.params.flat), ptr(self.bucket), ptr(self.adam_m), ptr(self.adam_v),
                                              ptr(self.ema), self.params.numel, 1.0, h["clip"], h["lr"], h["b1"], h["b2"],
                                              h["eps"], self.step_id, h["ema"], ptr(self.norm), ptr(self.ows),
                                              self.ows_bytes, stream_ptr()), "mac_clip_adam_ema_step")
        self.params.touch()

    def train_step_answers(self, key, batch, answers, global_batch):
        """One DP step on the reference's loss: cell forward -> output unit -> mean softmax-CE over the GLOBAL batch ->
        output-unit backward -> cell backward -> all-reduce -> clip/Adam/EMA.  Returns (logits, per-sample losses)."""
        from .autograd import mac_backward
        from .mac_cell import mac_network
        cell = self.cell_for(key, batch)
        cell._rw.clear()
        cell.seed = (self.base_seed * 1000003 + self.step_id * 7919 + self.rank * 104729 + 1) & 0x7FFFFFFFFFFFFFFF
        self.out.seed = cell.seed
        control, memory = mac_network(cell, self.L)
        self.bucket.zero_()
        gviews = self._views_of(self.bucket, self.params.specs, self.params.offsets)
        logits, losses, _ = self.out.forward(memory, getattr(cell, "vecQuestions", batch["vecQuestions"]), answers, step=self.step_id,
                                             loss_scale=1.0 / float(global_batch))
        d_mem, d_q = torch.zeros_like(memory), torch.zeros_like(memory)
        self.out.backward(gviews, d_mem, d_q)
        mac_backward(cell, None, d_mem, bucket=self.bucket, zero_bucket=False, d_vecq=d_q, tc=self.bwd_tc)
        self.apply()
        self.out.invalidate()
        return logits, losses

    def full_forward_backward(self, key, data, global_batch):
        """The reference's whole training graph (`MACnet.build`, model.py:774-821) on the local shard, gradients into the
        flat bucket (not yet reduced):  embeddings + bi-LSTM encoder -> stem -> netLength MAC steps -> output unit ->
        classifier -> mean softmax-CE over the GLOBAL batch, then the hand-written backward of each in reverse order.
        `data`: questions int32 [B,S] (0 = padding), questionLengths int32 [B], images fp32 [B,H,W,C] (NHWC: the
        reference transposes its NCHW feed first, model.py:68), answers int32 [B].  Returns (logits, per-sample losses)."""
        from .autograd import mac_backward
        from .mac_cell import mac_network
        if self.enc is None:
            raise RuntimeError("construct the trainer with classifier=, encoder= and stem=")
        seed = (self.base_seed * 1000003 + self.step_id * 7919 + self.rank * 104729 + 1) & 0x7FFFFFFFFFFFFFFF
        self.enc.seed = self.stem.seed = self.out.seed = seed
        words, cntx, vecq = self.enc.forward(data["questions"], data["questionLengths"], step=self.step_id,
                                             save_for_backward=True)
        kb = self.stem.forward(data["images"], keep=self.stem_dropout, step=self.step_id, save_for_backward=True)
        # the cell captures its inputs at construction (mac_cell.py:59-79): cell_for owns persistent buffers per key and
        # copies this step's encoder / stem outputs into them
        bufs = {"vecQuestions": vecq, "questionWords": words, "questionCntxWords": cntx, "knowledgeBase": kb,
                "questionLengths": data["questionLengths"]}
        cell = self.cell_for(key, bufs)
        cell._rw.clear()
        cell.seed = seed
        control, memory = mac_network(cell, self.L)
        self.bucket.zero_()
        gviews = self._views_of(self.bucket, self.params.specs, self.params.offsets)
        logits, losses, _ = self.out.forward(memory, getattr(cell, "vecQuestions", vecq), data["answers"], step=self.step_id,
                                             loss_scale=1.0 / float(global_batch))
        d_mem, d_q = torch.zeros_like(memory), torch.zeros_like(memory)
        self.out.backward(gviews, d_mem, d_q)
        g = mac_backward(cell, None, d_mem, bucket=self.bucket, zero_bucket=False, d_vecq=d_q, tc=self.bwd_tc)
        self.stem.backward(g["knowledgeBase"], gviews)
        if not self.cfg.controlContextual:
            raise NotImplementedError("the raw-word control inputs (controlContextual off) need wrdEmbDim == ctrlDim")
        self.enc.backward(g["questionCntxWords"], g["vecQuestions"], gviews)
        return logits, losses

    def train_step_full(self, key, data, global_batch):
        """One data-parallel step of the whole model: `full_forward_backward` -> all-reduce -> clip / Adam / EMA."""
        logits, losses = self.full_forward_backward(key, data, global_batch)
        self.apply()
        self.out.invalidate()
        self.stem._packed.clear()
        return logits, losses

    def train_step(self, key, batch, t_control, t_memory, global_batch):
        control, memory, _ = self.grads(key, batch, t_control, t_memory, global_batch)
        self.apply()
        return control, memory


def adam_reference(p, g, m, v, ema, step, lr=1e-4, clip=8.0, b1=0.9, b2=0.999, eps=1e-8, ema_decay=0.999):
    """numpy restatement of the fused step (tests): tf.clip_by_global_norm + tf.train.AdamOptimizer + EMA.apply."""
    p, g, m, v, ema = (np.asarray(a, np.float64) for a in (p, g, m, v, ema))
    norm = np.sqrt(np.sum(g * g))
    g = g * (clip / max(norm, clip))
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    p = p - lr_t * m / (np.sqrt(v) + eps)
    ema = ema_decay * ema + (1 - ema_decay) * p
    return p, m, v, ema, norm
