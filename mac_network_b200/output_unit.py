"""Output unit, classifier and answer loss of the reference model on the same CUDA primitives as the cell
("next" row 2): `MACnet.outputOp` (`model.py:512-528`, `outQuestion` on), `MACnet.classifier`
(`model.py:547-576` -> `ops.FCLayer`, `ops.py:349-359`), `addAnswerLossOp` (`model.py:593-596`).

    features = [memory, vecQuestions @ W_oq + b_oq]                       (2 * memDim)
    h        = act(dropout(features) @ W_fc0 + b_fc0) ...                 (outClassifierDims, act = RELU -> config.relu)
    logits   = dropout(h) @ W_fcK + b_fcK                                 (answerWordsNum)
    loss     = mean_b( logsumexp(logits_b) - logits_b[answer_b] )

It supplies dL/dmemory and dL/dvecQuestions to the cell's backward, so data-parallel training runs on the reference's
real loss.  Variable names follow the reference's scopes (siblings of "MACnetwork/" under "macModel/")."""
import collections
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import ACT, check, ptr, stream_ptr

SITE_OUTPUT = 16          # Philox site base for the output unit's dropouts (site + layer index)


def output_specs(ctrl_dim, mem_dim, hidden, n_answers):
    s = collections.OrderedDict()
    s["outputUnit/linearLayeroutQuestion/weights/weight"] = ((ctrl_dim, mem_dim), "xavier")
    s["outputUnit/linearLayeroutQuestion/biases/bias"] = ((mem_dim,), "zeros")
    dims = [2 * mem_dim] + list(hidden) + [n_answers]
    for i in range(len(dims) - 1):
        s["classifier/linearLayerfc_%d/weights/weight" % i] = ((dims[i], dims[i + 1]), "xavier")
        s["classifier/linearLayerfc_%d/biases/bias" % i] = ((dims[i + 1],), "zeros")
    return s


def init_output_params(specs, seed=0, dtype=np.float32, bias_scale=0.1):
    rng = np.random.RandomState(seed)
    out = collections.OrderedDict()
    for name, (shape, kind) in specs.items():
        if kind == "zeros":
            v = bias_scale * rng.standard_normal(shape)       # non-trivial biases so bias handling is exercised
        else:
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            v = rng.uniform(-lim, lim, size=shape)
        out[name] = np.asarray(v, dtype=dtype)
    return out


class OutputUnit(object):
    """Forward / loss / backward of the output unit on device tensors.  `params` / `grads`: dict name -> tensor
    (e.g. views into the trainer's flat buckets)."""

    def __init__(self, params, relu="ELU", keep=1.0, seed=0, version=None):
        """`version`: optional callable returning a counter that changes whenever the parameter values do
        (`MACParams.version`): the transposed weight copies of the backward are rebuilt when it moves."""
        self.lib = _lib.load()
        self.p = params
        self._version_fn, self._wt_version = version, None
        self.relu, self.keep, self.seed = relu, float(keep), int(seed)
        self.nfc = len([k for k in params if k.startswith("classifier/linearLayerfc_") and k.endswith("weights/weight")])
        for k, v in params.items():
            if k.endswith("weights/weight") and (v.shape[0] % 4 or v.shape[1] % 4):
                raise ValueError("%s is %s: the fp32 GEMM needs every dimension to be a multiple of 4 (pad the answer "
                                 "vocabulary / classifier width)" % (k, tuple(v.shape)))
        dev = next(iter(params.values())).device
        self.lws_bytes = 4096 + 32 * 64 * 2048 * 4
        self.lws = torch.zeros(self.lws_bytes, dtype=torch.uint8, device=dev)
        self._wt = {}

    def _new(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.lws.device)

    def _linear(self, xs, W, b, act=0):
        n, M = len(xs), xs[0].shape[0]
        y = self._new(M, W.shape[1])
        arr_p = (ctypes.c_void_p * n)(*[x.data_ptr() for x in xs])
        arr_k = (ctypes.c_int * n)(*[x.shape[1] for x in xs])
        arr_ld = (ctypes.c_int * n)(*[x.stride(0) for x in xs])
        check(self.lib.mac_linear_fwd(arr_p, arr_k, arr_ld, n, ptr(W), ptr(b), 0.0, act, ptr(y), y.stride(0), M, W.shape[1],
                                      ptr(self.lws), self.lws_bytes, stream_ptr()), "mac_linear_fwd")
        return y

    def _dropout(self, x, layer, step):
        if self.keep >= 1.0:
            return x
        out = self._new(*x.shape)
        check(self.lib.mac_dropout_fwd(ptr(x), self.keep, self.seed, SITE_OUTPUT + layer, step, ptr(out), x.numel(),
                                       stream_ptr()), "mac_dropout_fwd")
        return out

    def forward(self, memory, vecQuestions, answers, step=0, loss_scale=None):
        """Returns (logits, losses [B], dlogits [B, A]) -- dlogits = (softmax - onehot) * loss_scale (default 1/B)."""
        B = memory.shape[0]
        act = ACT["ELU"] if self.relu == "ELU" else ACT["RELU_STD"]
        self.eq = self._linear([vecQuestions], self.p["outputUnit/linearLayeroutQuestion/weights/weight"],
                               self.p["outputUnit/linearLayeroutQuestion/biases/bias"])
        self.memory, self.vecq, self.step = memory, vecQuestions, step
        self.inputs = []            # per layer: list of input segments after dropout
        xs = [memory, self.eq]
        x = None
        for i in range(self.nfc):
            W = self.p["classifier/linearLayerfc_%d/weights/weight" % i]
            b = self.p["classifier/linearLayerfc_%d/biases/bias" % i]
            if i == 0:
                # dropout over the concatenated features: one Philox stream over [B, 2*memDim], applied per segment
                if self.keep < 1.0:
                    cat = torch.cat(xs, dim=1)                                # plumbing: layout for the flat mask index
                    xs = [self._dropout(cat, 0, step)]
            else:
                xs = [self._dropout(x, i, step)]
            self.inputs.append(xs)
            x = self._linear(xs, W, b, act if i < self.nfc - 1 else 0)
            if i < self.nfc - 1:
                setattr(self, "_h%d" % i, x)
        self.logits = x
        A = x.shape[1]
        self.losses = self._new(B)
        self.dlogits = self._new(B, A)
        scale = (1.0 / B) if loss_scale is None else float(loss_scale)
        check(self.lib.mac_softmax_xent(ptr(self.logits), ptr(answers), ptr(self.losses), ptr(self.dlogits), scale, B, A,
                                        stream_ptr()), "mac_softmax_xent")
        return self.logits, self.losses, self.dlogits

    def invalidate(self):
        """Call after the parameters were updated in place (optimizer step): drops the cached transposes."""
        self._wt.clear()

    def _wt_of(self, name):
        v = self._version_fn() if self._version_fn is not None else None
        if v != self._wt_version:
            self._wt.clear()
            self._wt_version = v
        if name not in self._wt:
            self._wt[name] = self.p[name].t().contiguous()
        return self._wt[name]

    def backward(self, grads, d_memory, d_vecq):
        """Accumulates parameter gradients into `grads` (dict name -> tensor) and ADDS dL/dmemory, dL/dvecQuestions."""
        B = self.memory.shape[0]
        dy = self.dlogits
        act = ACT["ELU"] if self.relu == "ELU" else ACT["RELU_STD"]

        def lin_bwd(xs, wname, bname, dy, dxs, accum):
            n = len(xs)
            arr_x = (ctypes.c_void_p * n)(*[x.data_ptr() for x in xs])
            arr_k = (ctypes.c_int * n)(*[x.shape[1] for x in xs])
            arr_ld = (ctypes.c_int * n)(*[x.stride(0) for x in xs])
            arr_dx = (ctypes.c_void_p * n)(*[d.data_ptr() for d in dxs])
            arr_ldd = (ctypes.c_int * n)(*[d.stride(0) for d in dxs])
            arr_acc = (ctypes.c_int * n)(*accum)
            check(self.lib.mac_linear_bwd(arr_x, arr_k, arr_ld, n, ptr(self._wt_of(wname)), ptr(dy), dy.stride(0), arr_dx,
                                          arr_ldd, arr_acc, ptr(grads[wname]), ptr(grads[bname]), dy.shape[0], dy.shape[1],
                                          ptr(self.lws), self.lws_bytes, stream_ptr()), "mac_linear_bwd")
        for i in reversed(range(self.nfc)):
            xs = self.inputs[i]
            wn, bn = "classifier/linearLayerfc_%d/weights/weight" % i, "classifier/linearLayerfc_%d/biases/bias" % i
            if i == 0 and len(xs) == 2:
                dmem, deq = self._new(B, xs[0].shape[1]), self._new(B, xs[1].shape[1])
                lin_bwd(xs, wn, bn, dy, [dmem, deq], [0, 0])
            else:
                dx = self._new(B, xs[0].shape[1])
                lin_bwd(xs, wn, bn, dy, [dx], [0])
                if self.keep < 1.0:       # through the input dropout of this layer
                    check(self.lib.mac_dropout_fwd(ptr(dx), self.keep, self.seed, SITE_OUTPUT + i, self.step, ptr(dx),
                                                   dx.numel(), stream_ptr()), "dropout bwd")
                if i == 0:
                    md = self.memory.shape[1]
                    dmem, deq = dx[:, :md].contiguous(), dx[:, md:].contiguous()
                else:
                    h = getattr(self, "_h%d" % (i - 1))
                    dpre = self._new(*h.shape)
                    check(self.lib.mac_activation_bwd(ptr(h), ptr(dx), act, ptr(dpre), h.numel(), stream_ptr()), "act bwd")
                    dy = dpre
        check(self.lib.mac_axpy(ptr(d_memory), ptr(dmem), 1.0, dmem.numel(), stream_ptr()), "axpy")
        lin_bwd([self.vecq], "outputUnit/linearLayeroutQuestion/weights/weight",
                "outputUnit/linearLayeroutQuestion/biases/bias", deq, [d_vecq], [1])
