"""CPU ORACLE (test infrastructure only) -- numpy restatement of the reference MAC cell.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this file.  Nothing under `mac_network_b200/` does: the product path is the CUDA library and
fails loudly without it.

What it restates (all citations into /root/reference):
  * `mac_cell.py:133-187`  control unit        -> `MACOracle.control`
  * `mac_cell.py:209-277`  read unit           -> `MACOracle.read`
  * `mac_cell.py:305-375`  write unit          -> `MACOracle.write`
  * `mac_cell.py:420-480`  one reasoning step  -> `MACOracle.step`
  * `mac_cell.py:496-505, 539-592` state init  -> `MACOracle.zero_state`
  * `model.py:447-458`     netLength unroll    -> `MACOracle.run`
  * `ops.py:50-59, 65-78, 114-150, 161-187, 243-247, 298-333, 668-725, 1054-1067` -> the helpers below.

PARITY PINNING.  TensorFlow cannot be installed here, and the reference ships no tests or golden
vectors.  The restatement is instead pinned against the reference's *own Python code* executed on the
numpy TF1-API shim (`oracle/tf1_shim`, driven by `oracle/gen_golden.py`): the committed fixtures in
`tests/golden/*.npz` are outputs of the unmodified `/root/reference/mac_cell.py` + `ops.py`.  That
pins graph structure, variable naming and op order; TF's kernels themselves (matmul, softmax, elu,
dropout) are restated from their published definitions and are "parity unpinned" in that sense.

Arithmetic dtype is a constructor argument: float64 = arbiter, float32 = TF-like numerics.
"""
import collections
import numpy as np

MACCellTuple = collections.namedtuple("MACCellTuple", ("control", "memory"))
INF = 1e30          # ops.py:10


# ------------------------------------------------------------------ ops.py helpers
def softmax(x):
    m = np.max(x, axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=-1, keepdims=True)


def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def exp_mask(seq, lengths):
    """ops.py:243-247: seq + (1 - sequence_mask) * (-1e30)."""
    S = seq.shape[-1]
    valid = (np.arange(S)[None, :] < np.asarray(lengths)[:, None]).astype(seq.dtype)
    return seq + (1 - valid) * (-INF)


def att2smry(att, feats):
    """ops.py:149-150."""
    return np.sum(att[..., None] * feats, axis=-2)


class MACOracle(object):
    """Stateful like the reference cell (mac_cell.py:32-34): zero_state resets, step mutates."""

    def __init__(self, cfg, params, dtype=np.float64, prefix="MACnetwork/"):
        self.cfg = cfg
        self.dtype = dtype
        self.prefix = prefix
        self.p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
        self.uniforms = None      # iterator over uniform draws, in the reference's call order
        self.train = False        # the cell's `train` argument (mac_cell.py:64): only memoryBN reads it here

    # -------------------------------------------------------------- variable access
    def var(self, scope, name):
        return self.p[self.prefix + scope + name]

    def act(self, kind, x):
        """ops.py:161-187 (`activations` dict with the `config.relu` switch)."""
        if kind == "NON":
            return x
        if kind == "TANH":
            return np.tanh(x)
        if kind == "SIGMOID":
            return 1.0 / (1.0 + np.exp(-x))
        if kind == "ELU":
            return elu(x)
        if kind == "RELU":
            return elu(x) if self.cfg.relu == "ELU" else np.maximum(x, 0)
        raise ValueError(kind)

    def dropout(self, x, keep):
        """tf.nn.dropout (TF1): x/keep*floor(keep+U); exact identity at keep == 1 (no draw)."""
        if float(keep) == 1.0:
            return x
        u = np.asarray(next(self.uniforms), dtype=self.dtype)
        assert u.shape == x.shape, (u.shape, x.shape)
        return x / self.dtype(keep) * np.floor(self.dtype(keep) + u)

    def linear(self, x, scope, name, in_dim, out_dim, act="NON", dropout=1.0, bias=0.0):
        """ops.py:298-333: act(drop(x) @ W + b) and, when act != NON, the nested `name_2` layer
        (ops.py:325-328).  out_dim == 1 -> vector weight, scalar bias, row-dot (ops.py:304-305, 316-317)."""
        sc = scope + "linearLayer" + name + "/"
        W = self.var(sc, "weights/weight")
        b = self.var(sc, "biases/bias") + self.dtype(bias)
        x = self.dropout(x, dropout)
        if out_dim > 1:
            assert W.shape == (in_dim, out_dim), (sc, W.shape, in_dim, out_dim)
            y = np.matmul(x, W) + b
        else:
            assert W.shape == (in_dim,), (sc, W.shape, in_dim)
            y = np.sum(x * W, axis=-1) + b
        y = self.act(act, y)
        if act != "NON":
            y = self.linear(y, sc, name + "_2", out_dim, out_dim)
        return y

    def inter2att(self, inter, scope, dim, dropout=1.0, name=""):
        """ops.py:114-120, 140-144 (sumMod = LIN)."""
        logits = self.linear(inter, scope + "inter2att" + name + "/inter2logits/", "logits", dim, 1,
                             dropout=dropout)
        return softmax(logits)

    def mul(self, x, y, dim, scope, name, proj=None, inter_mod="MUL", concat=None):
        """ops.py:668-725.  `proj` = dict(dim, shared, dropout) or None; returns (out, outDim, projectedX)."""
        sc = scope + "mul" + name + "/"
        orig_x, orig_dim = x, dim
        proj_x = None
        if proj is not None:
            x = self.dropout(x, proj["dropout"])       # ops.py:678-679 (the dropout= arg is 1.0 at both call sites)
            y = self.dropout(y, proj["dropout"])
            xn, yn = ("proj", "proj") if proj["shared"] else ("projX", "projY")
            x = self.linear(x, sc, xn, dim, proj["dim"])
            y = self.linear(y, sc, yn, dim, proj["dim"])
            dim = proj["dim"]
            proj_x = x
        yb = y[..., None, :]                            # ops.py:694-697 (broadcast over the KB axis)
        if inter_mod == "MUL":
            mb = self.dtype(self.cfg.mulBias)
            out = (x + mb) * (yb + mb)
        elif inter_mod == "BL":
            out = np.matmul(x, self.var(sc, "weights/weight")) * yb + self.var(sc, "biases/bias")
        elif inter_mod == "ADD":
            out = np.tanh(x + yb)
        else:
            raise NotImplementedError(inter_mod)
        if concat and concat.get("x"):
            use_proj = concat.get("proj", False)
            cx, cd = (proj_x, dim) if use_proj else (orig_x, orig_dim)
            out = np.concatenate([out, cx], axis=-1)
            dim += cd
        return out, dim, proj_x

    # -------------------------------------------------------------- state init
    def init_state(self, name, dim, init_type, B):
        """mac_cell.py:496-505."""
        if init_type == "PRM":
            return np.tile(self.var("", name)[None, :], (B, 1))
        if init_type == "ZERO":
            return np.zeros((B, dim), dtype=self.dtype)
        return self.vecQuestions

    def zero_state(self, vecQuestions, questionWords, questionCntxWords, questionLengths, knowledgeBase,
                   memoryDropout=1.0, readDropout=1.0, writeDropout=1.0, uniforms=None):
        """mac_cell.py:59-79 (capture inputs) + 539-592."""
        c, t = self.cfg, self.dtype
        self.vecQuestions = np.asarray(vecQuestions, t)
        self.knowledgeBase = np.asarray(knowledgeBase, t)
        self.questionLengths = np.asarray(questionLengths)
        self.dropouts = {"memory": memoryDropout, "read": readDropout, "write": writeDropout}
        self.uniforms = iter(uniforms) if uniforms is not None else iter(())
        B = self.vecQuestions.shape[0]
        self.attentions = {"kb": [], "question": [], "self": [], "gate": []}
        c0 = self.init_state("initCtrl", c.ctrlDim, c.initCtrl, B)
        m0 = self.init_state("initMem", c.memDim, c.initMem, B)
        self.controls = c0[:, None, :]
        self.memories = m0[:, None, :]
        self.infos = m0[:, None, :]
        self.contControl = c0
        words = np.asarray(questionCntxWords if c.controlContextual else questionWords, t)
        self.inWords = self.outWords = words
        if c.controlInWordsProj or c.controlOutWordsProj:
            pw = self.linear(words, "", "wordsProj", c.ctrlDim, c.ctrlDim)
            self.inWords = pw if c.controlInWordsProj else words
            self.outWords = pw if c.controlOutWordsProj else words
        if c.memoryVariationalDropout:
            # ops.py:1054-1059: floor(keep + U) -- one [B,memDim] mask per forward.  The reference always
            # draws (keepProb is a tensor); at keep == 1 the mask is all-ones, so skip the draw then.
            keep = float(memoryDropout)
            if keep == 1.0:
                self.memDpMask = np.ones((B, c.memDim), dtype=t)
            else:
                u = np.asarray(next(self.uniforms), dtype=t)
                self.memDpMask = np.floor(t(keep) + u)
        self.iteration = 0
        self.trace = []
        return MACCellTuple(c0, m0)

    # -------------------------------------------------------------- units
    def control(self, controlInput, inWords, outWords, lengths, control, contControl, name=""):
        """mac_cell.py:133-187."""
        c = self.cfg
        sc = "MACCell/control" + name + "/"
        dim = c.ctrlDim
        new_cont = controlInput
        if c.controlFeedPrev:
            new_cont = control if c.controlFeedPrevAtt else contControl
            if c.controlFeedInputs:
                new_cont = np.concatenate([new_cont, controlInput], axis=-1)
                dim += c.ctrlDim
            new_cont = self.linear(new_cont, sc, "contControl", dim, c.ctrlDim, act=c.controlContAct)
            dim = c.ctrlDim
        inter = new_cont[:, None, :] * inWords
        if c.controlConcatWords:
            inter = np.concatenate([inter, inWords], axis=-1)
            dim += c.ctrlDim
        if c.controlProj:
            inter = self.linear(inter, sc, "", dim, c.ctrlDim, act=c.controlProjAct)
            dim = c.ctrlDim
        logits = self.linear(inter, sc + "inter2logits/", "logits", dim, 1)
        att = softmax(exp_mask(logits, lengths))
        self.attentions["question"].append(att)
        new_control = att2smry(att, outWords)
        if c.controlContinuous:
            new_control = new_cont
        return new_control, new_cont

    def read(self, knowledgeBase, memory, control, name=""):
        """mac_cell.py:209-277."""
        c, t = self.cfg, self.dtype
        sc = "MACCell/read" + name + "/"
        dim = c.memDim
        if c.memoryVariationalDropout:
            memory = memory / t(self.dropouts["memory"]) * self.memDpMask          # ops.py:1065-1067
        else:
            memory = self.dropout(memory, self.dropouts["memory"])
        proj = None
        if c.readProjInputs:
            proj = {"dim": c.attDim, "shared": c.readProjShared, "dropout": self.dropouts["read"]}
            dim = c.attDim
        inter, inter_dim, projectedKB = self.mul(
            knowledgeBase, memory, c.memDim, sc, "memInter", proj=proj, inter_mod=c.readMemAttType,
            concat={"x": c.readMemConcatKB, "proj": c.readMemConcatProj})
        if c.readMemProj:
            inter = self.linear(inter, sc, "memKbProj", inter_dim, dim, act=c.readMemAct)
        else:
            dim = inter_dim
        if c.readCtrl:
            inter, _, _ = self.mul(inter, control, dim, sc, "ctrlInter", inter_mod=c.readCtrlAttType,
                                   concat={"x": False})
            if c.readCtrlConcatKB:
                if c.readCtrlConcatProj:
                    added, added_dim = projectedKB, c.attDim
                else:
                    added, added_dim = knowledgeBase, c.memDim
                inter = np.concatenate([inter, added], axis=-1)
                dim += added_dim
            inter = self.act(c.readCtrlAct, inter)
        att = self.inter2att(inter, sc, dim, dropout=self.dropouts["read"])
        self.attentions["kb"].append(att)
        if c.readSmryKBProj:
            knowledgeBase = projectedKB
        return att2smry(att, knowledgeBase)

    def write(self, memory, info, control, contControl, name=""):
        """mac_cell.py:305-375."""
        c, t = self.cfg, self.dtype
        sc = "MACCell/write" + name + "/"
        if c.writeInfoProj:
            info = self.linear(info, sc, "info", c.memDim, c.memDim)
        info = self.act(c.writeInfoAct, info)
        if c.writeSelfAtt:
            self_control = contControl if c.writeSelfAttMod == "CONT" else control
            self_control = self.linear(self_control, sc, "ctrlProj", c.ctrlDim, c.ctrlDim)
            inter = self.controls * self_control[:, None, :]
            att = self.inter2att(inter, sc, c.ctrlDim, name="selfAttention")
            self.attentions["self"].append(att)
            self_smry = att2smry(att, self.memories)
        new_mem, dim = memory, c.memDim
        if c.writeInputs == "INFO":
            new_mem = info
        elif c.writeInputs == "SUM":
            new_mem = new_mem + info
        elif c.writeInputs == "BOTH":
            parts = [new_mem, info] + ([new_mem * info] if c.writeConcatMul else [])   # ops.py:65-78
            new_mem = np.concatenate(parts, axis=-1)
            dim = dim * len(parts)
        if c.writeSelfAtt:
            new_mem = np.concatenate([new_mem, self_smry], axis=-1)
            dim += c.memDim
        if c.writeMergeCtrl:
            new_mem = np.concatenate([new_mem, control], axis=-1)
            dim += c.memDim
        if c.writeMemProj or dim != c.memDim:
            new_mem = self.linear(new_mem, sc, "newMemory", dim, c.memDim)
        new_mem = self.act(c.writeMemAct, new_mem)
        if c.writeGate:
            z = self.linear(control, sc, "gate", c.ctrlDim, c.memDim, bias=c.writeGateBias)
            z = 1.0 / (1.0 + np.exp(-z))
            self.attentions["gate"].append(z)
            new_mem = new_mem * z + memory * (1 - z)
        if c.memoryBN:
            new_mem = self.batch_norm(new_mem, sc + "BatchNorm/")
        return new_mem

    def batch_norm(self, x, scope):
        """mac_cell.py:370-373: tf.contrib.layers.batch_norm(newMemory, decay=bnDecay, center=bnCenter, scale=bnScale,
        is_training=self.train, updates_collections=None), epsilon at its default 0.001.  TF 1.x semantics (fused path for
        a rank-2 input): training normalises with the batch mean and biased variance and moves the stored statistics by
        (1 - decay) towards the batch mean / the Bessel-corrected batch variance, once per call (i.e. per reasoning step);
        inference normalises with the stored statistics."""
        c, eps = self.cfg, 1e-3
        P = self.prefix + scope
        beta = self.p[P + "beta"] if c.bnCenter else 0.0
        gamma = self.p[P + "gamma"] if c.bnScale else 1.0
        if self.train:
            n = x.shape[0]
            mean = x.mean(axis=0)
            var = ((x - mean) ** 2).mean(axis=0)
            mm, mv = self.p[P + "moving_mean"], self.p[P + "moving_variance"]
            self.p[P + "moving_mean"] = mm - (mm - mean) * (1.0 - c.bnDecay)
            self.p[P + "moving_variance"] = mv - (mv - var * (float(n) / max(n - 1, 1))) * (1.0 - c.bnDecay)
        else:
            mean, var = self.p[P + "moving_mean"], self.p[P + "moving_variance"]
        return (x - mean) / np.sqrt(var + eps) * gamma + beta

    # -------------------------------------------------------------- one step / unroll
    def step(self, state):
        """mac_cell.py:420-480 with `self.iteration` set by the caller (model.py:454)."""
        c = self.cfg
        i = self.iteration
        control, memory = state
        in_name_u = ("qInput%d" % i) if c.controlInputUnshared else "qInputU"
        cell_name = str(i) if c.unsharedCells else ""
        ci = self.linear(self.vecQuestions, "MACCell/", "qInput", c.ctrlDim, c.ctrlDim)
        ci = self.act(c.controlInputAct, ci)
        ci = self.linear(ci, "MACCell/", in_name_u, c.ctrlDim, c.ctrlDim)
        new_control, self.contControl = self.control(ci, self.inWords, self.outWords, self.questionLengths,
                                                     control, self.contControl, name=cell_name)
        if c.controlWholeQ:
            new_control = self.vecQuestions
        info = self.read(self.knowledgeBase, memory, new_control, name=cell_name)
        if c.writeDropout < 1.0:          # python-level test on the *config* value (mac_cell.py:461)
            info = self.dropout(info, self.dropouts["write"])
        new_memory = self.write(memory, info, new_control, self.contControl, name=cell_name)
        self.controls = np.concatenate([self.controls, new_control[:, None, :]], axis=1)
        self.memories = np.concatenate([self.memories, new_memory[:, None, :]], axis=1)
        self.infos = np.concatenate([self.infos, info[:, None, :]], axis=1)
        self.trace.append({"control": new_control, "memory": new_memory, "info": info,
                           "contControl": self.contControl})
        return MACCellTuple(new_control, new_memory)

    def run(self, netLength, *inputs, **kw):
        """model.py:447-458: zero_state then the static netLength unroll."""
        state = self.zero_state(*inputs, **kw)
        for i in range(netLength):
            self.iteration = i
            state = self.step(state)
        return state

    def outputs(self):
        """Per-step tensors in the layout the golden fixtures use."""
        out = {k: np.stack([tr[k] for tr in self.trace]) for k in ("control", "memory", "info", "contControl")}
        for k, v in self.attentions.items():
            if v and k != "self":
                out["att_" + k] = np.stack(v)
            elif v:
                for i, a in enumerate(v):
                    out["att_self_%d" % i] = a
        return out
