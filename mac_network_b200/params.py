"""Parameters of the MAC cell, keyed by the reference's TensorFlow variable names.

The reference creates variables lazily through `tf.get_variable` inside nested `tf.variable_scope`s
(`ops.py:18-42, 298-333`; `mac_cell.py:136, 210, 306, 422, 498`).  The resulting names are the
checkpoint contract, so this module enumerates them explicitly for a given flag
set; `tests/test_oracle_golden.py` checks the enumeration against the names the unmodified reference
creates when run on the TF1 shim.

Initialisers: Xavier-uniform for every `weights/weight` including the 1-D logit vectors
(`ops.py:18-24`, fan_in = fan_out = n for shape [n]); zeros for biases (`ops.py:38-42`); N(0,1) for
`initMem` / `initCtrl` (`mac_cell.py:498-499`).
"""
import collections
import numpy as np

PREFIX = "MACnetwork/"          # model.py:431 (inside "macModel", model.py:774, which checkpoints prepend)


def _linear(specs, scope, name, in_dim, out_dim, act="NON"):
    sc = scope + "linearLayer" + name + "/"
    if out_dim > 1:
        specs[sc + "weights/weight"] = ((in_dim, out_dim), "xavier")
        specs[sc + "biases/bias"] = ((out_dim,), "zeros")
    else:
        specs[sc + "weights/weight"] = ((in_dim,), "xavier")
        specs[sc + "biases/bias"] = ((), "zeros")
    if act != "NON":                      # ops.py:325-328: nested "<name>_2" layer
        _linear(specs, sc, name + "_2", out_dim, out_dim)


def param_specs(cfg, netLength=None):
    """OrderedDict: full variable name -> (shape, initialiser kind)."""
    c = cfg
    L = c.netLength if netLength is None else netLength
    d, a = c.ctrlDim, c.attDim
    s = collections.OrderedDict()
    P = PREFIX
    if c.initCtrl == "PRM":
        s[P + "initCtrl"] = ((c.ctrlDim,), "normal")
    if c.initMem == "PRM":
        s[P + "initMem"] = ((c.memDim,), "normal")
    if c.controlInWordsProj or c.controlOutWordsProj:
        _linear(s, P, "wordsProj", d, d)                                   # mac_cell.py:578-581
    cell = P + "MACCell/"
    _linear(s, cell, "qInput", d, d)                                       # mac_cell.py:442-443
    if c.controlInputUnshared:
        for i in range(L):
            _linear(s, cell, "qInput%d" % i, d, d)                         # mac_cell.py:430-432, 447-448
    else:
        _linear(s, cell, "qInputU", d, d)
    for cell_name in ([str(i) for i in range(L)] if c.unsharedCells else [""]):
        # ---- control (mac_cell.py:133-187)
        sc = cell + "control" + cell_name + "/"
        dim = d
        if c.controlFeedPrev:
            if c.controlFeedInputs:
                dim += d
            _linear(s, sc, "contControl", dim, d, act=c.controlContAct)
            dim = d
        if c.controlConcatWords:
            dim += d
        if c.controlProj:
            _linear(s, sc, "", dim, d, act=c.controlProjAct)
            dim = d
        _linear(s, sc + "inter2logits/", "logits", dim, 1)
        # ---- read (mac_cell.py:209-277, ops.py:668-725)
        sc = cell + "read" + cell_name + "/"
        dim, inter_dim = c.memDim, c.memDim
        mul = sc + "mulmemInter/"
        if c.readProjInputs:
            if c.readProjShared:
                _linear(s, mul, "proj", c.memDim, a)
            else:
                _linear(s, mul, "projX", c.memDim, a)
                _linear(s, mul, "projY", c.memDim, a)
            dim = inter_dim = a
        if c.readMemAttType == "BL":
            s[mul + "weights/weight"] = ((inter_dim, inter_dim), "xavier")
            s[mul + "biases/bias"] = ((inter_dim,), "zeros")
        if c.readMemConcatKB:
            inter_dim += a if c.readMemConcatProj else c.memDim
        if c.readMemProj:
            _linear(s, sc, "memKbProj", inter_dim, dim, act=c.readMemAct)
        else:
            dim = inter_dim
        if c.readCtrl:
            if c.readCtrlAttType == "BL":
                s[sc + "mulctrlInter/weights/weight"] = ((dim, dim), "xavier")
                s[sc + "mulctrlInter/biases/bias"] = ((dim,), "zeros")
            if c.readCtrlConcatKB:
                dim += a if c.readCtrlConcatProj else c.memDim
        _linear(s, sc + "inter2att/inter2logits/", "logits", dim, 1)
        # ---- write (mac_cell.py:305-375)
        sc = cell + "write" + cell_name + "/"
        if c.writeInfoProj:
            _linear(s, sc, "info", c.memDim, c.memDim)
        if c.writeSelfAtt:
            _linear(s, sc, "ctrlProj", d, d)
            _linear(s, sc + "inter2attselfAttention/inter2logits/", "logits", d, 1)
        dim = c.memDim
        if c.writeInputs == "BOTH":
            dim *= 3 if c.writeConcatMul else 2
        if c.writeSelfAtt:
            dim += c.memDim
        if c.writeMergeCtrl:
            dim += c.memDim
        if c.writeMemProj or dim != c.memDim:
            _linear(s, sc, "newMemory", dim, c.memDim)
        if c.writeGate:
            _linear(s, sc, "gate", d, c.memDim)
        if c.memoryBN:
            # tf.contrib.layers.batch_norm under the write scope (mac_cell.py:370-373): beta / gamma only with
            # bnCenter / bnScale; the moving statistics are variables too (non-trainable: their gradient stays zero)
            bn = sc + "BatchNorm/"
            if c.bnCenter:
                s[bn + "beta"] = ((c.memDim,), "zeros")
            if c.bnScale:
                s[bn + "gamma"] = ((c.memDim,), "ones")
            s[bn + "moving_mean"] = ((c.memDim,), "zeros")
            s[bn + "moving_variance"] = ((c.memDim,), "ones")
    return s


def init_params(cfg, netLength=None, seed=0, dtype=np.float32):
    """Deterministic initial values (numpy legacy RandomState: stream frozen across numpy versions)."""
    rng = np.random.RandomState(seed)
    out = collections.OrderedDict()
    for name, (shape, kind) in param_specs(cfg, netLength).items():
        if kind == "zeros":
            v = np.zeros(shape)
        elif kind == "ones":
            v = np.ones(shape)
        elif kind == "normal":
            v = rng.standard_normal(shape)
        else:
            if len(shape) == 1:
                fan_in = fan_out = shape[0]
            else:
                fan_in, fan_out = shape
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            v = rng.uniform(-lim, lim, size=shape)
        out[name] = np.asarray(v, dtype=dtype)
    return out


def perturb_biases(params, seed=1, scale=0.1):
    """Tests/benchmarks: make the (zero-initialised) biases non-trivial so bias handling is exercised."""
    rng = np.random.RandomState(seed)
    out = collections.OrderedDict()
    for k, v in params.items():
        if k.endswith("biases/bias"):
            out[k] = np.asarray(scale * rng.standard_normal(v.shape), dtype=v.dtype)
        else:
            out[k] = v
    # batch-norm variables (memoryBN): their own stream, so that flag sets without them keep the values they always had
    rng_bn = np.random.RandomState(seed + 7919)
    for k, v in params.items():
        tail = k.rsplit("/", 1)[-1]
        if "/BatchNorm/" not in k:
            continue
        n = rng_bn.standard_normal(v.shape)
        if tail in ("beta", "moving_mean"):
            out[k] = np.asarray(v + 2 * scale * n, dtype=v.dtype)
        elif tail == "gamma":
            out[k] = np.asarray(v * (1.0 + 2 * scale * n), dtype=v.dtype)
        else:                                                    # moving_variance: positive
            out[k] = np.asarray(v * (0.5 + np.abs(n)), dtype=v.dtype)
    return out
