"""Eager numpy stand-in for the slice of the TensorFlow-1.x API that the
reference's hot path (`mac_cell.py`, `ops.py`, `mi_*_cell.py` import lines) touches.

TEST INFRASTRUCTURE ONLY.  TensorFlow is not installable in this image (no network,
no cp312 TF1 wheels), so `oracle/gen_golden.py` puts this directory on `sys.path`
*as* `tensorflow`, imports the UNMODIFIED reference modules from `/root/reference`
and runs them eagerly.  That pins the oracle (`oracle/mac_oracle.py`) against the
reference's own Python control flow: scope/variable naming, op order, concat order,
the nested "_2" layers, which tensors get dropout, what is appended to `attentions`.
What it cannot pin are TF's kernels themselves (matmul/softmax/elu/...): those are
restated here from their published definitions.

Everything is an `np.ndarray` in `WORK_DTYPE` (float64 for golden generation).
Nothing under `mac_network_b200/` may import this module.
"""
import collections
import contextlib
import numpy as np

WORK_DTYPE = np.float64

# dtype tokens the reference passes around (tf.zeros(..., dtype=tf.float32), tf.cast(x, tf.float32))
float32 = "float32"
float64 = "float64"
int32 = np.int32
int64 = np.int64


class Tensor(np.ndarray):
    """TF tensors are immutable: `x += y` in the reference rebinds, it never writes through an alias
    (e.g. `newMemory += info` at mac_cell.py:337 must not modify `memory`).  Augmented assignment on
    this ndarray subclass is therefore out-of-place; ufunc results keep the subclass."""
    def __iadd__(self, o):
        return np.add(self, o)

    def __isub__(self, o):
        return np.subtract(self, o)

    def __imul__(self, o):
        return np.multiply(self, o)

    def __itruediv__(self, o):
        return np.true_divide(self, o)


def _t(x, dtype=None):
    a = np.asarray(x, dtype=dtype)
    return a.view(Tensor)


def _np_dtype(dtype):
    if dtype in (float32, float64, None):
        return WORK_DTYPE
    return dtype


# --------------------------------------------------------------------------- variables
class _Store(object):
    def __init__(self):
        self.reset()

    def reset(self, values=None, seed=0):
        self.scope = []
        self.vars = {}            # full name -> ndarray (creation order preserved)
        self.provided = dict(values or {})
        self.rng = np.random.RandomState(seed)
        self.dropout_masks = []   # (site shape, keep) -> recorded masks, in call order
        self.uniform_draws = []
        self.mask_source = None   # optional iterator of pre-made masks


_store = _Store()


def reset_shim(values=None, seed=0, dtype=np.float64):
    global WORK_DTYPE
    WORK_DTYPE = dtype
    _store.reset(values, seed)
    return _store


def shim_store():
    return _store


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, reuse=None, **kw):
    name = name_or_scope if name_or_scope is not None else default_name
    _store.scope.append(name)
    try:
        yield name
    finally:
        _store.scope.pop()


def get_variable(name, shape=None, initializer=None, dtype=None, **kw):
    full = "/".join(_store.scope + [name])
    if full in _store.vars:
        return _store.vars[full]
    if shape is None and initializer is not None and not callable(initializer):
        shape = np.shape(initializer)          # tf.get_variable(name, initializer=<tensor>): the shape is the tensor's
    shape = tuple(int(s) for s in (shape if shape is not None else ()))
    if full in _store.provided:
        val = _t(_store.provided[full], dtype=WORK_DTYPE)
        assert val.shape == shape, (full, val.shape, shape)
    elif initializer is None:
        val = _t(_xavier_uniform()(shape, _store.rng), dtype=WORK_DTYPE)   # TF default: glorot_uniform_initializer
    elif not callable(initializer):
        val = _t(np.array(initializer, dtype=WORK_DTYPE))
    else:
        val = _t(initializer(shape, _store.rng), dtype=WORK_DTYPE)
    _store.vars[full] = val
    return val


def random_normal_initializer(mean=0.0, stddev=1.0):
    return lambda shape, rng: mean + stddev * rng.standard_normal(shape)


def zeros_initializer():
    return lambda shape, rng: np.zeros(shape)


def ones_initializer():
    return lambda shape, rng: np.ones(shape)


def constant_initializer(v):
    return lambda shape, rng: np.full(shape, v, dtype=np.float64)


def _xavier_uniform():
    # tf.contrib.layers.xavier_initializer(uniform=True): limit = sqrt(6/(fan_in+fan_out));
    # for a 1-D shape [n] TF's _compute_fans gives fan_in = fan_out = n.
    def init(shape, rng):
        if len(shape) == 0:
            fan_in = fan_out = 1
        elif len(shape) == 1:
            fan_in = fan_out = shape[0]
        else:
            recept = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            fan_in, fan_out = shape[-2] * recept, shape[-1] * recept
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, size=shape)
    return init


def trainable_variables():
    return list(_store.vars.values())


# --------------------------------------------------------------------------- basic ops
def shape(x):
    return np.array(np.shape(x), dtype=np.int64)


def fill(dims, value):
    return np.full(tuple(int(d) for d in dims), value)


def concat(values, axis=0):
    return _t(np.concatenate([np.asarray(v) for v in values], axis=axis))


def reshape(x, newshape):
    return np.reshape(x, tuple(int(s) for s in np.asarray(newshape).reshape(-1)))


def matmul(a, b):
    return np.matmul(a, b)


def expand_dims(x, axis):
    return np.expand_dims(x, axis)


def zeros(shape_, dtype=None):
    return _t(np.zeros(tuple(int(s) for s in shape_), dtype=_np_dtype(dtype)))


def zeros_like(x):
    return np.zeros_like(x)


def tile(x, multiples):
    return _t(np.tile(x, tuple(int(m) for m in multiples)))


def constant(v, dtype=None):
    return _t(v, dtype=_np_dtype(dtype))


def identity(x):
    return x


def stack(values, axis=0):
    return np.stack(values, axis=axis)


def reduce_sum(x, axis=None, keepdims=False):
    return np.sum(x, axis=axis, keepdims=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    return np.mean(x, axis=axis, keepdims=keepdims)


def tanh(x):
    return np.tanh(x)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def maximum(a, b):
    return np.maximum(a, b)


def floor(x):
    return np.floor(x)


def div(a, b):
    return a / b


def to_float(x):
    return _t(x, dtype=WORK_DTYPE)          # `t = to_float(p); t += u` (ops.py:1055-1056) rebinds


def cast(x, dtype):
    return np.asarray(x).astype(_np_dtype(dtype))


def sequence_mask(lengths, maxlen=None):
    lengths = np.asarray(lengths)
    maxlen = int(maxlen if maxlen is not None else lengths.max())
    return np.arange(maxlen)[None, :] < lengths[:, None]


def squared_difference(a, b):
    return (a - b) ** 2


def random_uniform(shape_, minval=0, maxval=1, dtype=None):
    shp = tuple(int(s) for s in shape_)
    if _store.mask_source is not None:
        u = np.asarray(next(_store.mask_source), dtype=WORK_DTYPE)
        assert u.shape == shp
    else:
        u = _store.rng.uniform(minval, maxval, size=shp)
    _store.uniform_draws.append(u)
    return _t(u, dtype=WORK_DTYPE)


# --------------------------------------------------------------------------- tf.nn
class _RNNCell(object):
    pass


LSTMStateTuple = collections.namedtuple("LSTMStateTuple", ("c", "h"))


class BasicLSTMCell(_RNNCell):
    """tf.nn.rnn_cell.BasicLSTMCell as published in TF 1.x `rnn_cell_impl.py`: variables `basic_lstm_cell/kernel`
    [input_depth + num_units, 4 * num_units] (default glorot-uniform initialiser) and `basic_lstm_cell/bias` (zeros);
        gate_inputs = concat([inputs, h], 1) @ kernel + bias;   i, j, f, o = split(gate_inputs, 4, axis=1)
        new_c = c * sigmoid(f + forget_bias) + sigmoid(i) * act(j);   new_h = act(new_c) * sigmoid(o)
    forget_bias defaults to 1.0, the activation to tanh, state_is_tuple to True."""

    def __init__(self, num_units, forget_bias=1.0, state_is_tuple=True, activation=None, reuse=None, name=None):
        self._num_units = int(num_units)
        self._forget_bias = float(forget_bias)
        self._activation = activation or tanh
        self._name = name or "basic_lstm_cell"

    @property
    def state_size(self):
        return LSTMStateTuple(self._num_units, self._num_units)

    @property
    def output_size(self):
        return self._num_units

    def zero_state(self, batch_size, dtype):
        z = np.zeros((int(batch_size), self._num_units), dtype=WORK_DTYPE)
        return LSTMStateTuple(_t(z), _t(z.copy()))

    def __call__(self, inputs, state):
        c, h = state
        n = self._num_units
        with variable_scope(self._name):
            kernel = get_variable("kernel", shape=(inputs.shape[1] + n, 4 * n))
            bias = get_variable("bias", shape=(4 * n,), initializer=zeros_initializer())
        gate_inputs = np.concatenate([np.asarray(inputs), np.asarray(h)], axis=1) @ kernel + bias
        i, j, f, o = np.split(gate_inputs, 4, axis=1)
        new_c = c * sigmoid(f + self._forget_bias) + sigmoid(i) * self._activation(j)
        new_h = self._activation(new_c) * sigmoid(o)
        return _t(new_h), LSTMStateTuple(_t(new_c), _t(new_h))


def _unsupported_cell(kind):
    class _Cell(_RNNCell):
        def __init__(self, *a, **k):
            raise NotImplementedError("%s is outside the restated slice (encType defaults to LSTM, config.py:262)" % kind)
    return _Cell


class _rnn_cell(object):
    RNNCell = _RNNCell
    LSTMStateTuple = LSTMStateTuple
    BasicLSTMCell = BasicLSTMCell
    BasicRNNCell = _unsupported_cell("BasicRNNCell")      # named in ops.createCell's table (ops.py:762-768)
    GRUCell = _unsupported_cell("GRUCell")
    LSTMCell = _unsupported_cell("LSTMCell")


def _reverse_sequence(x, lengths):
    """tf.reverse_sequence(x, lengths, seq_axis=1, batch_axis=0): the first lengths[b] steps reversed, the rest kept."""
    out = np.array(x, copy=True)
    for b, n in enumerate(np.asarray(lengths).astype(np.int64)):
        out[b, :n] = np.asarray(x)[b, :n][::-1]
    return out


def _dynamic_rnn(cell, inputs, sequence_length=None, initial_state=None, dtype=None, scope=None, **kw):
    """tf.nn.dynamic_rnn (batch-major) as published: step t runs the cell on every row; rows with t >= sequence_length
    emit a zero output and carry their state through unchanged (`_rnn_step` with `copy_through`); variables live in
    `scope or "rnn"`."""
    x = np.asarray(inputs)
    B, T = x.shape[0], x.shape[1]
    lengths = np.full((B,), T, np.int64) if sequence_length is None else np.asarray(sequence_length).astype(np.int64)
    state = initial_state if initial_state is not None else cell.zero_state(B, dtype)
    outs = []
    with variable_scope(scope or "rnn"):
        for t in range(T):
            out, new_state = cell(_t(x[:, t]), state)
            live = (t < lengths)[:, None]
            outs.append(np.where(live, out, 0.0))
            state = type(state)(*[_t(np.where(live, n_, o_)) for n_, o_ in zip(new_state, state)])
    return _t(np.stack(outs, axis=1)), state


def _bidirectional_dynamic_rnn(cell_fw, cell_bw, inputs, sequence_length=None, initial_state_fw=None,
                               initial_state_bw=None, dtype=None, scope=None, **kw):
    """tf.nn.bidirectional_dynamic_rnn as published: forward pass in scope "<bidirectional_rnn>/fw"; backward pass on
    `reverse_sequence(inputs, sequence_length)` in "<bidirectional_rnn>/bw", its outputs reversed back the same way."""
    with variable_scope(scope or "bidirectional_rnn"):
        out_fw, st_fw = _dynamic_rnn(cell_fw, inputs, sequence_length, initial_state_fw, dtype, scope="fw")
        x = np.asarray(inputs)
        lengths = np.full((x.shape[0],), x.shape[1]) if sequence_length is None else sequence_length
        out_bw, st_bw = _dynamic_rnn(cell_bw, _reverse_sequence(x, lengths), sequence_length, initial_state_bw, dtype,
                                     scope="bw")
        out_bw = _t(_reverse_sequence(out_bw, lengths))
    return (out_fw, out_bw), (st_fw, st_bw)


class _nn(object):
    rnn_cell = _rnn_cell
    dynamic_rnn = staticmethod(_dynamic_rnn)
    bidirectional_dynamic_rnn = staticmethod(_bidirectional_dynamic_rnn)

    @staticmethod
    def embedding_lookup(params, ids, **kw):
        return _t(np.asarray(params)[np.asarray(ids).astype(np.int64)])

    @staticmethod
    def softmax(x, axis=-1):
        # TF: exp(x - max) / sum(exp(x - max)) along the last axis
        m = np.max(x, axis=axis, keepdims=True)
        e = np.exp(x - m)
        return e / np.sum(e, axis=axis, keepdims=True)

    @staticmethod
    def elu(x):
        return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))

    @staticmethod
    def relu(x):
        return np.maximum(x, 0)

    @staticmethod
    def sigmoid(x):
        return sigmoid(x)

    @staticmethod
    def conv2d(inp, filter=None, strides=None, padding="SAME", **kw):
        # NHWC input, HWIO filter, SAME padding (TF: pad_total = k - 1 for stride 1, extra on the bottom/right)
        x, k = np.asarray(inp), np.asarray(filter)
        assert padding == "SAME" and tuple(strides) == (1, 1, 1, 1), "only what the default stem uses"
        kh, kw = k.shape[0], k.shape[1]
        pt, pl = (kh - 1) // 2, (kw - 1) // 2
        B, H, W, _ = x.shape
        xp = np.zeros((B, H + kh - 1, W + kw - 1, x.shape[3]), dtype=x.dtype)
        xp[:, pt:pt + H, pl:pl + W, :] = x
        out = np.zeros((B, H, W, k.shape[3]), dtype=x.dtype)
        for i in range(kh):
            for j in range(kw):
                out += np.einsum("bhwc,co->bhwo", xp[:, i:i + H, j:j + W, :], k[i, j])
        return _t(out)

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(labels=None, logits=None, **kw):
        # TF: -log_softmax(logits)[label], computed as logsumexp(logits) - logits[label]
        m = np.max(logits, axis=-1, keepdims=True)
        lse = m[..., 0] + np.log(np.sum(np.exp(logits - m), axis=-1))
        idx = np.asarray(labels).astype(np.int64)
        return _t(lse - np.take_along_axis(np.asarray(logits), idx[..., None], axis=-1)[..., 0])

    @staticmethod
    def dropout(x, keep_prob, **kw):
        # TF1: x / keep_prob * floor(keep_prob + U[0,1)).  keep_prob == 1.0 is an exact identity
        # (TF short-circuits a python-number keep_prob of 1 and the formula gives x anyway).
        keep = float(keep_prob)
        if keep == 1.0:
            return x
        u = random_uniform(np.shape(x))
        mask = np.floor(keep + u)
        _store.dropout_masks.append(mask)
        return x / keep * mask


nn = _nn


# --------------------------------------------------------------------------- tf.contrib
class _layers(object):
    xavier_initializer = staticmethod(_xavier_uniform)

    @staticmethod
    def batch_norm(inputs, decay=0.999, center=True, scale=False, epsilon=0.001, is_training=True,
                   updates_collections="update_ops", scope=None, **kw):
        """tf.contrib.layers.batch_norm on a [B, C] input, restated from TF 1.x contrib/layers/python/layers/layers.py:
        rank 2 takes the fused path (reshaped to [B,1,1,C]).  Training: batch mean and BIASED variance normalise; the moving
        mean / variance are updated by assign_moving_average(decay, zero_debias=False) with the batch mean and the
        Bessel-corrected variance FusedBatchNorm returns; `updates_collections=None` forces the update with the forward.
        Inference: the moving statistics normalise.  beta / gamma exist only with center / scale."""
        assert updates_collections is None, "the reference passes updates_collections=None (mac_cell.py:371-373)"
        x = np.asarray(inputs)
        C = x.shape[-1]
        with variable_scope(scope or "BatchNorm"):
            beta = get_variable("beta", [C], initializer=zeros_initializer()) if center else 0.0
            gamma = get_variable("gamma", [C], initializer=ones_initializer()) if scale else 1.0
            full = "/".join(_store.scope)
            mm = get_variable("moving_mean", [C], initializer=zeros_initializer())
            mv = get_variable("moving_variance", [C], initializer=ones_initializer())
            if bool(is_training):
                n = x.shape[0]
                mean = np.mean(x, axis=0)
                var = np.mean((x - mean) ** 2, axis=0)
                y = (x - mean) / np.sqrt(var + epsilon) * gamma + beta
                unbiased = var * (float(n) / max(n - 1, 1))
                _store.vars[full + "/moving_mean"] = _t(mm - (mm - mean) * (1.0 - decay))
                _store.vars[full + "/moving_variance"] = _t(mv - (mv - unbiased) * (1.0 - decay))
            else:
                y = (x - mm) / np.sqrt(mv + epsilon) * gamma + beta
        return _t(y)


class _rnn(object):
    DropoutWrapper = object


class _contrib(object):
    layers = _layers
    rnn = _rnn


contrib = _contrib
