"""An eager stand-in for the slice of TensorFlow 1.x that the reference's model-building code calls, so that the
reference's OWN graph code (code/imagebert_zk/pixelbert.py + model_triple.py, code/imagebert_lds/src/pixelmodel.py)
runs unmodified in the dev container (TF 1.12 / Python 2 cannot be installed here) and its outputs can pin the
restatement in oracle/imagebert.py (tools/make_golden.py --tf-shim -> tests/golden/{zk,lds}_ref_shim_*.npz).

What this pins and what it does not: the WIRING — which op consumes which tensor, in which order, under which
variable name, with which shapes, axes, masks, constants and quirks — is the reference's code, executed line by line.
The arithmetic of each TensorFlow op is restated here from the TF 1.12 API documentation (file:line of the call
sites in the docstrings below), on fp32 torch CPU tensors:
  slim.conv2d / layers.conv2d          NHWC, padding 'SAME' (total = k - 1, left = total // 2), activation_fn
                                       DEFAULTS TO relu, variables <scope>/weights [kh, kw, in, out], <scope>/biases
  contrib.layers.fully_connected       activation_fn defaults to relu (every call site passes None), variables
                                       <scope or "fully_connected">/weights [in, out], /biases
  tf.layers.dense                      variables <name or "dense">/kernel [in, out], /bias; activation after the bias
  contrib.layers.layer_norm            moments over the last axis (biased variance), variance_epsilon 1e-12,
                                       variables <scope or "LayerNorm">/beta, /gamma
  tf.nn.l2_normalize(x, axis, eps)     x * rsqrt(max(sum(x^2, axis), eps)), eps default 1e-12
  tf.layers.dropout(training=False)    identity;  tf.sequence_mask, one_hot, gather_nd, softmax, ... as documented
tf.get_variable returns the array stored under the full variable-scope path: a name the reference asks for that the
synthetic weight set (= the checkpoint importers' name map) does not hold is a KeyError, which pins the names too.
Test infrastructure only; nothing under the package imports it.
"""
import contextlib
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F


class TensorShape(list):
    def as_list(self):
        return list(self)

    @property
    def ndims(self):
        return len(self)


_DT = {"float32": torch.float32, "int32": torch.int32, "int64": torch.int64, "bool": torch.bool,
       "float64": torch.float64}


def _raw(x):
    return x.t if isinstance(x, Tensor) else x


def _t(x, dtype=None):
    """Anything the reference passes where TF accepts a tensor-like: Tensor, numpy array, python list / scalar."""
    if isinstance(x, Tensor):
        t = x.t
    elif isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        t = torch.stack([_t(e) for e in x])
    else:
        a = np.asarray(x)
        if a.dtype == np.float64 and dtype is None:
            a = a.astype(np.float32)          # TF's default float is float32
        if a.dtype == np.int64 and dtype is None and not isinstance(x, np.ndarray):
            a = a.astype(np.int32)            # python ints become int32 constants
        t = torch.from_numpy(np.ascontiguousarray(a))
    return t if dtype is None else t.to(dtype)


class Tensor:
    """Wrapper giving a torch tensor the few attributes of tf.Tensor the reference touches."""

    def __init__(self, t, name="shim"):
        self.t = t if isinstance(t, torch.Tensor) else _t(t)
        self.name = name

    @property
    def shape(self):
        return TensorShape(self.t.shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return self.t.dtype

    def _bin(self, other, fn, rev=False):
        a, b = self.t, _t(other)
        if b.dtype != a.dtype and not torch.is_floating_point(b) and torch.is_floating_point(a):
            b = b.to(a.dtype)
        elif b.dtype != a.dtype and b.dim() == 0:
            b = b.to(a.dtype)
        return Tensor(fn(b, a) if rev else fn(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __neg__(self): return Tensor(-self.t)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __lt__(self, o): return self._bin(o, torch.lt)

    def __getitem__(self, idx):
        return Tensor(self.t[idx])

    def numpy(self):
        return self.t.detach().numpy()


# ------------------------------------------------------------------------------------------------ variables
class _State:
    weights = {}
    scopes = []          # variable-scope path components
    used = set()
    layer_counts = {}    # (scope path, base layer name) -> uses, for TF's "dense", "dense_1", ... uniquification


def _scope_path(name=None):
    parts = [p for p in _State.scopes if p]
    if name:
        parts.append(name)
    return "/".join(parts)


class _VarScope:
    def __init__(self):
        self.name = _scope_path()

    def reuse_variables(self):
        pass


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, reuse=None, **_):
    """tf.variable_scope(name_or_scope, default_name): default_name is used only when name_or_scope is None
    (pixelbert.py:200 `tf.variable_scope("bert", scope)` with scope=None opens "bert")."""
    name = name_or_scope if name_or_scope is not None else default_name
    if isinstance(name, _VarScope):
        saved, _State.scopes = _State.scopes, name.name.split("/")
        try:
            yield name
        finally:
            _State.scopes = saved
        return
    _State.scopes.append(name)
    try:
        yield _VarScope()
    finally:
        _State.scopes.pop()


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **_):
    full = _scope_path(name)
    if full not in _State.weights:
        raise KeyError(f"the reference asks for variable '{full}', which the weight set does not hold")
    w = _State.weights[full]
    if shape is not None and list(w.shape) != [int(s) for s in shape]:
        raise ValueError(f"variable '{full}': reference shape {list(shape)}, weight set has {list(w.shape)}")
    _State.used.add(full)
    return Tensor(torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)), name=full + ":0")


def _unique_layer_scope(base):
    key = (_scope_path(), base)
    n = _State.layer_counts.get(key, 0)
    _State.layer_counts[key] = n + 1
    return base if n == 0 else f"{base}_{n}"


# ------------------------------------------------------------------------------------------------ ops
def reshape(x, shape, name=None):
    shape = [int(_raw(s)) if not isinstance(s, int) else s for s in (shape.as_list() if isinstance(shape, TensorShape) else list(shape))]
    return Tensor(_t(x).reshape(shape))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return Tensor(torch.matmul(a, b))


def cast(x, dtype, name=None):
    return Tensor(_t(x).to(dtype))


def concat(values, axis, name=None):
    ts = [_t(v) for v in values]
    if any(torch.is_floating_point(t) for t in ts):
        ts = [t.to(torch.float32) for t in ts]
    return Tensor(torch.cat(ts, dim=axis))


def expand_dims(x, axis=None, name=None, dim=None):
    axis = dim if axis is None else axis
    if isinstance(axis, (list, tuple)):
        axis = axis[0]
    return Tensor(_t(x).unsqueeze(axis))


def squeeze(x, axis=None, name=None, squeeze_dims=None):
    axis = squeeze_dims if axis is None else axis
    t = _t(x)
    if axis is None:
        return Tensor(t.squeeze())
    for a in sorted([axis] if isinstance(axis, int) else list(axis), reverse=True):
        t = t.squeeze(a)
    return Tensor(t)


def one_hot(indices, depth, dtype=torch.float32, **_):
    return Tensor(F.one_hot(_t(indices).long(), int(depth)).to(dtype))


def constant(value, dtype=None, shape=None, name=None):
    return Tensor(_t(value, dtype))


def _reduce(fn, x, axis, keepdims):
    t = _t(x)
    if axis is None:
        return Tensor(fn(t))
    return Tensor(fn(t, dim=axis, keepdim=bool(keepdims)))


def reduce_mean(x, axis=None, keepdims=False, name=None, keep_dims=None):
    return _reduce(torch.mean, x, axis, keepdims if keep_dims is None else keep_dims)


def reduce_sum(x, axis=None, keepdims=False, name=None, keep_dims=None):
    return _reduce(torch.sum, x, axis, keepdims if keep_dims is None else keep_dims)


def transpose(x, perm=None, name=None):
    t = _t(x)
    return Tensor(t.permute(*perm) if perm is not None else t.t())


def ones(shape, dtype=torch.float32, name=None):
    return Tensor(torch.ones([int(_raw(s)) for s in shape], dtype=dtype))


def zeros(shape, dtype=torch.float32, name=None):
    return Tensor(torch.zeros([int(_raw(s)) for s in shape], dtype=dtype))


def gather(params, indices, name=None, axis=0):
    return Tensor(torch.index_select(_t(params), axis, _t(indices).long().reshape(-1)).reshape(
        list(_t(indices).shape) + list(_t(params).shape[1:])) if axis == 0 else None)


def embedding_lookup(params, ids, name=None, **_):
    p, i = _t(params), _t(ids).long()
    return Tensor(p[i])


def gather_nd(params, indices, name=None):
    p, i = _t(params), _t(indices).long()
    return Tensor(p[tuple(i[..., k] for k in range(i.shape[-1]))])


def shape(x, name=None, out_type=None):
    return TensorShape(int(s) for s in _t(x).shape)


def range_(start, limit=None, delta=1, dtype=None, name=None):
    start = int(_raw(start)) if not isinstance(start, int) else start
    if limit is None:
        start, limit = 0, start
    return Tensor(torch.arange(start, int(_raw(limit)), delta, dtype=torch.int32))


def sequence_mask(lengths, maxlen=None, dtype=torch.bool, name=None):
    l = _t(lengths).long().reshape(-1, 1)
    return Tensor((torch.arange(int(maxlen)).reshape(1, -1) < l).to(dtype))


def clip_by_value(x, lo, hi, name=None):
    return Tensor(torch.clamp(_t(x), float(lo), float(hi)))


def l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2, axis), epsilon)) (model_triple.py:60, 64)."""
    axis = dim if axis is None else axis
    t = _t(x)
    ss = torch.sum(t * t, dim=axis, keepdim=True)
    return Tensor(t * torch.rsqrt(torch.clamp(ss, min=float(epsilon))))


def softmax(x, axis=-1, name=None, dim=None):
    return Tensor(torch.softmax(_t(x), dim=axis if dim is None else dim))


def log_softmax(x, axis=-1, name=None, dim=None):
    return Tensor(torch.log_softmax(_t(x), dim=axis if dim is None else dim))


def softmax_cross_entropy_with_logits(labels=None, logits=None, **_):
    return Tensor(-(torch.log_softmax(_t(logits), -1) * _t(labels)).sum(-1))


def bias_add(x, b, name=None):
    return Tensor(_t(x) + _t(b))


def dropout_nn(x, keep_prob=None, **_):
    if keep_prob is not None and float(keep_prob) != 1.0:
        raise RuntimeError("tf.nn.dropout with keep_prob < 1 reached at inference")
    return x


def layers_dropout(x, rate=0.5, training=False, **_):
    if training:
        raise RuntimeError("tf.layers.dropout(training=True) reached at inference")
    return x


def dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, name=None, reuse=None, **_):
    """tf.layers.dense (pixelbert.py:767-788, 913-917, ...): y = act(x . kernel + bias), kernel [in, units]."""
    scope = name if name is not None else _unique_layer_scope("dense")
    with variable_scope(scope):
        x = _t(inputs)
        k = get_variable("kernel", [x.shape[-1], units])
        y = torch.matmul(x, k.t)
        if use_bias:
            y = y + get_variable("bias", [units]).t
    out = Tensor(y)
    return activation(out) if activation is not None else out


def relu(x, name=None):
    return Tensor(torch.relu(_t(x)))


def fully_connected(inputs, num_outputs, activation_fn=relu, normalizer_fn=None, scope=None, reuse=None, **_):
    """tf.contrib.layers.fully_connected / slim.fully_connected: activation_fn DEFAULTS to relu (model_triple.py:191,
    pixelbert.py:451 pass None), variables weights [in, out] and biases under scope or "fully_connected"."""
    with variable_scope(scope if scope is not None else _unique_layer_scope("fully_connected")):
        x = _t(inputs)
        w = get_variable("weights", [x.shape[-1], num_outputs])
        y = torch.matmul(x, w.t) + get_variable("biases", [num_outputs]).t
    out = Tensor(y)
    return activation_fn(out) if activation_fn is not None else out


def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", activation_fn=relu, scope=None, **_):
    """slim.conv2d (model_triple.py:189, 193): NHWC input, weights [kh, kw, in, out], 'SAME' padding with the extra
    element on the right / bottom, activation_fn DEFAULTS to relu."""
    assert padding == "SAME" and stride in (1, [1, 1], (1, 1))
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    with variable_scope(scope if scope is not None else _unique_layer_scope("Conv")):
        x = _t(inputs)                                      # [N, H, W, C]
        w = get_variable("weights", [kh, kw, x.shape[-1], num_outputs]).t
        b = get_variable("biases", [num_outputs]).t
    xt = x.permute(0, 3, 1, 2)
    ph, pw = kh - 1, kw - 1
    xt = F.pad(xt, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    y = F.conv2d(xt, w.permute(3, 2, 0, 1), b).permute(0, 2, 3, 1)
    out = Tensor(y.contiguous())
    return activation_fn(out) if activation_fn is not None else out


def layer_norm(inputs, center=True, scale=True, begin_norm_axis=1, begin_params_axis=-1, scope=None, **_):
    """tf.contrib.layers.layer_norm (pixelbert.py:414-417): nn.moments over the normalised axes (biased variance),
    tf.nn.batch_normalization with variance_epsilon = 1e-12, beta / gamma over the last axis."""
    x = _t(inputs)
    assert begin_norm_axis in (-1, x.dim() - 1) and begin_params_axis in (-1, x.dim() - 1)
    with variable_scope(scope if scope is not None else _unique_layer_scope("LayerNorm")):
        beta = get_variable("beta", [x.shape[-1]]).t
        gamma = get_variable("gamma", [x.shape[-1]]).t
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return Tensor((x - mean) * torch.rsqrt(var + 1e-12) * gamma + beta)


class _Init:
    """Initialisers, ConfigProto (pixelbert.py:33-36 sets gpu_options on it at import time), Session: inert."""

    def __init__(self, *a, **k):
        self.gpu_options = types.SimpleNamespace()


@contextlib.contextmanager
def _noop_ctx(*a, **k):
    yield


def install(weights):
    """Registers the stand-in modules under the names the reference imports and binds the weight set."""
    _State.weights = dict(weights)
    _State.scopes, _State.used, _State.layer_counts = [], set(), {}

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    tf = mod("tensorflow")
    tf.Tensor = Tensor
    tf.float32, tf.float64, tf.int32, tf.int64, tf.bool = torch.float32, torch.float64, torch.int32, torch.int64, torch.bool
    tf.AUTO_REUSE = "AUTO_REUSE"
    tf.variable_scope, tf.get_variable = variable_scope, get_variable
    tf.get_variable_scope = lambda: _VarScope()
    for fn in (reshape, matmul, cast, concat, expand_dims, squeeze, one_hot, constant, reduce_mean, reduce_sum, transpose,
               ones, zeros, gather, gather_nd, shape, sequence_mask, clip_by_value):
        setattr(tf, fn.__name__, fn)
    tf.range = range_
    tf.tanh = lambda x, name=None: Tensor(torch.tanh(_t(x)))
    tf.pow = lambda x, y, name=None: Tensor(torch.pow(_t(x), y))
    tf.multiply = lambda a, b, name=None: Tensor(_t(a)) * b
    tf.subtract = lambda a, b, name=None: Tensor(_t(a)) - b
    tf.greater = lambda a, b, name=None: Tensor(_t(a)) > b
    tf.argmax = lambda x, axis=None, **_: Tensor(torch.argmax(_t(x), dim=axis))
    tf.slice = lambda x, begin, size, name=None: Tensor(_t(x)[tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))])
    tf.assert_less_equal = lambda a, b, **_: None
    tf.control_dependencies = _noop_ctx
    tf.zeros_initializer = tf.truncated_normal_initializer = tf.constant_initializer = _Init
    tf.ConfigProto = tf.Session = _Init
    tf.embedding_lookup = embedding_lookup
    tf.gfile = types.SimpleNamespace(GFile=lambda p, m="r": open(p, m))
    tf.train = types.SimpleNamespace(list_variables=lambda *_: [])
    tf.nn = types.SimpleNamespace(softmax=softmax, log_softmax=log_softmax, relu=relu, embedding_lookup=embedding_lookup,
                                  dropout=dropout_nn, bias_add=bias_add, l2_normalize=l2_normalize,
                                  softmax_cross_entropy_with_logits=softmax_cross_entropy_with_logits)
    tf.layers = types.SimpleNamespace(dense=dense, dropout=layers_dropout)

    contrib = mod("tensorflow.contrib")
    slim = mod("tensorflow.contrib.slim")
    slim.conv2d, slim.fully_connected = conv2d, fully_connected
    slim.arg_scope = _noop_ctx
    slim.batch_norm = slim.dropout = object()
    nets = mod("tensorflow.contrib.slim.nets")
    nets.resnet_v1 = mod("tensorflow.contrib.slim.nets.resnet_v1")
    slim.nets = nets
    layers = mod("tensorflow.contrib.layers")
    layers.fully_connected, layers.layer_norm = fully_connected, layer_norm
    layers.xavier_initializer = _Init
    contrib.slim, contrib.layers, contrib.rnn = slim, layers, mod("tensorflow.contrib.rnn")
    tf.contrib = contrib
    py = mod("tensorflow.python")
    ops = mod("tensorflow.python.ops")
    ops.variable_scope = mod("tensorflow.python.ops.variable_scope")
    ops.math_ops = mod("tensorflow.python.ops.math_ops")
    py.ops = ops
    tf.python = py
    return tf


def used_variables():
    return set(_State.used)
