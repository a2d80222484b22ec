"""A minimal TensorFlow-1.x API shim, eager, backed by torch-CPU.  TEST INFRASTRUCTURE ONLY.

Purpose: execute the reference's own source files -- /root/reference/utils/layers.py,
models/gat.py, models/base_gattn.py -- UNMODIFIED in this container (which has no TensorFlow), so
that the wiring the reference defines (which op feeds which, in which order, with which variables) is
what produces the golden fixtures under tests/golden/ref_*.npz, not our restatement of it.

What this does and does not pin.  Each ``tf.*`` entry point below implements the documented TF1
semantics of that op [external: TF 1.x API docs; defaults cited inline] on torch tensors of one run
dtype (fp64 = gold, fp32 = reference precision).  So the fixtures pin the reference's COMPOSITION
of ops bit-for-bit in structure -- variable creation order and names, where dropout sits relative to
f_1/f_2, the bias add before the softmax, the per-node softmax over meta-paths, the L2 over every
variable, Adam -- while the arithmetic of each primitive is torch's (IEEE fp64), not Eigen's.

Variables: TF1 creates them implicitly in call order with auto-generated names
(``conv1d``, ``conv1d_1`` ..., ``BiasAdd``, ``Variable``, ``dense``).  ``Store`` reproduces the
naming, and takes initial values either from a queue in creation order (``feed``) or draws them with
the initialiser TF would use (glorot-uniform kernels, zero biases, ``random_normal``).
Dropout: ``tf.nn.dropout`` takes its keep mask from ``Store.masks`` (a queue in call order) when one is
fed, so the reference's three dropout sites can be compared exactly against an implementation that is
handed the same masks.

Graph mode is not emulated: calls execute immediately.  ``opt.minimize(loss)`` therefore computes the
gradients with torch autograd (the analogue of TF's graph autodiff), stores them on the Store and applies
the TF1 Adam update in place.
"""
from __future__ import annotations

import math
import sys
import types
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------
# tensors
# ----------------------------------------------------------------------------------------------
class Dimension:
    def __init__(self, v):
        self.value = None if v is None else int(v)

    def __int__(self):
        return self.value

    __index__ = __int__

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dimension) else o)

    def __ne__(self, o):
        return not self.__eq__(o)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return f"Dimension({self.value})"


class TensorShape:
    def __init__(self, dims):
        self.dims = [Dimension(d) for d in dims]

    def __getitem__(self, i):
        return self.dims[i]

    def __len__(self):
        return len(self.dims)

    def as_list(self):
        return [d.value for d in self.dims]

    def __repr__(self):
        return f"TensorShape({self.as_list()})"


class _DType:
    def __init__(self, name, kind):
        self.name, self.kind = name, kind


float32 = _DType("float32", "float")     # "the run's float type": fp64 in a gold run, fp32 otherwise
float64 = _DType("float64", "float")
int32 = _DType("int32", "int32")
int64 = _DType("int64", "int64")
bool_ = _DType("bool", "bool")


def _unwrap(x):
    return x.t if isinstance(x, Tensor) else x


class Tensor:
    def __init__(self, t: torch.Tensor, name: str = ""):
        self.t = t
        self.name = name

    # --- shape protocol the reference uses: x.shape[-1], x.shape[2].value, set_shape ---
    @property
    def shape(self):
        return TensorShape(self.t.shape)

    def get_shape(self):
        return self.shape

    def set_shape(self, shape):
        want = [int(s) for s in shape]
        assert list(self.t.shape) == want, f"set_shape {want} on tensor of shape {list(self.t.shape)}"

    @property
    def dtype(self):
        return self.t.dtype

    def numpy(self):
        return self.t.detach().numpy()

    # --- arithmetic (TF broadcasting == numpy/torch broadcasting) ---
    def __add__(self, o):
        return Tensor(self.t + _coerce(o, self.t))

    __radd__ = __add__

    def __sub__(self, o):
        return Tensor(self.t - _coerce(o, self.t))

    def __rsub__(self, o):
        return Tensor(_coerce(o, self.t) - self.t)

    def __mul__(self, o):
        if isinstance(o, SparseTensor):
            return o.__mul__(self)
        return Tensor(self.t * _coerce(o, self.t))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return Tensor(self.t / _coerce(o, self.t))

    def __rtruediv__(self, o):
        return Tensor(_coerce(o, self.t) / self.t)

    def __neg__(self):
        return Tensor(-self.t)

    def __getitem__(self, idx):
        return Tensor(self.t[idx])

    def __repr__(self):
        return f"<shim Tensor {self.name} {tuple(self.t.shape)} {self.t.dtype}>"


def _coerce(o, like: torch.Tensor):
    o = _unwrap(o)
    if isinstance(o, torch.Tensor):
        return o
    if isinstance(o, Dimension):
        o = o.value
    return torch.as_tensor(o, dtype=like.dtype if like.is_floating_point() or not isinstance(o, float) else STORE.dtype)


class Variable(Tensor):
    """tf.Variable(initial_value): a trainable leaf.  Auto-named Variable, Variable_1, ... (TF1 default)."""

    def __init__(self, initial_value=None, trainable=True, name=None, _kind="Variable", _given_name=None):
        if _given_name is None:
            _given_name = STORE.unique(name or "Variable") + ":0"
        init = _unwrap(initial_value)
        value = STORE.initial(_given_name, _kind, tuple(init.shape), lambda: init)
        t = value.detach().clone().to(STORE.dtype).requires_grad_(trainable)
        super().__init__(t, _given_name)
        self.trainable = trainable
        STORE.variables.append(self)


class SparseTensor:
    """tf.SparseTensor(indices [nnz, rank] int64, values [nnz], dense_shape)."""

    def __init__(self, indices, values, dense_shape):
        self.indices = torch.as_tensor(_unwrap(indices), dtype=torch.int64)
        v = _unwrap(values)
        self.values_t = v if isinstance(v, torch.Tensor) else torch.as_tensor(v, dtype=STORE.dtype)
        ds = _unwrap(dense_shape)
        self.dense_shape_l = [int(d) for d in (ds.tolist() if hasattr(ds, "tolist") else ds)]

    @property
    def values(self):
        return Tensor(self.values_t)

    @property
    def dense_shape(self):
        return self.dense_shape_l

    def __mul__(self, dense):
        """sparse * dense (sparse_dense_cwise_mul): every stored value times the dense operand broadcast to
        the sparse shape, taken at the stored index.  Only the dense side may broadcast."""
        d = _unwrap(dense)
        if not isinstance(d, torch.Tensor):
            return SparseTensor(self.indices, self.values_t * d, self.dense_shape_l)
        d = d.broadcast_to(self.dense_shape_l)
        return SparseTensor(self.indices, self.values_t * d[tuple(self.indices.t())], self.dense_shape_l)

    __rmul__ = __mul__


# ----------------------------------------------------------------------------------------------
# the variable / dropout store (the analogue of TF's default graph)
# ----------------------------------------------------------------------------------------------
class Store:
    def __init__(self):
        self.reset()

    def reset(self, dtype=torch.float64, seed: int = 0):
        self.dtype = dtype
        self.variables: List[Variable] = []
        self.counters: Dict[str, int] = {}
        self.queue: Optional[List] = None     # initial values in creation order: (kind, array)
        self.masks: Optional[List] = None     # dropout keep masks in call order
        self.dropout_calls: List = []         # (shape, keep_prob) of every tf.nn.dropout call
        self.gen = torch.Generator().manual_seed(seed)
        self.last_grads: Optional[List[torch.Tensor]] = None

    def unique(self, base: str) -> str:
        """TF1 name uniquification in one scope: base, base_1, base_2, ..."""
        n = self.counters.get(base, 0)
        self.counters[base] = n + 1
        return base if n == 0 else f"{base}_{n}"

    def feed(self, values: Sequence):
        """Initial values for the variables about to be created, in creation order: (kind, array) pairs."""
        self.queue = list(values)

    def feed_masks(self, masks: Sequence):
        self.masks = [torch.as_tensor(m) for m in masks]

    def initial(self, name, kind, shape, default_fn):
        if self.queue is not None:
            assert self.queue, f"variable {name} {shape} created but the feed queue is empty"
            k, v = self.queue.pop(0)
            v = torch.as_tensor(np.asarray(v))
            assert k == kind, f"{name}: expected a {kind} initial value, feed has {k}"
            assert tuple(v.shape) == tuple(shape), f"{name}: feed shape {tuple(v.shape)} vs created {tuple(shape)}"
            return v
        return default_fn()

    def trainable(self):
        return [v for v in self.variables if v.trainable]


STORE = Store()


def _glorot_uniform(shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))      # tf.glorot_uniform_initializer (default kernel init of tf.layers)
    return (torch.rand(shape, generator=STORE.gen, dtype=torch.float64) * 2 - 1) * lim


def _new_var(name, kind, shape, default_fn):
    v = Variable.__new__(Variable)
    value = STORE.initial(name, kind, tuple(shape), default_fn)
    Tensor.__init__(v, value.detach().clone().to(STORE.dtype).requires_grad_(True), name)
    v.trainable = True
    STORE.variables.append(v)
    return v


# ----------------------------------------------------------------------------------------------
# tf.layers
# ----------------------------------------------------------------------------------------------
def _conv1d(inputs, filters, kernel_size, use_bias=True, activation=None, name=None, **kw):
    """tf.layers.conv1d, channels_last, stride 1, 'valid'.  Only kernel_size == 1 is used by the reference
    (utils/layers.py:20,23,24,40): a per-position matmul with kernel (1, in, filters), glorot-uniform over
    fan_in = 1*in, fan_out = 1*filters, optional zero-initialised bias (filters,)."""
    assert not kw, f"unsupported conv1d arguments {kw}"
    ks = kernel_size[0] if isinstance(kernel_size, (tuple, list)) else kernel_size
    assert int(ks) == 1, "the shim implements kernel_size=1 only"
    x = _unwrap(inputs)
    filters = int(filters)
    cin = x.shape[-1]
    scope = STORE.unique(name or "conv1d")
    kernel = _new_var(f"{scope}/kernel:0", "conv1d/kernel", (1, cin, filters),
                      lambda: _glorot_uniform((1, cin, filters), cin, filters))
    y = torch.matmul(x, kernel.t[0])
    if use_bias:
        bias = _new_var(f"{scope}/bias:0", "conv1d/bias", (filters,), lambda: torch.zeros(filters, dtype=torch.float64))
        y = y + bias.t
    out = Tensor(y)
    return activation(out) if activation is not None else out


def _dense(inputs, units, activation=None, use_bias=True, name=None, **kw):
    """tf.layers.dense: kernel (in, units) glorot-uniform, bias zeros (models/gat.py:68)."""
    assert not kw, f"unsupported dense arguments {kw}"
    x = _unwrap(inputs)
    cin = x.shape[-1]
    units = int(units)
    scope = STORE.unique(name or "dense")
    kernel = _new_var(f"{scope}/kernel:0", "dense/kernel", (cin, units), lambda: _glorot_uniform((cin, units), cin, units))
    y = torch.matmul(x, kernel.t)
    if use_bias:
        bias = _new_var(f"{scope}/bias:0", "dense/bias", (units,), lambda: torch.zeros(units, dtype=torch.float64))
        y = y + bias.t
    out = Tensor(y)
    return activation(out) if activation is not None else out


def _contrib_bias_add(inputs, **kw):
    """tf.contrib.layers.bias_add: variable_scope(None, 'BiasAdd'), variable 'biases' (last dim,), zeros init
    (utils/layers.py:35)."""
    assert not kw, f"unsupported bias_add arguments {kw}"
    x = _unwrap(inputs)
    n = x.shape[-1]
    scope = STORE.unique("BiasAdd")
    b = _new_var(f"{scope}/biases:0", "BiasAdd/biases", (n,), lambda: torch.zeros(n, dtype=torch.float64))
    return Tensor(x + b.t)


# ----------------------------------------------------------------------------------------------
# tf.nn
# ----------------------------------------------------------------------------------------------
def _dropout(x, keep_prob, **kw):
    """tf.nn.dropout(x, keep_prob) = x / keep_prob * floor(keep_prob + U[0,1)).  The keep mask comes from the
    Store's mask queue when one was fed (call order), else from the Store's generator."""
    assert not kw
    t = _unwrap(x)
    keep = float(_unwrap(keep_prob))
    STORE.dropout_calls.append((tuple(t.shape), keep))
    if STORE.masks is not None:
        assert STORE.masks, "tf.nn.dropout called but the mask queue is empty"
        m = STORE.masks.pop(0).to(t.dtype).reshape(t.shape)
    else:
        m = torch.floor(keep + torch.rand(t.shape, generator=STORE.gen, dtype=t.dtype))
    return Tensor(t / keep * m)


def _softmax(logits, axis=-1, name=None, dim=None):
    """tf.nn.softmax over the last axis: exp(x - max) / sum."""
    if dim is not None:
        axis = dim
    return Tensor(torch.softmax(_unwrap(logits), dim=axis))


def _leaky_relu(features, alpha=0.2, name=None):
    """tf.nn.leaky_relu: default alpha 0.2 (utils/layers.py:27 relies on it)."""
    return Tensor(torch.nn.functional.leaky_relu(_unwrap(features), alpha))


def _elu(features, name=None):
    return Tensor(torch.nn.functional.elu(_unwrap(features)))


def _l2_loss(t, name=None):
    """tf.nn.l2_loss = sum(t**2) / 2."""
    x = _unwrap(t)
    return Tensor((x * x).sum() / 2)


def _softmax_xent(logits=None, labels=None, dim=-1, name=None, _sentinel=None):
    """tf.nn.softmax_cross_entropy_with_logits: -sum(labels * log_softmax(logits)); labels are cast to the
    logits' type and (v1) receive no gradient."""
    lg = _unwrap(logits)
    lb = _unwrap(labels).detach().to(lg.dtype)
    return Tensor(-(lb * torch.log_softmax(lg, dim=dim)).sum(dim))


def _sparse_softmax_xent(labels=None, logits=None, name=None, _sentinel=None):
    lg = _unwrap(logits)
    lb = _unwrap(labels).long()
    return Tensor(-torch.log_softmax(lg, -1).gather(-1, lb.unsqueeze(-1)).squeeze(-1))


def _sigmoid_xent(logits=None, labels=None, name=None, _sentinel=None):
    lg = _unwrap(logits)
    lb = _unwrap(labels).to(lg.dtype)
    return Tensor(torch.clamp(lg, min=0) - lg * lb + torch.log1p(torch.exp(-lg.abs())))


# ----------------------------------------------------------------------------------------------
# plain ops
# ----------------------------------------------------------------------------------------------
def _tdtype(d):
    if isinstance(d, _DType):
        return {"float": STORE.dtype, "int32": torch.int32, "int64": torch.int64, "bool": torch.bool}[d.kind]
    return d


def transpose(a, perm=None, name=None):
    t = _unwrap(a)
    if perm is None:
        perm = list(range(t.dim()))[::-1]
    return Tensor(t.permute(*[int(p) for p in perm]))


def matmul(a, b, name=None):
    return Tensor(torch.matmul(_unwrap(a), _unwrap(b)))


def tensordot(a, b, axes, name=None):
    return Tensor(torch.tensordot(_unwrap(a), _unwrap(b), dims=axes))


def concat(values, axis, name=None):
    vals = [_unwrap(v) for v in values]
    if len(vals) == 0:
        raise ValueError("tf.concat of an empty list")         # what TF1 raises too (models/gat.py:165 quirk)
    return Tensor(torch.cat(vals, dim=int(axis)))


def squeeze(input, axis=None, name=None):
    t = _unwrap(input)
    return Tensor(t.squeeze() if axis is None else t.squeeze(axis))


def expand_dims(input, axis=None, name=None, dim=None):
    if axis is None:
        axis = dim
    return Tensor(_unwrap(input).unsqueeze(int(axis)))


def add_n(inputs, name=None):
    ts = [_unwrap(i) for i in inputs]
    out = ts[0]
    for t in ts[1:]:
        out = out + t
    return Tensor(out)


def reduce_sum(input_tensor, axis=None, keepdims=False, name=None):
    t = _unwrap(input_tensor)
    return Tensor(t.sum() if axis is None else t.sum(dim=axis, keepdim=keepdims))


def reduce_mean(input_tensor, axis=None, keepdims=False, name=None):
    t = _unwrap(input_tensor)
    return Tensor(t.mean() if axis is None else t.mean(dim=axis, keepdim=keepdims))


def tanh(x, name=None):
    return Tensor(torch.tanh(_unwrap(x)))


def cast(x, dtype, name=None):
    return Tensor(_unwrap(x).to(_tdtype(dtype)))


def equal(x, y, name=None):
    return Tensor(_unwrap(x) == _unwrap(y))


def argmax(input, axis=None, name=None):
    return Tensor(torch.argmax(_unwrap(input), dim=axis))


def reshape(tensor, shape, name=None):
    return Tensor(_unwrap(tensor).reshape([int(s) for s in shape]))


def multiply(x, y, name=None):
    return Tensor(_unwrap(x) * _unwrap(y))


def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    return Tensor(torch.randn([int(s) for s in shape], generator=STORE.gen, dtype=torch.float64) * stddev + mean)


def constant(value, dtype=None, name=None):
    t = torch.as_tensor(np.asarray(value))
    if t.is_floating_point():
        t = t.to(STORE.dtype)
    return Tensor(t if dtype is None else t.to(_tdtype(dtype)))


class name_scope:
    """name_scope only prefixes op names; tf.layers variable names ignore it."""

    def __init__(self, name, *a, **k):
        self.name = name

    def __enter__(self):
        return self.name

    def __exit__(self, *exc):
        return False


def trainable_variables():
    return STORE.trainable()


# ----------------------------------------------------------------------------------------------
# sparse ops (utils/layers.py:85-127)
# ----------------------------------------------------------------------------------------------
def sparse_add(a: SparseTensor, b: SparseTensor, thresh=0):
    """Sum of two SparseTensors over the union of their index sets, canonical (row-major) order; thresh=0
    keeps explicit zeros."""
    assert a.dense_shape_l == b.dense_shape_l
    idx = torch.cat([a.indices, b.indices], 0)
    val = torch.cat([a.values_t, b.values_t], 0)
    uniq, inv = torch.unique(idx, dim=0, return_inverse=True)       # sorted lexicographically
    out = torch.zeros(uniq.shape[0], dtype=val.dtype).index_add(0, inv, val)
    return SparseTensor(uniq, out, a.dense_shape_l)


def _row_keys(sp: SparseTensor):
    lead = sp.indices[:, :-1]
    mult = [1]
    for d in reversed(sp.dense_shape_l[1:-1]):
        mult.insert(0, mult[0] * d)
    mult = torch.tensor(mult[-lead.shape[1]:] if lead.shape[1] else [], dtype=torch.int64)
    return (lead * mult).sum(1) if lead.shape[1] else torch.zeros(sp.indices.shape[0], dtype=torch.int64)


def sparse_softmax(sp: SparseTensor, name=None):
    """tf.sparse_softmax: softmax along the last dimension over the STORED entries of each row only."""
    keys = _row_keys(sp)
    _, rows = torch.unique(keys, return_inverse=True)
    n = int(rows.max().item()) + 1 if rows.numel() else 0
    v = sp.values_t
    m = torch.full((n,), -math.inf, dtype=v.dtype).scatter_reduce(0, rows, v.detach(), reduce="amax")
    e = torch.exp(v - m[rows])
    den = torch.zeros(n, dtype=v.dtype).index_add(0, rows, e)
    return SparseTensor(sp.indices, e / den[rows], sp.dense_shape_l)


def sparse_reshape(sp: SparseTensor, shape, name=None):
    shape = [int(s) for s in shape]
    flat = torch.zeros(sp.indices.shape[0], dtype=torch.int64)
    for d, size in enumerate(sp.dense_shape_l):
        flat = flat * size + sp.indices[:, d]
    new = []
    for size in reversed(shape):
        new.insert(0, flat % size)
        flat = flat // size
    return SparseTensor(torch.stack(new, 1), sp.values_t, shape)


def sparse_tensor_dense_matmul(sp: SparseTensor, b, name=None):
    d = _unwrap(b)
    assert len(sp.dense_shape_l) == 2 and d.dim() == 2
    rows, cols = sp.indices[:, 0], sp.indices[:, 1]
    out = torch.zeros(sp.dense_shape_l[0], d.shape[1], dtype=d.dtype).index_add(0, rows, sp.values_t.unsqueeze(1) * d[cols])
    return Tensor(out)


# ----------------------------------------------------------------------------------------------
# tf.train.AdamOptimizer
# ----------------------------------------------------------------------------------------------
class AdamOptimizer:
    """TF1 Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMAs; var -= lr_t * m / (sqrt(v) + eps), eps OUTSIDE the
    bias correction (not Kingma's epsilon-hat)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **kw):
        self.lr, self.b1, self.b2, self.eps = float(_unwrap(learning_rate)), beta1, beta2, epsilon
        self.t = 0
        self.slots: Dict[int, tuple] = {}

    def minimize(self, loss, var_list=None, **kw):
        vs = var_list or STORE.trainable()
        grads = torch.autograd.grad(_unwrap(loss), [v.t for v in vs], allow_unused=True)
        STORE.last_grads = [g if g is not None else torch.zeros_like(v.t) for g, v in zip(grads, vs)]
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        with torch.no_grad():
            for v, g in zip(vs, STORE.last_grads):
                m, s = self.slots.get(id(v), (torch.zeros_like(v.t), torch.zeros_like(v.t)))
                m = self.b1 * m + (1 - self.b1) * g
                s = self.b2 * s + (1 - self.b2) * g * g
                self.slots[id(v)] = (m, s)
                v.t -= lr_t * m / (s.sqrt() + self.eps)
        return "train_op"


# ----------------------------------------------------------------------------------------------
# module assembly
# ----------------------------------------------------------------------------------------------
def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install():
    """Install the shim as ``tensorflow`` in sys.modules (and stub the one scipy module path that
    utils/process.py:5 imports and current scipy no longer has).  Idempotent."""
    tf = _module(
        "tensorflow",
        __version__="1.x-shim",
        Tensor=Tensor, Variable=Variable, SparseTensor=SparseTensor,
        float32=float32, float64=float64, int32=int32, int64=int64, bool=bool_,
        transpose=transpose, matmul=matmul, tensordot=tensordot, concat=concat, squeeze=squeeze,
        expand_dims=expand_dims, add_n=add_n, reduce_sum=reduce_sum, reduce_mean=reduce_mean, tanh=tanh,
        cast=cast, equal=equal, argmax=argmax, reshape=reshape, multiply=multiply, random_normal=random_normal,
        constant=constant, name_scope=name_scope, trainable_variables=trainable_variables,
        sparse_add=sparse_add, sparse_softmax=sparse_softmax, sparse_reshape=sparse_reshape,
        sparse_tensor_dense_matmul=sparse_tensor_dense_matmul,
    )
    tf.layers = _module("tensorflow.layers", conv1d=_conv1d, dense=_dense)
    tf.nn = _module("tensorflow.nn", dropout=_dropout, softmax=_softmax, leaky_relu=_leaky_relu, elu=_elu,
                    l2_loss=_l2_loss, softmax_cross_entropy_with_logits=_softmax_xent,
                    sparse_softmax_cross_entropy_with_logits=_sparse_softmax_xent,
                    sigmoid_cross_entropy_with_logits=_sigmoid_xent)
    tf.contrib = _module("tensorflow.contrib")
    tf.contrib.layers = _module("tensorflow.contrib.layers", bias_add=_contrib_bias_add)
    tf.train = _module("tensorflow.train", AdamOptimizer=AdamOptimizer)
    tf.array_ops = _module("tensorflow.array_ops", transpose=transpose)
    sys.modules["tensorflow"] = tf
    for sub in ("layers", "nn", "contrib", "train"):
        sys.modules[f"tensorflow.{sub}"] = getattr(tf, sub)
    sys.modules["tensorflow.contrib.layers"] = tf.contrib.layers
    # utils/process.py:5 ``from scipy.sparse.linalg.eigen.arpack import eigsh`` -- that private path is gone
    # from scipy >= 1.8; adj_to_bias (process.py:14-25) never touches it.  Stub the module path only.
    try:
        import scipy.sparse.linalg.eigen.arpack  # noqa: F401
    except Exception:
        from scipy.sparse.linalg import eigsh
        eigen = _module("scipy.sparse.linalg.eigen")
        arpack = _module("scipy.sparse.linalg.eigen.arpack", eigsh=eigsh)
        eigen.arpack = arpack
        sys.modules["scipy.sparse.linalg.eigen"] = eigen
        sys.modules["scipy.sparse.linalg.eigen.arpack"] = arpack
    return tf


def import_reference(root: str = "/root/reference"):
    """Import the reference's own modules, unmodified, from where they lie.  Returns
    (utils.layers, utils.process, models.gat, models.base_gattn)."""
    import importlib
    install()
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ("utils", "utils.layers", "utils.process", "models", "models.base_gattn", "models.gat"):
        sys.modules.pop(name, None)
    layers = importlib.import_module("utils.layers")
    process = importlib.import_module("utils.process")
    gat = importlib.import_module("models.gat")
    base = importlib.import_module("models.base_gattn")
    for m in (layers, process, gat, base):
        assert m.__file__.startswith(root), m.__file__
    return layers, process, gat, base
