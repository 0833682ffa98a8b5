"""`tensorflow` -- a minimal stand-in for the TensorFlow-1 graph API used by JamesBrofos/Stein.

NOT TensorFlow.  The reference builds its models as TF1 graphs (examples/*/main.py: placeholders
and variables under tf.variable_scope("model"), a scalar `log_p` tensor) and hands `log_p` to
SteinSampler, which differentiates it with tf.gradients once per particle
(stein/samplers/abstract_stein_sampler.py:49-55, stein/samplers/stein_sampler.py:59-68).
TensorFlow 1.12 cannot be installed here, so this package records the same graph with the same
names and lets PyTorch evaluate it: `stein_b200.log_p.GraphLogPosterior` turns a recorded `log_p`
into the batched score function of the sampler (torch.func.vmap(grad(.)) over the particles, on
the GPU).  It covers the symbol surface of SURVEY.md section A.5 -- what the reference library
and its three examples touch -- and nothing else.

Put the directory that holds this package (`compat/`) on PYTHONPATH to run the reference's
example scripts unchanged; it is deliberately not importable otherwise, so that it can never
shadow a real TensorFlow by accident.
"""
import collections
import contextlib

import numpy as np

__version__ = "1.12.0-stein-b200-shim"

float16, float32, float64, int32, int64 = "float16", "float32", "float64", "int32", "int64"
_NP_DTYPES = {"float16": np.float32, "float32": np.float32, "float64": np.float32,
              "int32": np.int64, "int64": np.int64, None: np.float32}


class GraphKeys:
    GLOBAL_VARIABLES = "variables"
    TRAINABLE_VARIABLES = "trainable_variables"


class Graph:
    """Name scopes, unique op names (TF1 rule: `name`, `name_1`, `name_2`, ...) and collections."""

    def __init__(self):
        self.scopes = []
        self.names = {}
        self.collections = {}
        self.nodes = 0

    def unique_name(self, base):
        full = "/".join(self.scopes + [base])
        k = self.names.get(full, 0)
        self.names[full] = k + 1
        return full if k == 0 else "%s_%d" % (full, k)

    def add_to_collection(self, key, value):
        self.collections.setdefault(key, []).append(value)

    def get_collection(self, key, scope=None):
        items = list(self.collections.get(key, []))
        if scope:
            prefix = scope.rstrip("/")
            items = [v for v in items if v.name == prefix or v.name.startswith(prefix + "/")
                     or v.name.startswith(prefix + ":")]
        return items


_default_graph = Graph()


def get_default_graph():
    return _default_graph


def reset_default_graph():
    global _default_graph
    _default_graph = Graph()


@contextlib.contextmanager
def variable_scope(name, *args, **kwargs):
    g = get_default_graph()
    g.scopes.append(str(name))
    try:
        yield name
    finally:
        g.scopes.pop()


name_scope = variable_scope


def get_collection(key, scope=None):
    return get_default_graph().get_collection(key, scope)


# ---- static shapes ---------------------------------------------------------------------------
class Dimension:
    def __init__(self, value):
        self.value = None if value is None else int(value)

    def __int__(self):
        return self.value

    def __eq__(self, other):
        return self.value == (other.value if isinstance(other, Dimension) else other)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return "Dimension(%r)" % (self.value,)


class TensorShape:
    def __init__(self, dims):
        self._dims = None if dims is None else [d if isinstance(d, Dimension) else Dimension(d) for d in dims]

    @property
    def dims(self):
        return self._dims

    @property
    def ndims(self):
        return None if self._dims is None else len(self._dims)

    def as_list(self):
        if self._dims is None:
            raise ValueError("as_list() is not defined on an unknown TensorShape.")
        return [d.value for d in self._dims]

    def __getitem__(self, i):
        if self._dims is None:
            raise ValueError("unknown TensorShape")
        return self._dims[i]

    def __len__(self):
        if self._dims is None:
            raise ValueError("unknown TensorShape")
        return len(self._dims)

    def __iter__(self):
        return iter(self._dims or [])

    def __repr__(self):
        return "TensorShape(%r)" % (None if self._dims is None else self.as_list(),)


def _broadcast_shape(a, b):
    if a is None or b is None:
        return None
    out = []
    for x, y in zip(([1] * (len(b) - len(a)) + list(a)), ([1] * (len(a) - len(b)) + list(b))):
        if x == 1:
            out.append(y)
        elif y == 1 or x == y:
            out.append(x)
        elif x is None or y is None:
            out.append(None)
        else:
            raise ValueError("Dimensions must be equal, but are %r and %r" % (x, y))
    return out


# ---- graph nodes -----------------------------------------------------------------------------
class Tensor:
    """A node of the recorded graph: `op` applied to `inputs` (Tensors) with `attrs`."""

    _stein_graph_node = True
    __array_priority__ = 100        # numpy scalars/arrays defer to the reflected operators

    def __init__(self, op, inputs=(), attrs=None, shape=None, base_name=None):
        self.graph = get_default_graph()
        self.op_type = op
        self.inputs = tuple(inputs)
        self.attrs = dict(attrs or {})
        self._shape = None if shape is None else list(shape)
        self.name = self.graph.unique_name(base_name or op) + ":0"
        self.graph.nodes += 1

    # static shape
    def get_shape(self):
        return TensorShape(self._shape)

    @property
    def shape(self):
        return TensorShape(self._shape)

    @property
    def dtype(self):
        return float32

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def __repr__(self):
        return "<tf.Tensor '%s' shape=%s>" % (self.name, self._shape)

    def __bool__(self):
        raise TypeError("a graph tensor has no truth value; evaluate it with Session.run")

    # arithmetic
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return subtract(self, o)
    def __rsub__(self, o): return subtract(o, self)
    def __mul__(self, o): return multiply(self, o)
    def __rmul__(self, o): return multiply(o, self)
    def __truediv__(self, o): return divide(self, o)
    def __rtruediv__(self, o): return divide(o, self)
    def __neg__(self): return negative(self)
    def __pow__(self, o): return pow(self, o)
    def __matmul__(self, o): return matmul(self, o)

    def __getitem__(self, key):
        keys = key if isinstance(key, tuple) else (key,)
        shape = None
        if self._shape is not None and len(keys) <= len(self._shape) and all(
                isinstance(k, (int, slice)) for k in keys):
            shape = []
            for k, dim in zip(keys, self._shape):
                if isinstance(k, slice):
                    shape.append(None if dim is None else len(range(*k.indices(dim))))
            shape += self._shape[len(keys):]
        return Tensor("StridedSlice", (self,), {"key": keys}, shape=shape)

    def eval(self, feed_dict=None, session=None):
        return (session or Session()).run(self, feed_dict)


class Variable(Tensor):
    """tf.Variable(initial_value): named `<scope>/Variable[_k]:0`, registered as trainable."""

    def __init__(self, initial_value, trainable=True, name=None, dtype=None):
        if isinstance(initial_value, Tensor):
            initial_value = _eval_constant(initial_value)
        value = np.array(initial_value, dtype=np.float32)
        super().__init__("Variable", shape=value.shape, base_name=name or "Variable")
        self.initial_value = value
        self.value = value.copy()           # what Session.run sees
        self.trainable = bool(trainable)
        self.graph.add_to_collection(GraphKeys.GLOBAL_VARIABLES, self)
        if trainable:
            self.graph.add_to_collection(GraphKeys.TRAINABLE_VARIABLES, self)

    @property
    def size(self):
        return int(np.prod(self._shape, dtype=np.int64)) if self._shape else 1

    def load(self, value, session=None):
        self.value = np.array(value, dtype=np.float32).reshape(self._shape)

    def __repr__(self):
        return "<tf.Variable '%s' shape=%s>" % (self.name, tuple(self._shape))


def placeholder(dtype=float32, shape=None, name=None):
    t = Tensor("Placeholder", shape=None if shape is None else list(shape), base_name=name or "Placeholder")
    t.attrs["dtype"] = dtype
    return t


def _const(value, dtype=None):
    """Python / NumPy value -> Const node: floating values become float32 (the graphs of the
    reference are float32 throughout), integers stay integers; Tensors pass through."""
    if isinstance(value, Tensor):
        return value
    arr = np.asarray(value)
    if dtype is not None:
        arr = arr.astype(_NP_DTYPES.get(dtype, np.float32))
    elif arr.dtype.kind == "f":
        arr = arr.astype(np.float32)
    elif arr.dtype.kind not in "iub":
        raise TypeError("cannot convert %r to a tensor" % (value,))
    return Tensor("Const", attrs={"value": arr}, shape=arr.shape)


constant = _const
convert_to_tensor = _const


def zeros(shape, dtype=float32, name=None):
    shape = [int(s) for s in shape]
    return Tensor("Const", attrs={"value": np.zeros(shape, np.float32)}, shape=shape, base_name=name or "zeros")


def ones(shape, dtype=float32, name=None):
    shape = [int(s) for s in shape]
    return Tensor("Const", attrs={"value": np.ones(shape, np.float32)}, shape=shape, base_name=name or "ones")


def zeros_like(x, name=None):
    return multiply(_const(x), 0.0)


def ones_like(x, name=None):
    return add(multiply(_const(x), 0.0), 1.0)


def _binary(op, a, b):
    a, b = _const(a), _const(b)
    try:
        shape = _broadcast_shape(a._shape, b._shape)
    except ValueError:
        raise
    return Tensor(op, (a, b), shape=shape)


def add(a, b, name=None): return _binary("Add", a, b)
def subtract(a, b, name=None): return _binary("Sub", a, b)
def multiply(a, b, name=None): return _binary("Mul", a, b)
def divide(a, b, name=None): return _binary("RealDiv", a, b)
def maximum(a, b, name=None): return _binary("Maximum", a, b)
def minimum(a, b, name=None): return _binary("Minimum", a, b)
def pow(a, b, name=None): return _binary("Pow", a, b)  # noqa: A001
div = truediv = divide


def _unary(op, x):
    x = _const(x)
    return Tensor(op, (x,), shape=x._shape)


def negative(x, name=None): return _unary("Neg", x)
def square(x, name=None): return _unary("Square", x)
def sqrt(x, name=None): return _unary("Sqrt", x)
def exp(x, name=None): return _unary("Exp", x)
def log(x, name=None): return _unary("Log", x)
def abs(x, name=None): return _unary("Abs", x)  # noqa: A001
def tanh(x, name=None): return _unary("Tanh", x)
def sigmoid(x, name=None): return _unary("Sigmoid", x)
def reciprocal(x, name=None): return _unary("Reciprocal", x)
def lgamma(x, name=None): return _unary("Lgamma", x)
def identity(x, name=None): return _unary("Identity", x)
def stop_gradient(x, name=None): return _unary("StopGradient", x)
def cast(x, dtype, name=None): return _unary("Identity", x)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _const(a), _const(b)
    shape = None
    if a._shape is not None and b._shape is not None and len(a._shape) == 2 and len(b._shape) == 2:
        m = a._shape[1] if transpose_a else a._shape[0]
        ka = a._shape[0] if transpose_a else a._shape[1]
        kb = b._shape[1] if transpose_b else b._shape[0]
        n = b._shape[0] if transpose_b else b._shape[1]
        if ka is not None and kb is not None and ka != kb:
            raise ValueError("Dimensions must be equal, but are %d and %d for 'MatMul'" % (ka, kb))
        shape = [m, n]
    return Tensor("MatMul", (a, b), {"ta": bool(transpose_a), "tb": bool(transpose_b)}, shape=shape)


def transpose(x, perm=None, name=None):
    x = _const(x)
    shape = None
    if x._shape is not None:
        p = list(perm) if perm is not None else list(range(len(x._shape)))[::-1]
        shape = [x._shape[i] for i in p]
    return Tensor("Transpose", (x,), {"perm": None if perm is None else [int(i) for i in perm]}, shape=shape)


def reshape(x, shape, name=None):
    x = _const(x)
    shape = [int(s) for s in shape]
    static = [None if s < 0 else s for s in shape]
    # like TensorFlow: a single -1 is inferred when the element count of the input is known
    if static.count(None) == 1 and x._shape is not None and all(d is not None for d in x._shape):
        total = int(np.prod(x._shape, dtype=np.int64)) if x._shape else 1
        rest = int(np.prod([s for s in static if s is not None], dtype=np.int64)) if len(static) > 1 else 1
        if rest > 0 and total % rest == 0:
            static[static.index(None)] = total // rest
    return Tensor("Reshape", (x,), {"shape": shape}, shape=static)


def squeeze(x, axis=None, name=None):
    return Tensor("Squeeze", (_const(x),), {"axis": axis})


def expand_dims(x, axis, name=None):
    return Tensor("ExpandDims", (_const(x),), {"axis": int(axis)})


def stack(values, axis=0, name=None):
    vals = [_const(v) for v in values]
    shape = None
    if vals and all(v._shape is not None for v in vals):
        shape = list(vals[0]._shape)
        shape.insert(axis if axis >= 0 else len(shape) + 1 + axis, len(vals))
    return Tensor("Stack", vals, {"axis": int(axis)}, shape=shape)


def concat(values, axis, name=None):
    return Tensor("Concat", [_const(v) for v in values], {"axis": int(axis)})


def _reduce(op, x, axis, keepdims):
    x = _const(x)
    if isinstance(axis, (list, tuple)):
        axis = [int(a) for a in axis]
    elif axis is not None:
        axis = [int(axis)]
    shape = None
    if x._shape is not None:
        nd = len(x._shape)
        ax = list(range(nd)) if axis is None else [a % nd for a in axis]
        shape = [(1 if i in ax else s) for i, s in enumerate(x._shape) if keepdims or i not in ax]
    return Tensor(op, (x,), {"axis": axis, "keepdims": bool(keepdims)}, shape=shape)


def reduce_sum(x, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    return _reduce("Sum", x, axis if axis is not None else reduction_indices,
                   keepdims if keep_dims is None else keep_dims)


def reduce_mean(x, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    return _reduce("Mean", x, axis if axis is not None else reduction_indices,
                   keepdims if keep_dims is None else keep_dims)


def reduce_max(x, axis=None, keepdims=False, name=None, reduction_indices=None, keep_dims=None):
    return _reduce("Max", x, axis if axis is not None else reduction_indices,
                   keepdims if keep_dims is None else keep_dims)


def gradients(ys, xs, name=None, stop_gradients=None):
    """d(sum of ys)/d(x) for every x, as graph tensors (evaluated with torch.autograd)."""
    ys = list(ys) if isinstance(ys, (list, tuple)) else [ys]
    xs = list(xs) if isinstance(xs, (list, tuple)) else [xs]
    return [Tensor("Gradient", [_const(y) for y in ys] + [x], {"n_ys": len(ys)}, shape=x._shape) for x in xs]


TopKV2 = collections.namedtuple("TopKV2", ["values", "indices"])


class _NN:
    @staticmethod
    def relu(x, name=None): return _unary("Relu", x)

    @staticmethod
    def sigmoid(x, name=None): return _unary("Sigmoid", x)

    @staticmethod
    def tanh(x, name=None): return _unary("Tanh", x)

    @staticmethod
    def softplus(x, name=None): return _unary("Softplus", x)

    @staticmethod
    def sigmoid_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
        if _sentinel is not None or labels is None or logits is None:
            raise ValueError("Only call `sigmoid_cross_entropy_with_logits` with named arguments (labels=..., logits=...)")
        return _binary("SigmoidCrossEntropyWithLogits", labels, logits)

    @staticmethod
    def top_k(x, k=1, sorted=True, name=None):  # noqa: A002
        x = _const(x)
        shape = None if x._shape is None else list(x._shape[:-1]) + [int(k)]
        values = Tensor("TopKValues", (x,), {"k": int(k)}, shape=shape)
        indices = Tensor("TopKIndices", (x,), {"k": int(k)}, shape=shape)
        return TopKV2(values, indices)


nn = _NN()


# ---- evaluation with PyTorch -------------------------------------------------------------------
def evaluate(fetches, values, device=None):
    """Evaluate graph tensors with PyTorch.  `values` maps Variables / placeholders to torch
    tensors (or anything torch.as_tensor takes); constants are created on `device` in float32.
    Differentiable end to end and safe under torch.func transforms (no data-dependent control
    flow, no in-place updates), which is how the sampler computes all particles' scores at once."""
    import torch
    single = isinstance(fetches, Tensor)
    cache = {}
    grad_nodes = [f for f in ([fetches] if single else fetches)
                  if isinstance(f, Tensor) and f.op_type == "Gradient"]

    def const(arr):
        dt = torch.float32 if arr.dtype.kind == "f" else torch.int64
        return torch.as_tensor(arr, dtype=dt, device=device)

    def ev(t):
        key = id(t)
        if key in cache:
            return cache[key]
        if t in values:
            out = values[t]
            if not isinstance(out, torch.Tensor):
                out = torch.as_tensor(np.asarray(out, dtype=np.float32), device=device)
        else:
            out = run_op(t)
        cache[key] = out
        return out

    def run_op(t):
        op, a = t.op_type, t.attrs
        if op == "Const":
            return const(a["value"])
        if op == "Variable":
            return const(t.value)
        if op == "Placeholder":
            raise ValueError("You must feed a value for placeholder tensor '%s'" % t.name)
        if op == "Gradient":
            return grad_op(t)
        x = [ev(i) for i in t.inputs]
        if op == "Add": return x[0] + x[1]
        if op == "Sub": return x[0] - x[1]
        if op == "Mul": return x[0] * x[1]
        if op == "RealDiv": return x[0] / x[1]
        if op == "Maximum": return torch.maximum(x[0], x[1])
        if op == "Minimum": return torch.minimum(x[0], x[1])
        if op == "Pow": return torch.pow(x[0], x[1])
        if op == "Neg": return -x[0]
        if op == "Square": return x[0] * x[0]
        if op == "Sqrt": return torch.sqrt(x[0])
        if op == "Exp": return torch.exp(x[0])
        if op == "Log": return torch.log(x[0])
        if op == "Abs": return torch.abs(x[0])
        if op == "Tanh": return torch.tanh(x[0])
        if op == "Sigmoid": return torch.sigmoid(x[0])
        if op == "Reciprocal": return torch.reciprocal(x[0])
        if op == "Lgamma": return torch.lgamma(x[0])
        if op == "Identity": return x[0]
        if op == "StopGradient": return x[0].detach()
        if op == "Relu": return torch.relu(x[0])
        if op == "Softplus": return torch.nn.functional.softplus(x[0])
        if op == "SigmoidCrossEntropyWithLogits":
            z, l = x[0], x[1]      # max(l, 0) - l z + log(1 + exp(-|l|)), TF's stable form
            return torch.relu(l) - l * z + torch.log1p(torch.exp(-torch.abs(l)))
        if op == "MatMul":
            p, q = x[0], x[1]
            if a["ta"]: p = p.transpose(-1, -2)
            if a["tb"]: q = q.transpose(-1, -2)
            return torch.matmul(p, q)
        if op == "Transpose":
            return x[0].permute(*(a["perm"] if a["perm"] is not None else range(x[0].dim())[::-1]))
        if op == "Reshape": return x[0].reshape(a["shape"])
        if op == "Squeeze":
            if a["axis"] is None: return x[0].squeeze()
            ax = a["axis"] if isinstance(a["axis"], (list, tuple)) else [a["axis"]]
            out = x[0]
            for k in sorted((i % out.dim() for i in ax), reverse=True):
                out = out.squeeze(k)
            return out
        if op == "ExpandDims": return x[0].unsqueeze(a["axis"])
        if op == "Stack": return torch.stack(x, dim=a["axis"])
        if op == "Concat": return torch.cat(x, dim=a["axis"])
        if op in ("Sum", "Mean", "Max"):
            ax, kd = a["axis"], a["keepdims"]
            if ax is None:
                ax = list(range(x[0].dim()))
            if not ax:
                return x[0]
            if op == "Sum": return x[0].sum(dim=ax, keepdim=kd)
            if op == "Mean": return x[0].mean(dim=ax, keepdim=kd)
            return torch.amax(x[0], dim=ax, keepdim=kd)
        if op == "StridedSlice": return x[0][a["key"]]
        if op == "TopKValues": return torch.topk(x[0], a["k"]).values
        if op == "TopKIndices": return torch.topk(x[0], a["k"]).indices
        raise NotImplementedError("tensorflow shim: op %r" % op)

    def grad_op(t):
        # Independent evaluation of the ys with the differentiated inputs as leaves.  All
        # gradient nodes of this call that share their ys (tf.gradients(y, [x1, ..., xn])) are
        # served by ONE backward pass.
        n_ys = t.attrs["n_ys"]
        ys = t.inputs[:n_ys]
        group = [g for g in grad_nodes if g.inputs[:g.attrs["n_ys"]] == ys] or [t]
        if t not in group:
            group.append(t)
        xs = [g.inputs[g.attrs["n_ys"]] for g in group]
        leaves = [ev(x).detach().clone().requires_grad_(True) for x in xs]
        sub = dict(values)
        sub.update(zip(xs, leaves))
        with torch.enable_grad():
            outs = evaluate(list(ys), sub, device=device)
            total = sum(o.sum() for o in outs)
            if total.requires_grad:
                grads = torch.autograd.grad(total, leaves, allow_unused=True)
            else:
                grads = [None] * len(leaves)
        for g, leaf, val in zip(group, leaves, grads):
            cache[id(g)] = torch.zeros_like(leaf) if val is None else val
        return cache[id(t)]

    out = [ev(_const(f)) for f in ([fetches] if single else fetches)]
    return out[0] if single else out


def _eval_constant(t):
    return evaluate(t, {}).detach().cpu().numpy()


class Session:
    """tf.Session stand-in: `run` evaluates on the CPU with the variables' current values."""

    def __init__(self, *args, **kwargs):
        self.graph = get_default_graph()

    def run(self, fetches, feed_dict=None):
        values = {}
        for k, v in (feed_dict or {}).items():
            values[k] = np.asarray(v, dtype=np.float32)
        flat = []

        def collect(f):
            if isinstance(f, (list, tuple)):
                for g in f:
                    collect(g)
            elif f is not None:
                flat.append(_const(f))

        collect(fetches)
        outs = iter(evaluate(flat, values)) if flat else iter(())

        def rebuild(f):
            if isinstance(f, (list, tuple)):
                out = [rebuild(g) for g in f]
                return out if isinstance(f, list) else (type(f)(*out) if hasattr(f, "_fields") else tuple(out))
            if f is None:
                return None
            return next(outs).detach().cpu().numpy()

        return rebuild(fetches)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


InteractiveSession = Session


def global_variables_initializer():
    return None


def trainable_variables(scope=None):
    return get_collection(GraphKeys.TRAINABLE_VARIABLES, scope)


from . import contrib  # noqa: E402,F401
