"""Model-side objects the sampler consumes in place of a TensorFlow graph.

The reference takes `log_p`, a TF1 tensor built under tf.variable_scope("model")
(examples/*/main.py), finds the trainable variables of that scope
(stein/samplers/abstract_stein_sampler.py:49-51) and differentiates log_p with
tf.gradients (:55), one particle at a time (stein/samplers/stein_sampler.py:59-68).
Without TensorFlow the same roles are played by:

  Variable     -- named, shaped slot (`.name`, `.get_shape().as_list()`), the
                  keys of `sampler.theta` and of the dict<->array converters;
  Placeholder  -- key of the feed dictionaries;
  LogPosterior -- owns the variables and knows how to produce the score matrix
                  S = grad_theta log p for ALL particles at once on the GPU.
"""
import numpy as np


class _Shape:
    def __init__(self, dims):
        self._dims = [int(x) for x in dims]

    def as_list(self):
        return list(self._dims)


class Variable:
    """Stand-in for tf.Variable: hashable, `.name` follows TF1's auto-naming
    ("model/Variable:0", "model/Variable_1:0", ...) so that the name sort of
    stein/utilities/converters.py:40 reproduces the reference's column order."""

    def __init__(self, name, shape):
        self.name = name
        self._shape = _Shape(shape)

    def get_shape(self):
        return self._shape

    @property
    def size(self):
        return int(np.prod(self._shape.as_list(), dtype=np.int64)) if self._shape.as_list() else 1

    def __repr__(self):
        return "<Variable %s shape=%s>" % (self.name, self._shape.as_list())


class Placeholder:
    def __init__(self, name, shape):
        self.name = name
        self.shape = tuple(shape)

    def __repr__(self):
        return "<Placeholder %s shape=%s>" % (self.name, self.shape)


class Output:
    """Symbolic handle on a tensor of the model that `function_posterior` can
    evaluate for every particle (e.g. logits, predictions)."""

    def __init__(self, model, name):
        self.model = model
        self.name = name

    def __repr__(self):
        return "<Output %s of %s>" % (self.name, type(self.model).__name__)


class LogPosterior:
    """Base class: variable bookkeeping + the hooks the sampler calls."""

    scope = "model"

    def __init__(self):
        self.model_vars = []
        self._n_placeholders = 0

    # -- graph-construction helpers -------------------------------------------
    def _variable(self, shape):
        k = len(self.model_vars)
        name = "%s/Variable%s:0" % (self.scope, "" if k == 0 else "_%d" % k)
        v = Variable(name, shape)
        self.model_vars.append(v)
        return v

    def _placeholder(self, shape):
        k = self._n_placeholders
        self._n_placeholders += 1
        return Placeholder("%s/Placeholder%s:0" % (self.scope, "" if k == 0 else "_%d" % k), shape)

    # -- what the sampler needs --------------------------------------------------
    @property
    def log_p(self):
        """What the examples hand to SteinSampler(n_particles, log_p, gd)."""
        return self

    @property
    def sorted_vars(self):
        return sorted(self.model_vars, key=lambda v: v.name)

    @property
    def n_params(self):
        return sum(v.size for v in self.model_vars)

    def column_slices(self):
        """{variable: (start, stop)} in the flat layout (converters.py:40-53)."""
        out, start = {}, 0
        for v in self.sorted_vars:
            out[v] = (start, start + v.size)
            start += v.size
        return out

    def scores(self, engine, batch_feed):
        """Write S = grad log p (all local particles) into engine.scores_dev."""
        raise NotImplementedError()

    def evaluate(self, output, engine, feed_dict):
        """Return a device tensor (n_local x M) with `output` for every particle."""
        raise NotImplementedError()


def feed_array(feed, placeholder, what):
    if placeholder not in feed:
        raise KeyError("feed dictionary lacks the %s placeholder %r" % (what, placeholder))
    return np.asarray(feed[placeholder])
