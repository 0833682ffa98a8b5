"""log_p given as a recorded TF1-style graph tensor (the reference's own model definitions).

The reference's examples build `log_p` with the TensorFlow-1 graph API and pass the tensor to
SteinSampler(n_particles, log_p, gd) (examples/*/main.py; stein/samplers/stein_sampler.py:16).
With the stand-in `tensorflow` package of compat/ on the path, the same scripts record the same
graph; this class is what the sampler makes of such a tensor:

  * model variables = trainable variables under scope "model", in creation order
    (stein/samplers/abstract_stein_sampler.py:49-51); the flat column order is the name sort of
    stein/utilities/converters.py:40 -- the shim names variables the TF1 way;
  * scores: torch.func.vmap(grad(log_p)) over the particle axis, one batched evaluation on the
    GPU in place of the n sequential sess.run calls of stein/samplers/stein_sampler.py:59-68;
  * function_posterior: the same graph evaluated for every particle
    (stein/samplers/abstract_stein_sampler.py:129-168);
  * graph recognition: when the graph BEHAVES like one of the reference's three example models
    (examples/linear_regression, logistic_regression, regression_neural_network: recognised by the
    shapes of its variables and placeholders and then VERIFIED numerically -- the closed-form CUDA
    score kernel must reproduce the autograd scores of the graph on the first particles), the
    scores come from stein_score_linear / _logistic / _bnn instead of torch autograd.  The
    training-set size baked into the graph as a constant is recovered from the scores themselves
    (they are affine in n_train resp. 1 / n_train).  Anything that does not verify stays on autograd.
"""
import os

import numpy as np

from .base import LogPosterior


def is_graph_tensor(obj):
    return bool(getattr(obj, "_stein_graph_node", False))


class GraphLogPosterior(LogPosterior):
    def __init__(self, log_p):
        super().__init__()
        if not is_graph_tensor(log_p):
            raise TypeError("GraphLogPosterior needs a tensor of the compat `tensorflow` package")
        import sys
        tf = sys.modules[type(log_p).__module__]      # the stand-in that recorded `log_p`
        self._tf = tf
        self.tensor = log_p
        self.model_vars = list(log_p.graph.get_collection(tf.GraphKeys.TRAINABLE_VARIABLES, self.scope))
        if not self.model_vars:
            raise ValueError('no trainable variable under tf.variable_scope("model") '
                             "(stein/samplers/abstract_stein_sampler.py:49-51 looks there)")
        self._fast = None          # None: not tried yet; False: stays on autograd; else (model, gX, gy)
        self.recognised = None     # "linear" | "logistic" | "bnn" once a closed-form kernel took over

    # -- graph evaluation (device agnostic: the tests run it on the CPU) ---------------------
    def _values(self, flat, feed):
        vals = dict(feed)
        for v, (a, b) in self.column_slices().items():
            vals[v] = flat[a:b].reshape(v.get_shape().as_list())
        return vals

    def _feed_tensors(self, feed_dict, to_device):
        feed = {}
        for k, val in (feed_dict or {}).items():
            if not is_graph_tensor(k):
                raise KeyError("feed dictionary key %r is not a placeholder of the graph" % (k,))
            feed[k] = to_device(np.ascontiguousarray(np.asarray(val, dtype=np.float32)))
        return feed

    def graph_scores(self, theta, feed):
        """theta: (n, n_params) torch tensor; feed: {placeholder: torch tensor} -> (n, n_params)."""
        from torch.func import grad, vmap
        tf = self._tf

        def one(flat):
            return tf.evaluate(self.tensor, self._values(flat, feed), device=flat.device).reshape(())

        return vmap(grad(one))(theta)

    def graph_values(self, tensor, theta, feed):
        """`tensor` for every particle, flattened: (n, size)."""
        from torch.func import vmap
        tf = self._tf
        return vmap(lambda flat: tf.evaluate(tensor, self._values(flat, feed), device=flat.device).reshape(-1))(theta)

    # -- graph recognition ------------------------------------------------------------------------
    def _candidates(self, batch_feed):
        """(kind, model, graph placeholder of X, of y) for every reading of the graph's signature as one
        of the three example models."""
        from .linear_regression import LinearRegression
        from .logistic_regression import LogisticRegression
        from .regression_neural_network import RegressionNeuralNetwork
        shapes = [v.get_shape().as_list() for v in self.sorted_vars]
        feeds = [(k, np.asarray(v)) for k, v in (batch_feed or {}).items()]
        if len(feeds) != 2 or any(a.ndim not in (1, 2) for _, a in feeds):
            return
        kind = F = H = None
        if len(shapes) == 1 and shapes[0] and (len(shapes[0]) == 1 or shapes[0][1:] == [1]):
            kind, F = "linear", shapes[0][0]
        elif len(shapes) == 2 and shapes[0] and shapes[0][1:] in ([], [1]) and shapes[1] == []:
            kind, F = "logistic", shapes[0][0]
        elif (len(shapes) == 6 and shapes[0] == [] and shapes[1] == [] and len(shapes[2]) == 2 and shapes[5] == []
              and shapes[3] == [shapes[2][1]] and shapes[4] in ([shapes[2][1], 1], [shapes[2][1]])):
            kind, (F, H) = "bnn", shapes[2]
        if kind is None:
            return
        for (gx, ax), (gy, ay) in (feeds, feeds[::-1]):
            cols_x = ax.shape[1] if ax.ndim == 2 else 1
            cols_y = ay.shape[1] if ay.ndim == 2 else 1
            if cols_x != F or cols_y != 1 or ax.shape[0] != ay.shape[0]:
                continue
            if kind == "linear":
                yield kind, LinearRegression(F), gx, gy
            elif kind == "logistic":
                yield kind, LogisticRegression(F, 1.0), gx, gy
            else:
                yield kind, RegressionNeuralNetwork(F, H, 1.0), gx, gy

    def _recognise(self, engine, batch_feed, S_graph, m):
        """Try the closed-form kernels against the autograd scores S_graph of the first m particles."""
        import torch
        if os.environ.get("STEIN_GRAPH_RECOGNITION", "1") == "0":
            return False
        d = self.n_params
        ref = S_graph[:m].double()
        colscale = ref.abs().amax(dim=0).clamp_min(1e-30)

        def kernel_scores(model, gx, gy, n_train=None):
            if n_train is not None:
                model.n_train = float(n_train)
            model.scores(engine, {model.X: batch_feed[gx], model.y: batch_feed[gy]})
            return engine.scores_dev[:m, :d].double().clone()

        for kind, model, gx, gy in self._candidates(batch_feed):
            B = float(np.asarray(batch_feed[gx]).shape[0])
            if kind == "logistic":        # S = prior + (n_train / B) lik: affine in n_train
                s0, s1 = kernel_scores(model, gx, gy, 0.0), kernel_scores(model, gx, gy, B)
                dirn = s1 - s0
                t = float(((ref - s0) * dirn).sum() / (dirn * dirn).sum().clamp_min(1e-300))
                if not np.isfinite(t) or t <= 0:
                    continue
                model.n_train = float(round(t * B)) if abs(round(t * B) - t * B) <= 1e-3 * t * B else t * B
            elif kind == "bnn":           # S = lik / B + prior / n_train: affine in 1 / n_train
                sa, sb = kernel_scores(model, gx, gy, 1.0), kernel_scores(model, gx, gy, 2.0)
                prior = 2.0 * (sa - sb)
                u = float(((ref - (sa - prior)) * prior).sum() / (prior * prior).sum().clamp_min(1e-300))
                if not np.isfinite(u) or u <= 0:
                    continue
                nt = 1.0 / u
                model.n_train = float(round(nt)) if abs(round(nt) - nt) <= 1e-3 * nt else nt
            got = kernel_scores(model, gx, gy)
            if bool((((got - ref).abs() / colscale).amax() <= 1e-4).item()):
                self.recognised = kind
                # the example graphs bake n_batch in as a constant next to n_train: the closed form is only
                # known to agree for batches of the size it was verified on (the linear model has no such scale)
                return model, gx, gy, (None if kind == "linear" else int(B))
        return False

    # -- hooks of the sampler ---------------------------------------------------------------------
    def scores(self, engine, batch_feed):
        import torch
        n, d = engine.n_local, self.n_params
        if self._fast:
            model, gx, gy, batch = self._fast
            if set(batch_feed.keys()) == {gx, gy} and batch in (None, int(np.asarray(batch_feed[gx]).shape[0])):
                model.scores(engine, {model.X: batch_feed[gx], model.y: batch_feed[gy]})
                return
        feed = self._feed_tensors(batch_feed, engine.ctx.dense)
        if self._fast is None:
            # first call: autograd scores of a few particles decide whether a closed-form kernel takes over
            m = min(n, 16)
            probe = self.graph_scores(engine.particles_dev[:m, :d].clone(), feed).to(torch.float32)
            self._fast = self._recognise(engine, batch_feed, probe, m)
            if self._fast:
                return self.scores(engine, batch_feed)
        S = self.graph_scores(engine.particles_dev[:n, :d], feed)
        engine.scores_dev[:n, :d] = S.to(torch.float32)

    def evaluate(self, output, engine, feed_dict):
        n, d = engine.n_local, self.n_params
        feed = self._feed_tensors(feed_dict, engine.ctx.dense)
        return self.graph_values(output, engine.particles_dev[:n, :d], feed)
