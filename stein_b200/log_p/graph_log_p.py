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
    (stein/samplers/abstract_stein_sampler.py:129-168).
"""
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

    # -- hooks of the sampler ---------------------------------------------------------------------
    def scores(self, engine, batch_feed):
        import torch
        n, d = engine.n_local, self.n_params
        feed = self._feed_tensors(batch_feed, engine.ctx.dense)
        S = self.graph_scores(engine.particles_dev[:n, :d], feed)
        engine.scores_dev[:n, :d] = S.to(torch.float32)

    def evaluate(self, output, engine, feed_dict):
        n, d = engine.n_local, self.n_params
        feed = self._feed_tensors(feed_dict, engine.ctx.dense)
        return self.graph_values(output, engine.particles_dev[:n, :d], feed)
