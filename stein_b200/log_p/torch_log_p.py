import numpy as np

from .base import LogPosterior, Output


class TorchLogPosterior(LogPosterior):
    """User-defined log-posterior with PyTorch as the autograd carrier.

    Replaces `tf.gradients(log_p, model_vars)` evaluated once per particle
    (stein/samplers/abstract_stein_sampler.py:55, stein_sampler.py:59-68) by
    torch.func.vmap(grad(log_p)) over the particle axis, on the GPU.

        model = TorchLogPosterior({"w": [F, 1], "log_alpha": []}, log_p_fn)
        def log_p_fn(params, feed):   # params: {name: tensor of the declared shape}
            ...                       # feed:   {placeholder name: tensor}
            return scalar_tensor

    Variables are created in the order of `var_shapes` (dict order) and named the
    TF1 way, so the flat layout follows the reference's name sort.  Placeholders
    are created with `model.placeholder(name, shape)`.
    """

    def __init__(self, var_shapes, log_p_fn, outputs=None):
        super().__init__()
        self._log_p_fn = log_p_fn
        self.vars = {}
        for name, shape in var_shapes.items():
            self.vars[name] = self._variable(list(shape))
        self.placeholders = {}
        self._outputs = dict(outputs or {})

    def placeholder(self, name, shape):
        p = self._placeholder(list(shape))
        self.placeholders[name] = p
        return p

    def output(self, name):
        return Output(self, name)

    def _unflatten(self, flat):
        slices = self.column_slices()
        return {name: flat[slices[v][0]:slices[v][1]].reshape(v.get_shape().as_list())
                for name, v in self.vars.items()}

    def _feed_tensors(self, engine, feed):
        by_obj = {p: n for n, p in self.placeholders.items()}
        out = {}
        for k, val in feed.items():
            name = by_obj.get(k, k if isinstance(k, str) else None)
            if name is None:
                raise KeyError("unknown placeholder %r" % (k,))
            out[name] = engine.ctx.dense(np.asarray(val, dtype=np.float32))
        return out

    def scores(self, engine, batch_feed):
        import torch
        from torch.func import grad, vmap
        feed = self._feed_tensors(engine, batch_feed)
        n, d = engine.n_local, self.n_params
        theta = engine.particles_dev[:n, :d]

        def one(flat):
            return self._log_p_fn(self._unflatten(flat), feed)

        S = vmap(grad(one))(theta)
        engine.scores_dev[:n, :d] = S.to(torch.float32)

    def evaluate(self, output, engine, feed_dict):
        from torch.func import vmap
        fn = self._outputs[output.name]
        feed = self._feed_tensors(engine, feed_dict)
        n, d = engine.n_local, self.n_params
        theta = engine.particles_dev[:n, :d]
        return vmap(lambda flat: fn(self._unflatten(flat), feed).reshape(-1))(theta)
