import numpy as np

from ..runtime import ptr
from .base import LogPosterior, Output, feed_array


class LinearRegression(LogPosterior):
    """Bayesian linear regression, unit-variance Gaussian likelihood, N(0,1) prior.

    Mirrors the graph of examples/linear_regression/main.py:18-31:
        y_hat = X w;  log_l = -0.5 sum (y_hat - y)^2;  log_p = log_l + sum N(w; 0, 1)
    Score for all particles: S = -(W X^T - y^T) X - W, computed by
    stein_score_linear (CUDA).  Full batch, no N/B rescale (as in the example).
    """

    def __init__(self, n_feats):
        super().__init__()
        self.n_feats = int(n_feats)
        self.X = self._placeholder([None, self.n_feats])       # model_X  (:20)
        self.y = self._placeholder([None, 1])                  # model_y  (:21)
        self.w = self._variable([self.n_feats, 1])             # model_w  (:22)
        self.y_hat = Output(self, "y_hat")                     # (:28)

    def scores(self, engine, batch_feed):
        ctx = engine.ctx
        Xd = ctx.dense(feed_array(batch_feed, self.X, "feature").reshape(-1, self.n_feats))
        yd = ctx.dense(feed_array(batch_feed, self.y, "target").reshape(-1))
        if yd.numel() != Xd.shape[0]:
            raise ValueError("X has %d rows but y has %d" % (Xd.shape[0], yd.numel()))
        ctx.check(ctx.lib.stein_score_linear(ctx.handle, ptr(engine.particles_dev), engine.n_local,
                                             self.n_feats, engine.ld, ptr(Xd), ptr(yd), Xd.shape[0],
                                             ptr(engine.scores_dev)))

    def evaluate(self, output, engine, feed_dict):
        import torch
        ctx = engine.ctx
        Xt = ctx.dense(feed_array(feed_dict, self.X, "feature").reshape(-1, self.n_feats))
        out = torch.empty((engine.n_local, Xt.shape[0]), dtype=torch.float32, device=Xt.device)
        ctx.check(ctx.lib.stein_predict_linear(ctx.handle, ptr(engine.particles_dev), engine.n_local,
                                               self.n_feats, engine.ld, ptr(Xt), Xt.shape[0], ptr(out)))
        return out
