from ..runtime import ptr
from .base import LogPosterior, Output, feed_array


class LogisticRegression(LogPosterior):
    """Bayesian logistic regression with a Gamma hyper-prior on the weight precision.

    Mirrors the graph of examples/logistic_regression/main.py:23-49:
        alpha = exp(log_alpha);  w ~ N(0, 1/sqrt(alpha));  alpha ~ Gamma(1, 0.01)
        log_l = -sum sigmoid_cross_entropy(labels=y, logits=X w)
        log_p = log_l * (n_train / n_batch) + sum log N(w) + log Gamma(alpha)
    (the Gamma density is evaluated at alpha with no Jacobian term, :37,:48).
    Flat layout [w (F), log_alpha]; scores by stein_score_logistic (CUDA).
    """

    def __init__(self, n_feats, n_train, prior_a=1.0, prior_b=0.01):
        super().__init__()
        self.n_feats, self.n_train = int(n_feats), float(n_train)
        self.prior_a, self.prior_b = float(prior_a), float(prior_b)
        self.X = self._placeholder([None, self.n_feats])       # model_X         (:26)
        self.y = self._placeholder([None, 1])                  # model_y         (:27)
        self.w = self._variable([self.n_feats, 1])             # model_w         (:28)
        self.log_alpha = self._variable([])                    # model_log_alpha (:29)
        self.logits = Output(self, "logits")                   # (:40)

    def scores(self, engine, batch_feed):
        ctx = engine.ctx
        Xb = ctx.dense(feed_array(batch_feed, self.X, "feature").reshape(-1, self.n_feats))
        yb = ctx.dense(feed_array(batch_feed, self.y, "label").reshape(-1))
        if yb.numel() != Xb.shape[0]:
            raise ValueError("X has %d rows but y has %d" % (Xb.shape[0], yb.numel()))
        ctx.check(ctx.lib.stein_score_logistic(
            ctx.handle, ptr(engine.particles_dev), engine.n_local, self.n_feats, engine.ld, ptr(Xb),
            ptr(yb), Xb.shape[0], self.n_train, self.prior_a, self.prior_b, ptr(engine.scores_dev)))

    def evaluate(self, output, engine, feed_dict):
        import torch
        ctx = engine.ctx
        Xt = ctx.dense(feed_array(feed_dict, self.X, "feature").reshape(-1, self.n_feats))
        out = torch.empty((engine.n_local, Xt.shape[0]), dtype=torch.float32, device=Xt.device)
        ctx.check(ctx.lib.stein_predict_linear(ctx.handle, ptr(engine.particles_dev), engine.n_local,
                                               self.n_feats, engine.ld, ptr(Xt), Xt.shape[0], ptr(out)))
        return out
