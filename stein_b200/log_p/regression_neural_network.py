from ..runtime import ptr
from .base import LogPosterior, Output, feed_array


class RegressionNeuralNetwork(LogPosterior):
    """One-hidden-layer ReLU Bayesian neural network for regression.

    Mirrors the graph of examples/regression_neural_network/main.py:29-85:
        lambda = exp(log_lambda), gamma = exp(log_gamma), both ~ Gamma(alpha, beta)
        pred = relu(X w_1 + b_1) w_2 + b_2;   y ~ N(pred, 1/sqrt(gamma))
        w_1, b_1, w_2, b_2 ~ N(0, 1/sqrt(lambda))
        log_p = (log_l * n_train / n_batch + priors) / n_train
    Variable creation order log_lambda, log_gamma, w_1, b_1, w_2, b_2 (:35-42)
    gives the flat layout [ll, lg, w_1 (F*H row-major), b_1, w_2, b_2].
    Scores by stein_score_bnn, predictions by stein_predict_bnn (CUDA).
    """

    def __init__(self, n_feats, n_hidden, n_train, alpha=1.0, beta=0.01):
        super().__init__()
        self.n_feats, self.n_hidden, self.n_train = int(n_feats), int(n_hidden), float(n_train)
        self.alpha, self.beta = float(alpha), float(beta)
        self.X = self._placeholder([None, self.n_feats])               # model_X (:31)
        self.y = self._placeholder([None, 1])                          # model_y (:32)
        self.log_lambda = self._variable([])                           # (:35)
        self.log_gamma = self._variable([])                            # (:36)
        self.w_1 = self._variable([self.n_feats, self.n_hidden])       # (:39)
        self.b_1 = self._variable([self.n_hidden])                     # (:40)
        self.w_2 = self._variable([self.n_hidden, 1])                  # (:41)
        self.b_2 = self._variable([])                                  # (:42)
        self.pred = Output(self, "pred")                               # (:46-48)

    def scores(self, engine, batch_feed):
        ctx = engine.ctx
        Xb = ctx.dense(feed_array(batch_feed, self.X, "feature").reshape(-1, self.n_feats))
        yb = ctx.dense(feed_array(batch_feed, self.y, "target").reshape(-1))
        if yb.numel() != Xb.shape[0]:
            raise ValueError("X has %d rows but y has %d" % (Xb.shape[0], yb.numel()))
        ctx.check(ctx.lib.stein_score_bnn(
            ctx.handle, ptr(engine.particles_dev), engine.n_local, self.n_feats, self.n_hidden,
            engine.ld, ptr(Xb), ptr(yb), Xb.shape[0], self.n_train, self.alpha, self.beta,
            ptr(engine.scores_dev)))

    def evaluate(self, output, engine, feed_dict):
        import torch
        ctx = engine.ctx
        Xt = ctx.dense(feed_array(feed_dict, self.X, "feature").reshape(-1, self.n_feats))
        out = torch.empty((engine.n_local, Xt.shape[0]), dtype=torch.float32, device=Xt.device)
        ctx.check(ctx.lib.stein_predict_bnn(ctx.handle, ptr(engine.particles_dev), engine.n_local,
                                            self.n_feats, self.n_hidden, engine.ld, ptr(Xt), Xt.shape[0],
                                            ptr(out)))
        return out
