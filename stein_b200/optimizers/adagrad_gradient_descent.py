from ..runtime import ptr
from ._fused import FusedGradientDescent


class AdagradGradientDescent(FusedGradientDescent):
    """"Adagrad" step rule of the reference (an RMSProp-style running average).

    Mirrors stein/optimizers/adagrad_gradient_descent.py:13-44: first call
    `hist = phi**2` (:37-38), then `hist = alpha*hist + (1-alpha)*phi**2` (:40);
    step `phi / (1e-6 + sqrt(hist)) * learning_rate` (:44).  `decay` is accepted
    and stored but never applied -- the reference class has no
    `learning_rate *= decay` line.  Arithmetic: stein_clip_adagrad_step (CUDA).
    """
    _kind = "adagrad"

    def __init__(self, learning_rate=1e-3, decay=1., alpha=0.9):
        super().__init__(learning_rate, decay)
        self.alpha = alpha

    def _hyper(self):
        return dict(optimizer="adagrad", learning_rate=self.learning_rate, decay=self.decay,
                    p1=self.alpha, p2=0.0)

    @property
    def hist(self):
        return self._moment(0)

    def _launch(self, ctx, X, p):
        ctx.check(ctx.lib.stein_clip_adagrad_step(
            ctx.handle, ptr(X), ptr(p), ptr(self._dev["m1"]), X.numel(), ptr(self._dev["zero"]),
            float(self.learning_rate), float(self.alpha), int(self.n_iters)))
