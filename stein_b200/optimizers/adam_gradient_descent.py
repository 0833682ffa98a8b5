from ..runtime import ptr
from ._fused import FusedGradientDescent


class AdamGradientDescent(FusedGradientDescent):
    """Adam step rule with the reference's exact semantics.

    Mirrors stein/optimizers/adam_gradient_descent.py:15-58, quirks included:
    the first call sets `mu = phi`, `nu = phi**2` (not `(1-beta) * phi`, :45-46),
    bias correction uses the already-incremented `n_iters` (:52-54), epsilon 1e-8
    sits outside the square root (:55) and `learning_rate *= decay` happens after
    the step is computed (:56).  Arithmetic: stein_clip_adam_step (CUDA, fp32).
    """
    _kind = "adam"

    def __init__(self, learning_rate=1e-3, decay=1., beta_1=0.9, beta_2=0.999):
        super().__init__(learning_rate, decay)
        self.beta_1 = beta_1
        self.beta_2 = beta_2

    def _hyper(self):
        return dict(optimizer="adam", learning_rate=self.learning_rate, decay=self.decay,
                    p1=self.beta_1, p2=self.beta_2)

    @property
    def mu(self):
        return self._moment(0)

    @property
    def nu(self):
        return self._moment(1)

    def _launch(self, ctx, X, p):
        ctx.check(ctx.lib.stein_clip_adam_step(
            ctx.handle, ptr(X), ptr(p), ptr(self._dev["m1"]), ptr(self._dev["m2"]), X.numel(),
            ptr(self._dev["zero"]), float(self.learning_rate), float(self.beta_1), float(self.beta_2),
            int(self.n_iters)))
