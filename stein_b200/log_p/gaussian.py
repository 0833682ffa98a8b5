import numpy as np

from ..runtime import ptr
from .base import LogPosterior


class GaussianMixtureTarget(LogPosterior):
    """Synthetic data-free targets of BASELINE.json configs D and E: an
    equal-weight mixture of isotropic Gaussians N(mu_k, sigma2 I) over one flat
    variable of dimension d.  `means=None` is the standard normal target
    (score S = -X / sigma2).  Scores by stein_score_gaussian_mixture (CUDA).

    Not in the reference (its examples are the three regression models); it is
    the score function the headline benchmark shapes are defined on
    (SURVEY.md section 8d).
    """

    def __init__(self, dim, means=None, sigma2=1.0):
        super().__init__()
        self.dim, self.sigma2 = int(dim), float(sigma2)
        self.x = self._variable([self.dim])
        self.means = None if means is None else np.ascontiguousarray(means, dtype=np.float32)
        if self.means is not None and self.means.shape[1] != self.dim:
            raise ValueError("means must be (n_components x dim)")
        self._means_dev = None

    def scores(self, engine, batch_feed=None):
        ctx = engine.ctx
        if self.means is not None and self._means_dev is None:
            self._means_dev = ctx.dense(self.means)
        ncomp = 1 if self.means is None else self.means.shape[0]
        ctx.check(ctx.lib.stein_score_gaussian_mixture(
            ctx.handle, ptr(engine.particles_dev), engine.n_local, self.dim, engine.ld,
            None if self.means is None else ptr(self._means_dev), ncomp, self.sigma2,
            ptr(engine.scores_dev)))
