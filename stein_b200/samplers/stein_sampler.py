import numpy as np

from ..kernels import SquaredExponentialKernel
from ..optimizers._fused import FusedGradientDescent
from ..utilities.converters import convert_dictionary_to_array
from .abstract_stein_sampler import AbstractSteinSampler


class SteinSampler(AbstractSteinSampler):
    """SVGD sampler; drop-in for stein/samplers/stein_sampler.py:8-78.

    `train_on_batch(batch_feed)`: scores for all particles in one kernel (instead
    of the reference's n sequential sess.run calls, :59-68), then median ->
    bandwidth -> fused phi -> clip -> optimizer step, all on the device.
    """

    def __init__(self, n_particles, log_p, gd, theta=None):
        super().__init__(n_particles, log_p, theta)
        self.gd = gd
        self.kernel = SquaredExponentialKernel(self.n_particles, self.sess)
        self._make_engine(gd)

    def train_on_batch(self, batch_feed):
        e = self._engine
        e.ctx.sync_stream()
        self.log_p.scores(e, batch_feed)                 # S stays on the device
        if isinstance(self.gd, FusedGradientDescent) and isinstance(self.kernel, SquaredExponentialKernel):
            self._sync_kernel_bandwidth()
            self.gd._push_hyper()          # the reference reads lr / decay / betas at every update()
            e.step()
            self.gd._after_engine_step()
        elif isinstance(self.gd, FusedGradientDescent):
            # another kernel operator assigned to `self.kernel` (abstract_kernel.py:45-62)
            self._plugin_kernel_step()
            self.gd._push_hyper()
            e.apply_phi()
            self.gd._after_engine_step()
        else:
            self._device_phi_only()
            phi = e.get_phi(np.float64)
            phi *= 10. / max(10., np.linalg.norm(phi))
            e.set_particles(e.get_particles(np.float64) + self.gd.update(phi))
        self._theta_cache = None

    @property
    def samples(self):
        """(n x n_params) float64 matrix of the particles (:73-78)."""
        return convert_dictionary_to_array(self.theta)[0]
