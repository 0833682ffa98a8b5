import ctypes
from abc import abstractmethod

import numpy as np

from ..runtime import context, ptr


class AbstractKernel:
    """Kernel operator of the sampler (plugin point of the reference).

    Mirrors stein/kernels/abstract_kernel.py:7-62.  The reference builds, once,
    a TensorFlow graph with n fp32 placeholders, the distance matrix
    D = r + r^T - 2 T T^T (:33-35), its median (:38) and the bandwidth
    sqrt(median / ln n) (:40).  Here the same quantities are computed per call by
    libstein_b200.so: `sqdist_median(theta)` never forms D, `bandwidth` holds the
    value from the most recent call.

    `sess` is kept for signature compatibility (the reference passes its
    tf.Session); it is unused -- the CUDA context plays that role.
    """

    def __init__(self, n_particles, sess=None, bandwidth=None):
        """`bandwidth` (not in the reference): a fixed h > 0 instead of the median
        heuristic; the samplers then skip the median (stein_engine_set_bandwidth)."""
        self.n_particles = n_particles
        self.sess = sess
        if bandwidth is not None and not (float(bandwidth) > 0.0 and np.isfinite(bandwidth)):
            raise ValueError("bandwidth must be positive and finite")
        self.fixed_bandwidth = None if bandwidth is None else np.float32(bandwidth)
        self.bandwidth = self.fixed_bandwidth   # fp32, set by kernel_and_grad / compute_bandwidth
        self.median = None

    def _device_particles(self, theta):
        import torch
        ctx = context()
        a = np.asarray(theta, dtype=np.float32)
        if a.shape[0] != self.n_particles:
            raise ValueError("theta has %d rows, kernel was built for %d particles"
                             % (a.shape[0], self.n_particles))
        X = ctx.to_padded(a)
        r = torch.empty(X.shape[0], dtype=torch.float32, device=X.device)
        ctx.check(ctx.lib.stein_row_norms(ctx.handle, ptr(X), a.shape[0], a.shape[1], X.shape[1], ptr(r)))
        return ctx, a.shape, X, r

    def compute_bandwidth(self, theta):
        """abstract_kernel.py:38-40 for the rows of `theta` (float32 like the
        reference's placeholders).  Returns the fp32 bandwidth h."""
        ctx, (n, d), X, r = self._device_particles(theta)
        return self._bandwidth_dev(ctx, X, r, n, d)

    def _bandwidth_dev(self, ctx, X, r, n, d):
        if self.fixed_bandwidth is not None:
            self.bandwidth = self.fixed_bandwidth
            return self.bandwidth
        med = ctypes.c_float()
        ctx.check(ctx.lib.stein_median_sqdist(ctx.handle, ptr(X), ptr(r), n, d, X.shape[1],
                                              ctypes.byref(med), None, None))
        self.median = np.float32(med.value)
        self.bandwidth = np.float32(ctx.lib.stein_bandwidth(med.value, n))
        return self.bandwidth

    @abstractmethod
    def kernel_and_grad(self, theta):
        """Returns (K, dK): the n x n kernel matrix of the rows of `theta` and
        the n x d gradient term (abstract_kernel.py:45-62)."""
        raise NotImplementedError()
