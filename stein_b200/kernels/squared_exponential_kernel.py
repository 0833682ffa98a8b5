import numpy as np

from ..runtime import ptr
from .abstract_kernel import AbstractKernel


class SquaredExponentialKernel(AbstractKernel):
    """Isotropic squared-exponential kernel with the median-heuristic bandwidth.

    Mirrors stein/kernels/squared_exponential_kernel.py:6-35:
        K  = exp(-D / bandwidth**2 / 2)                                  (:22)
        dK = -0.5 * vstack(tf.gradients(K, theta_i))                     (:23, :32)
           = (x_i sum_j K_ij - sum_j K_ij x_j) / bandwidth**2
    `kernel_and_grad` returns dense fp32 arrays like the reference; it exists for
    API compatibility and small n (it allocates n x n).  The sampler's hot path
    goes through the fused phi kernel instead and never materialises K.
    """

    def kernel_and_grad(self, theta):
        import torch
        ctx, (n, d), X, r = self._device_particles(theta)
        bw = self._bandwidth_dev(ctx, X, r, n, d)
        rows, ld = X.shape
        K = torch.empty((rows, rows), dtype=torch.float32, device=X.device)
        dK = torch.empty((rows, ld), dtype=torch.float32, device=X.device)
        ws = torch.empty(rows * ld + rows, dtype=torch.float32, device=X.device)
        ctx.check(ctx.lib.stein_kernel_and_grad(ctx.handle, ptr(X), ptr(r), n, d, ld, float(bw), ptr(K),
                                                rows, ptr(dK), ptr(ws), ws.numel() * 4))
        return K[:n, :n].cpu().numpy(), dK[:n, :d].cpu().numpy()
