import numpy as np

from ..runtime import ptr
from .abstract_kernel import AbstractKernel


class InverseMultiquadricKernel(AbstractKernel):
    """Inverse multiquadric kernel with the median-heuristic bandwidth of the reference:

        K  = (1 + D / bandwidth**2) ** beta,   beta < 0 (default -1/2)
        dK = -0.5 * vstack(tf.gradients(K, theta_i))        (the recipe of
             stein/kernels/squared_exponential_kernel.py:23, :32 applied to this K)
           = (-2 beta / bandwidth**2) (x_i sum_j G_ij - sum_j G_ij x_j),  G = (1 + D / bandwidth**2) ** (beta - 1)

    Not in the reference, which ships only the squared-exponential kernel: this is a second kernel
    family through its plugin point AbstractKernel.kernel_and_grad (stein/kernels/abstract_kernel.py:45-62;
    SURVEY.md section 8 f4).  The heavy-tailed IMQ kernel is the usual alternative for SVGD in higher
    dimensions.  Assign an instance to `sampler.kernel`; the sampler then evaluates phi = (K S + dK) / n
    from this operator (stein_phi_from_kernel) instead of the fused squared-exponential path.
    Dense n x n like the reference's kernel_and_grad: meant for the example sizes (n up to a few thousand).
    """

    def __init__(self, n_particles, sess=None, bandwidth=None, beta=-0.5):
        super().__init__(n_particles, sess, bandwidth)
        if not beta < 0:
            raise ValueError("beta must be negative")
        self.beta = float(beta)

    def kernel_and_grad_dev(self, ctx, X, r, n, d):
        """Device version: padded X (rows x ld) and its row norms -> (K rows x rows, dK rows x ld) tensors."""
        import torch
        bw = self._bandwidth_dev(ctx, X, r, n, d)
        rows, ld = X.shape
        K = torch.empty((rows, rows), dtype=torch.float32, device=X.device)
        dK = torch.empty((rows, ld), dtype=torch.float32, device=X.device)
        ws = torch.empty(rows * ld + rows, dtype=torch.float32, device=X.device)
        ctx.check(ctx.lib.stein_imq_kernel_and_grad(ctx.handle, ptr(X), ptr(r), n, d, ld, float(bw), self.beta, ptr(K),
                                                    rows, ptr(dK), ptr(ws), ws.numel() * 4))
        return K, dK

    def kernel_and_grad(self, theta):
        ctx, (n, d), X, r = self._device_particles(theta)
        K, dK = self.kernel_and_grad_dev(ctx, X, r, n, d)
        return K[:n, :n].cpu().numpy(), dK[:n, :d].cpu().numpy()
