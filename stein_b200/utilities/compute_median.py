"""Median of all entries of an array by the reference's rule.

Mirrors stein/utilities/compute_median.py:4-16: `top_k(V, dim//2 + 1)`; an even
count takes the fp32 mean of the two middle values, an odd count the middle one.
Runs on the GPU (radix select in libstein_b200.so).  The sampler itself never
materialises D: it uses `median_sqdist` below (stein_median_sqdist in the C ABI).
"""
import ctypes

import numpy as np

from ..runtime import context, ptr


def compute_median(D):
    """D: array-like (any shape; NumPy or torch tensor).  Returns np.float32."""
    ctx = context()
    try:
        import torch
        is_tensor = isinstance(D, torch.Tensor)
    except ImportError:                                   # pragma: no cover
        is_tensor = False
    if is_tensor:
        v = D.detach().to("cuda:%d" % ctx.device, dtype=torch.float32).reshape(-1).contiguous()
    else:
        v = ctx.dense(np.asarray(D, dtype=np.float32).reshape(-1))
    med = ctypes.c_float()
    ctx.check(ctx.lib.stein_median_values(ctx.handle, ptr(v), v.numel(), ctypes.byref(med)))
    return np.float32(med.value)


def median_sqdist(theta, return_middle=False):
    """Exact median of the n x n squared-distance matrix of the rows of `theta`
    (stein/kernels/abstract_kernel.py:33-38) without forming it."""
    ctx = context()
    a = np.asarray(theta, dtype=np.float32)
    n, d = a.shape
    # like stein_engine_create: many particles of up to 1 024 coordinates get a leading dimension the
    # tensor-core kernels take (the zero pad columns are extra coordinates: same distances)
    ld = None
    if n >= 2048 and d <= 1024:
        ld = 128 if d <= 128 else -(-d // 256) * 256
    X = ctx.to_padded(a, ld)
    import torch
    r = torch.empty(X.shape[0], dtype=torch.float32, device=X.device)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, ptr(X), n, d, X.shape[1], ptr(r)))
    med, mid, sweeps = ctypes.c_float(), (ctypes.c_float * 2)(), ctypes.c_int32()
    ctx.check(ctx.lib.stein_median_sqdist(ctx.handle, ptr(X), ptr(r), n, d, X.shape[1],
                                          ctypes.byref(med), mid, ctypes.byref(sweeps)))
    if return_middle:
        return np.float32(med.value), (np.float32(mid[0]), np.float32(mid[1])), sweeps.value
    return np.float32(med.value)
