"""Error of phi against the float64 evaluation of the reference formula
(stein/kernels/squared_exponential_kernel.py:22-35, stein/samplers/abstract_stein_sampler.py:105)
for every phi implementation on clouds of different conditioning: one Gaussian cloud, an offset
cloud, two clusters at +-3 / +-10, a 60/40 split at +-1, the config-E mixture.  Also the error of
the fp32 oracle itself (the reference's arithmetic).  Run on a B200:  python tools/phi_conditioning_study.py
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import svgd_oracle as orc  # noqa: E402


def clouds(n, d, rng):
    Z = rng.standard_normal((n, d))
    half = (rng.random(n) < 0.5)[:, None]          # random membership: clusters of unequal size
    even = (np.arange(n) % 2 == 0)[:, None]
    yield "gauss", Z
    yield "clusters_pm3_exact5050", np.where(even, 3.0, -3.0) + 0.3 * Z
    yield "offset10_sd0.1", 10.0 + 0.1 * Z
    yield "clusters_pm3_sd0.3", np.where(half, 3.0, -3.0) + 0.3 * Z
    yield "clusters_pm10_sd0.1", np.where(half, 10.0, -10.0) + 0.1 * Z
    sixty = (rng.random(n) < 0.6)[:, None]
    yield "split6040_pm1_sd0.05", np.where(sixty, 1.0, -1.0) + 0.05 * Z
    means = np.zeros((4, d))
    means[0, 0], means[1, 0], means[2, 1], means[3, 1] = 2, -2, 2, -2
    yield "gmm_configE", means[rng.integers(0, 4, n)] + Z
    yield "clusters_1axis_pm20_sd1", np.concatenate([np.where(half, 20.0, -20.0), np.zeros((n, d - 1))], 1) + Z


def phi_float64(X, S, bw):
    X64, S64 = X.astype(np.float64), S.astype(np.float64)
    r = (X64 ** 2).sum(1)
    D = r[:, None] + r[None, :] - 2 * X64 @ X64.T
    h2 = float(bw) ** 2
    K = np.exp(-D / h2 / 2)
    dK = (X64 * K.sum(1)[:, None] - K @ X64) / h2
    return (K @ S64 + dK) / X.shape[0]


def main():
    import torch
    from stein_b200 import _lib
    from stein_b200.runtime import context
    ctx = context()
    n, d = int(os.environ.get("N", 3072)), int(os.environ.get("D", 256))
    rng = np.random.default_rng(5)
    impls = [("dense", _lib.PHI_DENSE_SIMT), ("bf16x3", _lib.PHI_FLASH_TC2), ("fast", _lib.PHI_FLASH_TC4),
             ("precise", _lib.PHI_FLASH_TC5), ("auto", _lib.PHI_AUTO)]
    print("%-26s %9s %9s | %s | %s" % ("cloud", "kappa", "pred", " ".join("%9s" % k for k, _ in impls + [("oracle", 0)]),
                                    "route"))
    vp = ctypes.c_void_p
    for name, X in clouds(n, d, rng):
        X = X.astype(np.float32)
        S = (rng.standard_normal((n, d)) - X).astype(np.float32)
        Xd, Sd = ctx.to_padded(X), ctx.to_padded(S)
        rows, ld = Xd.shape
        r = torch.empty(rows, dtype=torch.float32, device=Xd.device)
        ctx.check(ctx.lib.stein_row_norms(ctx.handle, vp(Xd.data_ptr()), n, d, ld, vp(r.data_ptr())))
        med = ctypes.c_float()
        ctx.check(ctx.lib.stein_median_sqdist(ctx.handle, vp(Xd.data_ptr()), vp(r.data_ptr()), n, d, ld,
                                              ctypes.byref(med), None, None))
        bw = ctx.lib.stein_bandwidth(med.value, n)
        ref = phi_float64(X, S, bw)
        errs, route = [], None
        for _, code in impls:
            ctx.set_phi_impl(code)
            nb = int(ctx.lib.stein_phi_workspace_bytes(ctx.handle, n, n, d))
            ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
            phi = torch.zeros_like(Xd)
            sumsq = torch.zeros(1, dtype=torch.float64, device=Xd.device)
            ctx.check(ctx.lib.stein_phi(ctx.handle, vp(Xd.data_ptr()), vp(Sd.data_ptr()), vp(r.data_ptr()), n, d, ld, 0, n,
                                        bw, vp(ws.data_ptr()), nb, vp(phi.data_ptr()), vp(sumsq.data_ptr())))
            got = phi.cpu().numpy()[:n, :d].astype(np.float64)
            errs.append(np.abs(got - ref).max() / np.abs(ref).max())
            if code == _lib.PHI_AUTO:
                route = ctx.phi_route()
        ctx.set_phi_guard_tol(1.0)        # kappa of this cloud (a guarded call that stays on the fast route)
        ctx.set_phi_guard_tol(5e-5)
        ctx.set_phi_impl(_lib.PHI_AUTO)
        o = orc.compute_phi(X, S.astype(np.float64))
        errs.append(np.abs(o - ref).max() / np.abs(ref).max())
        print("%-26s %9.3g %9.2e | %s | %s" % (name, route["kappa"], route["predicted_fast_error"],
                                               " ".join("%9.2e" % e for e in errs), route["route"]), flush=True)


if __name__ == "__main__":
    main()
