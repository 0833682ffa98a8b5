"""Accuracy of the phi paths on a particle cloud away from the origin, against a float64
evaluation of the same formula (bandwidth taken from the oracle)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import svgd_oracle as orc  # noqa: E402
from stein_b200 import _lib  # noqa: E402
from stein_b200.runtime import context  # noqa: E402
import test_gpu_parity as T  # noqa: E402


def truth(X, S, bw):
    X64, S64 = X.astype(np.float64), S.astype(np.float64)
    r = (X64 ** 2).sum(1)
    D = r[:, None] + r[None, :] - 2 * X64 @ X64.T
    h2 = float(bw) ** 2
    K = np.exp(-D / h2 / 2)
    dK = (X64 * K.sum(1)[:, None] - K @ X64) / h2
    return (K @ S64 + dK) / X.shape[0]


def main():
    ctx = context(0)
    n, d = 2000, 256
    rng = np.random.default_rng(99)
    Z = rng.standard_normal((n, d))
    for off in (0.0, 1.0, 3.0, 10.0):
        X = (off + 0.1 * Z).astype(np.float32)
        S = (rng.standard_normal((n, d)) - 10.0 * Z).astype(np.float32)
        bw = orc.kernel_and_grad(X)[2]
        ref64 = truth(X, S, bw)
        o = orc.compute_phi(X, S.astype(np.float64))
        line = "offset %5.1f  oracle-vs-f64 %.2e" % (off, T._rel(o, ref64)[1])
        for name, impl in (("dense", _lib.PHI_DENSE_SIMT), ("flash", _lib.PHI_FLASH_TC), ("pair", _lib.PHI_FLASH_TC2)):
            phi, _, _ = T._phi_gpu(ctx, X, S, impl)
            line += "  %s-vs-f64 %.2e vs-oracle %.2e" % (name, T._rel(phi, ref64)[1], T._rel(phi, o)[1])
        print(line, flush=True)


if __name__ == "__main__":
    main()
