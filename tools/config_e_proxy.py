"""Config-E-shaped run on one GPU at reduced n (d = 1024 > 256 takes the FFMA paths today):
timing of one iteration and a spot check of phi rows against the C oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from oracle import svgd_oracle as orc  # noqa: E402
from stein_b200.engine import SvgdEngine  # noqa: E402


def main():
    n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 32768, 1024
    rng = np.random.default_rng(3)
    X = rng.standard_normal((n, d)).astype(np.float32)
    means = np.zeros((4, d), np.float32)
    means[0, 0], means[1, 0], means[2, 1], means[3, 1] = 2, -2, 2, -2
    # equal-weight mixture of 4 unit Gaussians: S_i = sum_k r_ik (mu_k - x_i)
    logit = -0.5 * ((X[:, None, :2] - means[None, :, :2]) ** 2).sum(-1)
    resp = np.exp(logit - logit.max(1, keepdims=True))
    resp /= resp.sum(1, keepdims=True)
    S = (resp @ means - X).astype(np.float32)
    eng = SvgdEngine(n, d, "adam", learning_rate=0.05)
    eng.set_particles(X)
    eng.set_scores(S)
    eng.step()                                   # warm-up (allocations, module load)
    eng.set_particles(X)
    eng.set_scores(S)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    info = eng.last()
    phi = eng.get_phi(np.float64)
    rows = [0, n // 2, n - 1]
    err = 0.0
    for i in rows:
        ref, _ = orc.phi_rows_c(X, S, np.float32(info["bandwidth"]), i, i + 1)
        err = max(err, np.abs(phi[i] - ref[0]).max() / np.abs(ref[0]).max())
    print("n=%d d=%d: %.3f s per iteration (%d median sweeps), bandwidth %.6f, phi rows vs oracle %.2e"
          % (n, d, dt, info["sweeps"], info["bandwidth"], err))


if __name__ == "__main__":
    main()
