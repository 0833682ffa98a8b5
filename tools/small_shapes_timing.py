"""Per-iteration wall time of train_on_batch at the example shapes of BASELINE.json (configs A-C)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from stein.log_p import LinearRegression, LogisticRegression, RegressionNeuralNetwork  # noqa: E402
from stein.optimizers import AdamGradientDescent  # noqa: E402
from stein.samplers import SteinSampler  # noqa: E402


def run(name, model, n_particles, feed_fn, iters=int(os.environ.get("SMALL_ITERS", "200"))):
    sampler = SteinSampler(n_particles, model.log_p, AdamGradientDescent(learning_rate=1e-1))
    for _ in range(int(os.environ.get("SMALL_WARM", "20"))):
        sampler.train_on_batch(feed_fn())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        sampler.train_on_batch(feed_fn())
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    info = sampler.engine.last()
    print("%-40s n=%5d d=%5d  %.3f ms/iteration  (median sweeps %d)" %
          (name, n_particles, model.n_params, dt * 1e3, info["sweeps"]), flush=True)


def main():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((1000, 10)).astype(np.float32)
    y = (X @ rng.standard_normal((10, 1)) * 5 + 0.3 * rng.standard_normal((1000, 1))).astype(np.float32)
    m = LinearRegression(10)
    run("A linear 1000x10, full batch", m, 100, lambda: {m.X: X, m.y: y})
    Xl = rng.standard_normal((100000, 54)).astype(np.float32)
    yl = (rng.random((100000, 1)) < 0.5).astype(np.float32)
    ml = LogisticRegression(54, 464809)
    def feed_l():
        b = rng.choice(100000, 50, replace=False)
        return {ml.X: Xl[b], ml.y: yl[b]}
    run("B logistic 54 feats, batch 50", ml, 1024, feed_l)
    for F, nm in ((13, "C bnn boston 13 feats H=50, batch 100"), (90, "C bnn yearmsd 90 feats H=50, batch 100")):
        Xb = rng.standard_normal((5000, F)).astype(np.float32)
        yb = rng.standard_normal((5000, 1)).astype(np.float32)
        mb = RegressionNeuralNetwork(F, 50, 5000)
        def feed_b(Xb=Xb, yb=yb, mb=mb):
            b = rng.choice(5000, 100, replace=False)
            return {mb.X: Xb[b], mb.y: yb[b]}
        run(nm, mb, 512, feed_b)


if __name__ == "__main__":
    main()
