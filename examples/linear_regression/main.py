"""Bayesian linear regression with SVGD -- the stein_b200 counterpart of the reference's
examples/linear_regression/main.py (same model, particle count, optimizer and iteration count).

The reference builds `log_p` as a TensorFlow-1 graph (examples/linear_regression/main.py:18-31);
TensorFlow 1.12 cannot run here, so the same graph is provided as
`stein_b200.log_p.LinearRegression`, whose scores for all particles are one batched CUDA kernel.
Data: BASELINE.json config A (synthetic 1000 x 10 from the reference's generator recipe,
examples/linear_regression/data/generator.py:5-9) unless --csv points at a directory with the
reference's data_X.csv / data_y.csv / data_w.csv.  The plot of the reference (:57-66) is omitted.

    python examples/linear_regression/main.py [--particles 100] [--iters 500] [--csv DIR]
"""
import argparse
import os
import sys
from time import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from stein.log_p import LinearRegression  # noqa: E402
from stein.optimizers import AdamGradientDescent  # noqa: E402
from stein.samplers import SteinSampler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=100)
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--csv", default=None)
    args = ap.parse_args()

    if args.csv:
        data_X = np.loadtxt(os.path.join(args.csv, "data_X.csv"), delimiter=",")
        if data_X.ndim == 1:
            data_X = np.atleast_2d(data_X).T
        data_w = np.atleast_2d(np.loadtxt(os.path.join(args.csv, "data_w.csv"), delimiter=",")).T
        data_y = np.atleast_2d(np.loadtxt(os.path.join(args.csv, "data_y.csv"), delimiter=",")).T
    else:
        rng = np.random.default_rng(0)
        n, k = 1000, 10
        data_X = rng.normal(size=(n, k))
        data_w = 5.0 * rng.normal(size=(k, 1))
        data_y = rng.normal(data_X.dot(data_w), 0.3)
    n_samples, n_feats = data_X.shape

    model = LinearRegression(n_feats)
    start_time = time()
    gd = AdamGradientDescent(learning_rate=1e-1)
    sampler = SteinSampler(args.particles, model.log_p, gd)
    for i in range(args.iters):
        sampler.train_on_batch({model.X: data_X, model.y: data_y})

    est = np.array(list(sampler.theta.values()))[0].mean(axis=0).ravel()
    # conjugate model: exact posterior mean (X^T X + I)^-1 X^T y
    exact = np.linalg.solve(data_X.T @ data_X + np.eye(n_feats), data_X.T @ data_y).ravel()
    print("True coefficients: {}".format(data_w.ravel()))
    print("Est. coefficients: {}".format(est))
    print("Exact post. mean : {}".format(exact))
    print("Time elapsed: {:.3f} s for {} iterations".format(time() - start_time, args.iters))


if __name__ == "__main__":
    main()
