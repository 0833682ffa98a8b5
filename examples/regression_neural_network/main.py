"""Bayesian neural-network regression with SVGD -- the stein_b200 counterpart of the reference's
examples/regression_neural_network/main.py (one ReLU hidden layer, Gamma(1, 0.01) priors on the
noise and weight precisions, Adam lr 0.1 with decay 0.999, train MSE printed every 1000
iterations; the plot of the reference is omitted).

Shapes: --shape toy is the reference's own 20-point 1-D problem (H = 100, 20 particles);
boston / yearmsd are BASELINE.json config C (H = 50, 512 particles, synthetic 506 x 13 and
515 345 x 90 data, minibatch 100).

    python examples/regression_neural_network/main.py [--shape toy|boston|yearmsd] [--iters N]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from stein.log_p import RegressionNeuralNetwork  # noqa: E402
from stein.optimizers import AdamGradientDescent  # noqa: E402
from stein.samplers import SteinSampler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="toy", choices=["toy", "boston", "yearmsd"])
    ap.add_argument("--iters", type=int, default=10000)
    args = ap.parse_args()
    rng = np.random.default_rng(0)

    if args.shape == "toy":          # examples/regression_neural_network/main.py:12-20
        n_train, n_feats, n_hidden, n_particles, n_batch = 20, 1, 100, 20, 20
        X = rng.uniform(-4.0, 4.0, size=(n_train, n_feats)).astype(np.float32)
        y = (np.sin(X) + 0.1 * rng.standard_normal(X.shape)).astype(np.float32)
    else:
        rows, n_feats = (506, 13) if args.shape == "boston" else (515345, 90)
        n_hidden, n_particles, n_batch = 50, 512, 100
        X = rng.standard_normal((rows, n_feats)).astype(np.float32)
        w = rng.standard_normal((n_feats, 1)).astype(np.float32) / np.sqrt(n_feats)
        y = np.tanh(X @ w) + 0.1 * rng.standard_normal((rows, 1)).astype(np.float32)
        y = ((y - y.mean()) / y.std()).astype(np.float32)
        n_train = rows

    model = RegressionNeuralNetwork(n_feats, n_hidden, n_train)
    gd = AdamGradientDescent(learning_rate=1e-1, decay=0.999)
    sampler = SteinSampler(n_particles, model.log_p, gd)
    X_eval, y_eval = X[:5000], y[:5000]
    for i in range(args.iters):
        batch = rng.choice(n_train, n_batch, replace=False) if n_batch < n_train else np.arange(n_train)
        sampler.train_on_batch({model.X: X[batch], model.y: y[batch]})
        if i % 1000 == 0:
            pred = sampler.function_posterior(model.pred, {model.X: X_eval, model.y: y_eval}).mean(axis=0)
            print("Iteration {} / {}: MSE {:.4f}".format(i, args.iters, np.mean((pred - y_eval.ravel()) ** 2)))
    pred = sampler.function_posterior(model.pred, {model.X: X_eval, model.y: y_eval}).mean(axis=0)
    print("Final MSE: {:.4f}".format(np.mean((pred - y_eval.ravel()) ** 2)))


if __name__ == "__main__":
    main()
