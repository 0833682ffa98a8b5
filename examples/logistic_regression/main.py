"""Bayesian logistic regression with SVGD -- the stein_b200 counterpart of the reference's
examples/logistic_regression/main.py (same model: Gamma(1, 0.01) hyper-prior on the weight
precision, minibatches of 50, Adam lr 0.1, accuracy on a 20 % test split every 100 iterations).

The covertype data of the reference is not shipped (.MISSING_LARGE_BLOBS); BASELINE.json config B
asks for synthetic data of its shape, 581 012 x 54, and 1 024 particles.

    python examples/logistic_regression/main.py [--particles 1024] [--iters 6000] [--rows 581012]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from stein.log_p import LogisticRegression  # noqa: E402
from stein.optimizers import AdamGradientDescent  # noqa: E402
from stein.samplers import SteinSampler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=6000)
    ap.add_argument("--rows", type=int, default=581012)
    ap.add_argument("--eval-rows", type=int, default=20000, help="test rows used for the accuracy print")
    args = ap.parse_args()

    rng = np.random.default_rng(0)
    n_feats, n_batch = 54, 50
    data_X = rng.standard_normal((args.rows, n_feats)).astype(np.float32)
    w_true = rng.standard_normal((n_feats, 1)).astype(np.float32)
    data_y = (rng.random((args.rows, 1)) < 1.0 / (1.0 + np.exp(-data_X @ w_true))).astype(np.float32)
    n_test = args.rows // 5                                   # train_test_split(test_size=0.2)
    perm = rng.permutation(args.rows)
    X_test, y_test = data_X[perm[:n_test]], data_y[perm[:n_test]]
    X_train, y_train = data_X[perm[n_test:]], data_y[perm[n_test:]]
    n_train = X_train.shape[0]
    X_eval, y_eval = X_test[:args.eval_rows], y_test[:args.eval_rows]

    model = LogisticRegression(n_feats, n_train)

    def evaluate(sampler):
        logits_pred = sampler.function_posterior(model.logits, {model.X: X_eval, model.y: y_eval})
        avg_pred = logits_pred.mean(axis=0) > 0.0
        return np.mean(avg_pred == y_eval.ravel())

    n_prog = 100
    gd = AdamGradientDescent(learning_rate=1e-1)
    sampler = SteinSampler(args.particles, model.log_p, gd)
    for i in range(args.iters):
        if i % n_prog == 0:
            print("Iteration {} / {}: {:4f}".format(i, args.iters, evaluate(sampler)))
        batch = rng.choice(n_train, n_batch, replace=False)
        sampler.train_on_batch({model.X: X_train[batch], model.y: y_train[batch]})
    print("Final accuracy: {:4f}".format(evaluate(sampler)))


if __name__ == "__main__":
    main()
