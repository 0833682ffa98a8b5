"""Posterior-level checks of the logistic-regression and BNN examples at the BASELINE.json shapes
(configs B and C): a real run of a few hundred SVGD iterations on the GPU path against the CPU oracle
run on the same data, the same minibatch stream and the same initial particles.

Individual particles of two fp32 implementations drift apart over hundreds of Adam steps (Adam
normalises phi, so rounding noise in near-zero entries is amplified); what must agree is what the
reference's examples report: the predictive accuracy of the logistic model
(examples/logistic_regression/main.py:52-61) and the test error of the BNN
(examples/regression_neural_network/main.py:95-102), plus the posterior mean of the weights.
"""
import numpy as np
import pytest

from oracle import svgd_oracle as orc

pytestmark = pytest.mark.gpu


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def test_logistic_posterior_matches_the_oracle_run():
    """Config B: 1 024 particles, synthetic covertype-shape data 581 012 x 54 (80/20 split as
    examples/logistic_regression/main.py:14-16), minibatch 50, Adam lr 0.1, 300 iterations."""
    from stein_b200.log_p import LogisticRegression
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.samplers import SteinSampler
    rng = np.random.default_rng(0)
    N_all, F, n, iters, B = 581012, 54, 1024, 300, 50
    Xd = rng.standard_normal((N_all, F)).astype(np.float32)
    w_true = rng.standard_normal(F)
    y = (rng.random(N_all) < _sigmoid(Xd @ w_true)).astype(np.float32)[:, None]
    n_train = int(0.8 * N_all)
    Xtr, ytr, Xte, yte = Xd[:n_train], y[:n_train], Xd[n_train:n_train + 20000], y[n_train:n_train + 20000]
    batches = np.random.default_rng(2).integers(0, n_train, size=(iters, B))

    model = LogisticRegression(F, n_train)
    np.random.seed(3)
    sampler = SteinSampler(n, model.log_p, AdamGradientDescent(learning_rate=1e-1))
    theta = sampler.samples.copy()
    gd = orc.AdamGradientDescent(learning_rate=1e-1)
    for it in range(iters):
        idx = batches[it]
        Xb, yb = Xtr[idx], ytr[idx]
        sampler.train_on_batch({model.X: Xb, model.y: yb})
        theta, _ = orc.update_particles(theta, orc.score_logistic(theta, Xb, yb, n_train), gd)
    got = sampler.samples

    def accuracy(th):          # main.py:57-61: mean over particles of the predictive probability, thresholded
        p = _sigmoid(th[:, :F] @ Xte.T.astype(np.float64)).mean(axis=0)
        return float(((p > 0.5) == (yte[:, 0] > 0.5)).mean())

    logits = sampler.function_posterior(model.logits, {model.X: Xte})          # GPU: all particles at once
    acc_gpu_api = float(((_sigmoid(logits).mean(axis=0) > 0.5) == (yte[:, 0] > 0.5)).mean())
    acc_gpu, acc_ref = accuracy(got), accuracy(theta)
    print("logistic: accuracy gpu %.4f (function_posterior %.4f) oracle %.4f" % (acc_gpu, acc_gpu_api, acc_ref))
    assert abs(acc_gpu_api - acc_gpu) <= 1e-3
    assert abs(acc_gpu - acc_ref) <= 5e-3
    assert acc_gpu > 0.8                      # the run learned something (data are linearly generated)
    # posterior mean of the weights: within a fraction of the posterior spread of the oracle's particles
    sd = theta[:, :F].std(axis=0).mean()
    assert np.abs(got[:, :F].mean(0) - theta[:, :F].mean(0)).max() <= 0.5 * sd + 1e-3
    assert abs(got[:, F].mean() - theta[:, F].mean()) <= 0.1 * max(1.0, abs(theta[:, F].mean()))


@pytest.mark.parametrize("shape", ["boston", "yearmsd_reduced"])
def test_bnn_posterior_matches_the_oracle_run(shape):
    """Config C: 512 particles, one hidden layer of 50 units; Boston-housing shape 506 x 13 (d = 753) and a
    reduced YearMSD shape 20 000 x 90 (d = 4 603; the full 515 345 rows only change the minibatch
    pool); minibatch 100, Adam lr 0.1 decay 0.999 (examples/regression_neural_network/main.py:88)."""
    from stein_b200.log_p import RegressionNeuralNetwork
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.samplers import SteinSampler
    rng = np.random.default_rng(1)
    N, F = (506, 13) if shape == "boston" else (20000, 90)
    H, n, B = 50, 512, 100
    iters = 200 if shape == "boston" else 60
    Xd = rng.standard_normal((N, F)).astype(np.float32)
    w1 = rng.standard_normal((F, 4)) / np.sqrt(F)
    yd = (np.tanh(Xd @ w1).sum(1) + 0.1 * rng.standard_normal(N)).astype(np.float32)
    yd = ((yd - yd.mean()) / yd.std()).astype(np.float32)[:, None]
    n_train = int(0.9 * N)
    Xtr, ytr, Xte, yte = Xd[:n_train], yd[:n_train], Xd[n_train:], yd[n_train:]
    batches = np.random.default_rng(2).integers(0, n_train, size=(iters, B))

    model = RegressionNeuralNetwork(F, H, n_train)
    np.random.seed(4)
    sampler = SteinSampler(n, model.log_p, AdamGradientDescent(learning_rate=1e-1, decay=0.999))
    theta = sampler.samples.copy()
    gd = orc.AdamGradientDescent(learning_rate=1e-1, decay=0.999)
    for it in range(iters):
        idx = batches[it]
        Xb, yb = Xtr[idx], ytr[idx]
        sampler.train_on_batch({model.X: Xb, model.y: yb})
        theta, _ = orc.update_particles(theta, orc.score_bnn(theta, Xb, yb, n_train, F, H), gd)
    got = sampler.samples

    def rmse(th):              # main.py:100-102: posterior-mean prediction against the held-out targets
        pred = orc.bnn_predict(th, Xte, F, H).mean(axis=0)
        return float(np.sqrt(((pred - yte[:, 0]) ** 2).mean()))

    pred_api = sampler.function_posterior(model.pred, {model.X: Xte}).mean(axis=0)
    rmse_api = float(np.sqrt(((pred_api - yte[:, 0]) ** 2).mean()))
    r_gpu, r_ref = rmse(got), rmse(theta)
    print("bnn %s: test rmse gpu %.4f (function_posterior %.4f) oracle %.4f" % (shape, r_gpu, rmse_api, r_ref))
    assert abs(rmse_api - r_gpu) <= 1e-3
    assert abs(r_gpu - r_ref) <= 0.03 * max(r_ref, 0.1)
    assert r_gpu < 1.0                        # better than predicting the mean of the standardised targets
