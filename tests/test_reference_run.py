"""The oracle against tests/golden/reference_run.npz: outputs of the REFERENCE'S OWN library code
(kernels, compute_median, samplers, optimizers, imported unmodified from the reference tree by
tests/golden/make_golden_reference_run.py) executed on the TF1 stand-in of compat/.  This pins the
parts of the path that lived behind TensorFlow in the reference -- D, the top_k median rule, the
bandwidth, K, the tf.gradients-based dK with its -0.5 post-scale, phi, the clip, the per-particle
score loop and whole trajectories -- up to float32 rounding of the individual ops (the stand-in
evaluates them with PyTorch, not with TensorFlow's kernels), hence tolerances instead of bits."""
import os

import numpy as np
import pytest

from oracle import svgd_oracle as orc


@pytest.fixture(scope="module")
def ref(golden_dir):
    return np.load(os.path.join(golden_dir, "reference_run.npz"))


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_kernel_operator_matches_the_reference_run(ref):
    """stein/kernels/squared_exponential_kernel.py:25-35 through the reference's own class."""
    for i, (n, d) in enumerate(ref["kernel_shapes"]):
        theta = ref["kernel%d_theta" % i]
        K, dK, bw = orc.kernel_and_grad(theta)
        assert K.shape == (n, n) and dK.shape == (n, d)
        assert _rel(orc.sqdist_chain(theta.astype(np.float32)), ref["kernel%d_D" % i]) < 2e-6
        assert abs(float(bw) - float(ref["kernel%d_bandwidth" % i])) <= 2e-6 * float(bw)
        assert np.abs(K - ref["kernel%d_K" % i]).max() < 2e-6
        assert _rel(dK, ref["kernel%d_dK" % i]) < 1e-5


def test_kernel_operator_with_ties_and_wide_norm_range(ref):
    """Duplicated particles (D = 0 off the diagonal), two tight clusters (the two middle values
    differ), norms over four decades with an odd n*n: same bandwidth, K and dK as the reference."""
    for i in range(int(ref["n_ties"])):
        theta = ref["tie%d_theta" % i]
        with np.errstate(divide="ignore", invalid="ignore"):
            K, dK, bw = orc.kernel_and_grad(theta)
        if float(ref["tie%d_bandwidth" % i]) == 0.0:
            # more than half of the distances are zero: the reference divides by a zero bandwidth
            # and returns NaN wherever D = 0 (the product refuses this input instead, engine.cu)
            assert float(bw) == 0.0
            assert (np.isnan(K) == np.isnan(ref["tie%d_K" % i])).all() and np.isnan(K).any()
            assert np.array_equal(K[~np.isnan(K)], ref["tie%d_K" % i][~np.isnan(K)])
            continue
        assert abs(float(bw) - float(ref["tie%d_bandwidth" % i])) <= 4e-6 * float(bw), i
        # r_i + r_j - 2 x_i.x_j in float32 carries an absolute error of a few ulp(max r) whatever
        # the summation order; it reaches K through exp(-D / 2h^2)
        rmax = float((theta.astype(np.float32) ** 2).sum(axis=1).max())
        atol = 5e-6 + 8.0 * 2.0 ** -24 * rmax / float(bw) ** 2
        assert np.abs(K - ref["tie%d_K" % i]).max() < atol, i
        assert _rel(dK, ref["tie%d_dK" % i]) < 2e-5, i


def test_median_rule_matches_the_reference_run(ref):
    """stein/utilities/compute_median.py:4-16 (top_k with k = n*n/2 + 1; mean of the two middle
    values for an even count): identical values in, identical bits out."""
    for i in range(int(ref["n_median"])):
        got = orc.compute_median(ref["median%d_in" % i])
        assert np.float32(got).tobytes() == np.float32(ref["median%d_out" % i]).tobytes()


CASES = {
    "linear": dict(gd=lambda: orc.AdamGradientDescent(1e-1)),
    "logistic": dict(gd=lambda: orc.AdamGradientDescent(1e-1)),
    "bnn_adam": dict(gd=lambda: orc.AdamGradientDescent(1e-1, 0.999)),
    "bnn_adagrad": dict(gd=lambda: orc.AdagradGradientDescent(5e-2, 0.5, 0.9)),
}


def reference_run_feeds(ref, golden_dir, tag):
    """(score function of the oracle, list of minibatches, prediction function) of a case."""
    if tag == "linear":
        g = np.load(os.path.join(golden_dir, "linear_regression.npz"))
        X, y = g["X"].astype(np.float32), g["y"].reshape(-1, 1).astype(np.float32)
        return (lambda th, b: orc.score_linear(th, *b)), [(X, y)] * 6, (lambda th: th @ X[:7].T.astype(np.float64))
    if tag == "logistic":
        X, y, idx = ref["logistic_X"].astype(np.float32), ref["logistic_y"].astype(np.float32), ref["logistic_batches"]
        F = X.shape[1]
        return ((lambda th, b: orc.score_logistic(th, b[0], b[1], X.shape[0])), [(X[i], y[i]) for i in idx],
                (lambda th: th[:, :F] @ X[:9].T.astype(np.float64)))
    X, y, idx = ref["bnn_X"].astype(np.float32), ref["bnn_y"].astype(np.float32), ref["bnn_batches"]
    F, H = 2, 5
    return ((lambda th, b: orc.score_bnn(th, b[0], b[1], X.shape[0], F, H)), [(X[i], y[i]) for i in idx],
            (lambda th: orc.bnn_predict(th, X[:6], F, H)))


@pytest.mark.parametrize("tag", sorted(CASES))
def test_trajectories_match_the_reference_run(ref, golden_dir, tag):
    """The reference's SteinSampler.train_on_batch (per-particle sess.run of tf.gradients, then
    update_particles) for 5-6 iterations from recorded particles, against the oracle's closed-form
    scores + kernel + clip + optimizer restatement."""
    score, batches, predict = reference_run_feeds(ref, golden_dir, tag)
    traj = ref[tag + "_traj"]
    assert traj.shape[0] == len(batches) + 1
    theta = traj[0].copy()
    assert _rel(score(theta, batches[0]), ref[tag + "_scores0"]) < 2e-5         # stein_sampler.py:59-68
    assert _rel(orc.compute_phi(theta, ref[tag + "_scores0"]), ref[tag + "_phi0"]) < 1e-5   # :76-105
    gd = CASES[tag]["gd"]()
    for it, b in enumerate(batches):
        theta, _ = orc.update_particles(theta, score(theta, b), gd)
        assert _rel(theta, traj[it + 1]) < 5e-5, (tag, it)
    if tag + "_posterior" in ref.files:                                           # :129-168
        assert _rel(predict(traj[-1]), ref[tag + "_posterior"]) < 1e-5
        assert _rel(predict(traj[-1]).mean(axis=0), ref[tag + "_posterior_mean"]) < 1e-5
