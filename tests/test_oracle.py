"""CPU tests of the oracle itself (test infrastructure must be right first).

Pins: reference-generated golden vectors for the pure-NumPy parts of the
reference (optimizers, converters; tests/golden/make_golden.py) and the analytic
posterior of the reference's shipped linear-regression data.  What lived inside
TensorFlow in the reference is pinned separately, against a run of the reference's
own library on the TF1 stand-in (tests/test_reference_run.py); here the oracle is
additionally cross-checked against an independent autograd restatement and between
its two implementations (NumPy / C).
"""
import os
import types

import numpy as np
import pytest

from oracle import svgd_oracle as orc


class FakeVar:
    def __init__(self, name, shape):
        self.name, self._shape = name, list(shape)

    def get_shape(self):
        return types.SimpleNamespace(as_list=lambda: list(self._shape))


def test_optimizers_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "optimizers.npz"))
    phis = g["phis"]
    for tag, gd in (("adam", orc.AdamGradientDescent(learning_rate=0.1, decay=0.999)),
                    ("adam_default", orc.AdamGradientDescent()),
                    ("adagrad", orc.AdagradGradientDescent(learning_rate=0.05, decay=0.5, alpha=0.9))):
        for t in range(phis.shape[0]):
            np.testing.assert_array_equal(gd.update(phis[t].copy()), g[tag + "_updates"][t])
        assert gd.learning_rate == float(g[tag + "_final_lr"])
        assert gd.n_iters == int(g[tag + "_n_iters"])


def test_adam_first_step_quirk():
    # adam_gradient_descent.py:45-55: mu=phi, nu=phi^2 -> lr * 10 phi / (1e-8 + sqrt(1000) |phi|)
    gd = orc.AdamGradientDescent(learning_rate=0.1)
    step = gd.update(np.array([[2.0, -3.0]]))
    np.testing.assert_allclose(step, 0.1 * np.sqrt(0.001) / 0.1 * np.sign([[2.0, -3.0]]), rtol=1e-6)


def test_converters_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "converters.npz"))
    shapes = [[int(x) for x in s.split(",")] if s else [] for s in g["shapes"]]
    vs = [FakeVar(str(nm), sh) for nm, sh in zip(g["names"], shapes)]
    dictionary = {v: g["value_%d" % i] for i, v in enumerate(vs)}
    array, access = orc.convert_dictionary_to_array(dictionary)
    np.testing.assert_array_equal(array, g["array"])
    assert [access[v][0] for v in vs] == list(g["starts"])
    assert [access[v][1] for v in vs] == list(g["stops"])
    back = orc.convert_array_to_dictionary(array, access)
    for v in vs:
        np.testing.assert_array_equal(back[v], dictionary[v])


@pytest.mark.parametrize("n,d", [(6, 3), (7, 4), (50, 1), (100, 10), (129, 33)])
def test_chain_distance_close_to_blas_and_symmetric(n, d):
    rng = np.random.default_rng(n * 100 + d)
    T = rng.standard_normal((n, d))
    D1, D2 = orc.sqdist(T), orc.sqdist_chain(T)
    assert np.array_equal(D2, D2.T)
    assert np.all(np.diag(D2) == 0.0)
    np.testing.assert_allclose(D1, D2, atol=2e-5 * max(1.0, np.abs(D2).max()))
    # the fma chain against a float64-emulated fma (exact product, one rounding per step)
    T32 = T.astype(np.float32)
    acc = np.zeros((n, n), np.float32)
    for k in range(d):
        acc = (np.outer(T32[:, k].astype(np.float64), T32[:, k].astype(np.float64)) + acc).astype(np.float32)
    r = np.diag(acc)
    Dref = ((r[:, None] + r[None, :]).astype(np.float32) - np.float32(2) * acc).astype(np.float32)
    # float64 emulation can double-round in rare cases; allow a few ulps on a tiny fraction
    mism = np.mean(Dref != D2)
    assert mism < 1e-3


@pytest.mark.parametrize("n,d", [(2, 1), (3, 2), (7, 3), (50, 1), (100, 10), (257, 17), (512, 64)])
def test_median_routes_agree(n, d):
    rng = np.random.default_rng(n + d)
    T = rng.standard_normal((n, d)) * 0.5
    D = orc.sqdist_chain(T)
    m_np = orc.compute_median(D)
    m_c, mid = orc.median_chain(T)
    m_r, mid_r = orc.median_chain(T, radix=True)
    assert m_np == m_c == m_r
    assert mid == mid_r
    v = np.sort(D.reshape(-1))
    dim = n * n
    if dim % 2 == 0:
        assert (mid[0], mid[1]) == (v[dim // 2 - 1], v[dim // 2])
    else:
        assert mid[0] == mid[1] == v[dim // 2]


def test_median_with_duplicates_and_tf_topk_rule():
    # compute_median.py:9-15 on a hand-made vector: top_k(V, dim//2+1)
    V = np.array([[5., 1.], [3., 3.]], np.float32)       # even: (3+3)/2
    assert orc.compute_median(V) == np.float32(3.0)
    V = np.array([4., 1., 9.], np.float32)               # odd: middle
    assert orc.compute_median(V) == np.float32(4.0)
    V = np.array([1., 2., 4., 8.], np.float32)           # even: (2+4)/2
    assert orc.compute_median(V) == np.float32(3.0)
    T = np.zeros((6, 3))
    T[3:] = 1.0                                          # two clusters: D in {0, 3}
    m, mid = orc.median_chain(T)
    assert (mid[0], mid[1]) == (0.0, 3.0) and m == np.float32(1.5)


def test_kernel_and_grad_against_autograd():
    import torch
    rng = np.random.default_rng(3)
    for n, d in [(6, 3), (9, 4), (40, 2)]:
        T = rng.standard_normal((n, d))
        K, dK, bw = orc.kernel_and_grad(T, chain=False)
        t = torch.tensor(T.astype(np.float32).astype(np.float64), requires_grad=True)
        r = (t * t).sum(1, keepdim=True)
        D = r + r.T - 2 * t @ t.T
        Kt = torch.exp(-D / float(bw) ** 2 / 2.0)
        (g,) = torch.autograd.grad(Kt.sum(), t)
        np.testing.assert_allclose(K, Kt.detach().numpy(), atol=2e-6)
        # squared_exponential_kernel.py:32 -- grads = -0.5 * vstack(tf.gradients(K, theta))
        np.testing.assert_allclose(dK, -0.5 * g.numpy(), rtol=2e-5, atol=2e-6)


def test_phi_numpy_vs_c_rows():
    rng = np.random.default_rng(5)
    n, d = 150, 7
    T, S = rng.standard_normal((n, d)), rng.standard_normal((n, d))
    phi = orc.compute_phi(T, S)
    bw = orc.kernel_and_grad(T)[2]
    rows, ks = orc.phi_rows_c(T, S, bw, 20, 90)
    np.testing.assert_allclose(rows, phi[20:90], rtol=1e-5, atol=1e-7)


def test_scores_closed_form_vs_autograd():
    rng = np.random.default_rng(0)
    n, F, N, H = 5, 4, 23, 6
    X = rng.standard_normal((N, F))
    y = rng.standard_normal(N)
    th = rng.standard_normal((n, F))
    np.testing.assert_allclose(orc.score_linear(th, X, y),
                               orc.score_autograd(orc.log_p_linear_torch, th, X, y), rtol=1e-10)
    th = rng.standard_normal((n, F + 1)) * 0.5
    yb = (rng.random(N) > 0.5).astype(float)
    np.testing.assert_allclose(orc.score_logistic(th, X, yb, 1000.0),
                               orc.score_autograd(orc.log_p_logistic_torch, th, X, yb, 1000.0), rtol=1e-9)
    d = 2 + F * H + 2 * H + 1
    th = rng.standard_normal((n, d)) * 0.3
    np.testing.assert_allclose(orc.score_bnn(th, X, y, 500.0, F, H),
                               orc.score_autograd(orc.log_p_bnn_torch, th, X, y, 500.0, F, H),
                               rtol=1e-8, atol=1e-12)


def test_linear_regression_known_answer(golden_dir):
    """The reference's only shipped fixture: examples/linear_regression/data/*.csv.
    50 particles, Adam lr 0.1, 500 full-batch iterations (main.py:36-44) must land
    on the analytic posterior N(0.38394912, 0.0319173^2)."""
    g = np.load(os.path.join(golden_dir, "linear_regression.npz"))
    X, y = g["X"], g["y"]
    np.random.seed(0)
    theta = np.random.normal(size=(50, 1)) * 0.01         # abstract_stein_sampler.py:69-74
    gd = orc.AdamGradientDescent(learning_rate=1e-1)
    for _ in range(500):
        S = orc.score_linear(theta, X.astype(np.float32), y.astype(np.float32))
        theta, _ = orc.update_particles(theta, S, gd)
    assert abs(theta.mean() - g["post_mean"][0]) < 5e-3
    sd = np.sqrt(g["post_cov"][0, 0])
    assert 0.5 * sd < theta.std() < 1.5 * sd
    assert abs(g["post_mean"][0] - 0.38394912) < 1e-7 and abs(g["post_cov"][0, 0] - 0.00101872) < 1e-7


def test_blocked_iteration_matches_dense():
    rng = np.random.default_rng(9)
    n, d = 300, 12
    T, S = rng.standard_normal((n, d)), rng.standard_normal((n, d))
    phi_b, bw_b = orc.iteration_blocked_numpy(T, S, row_block=128)
    phi = orc.compute_phi(T, S, chain=False)
    bw = orc.kernel_and_grad(T, chain=False)[2]
    assert bw == bw_b
    np.testing.assert_allclose(phi_b, phi, rtol=1e-5, atol=1e-8)


def test_imq_kernel_and_grad_against_autograd():
    """The inverse multiquadric operator of the oracle (plugin point abstract_kernel.py:45-62): its
    closed-form dK equals the reference's recipe -0.5 * d(sum K)/d theta_i
    (squared_exponential_kernel.py:23,32) evaluated by autograd, bandwidth held fixed (stop_gradient,
    abstract_kernel.py:40)."""
    import torch
    rng = np.random.default_rng(5)
    theta = rng.standard_normal((23, 4)).astype(np.float32)
    for beta in (-0.5, -1.0):
        K, dK, h = orc.imq_kernel_and_grad(theta, beta=beta)
        T = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
        r = (T * T).sum(1, keepdim=True)
        D = r + r.T - 2 * T @ T.T
        Kt = (1.0 + D / float(h) ** 2) ** beta
        (g,) = torch.autograd.grad(Kt.sum(), T)
        np.testing.assert_allclose(K, Kt.detach().numpy(), rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(dK, -0.5 * g.numpy(), rtol=1e-4, atol=2e-5 * np.abs(g.numpy()).max())
