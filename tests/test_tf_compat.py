"""CPU tests of the TensorFlow-1 stand-in (compat/tensorflow) and of GraphLogPosterior: the graphs
of the reference's three examples, written against the stand-in the way the reference writes them
against TensorFlow, give the oracle's closed-form scores (oracle/svgd_oracle.py)."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.join(REPO, "compat") not in sys.path:
    sys.path.insert(0, os.path.join(REPO, "compat"))

import tensorflow as tf  # noqa: E402  (the stand-in)
from tensorflow.contrib.distributions import Gamma, Normal  # noqa: E402

from oracle import svgd_oracle as orc  # noqa: E402
from stein_b200.log_p import GraphLogPosterior  # noqa: E402
from stein_b200.utilities.converters import convert_dictionary_to_array  # noqa: E402

torch = pytest.importorskip("torch")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def linear_graph(F):
    """The model of examples/linear_regression/main.py:18-31."""
    tf.reset_default_graph()
    with tf.variable_scope("model"):
        X = tf.placeholder(tf.float32, shape=[None, F])
        y = tf.placeholder(tf.float32, shape=[None, 1])
        w = tf.Variable(tf.zeros([F, 1]))
        with tf.variable_scope("priors"):
            prior = Normal(tf.zeros([F, 1]), 1.)
        with tf.variable_scope("likelihood"):
            y_hat = tf.matmul(X, w)
            log_l = -0.5 * tf.reduce_sum(tf.square(y_hat - y))
        log_p = log_l + tf.reduce_sum(prior.log_prob(w))
    return dict(X=X, y=y, w=w, y_hat=y_hat, log_p=log_p)


def logistic_graph(F, n_train, n_batch):
    """The model of examples/logistic_regression/main.py:23-49."""
    tf.reset_default_graph()
    with tf.variable_scope("model"):
        X = tf.placeholder(tf.float32, shape=[None, F])
        y = tf.placeholder(tf.float32, shape=[None, 1])
        w = tf.Variable(tf.zeros([F, 1]))
        log_alpha = tf.Variable(tf.zeros([]))
        alpha = tf.exp(log_alpha)
        with tf.variable_scope("priors"):
            w_prior = Normal(tf.zeros([F, 1]), tf.reciprocal(tf.sqrt(alpha)))
            alpha_prior = Gamma(1., 0.01)
        with tf.variable_scope("likelihood"):
            logits = tf.matmul(X, w)
            log_l = -tf.reduce_sum(tf.nn.sigmoid_cross_entropy_with_logits(labels=y, logits=logits))
        log_p = (log_l * (n_train / n_batch) + tf.reduce_sum(w_prior.log_prob(w)) + alpha_prior.log_prob(alpha))
    return dict(X=X, y=y, w=w, log_alpha=log_alpha, logits=logits, log_p=log_p)


def bnn_graph(F, H, n_train, n_batch, a=1., b=0.01):
    """The model of examples/regression_neural_network/main.py:29-85."""
    tf.reset_default_graph()
    with tf.variable_scope("model"):
        X = tf.placeholder(tf.float32, shape=[None, F])
        y = tf.placeholder(tf.float32, shape=[None, 1])
        log_lambda = tf.Variable(tf.zeros([]))
        log_gamma = tf.Variable(tf.zeros([]))
        lam, gam = tf.exp(log_lambda), tf.exp(log_gamma)
        w1 = tf.Variable(tf.zeros([F, H]))
        b1 = tf.Variable(tf.zeros([H]))
        w2 = tf.Variable(tf.zeros([H, 1]))
        b2 = tf.Variable(tf.zeros([]))
        with tf.variable_scope("prediction"):
            pred = tf.matmul(tf.nn.relu(tf.matmul(X, w1) + b1), w2) + b2
        with tf.variable_scope("likelihood"):
            log_l = tf.reduce_sum(Normal(pred, tf.reciprocal(tf.sqrt(gam))).log_prob(y))
        with tf.variable_scope("priors"):
            sd = tf.reciprocal(tf.sqrt(lam))
            priors = [Normal(tf.zeros([F, H]), sd).log_prob(w1), Normal(tf.zeros([H, 1]), sd).log_prob(w2),
                      Normal(tf.zeros([H]), sd).log_prob(b1)]
            prior_b2 = Normal(tf.zeros([]), sd).log_prob(b2)
        log_p = (log_l * n_train / n_batch + Gamma(a, b).log_prob(lam) + Gamma(a, b).log_prob(gam) +
                 tf.reduce_sum(priors[0]) + tf.reduce_sum(priors[1]) + tf.reduce_sum(priors[2]) + prior_b2) / n_train
    return dict(X=X, y=y, vars=[log_lambda, log_gamma, w1, b1, w2, b2], pred=pred, log_p=log_p)


def test_names_shapes_and_collections_follow_tf1():
    tf.reset_default_graph()
    with tf.variable_scope("model"):
        vs = [tf.Variable(tf.zeros([2, 3])) for _ in range(12)]
        ph = tf.placeholder(tf.float32, shape=[None, 7])
        with tf.variable_scope("inner"):
            inner = tf.Variable(tf.zeros([]))
    outside = tf.Variable(tf.zeros([4]))
    assert [v.name for v in vs[:3]] == ["model/Variable:0", "model/Variable_1:0", "model/Variable_2:0"]
    assert inner.name == "model/inner/Variable:0" and outside.name == "Variable:0"
    # abstract_stein_sampler.py:49-51: trainable variables of the scope, creation order
    got = tf.get_collection(tf.GraphKeys.TRAINABLE_VARIABLES, "model")
    assert got == vs + [inner]
    # converters.py:40: the name sort puts Variable_10 before Variable_2
    names = sorted(v.name for v in vs)
    assert names.index("model/Variable_10:0") < names.index("model/Variable_2:0")
    assert vs[0].get_shape().as_list() == [2, 3] and inner.get_shape().as_list() == []
    assert ph.shape[1].value == 7 and ph.shape[0].value is None and ph.get_shape().as_list() == [None, 7]
    with pytest.raises(ValueError):
        tf.matmul(vs[0], vs[1])                       # 3 != 2
    with pytest.raises(ValueError):
        tf.nn.sigmoid_cross_entropy_with_logits(vs[0], vs[1])
    with pytest.raises(ValueError):
        tf.Session().run(ph + 1.0)                    # placeholder not fed


def test_session_run_and_gradients():
    g = linear_graph(3)
    rng = np.random.default_rng(0)
    Xb, yb, w0 = rng.standard_normal((9, 3)), rng.standard_normal((9, 1)), rng.standard_normal((3, 1))
    g["w"].load(w0)
    sess = tf.Session()
    val, grad = sess.run([g["log_p"], tf.gradients(g["log_p"], [g["w"]])[0]], {g["X"]: Xb, g["y"]: yb})
    ref = -0.5 * ((Xb @ w0 - yb) ** 2).sum() + (-0.5 * w0 ** 2 - 0.5 * np.log(2 * np.pi)).sum()
    assert abs(val - ref) < 1e-4 * abs(ref)
    assert _rel(grad.ravel(), orc.score_linear(w0.T, Xb, yb).ravel()) < 1e-5
    # stop_gradient, reduce_mean, stack, transpose, reshape, top_k: the symbols of the kernel graph
    t = tf.stack([tf.reshape(tf.transpose(g["w"]), [3]), tf.stop_gradient(tf.reshape(g["w"], [3])) * 2.0])
    vals, _ = tf.nn.top_k(tf.reshape(t, [-1]), k=4)
    out = sess.run([tf.reduce_mean(t, axis=0), vals])
    np.testing.assert_allclose(out[0], 1.5 * w0.ravel(), rtol=1e-6)
    np.testing.assert_allclose(out[1], np.sort(np.concatenate([w0.ravel(), 2 * w0.ravel()]))[::-1][:4], rtol=1e-6)


@pytest.mark.parametrize("n,F,N", [(7, 1, 40), (5, 10, 64)])
def test_linear_graph_scores_match_oracle(n, F, N):
    g = linear_graph(F)
    model = GraphLogPosterior(g["log_p"])
    assert model.model_vars == [g["w"]] and model.n_params == F
    rng = np.random.default_rng(1)
    theta = rng.standard_normal((n, F)).astype(np.float32)
    Xb, yb = rng.standard_normal((N, F)).astype(np.float32), rng.standard_normal((N, 1)).astype(np.float32)
    feed = model._feed_tensors({g["X"]: Xb, g["y"]: yb}, torch.as_tensor)
    S = model.graph_scores(torch.as_tensor(theta), feed).numpy()
    assert _rel(S, orc.score_linear(theta, Xb, yb)) < 2e-5
    yh = model.graph_values(g["y_hat"], torch.as_tensor(theta), {g["X"]: feed[g["X"]]}).numpy()
    assert yh.shape == (n, N) and _rel(yh, theta.astype(np.float64) @ Xb.T) < 1e-5


def test_logistic_graph_scores_match_oracle():
    F, B, N, n = 54, 50, 464809, 6
    g = logistic_graph(F, N, B)
    model = GraphLogPosterior(g["log_p"])
    # column order = name sort = [w (F), log_alpha]
    sl = model.column_slices()
    assert sl[g["w"]] == (0, F) and sl[g["log_alpha"]] == (F, F + 1)
    rng = np.random.default_rng(2)
    theta = (rng.standard_normal((n, F + 1)) * 0.3).astype(np.float32)
    Xb = rng.standard_normal((B, F)).astype(np.float32)
    yb = (rng.random((B, 1)) > 0.5).astype(np.float32)
    feed = model._feed_tensors({g["X"]: Xb, g["y"]: yb}, torch.as_tensor)
    S = model.graph_scores(torch.as_tensor(theta), feed).numpy()
    assert _rel(S, orc.score_logistic(theta, Xb, yb, N)) < 5e-5
    logits = model.graph_values(g["logits"], torch.as_tensor(theta), feed).numpy()
    assert _rel(logits, theta[:, :F].astype(np.float64) @ Xb.T) < 1e-5


def test_bnn_graph_scores_match_oracle():
    F, H, B, N, n = 13, 50, 20, 506, 4
    g = bnn_graph(F, H, N, B)
    model = GraphLogPosterior(g["log_p"])
    assert model.n_params == 2 + F * H + 2 * H + 1
    # the dict <-> array converters accept the stand-in's variables (name, get_shape().as_list())
    rng = np.random.default_rng(3)
    theta_dict = {v: rng.standard_normal([n] + v.get_shape().as_list()) * 0.3 for v in model.model_vars}
    theta, access = convert_dictionary_to_array(theta_dict)
    assert {v: tuple(ix) for v, ix in access.items()} == {v: tuple(ix) for v, ix in model.column_slices().items()}
    theta = theta.astype(np.float32)
    Xb = rng.standard_normal((B, F)).astype(np.float32)
    yb = rng.standard_normal((B, 1)).astype(np.float32)
    feed = model._feed_tensors({g["X"]: Xb, g["y"]: yb}, torch.as_tensor)
    S = model.graph_scores(torch.as_tensor(theta), feed).numpy()
    assert _rel(S, orc.score_bnn(theta, Xb, yb, N, F, H)) < 5e-5
    pred = model.graph_values(g["pred"], torch.as_tensor(theta), {g["X"]: feed[g["X"]]}).numpy()
    assert _rel(pred, orc.bnn_predict(theta, Xb, F, H)) < 1e-5


def test_sampler_rejects_other_objects_and_graphs_without_model_scope():
    from stein_b200.samplers import SteinSampler
    from stein_b200.optimizers import AdamGradientDescent
    with pytest.raises(TypeError):
        SteinSampler(4, object(), AdamGradientDescent())
    tf.reset_default_graph()
    v = tf.Variable(tf.zeros([2]))            # not under "model"
    with pytest.raises(ValueError):
        GraphLogPosterior(tf.reduce_sum(v))


# --------------------------------------------------------------------------- #
# the reference's own example scripts, unmodified, up to the sampler call       #
# --------------------------------------------------------------------------- #
REFERENCE = "/root/reference/examples"


class _Captured(Exception):
    pass


def _run_reference_script(monkeypatch, example, extra_modules=()):
    """Runs examples/<example>/main.py of the reference AS IS (stand-in tensorflow, `stein`
    alias package) with a SteinSampler that stops the script at its constructor and hands back
    (n_particles, log_p, gd).  Needs the reference tree: skipped where it does not exist."""
    import runpy
    import types
    script = os.path.join(REFERENCE, example, "main.py")
    if not os.path.exists(script):
        pytest.skip("reference tree not present")
    import stein.samplers as samplers_mod
    got = {}

    class StopAtSampler:
        def __init__(self, n_particles, log_p, gd, theta=None):
            got.update(n_particles=n_particles, log_p=log_p, gd=gd)
            raise _Captured()

    monkeypatch.setattr(samplers_mod, "SteinSampler", StopAtSampler)
    plt = types.ModuleType("matplotlib.pyplot")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    for name, mod in extra_modules:
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.chdir(os.path.dirname(script))
    tf.reset_default_graph()
    np.random.seed(0)
    with pytest.raises(_Captured):
        runpy.run_path(script, run_name="__main__")
    return got


def test_reference_linear_example_builds_its_graph_on_the_stand_in(monkeypatch, golden_dir):
    got = _run_reference_script(monkeypatch, "linear_regression")
    assert got["n_particles"] == 50 and got["gd"].learning_rate == 1e-1
    model = GraphLogPosterior(got["log_p"])
    assert [v.name for v in model.model_vars] == ["model/Variable:0"]
    g = np.load(os.path.join(golden_dir, "linear_regression.npz"))       # the same CSVs
    X, y = g["X"].astype(np.float32), g["y"].reshape(-1, 1).astype(np.float32)
    phs = [t for t in _placeholders(got["log_p"])]
    assert [p.get_shape().as_list() for p in phs] == [[None, 1], [None, 1]]
    theta = np.random.default_rng(5).standard_normal((6, 1)).astype(np.float32)
    feed = model._feed_tensors({phs[0]: X, phs[1]: y}, torch.as_tensor)
    S = model.graph_scores(torch.as_tensor(theta), feed).numpy()
    assert _rel(S, orc.score_linear(theta, X, y)) < 2e-5


def test_reference_bnn_example_builds_its_graph_on_the_stand_in(monkeypatch):
    got = _run_reference_script(monkeypatch, "regression_neural_network")
    assert got["n_particles"] == 20 and got["gd"].decay == 0.999
    model = GraphLogPosterior(got["log_p"])
    F, H, B, N = 1, 100, 20, 20                       # main.py:12-20
    assert model.n_params == 2 + F * H + 2 * H + 1
    phs = _placeholders(got["log_p"])
    rng = np.random.default_rng(6)
    theta = (rng.standard_normal((3, model.n_params)) * 0.3).astype(np.float32)
    Xb, yb = rng.random((B, F)).astype(np.float32), rng.standard_normal((B, 1)).astype(np.float32)
    feed = model._feed_tensors({phs[0]: Xb, phs[1]: yb}, torch.as_tensor)
    S = model.graph_scores(torch.as_tensor(theta), feed).numpy()
    assert _rel(S, orc.score_bnn(theta, Xb, yb, N, F, H)) < 5e-5


def test_reference_logistic_example_builds_its_graph_on_the_stand_in(monkeypatch):
    import types
    import scipy.io
    rng = np.random.default_rng(7)
    N_all, F = 500, 54                                 # covertype.mat is not shipped: same layout, synthetic
    data = np.concatenate([rng.integers(1, 3, size=(N_all, 1)).astype(np.float64),
                           rng.standard_normal((N_all, F))], axis=1)
    fake_io = types.ModuleType("scipy.io")
    fake_io.loadmat = lambda path: {"covtype": data.copy()}
    monkeypatch.setattr(scipy, "io", fake_io)
    got = _run_reference_script(monkeypatch, "logistic_regression", extra_modules=[("scipy.io", fake_io)])
    assert got["n_particles"] == 100
    model = GraphLogPosterior(got["log_p"])
    assert model.n_params == F + 1
    phs = _placeholders(got["log_p"])
    n_train, B = 400, 50                               # 80 % split of 500; main.py:18
    theta = (rng.standard_normal((4, F + 1)) * 0.3).astype(np.float32)
    Xb = rng.standard_normal((B, F)).astype(np.float32)
    yb = (rng.random((B, 1)) > 0.5).astype(np.float32)
    feed = model._feed_tensors({phs[0]: Xb, phs[1]: yb}, torch.as_tensor)
    S = model.graph_scores(torch.as_tensor(theta), feed).numpy()
    assert _rel(S, orc.score_logistic(theta, Xb, yb, n_train)) < 5e-5


def _placeholders(t):
    """Placeholders below a tensor, in creation (name) order."""
    seen, out, todo = set(), [], [t]
    while todo:
        x = todo.pop()
        if id(x) in seen:
            continue
        seen.add(id(x))
        if x.op_type == "Placeholder":
            out.append(x)
        todo.extend(x.inputs)
    key = lambda p: (len(p.name), p.name)
    return sorted(out, key=key)


def test_reference_library_reproduces_its_recorded_run(golden_dir):
    """tests/golden/reference_run.npz is reproducible: the reference's SquaredExponentialKernel and
    compute_median, imported unmodified and run on the stand-in, give the recorded values again
    (guards the stand-in; skipped where the reference tree does not exist)."""
    import importlib
    import importlib.util
    root = "/root/reference/stein"
    if not os.path.exists(os.path.join(root, "__init__.py")):
        pytest.skip("reference tree not present")
    spec = importlib.util.spec_from_file_location("_ref_stein_t", os.path.join(root, "__init__.py"),
                                                  submodule_search_locations=[root])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["_ref_stein_t"] = pkg
    spec.loader.exec_module(pkg)
    kernels = importlib.import_module("_ref_stein_t.kernels")
    utilities = importlib.import_module("_ref_stein_t.utilities")
    ref = np.load(os.path.join(golden_dir, "reference_run.npz"))
    for i, (n, d) in enumerate(ref["kernel_shapes"][:4]):
        tf.reset_default_graph()
        sess = tf.Session()
        k = kernels.SquaredExponentialKernel(int(n), sess)
        K, dK = k.kernel_and_grad(ref["kernel%d_theta" % i])
        np.testing.assert_allclose(K, ref["kernel%d_K" % i], rtol=0, atol=1e-7)
        np.testing.assert_allclose(dK, ref["kernel%d_dK" % i], rtol=1e-6, atol=1e-9)
    for i in range(int(ref["n_median"])):
        tf.reset_default_graph()
        m = utilities.compute_median(tf.constant(ref["median%d_in" % i]))
        assert np.float32(tf.Session().run(m)).tobytes() == np.float32(ref["median%d_out" % i]).tobytes()
