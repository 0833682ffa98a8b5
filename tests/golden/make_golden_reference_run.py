"""Generates tests/golden/reference_run.npz by RUNNING THE REFERENCE'S OWN LIBRARY CODE.

/root/reference/stein (kernels, compute_median, samplers, optimizers, converters) is imported
unmodified under the alias `_ref_stein`, with `tensorflow` resolved to the TF1 stand-in of
compat/ (graph recording; the ops are evaluated by PyTorch on the CPU in float32).  What runs is
the reference's Python: its distance / median / bandwidth graph
(stein/kernels/abstract_kernel.py:30-40, stein/utilities/compute_median.py:4-16), K and the
tf.gradients-based dK with the -0.5 post-scale (stein/kernels/squared_exponential_kernel.py:22-35),
compute_phi / clip / update (stein/samplers/abstract_stein_sampler.py:76-127), the per-particle
score loop (stein/samplers/stein_sampler.py:50-71) and both optimizers.  What does NOT run is
TensorFlow's own kernels: results agree with a real TF 1.12 run up to float32 rounding of the
individual ops, which is why the tests that read this file use tolerances, not bit equality.

Run once in the build container (needs /root/reference); the output is committed and the tests
never read /root/reference.

    python tests/golden/make_golden_reference_run.py
"""
import importlib
import importlib.util
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/stein"
sys.path[:0] = [os.path.join(REPO, "compat"), REPO, os.path.join(REPO, "tests")]

import tensorflow as tf  # noqa: E402  (the stand-in)

from test_tf_compat import bnn_graph, linear_graph, logistic_graph  # noqa: E402


def load_reference():
    spec = importlib.util.spec_from_file_location("_ref_stein", os.path.join(REF, "__init__.py"),
                                                  submodule_search_locations=[REF])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["_ref_stein"] = pkg
    spec.loader.exec_module(pkg)
    return {m: importlib.import_module("_ref_stein." + m) for m in ("kernels", "samplers", "optimizers", "utilities")}


def main():
    ref = load_reference()
    out = {}
    rng = np.random.default_rng(20261018)

    # ---- kernel operator: K, dK, bandwidth ---------------------------------------------------
    shapes = [(2, 1), (7, 3), (20, 5), (33, 10), (50, 1), (64, 55)]
    out["kernel_shapes"] = np.array(shapes)
    for i, (n, d) in enumerate(shapes):
        tf.reset_default_graph()
        sess = tf.Session()
        k = ref["kernels"].SquaredExponentialKernel(n, sess)
        theta = rng.standard_normal((n, d)) * (0.01 if i % 2 else 1.0)
        K, dK = k.kernel_and_grad(theta)
        feed = {k.theta[j]: theta[j] for j in range(n)}
        bw, D = sess.run([k.bandwidth, k.D], feed)
        out["kernel%d_theta" % i], out["kernel%d_K" % i], out["kernel%d_dK" % i] = theta, K, dK
        out["kernel%d_bandwidth" % i], out["kernel%d_D" % i] = bw, D

    # ---- ties and dynamic range (read by the CPU oracle test only) ------------------------------
    rng_t = np.random.default_rng(77)        # own stream: the draws of the other cases do not move
    base = rng_t.standard_normal((10, 4))
    ties = [np.repeat(base, 2, axis=0),                                        # duplicated particles
            np.concatenate([np.zeros((5, 3)), np.ones((5, 3))]),               # two tight clusters: middle values 0 and 3
            np.concatenate([np.zeros((5, 3)), np.ones((6, 3))]),               # degenerate: median 0 -> bandwidth 0 -> NaN
            rng_t.standard_normal((21, 6)) * np.logspace(-2, 2, 21)[:, None]]    # wide range of norms, odd n*n
    out["n_ties"] = np.array(len(ties))
    for i, theta in enumerate(ties):
        n = theta.shape[0]
        tf.reset_default_graph()
        sess = tf.Session()
        k = ref["kernels"].SquaredExponentialKernel(n, sess)
        K, dK = k.kernel_and_grad(theta)
        bw = sess.run(k.bandwidth, {k.theta[j]: theta[j] for j in range(n)})
        out["tie%d_theta" % i], out["tie%d_K" % i], out["tie%d_dK" % i], out["tie%d_bandwidth" % i] = theta, K, dK, bw

    # ---- compute_median on explicit inputs (even / odd length, ties) ----------------------------
    meds = [np.array([[3., 1.], [2., 5.]]), np.arange(9, dtype=np.float64).reshape(3, 3)[::-1] * 0.5,
            np.array([[0., 0., 1., 1.]]), rng.standard_normal((5, 5)), rng.standard_normal((6, 6))]
    out["n_median"] = np.array(len(meds))
    for i, D in enumerate(meds):
        tf.reset_default_graph()
        m = ref["utilities"].compute_median(tf.constant(D.astype(np.float32)))
        out["median%d_in" % i], out["median%d_out" % i] = D.astype(np.float32), tf.Session().run(m)

    # ---- trajectories -----------------------------------------------------------------------------
    def run(tag, graph, n_particles, gd, feeds, posterior=None):
        np.random.seed(zlib.crc32(tag.encode()))
        s = ref["samplers"].SteinSampler(n_particles, graph["log_p"], gd)
        traj = [s.samples.copy()]
        phi0 = None
        for it, feed in enumerate(feeds):
            if it == 0:   # phi of the first iteration, through the reference's compute_phi
                grads = {v: np.zeros([n_particles] + v.get_shape().as_list()) for v in s.model_vars}
                for p in range(n_particles):
                    tfeed = {v: s.theta[v][p] for v in s.model_vars}
                    tfeed.update(feed)
                    for v, gval in zip(s.model_vars, s.sess.run(s.grad_log_p, tfeed)):
                        grads[v][p] = gval
                S0 = ref["utilities"].convert_dictionary_to_array(grads)[0]
                phi0 = s.compute_phi(s.samples, S0)
                out[tag + "_scores0"], out[tag + "_phi0"] = S0, phi0
            s.train_on_batch(feed)
            traj.append(s.samples.copy())
        out[tag + "_traj"] = np.stack(traj)
        if posterior is not None:
            out[tag + "_posterior"] = s.function_posterior(posterior[0], posterior[1])
            out[tag + "_posterior_mean"] = s.function_posterior(posterior[0], posterior[1], axis=0)

    Adam, Adagrad = ref["optimizers"].AdamGradientDescent, ref["optimizers"].AdagradGradientDescent

    lin = np.load(os.path.join(HERE, "linear_regression.npz"))           # the reference's CSVs
    X, y = lin["X"], lin["y"].reshape(-1, 1)
    g = linear_graph(X.shape[1])
    run("linear", g, 50, Adam(learning_rate=1e-1), [{g["X"]: X, g["y"]: y}] * 6,
        posterior=(g["y_hat"], {g["X"]: X[:7]}))

    F, N, B = 6, 200, 10
    Xl = rng.standard_normal((N, F))
    yl = (rng.random((N, 1)) < 1.0 / (1.0 + np.exp(-Xl @ rng.standard_normal((F, 1))))).astype(np.float64)
    idx = [rng.choice(N, B, replace=False) for _ in range(5)]
    out["logistic_X"], out["logistic_y"], out["logistic_batches"] = Xl, yl, np.stack(idx)
    g = logistic_graph(F, N, B)
    run("logistic", g, 16, Adam(learning_rate=1e-1), [{g["X"]: Xl[i], g["y"]: yl[i]} for i in idx],
        posterior=(g["logits"], {g["X"]: Xl[:9], g["y"]: yl[:9]}))

    F, H, N, B = 2, 5, 40, 8
    Xb = rng.random((N, F))
    yb = np.cos(3 * Xb[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    idx = [rng.choice(N, B, replace=False) for _ in range(5)]
    out["bnn_X"], out["bnn_y"], out["bnn_batches"] = Xb, yb, np.stack(idx)
    g = bnn_graph(F, H, N, B)
    run("bnn_adam", g, 12, Adam(learning_rate=1e-1, decay=0.999), [{g["X"]: Xb[i], g["y"]: yb[i]} for i in idx],
        posterior=(g["pred"], {g["X"]: Xb[:6]}))
    g = bnn_graph(F, H, N, B)
    run("bnn_adagrad", g, 12, Adagrad(learning_rate=5e-2, decay=0.5, alpha=0.9),
        [{g["X"]: Xb[i], g["y"]: yb[i]} for i in idx])

    path = os.path.join(HERE, "reference_run.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f KiB)" % (path, len(out), os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
