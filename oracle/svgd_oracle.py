"""oracle/svgd_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of the reference's SVGD iteration (JamesBrofos/Stein), one
function per reference function, each citing the file:line it follows.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; nothing under ``stein_b200/``
does.

Pinning.  The reference ships no tests and no golden vectors, and TensorFlow
1.12 cannot run here.  Pinned against THE REFERENCE'S OWN CODE, run in this
container with the scripts under ``tests/golden/`` (outputs committed there):
  * the two optimizers and the dict<->array converters (pure NumPy in the
    reference): ``make_golden.py`` -> ``optimizers.npz``, ``converters.npz``;
  * the whole library -- kernels (D, top_k median, bandwidth, K, the
    tf.gradients-based dK and its -0.5 post-scale), compute_phi, the clip, the
    per-particle score loop, trajectories, function_posterior -- imported
    unmodified from /root/reference/stein and executed on the TF1 graph-API
    stand-in of ``compat/`` (graph recorded, ops evaluated by PyTorch in fp32):
    ``make_golden_reference_run.py`` -> ``reference_run.npz``, checked by
    ``tests/test_reference_run.py`` to 1e-5..5e-5 (median rule: same bits);
  * the linear-regression example's shipped data, whose analytic posterior is
    the known answer.
STILL UNPINNED: the float32 rounding of TensorFlow's own op kernels (SGEMM
accumulation order, exp) -- the stand-in is not TensorFlow -- so "bit-exact
median" is defined on the contract arithmetic below, not on a TF 1.12 run.

Two flavours of the distance matrix are offered:
  * ``sqdist``        -- the literal NumPy line (BLAS sgemm; summation order is
                         whatever the BLAS does), used for the "what would the
                         reference's CPU path cost" baseline;
  * ``sqdist_chain``  -- the contract arithmetic of oracle/svgd_oracle.c (fixed
                         fma order), the one the CUDA path reproduces bit for
                         bit, used for the median parity tests.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile oracle/svgd_oracle.c -> oracle/libsvgd_oracle.so (gcc, OpenMP)."""
    so = os.path.join(_HERE, "libsvgd_oracle.so")
    src = os.path.join(_HERE, "svgd_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def clib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        i64 = ctypes.c_int64
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_row_norms.argtypes = [fp, i64, i64, fp]
        L.oracle_sqdist_chain.argtypes = [fp, i64, i64, fp]
        L.oracle_median_chain.argtypes = [fp, i64, i64, fp]
        L.oracle_median_chain.restype = ctypes.c_float
        L.oracle_median_chain_radix.argtypes = [fp, i64, i64, fp]
        L.oracle_median_chain_radix.restype = ctypes.c_float
        L.oracle_bandwidth.argtypes = [ctypes.c_float, i64]
        L.oracle_bandwidth.restype = ctypes.c_float
        L.oracle_phi_rows.argtypes = [fp, fp, i64, i64, ctypes.c_float, i64, i64, dp, fp]
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# --------------------------------------------------------------------------- #
# distance matrix, median, bandwidth                                          #
# --------------------------------------------------------------------------- #
def sqdist(theta):
    """stein/kernels/abstract_kernel.py:33-35.  Placeholders are tf.float32, so
    the float64 particles are down-cast at the feed (abstract_kernel.py:31)."""
    T = _f32(theta)
    r = np.sum(T * T, axis=1).reshape(-1, 1)
    return r + r.T - 2.0 * (T @ T.T)


def sqdist_chain(theta):
    """abstract_kernel.py:33-35 in the contract arithmetic (svgd_oracle.c)."""
    T = _f32(theta)
    n, d = T.shape
    D = np.empty((n, n), np.float32)
    clib().oracle_sqdist_chain(_fp(T), n, d, _fp(D))
    return D


def row_norms_chain(theta):
    T = _f32(theta)
    r = np.empty(T.shape[0], np.float32)
    clib().oracle_row_norms(_fp(T), T.shape[0], T.shape[1], _fp(r))
    return r


def compute_median(D):
    """stein/utilities/compute_median.py:4-16.  top_k(V, dim//2+1): the last
    value is the (dim//2+1)-th largest; even dim -> mean of the last two."""
    V = np.asarray(D, dtype=np.float32).reshape(-1)
    dim = V.shape[0]
    m = dim // 2 + 1
    # m-th largest == (dim-m)-th smallest (0-based)
    if dim % 2 == 0:
        part = np.partition(V, [dim - m, dim - m + 1])
        two = part[[dim - m + 1, dim - m]]          # values[m-2:], descending
        return np.float32((two[0] + two[1]) / np.float32(2.0))
    return np.float32(np.partition(V, dim - m)[dim - m])


def median_chain(theta, radix=False):
    """Median of the contract-arithmetic D without NumPy in the loop.  Returns
    (median, (lower_middle, upper_middle))."""
    T = _f32(theta)
    mid = np.zeros(2, np.float32)
    fn = clib().oracle_median_chain_radix if radix else clib().oracle_median_chain
    med = fn(_fp(T), T.shape[0], T.shape[1], _fp(mid))
    return np.float32(med), (np.float32(mid[0]), np.float32(mid[1]))


def bandwidth(med, n_particles):
    """abstract_kernel.py:40: sqrt(m / np.log(n)); fp32 graph arithmetic."""
    return np.sqrt(np.float32(med) / np.float32(np.log(n_particles)), dtype=np.float32)


# --------------------------------------------------------------------------- #
# kernel, its gradient, phi                                                   #
# --------------------------------------------------------------------------- #
def kernel_and_grad(theta, chain=True):
    """stein/kernels/squared_exponential_kernel.py:22-35.

    K  = exp(-D / square(bandwidth) / 2)                         (:22)
    dK = -0.5 * vstack(tf.gradients(K, theta_i))                 (:23, :32)
       = (x_i * sum_j K_ij - sum_j K_ij x_j) / h^2   (bandwidth is
         stop_gradient'ed, abstract_kernel.py:40; derivation SURVEY.md sec. 0)
    Returns (K fp32 n x n, dK fp32 n x d, bandwidth fp32)."""
    T = _f32(theta)
    n = T.shape[0]
    D = sqdist_chain(T) if chain else sqdist(T)
    bw = bandwidth(compute_median(D), n)
    h2 = np.float32(bw * bw)
    K = np.exp(-D / h2 / np.float32(2.0)).astype(np.float32)
    dK = ((T * K.sum(axis=1, dtype=np.float32)[:, None] - K @ T) / h2).astype(np.float32)
    return K, dK, bw


def imq_kernel_and_grad(theta, beta=-0.5, chain=True):
    """An inverse multiquadric operator behind the reference's plugin point
    (stein/kernels/abstract_kernel.py:45-62), built with the reference's recipes: the distance matrix and
    bandwidth of abstract_kernel.py:33-40, K = (1 + D / h^2)^beta in fp32, and
    dK = -0.5 * d(sum K)/d theta_i (squared_exponential_kernel.py:23,32)
       = (-2 beta / h^2) (x_i sum_j G_ij - sum_j G_ij x_j),  G = (1 + D / h^2)^(beta - 1).
    Returns (K fp32, dK fp32, bandwidth fp32).  tests/test_oracle.py checks the closed form of dK
    against torch autograd of sum(K)."""
    T = _f32(theta)
    D = sqdist_chain(T) if chain else sqdist(T)
    h = bandwidth(compute_median(D), T.shape[0])
    h2 = np.float32(h) * np.float32(h)
    base = np.float32(1.0) + np.maximum(D, np.float32(0.0)) / h2
    K = np.power(base, np.float32(beta)).astype(np.float32)
    G = np.power(base, np.float32(beta - 1.0)).astype(np.float32)
    coef = np.float32(-2.0 * beta) / h2
    dK = ((T * G.sum(axis=1, dtype=np.float32)[:, None] - G @ T) * coef).astype(np.float32)
    return K, dK, h


def compute_phi(theta_array, grads_array, chain=True):
    """stein/samplers/abstract_stein_sampler.py:100-105: float64 NumPy GEMM of
    the fp32 K against the (float64) score matrix, plus the fp32 dK, over n."""
    n_particles = grads_array.shape[0]
    K, dK, _ = kernel_and_grad(theta_array, chain=chain)
    return (K.dot(np.asarray(grads_array, dtype=np.float64)) + dK) / n_particles


def phi_rows_c(theta, grads, bw, i0, i1):
    """Row block [i0, i1) of compute_phi through the C oracle (no n x n)."""
    T = _f32(theta)
    S = _f32(grads)
    n, d = T.shape
    out = np.empty((i1 - i0, d), np.float64)
    ks = np.empty(i1 - i0, np.float32)
    clib().oracle_phi_rows(_fp(T), _fp(S), n, d, ctypes.c_float(float(bw)), i0, i1,
                           out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), _fp(ks))
    return out, ks


def clip(phi):
    """abstract_stein_sampler.py:125."""
    return phi * (10.0 / max(10.0, np.linalg.norm(phi)))


# --------------------------------------------------------------------------- #
# optimizers (stein/optimizers/*.py), restated                                #
# --------------------------------------------------------------------------- #
class AdamGradientDescent:
    """stein/optimizers/adam_gradient_descent.py:15-58 (first call sets mu=phi,
    nu=phi**2; eps outside the sqrt; learning-rate decay after the step)."""

    def __init__(self, learning_rate=1e-3, decay=1., beta_1=0.9, beta_2=0.999):
        self.learning_rate, self.decay, self.n_iters = learning_rate, decay, 0
        self.beta_1, self.beta_2 = beta_1, beta_2

    def update(self, phi):
        if self.n_iters == 0:
            self.mu, self.nu = phi, phi ** 2
        else:
            self.mu = self.beta_1 * self.mu + (1. - self.beta_1) * phi
            self.nu = self.beta_2 * self.nu + (1. - self.beta_2) * phi ** 2
        self.n_iters += 1
        mup = self.mu / (1. - self.beta_1 ** self.n_iters)
        nup = self.nu / (1. - self.beta_2 ** self.n_iters)
        grad = mup / (1e-8 + np.sqrt(nup)) * self.learning_rate
        self.learning_rate *= self.decay
        return grad


class AdagradGradientDescent:
    """stein/optimizers/adagrad_gradient_descent.py:13-44 (`decay` is stored but
    never applied; no bias correction)."""

    def __init__(self, learning_rate=1e-3, decay=1., alpha=0.9):
        self.learning_rate, self.decay, self.n_iters = learning_rate, decay, 0
        self.alpha = alpha

    def update(self, phi):
        if self.n_iters == 0:
            self.hist = phi ** 2
        else:
            self.hist = self.alpha * self.hist + (1. - self.alpha) * phi ** 2
        self.n_iters += 1
        return phi / (1e-6 + np.sqrt(self.hist)) * self.learning_rate


def update_particles(theta_array, grads_array, gd, chain=True):
    """abstract_stein_sampler.py:121-127 on the flat (n x d) float64 layout.
    Returns (new theta_array, clipped phi)."""
    phi = compute_phi(theta_array, grads_array, chain=chain)
    phi = clip(phi)
    return theta_array + gd.update(phi), phi


# --------------------------------------------------------------------------- #
# dict <-> array layout (stein/utilities/converters.py)                       #
# --------------------------------------------------------------------------- #
def convert_dictionary_to_array(dictionary):
    """converters.py:30-55: keys sorted by `.name`, each value reshaped to
    (n, prod(shape[1:])) row-major and concatenated along columns."""
    n_particles = next(iter(dictionary.values())).shape[0]
    n_params = sum(int(np.prod(v.shape[1:])) for v in dictionary.values())
    array = np.zeros((n_particles, n_params))
    access, index = {}, 0
    for v in sorted(dictionary.keys(), key=lambda x: x.name):
        value = dictionary[v]
        dim = int(np.prod(value.shape[1:]))
        array[:, index:index + dim] = np.reshape(value, (n_particles, dim))
        access[v] = (index, index + dim)
        index += dim
    return array, access


def convert_array_to_dictionary(array, access_indices):
    """converters.py:77-89."""
    n = array.shape[0]
    return {v: np.reshape(array[:, a:b], [n] + v.get_shape().as_list())
            for v, (a, b) in access_indices.items()}


# --------------------------------------------------------------------------- #
# closed-form scores of the three example models, all particles#
# at once, float64.  W etc. are (n, ...) stacks in the flat layout of A.2.     #
# --------------------------------------------------------------------------- #
def score_linear(theta, X, y):
    """examples/linear_regression/main.py:25-31.
    log p = -0.5 ||Xw - y||^2 + sum N(w; 0, 1)  =>  grad = -X^T (Xw - y) - w."""
    W = np.asarray(theta, np.float64)                      # (n, F)
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(1, -1)
    R = W @ X.T - y                                        # (n, N)
    return -(R @ X) - W


def score_logistic(theta, X, y, n_train, a=1.0, b=0.01):
    """examples/logistic_regression/main.py:28-49; layout [w (F), log_alpha]."""
    th = np.asarray(theta, np.float64)
    W, la = th[:, :-1], th[:, -1]
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(1, -1)
    F = W.shape[1]
    alpha = np.exp(la)
    scale = float(n_train) / X.shape[0]
    Z = W @ X.T
    sig = 1.0 / (1.0 + np.exp(-Z))
    gw = scale * ((y - sig) @ X) - alpha[:, None] * W
    gla = 0.5 * F - 0.5 * alpha * (W * W).sum(1) + (a - 1.0) - b * alpha
    return np.concatenate([gw, gla[:, None]], axis=1)


def score_gmm(theta, means=None, sigma2=1.0):
    """Score of the synthetic targets of BASELINE.json configs D / E: an
    equal-weight mixture of isotropic Gaussians N(mu_k, sigma2 I),
        S_i = sum_k r_ik (mu_k - x_i) / sigma2,   r_ik = softmax_k(-|x_i - mu_k|^2 / (2 sigma2)).
    means=None: the standard normal target, S = -X / sigma2.  float64 throughout (a closed form,
    not a reference function: the reference has no such model; the role is that of the per-particle
    tf.gradients(log_p) loop of stein/samplers/stein_sampler.py:59-68)."""
    X = np.asarray(theta, np.float64)
    if means is None:
        return -X / sigma2
    M = np.asarray(means, np.float64)
    d2 = ((X[:, None, :] - M[None, :, :]) ** 2).sum(-1)
    logit = -d2 / (2.0 * sigma2)
    logit -= logit.max(axis=1, keepdims=True)
    R = np.exp(logit)
    R /= R.sum(axis=1, keepdims=True)
    return (R @ M - X) / sigma2


def bnn_unpack(theta, F, H):
    """Flat layout of examples/regression_neural_network/main.py:35-42 under the
    name sort of converters.py:40: [log_lambda, log_gamma, w1 (F*H), b1 (H),
    w2 (H), b2]."""
    th = np.asarray(theta, np.float64)
    n = th.shape[0]
    o = 2
    w1 = th[:, o:o + F * H].reshape(n, F, H); o += F * H
    b1 = th[:, o:o + H]; o += H
    w2 = th[:, o:o + H]; o += H
    b2 = th[:, o]
    return th[:, 0], th[:, 1], w1, b1, w2, b2


def bnn_predict(theta, X, F, H):
    """examples/regression_neural_network/main.py:46-48, per particle: (n, B)."""
    _, _, w1, b1, w2, b2 = bnn_unpack(theta, F, H)
    X = np.asarray(X, np.float64)
    Z = np.einsum("bf,nfh->nbh", X, w1) + b1[:, None, :]
    A = np.maximum(Z, 0.0)
    return np.einsum("nbh,nh->nb", A, w2) + b2[:, None]


def score_bnn(theta, X, y, n_train, F, H, a=1.0, b=0.01):
    """examples/regression_neural_network/main.py:35-85, closed form."""
    ll, lg, w1, b1, w2, b2 = bnn_unpack(theta, F, H)
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(1, -1)
    B = X.shape[0]
    N = float(n_train)
    lam, gam = np.exp(ll), np.exp(lg)
    Z = np.einsum("bf,nfh->nbh", X, w1) + b1[:, None, :]
    A = np.maximum(Z, 0.0)
    pred = np.einsum("nbh,nh->nb", A, w2) + b2[:, None]
    res = y - pred                                          # (n, B)
    delta = (N / B) * gam[:, None] * res
    g_w2 = np.einsum("nbh,nb->nh", A, delta) - lam[:, None] * w2
    g_b2 = delta.sum(1) - lam * b2
    dZ = delta[:, :, None] * w2[:, None, :] * (Z > 0)
    g_w1 = np.einsum("bf,nbh->nfh", X, dZ) - lam[:, None, None] * w1
    g_b1 = dZ.sum(1) - lam[:, None] * b1
    g_lg = (N / B) * (0.5 * B - 0.5 * gam * (res ** 2).sum(1)) + (a - 1.0) - b * gam
    Pw = F * H + 2 * H + 1
    sumsq = (w1 ** 2).sum((1, 2)) + (b1 ** 2).sum(1) + (w2 ** 2).sum(1) + b2 ** 2
    g_ll = 0.5 * Pw - 0.5 * lam * sumsq + (a - 1.0) - b * lam
    n = theta.shape[0]
    out = np.concatenate([g_ll[:, None], g_lg[:, None], g_w1.reshape(n, -1), g_b1, g_w2,
                          g_b2[:, None]], axis=1)
    return out / N


# --------------------------------------------------------------------------- #
# literal log_p restatements (torch autograd stands in for tf.gradients,      #
# abstract_stein_sampler.py:55) used to validate the closed forms above       #
# --------------------------------------------------------------------------- #
def _normal_logprob(x, mu, sigma):
    return -0.5 * math.log(2 * math.pi) - sigma.log() - (x - mu) ** 2 / (2 * sigma ** 2)


def _gamma_logprob(x, a, b):
    return a * math.log(b) - math.lgamma(a) + (a - 1.0) * x.log() - b * x


def log_p_linear_torch(w, X, y):
    """examples/linear_regression/main.py:25-31 for one particle (w: F)."""
    import torch
    y_hat = X @ w
    log_l = -0.5 * ((y_hat - y) ** 2).sum()
    return log_l + _normal_logprob(w, torch.zeros_like(w), torch.ones_like(w)).sum()


def log_p_logistic_torch(th, X, y, n_train):
    """examples/logistic_regression/main.py:28-49 for one particle."""
    import torch
    w, la = th[:-1], th[-1]
    alpha = la.exp()
    logits = X @ w
    xent = torch.clamp(logits, min=0) - logits * y + torch.log1p(torch.exp(-logits.abs()))
    log_l = -xent.sum()
    prior_w = _normal_logprob(w, torch.zeros_like(w), 1.0 / alpha.sqrt()).sum()
    return log_l * (n_train / X.shape[0]) + prior_w + _gamma_logprob(alpha, 1.0, 0.01)


def log_p_bnn_torch(th, X, y, n_train, F, H):
    """examples/regression_neural_network/main.py:35-85 for one particle."""
    import torch
    o = 2
    ll, lg = th[0], th[1]
    w1 = th[o:o + F * H].reshape(F, H); o += F * H
    b1 = th[o:o + H]; o += H
    w2 = th[o:o + H]; o += H
    b2 = th[o]
    lam, gam = ll.exp(), lg.exp()
    pred = torch.relu(X @ w1 + b1) @ w2 + b2
    log_l = _normal_logprob(y, pred, 1.0 / gam.sqrt()).sum()
    sd = 1.0 / lam.sqrt()
    pri = sum(_normal_logprob(t, torch.zeros_like(t), sd).sum() for t in (w1, b1, w2, b2))
    return (log_l * n_train / X.shape[0] + _gamma_logprob(lam, 1.0, 0.01)
            + _gamma_logprob(gam, 1.0, 0.01) + pri) / n_train


def score_autograd(log_p_one, theta, *args):
    """The per-particle loop of stein/samplers/stein_sampler.py:59-68."""
    import torch
    th = torch.as_tensor(np.asarray(theta, np.float64))
    targs = [torch.as_tensor(np.asarray(a, np.float64)) if isinstance(a, np.ndarray) else a
             for a in args]
    out = np.zeros_like(np.asarray(theta, np.float64))
    for i in range(th.shape[0]):
        t = th[i].clone().requires_grad_(True)
        lp = log_p_one(t, *targs)
        (g,) = torch.autograd.grad(lp, t)
        out[i] = g.numpy()
    return out


# --------------------------------------------------------------------------- #
# CPU baseline: one full iteration, row-blocked so n x n is never resident    #
# --------------------------------------------------------------------------- #
def iteration_blocked_numpy(theta, grads, row_block=2048, rows=None):
    """The reference's per-iteration math (fp32 D/K/dK via BLAS sgemm, exact
    median, float64 K.dot(S)) restated with NumPy/BLAS and blocked over rows so
    that it can run at n = 65 536.  `rows` limits the work to the first `rows`
    rows (all n columns) -- the bounded sample bench.py times.  The median of
    the sample is taken over the sampled rows x all columns (two-pass: the
    blocks are kept in fp32, rows*n*4 bytes).  Returns (phi rows, bandwidth)."""
    T = _f32(theta)
    S = np.asarray(grads, np.float64)
    n, d = T.shape
    rows = n if rows is None else min(rows, n)
    r = np.sum(T * T, axis=1)
    blocks = []
    for i0 in range(0, rows, row_block):
        i1 = min(rows, i0 + row_block)
        blocks.append(r[i0:i1, None] + r[None, :] - 2.0 * (T[i0:i1] @ T.T))
    med = compute_median(np.concatenate([b.reshape(-1) for b in blocks]))
    bw = bandwidth(med, n)
    h2 = np.float32(bw * bw)
    phi = np.empty((rows, d), np.float64)
    for bi, i0 in enumerate(range(0, rows, row_block)):
        i1 = min(rows, i0 + row_block)
        K = np.exp(-blocks[bi] / h2 / np.float32(2.0))
        dK = (T[i0:i1] * K.sum(axis=1, dtype=np.float32)[:, None] - K @ T) / h2
        phi[i0:i1] = (K.dot(S) + dK) / n
    return phi, bw
