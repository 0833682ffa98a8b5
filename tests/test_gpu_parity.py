"""GPU parity tests: every CUDA kernel, called through the C ABI, against the
CPU oracle on identical seeded inputs.

Bars (north_star): the median / bandwidth are BIT-EXACT; phi and the post-step
particles agree to a relative tolerance of 1e-4 (normwise, and elementwise
against the largest entry); integer work (histogram counts) is exact.
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import svgd_oracle as orc

pytestmark = pytest.mark.gpu

RTOL_PHI = 1e-4      # north_star: "phi and post-step particles within ... 1e-4"


@pytest.fixture(scope="module")
def ctx():
    from stein_b200.runtime import context
    return context()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300), np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _assert_close(a, b, tol=RTOL_PHI):
    fro, mx = _rel(a, b)
    assert fro <= tol and mx <= tol, (fro, mx)


def _particles(n, d, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal((n, d)) * scale).astype(np.float32)


# --------------------------------------------------------------------------- #
# kernel (2): norms, histogram sweep, exact median, bandwidth                  #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("n,d", [(2, 1), (7, 3), (100, 10), (129, 33), (1024, 55), (300, 753)])
def test_row_norms_bit_exact(ctx, n, d):
    import torch
    X = _particles(n, d, n + d)
    Xd = ctx.to_padded(X)
    r = torch.empty(Xd.shape[0], dtype=torch.float32, device=Xd.device)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, _ptr(Xd), n, d, Xd.shape[1], _ptr(r)))
    got = r.cpu().numpy()
    np.testing.assert_array_equal(got[:n], orc.row_norms_chain(X))
    assert np.all(got[n:] == 0)


MEDIAN_CASES = [(2, 1, 1.0), (3, 2, 1.0), (7, 3, 1.0), (50, 1, 0.01), (100, 10, 0.01), (127, 5, 1.0),
                (128, 32, 1.0), (129, 33, 1.0), (257, 17, 3.0), (512, 64, 1.0), (1000, 55, 0.5),
                (1024, 256, 1.0), (1500, 9, 2.0), (2048, 30, 1.0), (2049, 20, 1.0)]


@pytest.mark.parametrize("n,d,scale", MEDIAN_CASES)
def test_median_bit_exact_small(ctx, n, d, scale):
    from stein_b200.utilities import median_sqdist
    X = _particles(n, d, 7 * n + d, scale)
    med, mid, sweeps = median_sqdist(X, return_middle=True)
    m_ref, mid_ref = orc.median_chain(X)
    assert med.tobytes() == m_ref.tobytes()
    assert (mid[0].tobytes(), mid[1].tobytes()) == (mid_ref[0].tobytes(), mid_ref[1].tobytes())
    bw = np.float32(ctx.lib.stein_bandwidth(ctypes.c_float(float(med)), n))
    assert bw.tobytes() == orc.bandwidth(m_ref, n).tobytes()


def test_median_edge_cases(ctx):
    from stein_b200.utilities import median_sqdist
    # duplicated particles: D == 0 off the diagonal too
    X = np.repeat(_particles(5, 4, 1), 40, axis=0)
    assert median_sqdist(X).tobytes() == orc.median_chain(X)[0].tobytes()
    # two tight clusters: the two middle values differ by orders of magnitude
    X = np.concatenate([np.zeros((64, 3), np.float32), np.ones((64, 3), np.float32)])
    med, mid, _ = median_sqdist(X, return_middle=True)
    assert (float(mid[0]), float(mid[1]), float(med)) == (0.0, 3.0, 1.5)
    # large dynamic range, odd n*n
    X = _particles(301, 9, 2) * np.logspace(-3, 3, 301, dtype=np.float32)[:, None]
    assert median_sqdist(X).tobytes() == orc.median_chain(X)[0].tobytes()


@pytest.mark.parametrize("n,d", [(4096, 64), (4097, 32), (6000, 256)])
def test_median_bit_exact_with_pilot_window(ctx, n, d):
    """n*n >= 2^24 takes the sampled-window route (1-2 sweeps); the oracle takes
    its radix route.  Same bits."""
    from stein_b200.utilities import median_sqdist
    X = _particles(n, d, n)
    med, mid, sweeps = median_sqdist(X, return_middle=True)
    m_ref, mid_ref = orc.median_chain(X, radix=True)
    assert med.tobytes() == m_ref.tobytes()
    assert (mid[0].tobytes(), mid[1].tobytes()) == (mid_ref[0].tobytes(), mid_ref[1].tobytes())
    assert sweeps <= 3


def _median_with(ctx, X, impl):
    from stein_b200.utilities import median_sqdist
    ctx.set_median_impl(impl)
    try:
        return median_sqdist(X, return_middle=True)
    finally:
        ctx.set_median_impl(0)


@pytest.mark.parametrize("n,d,kind", [(4096, 256, "gauss"), (5000, 128, "gauss"), (4500, 250, "small"),
                                      (6000, 256, "clusters"), (4100, 256, "dup"), (4225, 250, "gauss")])
def test_median_tensor_core_route_bit_exact(ctx, n, d, kind):
    """tcgen05 filter sweep (CTA-pair kernel and single-CTA kernel) + contract recomputation
    of the candidates == the all-FFMA route == the oracle (radix route), bit for bit.
    n = 4225 has an odd number of 128-row tiles (the last pair row is half empty)."""
    from stein_b200 import _lib
    rng = np.random.default_rng(n + d)
    X = rng.standard_normal((n, d)).astype(np.float32)
    if kind == "small":
        X *= 0.01
    elif kind == "clusters":
        X = (X * 0.05 + rng.integers(0, 3, size=(n, 1)) * 2.0).astype(np.float32)
    elif kind == "dup":
        X[n // 2:] = X[:n - n // 2]
    tc = _median_with(ctx, X, _lib.MEDIAN_TC)
    tc1 = _median_with(ctx, X, _lib.MEDIAN_TC1)
    ff = _median_with(ctx, X, _lib.MEDIAN_FFMA)
    assert tc[0].tobytes() == ff[0].tobytes() == tc1[0].tobytes()
    assert (tc1[1][0].tobytes(), tc1[1][1].tobytes()) == (ff[1][0].tobytes(), ff[1][1].tobytes())
    assert tc1[2] == 1
    assert (tc[1][0].tobytes(), tc[1][1].tobytes()) == (ff[1][0].tobytes(), ff[1][1].tobytes())
    if n <= 4500:
        m_ref, mid_ref = orc.median_chain(X, radix=True)
        assert tc[0].tobytes() == m_ref.tobytes()
        assert (tc[1][0].tobytes(), tc[1][1].tobytes()) == (mid_ref[0].tobytes(), mid_ref[1].tobytes())
    assert tc[2] == 1          # one distance sweep


def test_histogram_sweep_counts_are_exact(ctx):
    import torch
    n, d = 1000, 24
    X = _particles(n, d, 11)
    Xd = ctx.to_padded(X)
    r = torch.empty(Xd.shape[0], dtype=torch.float32, device=Xd.device)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, _ptr(Xd), n, d, Xd.shape[1], _ptr(r)))
    D = orc.sqdist_chain(X)
    keys = np.array([ctx.lib.stein_float_to_key(ctypes.c_float(float(v))) for v in np.unique(D)], np.uint64)
    key_of = dict(zip(np.unique(D).tolist(), keys.tolist()))
    allkeys = np.vectorize(key_of.get)(D).astype(np.uint64).reshape(-1)
    nt = ctx.lib.stein_num_tiles(n)
    for key_lo, shift, nbins in [(0, 18, 16384), (int(np.median(allkeys)) - 5000, 0, 16384),
                                 (int(np.median(allkeys)) - 100000, 5, 8000)]:
        counts = torch.zeros(nbins + 1, dtype=torch.int64, device=Xd.device)
        # split the tile range in two calls: counts accumulate
        for a, b in [(0, nt // 3), (nt // 3, nt)]:
            ctx.check(ctx.lib.stein_sqdist_hist(ctx.handle, _ptr(Xd), _ptr(r), n, d, Xd.shape[1], a, b,
                                                key_lo, shift, nbins, _ptr(counts)))
        got = counts.cpu().numpy().astype(np.uint64)
        ref = np.zeros(nbins + 1, np.uint64)
        ref[0] = np.sum(allkeys < key_lo)
        inw = allkeys[allkeys >= key_lo]
        b = (inw - np.uint64(key_lo)) >> np.uint64(shift)
        np.add.at(ref, 1 + b[b < nbins].astype(np.int64), 1)
        np.testing.assert_array_equal(got, ref)


def test_compute_median_values(ctx):
    from stein_b200.utilities import compute_median
    rng = np.random.default_rng(3)
    for shape in [(7,), (6, 6), (101, 101), (64, 64)]:
        V = rng.standard_normal(shape).astype(np.float32)
        assert compute_median(V).tobytes() == orc.compute_median(V).tobytes()


# --------------------------------------------------------------------------- #
# kernel (3): K, dK, phi                                                        #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("n,d", [(6, 3), (50, 1), (100, 10), (129, 33), (300, 55)])
def test_kernel_and_grad_matches_oracle(ctx, n, d):
    from stein_b200.kernels import SquaredExponentialKernel
    X = _particles(n, d, n * d)
    kern = SquaredExponentialKernel(n, None)
    K, dK = kern.kernel_and_grad(X.astype(np.float64))
    K_ref, dK_ref, bw_ref = orc.kernel_and_grad(X)
    assert kern.bandwidth.tobytes() == bw_ref.tobytes()
    assert K.shape == (n, n) and dK.shape == (n, d) and K.dtype == np.float32
    np.testing.assert_allclose(K, K_ref, atol=2e-6, rtol=1e-5)
    _assert_close(dK, dK_ref, 2e-5)


def _phi_gpu(ctx, X, S, impl, ld=None):
    """stein_row_norms -> stein_median_sqdist -> stein_phi through the C ABI."""
    import torch
    from stein_b200 import _lib
    n, d = X.shape
    Xd, Sd = ctx.to_padded(X, ld), ctx.to_padded(S, ld)
    rows, ld = Xd.shape
    r = torch.empty(rows, dtype=torch.float32, device=Xd.device)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, _ptr(Xd), n, d, ld, _ptr(r)))
    med = ctypes.c_float()
    ctx.check(ctx.lib.stein_median_sqdist(ctx.handle, _ptr(Xd), _ptr(r), n, d, ld, ctypes.byref(med), None, None))
    bw = ctx.lib.stein_bandwidth(med.value, n)
    ctx.set_phi_impl(impl)
    try:
        nb = int(ctx.lib.stein_phi_workspace_bytes(ctx.handle, n, n, ld))
        ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
        phi = torch.full_like(Xd, float("nan"))
        sumsq = torch.zeros(1, dtype=torch.float64, device=Xd.device)
        ctx.check(ctx.lib.stein_phi(ctx.handle, _ptr(Xd), _ptr(Sd), _ptr(r), n, d, ld, 0, n, bw, _ptr(ws), nb,
                                    _ptr(phi), _ptr(sumsq)))
    finally:
        ctx.set_phi_impl(_lib.PHI_AUTO)
    full = phi.cpu().numpy()
    assert np.all(full[n:] == 0) and np.all(full[:, d:] == 0), "pad must stay zero"
    return full[:n, :d].astype(np.float64), float(sumsq.item()), np.float32(bw)


PHI_CASES = [(6, 3, 1.0), (50, 1, 0.01), (100, 10, 0.01), (129, 33, 1.0), (512, 64, 1.0), (1000, 55, 0.3),
             (640, 256, 1.0), (300, 753, 0.05)]


@pytest.mark.parametrize("n,d,scale", PHI_CASES)
def test_phi_dense_matches_oracle(ctx, n, d, scale):
    from stein_b200 import _lib
    X = _particles(n, d, 3 * n + d, scale)
    S = _particles(n, d, 5 * n + d, 1.0) - X          # score of a shifted Gaussian target
    phi, sumsq, bw = _phi_gpu(ctx, X, S, _lib.PHI_DENSE_SIMT)
    ref = orc.compute_phi(X, S.astype(np.float64))
    assert bw.tobytes() == orc.kernel_and_grad(X)[2].tobytes()
    _assert_close(phi, ref)
    assert abs(sumsq - (phi ** 2).sum()) <= 1e-6 * (phi ** 2).sum()


FLASH_CASES = [(128, 256, 1.0), (130, 256, 1.0), (640, 256, 1.0), (1000, 128, 0.3), (3000, 256, 1.0),
               (2500, 250, 0.01), (4096, 100, 1.0)]


@pytest.mark.parametrize("n,d,scale", FLASH_CASES)
def test_phi_flash_tcgen05_matches_oracle(ctx, n, d, scale):
    """The tcgen05/TMEM/TMA fused kernel (single-CTA variant: three BF16 passes per GEMM, FP32 accumulate):
    1e-4 relative to the oracle, and to the FFMA dense path on the same device."""
    from stein_b200 import _lib
    X = _particles(n, d, 3 * n + d, scale)
    S = _particles(n, d, 5 * n + d, 1.0) - X
    phi, sumsq, bw = _phi_gpu(ctx, X, S, _lib.PHI_FLASH_TC)
    ref = orc.compute_phi(X, S.astype(np.float64))
    _assert_close(phi, ref)
    dense, _, _ = _phi_gpu(ctx, X, S, _lib.PHI_DENSE_SIMT)
    _assert_close(phi, dense)
    assert abs(sumsq - (phi ** 2).sum()) <= 1e-6 * (phi ** 2).sum()


@pytest.mark.parametrize("n,d,scale", [(128, 256, 1.0), (130, 256, 1.0), (640, 256, 1.0), (3000, 256, 1.0),
                                       (2500, 250, 0.01), (5000, 230, 1.0), (19000, 256, 1.0)])
def test_phi_flash_cta_pair_matches_oracle(ctx, n, d, scale):
    """cta_group::2 variant (CTA pairs, M = 256): same bar as the single-CTA kernel."""
    from stein_b200 import _lib
    X = _particles(n, d, 3 * n + d, scale)
    S = _particles(n, d, 5 * n + d, 1.0) - X
    phi, sumsq, bw = _phi_gpu(ctx, X, S, _lib.PHI_FLASH_TC2)
    one, _, _ = _phi_gpu(ctx, X, S, _lib.PHI_FLASH_TC)
    _assert_close(phi, one, 2e-5)
    if n <= 5000:
        _assert_close(phi, orc.compute_phi(X, S.astype(np.float64)))
    assert abs(sumsq - (phi ** 2).sum()) <= 1e-6 * (phi ** 2).sum()


@pytest.mark.parametrize("mode", ["gemm2", "both"])
@pytest.mark.parametrize("n,d,scale", [(128, 256, 1.0), (640, 256, 1.0), (3000, 256, 1.0), (2500, 250, 0.01),
                                       (5000, 230, 1.0), (2000, 256, 1e-4), (2000, 256, 300.0)])
def test_phi_flash_mixed_precision_matches_oracle(ctx, n, d, scale, mode):
    """CTA-pair kernel with GEMM2 (and GEMM1) as one FP16 pass + two FP8 passes instead of three
    BF16 passes: same 1e-4 bar against the oracle; with scores of very different column scales and
    particles of very different overall scales (the FP16 / FP8 operands are scaled by powers of two)."""
    from stein_b200 import _lib
    X = _particles(n, d, 3 * n + d, scale)
    S = _particles(n, d, 5 * n + d, 1.0) - X
    S *= np.logspace(-3, 4, d, dtype=np.float32)[None, :]
    phi, sumsq, bw = _phi_gpu(ctx, X, S, _lib.PHI_FLASH_TC3 if mode == "gemm2" else _lib.PHI_FLASH_TC4)
    ref = orc.compute_phi(X, S.astype(np.float64))
    fro, mx = _rel(phi, ref)
    print("mixed %s: fro %.2e max %.2e" % (mode, fro, mx))
    # column-wise: every column must be right relative to its own scale
    col = np.abs(phi - ref).max(axis=0) / np.abs(ref).max(axis=0)
    assert col.max() <= RTOL_PHI, col.max()
    assert abs(sumsq - (phi ** 2).sum()) <= 1e-6 * (phi ** 2).sum()


def _phi_float64(X, S, bw):
    """The reference formula (squared_exponential_kernel.py:22-35, abstract_stein_sampler.py:105)
    evaluated in float64 throughout, with the given bandwidth."""
    X64, S64 = X.astype(np.float64), S.astype(np.float64)
    r = (X64 ** 2).sum(1)
    D = r[:, None] + r[None, :] - 2 * X64 @ X64.T
    h2 = float(bw) ** 2
    K = np.exp(-D / h2 / 2)
    dK = (X64 * K.sum(1)[:, None] - K @ X64) / h2
    return (K @ S64 + dK) / X.shape[0]


@pytest.mark.parametrize("impl", ["flash", "pair", "pair_f8", "pair_f8x2"])
@pytest.mark.parametrize("offset", [1.0, 10.0])
def test_phi_flash_offset_cloud(ctx, impl, offset):
    """A particle cloud away from the origin (|mean|^2 >> spread^2).  The Gram form of the
    distances loses |x|^2 / D of its precision; the reference (and the oracle, and the FFMA
    dense path) form D in fp32 and are themselves 1e-4 .. 1e-2 away from the exact value of the
    formula here (tools/offset_study.py).  The flash kernels work on centred particles: the
    bandwidth is still the reference's bit for bit, and phi matches the float64 evaluation of
    the reference formula to 1e-4 at any offset."""
    from stein_b200 import _lib
    n, d = 2000, 256
    rng = np.random.default_rng(99)
    Z = rng.standard_normal((n, d))
    X = (offset + 0.1 * Z).astype(np.float32)
    S = (rng.standard_normal((n, d)) - 10.0 * Z).astype(np.float32)
    code = {"flash": _lib.PHI_FLASH_TC, "pair": _lib.PHI_FLASH_TC2, "pair_f8": _lib.PHI_FLASH_TC3,
            "pair_f8x2": _lib.PHI_FLASH_TC4}[impl]
    phi, sumsq, bw = _phi_gpu(ctx, X, S, code)
    assert bw.tobytes() == orc.kernel_and_grad(X)[2].tobytes()
    _assert_close(phi, _phi_float64(X, S, bw))


@pytest.mark.parametrize("n,d", [(128, 256), (300, 128), (1000, 256)])
def test_flash_gram_tiles_are_accurate(ctx, n, d):
    """GEMM1 of the flash kernel (3-pass BF16 split on tcgen05) against float64 X X^T."""
    import torch
    X = _particles(n, d, n + 3 * d)
    Xd, Sd = ctx.to_padded(X), ctx.to_padded(np.zeros_like(X))
    rows, ld = Xd.shape
    r = torch.empty(rows, dtype=torch.float32, device=Xd.device)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, _ptr(Xd), n, d, ld, _ptr(r)))
    nb = int(ctx.lib.stein_phi_workspace_bytes(ctx.handle, n, n, d))
    ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
    phi = torch.empty_like(Xd)
    sumsq = torch.zeros(1, dtype=torch.float64, device=Xd.device)
    G = torch.zeros((rows, rows), dtype=torch.float32, device=Xd.device)
    fn = ctx.lib.stein_debug_flash_gram
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64] * 3 + [ctypes.c_float, ctypes.c_void_p,
                                                                  ctypes.c_int64] + [ctypes.c_void_p] * 3
    ctx.check(fn(ctx.handle, _ptr(Xd), _ptr(Sd), _ptr(r), n, d, ld, 10.0, _ptr(ws), nb, _ptr(phi), _ptr(sumsq),
                 _ptr(G)))
    got = G.cpu().numpy()[:n, :n].astype(np.float64)
    ref = X.astype(np.float64) @ X.astype(np.float64).T
    scale = np.sqrt(np.outer((X.astype(np.float64) ** 2).sum(1), (X.astype(np.float64) ** 2).sum(1)))
    err = np.abs(got - ref) / scale            # relative to |x_i| |x_j|
    print("gram max rel err", err.max(), "rms", np.sqrt((err ** 2).mean()))
    assert err.max() < 2.0 ** -15


def test_compute_phi_api(ctx):
    """AbstractSteinSampler.compute_phi(theta_array, grads_array) on host arrays."""
    from stein_b200.log_p import LinearRegression
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.samplers import SteinSampler
    n, F = 100, 10
    np.random.seed(0)
    sampler = SteinSampler(n, LinearRegression(F).log_p, AdamGradientDescent(0.1))
    X, S = _particles(n, F, 1).astype(np.float64), _particles(n, F, 2).astype(np.float64)
    _assert_close(sampler.compute_phi(X, S), orc.compute_phi(X, S))


# --------------------------------------------------------------------------- #
# kernel (4): clip + optimizers, against the reference's own outputs           #
# --------------------------------------------------------------------------- #
def test_optimizers_match_reference_golden(ctx, golden_dir):
    from stein_b200.optimizers import AdagradGradientDescent, AdamGradientDescent
    g = np.load(os.path.join(golden_dir, "optimizers.npz"))
    phis = g["phis"]
    for tag, gd in (("adam", AdamGradientDescent(learning_rate=0.1, decay=0.999)),
                    ("adam_default", AdamGradientDescent()),
                    ("adagrad", AdagradGradientDescent(learning_rate=0.05, decay=0.5, alpha=0.9))):
        for t in range(phis.shape[0]):
            step = gd.update(phis[t].copy())
            ref = g[tag + "_updates"][t]
            # fp32 moments on the device vs the reference's float64: normwise 1e-4 is the
            # contract; elementwise the only slack needed is where mu cancels
            _assert_close(step, ref, 2e-5)
            np.testing.assert_allclose(step, ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())
        assert gd.n_iters == int(g[tag + "_n_iters"])
        assert abs(gd.learning_rate - float(g[tag + "_final_lr"])) <= 1e-15
    # moments are exposed like the reference's attributes
    gd = AdamGradientDescent(0.1)
    gd.update(phis[0])
    np.testing.assert_allclose(gd.mu, phis[0], rtol=1e-6)
    np.testing.assert_allclose(gd.nu, phis[0] ** 2, rtol=1e-6)


def test_clip_scale(ctx):
    """phi *= 10 / max(10, ||phi||_F) (abstract_stein_sampler.py:125) inside the step kernel."""
    import torch
    n, d = 64, 32
    rng = np.random.default_rng(0)
    phi = rng.standard_normal((n, d)) * 3.0                       # ||phi|| >> 10
    P = ctx.to_padded(phi)
    X = torch.zeros_like(P)
    hist = torch.zeros_like(P)
    sumsq = torch.tensor([float((phi.astype(np.float32).astype(np.float64) ** 2).sum())], dtype=torch.float64,
                         device=P.device)
    ctx.check(ctx.lib.stein_clip_adagrad_step(ctx.handle, _ptr(X), _ptr(P), _ptr(hist), X.numel(), _ptr(sumsq),
                                              0.05, 0.9, 0))
    gd = orc.AdagradGradientDescent(0.05, 1.0, 0.9)
    ref = gd.update(orc.clip(phi))
    np.testing.assert_allclose(X[:n, :d].cpu().numpy(), ref, rtol=2e-5)


# --------------------------------------------------------------------------- #
# kernel (1): scores and predictions of the three built-in models              #
# --------------------------------------------------------------------------- #
def _score_gpu(ctx, fn, theta, *args):
    import torch
    n, d = theta.shape
    Th = ctx.to_padded(theta)
    S = torch.zeros_like(Th)
    fn(Th, S)
    out = S.cpu().numpy()
    assert np.all(out[n:] == 0) and np.all(out[:, d:] == 0)
    return out[:n, :d]


@pytest.mark.parametrize("n,F,N", [(100, 10, 1000), (50, 1, 1000), (3, 70, 2500), (17, 300, 64)])
def test_score_linear(ctx, n, F, N):
    rng = np.random.default_rng(F)
    Xd = rng.standard_normal((N, F)).astype(np.float32)
    w = rng.standard_normal(F) * 2
    y = (Xd @ w + 0.3 * rng.standard_normal(N)).astype(np.float32)
    theta = (rng.standard_normal((n, F)) * 0.5).astype(np.float32)
    Xg, yg = ctx.dense(Xd), ctx.dense(y)
    got = _score_gpu(ctx, lambda Th, S: ctx.check(ctx.lib.stein_score_linear(
        ctx.handle, _ptr(Th), n, F, Th.shape[1], _ptr(Xg), _ptr(yg), N, _ptr(S))), theta)
    _assert_close(got, orc.score_linear(theta, Xd, y), 2e-5)


@pytest.mark.parametrize("n,F,B", [(1024, 54, 50), (100, 54, 50), (5, 3, 7), (33, 130, 300)])
def test_score_logistic(ctx, n, F, B):
    rng = np.random.default_rng(B)
    Xb = rng.standard_normal((B, F)).astype(np.float32)
    yb = (rng.random(B) > 0.5).astype(np.float32)
    theta = (rng.standard_normal((n, F + 1)) * 0.3).astype(np.float32)
    Xg, yg = ctx.dense(Xb), ctx.dense(yb)
    got = _score_gpu(ctx, lambda Th, S: ctx.check(ctx.lib.stein_score_logistic(
        ctx.handle, _ptr(Th), n, F, Th.shape[1], _ptr(Xg), _ptr(yg), B, 464809.0, 1.0, 0.01, _ptr(S))), theta)
    _assert_close(got, orc.score_logistic(theta, Xb, yb, 464809.0), 2e-5)


@pytest.mark.parametrize("n,F,H,B,N", [(512, 13, 50, 100, 506), (20, 1, 100, 20, 20), (64, 90, 50, 100, 515345),
                                       (7, 3, 5, 9, 40)])
def test_score_bnn(ctx, n, F, H, B, N):
    rng = np.random.default_rng(H)
    Xb = rng.standard_normal((B, F)).astype(np.float32)
    yb = rng.standard_normal(B).astype(np.float32)
    d = 2 + F * H + 2 * H + 1
    theta = (rng.standard_normal((n, d)) * 0.2).astype(np.float32)
    Xg, yg = ctx.dense(Xb), ctx.dense(yb)
    got = _score_gpu(ctx, lambda Th, S: ctx.check(ctx.lib.stein_score_bnn(
        ctx.handle, _ptr(Th), n, F, H, Th.shape[1], _ptr(Xg), _ptr(yg), B, float(N), 1.0, 0.01, _ptr(S))), theta)
    _assert_close(got, orc.score_bnn(theta, Xb, yb, float(N), F, H), 5e-5)


def test_predictions(ctx):
    import torch
    rng = np.random.default_rng(4)
    n, F, H, N = 33, 13, 50, 700
    Xt = rng.standard_normal((N, F)).astype(np.float32)
    Xg = ctx.dense(Xt)
    th = (rng.standard_normal((n, F + 1)) * 0.3).astype(np.float32)
    Th = ctx.to_padded(th)
    out = torch.empty((n, N), dtype=torch.float32, device=Th.device)
    ctx.check(ctx.lib.stein_predict_linear(ctx.handle, _ptr(Th), n, F, Th.shape[1], _ptr(Xg), N, _ptr(out)))
    _assert_close(out.cpu().numpy(), th[:, :F].astype(np.float64) @ Xt.T.astype(np.float64), 1e-5)
    d = 2 + F * H + 2 * H + 1
    th = (rng.standard_normal((n, d)) * 0.3).astype(np.float32)
    Th = ctx.to_padded(th)
    ctx.check(ctx.lib.stein_predict_bnn(ctx.handle, _ptr(Th), n, F, H, Th.shape[1], _ptr(Xg), N, _ptr(out)))
    _assert_close(out.cpu().numpy(), orc.bnn_predict(th, Xt, F, H), 1e-5)


# --------------------------------------------------------------------------- #
# whole iteration through the engine / sampler                                  #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("opt", ["adam", "adagrad"])
@pytest.mark.parametrize("n,d", [(100, 10), (257, 33)])
def test_engine_steps_match_oracle(ctx, opt, n, d):
    """3 consecutive update_particles() with host buffers (float64 like the
    reference's arrays): bandwidth bit-exact each step, particles within 1e-4."""
    from stein_b200.engine import SvgdEngine
    X0 = _particles(n, d, 1, 0.5).astype(np.float64)
    mean = np.random.default_rng(2).standard_normal(d)
    if opt == "adam":
        eng = SvgdEngine(n, d, "adam", learning_rate=0.1, decay=0.999)
        gd = orc.AdamGradientDescent(0.1, 0.999)
    else:
        eng = SvgdEngine(n, d, "adagrad", learning_rate=0.05, p1=0.9)
        gd = orc.AdagradGradientDescent(0.05, 1.0, 0.9)
    eng.set_particles(X0)
    X_ref = X0.copy()
    X_gpu = np.empty_like(X0)
    for it in range(3):
        # the oracle is stepped from the GPU's own (fp32) particles so that each
        # step is compared on identical inputs
        X_in = eng.get_particles(np.float64)
        S = (mean - X_in) * 3.0
        bw_ref = orc.kernel_and_grad(X_in)[2]
        X_ref, phi_ref = orc.update_particles(X_in, S, gd)
        eng.update_particles_host(np.ascontiguousarray(S), X_gpu)
        last = eng.last()
        assert np.float32(last["bandwidth"]).tobytes() == bw_ref.tobytes()
        _assert_close(eng.get_phi(), orc.compute_phi(X_in, S))
        assert abs(last["phi_norm"] - np.linalg.norm(orc.compute_phi(X_in, S))) <= 1e-4 * last["phi_norm"]
        _assert_close(X_gpu, X_ref)
        # keep the oracle optimizer's moments in step with the device (fp32) ones
        st = eng.get_state()
        assert st["n_iters"] == gd.n_iters and abs(st["learning_rate"] - gd.learning_rate) < 1e-15
    eng.close()


@pytest.mark.parametrize("n,d", [(300, 20), (2500, 256)])
def test_engine_fixed_bandwidth(ctx, n, d):
    """stein_engine_set_bandwidth: the step uses exactly the given h (no median), phi and the
    particles match the oracle evaluated with that h; 0 switches back to the median heuristic."""
    from stein_b200.engine import SvgdEngine
    X0 = _particles(n, d, 21, 0.6).astype(np.float64)
    mean = np.random.default_rng(22).standard_normal(d)
    eng = SvgdEngine(n, d, "adam", learning_rate=0.05)
    gd = orc.AdamGradientDescent(0.05)
    eng.set_particles(X0)
    h = np.float32(0.8 * np.sqrt(d))
    eng.set_bandwidth(h)
    X_gpu = np.empty_like(X0)
    for it in range(2):
        X_in = eng.get_particles(np.float64)
        S = (mean - X_in) * 1.5
        phi_ref, _ = orc.phi_rows_c(X_in, S, h, 0, n)
        eng.update_particles_host(np.ascontiguousarray(S), X_gpu)
        last = eng.last()
        assert np.float32(last["bandwidth"]).tobytes() == h.tobytes()
        assert np.isnan(last["median"]) and last["sweeps"] == 0
        _assert_close(eng.get_phi(), phi_ref)
        X_ref = X_in + gd.update(orc.clip(phi_ref))
        ok = np.abs(phi_ref) > 1e-4 * np.abs(phi_ref).max()
        assert np.abs(X_gpu - X_ref)[ok].max() <= RTOL_PHI * np.abs(X_ref).max()
    eng.set_bandwidth(None)
    X_in = eng.get_particles(np.float64)
    eng.update_particles_host(np.ascontiguousarray((mean - X_in) * 1.5), X_gpu)
    assert np.float32(eng.last()["bandwidth"]).tobytes() == orc.kernel_and_grad(X_in)[2].tobytes()
    with pytest.raises(Exception):
        eng.set_bandwidth(-1.0)
    eng.close()


@pytest.mark.parametrize("n,d,ld", [(3000, 55, 128), (4500, 200, 256), (2048, 130, 256)])
def test_engine_pads_rows_for_the_tensor_core_kernels(ctx, n, d, ld):
    """With >= 2048 particles of up to 256 coordinates the engine pads the rows to 128 / 256
    floats, so the tcgen05 kernels (phi; the median too when n*n >= 2^24) run for any such d.
    Two update_particles(): bandwidth bit-exact, particles within 1e-4 of the oracle."""
    from stein_b200.engine import SvgdEngine
    X0 = _particles(n, d, 11, 0.7).astype(np.float64)
    mean = np.random.default_rng(12).standard_normal(d)
    eng = SvgdEngine(n, d, "adam", learning_rate=0.1)
    assert eng.ld == ld
    gd = orc.AdamGradientDescent(0.1)
    eng.set_particles(X0)
    X_gpu = np.empty_like(X0)
    for it in range(2):
        X_in = eng.get_particles(np.float64)
        S = (mean - X_in) * 2.0
        bw_ref = orc.kernel_and_grad(X_in)[2]
        X_ref, _ = orc.update_particles(X_in, S, gd)
        eng.update_particles_host(np.ascontiguousarray(S), X_gpu)
        last = eng.last()
        assert np.float32(last["bandwidth"]).tobytes() == bw_ref.tobytes()
        phi_ref = orc.compute_phi(X_in, S)
        _assert_close(eng.get_phi(), phi_ref)
        # Adam's first steps are close to lr * sign(phi): where phi is within the kernels' error
        # of zero the step is ill-conditioned in ANY arithmetic, so those few entries are left out
        ok = np.abs(phi_ref) > 1e-4 * np.abs(phi_ref).max()
        assert ok.mean() > 0.98
        assert np.abs(X_gpu - X_ref)[ok].max() <= RTOL_PHI * np.abs(X_ref).max()
        if n * n >= 1 << 24:
            assert last["sweeps"] == 1          # tensor-core median route
    eng.close()


def test_sampler_linear_regression_known_answer(ctx, golden_dir):
    """examples/linear_regression/main.py on the reference's shipped data: 50
    particles, Adam lr 0.1, 500 iterations -> analytic posterior."""
    from stein_b200.log_p import LinearRegression
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.samplers import SteinSampler
    g = np.load(os.path.join(golden_dir, "linear_regression.npz"))
    X, y = g["X"], g["y"].reshape(-1, 1)
    model = LinearRegression(X.shape[1])
    np.random.seed(0)
    sampler = SteinSampler(50, model.log_p, AdamGradientDescent(learning_rate=1e-1))
    assert sampler.theta[model.w].shape == (50, 1, 1)
    # same trajectory on the oracle from the same start
    theta_ref = sampler.samples.copy()
    gd_ref = orc.AdamGradientDescent(learning_rate=1e-1)
    centre, spread = [], []
    for it in range(500):
        sampler.train_on_batch({model.X: X, model.y: y})
        if it < 5:
            S = orc.score_linear(theta_ref, X.astype(np.float32), y.astype(np.float32))
            theta_ref, _ = orc.update_particles(theta_ref, S, gd_ref)
            _assert_close(sampler.samples, theta_ref, 2e-4)
        if it >= 400:
            centre.append(sampler.samples.mean())
            spread.append(sampler.samples.std())
    est = np.array(list(sampler.theta.values()))[0].mean(axis=0).ravel()      # main.py:51
    sd = np.sqrt(g["post_cov"][0, 0])
    # Adam at lr = 0.1 keeps the cloud swinging around the posterior: on the oracle (4 seeds) the mean of ONE
    # iteration wanders over +-0.4 sd and the spread over 0.7 .. 1.5 sd -- any rounding-level change picks another
    # phase -- while their averages over the last 100 iterations sit within 0.012 sd of the analytic mean and at
    # 1.03 .. 1.11 sd.  The averages are the known answer; the last iterate only has to be in the swing.
    assert abs(np.mean(centre) - g["post_mean"][0]) < 0.1 * sd, (np.mean(centre) - g["post_mean"][0]) / sd
    assert 0.85 * sd < np.mean(spread) < 1.35 * sd, np.mean(spread) / sd
    assert abs(est[0] - g["post_mean"][0]) < 0.6 * sd
    assert 0.4 * sd < sampler.samples.std() < 2.0 * sd
    pred = sampler.function_posterior(model.y_hat, {model.X: X[:7]}, axis=0)
    np.testing.assert_allclose(pred, X[:7, 0] * sampler.samples.mean(), rtol=1e-4, atol=1e-6)


def test_sampler_follows_a_fixed_bandwidth_kernel(ctx, golden_dir):
    """The kernel object is the plugin point (abstract_kernel.py:45-62): a squared-exponential
    kernel built with `bandwidth=h` makes the sampler's engine skip the median; putting the
    default kernel back restores the heuristic."""
    from stein_b200.kernels import SquaredExponentialKernel
    from stein_b200.log_p import LinearRegression
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.samplers import SteinSampler
    g = np.load(os.path.join(golden_dir, "linear_regression.npz"))
    X, y = g["X"], g["y"].reshape(-1, 1)
    model = LinearRegression(X.shape[1])
    np.random.seed(3)
    sampler = SteinSampler(40, model.log_p, AdamGradientDescent(learning_rate=1e-2))
    sampler.kernel = SquaredExponentialKernel(40, bandwidth=0.25)
    th = sampler.samples
    S = orc.score_linear(th, X.astype(np.float32), y.astype(np.float32))
    phi_ref, _ = orc.phi_rows_c(th, S, np.float32(0.25), 0, 40)
    ref = th + orc.AdamGradientDescent(learning_rate=1e-2).update(orc.clip(phi_ref))
    sampler.train_on_batch({model.X: X, model.y: y})
    assert sampler.engine.last()["bandwidth"] == 0.25 and sampler.engine.last()["sweeps"] == 0
    _assert_close(sampler.samples, ref, 2e-4)
    assert sampler.kernel.compute_bandwidth(sampler.samples) == np.float32(0.25)
    sampler.kernel = SquaredExponentialKernel(40)
    th = sampler.samples
    sampler.train_on_batch({model.X: X, model.y: y})
    assert np.float32(sampler.engine.last()["bandwidth"]).tobytes() == orc.kernel_and_grad(th)[2].tobytes()
    with pytest.raises(ValueError):
        SquaredExponentialKernel(40, bandwidth=0.0)


def test_sampler_logistic_and_bnn_trajectories(ctx):
    """Short trajectories of the other two examples against the oracle (shared seeds)."""
    from stein_b200.log_p import LogisticRegression, RegressionNeuralNetwork
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.samplers import SteinSampler
    rng = np.random.default_rng(0)
    # logistic: covertype-shaped synthetic minibatches
    F, n, N = 54, 128, 464809
    model = LogisticRegression(F, N)
    np.random.seed(1)
    sampler = SteinSampler(n, model.log_p, AdamGradientDescent(learning_rate=1e-1))
    gd = orc.AdamGradientDescent(learning_rate=1e-1)
    for it in range(4):
        Xb = rng.standard_normal((50, F)).astype(np.float32)
        yb = (rng.random((50, 1)) > 0.5).astype(np.float32)
        th = sampler.samples
        ref, _ = orc.update_particles(th, orc.score_logistic(th, Xb, yb, N), gd)
        sampler.train_on_batch({model.X: Xb, model.y: yb})
        _assert_close(sampler.samples, ref, 2e-4)
    logits = sampler.function_posterior(model.logits, {model.X: Xb})
    _assert_close(logits, sampler.samples[:, :F] @ Xb.T.astype(np.float64), 1e-5)
    # BNN, Boston-shaped
    F, H, n, N = 13, 50, 64, 506
    model = RegressionNeuralNetwork(F, H, N)
    np.random.seed(2)
    sampler = SteinSampler(n, model.log_p, AdamGradientDescent(learning_rate=1e-1, decay=0.999))
    gd = orc.AdamGradientDescent(learning_rate=1e-1, decay=0.999)
    for it in range(4):
        Xb = rng.standard_normal((100, F)).astype(np.float32)
        yb = rng.standard_normal((100, 1)).astype(np.float32)
        th = sampler.samples
        ref, _ = orc.update_particles(th, orc.score_bnn(th, Xb, yb, N, F, H), gd)
        sampler.train_on_batch({model.X: Xb, model.y: yb})
        _assert_close(sampler.samples, ref, 2e-4)
    pred = sampler.function_posterior(model.pred, {model.X: Xb})
    _assert_close(pred, orc.bnn_predict(sampler.samples, Xb, F, H), 1e-4)


def test_torch_log_posterior_carrier(ctx):
    """User log_p through torch autograd (vmap(grad)) reproduces the fused logistic scores."""
    import torch
    from stein_b200.log_p import LogisticRegression, TorchLogPosterior
    from stein_b200.optimizers import AdagradGradientDescent
    from stein_b200.samplers import SteinSampler
    F, n, N = 6, 40, 1000.0

    def log_p(params, feed):
        w, la = params["w"].reshape(-1), params["log_alpha"]
        z = feed["X"] @ w
        yv = feed["y"].reshape(-1)
        ll = -(torch.clamp(z, min=0) - z * yv + torch.log1p(torch.exp(-z.abs()))).sum()
        alpha = la.exp()
        prior = (0.5 * la - 0.5 * alpha * w ** 2).sum() + (-0.01 * alpha)
        return ll * (N / feed["X"].shape[0]) + prior

    tm = TorchLogPosterior({"w": [F, 1], "log_alpha": []}, log_p)
    pX, py = tm.placeholder("X", [None, F]), tm.placeholder("y", [None, 1])
    bm = LogisticRegression(F, N)
    rng = np.random.default_rng(0)
    Xb = rng.standard_normal((25, F)).astype(np.float32)
    yb = (rng.random((25, 1)) > 0.5).astype(np.float32)
    theta0 = rng.standard_normal((n, F + 1)) * 0.3
    th_t = {tm.vars["w"]: theta0[:, :F].reshape(n, F, 1), tm.vars["log_alpha"]: theta0[:, F]}
    th_b = {bm.w: theta0[:, :F].reshape(n, F, 1), bm.log_alpha: theta0[:, F]}
    s1 = SteinSampler(n, tm.log_p, AdagradGradientDescent(0.05), theta=th_t)
    s2 = SteinSampler(n, bm.log_p, AdagradGradientDescent(0.05), theta=th_b)
    for _ in range(3):
        s1.train_on_batch({pX: Xb, py: yb})
        s2.train_on_batch({bm.X: Xb, bm.y: yb})
    _assert_close(s1.samples, s2.samples, 1e-4)


# --------------------------------------------------------------------------- #
# full-size checks (BASELINE.json config D: n = 65 536, d = 256)               #
# --------------------------------------------------------------------------- #
def test_full_size_properties(ctx):
    """At n = 65 536, d = 256 the oracle cannot materialise anything n x n, so:
    (a) the histogram of one full sweep sums to exactly n*n; (b) the median is
    bracketed by the oracle's medians of row-sampled sub-problems and its two
    middle values are adjacent in the sweep's counts; (c) sampled phi rows match
    the C oracle's rows (full columns); (d) sum_i dK_i = 0: with S = 0 the
    column sums of n*phi vanish relative to their scale."""
    import torch
    from stein_b200.engine import SvgdEngine
    n, d = 65536, 256
    rng = np.random.default_rng(1)
    X = rng.standard_normal((n, d)).astype(np.float32)
    S = -X
    eng = SvgdEngine(n, d, "adam", learning_rate=0.1)
    eng.set_particles(X)
    eng.set_scores(S)
    Xd = eng.particles_dev
    r = torch.empty(n, dtype=torch.float32, device=Xd.device)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, _ptr(Xd), n, d, d, _ptr(r)))
    # (a)
    counts = torch.zeros(16385, dtype=torch.int64, device=Xd.device)
    ctx.check(ctx.lib.stein_sqdist_hist(ctx.handle, _ptr(Xd), _ptr(r), n, d, d, 0, ctx.lib.stein_num_tiles(n),
                                        0, 18, 16384, _ptr(counts)))
    assert int(counts.sum().item()) == n * n
    # (b)
    med, mid, sweeps = ctypes.c_float(), (ctypes.c_float * 2)(), ctypes.c_int32()
    ctx.check(ctx.lib.stein_median_sqdist(ctx.handle, _ptr(Xd), _ptr(r), n, d, d, ctypes.byref(med), mid,
                                          ctypes.byref(sweeps)))
    assert mid[0] <= med.value <= mid[1] and sweeps.value <= 3
    # the tensor-core route (AUTO at this shape) and the all-FFMA route agree bit for bit
    from stein_b200 import _lib
    ctx.set_median_impl(_lib.MEDIAN_FFMA)
    med_f, mid_f = ctypes.c_float(), (ctypes.c_float * 2)()
    ctx.check(ctx.lib.stein_median_sqdist(ctx.handle, _ptr(Xd), _ptr(r), n, d, d, ctypes.byref(med_f), mid_f, None))
    ctx.set_median_impl(0)
    assert (med_f.value, mid_f[0], mid_f[1]) == (med.value, mid[0], mid[1])
    sub = [float(orc.median_chain(X[rng.choice(n, 2048, replace=False)])[0]) for _ in range(3)]
    assert min(sub) * 0.995 < med.value < max(sub) * 1.005
    k0, k1 = ctx.lib.stein_float_to_key(mid[0]), ctx.lib.stein_float_to_key(mid[1])
    c2 = torch.zeros(3, dtype=torch.int64, device=Xd.device)
    ctx.check(ctx.lib.stein_sqdist_hist(ctx.handle, _ptr(Xd), _ptr(r), n, d, d, 0, ctx.lib.stein_num_tiles(n),
                                        k0, 0, 1, _ptr(c2)))
    below, at = int(c2[0].item()), int(c2[1].item())
    assert below <= n * n // 2 - 1 < below + at           # rank n^2/2-1 is the value mid[0]
    if k1 != k0:
        assert below + at == n * n // 2                   # and rank n^2/2 is the next value present
    # (c) + (d)
    eng.step()
    info = eng.last()
    assert np.float32(info["median"]).tobytes() == np.float32(med.value).tobytes()
    phi = eng.get_phi(np.float64)
    rows = [0, 1, 777, 32768, 65535]
    bw = np.float32(info["bandwidth"])
    for i in rows:
        ref, _ = orc.phi_rows_c(X, S, bw, i, i + 1)
        _assert_close(phi[i], ref[0])
    eng.set_particles(X)
    eng.set_scores(np.zeros_like(X))
    eng.step()
    phi0 = eng.get_phi(np.float64)
    assert np.abs(phi0.sum(axis=0)).max() <= 1e-4 * np.abs(phi0).sum(axis=0).max()
    eng.close()


# --------------------------------------------------------------------------- #
# medium-size trajectories on the tensor-core paths                            #
# --------------------------------------------------------------------------- #
def _run_engine(n, d, steps, phi_impl, seed=5, lr=0.05):
    from stein_b200.engine import SvgdEngine
    from stein_b200.runtime import context
    ctx = context()
    ctx.set_phi_impl(phi_impl)
    try:
        rng = np.random.default_rng(seed)
        X = (1.5 + 2.0 * rng.standard_normal((n, d))).astype(np.float32)     # off-centre, too wide
        eng = SvgdEngine(n, d, "adam", learning_rate=lr)
        eng.set_particles(X)
        sweeps, bws = [], []
        for _ in range(steps):
            eng.set_scores(-eng.get_particles(np.float32))                    # standard normal target
            eng.step()
            info = eng.last()
            sweeps.append(info["sweeps"])
            bws.append(info["bandwidth"])
        out = eng.get_particles(np.float64)
        eng.close()
        return out, sweeps, bws
    finally:
        ctx.set_phi_impl(0)


def test_trajectory_is_deterministic_and_stable(ctx):
    """40 SVGD iterations at n = 8192, d = 256 on the default (mixed-precision, CTA-pair) path:
    bit-identical when repeated; one tensor-core sweep per median every time; finite and moving
    towards the target; and within 1e-3 of the same trajectory on the BF16x3 pair kernel (the two
    arithmetics differ by ~1e-5 per step)."""
    from stein_b200 import _lib
    n, d, steps = 8192, 256, 40
    a, sweeps, bws = _run_engine(n, d, steps, _lib.PHI_AUTO)
    b, _, bws_b = _run_engine(n, d, steps, _lib.PHI_AUTO)
    assert np.array_equal(a, b) and bws == bws_b
    assert all(s == 1 for s in sweeps), sweeps
    assert np.isfinite(a).all() and all(np.isfinite(bws))
    assert abs(a.mean()) < 1.5 and a.std() < 2.0 + 1e-6            # started at mean 1.5, sd 2.0
    c, _, bws_c = _run_engine(n, d, steps, _lib.PHI_FLASH_TC2)
    assert np.abs(a - c).max() <= 1e-3 * np.abs(c).max()
    assert np.allclose(bws, bws_c, rtol=1e-4)


def test_median_window_hint_miss_falls_back(ctx):
    """Within an engine the pilot histogram and the sweep window of iteration t are taken around
    the window of iteration t-1 without a host round trip.  When the particles jump (here: the same
    cloud blown up 10x between two steps) the pilot ranks fall outside that histogram; the
    iteration must notice, take the generic route and still return the exact median."""
    from stein_b200.engine import SvgdEngine
    n, d = 4096, 256
    X = _particles(n, d, 21)
    eng = SvgdEngine(n, d, "adam", learning_rate=1e-3)
    sweeps = []
    for scale in (1.0, 1.0, 10.0, 10.0, 0.05):
        Xs = (X * scale).astype(np.float32)
        eng.set_particles(Xs)
        eng.set_scores(-Xs)
        Xin = eng.get_particles(np.float32)
        eng.step()
        info = eng.last()
        m_ref, _ = orc.median_chain(Xin, radix=True)
        assert np.float32(info["median"]).tobytes() == m_ref.tobytes(), scale
        sweeps.append(info["sweeps"])
    eng.close()
    # first step: no hint; second: hint hits; after each jump the speculative sweep is wasted once
    assert sweeps[0] == 1 and sweeps[1] == 1 and sweeps[2] == 2 and sweeps[3] == 1 and sweeps[4] == 2, sweeps


def test_sampler_takes_a_tf1_graph_like_the_reference(ctx, golden_dir):
    """SteinSampler(n_particles, log_p, gd) with `log_p` a graph tensor recorded by the
    TensorFlow-1 stand-in of compat/ (what the reference's scripts pass): same trajectory as the
    built-in model class from the same start, `theta` keyed by the graph's variables,
    function_posterior on any tensor of the graph."""
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat")
    if compat not in sys.path:
        sys.path.insert(0, compat)
    import tensorflow as tf
    from tensorflow.contrib.distributions import Normal
    from stein_b200.log_p import LinearRegression
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.samplers import SteinSampler
    g = np.load(os.path.join(golden_dir, "linear_regression.npz"))
    X, y = g["X"], g["y"].reshape(-1, 1)
    tf.reset_default_graph()
    with tf.variable_scope("model"):
        model_X = tf.placeholder(tf.float32, shape=[None, X.shape[1]])
        model_y = tf.placeholder(tf.float32, shape=[None, 1])
        model_w = tf.Variable(tf.zeros([X.shape[1], 1]))
        y_hat = tf.matmul(model_X, model_w)
        log_p = (-0.5 * tf.reduce_sum(tf.square(y_hat - model_y)) +
                 tf.reduce_sum(Normal(tf.zeros([X.shape[1], 1]), 1.).log_prob(model_w)))
    np.random.seed(4)
    sampler = SteinSampler(50, log_p, AdamGradientDescent(learning_rate=1e-1))
    assert sampler.model_vars == [model_w] and sampler.theta[model_w].shape == (50, 1, 1)
    builtin = LinearRegression(X.shape[1])
    np.random.seed(4)
    twin = SteinSampler(50, builtin.log_p, AdamGradientDescent(learning_rate=1e-1))
    np.testing.assert_array_equal(sampler.samples, twin.samples)
    for it in range(5):
        sampler.train_on_batch({model_X: X, model_y: y})
        twin.train_on_batch({builtin.X: X, builtin.y: y})
        _assert_close(sampler.samples, twin.samples, 2e-4)
    pred = sampler.function_posterior(y_hat, {model_X: X[:7]}, axis=0)
    np.testing.assert_allclose(pred, twin.function_posterior(builtin.y_hat, {builtin.X: X[:7]}, axis=0),
                               rtol=1e-3, atol=1e-6)
    full = sampler.function_posterior(y_hat, {model_X: X[:7]})
    assert full.shape == (50, 7)


# --------------------------------------------------------------------------- #
# against tests/golden/reference_run.npz: the reference's own library code,     #
# executed on the TF1 stand-in (tests/golden/make_golden_reference_run.py)      #
# --------------------------------------------------------------------------- #
def test_kernel_and_grad_matches_the_reference_run(ctx, golden_dir):
    """stein/kernels/squared_exponential_kernel.py:25-35 run by the reference's own class."""
    from stein_b200.kernels import SquaredExponentialKernel
    ref = np.load(os.path.join(golden_dir, "reference_run.npz"))
    for i, (n, d) in enumerate(ref["kernel_shapes"]):
        theta = ref["kernel%d_theta" % i]
        kern = SquaredExponentialKernel(int(n), None)
        K, dK = kern.kernel_and_grad(theta)
        assert abs(float(kern.bandwidth) - float(ref["kernel%d_bandwidth" % i])) <= 2e-6 * float(kern.bandwidth)
        np.testing.assert_allclose(K, ref["kernel%d_K" % i], atol=4e-6, rtol=2e-5)
        _assert_close(dK, ref["kernel%d_dK" % i], 3e-5)


@pytest.mark.parametrize("tag", ["linear", "logistic", "bnn_adam", "bnn_adagrad"])
def test_first_iteration_matches_the_reference_run(ctx, golden_dir, tag):
    """One SteinSampler.train_on_batch from the particles the reference run started from: scores
    (the reference's per-particle tf.gradients loop), phi (its compute_phi) and the particles
    after its clip + optimizer step.  Entries whose phi is within the kernels' error of zero are
    left out of the particle comparison (the first Adam / Adagrad step is lr * c * sign(phi))."""
    from stein_b200.log_p import LinearRegression, LogisticRegression, RegressionNeuralNetwork
    from stein_b200.optimizers import AdagradGradientDescent, AdamGradientDescent
    from stein_b200.samplers import SteinSampler
    from stein_b200.utilities import convert_array_to_dictionary
    ref = np.load(os.path.join(golden_dir, "reference_run.npz"))
    traj = ref[tag + "_traj"]
    n, d = traj[0].shape
    if tag == "linear":
        g = np.load(os.path.join(golden_dir, "linear_regression.npz"))
        model = LinearRegression(g["X"].shape[1])
        feed = {model.X: g["X"], model.y: g["y"].reshape(-1, 1)}
        gd = AdamGradientDescent(learning_rate=1e-1)
    elif tag == "logistic":
        X, y, idx = ref["logistic_X"], ref["logistic_y"], ref["logistic_batches"][0]
        model = LogisticRegression(X.shape[1], X.shape[0])
        feed = {model.X: X[idx], model.y: y[idx]}
        gd = AdamGradientDescent(learning_rate=1e-1)
    else:
        X, y, idx = ref["bnn_X"], ref["bnn_y"], ref["bnn_batches"][0]
        model = RegressionNeuralNetwork(X.shape[1], 5, X.shape[0])
        feed = {model.X: X[idx], model.y: y[idx]}
        gd = (AdamGradientDescent(learning_rate=1e-1, decay=0.999) if tag == "bnn_adam"
              else AdagradGradientDescent(learning_rate=5e-2, decay=0.5, alpha=0.9))
    assert model.n_params == d
    theta0 = convert_array_to_dictionary(traj[0], model.column_slices())
    sampler = SteinSampler(int(n), model.log_p, gd, theta=theta0)
    _assert_close(sampler.samples, traj[0], 1e-6)
    sampler.train_on_batch(feed)
    eng = sampler.engine
    _assert_close(eng.scores_dev[:n, :d].cpu().numpy(), ref[tag + "_scores0"], 1e-4)
    phi_ref = ref[tag + "_phi0"]
    _assert_close(eng.get_phi(), phi_ref)
    ok = np.abs(phi_ref) > 1e-4 * np.abs(phi_ref).max()
    assert ok.mean() > 0.9
    assert np.abs(sampler.samples - traj[1])[ok].max() <= RTOL_PHI * np.abs(traj[1]).max()


# --------------------------------------------------------------------------- #
# round 2: conditioning guard, shard-shaped calls, mixture score               #
# --------------------------------------------------------------------------- #
def _cloud(kind, n, d, rng):
    """The badly conditioned clouds of VERDICT round 1 ("what's weak" #1)."""
    Z = rng.standard_normal((n, d))
    # random membership: the clusters differ in size, so the median is an intra-cluster distance
    # (an exact 50/50 split puts it half way between the largest intra- and the smallest inter-cluster one)
    half = (rng.random(n) < 0.5)[:, None]
    if kind == "gauss":
        return Z
    if kind == "pm3":
        return np.where(half, 3.0, -3.0) + 0.3 * Z
    if kind == "pm10":
        return np.where(half, 10.0, -10.0) + 0.1 * Z
    if kind == "split6040":
        return np.where((rng.random(n) < 0.6)[:, None], 1.0, -1.0) + 0.05 * Z
    if kind == "gmm":
        means = np.zeros((4, d))
        means[0, 0], means[1, 0], means[2, 1], means[3, 1] = 2, -2, 2, -2
        return means[rng.integers(0, 4, n)] + Z
    raise ValueError(kind)


@pytest.mark.parametrize("kind,n", [("gauss", 2048), ("pm3", 2048), ("split6040", 2560), ("gmm", 3000), ("pm10", 2048)])
def test_phi_auto_route_on_multimodal_clouds(ctx, kind, n):
    """The AUTO route (conditioning guard: fast FP16+FP8 kernel, precise FP16x3 kernel, or the FP32
    FFMA path) on multi-modal clouds, d = 256, against the float64 evaluation of the reference formula
    (squared_exponential_kernel.py:22-35, abstract_stein_sampler.py:105) and against the fp32 oracle.

    Bar: 1e-4 relative wherever the reference's own fp32 arithmetic reaches it.  On the worst clouds
    (kappa = max|x - mean|^2 / h^2 of 10^3 .. 10^4) the fp32 Gram form of abstract_kernel.py:33-35
    itself is further than 1e-4 from float64 (the oracle's own error is measured here); there the
    library must stay within twice the oracle's error -- the reference's arithmetic is the contract --
    and the guard must have left the tensor-core routes."""
    from stein_b200 import _lib
    d = 256
    rng = np.random.default_rng(5)
    X = _cloud(kind, n, d, rng).astype(np.float32)
    S = (rng.standard_normal((n, d)) - X).astype(np.float32)
    phi, sumsq, bw = _phi_gpu(ctx, X, S, _lib.PHI_AUTO)
    route = ctx.phi_route()
    assert bw.tobytes() == orc.kernel_and_grad(X)[2].tobytes()
    ref64 = _phi_float64(X, S, bw)
    oracle = orc.compute_phi(X, S.astype(np.float64))
    err = np.abs(phi - ref64).max() / np.abs(ref64).max()
    err_oracle = np.abs(oracle - ref64).max() / np.abs(ref64).max()
    print("%s: kappa %.3g route %s predicted(fast) %.2e | err vs float64 %.2e (oracle's own %.2e)"
          % (kind, route["kappa"], route["route"], route["predicted_fast_error"], err, err_oracle))
    assert route["route"] == ("fast" if kind in ("gauss", "gmm") else "ffma"), route
    assert err <= max(RTOL_PHI, 2.0 * err_oracle), (err, err_oracle, route)
    if kind in ("gauss", "gmm"):
        assert err <= RTOL_PHI
        _assert_close(phi, oracle)


@pytest.mark.parametrize("tol,want", [(0.0, "ffma"), (2.0e-5, "precise"), (1.0, "fast")])
def test_phi_guard_takes_the_fastest_route_within_tolerance(ctx, tol, want):
    """The three routes of the guard on one Gaussian cloud (kappa ~ 5): forcing the tolerance walks
    through them; each stays within 1e-4 of the oracle."""
    from stein_b200 import _lib
    n, d = 2304, 256
    X = _particles(n, d, 31)
    S = _particles(n, d, 32) - X
    ctx.set_phi_guard_tol(tol)
    try:
        phi, _, _ = _phi_gpu(ctx, X, S, _lib.PHI_AUTO)
        route = ctx.phi_route()
    finally:
        ctx.set_phi_guard_tol(5e-5)
    assert route["route"] == want, route
    _assert_close(phi, orc.compute_phi(X, S.astype(np.float64)))


@pytest.mark.parametrize("n_local,row_begin", [(8192, 0), (8192, 16384), (8192, 57344), (128, 32768), (2944, 62592)])
def test_phi_shard_shaped_calls_match_oracle_rows(ctx, n_local, row_begin):
    """What a particle-sharded rank runs (emulated on one GPU): stein_phi for
    the row block [row_begin, row_begin + n_local) of n = 65 536 particles against all columns --
    every row tile is a "leftover" of the schedule (fewer row tiles than clusters) -- compared with
    the C oracle on 256 random rows of the block, in the default (guarded FP16 + FP8) mode."""
    import torch
    from stein_b200 import _lib
    n, d = 65536, 256
    X = _particles(n, d, 11)
    S = (-X + 0.1 * _particles(n, d, 12)).astype(np.float32)
    Xd, Sd = ctx.to_padded(X), ctx.to_padded(S)
    rows, ld = Xd.shape
    r = torch.empty(rows, dtype=torch.float32, device=Xd.device)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, _ptr(Xd), n, d, ld, _ptr(r)))
    bw = np.float32(3.3971)          # any fixed bandwidth: phi is compared at the same h on both sides
    n_loc = min(n_local, n - row_begin)
    nb = int(ctx.lib.stein_phi_workspace_bytes(ctx.handle, n_loc, n, d))
    ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
    prow = ctx.rows_padded(n_loc)
    phi = torch.full((prow, ld), float("nan"), dtype=torch.float32, device=Xd.device)
    sumsq = torch.zeros(1, dtype=torch.float64, device=Xd.device)
    ctx.check(ctx.lib.stein_phi(ctx.handle, _ptr(Xd), _ptr(Sd), _ptr(r), n, d, ld, row_begin, n_loc, float(bw),
                                _ptr(ws), nb, _ptr(phi), _ptr(sumsq)))
    got = phi.cpu().numpy()
    assert np.all(got[n_loc:] == 0), "pad rows must stay zero"
    rng = np.random.default_rng(row_begin + n_local)
    pick = np.sort(rng.choice(n_loc, size=min(256, n_loc), replace=False))
    scale = 0.0
    worst = 0.0
    for i in pick:
        ref, _ = orc.phi_rows_c(X, S, bw, row_begin + int(i), row_begin + int(i) + 1)
        scale = max(scale, np.abs(ref).max())
        worst = max(worst, np.abs(got[i, :d] - ref[0]).max())
    assert worst <= RTOL_PHI * scale, (worst, scale)
    assert abs(float(sumsq.item()) - (got.astype(np.float64) ** 2).sum()) <= 1e-6 * float(sumsq.item())


def test_tile_range_histograms_sum_to_the_unsharded_one(ctx):
    """The median's distributed sweep: contiguous tile ranges (one per rank, stein_b200.distributed
    .shard_tiles) accumulate to exactly the counts of the single-range sweep (integer work: exact)."""
    import torch
    from stein_b200.distributed import shard_tiles
    n, d = 3000, 64
    X = _particles(n, d, 21)
    Xd = ctx.to_padded(X)
    rows, ld = Xd.shape
    r = torch.empty(rows, dtype=torch.float32, device=Xd.device)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, _ptr(Xd), n, d, ld, _ptr(r)))
    nt = int(ctx.lib.stein_num_tiles(n))
    klo, shift, nbins = 0, 18, 16384

    def sweep(t0, t1, into):
        ctx.check(ctx.lib.stein_sqdist_hist(ctx.handle, _ptr(Xd), _ptr(r), n, d, ld, t0, t1, klo, shift, nbins, _ptr(into)))

    whole = torch.zeros(nbins + 1, dtype=torch.int64, device=Xd.device)
    sweep(0, nt, whole)
    for world in (2, 3, 8):
        acc = torch.zeros_like(whole)
        for rank in range(world):
            t0, t1 = shard_tiles(n, world, rank)
            part = torch.zeros_like(whole)
            sweep(t0, t1, part)
            acc += part
        assert torch.equal(acc, whole), world
    assert int(whole.sum().item()) == n * n


@pytest.mark.parametrize("n,d,ncomp", [(1000, 256, 4), (300, 1024, 4), (129, 33, 3), (64, 8, 1)])
def test_score_gaussian_mixture(ctx, n, d, ncomp):
    """stein_score_gaussian_mixture (the score inside bench.py's timed region, configs D / E)
    against the float64 closed form of the oracle."""
    import torch
    rng = np.random.default_rng(n + d)
    X = (1.5 * rng.standard_normal((n, d))).astype(np.float32)
    means = None
    if ncomp > 1:
        means = np.zeros((ncomp, d), np.float32)
        for k in range(ncomp):
            means[k, (k // 2) % d] = 2.0 if k % 2 == 0 else -2.0
    Xd = ctx.to_padded(X)
    Sd = torch.full_like(Xd, float("nan"))
    md = ctx.dense(means) if means is not None else None
    sigma2 = 1.3
    ctx.check(ctx.lib.stein_score_gaussian_mixture(ctx.handle, _ptr(Xd), n, d, Xd.shape[1],
                                                   _ptr(md) if md is not None else None, ncomp, sigma2, _ptr(Sd)))
    got = Sd.cpu().numpy()[:n, :d]
    ref = orc.score_gmm(X, means, sigma2)
    assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()


def test_custom_gradient_descent_uses_the_engine_phi(ctx):
    """A user-defined AbstractGradientDescent (abstract_gradient_descent.py:32-52) on >= 2 048 particles of
    few coordinates: the engine pads the rows to 128 floats and stein_engine_phi_only must run on the
    engine's own workspace / leading dimension (the old path sized the workspace for
    stein_ld(d) and failed with "phi workspace too small")."""
    from stein_b200.log_p import LinearRegression
    from stein_b200.optimizers.abstract_gradient_descent import AbstractGradientDescent
    from stein_b200.samplers import SteinSampler

    class PlainAscent(AbstractGradientDescent):
        def update(self, phi):
            self.n_iters += 1
            return self.learning_rate * phi

    n, F, N = 2304, 10, 64
    rng = np.random.default_rng(4)
    Xd = rng.standard_normal((N, F)).astype(np.float32)
    yd = rng.standard_normal((N, 1)).astype(np.float32)
    model = LinearRegression(F)
    np.random.seed(1)
    sampler = SteinSampler(n, model.log_p, PlainAscent(0.05, 1.0))
    theta0 = sampler.samples
    sampler.train_on_batch({model.X: Xd, model.y: yd})
    got = sampler.samples
    S = orc.score_linear(theta0, Xd, yd)
    phi = orc.compute_phi(theta0.astype(np.float32), S)
    phi *= 10. / max(10., np.linalg.norm(phi))
    ref = theta0 + 0.05 * phi
    assert np.abs((got - theta0) - (ref - theta0)).max() <= RTOL_PHI * np.abs(ref - theta0).max()


def test_learning_rate_schedule_reaches_the_engine(ctx):
    """`gd.learning_rate = x` between iterations (the reference reads it at every update(),
    adam_gradient_descent.py:55) must change the next engine step."""
    from stein_b200.log_p import LinearRegression
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.samplers import SteinSampler
    n, F, N = 64, 5, 32
    rng = np.random.default_rng(8)
    Xd = rng.standard_normal((N, F)).astype(np.float32)
    yd = rng.standard_normal((N, 1)).astype(np.float32)
    model = LinearRegression(F)
    np.random.seed(2)
    gd = AdamGradientDescent(learning_rate=0.1, decay=0.9)
    sampler = SteinSampler(n, model.log_p, gd)
    ogd = orc.AdamGradientDescent(learning_rate=0.1, decay=0.9)
    theta = sampler.samples
    for it in range(4):
        if it == 2:
            gd.learning_rate = 0.5
            ogd.learning_rate = 0.5
        theta, _ = orc.update_particles(theta, orc.score_linear(theta, Xd, yd), ogd)
        sampler.train_on_batch({model.X: Xd, model.y: yd})
        assert abs(gd.learning_rate - ogd.learning_rate) <= 1e-15
    assert np.abs(sampler.samples - theta).max() <= RTOL_PHI * np.abs(theta).max()


# --------------------------------------------------------------------------- #
# round 2: more than 256 coordinates on the tensor cores (config E shapes)     #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("mode", ["fast", "precise", "auto"])
@pytest.mark.parametrize("n,d", [(1000, 512), (4096, 512), (2500, 700), (4096, 1024), (5000, 1000)])
def test_phi_panel_kernels_match_oracle(ctx, n, d, mode):
    """Leading dimension 512 / 768 / 1024: the K-streaming panel kernels (panel_gemm.cuh, phi_panel.cuh)
    against the oracle (abstract_stein_sampler.py:100-105), 1e-4, pads zero, sum(phi^2) consistent."""
    from stein_b200 import _lib
    X = _particles(n, d, 3 * n + d)
    S = _particles(n, d, 5 * n + d) - X
    code = {"fast": _lib.PHI_FLASH_TC4, "precise": _lib.PHI_FLASH_TC5, "auto": _lib.PHI_AUTO}[mode]
    phi, sumsq, bw = _phi_gpu(ctx, X, S, code, ld=-(-d // 256) * 256)      # what the engine pads such rows to
    ref = orc.compute_phi(X, S.astype(np.float64))
    assert bw.tobytes() == orc.kernel_and_grad(X)[2].tobytes()
    fro, mx = _rel(phi, ref)
    print("panel %s n=%d d=%d: fro %.2e max %.2e" % (mode, n, d, fro, mx))
    _assert_close(phi, ref)
    assert abs(sumsq - (phi ** 2).sum()) <= 1e-6 * (phi ** 2).sum()
    if mode == "auto":
        assert ctx.phi_route()["route"] == "fast"


def test_phi_panel_config_e_shape_and_shard(ctx):
    """n = 8 192, d = 1 024 on the Gaussian-mixture cloud of BASELINE.json config E, as one block and as the
    shard-shaped call of rank 1 of 4 (row_begin = 2 048), against the oracle rows."""
    import torch
    from stein_b200 import _lib
    n, d = 8192, 1024
    rng = np.random.default_rng(77)
    X = _cloud("gmm", n, d, rng).astype(np.float32)
    means = np.zeros((4, d), np.float32)
    means[0, 0], means[1, 0], means[2, 1], means[3, 1] = 2, -2, 2, -2
    S = orc.score_gmm(X, means).astype(np.float32)
    phi, sumsq, bw = _phi_gpu(ctx, X, S, _lib.PHI_AUTO)
    ref = orc.compute_phi(X, S.astype(np.float64))
    _assert_close(phi, ref)
    # shard-shaped: rows [2048, 4096) against all columns
    Xd, Sd = ctx.to_padded(X), ctx.to_padded(S)
    rows, ld = Xd.shape
    r = torch.empty(rows, dtype=torch.float32, device=Xd.device)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, _ptr(Xd), n, d, ld, _ptr(r)))
    nb = int(ctx.lib.stein_phi_workspace_bytes(ctx.handle, 2048, n, d))
    ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
    out = torch.full((2048, ld), float("nan"), dtype=torch.float32, device=Xd.device)
    ss = torch.zeros(1, dtype=torch.float64, device=Xd.device)
    ctx.check(ctx.lib.stein_phi(ctx.handle, _ptr(Xd), _ptr(Sd), _ptr(r), n, d, ld, 2048, 2048, float(bw), _ptr(ws), nb,
                                _ptr(out), _ptr(ss)))
    _assert_close(out.cpu().numpy()[:, :d], ref[2048:4096])


@pytest.mark.parametrize("n,d,kind", [(4096, 512, "gauss"), (4096, 1024, "gauss"), (4225, 700, "gauss"),
                                      (5000, 1024, "gmm"), (4500, 1000, "small")])
def test_median_tensor_core_route_beyond_256_coordinates(ctx, n, d, kind):
    """Leading dimension 512 / 768 / 1024: the K-streaming tcgen05 sweep (Sweep3Policy on the main loop of
    panel_gemm.cuh) + contract recomputation of the candidates == the all-FFMA route == the C oracle,
    bit for bit (compute_median.py:4-16 over abstract_kernel.py:33-35); one distance sweep."""
    from stein_b200 import _lib
    rng = np.random.default_rng(n + d)
    if kind == "gmm":
        X = _cloud("gmm", n, d, rng).astype(np.float32)
    else:
        X = rng.standard_normal((n, d)).astype(np.float32) * (0.01 if kind == "small" else 1.0)
    tc = _median_with(ctx, X, _lib.MEDIAN_TC)
    ff = _median_with(ctx, X, _lib.MEDIAN_FFMA)
    assert tc[0].tobytes() == ff[0].tobytes()
    assert (tc[1][0].tobytes(), tc[1][1].tobytes()) == (ff[1][0].tobytes(), ff[1][1].tobytes())
    assert tc[2] == 1
    if n <= 4500:
        m_ref, mid_ref = orc.median_chain(X, radix=True)
        assert tc[0].tobytes() == m_ref.tobytes()
        assert (tc[1][0].tobytes(), tc[1][1].tobytes()) == (mid_ref[0].tobytes(), mid_ref[1].tobytes())


# --------------------------------------------------------------------------- #
# round 2: other kernel operators through the plugin point #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("n,d,beta", [(50, 3, -0.5), (300, 55, -0.5), (129, 33, -1.0)])
def test_imq_kernel_and_grad_matches_oracle(ctx, n, d, beta):
    from stein_b200.kernels import InverseMultiquadricKernel
    X = _particles(n, d, n + d, 0.7)
    kern = InverseMultiquadricKernel(n, beta=beta)
    K, dK = kern.kernel_and_grad(X)
    K_ref, dK_ref, h = orc.imq_kernel_and_grad(X, beta=beta)
    assert np.float32(kern.bandwidth).tobytes() == h.tobytes()
    assert np.abs(K - K_ref).max() <= 2e-5
    assert np.abs(dK - dK_ref).max() <= 1e-4 * np.abs(dK_ref).max()


@pytest.mark.parametrize("kind", ["imq_device", "user_numpy"])
def test_sampler_runs_another_kernel_operator(ctx, kind):
    """`sampler.kernel = <another AbstractKernel>` (the reference's plugin point): the iteration becomes
    phi = (K S + dK) / n from that operator (abstract_stein_sampler.py:100-105), clip, optimizer --
    with the built-in IMQ operator (CUDA) and with a user-defined operator that returns NumPy arrays."""
    from stein_b200.kernels import AbstractKernel, InverseMultiquadricKernel
    from stein_b200.log_p import LinearRegression
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.samplers import SteinSampler

    class NumpyIMQ(AbstractKernel):            # what a user of the reference would write
        def kernel_and_grad(self, theta):
            K, dK, _ = orc.imq_kernel_and_grad(theta, beta=-0.5)
            return K, dK

    n, F, N = 200, 7, 64
    rng = np.random.default_rng(6)
    Xd = rng.standard_normal((N, F)).astype(np.float32)
    yd = rng.standard_normal((N, 1)).astype(np.float32)
    model = LinearRegression(F)
    np.random.seed(5)
    sampler = SteinSampler(n, model.log_p, AdamGradientDescent(learning_rate=0.05))
    sampler.kernel = InverseMultiquadricKernel(n) if kind == "imq_device" else NumpyIMQ(n)
    theta = sampler.samples.copy()
    gd = orc.AdamGradientDescent(learning_rate=0.05)
    for _ in range(3):
        S = orc.score_linear(theta, Xd, yd)
        K, dK, _ = orc.imq_kernel_and_grad(theta, beta=-0.5)
        phi = orc.clip((K.astype(np.float64) @ S + dK) / n)
        theta = theta + gd.update(phi)
        sampler.train_on_batch({model.X: Xd, model.y: yd})
    _assert_close(sampler.samples, theta, 2e-4)
    # compute_phi(theta, grads) with the plugged operator
    S = orc.score_linear(theta, Xd, yd)
    K, dK, _ = orc.imq_kernel_and_grad(theta, beta=-0.5)
    _assert_close(sampler.compute_phi(theta, S), (K.astype(np.float64) @ S + dK) / n)


# --------------------------------------------------------------------------- #
# round 2: graph recognition (the reference's scripts leave torch autograd)     #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("kind", ["linear", "linear_1feature", "logistic", "bnn", "unknown"])
def test_graph_recognition_dispatches_to_the_score_kernels(ctx, kind):
    """A `log_p` tensor recorded by the TF1 stand-in from one of the reference's three example graphs
    (examples/*/main.py, restated in tests/test_tf_compat.py) is recognised -- by signature, then verified
    against its own autograd scores -- and served by stein_score_linear / _logistic / _bnn; the training-set
    size baked into the graph is recovered.  Same scores as the autograd path to 2e-5; a graph that is
    none of the three stays on autograd."""
    import sys
    import torch
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat")
    if compat not in sys.path:
        sys.path.insert(0, compat)
    import tensorflow as tf
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import test_tf_compat as graphs
    from stein_b200.optimizers import AdamGradientDescent
    from stein_b200.samplers import SteinSampler
    rng = np.random.default_rng(9)
    n, B = 96, 40
    if kind in ("linear", "linear_1feature"):
        F = 1 if kind == "linear_1feature" else 6
        g = graphs.linear_graph(F)
    elif kind == "logistic":
        F = 9
        g = graphs.logistic_graph(F, 464809.0, float(B))
    elif kind == "bnn":
        F, H = 5, 12
        g = graphs.bnn_graph(F, H, 455.0, float(B))
    else:
        F = 4
        tf.reset_default_graph()
        with tf.variable_scope("model"):
            X = tf.placeholder(tf.float32, shape=[None, F])
            y = tf.placeholder(tf.float32, shape=[None, 1])
            w = tf.Variable(tf.zeros([F, 1]))
            log_p = -tf.reduce_sum(tf.square(tf.matmul(X, w) - y)) - 3.0 * tf.reduce_sum(tf.square(w))   # not the example
        g = dict(X=X, y=y, log_p=log_p)
    Xb = rng.standard_normal((B, F)).astype(np.float32)
    yb = (rng.random((B, 1)) > 0.5).astype(np.float32) if kind == "logistic" else rng.standard_normal((B, 1)).astype(np.float32)
    np.random.seed(7)
    sampler = SteinSampler(n, g["log_p"], AdamGradientDescent(learning_rate=1e-2))
    lp, e = sampler.log_p, sampler.engine
    feed = {g["X"]: Xb, g["y"]: yb}
    lp.scores(e, feed)                                   # first call: probe + recognition
    d = lp.n_params
    S_used = e.scores_dev[:n, :d].clone()
    S_auto = lp.graph_scores(e.particles_dev[:n, :d], lp._feed_tensors(feed, e.ctx.dense)).to(torch.float32)
    want = {"linear": "linear", "linear_1feature": "linear", "logistic": "logistic", "bnn": "bnn", "unknown": None}[kind]
    assert lp.recognised == want, lp.recognised
    scale = S_auto.abs().amax(dim=0).clamp_min(1e-30)
    assert float(((S_used - S_auto).abs() / scale).amax()) <= 2e-5
    if kind == "logistic":
        assert lp._fast[0].n_train == 464809.0
    if kind == "bnn":
        assert lp._fast[0].n_train == 455.0
    # a second batch of another size goes through the same kernel
    Xb2, yb2 = Xb[:17], yb[:17]
    lp.scores(e, {g["X"]: Xb2, g["y"]: yb2})
    S2 = e.scores_dev[:n, :d].clone()
    S2_auto = lp.graph_scores(e.particles_dev[:n, :d], lp._feed_tensors({g["X"]: Xb2, g["y"]: yb2}, e.ctx.dense))
    scale2 = S2_auto.abs().amax(dim=0).clamp_min(1e-30)
    # (the logistic / bnn graphs bake n_batch in: a batch of another size goes back to autograd -- exact either way)
    assert float(((S2 - S2_auto.to(torch.float32)).abs() / scale2).amax()) <= 2e-5


def test_median_pilotless_steady_state_is_bit_exact(ctx):
    """While the median drifts slowly (tiny steps) an engine's iterations skip the pilot sample and reuse the
    last window, recentred on the last exact median; the result is still the oracle's, bit for bit, and a
    jump of the particles falls back to the pilot routes (compute_median.py:4-16)."""
    from stein_b200.engine import SvgdEngine
    n, d = 4608, 256
    X = _particles(n, d, 41)
    eng = SvgdEngine(n, d, "adam", learning_rate=1e-6)
    hits0, miss0 = ctypes.c_longlong(), ctypes.c_longlong()
    ctx.lib.stein_debug_median_direct_stats(ctypes.byref(hits0), ctypes.byref(miss0))
    eng.set_particles(X)
    for it in range(6):
        Xin = eng.get_particles(np.float32)
        eng.set_scores(-Xin)
        eng.step()
        m_ref, _ = orc.median_chain(Xin, radix=True)
        assert np.float32(eng.last()["median"]).tobytes() == m_ref.tobytes(), it
    hits, miss = ctypes.c_longlong(), ctypes.c_longlong()
    ctx.lib.stein_debug_median_direct_stats(ctypes.byref(hits), ctypes.byref(miss))
    assert hits.value - hits0.value >= 3 and miss.value == miss0.value, (hits.value - hits0.value, miss.value - miss0.value)
    # a jump: the pilot-less attempt misses once, the step still returns the exact median
    eng.particles_dev.mul_(3.0)
    Xin = eng.get_particles(np.float32)
    eng.set_scores(-Xin)
    eng.step()
    m_ref, _ = orc.median_chain(Xin, radix=True)
    assert np.float32(eng.last()["median"]).tobytes() == m_ref.tobytes()
    ctx.lib.stein_debug_median_direct_stats(ctypes.byref(hits), ctypes.byref(miss))
    assert miss.value - miss0.value == 1
    eng.close()


def test_update_particles_host_prefetches_the_next_median(ctx):
    """update_particles_host enqueues the next iteration's head and median behind the optimizer kernel (they need
    only the particles) so that they run while the particles cross PCIe.  Same bits as the plain step sequence:
    particles after every iteration, median and bandwidth; exact against the oracle's median rule
    (abstract_stein_sampler.py:107-127, compute_median.py:4-16); dropped when the particles are replaced."""
    from stein_b200.engine import SvgdEngine
    n, d, iters = 4608, 256, 8
    X = _particles(n, d, 43)

    def run(prefetch, device_bw=True):
        eng = SvgdEngine(n, d, "adam", learning_rate=1e-4)
        eng.set_prefetch(prefetch)
        eng.set_device_bandwidth(device_bw)
        eng.set_particles(X)
        out, meds = [], []
        Xin = X.copy()
        for it in range(iters):
            Xout = np.empty_like(Xin)
            eng.update_particles_host(-Xin, Xout)
            info = eng.last()
            if it in (0, iters - 1) or prefetch:
                m_ref, _ = orc.median_chain(Xin, radix=True)
                assert np.float32(info["median"]).tobytes() == m_ref.tobytes(), (prefetch, it)
            out.append(Xout)
            meds.append((info["median"], info["bandwidth"], info["sweeps"]))
            Xin = Xout
        stats = eng.prefetch_stats()
        return eng, out, meds, stats

    eng, out_a, meds_a, stats_a = run(True)
    assert stats_a["begun"] >= 3 and stats_a["used"] == stats_a["begun"] - 1, stats_a   # (the last one is still pending)
    # ... and phi ran ahead of the host's median result, on the device-side select (verified by the host each time)
    dbw = eng.device_bandwidth_stats()
    assert dbw["used"] >= 3 and dbw["redone"] == 0, dbw
    # the pending median belongs to the particles it was computed from: new particles drop it ...
    Y = _particles(n, d, 44, 1.7)
    eng.set_particles(Y)
    Yout = np.empty_like(Y)
    eng.update_particles_host(-Y, Yout)
    m_ref, _ = orc.median_chain(Y, radix=True)
    assert np.float32(eng.last()["median"]).tobytes() == m_ref.tobytes()
    # ... and so does a write through the device view that is announced
    for _ in range(3):
        Yin = Yout
        Yout = np.empty_like(Yin)
        eng.update_particles_host(-Yin, Yout)
    assert eng.prefetch_stats()["begun"] > stats_a["begun"]
    eng.particles_dev.mul_(1.5)
    eng.particles_changed()
    Z = eng.get_particles(np.float32)
    eng.set_scores(-Z)
    eng.step()
    m_ref, _ = orc.median_chain(Z, radix=True)
    assert np.float32(eng.last()["median"]).tobytes() == m_ref.tobytes()
    eng.close()

    eng, out_b, meds_b, stats_b = run(False)
    assert eng.device_bandwidth_stats()["used"] >= 3
    eng.close()
    assert stats_b == {"begun": 0, "used": 0}
    assert meds_a == meds_b
    for a, b in zip(out_a, out_b):
        np.testing.assert_array_equal(a, b)
    # the plain sequence (host collects the median, then launches phi): same bits again
    eng, out_c, meds_c, stats_c = run(False, device_bw=False)
    assert eng.device_bandwidth_stats() == {"used": 0, "redone": 0} and stats_c == {"begun": 0, "used": 0}
    eng.close()
    assert meds_a == meds_c
    for a, c in zip(out_a, out_c):
        np.testing.assert_array_equal(a, c)


def test_two_engines_interleaved_keep_their_own_median_history(ctx):
    """Two engines stepping alternately in one process: each keeps its own median history (window hint, pilot-less
    steady state) and the prefetched median of one is void once the other has used the arena.  Every engine's
    trajectory has the bits of its solo run; the medians are the oracle's (compute_median.py:4-16)."""
    from stein_b200.engine import SvgdEngine
    n, d, iters = 4608, 256, 7
    X0 = {"a": _particles(n, d, 51), "b": _particles(n, d, 52, 2.0)}

    def solo(X):
        eng = SvgdEngine(n, d, "adam", learning_rate=1e-5)
        eng.set_particles(X)
        outs, meds, Xin = [], [], X
        for _ in range(iters):
            Xout = np.empty_like(Xin)
            eng.update_particles_host(-Xin, Xout)
            outs.append(Xout)
            meds.append(eng.last()["median"])
            Xin = Xout
        eng.close()
        return outs, meds

    ref = {k: solo(X) for k, X in X0.items()}
    hits0, miss0 = ctypes.c_longlong(), ctypes.c_longlong()
    ctx.lib.stein_debug_median_direct_stats(ctypes.byref(hits0), ctypes.byref(miss0))
    engs = {k: SvgdEngine(n, d, "adam", learning_rate=1e-5) for k in X0}
    cur = dict(X0)
    for k in engs:
        engs[k].set_particles(X0[k])
    for it in range(iters):
        for k in ("a", "b"):
            Xout = np.empty_like(cur[k])
            engs[k].update_particles_host(-cur[k], Xout)
            assert np.float32(engs[k].last()["median"]).tobytes() == np.float32(ref[k][1][it]).tobytes(), (k, it)
            np.testing.assert_array_equal(Xout, ref[k][0][it])
            if it == iters - 1:
                m_ref, _ = orc.median_chain(cur[k], radix=True)
                assert np.float32(engs[k].last()["median"]).tobytes() == m_ref.tobytes()
            cur[k] = Xout
    hits, miss = ctypes.c_longlong(), ctypes.c_longlong()
    ctx.lib.stein_debug_median_direct_stats(ctypes.byref(hits), ctypes.byref(miss))
    assert hits.value - hits0.value >= 4 and miss.value == miss0.value       # both reached the pilot-less state
    for e in engs.values():
        e.close()
