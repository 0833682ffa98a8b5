"""GPU experiment (not a test): where does the flash-phi error come from?
Usage: python tools/phi_error_study.py   (on a B200 box)"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import svgd_oracle as orc  # noqa: E402
from stein_b200 import _lib  # noqa: E402
from stein_b200.runtime import context  # noqa: E402

ctx = context()


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def phi_gpu(X, S, bw, impl):
    n, d = X.shape
    Xd, Sd = ctx.to_padded(X), ctx.to_padded(S)
    rows, ld = Xd.shape
    r = torch.empty(rows, dtype=torch.float32, device=Xd.device)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, P(Xd), n, d, ld, P(r)))
    ctx.set_phi_impl(impl)
    nb = int(ctx.lib.stein_phi_workspace_bytes(ctx.handle, n, n, d))
    ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
    phi = torch.empty_like(Xd)
    sumsq = torch.zeros(1, dtype=torch.float64, device=Xd.device)
    ctx.check(ctx.lib.stein_phi(ctx.handle, P(Xd), P(Sd), P(r), n, d, ld, 0, n, float(bw), P(ws), nb, P(phi),
                                P(sumsq)))
    ctx.set_phi_impl(0)
    return phi.cpu().numpy()[:n, :d].astype(np.float64)


def phi_ref(X, S, bw):
    X64, S64 = X.astype(np.float64), S.astype(np.float64)
    r = (X64 ** 2).sum(1)
    D = r[:, None] + r[None, :] - 2 * X64 @ X64.T
    h2 = float(np.float32(bw) * np.float32(bw))
    K = np.exp(-D / h2 / 2)
    return (K @ S64 + (X64 * K.sum(1)[:, None] - K @ X64) / h2) / X.shape[0]


def report(tag, a, b):
    rel = np.linalg.norm(a - b) / np.linalg.norm(b)
    # best scalar c with a ~ c*b
    c = (a * b).sum() / (b * b).sum()
    res = np.linalg.norm(a - c * b) / np.linalg.norm(b)
    print("%-44s fro %.3e   scale-1 %+.3e   residual after scale %.3e" % (tag, rel, c - 1, res))


rng = np.random.default_rng(0)
for n, d in [(128, 256), (1024, 256), (4096, 256)]:
    X = rng.standard_normal((n, d)).astype(np.float32)
    S = rng.standard_normal((n, d)).astype(np.float32)
    bw = float(orc.bandwidth(np.float32(2 * d), n))
    ref = phi_ref(X, S, bw)
    report("n=%d random S, dense" % n, phi_gpu(X, S, bw, 1), ref)
    report("n=%d random S, flash" % n, phi_gpu(X, S, bw, 2), ref)
    # E1: exactly representable operands: X = 0 (K = 1), S small integers
    X0 = np.zeros_like(X)
    Si = rng.integers(-3, 4, size=(n, d)).astype(np.float32)
    report("n=%d X=0, integer S, flash" % n, phi_gpu(X0, Si, 1.0, 2), phi_ref(X0, Si, 1.0))
    # E2: K = 1, full-mantissa S
    report("n=%d X=0, random S, flash" % n, phi_gpu(X0, S, 1.0, 2), phi_ref(X0, S, 1.0))
    # E3: random X, S = 1 (O = ksum): tests P
    S1 = np.ones_like(X)
    report("n=%d random X, S=1, flash" % n, phi_gpu(X, S1, bw, 2), phi_ref(X, S1, bw))
    report("n=%d random X, S=-X, flash" % n, phi_gpu(X, -X, bw, 2), phi_ref(X, -X, bw))
    report("n=%d random X, S=-X, dense" % n, phi_gpu(X, -X, bw, 1), phi_ref(X, -X, bw))
