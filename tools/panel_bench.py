"""Timing of the panel phi kernels and of the wide median sweep on one GPU at config-E-like shapes:
  python tools/panel_bench.py [n] [d] [reps]     (environment STEIN_PANEL_TILES = P block budget in tiles)"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from stein_b200.runtime import context
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    n_local = int(sys.argv[4]) if len(sys.argv) > 4 else n          # rows of one rank's shard (columns: all n)
    ctx = context()
    vp = ctypes.c_void_p
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    rows, ld = ctx.rows_padded(n), -(-d // 256) * 256
    X = torch.zeros((rows, ld), device=dev)
    X[:n, :d] = torch.randn((n, d), generator=g, device=dev)
    S = -X
    r = torch.empty(rows, dtype=torch.float32, device=dev)
    ctx.check(ctx.lib.stein_row_norms(ctx.handle, vp(X.data_ptr()), n, d, ld, vp(r.data_ptr())))
    med = ctypes.c_float()
    sw = ctypes.c_int32()
    ctx.check(ctx.lib.stein_ctx_profile_enable(ctx.handle, 1))
    ms, cnt = ctypes.c_double(), ctypes.c_int64()
    if os.environ.get("STEIN_SKIP_MEDIAN") == "1":
        med.value = 2.0 * d          # E|x - y|^2 of a standard normal cloud: skips the exact median (phi timing only)
    else:
        time_median(ctx, X, r, n, d, ld, med, sw, ms, cnt)
    run_phi(ctx, X, S, r, n, d, ld, n_local, reps, med, ms, cnt)


def time_median(ctx, X, r, n, d, ld, med, sw, ms, cnt):
    import torch
    vp = ctypes.c_void_p
    t0 = time.perf_counter()
    ctx.check(ctx.lib.stein_median_sqdist(ctx.handle, vp(X.data_ptr()), vp(r.data_ptr()), n, d, ld, ctypes.byref(med), None,
                                          ctypes.byref(sw)))
    torch.cuda.synchronize()
    t_med = time.perf_counter() - t0
    ctx.check(ctx.lib.stein_ctx_profile_read(ctx.handle, 1, ctypes.byref(ms), ctypes.byref(cnt)))
    print("median %.6g: %.1f ms wall (%d sweeps, sweep kernels %.2f ms = %.0f TFLOP/s executed in 3 FP16 passes)"
          % (med.value, t_med * 1e3, sw.value, ms.value, 3 * n * n * ld / (ms.value * 1e-3) / 1e12 if ms.value else 0))


def run_phi(ctx, X, S, r, n, d, ld, n_local, reps, med, ms, cnt):
    import torch
    vp = ctypes.c_void_p
    dev = X.device
    bw = ctx.lib.stein_bandwidth(med.value, n)
    nb = int(ctx.lib.stein_phi_workspace_bytes(ctx.handle, n_local, n, ld))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    phi = torch.empty((ctx.rows_padded(n_local), ld), device=dev)
    sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
    for rep in range(reps):
        ctx.lib.stein_ctx_profile_read(ctx.handle, 0, None, None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.check(ctx.lib.stein_phi(ctx.handle, vp(X.data_ptr()), vp(S.data_ptr()), vp(r.data_ptr()), n, d, ld, 0, n_local, bw,
                                    vp(ws.data_ptr()), nb, vp(phi.data_ptr()), vp(sumsq.data_ptr())))
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        ctx.check(ctx.lib.stein_ctx_profile_read(ctx.handle, 0, ctypes.byref(ms), ctypes.byref(cnt)))
        alg = 2.0 * n_local * n * (3 * d + 1)
        print("phi: %.2f ms wall, panel kernels %.2f ms: %.0f TFLOP/s algorithmic, %.0f executed (bf16-pass equivalents), "
              "route %s" % (t * 1e3, ms.value, alg / (ms.value * 1e-3) / 1e12, 8.0 * n_local * n * ld / (ms.value * 1e-3) / 1e12,
                            ctx.phi_route()["route"]), flush=True)


if __name__ == "__main__":
    main()
