"""Is the tensor-core Gram matrix biased?  Raw GEMM1 output of the single-CTA flash kernel
(3 BF16 passes, fp32 accumulation in tensor memory) against float64 X X^T on a two-cluster cloud:
signed relative error of the large same-cluster entries (mean = bias, std = noise)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from stein_b200.runtime import context
    ctx = context()
    vp = ctypes.c_void_p
    for n, d in [(1024, 256), (1024, 128)]:
        rng = np.random.default_rng(1)
        Z = rng.standard_normal((n, d))
        X = (np.where((rng.random(n) < 0.6)[:, None], 0.8, -1.2) + 0.05 * Z).astype(np.float32)
        Xd, Sd = ctx.to_padded(X), ctx.to_padded(np.zeros_like(X))
        rows, ld = Xd.shape
        r = torch.empty(rows, dtype=torch.float32, device=Xd.device)
        ctx.check(ctx.lib.stein_row_norms(ctx.handle, vp(Xd.data_ptr()), n, d, ld, vp(r.data_ptr())))
        nb = int(ctx.lib.stein_phi_workspace_bytes(ctx.handle, n, n, d))
        ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
        phi = torch.empty_like(Xd)
        sumsq = torch.zeros(1, dtype=torch.float64, device=Xd.device)
        G = torch.zeros((rows, rows), dtype=torch.float32, device=Xd.device)
        fn = ctx.lib.stein_debug_flash_gram
        fn.restype = ctypes.c_int
        fn.argtypes = [vp] * 4 + [ctypes.c_int64] * 3 + [ctypes.c_float, vp, ctypes.c_int64] + [vp] * 3
        ctx.check(fn(ctx.handle, vp(Xd.data_ptr()), vp(Sd.data_ptr()), vp(r.data_ptr()), n, d, ld, 10.0, vp(ws.data_ptr()), nb,
                     vp(phi.data_ptr()), vp(sumsq.data_ptr()), vp(G.data_ptr())))
        got = G.cpu().numpy()[:n, :n].astype(np.float64)
        X64 = X.astype(np.float64)
        ref = X64 @ X64.T
        # what an exact accumulation of the 3 BF16 passes would give (isolates the accumulation error)
        t = torch.from_numpy(X)
        hi = t.bfloat16().float()
        lo = (t - hi).bfloat16().float()
        hi64, lo64 = hi.double().numpy(), lo.double().numpy()
        split = hi64 @ hi64.T + lo64 @ hi64.T + hi64 @ lo64.T
        big = ref > 0.5 * ref.max()
        rel_total = (got - ref)[big] / ref[big]
        rel_acc = (got - split)[big] / ref[big]
        print("n=%d d=%d: entries %d | total error mean %.3e std %.3e | accumulation-only error mean %.3e std %.3e "
              "(2^-24 = %.2e, MMAs per entry = %d)" % (n, d, big.sum(), rel_total.mean(), rel_total.std(), rel_acc.mean(),
                                                        rel_acc.std(), 2.0 ** -24, 3 * d // 16))


if __name__ == "__main__":
    main()
