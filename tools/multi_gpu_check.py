"""torchrun --nproc-per-node N tools/multi_gpu_check.py
Particle-sharded SVGD steps on N GPUs (NCCL) against the single-process CPU oracle:
bandwidth bit-exact on every rank, post-step particles within 1e-4."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import svgd_oracle as orc  # noqa: E402
from stein_b200.distributed import make_comm  # noqa: E402
from stein_b200.engine import SvgdEngine  # noqa: E402
from stein_b200.runtime import context  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = context(local)
    make_comm(ctx)
    ok = True
    cases = [(1000, 33, True), (5000, 256, True), (5000, 256, False), (4097, 128, True), (4200, 1024, True),
             (4500, 600, True)]
    if os.environ.get("MGC_QUICK"):
        cases = [(1000, 33, True), (5000, 256, True), (4200, 1024, True)]
    for n, d, push in cases:
        rng = np.random.default_rng(n)
        X = rng.standard_normal((n, d)).astype(np.float32).astype(np.float64)
        mean = rng.standard_normal(d)
        eng = SvgdEngine(n, d, "adam", learning_rate=0.1, decay=0.999, ctx=ctx, peer_push=push)
        b, nl = eng.row_begin, eng.n_local
        eng.set_particles(np.ascontiguousarray(X[b:b + nl]))
        gd = orc.AdamGradientDescent(0.1, 0.999)
        Xref = X.copy()
        for it in range(6):
            S = (mean - Xref) * 2.0
            bw_ref = orc.kernel_and_grad(Xref)[2] if n <= 1500 else None
            out = np.empty((nl, d))
            eng.update_particles_host(np.ascontiguousarray(S[b:b + nl]), out)
            info = eng.last()
            if n <= 1500:
                Xref, _ = orc.update_particles(Xref, S, gd)
                good = np.float32(info["bandwidth"]).tobytes() == bw_ref.tobytes()
                err = np.abs(out - Xref[b:b + nl]).max() / np.abs(Xref).max()
                good = good and err < 1e-4
            else:
                # larger n: compare ranks against each other through the oracle's row blocks
                if nl > 0:
                    rows, _ = orc.phi_rows_c(Xref, S, np.float32(info["bandwidth"]), b, min(b + 64, b + nl))
                    phi = eng.get_phi(np.float64)[:rows.shape[0]]
                    err = np.abs(phi - rows).max() / np.abs(rows).max()
                else:       # more ranks than 128-row tiles: this rank owns no particle
                    err = 0.0
                good = err < 1e-4
                # the exact median of the reference's rule (compute_median.py:4-16), bit for bit
                m_ref, _ = orc.median_chain(Xref.astype(np.float32), radix=True)
                good = good and np.float32(info["median"]).tobytes() == m_ref.tobytes()
                # all ranks must hold the same bandwidth bits
                t = torch.tensor([np.float32(info["bandwidth"]).view(np.int32).item()], device="cuda")
                lo, hi = t.clone(), t.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                good = good and lo.item() == hi.item()
                # advance the reference with the gathered device particles
                full = [None] * world
                dist.all_gather_object(full, out)
                Xref = np.concatenate(full, axis=0)
            print("rank %d n=%d d=%d push=%s iter %d: bandwidth %.9g sweeps %d err %.2e %s"
                  % (rank, n, d, eng.peer_push, it, info["bandwidth"], info["sweeps"], err, "ok" if good else "FAIL"),
                  flush=True)
            ok = ok and good
        if rank == 0:
            print("n=%d d=%d push=%s: prefetched medians %r, phi ahead of the host's median %r"
                  % (n, d, eng.peer_push, eng.prefetch_stats(), eng.device_bandwidth_stats()), flush=True)
        eng.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
