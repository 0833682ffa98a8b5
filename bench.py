#!/usr/bin/env python
"""bench.py -- SVGD iterations/s and kernel interactions/s on BASELINE.json's
headline configuration: synthetic standard-normal target, n = 65 536 particles,
d = 256, 1..8 B200 (particles row-sharded, total work fixed => strong scaling).

One "step" = one full SVGD iteration: score kernel -> (all-gather) -> row norms
-> exact median / bandwidth -> phi -> (all-reduce) -> clip + Adam step.

  python bench.py --gpus N --steps K --warmup W             (our arm)
  python bench.py --impl reference --gpus N --steps K ...    (CPU restatement of the
        reference's path on the host cores; the reference itself needs TensorFlow 1.12
        and cannot run -- DESIGN.md "Oracle")

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_PARTICLES, DIM = 65536, 256
METRIC, UNIT = "svgd_iterations_per_sec", "iterations/s"


def algorithmic_flops_phi(n, d):
    """one T.T^T, one K.S, one K.X, one row sum = 2 n^2 (3d + 1)."""
    return 2.0 * n * n * (3 * d + 1)


def ncu_traffic(kernel_substr):
    """DRAM bytes per launch (read + write) of the dominant kernel from the committed ncu
    --set full summary (profiles/, written by tools/ncu_summary.py), or None."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")
    try:
        data = json.load(open(path))
    except (OSError, ValueError):
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for rec in data.get("launches", []):
        if kernel_substr in rec.get("kernel", ""):
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if key not in rec:
                    return None, None
                tot += float(rec[key]) * scale.get(rec.get("_units", {}).get(key, "byte"), 1.0)
            return tot, os.path.relpath(path, ROOT)
    return None, None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        # the busiest half of the samples = "under load"
        sm_sorted = sorted(sm)
        return {"sm_mhz": (sm_sorted[len(sm_sorted) // 2] if sm_sorted else None),
                "sm_max_mhz": (max(mx) if mx else None), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_baseline(n, d, target_seconds=15.0):
    """The NumPy/BLAS restatement of the reference's iteration (oracle/, kind "port"),
    timed on a bounded row slab (rows x all n columns) and scaled by n/rows."""
    from oracle import svgd_oracle as orc
    try:
        from threadpoolctl import threadpool_info
        blas_threads = max([p.get("num_threads", 1) for p in threadpool_info()
                            if p.get("user_api") == "blas"] or [os.cpu_count() or 1])
    except Exception:
        blas_threads = os.cpu_count() or 1
    rng = np.random.default_rng(1)
    X = rng.standard_normal((n, d)).astype(np.float32)
    S = (-X).astype(np.float64)
    rows = 256
    t0 = time.perf_counter()
    orc.iteration_blocked_numpy(X, S, row_block=256, rows=rows)
    t_probe = time.perf_counter() - t0
    rows = int(min(n, max(256, (target_seconds / max(t_probe, 1e-3)) * 256 // 256 * 256)))
    rows = min(rows, 8192)          # rows * n * 4 B of fp32 distances stay resident
    t0 = time.perf_counter()
    orc.iteration_blocked_numpy(X, S, row_block=1024, rows=rows)
    t = time.perf_counter() - t0
    it_s = 1.0 / (t * n / rows)
    return {"value": it_s, "unit": UNIT, "cores": int(blas_threads), "kind": "port",
            "host_cpus": os.cpu_count(),
            "sample": "rows 0..%d of %d x all %d columns (fp32 sgemm distances, exact median of the "
                      "slab, exp, float64 K.dot(S)); %.2f s measured, scaled by n/rows" % (rows, n, n, t),
            "interactions_per_sec": it_s * n * n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, d = args.n, args.d
    vals = []
    sample = None
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(n, d, target_seconds=args.ref_seconds)
        sample = cb
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32/f64",
            "data": "synthetic",
            "config": {"workload": "gaussian_target_n%d_d%d" % (n, d), "n_particles": n, "dim": d,
                       "note": "reference needs TensorFlow 1.12 (not installable); this is the NumPy/BLAS "
                               "port of its per-iteration math on the host cores"},
            "interactions_per_sec": v * n * n,
            "cpu_baseline": {k: sample[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["cpu_baseline"]["value"] = v
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from stein_b200 import _lib
    from stein_b200.engine import SvgdEngine
    from stein_b200.log_p import GaussianMixtureTarget
    from stein_b200.runtime import context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = context(local_rank)
    if world > 1:
        from stein_b200.distributed import make_comm
        make_comm(ctx)
    if args.phi_impl:
        ctx.set_phi_impl({"auto": 0, "dense": 1, "flash": 2, "flash2": 3, "flash3": 4, "flash4": 5}[args.phi_impl])

    n, d = args.n, args.d
    eng = SvgdEngine(n, d, "adam", learning_rate=1e-2, ctx=ctx)
    rng = np.random.default_rng(1)
    X0 = rng.standard_normal((n, d)).astype(np.float32)
    X_local = np.ascontiguousarray(X0[eng.row_begin:eng.row_begin + eng.n_local])
    eng.set_particles(X_local)
    model = GaussianMixtureTarget(d)
    dev = torch.device("cuda", local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        model.scores(eng)          # S = -X on the device
        eng.step()

    for _ in range(args.warmup):
        one_step()
    barrier()

    # ---- device-resident timing: K steps, each bracketed by events, L2 flushed between
    ctx.check(ctx.lib.stein_ctx_profile_enable(ctx.handle, 1))
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()
        a.record()
        one_step()
        b.record()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    import ctypes
    phi_ms, phi_n = ctypes.c_double(), ctypes.c_int64()
    sw_ms, sw_n = ctypes.c_double(), ctypes.c_int64()
    ctx.check(ctx.lib.stein_ctx_profile_read(ctx.handle, 0, ctypes.byref(phi_ms), ctypes.byref(phi_n)))
    ctx.check(ctx.lib.stein_ctx_profile_read(ctx.handle, 1, ctypes.byref(sw_ms), ctypes.byref(sw_n)))
    ctx.check(ctx.lib.stein_ctx_profile_enable(ctx.handle, 0))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    info = eng.last()

    # ---- end to end through the host-buffer entry point (what a NumPy caller of
    # update_particles(grads_array) sees): pinned fp32 scores in, particles out
    S_host = torch.from_numpy(-X_local).pin_memory()
    X_out = torch.empty_like(S_host).pin_memory()
    S_np, X_np = S_host.numpy(), X_out.numpy()
    for _ in range(2):
        eng.update_particles_host(S_np, X_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.update_particles_host(S_np, X_np)       # returns after the D2H copy completed
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    n_local, peer_push = eng.n_local, eng.peer_push
    eng.close()          # collective when the peers are connected (every rank is idle here)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    ms_per_step = total_ms / args.steps
    value = 1000.0 / ms_per_step
    # roofline of the dominant kernel (phi): algorithmic FLOPs of the LOCAL row block
    f_phi = algorithmic_flops_phi(n, d) * (n_local / float(n))
    phi_avg_ms = phi_ms.value / max(phi_n.value, 1)
    achieved = f_phi / (phi_avg_ms * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    code = args_phi_impl_code(ctx, args)
    if code == 0:       # what AUTO resolves to (stein_b200/csrc/ctx.cu pick_phi_impl)
        ldp = (d + 31) // 32 * 32
        code = 5 if ldp == 256 else (2 if ldp == 128 else 1)
    impl = {1: "dense_simt_fp32", 2: "flash_tcgen05", 3: "flash_tcgen05_cta_pair",
            4: "flash_tcgen05_cta_pair_fp8_gemm2", 5: "flash_tcgen05_cta_pair_fp16_fp8"}[code]
    executed = {1: "FP32 FFMA", 2: "3 BF16 passes per GEMM", 3: "3 BF16 passes per GEMM",
                4: "GEMM1 3 BF16 passes; GEMM2 1 FP16 + 2 FP8 passes",
                5: "1 FP16 + 2 FP8 passes per GEMM (= 2 BF16-pass equivalents of tensor time each)"}[code]
    traffic, traffic_src = ncu_traffic("flash_phi2_kernel") if (code >= 3 and world == 1 and (n, d) == (N_PARTICLES, DIM)) \
        else (None, None)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gaussian_target_n%d_d%d" % (n, d), "n_particles": n, "dim": d,
                   "optimizer": "adam", "parallelism": "particle_rows_x%d" % world,
                   "collectives": ("none" if world == 1 else
                                   "scores: NCCL all-gather on a side stream; particles: pushed to the peers by "
                                   "the optimizer kernel; small all-reduces: %s"
                                   % ("one kernel each over NVLink peer memory"
                                      if os.environ.get("STEIN_PEER_REDUCE", "1") != "0" else "NCCL")
                                   if peer_push else "NCCL (library-driven)"),
                   "phi_impl": impl, "median_sweeps_last_step": info["sweeps"],
                   "l2": "256 MiB buffer written between timed steps (L2 flush); per-step working set "
                         "is also > 126 MB"},
        "interactions_per_sec": value * n * n,
        "gpu_launches": int(launches),
        "wall_s_timed_region": wall,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "phi (%s)" % impl, "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peaks["source"] + ", bf16 dense sustained",
                     "executed_arithmetic": executed,
                     "algorithmic_flops_per_launch": f_phi, "avg_launch_ms": phi_avg_ms,
                     "launches_timed": int(phi_n.value),
                     "share_of_step": phi_ms.value / total_ms if world == 1 else None,
                     "median_sweep_ms_per_step": sw_ms.value / args.steps,
                     "median_sweeps_timed": int(sw_n.value)},
        "e2e": {"value": args.steps / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": int(S_np.nbytes), "d2h_bytes_per_step": int(X_np.nbytes),
                "call": "stein_engine_update_particles_host (pinned fp32 scores H2D -> iteration -> "
                        "particles D2H), max over ranks of wall time"},
        "bandwidth_last_step": info["bandwidth"],
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(n, d, target_seconds=args.ref_seconds)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def args_phi_impl_code(ctx, args):
    if args.phi_impl:
        return {"auto": 0, "dense": 1, "flash": 2, "flash2": 3, "flash3": 4, "flash4": 5}[args.phi_impl]
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_PARTICLES)
    ap.add_argument("--d", type=int, default=DIM)
    ap.add_argument("--phi-impl", default=None, choices=[None, "auto", "dense", "flash", "flash2", "flash3", "flash4"])
    ap.add_argument("--ref-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
