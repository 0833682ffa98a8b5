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
    path = os.path.join(ROOT, "profiles", "r02_ncu_full_summary.json")
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
    """SM clock and throttle reasons DURING the timed region: NVML polled from a thread every few milliseconds
    (the timed region is only 0.1-0.2 s long: `nvidia-smi -lms` needs longer than that to print its first line);
    nvidia-smi is the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.thread = index, [], None, None, None
        self.stop_flag = threading.Event()
        self.samples = []          # (sm MHz, reasons bitmask)
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = int(reasons_fn(self.handle)) if reasons_fn else 0
                self.samples.append((mhz, mask))
            except Exception:
                pass
            time.sleep(0.004)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            sm = sorted(m for m, _ in self.samples)
            mask = 0
            for _, k in self.samples:
                mask |= k
            return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_min_mhz": (sm[0] if sm else None),
                    "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": sorted(nm for nm, bit in self.BITS.items() if mask & bit), "source": "nvml, 4 ms period"}
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
                "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


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


def workload_config(n, d, world):
    """The `config` object both arms print (same keys, same values for the same workload)."""
    return {"workload": "gaussian_target_n%d_d%d" % (n, d), "n_particles": n, "dim": d,
            "optimizer": "adam", "parallelism": "particle_rows_x%d" % world}


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
            "config": workload_config(n, d, args.gpus),
            "note": "reference needs TensorFlow 1.12 (not installable); this is the NumPy/BLAS port of its "
                    "per-iteration math on the host cores, timed on a bounded row slab and scaled by n/rows",
            "interactions_per_sec": v * n * n,
            "cpu_baseline": {k: sample[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["cpu_baseline"]["value"] = v
    print(json.dumps(line))


def bind_to_gpu_numa_node(local_rank):
    """Best effort: run this rank (and first-touch its pinned buffers) on the NUMA node its GPU hangs off.
    With N ranks streaming particles over PCIe at once, buffers that all live on one socket share that
    socket's links and the inter-socket fabric.  Returns a short description for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:          # 00000000:1b:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "numa node %d (%d cpus)" % (node, len(cpus))
        return "numa node %d (no allowed cpu)" % node
    except Exception as exc:       # containers without /sys, no NVML, ...
        return "not bound (%s)" % type(exc).__name__


def run_ours(args):
    numa = bind_to_gpu_numa_node(int(os.environ.get("LOCAL_RANK", "0")))
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

    def timed_once():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        one_step()
        b.record()
        barrier()
        return a.elapsed_time(b)

    # cold numbers (reported, not part of `value`): the very first iteration of the engine (no window
    # hint: host-driven median, first-use allocations) ...
    cold_first_ms = timed_once()
    for _ in range(max(args.warmup - 1, 0)):
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

    def region(k):
        ms, cnt = ctypes.c_double(), ctypes.c_int64()
        ctx.check(ctx.lib.stein_ctx_profile_read(ctx.handle, k, ctypes.byref(ms), ctypes.byref(cnt)))
        return ms, cnt

    phi_ms, phi_n = region(0)
    sw_ms, sw_n = region(1)
    # per-phase timeline: a second pass of the same K steps with every region timed (28 more event records per
    # iteration -- kept out of the pass that `value` comes from)
    ctx.check(ctx.lib.stein_ctx_profile_enable(ctx.handle, 2))
    barrier()
    tl = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in tl:
        flush.zero_()
        a.record()
        one_step()
        b.record()
    barrier()
    tl_ms_per_step = sum(a.elapsed_time(b) for a, b in tl) / args.steps
    tl_phi, tl_sweep = region(0)[0].value / args.steps, region(1)[0].value / args.steps
    regions = {name: region(k)[0].value / args.steps for k, name in
               [(2, "median"), (3, "phi_prep"), (4, "phi_tail"), (5, "step_push"), (6, "collectives"), (7, "head")]}
    ctx.check(ctx.lib.stein_ctx_profile_enable(ctx.handle, 0))
    route = ctx.phi_route()
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
    for _ in range(max(args.warmup, 2)):
        eng.update_particles_host(S_np, X_np)
    barrier()
    pf0 = eng.prefetch_stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        X_np[:1, :1] = np.nan                       # (the call must deliver this step's particles)
        eng.update_particles_host(S_np, X_np)       # returns after the D2H copy completed
        assert X_np[0, 0] == X_np[0, 0]
    barrier()
    e2e_s = time.perf_counter() - t0
    pf1 = eng.prefetch_stats()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    # where the end-to-end time goes: the particle download alone (PCIe), same buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        eng.get_particles(np.float32, out=X_np)
    d2h_ms = (time.perf_counter() - t0) / 5 * 1e3
    # ... and a step whose median window hint misses (the particles jumped: the speculative sweep
    # around the old window is wasted and the host-driven route runs, 2 sweeps)
    eng.set_particles(eng.get_particles(np.float32) * np.float32(1.25))      # collective on a sharded engine
    hint_miss_ms = timed_once()
    hint_miss_sweeps = eng.last()["sweeps"]
    eng.set_particles(eng.get_particles(np.float32) * np.float32(0.8))
    one_step()

    config_e = run_config_e(args, ctx, world, rank, dev, barrier) if args.config_e_steps > 0 else None
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
    ldp = (d + 31) // 32 * 32
    if code == 0:       # what AUTO resolves to (stein_b200/csrc/ctx.cu pick_phi_impl + the conditioning guard)
        code = {"fast": 5, "precise": 6, "ffma": 1}.get(route["route"], 5) if ldp == 256 else (2 if ldp == 128 else 1)
    impl = {1: "dense_simt_fp32", 2: "flash_tcgen05", 3: "flash_tcgen05_cta_pair",
            4: "flash_tcgen05_cta_pair_fp8_gemm2", 5: "flash_tcgen05_cta_pair_fp16_fp8",
            6: "flash_tcgen05_cta_pair_fp16x3"}[code]
    executed = {1: "FP32 FFMA", 2: "3 BF16 passes per GEMM", 3: "3 BF16 passes per GEMM",
                4: "GEMM1 3 BF16 passes; GEMM2 1 FP16 + 2 FP8 passes",
                5: "1 FP16 + 2 FP8 passes per GEMM (= 2 BF16-pass equivalents of tensor time each)",
                6: "3 FP16 passes per GEMM"}[code]
    # tensor work actually issued, in BF16-pass equivalents: 2 GEMMs of n_local x n x d, passes as above
    passes = {1: 0.0, 2: 3.0, 3: 3.0, 4: 2.5, 5: 2.0, 6: 3.0}[code]
    executed_tflops = passes * 2.0 * 2.0 * n_local * n * (ldp if code >= 2 else d) / (phi_ms.value / max(phi_n.value, 1) * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic("flash_phi2_kernel") if (code >= 3 and world == 1 and (n, d) == (N_PARTICLES, DIM)) \
        else (None, None)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n, d, world),
        "details": {"collectives": ("none" if world == 1 else
                                   "scores: NCCL all-gather on a side stream; particles: pushed to the peers by "
                                   "the optimizer kernel; small all-reduces: %s"
                                   % ("one kernel each over NVLink peer memory"
                                      if os.environ.get("STEIN_PEER_REDUCE", "1") != "0" else "NCCL")
                                   if peer_push else "NCCL (library-driven)"),
                    "phi_impl": impl, "phi_route": route, "median_sweeps_last_step": info["sweeps"],
                    "l2": "256 MiB buffer written between timed steps (L2 flush); per-step working set "
                          "is also > 126 MB"},
        "interactions_per_sec": value * n * n,
        "gpu_launches": int(launches),
        "wall_s_timed_region": wall,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "phi (%s)" % impl, "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak,
                     "frac_burst": achieved / peaks["bf16_tflops"], "peak_burst": peaks["bf16_tflops"],
                     "traffic": traffic, "traffic_source": traffic_src,
                     "traffic_kind": "static (one ncu --set full capture of this kernel at this shape, committed "
                                     "under profiles/; not re-measured by this run)" if traffic is not None else None,
                     "peak_source": peaks["source"] + ": `peak` = bf16 dense sustained (seconds-long loop under the "
                                    "power cap), `peak_burst` = best of 10 in isolation; the timed region here is "
                                    "short, so frac_burst is the conservative reading",
                     "executed_arithmetic": executed, "executed_bf16_equiv_tflops": executed_tflops,
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
        # device time per iteration by phase (CUDA events on the ctx stream, rank 0; a SEPARATE pass of K steps with
        # all regions timed, `step_total_in_this_pass` per iteration): head = barrier /
        # all-gather of the particles + row norms, error budgets, scale and FP16 split of the median (one
        # read of the particles); median = the rest of the median call including its host round
        # trip (the sweep is part of it); phi_prep = centring, guard, operand arrays; phi = main kernel;
        # phi_tail = finalize + sum(phi^2); step_push = clip + optimizer (+ peer push); collectives = the
        # all-reduce kernels (already contained in median / phi_tail / head); idle = the rest of the step
        "phases_ms": dict(regions, phi=tl_phi, sweep=tl_sweep, step_total_in_this_pass=tl_ms_per_step,
                          idle=tl_ms_per_step - (regions["head"] + regions["median"] + regions["phi_prep"] +
                                                 tl_phi + regions["phi_tail"] + regions["step_push"])),
        "cold": {"first_iteration_ms": cold_first_ms, "hint_miss_step_ms": hint_miss_ms,
                 "hint_miss_sweeps": hint_miss_sweeps,
                 "note": "not part of `value`: the timed steps ride the previous iteration's median window "
                         "(1 sweep); a step after a jump of the particles pays 2 sweeps"},
    }
    line["e2e"]["d2h_ms_alone"] = d2h_ms
    line["e2e"]["host_placement"] = numa
    line["e2e"]["prefetched_medians"] = pf1["used"] - pf0["used"]
    line["e2e"]["note"] = ("synchronous contract (the caller holds the new particles when the call returns). The "
                           "%.1f MB download crosses PCIe on a copy stream while the ctx stream already runs the "
                           "NEXT iteration's row norms, operand preparation and median (they need only the "
                           "particles, not the caller's next scores); the score upload of the next call overlaps "
                           "that median too. `prefetched_medians` of the timed calls collected a median enqueued "
                           "by the previous call." % (X_np.nbytes / 1e6))
    if config_e is not None:
        line["config_e"] = config_e
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(n, d, target_seconds=args.ref_seconds)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_config_e(args, ctx, world, rank, dev, barrier):
    """BASELINE.json configs[4]: Gaussian-mixture target (means +-2 e1, +-2 e2, sigma = 1), n = 262 144 particles,
    d = 1 024, particle rows sharded over the ranks, global exact median bandwidth.  A sub-record of the
    bench line (the headline `value` stays on config D).  Never raises: an error is reported in the record."""
    import ctypes
    import torch
    rec = {"workload": "gaussian_mixture_n%d_d%d" % (args.config_e_n, args.config_e_d), "n_particles": args.config_e_n,
           "dim": args.config_e_d, "n_gpus": world, "steps": args.config_e_steps}
    try:
        from stein_b200.engine import SvgdEngine
        from stein_b200.log_p import GaussianMixtureTarget
        n, d = args.config_e_n, args.config_e_d
        eng = SvgdEngine(n, d, "adam", learning_rate=1e-2, ctx=ctx)
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + rank)
        Xl = eng.particles_dev
        Xl[:eng.n_local, :d] = torch.randn((eng.n_local, d), generator=gen, device=dev, dtype=torch.float32)
        means = np.zeros((4, d), np.float32)
        means[0, 0], means[1, 0], means[2, 1], means[3, 1] = 2, -2, 2, -2
        model = GaussianMixtureTarget(d, means=means)

        def step():
            model.scores(eng)
            eng.step()

        step()                                  # cold: host-driven median, allocations
        barrier()
        ctx.check(ctx.lib.stein_ctx_profile_enable(ctx.handle, 2))
        for k in range(8):
            ctx.lib.stein_ctx_profile_read(ctx.handle, k, None, None)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for _ in range(args.config_e_steps):
            step()
        b.record()
        barrier()
        ms = a.elapsed_time(b) / args.config_e_steps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

        def region(k):
            v, c = ctypes.c_double(), ctypes.c_int64()
            ctx.check(ctx.lib.stein_ctx_profile_read(ctx.handle, k, ctypes.byref(v), ctypes.byref(c)))
            return v.value / args.config_e_steps
        phi_ms, sweep_ms, median_ms = region(0), region(1), region(2)
        ctx.check(ctx.lib.stein_ctx_profile_enable(ctx.handle, 0))
        info, route = eng.last(), ctx.phi_route()
        n_local = eng.n_local
        eng.close()
        peaks = load_peaks()
        f_phi = algorithmic_flops_phi(n, d) * (n_local / float(n))
        ach = f_phi / (phi_ms * 1e-3) / 1e12
        rec.update({
            "ms_per_step": ms, "value": 1000.0 / ms, "unit": UNIT, "interactions_per_sec": 1000.0 / ms * n * n,
            "bandwidth_last_step": info["bandwidth"], "median_sweeps_last_step": info["sweeps"], "phi_route": route,
            "phases_ms": {"phi_panel_kernels": phi_ms, "median": median_ms, "sweep": sweep_ms},
            "roofline": {"bound": "tensor", "kernel": "phi panel kernels (exp-GEMM + P.Y GEMM, FP16 + 2 FP8 passes each)",
                         "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": ach / peaks["bf16_tflops_sustained"], "frac_burst": ach / peaks["bf16_tflops"],
                         "algorithmic_flops_per_launch_group": f_phi, "traffic": None},
            "data": "synthetic: X ~ N(0, I) per rank, scores by stein_score_gaussian_mixture on the device",
        })
    except Exception as exc:          # the headline record must survive a failure here
        rec["error"] = "%s: %s" % (type(exc).__name__, exc)
    return rec


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
    ap.add_argument("--config-e-steps", type=int, default=2,
                    help="timed iterations of the config-E sub-record (n = 262 144, d = 1 024); 0 = skip")
    ap.add_argument("--config-e-n", type=int, default=262144)
    ap.add_argument("--config-e-d", type=int, default=1024)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
