"""Fine-grained timeline of one SVGD iteration at the bench configuration (n = 65 536, d = 256) on 1..8 GPUs:
labelled CUDA events after each stage of the step (stein_ctx_trace_enable / _read), averaged over K steps, printed
for every rank.  `python tools/step_trace.py` or under torchrun.  Timing aid, not a benchmark (the event records
themselves cost a few microseconds each)."""
import collections
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from stein_b200.engine import SvgdEngine
    from stein_b200.log_p import GaussianMixtureTarget
    from stein_b200.runtime import context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps = int(os.environ.get("TRACE_STEPS", "10"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = context(local_rank)
    if world > 1:
        from stein_b200.distributed import make_comm
        make_comm(ctx)
    n, d = 65536, 256
    eng = SvgdEngine(n, d, "adam", learning_rate=1e-2, ctx=ctx)
    X0 = np.random.default_rng(1).standard_normal((n, d)).astype(np.float32)
    eng.set_particles(np.ascontiguousarray(X0[eng.row_begin:eng.row_begin + eng.n_local]))
    model = GaussianMixtureTarget(d)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(5):
        model.scores(eng)
        eng.step()
    barrier()
    ctx.check(ctx.lib.stein_ctx_trace_enable(ctx.handle, 1))
    evs = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        model.scores(eng)
        eng.step()
        b.record()
        evs.append((a, b))
    barrier()
    buf = ctypes.create_string_buffer(1 << 20)
    ctx.check(ctx.lib.stein_ctx_trace_read(ctx.handle, buf, len(buf)))
    ctx.check(ctx.lib.stein_ctx_trace_enable(ctx.handle, 0))
    step_ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    # marks in enqueue order; "step:begin" starts an iteration (its own interval is the gap since the last step)
    order, acc = [], collections.OrderedDict()
    k = 0
    for line in buf.value.decode().splitlines():
        label, ms = line.split("\t")
        if label == "step:begin":
            k = 0
            continue
        key = "%02d %s" % (k, label)
        acc.setdefault(key, []).append(float(ms))
        k += 1
    rows = [(key, sum(v) / len(v)) for key, v in acc.items()]
    out = {"rank": rank, "world": world, "step_ms": step_ms, "sum_marks_ms": sum(ms for _, ms in rows),
           "marks_ms": rows}
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, out)
    else:
        gathered = [out]
    if rank == 0:
        print("step (events around scores + step, mean of %d): " % steps
              + " ".join("%.3f" % g["step_ms"] for g in gathered))
        keys = [k for k, _ in gathered[0]["marks_ms"]]
        for i, key in enumerate(keys):
            vals = [g["marks_ms"][i][1] if i < len(g["marks_ms"]) else float("nan") for g in gathered]
            print("%-44s max %.4f  mean %.4f  | %s" % (key, max(vals), sum(vals) / len(vals),
                                                      " ".join("%.3f" % v for v in vals)))
        print(json.dumps({"world": world, "step_ms": [g["step_ms"] for g in gathered],
                          "marks_mean_over_ranks": [[k, float(np.mean([g["marks_ms"][i][1] for g in gathered]))]
                                                    for i, k in enumerate(keys)]}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
