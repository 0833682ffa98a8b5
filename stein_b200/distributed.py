"""Collective hooks (include/stein_b200.h `stein_comm`) backed by
torch.distributed -- NCCL over NVLink/NVSwitch on a B200 node.

The engine is particle-row sharded: per iteration one
all-gather of the particle and score shards, one u64 all-reduce per median
sweep, one f64 all-reduce of sum(phi^2).  No other cross-GPU traffic.
"""
import ctypes

from . import _lib


class _DevMem:
    """View of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {
            "shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def make_nccl_comm(ctx, group=None):
    """Collectives driven by the library itself: an NCCL communicator created from a unique
    id that rank 0 generates and torch.distributed broadcasts once.  Afterwards no Python runs
    inside an iteration (a Python-backed hook costs more than the collective at these sizes)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [None, None]      # ids of the main and of the side communicator
    if rank == 0:
        for k in range(2):
            buf = ctypes.create_string_buffer(128)
            ctx.check(ctx.lib.stein_nccl_unique_id(buf))
            box[k] = buf.raw
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.check(ctx.lib.stein_ctx_init_nccl(ctx.handle, rank, world, ctypes.c_char_p(box[0]),
                                          ctypes.c_char_p(box[1])))
    ctx._native_comm_group = (group, rank, world)      # engines exchange their IPC handles over it
    return rank, world


def connect_peers(engine):
    """Exchange the CUDA-IPC handles of the engines' particle buffers (all ranks, one node) so that
    the optimizer kernel pushes updated rows straight into the peers' buffers.  Collective.
    Returns True when every rank could open every handle; otherwise the engines keep using the
    all-gather hook."""
    import torch.distributed as dist
    ctx = engine.ctx
    info = getattr(ctx, "_native_comm_group", None)
    if info is None:
        return False
    group, rank, world = info
    buf = ctypes.create_string_buffer(64)
    ctx.check(ctx.lib.stein_engine_ipc_handle(engine.handle, buf))
    handles = [None] * world
    dist.all_gather_object(handles, buf.raw, group=group)
    rc = ctx.lib.stein_engine_set_peer_handles(engine.handle, ctypes.c_char_p(b"".join(handles)))
    # all or nothing: a rank that could not open its peers must not skip the all-gather alone
    ok = [None] * world
    dist.all_gather_object(ok, rc == 0, group=group)
    if not all(ok):
        ctx.check(ctx.lib.stein_engine_set_peer_handles(engine.handle, None))
        return False
    return True


def quiesce_peers(engine):
    """Before an engine with connected peers is freed: drain this rank's stream, then wait for
    every rank (their kernels write into this engine's particle buffer and mailbox).  Best
    effort at interpreter shutdown, when the process group may already be gone."""
    import torch
    import torch.distributed as dist
    info = getattr(engine.ctx, "_native_comm_group", None)
    try:
        torch.cuda.synchronize(engine.ctx.device)
        if info is not None and dist.is_initialized():
            dist.barrier(group=info[0])
    except Exception:
        pass


def make_comm(ctx, group=None, native=None):
    """Install collective hooks on `ctx` for the default (or given) process group.
    native=True (default on an NCCL group): the library's own NCCL transport
    (`make_nccl_comm`); native=False: hooks backed by torch.distributed calls (any backend
    that works on device tensors).  Returns what was installed."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    if native is None:
        native = dist.get_backend(group) == "nccl"
    if native:
        return make_nccl_comm(ctx, group)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = "cuda:%d" % ctx.device

    def _tensor(ptr, count, typestr):
        return torch.as_tensor(_DevMem(ptr, count, typestr), device=dev)

    def allreduce_u64(_user, buf, count):
        try:
            # counts stay far below 2^63, so the int64 sum has the same bits
            dist.all_reduce(_tensor(buf, count, "<i8"), group=group)
            return 0
        except Exception:       # surfaced by the C side as STEIN_ERR_COMM
            import traceback
            traceback.print_exc()
            return 1

    def allreduce_f64(_user, buf, count):
        try:
            dist.all_reduce(_tensor(buf, count, "<f8"), group=group)
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return 1

    def allgather_f32(_user, send, recv, count):
        try:
            out = _tensor(recv, count * world, "<f4")
            # `send` is the rank's own slice of `recv` (in-place all-gather)
            inp = out[rank * count:(rank + 1) * count] if send == recv + 4 * rank * count \
                else _tensor(send, count, "<f4")
            dist.all_gather_into_tensor(out, inp, group=group)
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return 1

    comm = _lib.SteinComm(rank, world, None, _lib.HOOK(allreduce_u64), _lib.HOOK(allreduce_f64),
                          _lib.HOOK_GATHER(allgather_f32), _lib.HOOK_GATHER_ON())
    ctx.check(ctx.lib.stein_ctx_set_comm(ctx.handle, ctypes.byref(comm)))
    ctx._comm_keepalive = comm      # the C side copied the struct; keep the callbacks alive
    return comm


def shard_rows(n_total, world, rank, tile=128):
    """Row range owned by `rank`: q = tile-padded ceil(n/world) rows per rank,
    global rows [rank*q, min(n, (rank+1)*q)).  Mirrors stein_engine_create."""
    per_rank = -(-n_total // world)              # ceil(n / world)
    q = -(-per_rank // tile) * tile              # padded to whole tiles
    begin = rank * q
    n_local = max(0, min(q, n_total - begin))
    return begin, n_local, q


def shard_tiles(n_total, world, rank, tile=128):
    """Upper-triangular distance-tile range swept by `rank` in the median
    (mirrors stein_median_sqdist): contiguous, equal-cost chunks."""
    T = -(-n_total // tile)
    ntiles = T * (T + 1) // 2
    return ntiles * rank // world, ntiles * (rank + 1) // world
