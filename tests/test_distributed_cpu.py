"""World-size-2 `gloo` tests (CPU) of the N > 1 host logic: the row / tile shard plans,
the radix-select narrowing driven by ALL-REDUCED histograms (the product's own
stein_median_narrow, a pure host function of libstein_b200.so), the all-gathered particle
layout and the all-reduced Frobenius norm.  The arithmetic inside each rank is done by the
oracle (this is a test: the product has no CPU compute path); what is checked is that the
sharded protocol reproduces the unsharded result bit for bit (median) / to rounding (phi).
"""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import svgd_oracle as orc
from stein_b200 import _lib
from stein_b200.distributed import shard_rows, shard_tiles

TILE = 128


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _tile_hist(lib, D, n, t0, t1, key_lo, shift, nbins):
    """What stein_sqdist_hist accumulates for tiles [t0, t1): upper-triangular 128x128
    tiles, off-diagonal tiles weigh 2."""
    counts = np.zeros(nbins + 1, np.int64)
    I, J = ctypes.c_int32(), ctypes.c_int32()
    for t in range(t0, t1):
        assert lib.stein_tile_coords(t, n, ctypes.byref(I), ctypes.byref(J)) == 0
        blk = D[I.value * TILE:(I.value + 1) * TILE, J.value * TILE:(J.value + 1) * TILE].reshape(-1)
        w = 1 if I.value == J.value else 2
        bits = (blk + np.float32(0)).view(np.uint32).astype(np.uint64)
        keys = np.where(bits & 0x80000000, ~bits & 0xFFFFFFFF, bits | 0x80000000).astype(np.uint64)
        counts[0] += w * int(np.sum(keys < key_lo))
        inw = keys[keys >= key_lo]
        b = (inw - np.uint64(key_lo)) >> np.uint64(shift)
        np.add.at(counts, 1 + b[b < nbins].astype(np.int64), w)
    return counts


def _worker(rank, world, port, n, d, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _lib.load()
    rng = np.random.default_rng(123)
    X = rng.standard_normal((n, d)).astype(np.float32)
    S = rng.standard_normal((n, d)).astype(np.float32)

    # --- all-gather of the row shards reproduces the full (padded) particle matrix
    begin, n_local, q = shard_rows(n, world, rank)
    ld = lib.stein_ld(d)
    local = torch.zeros((q, ld), dtype=torch.float32)
    local[:n_local, :d] = torch.from_numpy(X[begin:begin + n_local])
    full = torch.zeros((q * world, ld), dtype=torch.float32)
    dist.all_gather_into_tensor(full, local)
    assert np.array_equal(full.numpy()[:n, :d], X) and not full.numpy()[n:].any()

    # --- sharded exact median: each rank histograms its tile range, histograms are all-reduced,
    #     every rank narrows identically with the product's stein_median_narrow
    D = orc.sqdist_chain(X)
    t0, t1 = shard_tiles(n, world, rank)
    dim = n * n
    ranks = [dim // 2 - 1, dim // 2] if dim % 2 == 0 else [dim // 2, dim // 2]
    found = []
    for target in ranks:
        key_lo, shift, nbins = 0, 18, 16384
        for _ in range(4):
            counts = torch.from_numpy(_tile_hist(lib, D, n, t0, t1, key_lo, shift, nbins))
            dist.all_reduce(counts)
            c = counts.numpy().astype(np.uint64)
            ko, lo, sh, nb = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
            rc = lib.stein_median_narrow(c.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), key_lo, shift, nbins,
                                         target, ctypes.byref(ko), ctypes.byref(lo), ctypes.byref(sh), ctypes.byref(nb))
            assert rc in (0, 1)
            if rc == 1:
                found.append(lib.stein_key_to_float(ko.value))
                break
            key_lo, shift, nbins = lo.value, sh.value, nb.value
    lo_v, hi_v = np.float32(found[0]), np.float32(found[1])
    med = np.float32((lo_v + hi_v) / np.float32(2)) if dim % 2 == 0 else lo_v
    med_ref, mid_ref = orc.median_chain(X)
    assert med.tobytes() == med_ref.tobytes() and (lo_v, hi_v) == mid_ref
    bw = np.float32(lib.stein_bandwidth(ctypes.c_float(float(med)), n))
    assert bw.tobytes() == orc.bandwidth(med_ref, n).tobytes()

    # --- sharded phi + all-reduced Frobenius norm + clip: equals the unsharded update
    phi_local, _ = orc.phi_rows_c(X, S, bw, begin, begin + n_local)
    ss = torch.tensor([float((phi_local ** 2).sum())], dtype=torch.float64)
    dist.all_reduce(ss)
    phi_full = orc.compute_phi(X, S)
    assert abs(np.sqrt(ss.item()) - np.linalg.norm(phi_full)) <= 1e-6 * np.linalg.norm(phi_full)
    scale = 10.0 / max(10.0, np.sqrt(ss.item()))
    ref = orc.clip(phi_full)[begin:begin + n_local]
    assert np.abs(phi_local * scale - ref).max() <= 2e-6 * np.abs(phi_full).max()
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.array([1]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,d", [(300, 7), (257, 33)])
def test_two_rank_protocol_matches_single_rank(tmp_path, n, d):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, d, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok_%d.npy" % r)) for r in range(world))
