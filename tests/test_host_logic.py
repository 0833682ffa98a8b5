"""CPU tests: the C-ABI library loads and exports every declared symbol, and the
host-side logic (layout converters, key order, tile enumeration, radix-select
narrowing, sharding plans) is correct.  No CUDA compute is called here."""
import ctypes
import os
import re
import types

import numpy as np
import pytest

from oracle import svgd_oracle as orc
from stein_b200 import _lib
from stein_b200.distributed import shard_rows, shard_tiles
from stein_b200.utilities.converters import convert_array_to_dictionary, convert_dictionary_to_array

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "stein_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(stein_[a-z0-9_]+)\s*\(", header))
    declared -= {"stein_comm"}
    assert len(declared) >= 40
    for name in sorted(declared):
        assert hasattr(lib, name), "libstein_b200.so lacks %s" % name
    # and the Python binding table covers the header one to one
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.stein_ctx_create(ctypes.byref(h), 0, None)
    assert rc == -2 and b"no CPU fallback" in lib.stein_last_error(None)
    from stein_b200.runtime import context
    with pytest.raises(_lib.SteinLibraryError):
        context()


def test_layout_helpers(lib):
    assert [lib.stein_ld(d) for d in (1, 10, 32, 33, 55, 256, 753)] == [32, 32, 32, 64, 64, 256, 768]
    assert [lib.stein_rows_padded(n) for n in (1, 100, 128, 129, 65536)] == [128, 128, 128, 256, 65536]


def test_key_transform_is_monotone_and_invertible(lib):
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.standard_normal(2000).astype(np.float32) * 10.0 ** rng.integers(-20, 20, 2000),
                        np.array([0.0, -0.0, 1e-45, -1e-45, 3.4e38, -3.4e38], np.float32)]).astype(np.float32)
    keys = np.array([lib.stein_float_to_key(ctypes.c_float(float(x))) for x in v], dtype=np.uint64)
    order = np.argsort(v, kind="stable")
    assert np.all(np.diff(keys[order].astype(np.int64)) >= 0)
    back = np.array([lib.stein_key_to_float(ctypes.c_uint32(int(k))) for k in keys], np.float32)
    np.testing.assert_array_equal(back, v + np.float32(0.0))
    assert lib.stein_float_to_key(ctypes.c_float(-0.0)) == lib.stein_float_to_key(ctypes.c_float(0.0))


def test_bandwidth_matches_oracle(lib):
    for med, n in [(512.0, 65536), (0.73801529, 50), (18.559925, 100), (1e-6, 2), (3.0, 262144)]:
        got = lib.stein_bandwidth(ctypes.c_float(med), n)
        assert np.float32(got) == orc.bandwidth(np.float32(med), n)


@pytest.mark.parametrize("n", [1, 100, 128, 129, 1000, 5000, 65536, 262144])
def test_tile_enumeration(lib, n):
    T = -(-n // 128)
    nt = lib.stein_num_tiles(n)
    assert nt == T * (T + 1) // 2
    I, J = ctypes.c_int32(), ctypes.c_int32()
    ts = list(range(min(nt, 300))) + list(range(max(0, nt - 300), nt)) + \
        list(np.random.default_rng(n).integers(0, nt, 300))
    for t in ts:
        assert lib.stein_tile_coords(int(t), n, ctypes.byref(I), ctypes.byref(J)) == 0
        i, j = I.value, J.value
        assert 0 <= i <= j < T
        assert i * T - i * (i - 1) // 2 + (j - i) == t
    assert lib.stein_tile_coords(nt, n, ctypes.byref(I), ctypes.byref(J)) != 0


def _hist(keys, key_lo, shift, nbins):
    counts = np.zeros(nbins + 1, np.uint64)
    counts[0] = np.sum(keys < key_lo)
    inw = keys[keys >= key_lo]
    b = (inw - key_lo) >> np.uint64(shift)
    b = b[b < nbins]
    np.add.at(counts, 1 + b.astype(np.int64), 1)
    return counts


def _select(lib, keys, rank, window):
    """Drive stein_median_narrow exactly like stein_median_sqdist does."""
    key_lo, shift, nbins = window
    full = (0, 18, 16384)
    sweeps = 0
    for _ in range(10):
        counts = _hist(keys, key_lo, shift, nbins)
        sweeps += 1
        ko, lo, sh, nb = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        rc = lib.stein_median_narrow(counts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), key_lo, shift,
                                     nbins, rank, ctypes.byref(ko), ctypes.byref(lo), ctypes.byref(sh),
                                     ctypes.byref(nb))
        if rc == 1:
            return ko.value, sweeps
        if rc == 0:
            key_lo, shift, nbins = lo.value, sh.value, nb.value
        else:
            assert (key_lo, shift, nbins) != full
            key_lo, shift, nbins = full
    raise AssertionError("did not converge")


def test_narrowing_selects_exact_rank(lib):
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.standard_normal(20000).astype(np.float32) * 3 + 500,
                           np.zeros(300, np.float32), -rng.random(50).astype(np.float32) * 1e-6])
    keys = np.array([lib.stein_float_to_key(ctypes.c_float(float(x))) for x in vals], dtype=np.uint64)
    skeys = np.sort(keys)
    for rank in [0, 1, 299, 350, len(keys) // 2 - 1, len(keys) // 2, len(keys) - 1]:
        got, sweeps = _select(lib, keys, rank, (0, 18, 16384))
        assert got == skeys[rank] and sweeps == 3
    # a pilot-style window around the median: one or two sweeps
    lo, hi = int(skeys[len(keys) // 2 - 400]), int(skeys[len(keys) // 2 + 400])
    span = hi - lo + 1
    sh = 0
    while ((span - 1) >> sh) >= 16384:
        sh += 1
    got, sweeps = _select(lib, keys, len(keys) // 2, (lo, sh, ((span - 1) >> sh) + 1))
    assert got == skeys[len(keys) // 2] and sweeps <= 2
    # a window that misses the rank falls back to the full range and still succeeds
    got, sweeps = _select(lib, keys, 10, (lo, sh, ((span - 1) >> sh) + 1))
    assert got == skeys[10]
    got, sweeps = _select(lib, keys, len(keys) - 5, (lo, sh, ((span - 1) >> sh) + 1))
    assert got == skeys[len(keys) - 5]


class FakeVar:
    def __init__(self, name, shape):
        self.name, self._shape = name, list(shape)

    def get_shape(self):
        return types.SimpleNamespace(as_list=lambda: list(self._shape))


def test_product_converters_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "converters.npz"))
    shapes = [[int(x) for x in s.split(",")] if s else [] for s in g["shapes"]]
    vs = [FakeVar(str(nm), sh) for nm, sh in zip(g["names"], shapes)]
    dictionary = {v: g["value_%d" % i] for i, v in enumerate(vs)}
    array, access = convert_dictionary_to_array(dictionary)
    np.testing.assert_array_equal(array, g["array"])
    assert array.dtype == np.float64
    assert [access[v] for v in vs] == list(zip(g["starts"], g["stops"]))
    back = convert_array_to_dictionary(array, access)
    for v in vs:
        np.testing.assert_array_equal(back[v], dictionary[v])
        assert back[v].shape == dictionary[v].shape


def test_model_layouts_follow_the_name_sort():
    from stein_b200.log_p import LinearRegression, LogisticRegression, RegressionNeuralNetwork
    m = LogisticRegression(54, 464809)
    assert m.n_params == 55 and m.column_slices()[m.w] == (0, 54) and m.column_slices()[m.log_alpha] == (54, 55)
    b = RegressionNeuralNetwork(13, 50, 506)
    sl = b.column_slices()
    assert b.n_params == 753
    assert [sl[v] for v in (b.log_lambda, b.log_gamma, b.w_1, b.b_1, b.w_2, b.b_2)] == \
        [(0, 1), (1, 2), (2, 652), (652, 702), (702, 752), (752, 753)]
    assert RegressionNeuralNetwork(90, 50, 515345).n_params == 4603
    assert LinearRegression(10).n_params == 10
    assert [v.name for v in b.model_vars][:2] == ["model/Variable:0", "model/Variable_1:0"]


@pytest.mark.parametrize("n,world", [(65536, 8), (65536, 1), (1000, 2), (100, 4), (262144, 8), (129, 2)])
def test_shard_plans_cover_everything_once(n, world):
    rows = [shard_rows(n, world, r) for r in range(world)]
    q = rows[0][2]
    assert q % 128 == 0 and q * world >= n
    covered = []
    for r, (b, nl, qq) in enumerate(rows):
        assert qq == q and b == r * q
        covered += list(range(b, b + nl))
    assert covered == list(range(n))
    tiles = [shard_tiles(n, world, r) for r in range(world)]
    T = -(-n // 128)
    assert tiles[0][0] == 0 and tiles[-1][1] == T * (T + 1) // 2
    assert all(tiles[r][1] == tiles[r + 1][0] for r in range(world - 1))
    sizes = [b - a for a, b in tiles]
    assert max(sizes) - min(sizes) <= 1


# --------------------------------------------------------------------------- #
# work schedules of the tensor-core kernels (pure host functions of the .so)   #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("nI,nJ,G", [(256, 512, 74), (512, 512, 148), (128, 512, 74), (64, 512, 74), (32, 512, 74),
                                     (1, 1, 74), (1, 40, 74), (75, 3, 74), (3, 1000, 148), (147, 17, 148),
                                     (149, 17, 148), (10, 5, 1)])
def test_phi_tile_schedule_covers_every_tile_pair_once(lib, nI, nJ, G):
    """TileSchedule / SegWalk (csrc/phi_tc.cu): over all units, every (row tile, column tile) pair
    is visited exactly once; the partial results of a row tile land in distinct slots
    0..nslots-1 with nslots <= 8; and the units' loads differ by at most one chunk."""
    fn = lib.stein_debug_tile_schedule
    fn.restype = ctypes.c_int
    IntP = ctypes.POINTER(ctypes.c_int)
    fn.argtypes = [ctypes.c_int] * 5 + [IntP] * 5
    cover = np.zeros((nI, nJ), np.int32)
    slots = [set() for _ in range(nI)]
    nslots = (ctypes.c_int * nI)()
    loads = []
    cap = nI // G + 4
    for c in range(G):
        t, j0, j1, sl = ((ctypes.c_int * cap)() for _ in range(4))
        n = fn(nI, nJ, G, c, cap, t, j0, j1, sl, nslots)
        assert n <= cap
        load = 0
        for k in range(n):
            assert 0 <= t[k] < nI and 0 <= j0[k] < j1[k] <= nJ
            cover[t[k], j0[k]:j1[k]] += 1
            assert sl[k] not in slots[t[k]]
            slots[t[k]].add(sl[k])
            load += j1[k] - j0[k]
        loads.append(load)
    assert (cover == 1).all()
    for i in range(nI):
        assert slots[i] == set(range(nslots[i])) and 1 <= nslots[i] <= 8
    rem = nI % G
    chunk = max(-(-rem * nJ // G), -(-nJ // 7), 1)
    assert max(loads) - min(l for l in loads if l > 0 or rem == 0) <= chunk


@pytest.mark.parametrize("nI,nJ,G", [(256, 512, 74), (128, 512, 74), (64, 512, 74), (32, 512, 74), (16, 512, 74)])
def test_phi_leftover_chunks_walk_the_columns_together(lib, nI, nJ, G):
    """The leftover row tiles (all of them on a sharded run) are cut into one chunk per unit.  A
    unit walks the columns of its chunk in ASCENDING order, so at every step all units are within
    nJ - chunk columns of each other and a column tile is served from L2 after its first fetch."""
    fn = lib.stein_debug_tile_schedule
    fn.restype = ctypes.c_int
    IntP = ctypes.POINTER(ctypes.c_int)
    fn.argtypes = [ctypes.c_int] * 5 + [IntP] * 5
    rounds, rem = nI // G, nI % G
    chunk = max(-(-rem * nJ // G), -(-nJ // 7), 1)
    cap = rounds + 4
    walks = []
    for c in range(G):
        t, j0, j1, sl = ((ctypes.c_int * cap)() for _ in range(4))
        n = fn(nI, nJ, G, c, cap, t, j0, j1, sl, None)
        cols = [j for k in range(rounds, n) for j in range(j0[k], j1[k])]
        assert cols == sorted(cols) and len(set(cols)) == len(cols)
        if cols:
            walks.append(cols)
    assert all(len(w) == chunk for w in walks[:-1])        # only the last chunk may be shorter
    steps = max(len(w) for w in walks)
    for s in range(steps):
        at = [w[s] for w in walks if s < len(w)]
        assert max(at) - min(at) <= nJ - len(walks[-1]) + 1
        assert max(at[:-1] or [0]) - min(at[:-1] or [0]) <= nJ - chunk + 1


@pytest.mark.parametrize("T", [1, 2, 3, 4, 33, 40, 512])
def test_pair_sweep_tile_enumeration(lib, T):
    """The CTA-pair median sweep walks pair rows I2 and column tiles J >= 2 I2 in row-major order
    (csrc/median_tc.cu pair_tile): together with the weights J > I -> 2, J == I -> 1, J < I -> 0
    every unordered tile pair {I, J} of the T x T tile grid is weighted as in the full matrix."""
    num = lib.stein_debug_num_pair_tiles
    num.restype, num.argtypes = ctypes.c_longlong, [ctypes.c_longlong]
    pt = lib.stein_debug_pair_tile
    pt.restype, pt.argtypes = None, [ctypes.c_longlong, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_int)]
    nt = num(T)
    assert nt == sum(T - 2 * i2 for i2 in range((T + 1) // 2))
    weight = np.zeros((T + 1, T), np.int64)
    I2, J = ctypes.c_int(), ctypes.c_int()
    prev = (-1, -1)
    for t in range(nt):
        pt(t, T, ctypes.byref(I2), ctypes.byref(J))
        assert (I2.value, J.value) > prev and 2 * I2.value <= J.value < T
        prev = (I2.value, J.value)
        for rank in (0, 1):
            I = 2 * I2.value + rank
            weight[I, J.value] += 2 if J.value > I else (1 if J.value == I else 0)
    assert weight[T].sum() == 0                            # the phantom row tile of an odd T weighs nothing
    full = weight[:T]
    assert full.sum() == T * T                             # = all entries of the full (symmetric) matrix
    iu, il = np.triu_indices(T, 1), np.tril_indices(T, -1)
    assert (np.diag(full) == 1).all() and (full[iu] == 2).all() and (full[il] == 0).all()


@pytest.mark.parametrize("T,parts,beta", [(512, 74, 1.0), (512, 8 * 74, 1.0), (36, 2 * 74, 1.0), (513, 4 * 74, 2.5),
                                          (128, 74, 0.5)])
def test_sweep_chunks_are_a_cost_balanced_partition(lib, T, parts, beta):
    """The CTA-pair sweep deals its tiles to world x clusters chunks of equal cost, tiles + beta per pair row
    entered (csrc/median_tc.cu sweep_chunk_bounds): the chunks are a partition of the tile sequence in order,
    and no chunk costs more than the ideal share plus one row change and one tile."""
    f = lib.stein_debug_sweep_chunk_bounds
    f.restype, f.argtypes = None, [ctypes.c_longlong, ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_longlong)]
    out = (ctypes.c_longlong * (parts + 1))()
    f(T, parts, beta, out)
    b = np.array(list(out))
    rows = [T - 2 * i for i in range((T + 1) // 2)]
    nt = sum(rows)
    assert b[0] == 0 and b[-1] == nt and (np.diff(b) >= 0).all()
    starts = np.concatenate([[0], np.cumsum(rows)])          # first tile of every pair row
    def cost(a, e):
        if e <= a:
            return 0.0
        first, last = np.searchsorted(starts, a, "right") - 1, np.searchsorted(starts, e - 1, "right") - 1
        return (e - a) + beta * (last - first + 1)
    costs = np.array([cost(b[k], b[k + 1]) for k in range(parts)])
    ideal = (nt + beta * len(rows)) / parts
    assert costs.max() <= ideal + 2 * beta + 2, (costs.max(), ideal)
    # the even split of the tiles it replaces leaves its last chunk with far more row changes
    even = np.array([cost(nt * k // parts, nt * (k + 1) // parts) for k in range(parts)])
    assert costs.max() <= even.max() + 1e-9


def test_kernel_plugin_bandwidth_argument():
    """abstract_kernel.py:17 takes (n_particles, sess); the optional `bandwidth` selects the
    fixed-bandwidth kernel and is validated on the host."""
    from stein_b200.kernels import AbstractKernel, SquaredExponentialKernel
    k = SquaredExponentialKernel(16, None)
    assert k.fixed_bandwidth is None and k.bandwidth is None and k.n_particles == 16
    k = SquaredExponentialKernel(16, bandwidth=0.5)
    assert k.fixed_bandwidth == np.float32(0.5) and k.bandwidth == np.float32(0.5)
    for bad in (0.0, -1.0, float("nan"), float("inf")):
        with pytest.raises(ValueError):
            SquaredExponentialKernel(16, bandwidth=bad)
    with pytest.raises(NotImplementedError):
        AbstractKernel.kernel_and_grad(k, np.zeros((16, 2)))
