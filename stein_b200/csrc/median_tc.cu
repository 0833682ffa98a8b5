// median_tc.cu -- kernel (2), tensor-core route: the exact median of the n x n
// squared-distance matrix with ONE tcgen05 sweep over the upper-triangular tiles.
//
// Reference: stein/kernels/abstract_kernel.py:33-38, stein/utilities/compute_median.py:4-16.
//
// The answer must be the order statistic of the CONTRACT-arithmetic distances
// (fp32 fma chain, see median.cu / oracle/svgd_oracle.c) -- tensor cores cannot
// produce those bits.  They are used as a FILTER instead:
//   1. the sweep computes D~ = r_i + r_j - 2 g~ with g~ from a 3-pass FP16-split GEMM
//      (|D~ - D| <= eps_ij = e_i + e_j, a worst-case bound with a per-row error budget, see err_budget_kernel);
//      pairs that are certainly below the pilot window [wlo, whi] are counted, pairs
//      certainly above are dropped, the rest (~0.7 %) are appended to a list with D~;
//   2. a radix select over the listed D~ gives t~, the D~ value at the target rank;
//      the exact median lies within delta = max eps of it;
//   3. entries certainly below t~ - delta are counted, entries that may lie within
//      [t~ - delta, t~ + delta] (~0.07 %) get their distance recomputed in contract
//      arithmetic (FFMA chain, pair_chain.cuh) and the exact rank is selected among those keys.
// Every step checks that the target rank is bracketed; if not (pilot window missed,
// list overflow) the caller falls back to the all-FFMA route of median.cu.
#include <map>
#include <cuda_fp16.h>
#include <math.h>

#include <algorithm>
#include <functional>
#include <vector>

#include "pair_chain.cuh"
#include "panel_gemm.cuh"
#include "tc_common.cuh"

namespace stein {

using namespace tc;

constexpr int SW_THREADS = 384;
constexpr int SW_EPI_THREADS = 256;
constexpr int SW_STAGES = 12;
constexpr uint32_t SW_UNIT_BYTES = 128 * 128;
constexpr int SW_EPI_WARPS = SW_EPI_THREADS / 32;
constexpr int SW2_EPI_WG = 4;               // pair kernel: four classification warpgroups, one 32-column chunk per warp
constexpr int SW2_THREADS = 128 + 128 * SW2_EPI_WG;
constexpr int SW_WARP_CAP = 128;            // entries a classification warp of the 8-warp layout stages between flushes
// (the tail below is sized for 8 warps x 128 entries x 64 columns = 16 warps x 64 entries x 32 columns)
// shared-memory tail after the operand buffers (both sweep kernels): barriers at +0 (256 B),
// tensor-memory base at +256, per-warp staging counters at +288, then the arrays below
constexpr size_t SW_TAIL_COUNTS = 288;
constexpr size_t SW_TAIL_BUF = 384;                                             // PairEntry [warps][entries] (counts: up to 16 warps x 4 B from +288)
constexpr size_t SW_TAIL_COL = SW_TAIL_BUF + (size_t)SW_EPI_WARPS * SW_WARP_CAP * 12;   // float [warps][64]
constexpr size_t SW_TAIL_HIST = SW_TAIL_COL + (size_t)SW_EPI_WARPS * 64 * 4;    // u32 [SW_HIST_BINS]
constexpr int SW_HIST_BINS = 4096;          // histogram of the listed D~ (locates t~ without extra passes)
constexpr size_t SW_TAIL_BYTES = SW_TAIL_HIST + (size_t)SW_HIST_BINS * 4 + 64;
// 32-bit words of the device-picked band parameters (pick_band_kernel)
constexpr int BP_TLO = 0, BP_THI = 1, BP_DELTA = 2, BP_EPS_ABS = 3, BP_KLO = 4, BP_SHIFT = 5, BP_NBINS = 6,
              BP_STATUS = 7;   // 0 = usable
constexpr uint32_t SW_TMEM_COLS = 512;
constexpr uint32_t SW_TMEM_AH = 256, SW_TMEM_AL = 384;   // A operand (row tile, FP16 hi / lo) in TMEM

struct PairEntry {
    uint32_t i;
    uint32_t jw;    // bit 31: weight 2 (off-diagonal tile), else weight 1
    float dt;       // tensor-core distance D~
};

struct SweepParams {
    int T;                        // tiles per side
    long long t_begin, t_end;     // upper-triangular tile range of this launch
    int kblocks;                  // DP / 64
    long long n;
    const float *r;
    const uint32_t *Xh, *Xl;      // FP16 split of s X viewed as 32-bit words (DP / 2 per row)
    float wlo, whi;
    int direct_window;            // host-side note: the window was not derived from a pilot (median_tc_direct_ok)
    const float *e;               // per-row error budget e_i: |D~_ij - D_ij| <= e_i + e_j (err_budget_kernel)
    const float *emax;            // device: max_i e_i
    const float *window;          // device [wlo, whi] replacing the two fields above when non-NULL (set by
                                  // pilot_pick_kernel without a host round trip)
    const float *rmax;            // device: max_i r_i (bounds the column part of the error term)
    const float *scale;           // device: [s, s^2, 1/s^2], s = power of two applied to X before the FP16 split
    // listed D~ -> bin floor((D~ - hlo) * hscale), clamped to [0, SW_HIST_BINS); hlo / hscale are
    // derived in the kernel from the window and rmax (a device value)
    unsigned long long *hist;     // [SW_HIST_BINS] weighted counts of the listed D~ (accumulated)
    unsigned long long *cnt_below, *cnt_listed, *cnt_len;   // weighted below / listed counts, list length
    const float *hparams_out;     // [2]: (hlo, hscale), written by hparams_kernel before the sweep
    int *overflow;
    PairEntry *list;
    unsigned long long list_cap;
};

struct SweepBarriers {
    uint64_t full[SW_STAGES], empty[SW_STAGES];
    uint64_t a_full, a_empty;
    uint64_t s_full[2], s_empty[2];
};

// v[c] for a run-time c without spilling the array to local memory: binary select tree
__device__ __forceinline__ uint32_t select32(const uint32_t (&v)[32], int c) {
    uint32_t a[16], b[8], d[4], e[2];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = (c & 1) ? v[2 * k + 1] : v[2 * k];
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] = (c & 2) ? a[2 * k + 1] : a[2 * k];
#pragma unroll
    for (int k = 0; k < 4; ++k) d[k] = (c & 4) ? b[2 * k + 1] : b[2 * k];
#pragma unroll
    for (int k = 0; k < 2; ++k) e[k] = (c & 8) ? d[2 * k + 1] : d[2 * k];
    return (c & 16) ? e[1] : e[0];
}

// monotone (non-decreasing in dt) bin of a listed D~
__device__ __forceinline__ int hist_bin(float dt, float hlo, float hscale) {
    const float x = (dt - hlo) * hscale;
    return x <= 0.0f ? 0 : (x >= (float)(SW_HIST_BINS - 1) ? SW_HIST_BINS - 1 : (int)x);
}

// successor of tile (I, J) in the row-major order of the upper triangle
__device__ __forceinline__ void next_tile(int &I, int &J, int T) {
    if (++J == T) {
        ++I;
        J = I;
    }
}

// ---- classification of one 128 x 128 tile of g = x_i . x_j (shared by both sweep kernels) ----
// With D~ = r_i + r_j - 2 g and |D~ - D| <= e_i + e_j:
//   certainly above the window  <=>  g < (r_i - e_i)/2 + (r_j - e_j)/2 - whi/2
//   certainly below the window  <=>  g > (r_i + e_i)/2 + (r_j + e_j)/2 - wlo/2
// Written with t = g - B_j, B_j = (r_j - e_j)/2 (one shared-memory value per column):
//   above  <=>  t < lo_i,   lo_i = (r_i - e_i)/2 - whi/2
//   below  <=>  t > hi_i,   hi_i = (r_i + e_i)/2 + emax - wlo/2   (emax >= e_j: conservative)
// and everything else is listed.  lo_i / hi_i are moved outwards by `slack` (>> the fp32 rounding
// of these few operations); that only enlarges the listed set, whose members are resolved from
// their own D~ later.  "below" and "above" are decided by the sign of one subtraction each, so
// the three classes are an exact partition of the pairs (t = -0.0 cannot occur with lo_i > 0;
// a zero difference has a clear sign bit and counts as listed).
// Each classification warp is self-contained: it derives the column terms of its own 64 columns
// and stages its listed pairs in its own shared-memory buffer, so the warps of a CTA never wait
// for each other (the block-wide named barriers of the first version cost 10 % of the sweep).
template <int NWG>   // classification warpgroups: 2 (each warp two 32-column chunks) or 4 (one chunk)
struct TileClassifier {
    static constexpr int CHUNKS = 4 / NWG;           // 32-column chunks per warp
    static constexpr int WCOLS = 32 * CHUNKS;        // columns per warp
    static constexpr int WCAP = SW_WARP_CAP * 2 / NWG;   // staged entries per warp
    static constexpr int ETHREADS = 128 * NWG;
    const SweepParams &p;
    PairEntry *wBuf;             // [WCAP] staging of this warp
    float *wCol;                 // [WCOLS] column terms B_j of this warp's chunk(s)
    unsigned int *sHist, *wCount;
    int wg, row, lane;
    uint32_t lane_addr;
    float rmax, emax, hlo, hscale, s2, inv_s2, wlo, whi;
    unsigned int below, listed;

    __device__ TileClassifier(const SweepParams &p_, uint8_t *tail, int warp, int lane_)
        : p(p_), lane(lane_), below(0u), listed(0u) {
        const int ew = warp - 4;                     // classification warp index
        wCount = reinterpret_cast<unsigned int *>(tail + SW_TAIL_COUNTS) + ew;
        wBuf = reinterpret_cast<PairEntry *>(tail + SW_TAIL_BUF) + (size_t)ew * WCAP;
        wCol = reinterpret_cast<float *>(tail + SW_TAIL_COL) + ew * WCOLS;
        sHist = reinterpret_cast<unsigned int *>(tail + SW_TAIL_HIST);
        const int q = warp & 3;
        wg = ew >> 2;
        row = q * 32 + lane;
        lane_addr = (uint32_t)(q * 32) << 16;
        rmax = __ldg(p.rmax);
        emax = __ldg(p.emax);
        wlo = p.window ? __ldg(p.window) : p.wlo;
        whi = p.window ? __ldg(p.window + 1) : p.whi;
        // g arrives scaled by s^2 (the operands are s X): the thresholds are scaled instead of g --
        // exact, s is a power of two
        s2 = __ldg(p.scale + 1);
        inv_s2 = __ldg(p.scale + 2);
        // histogram range of the listed D~: the window widened by the largest possible error term -- computed once
        // per rank by hparams_kernel (one source for the sweep, pick_band_kernel and the host)
        hlo = __ldg(p.hparams_out);
        hscale = __ldg(p.hparams_out + 1);
        if (lane == 0) *wCount = 0u;
        __syncwarp();
    }
    // this thread's row i against the 128 columns of tile J; the g tile sits in TMEM at `s_tmem`
    // (column 0 of the buffer, lane field 0); w = weight of the tile (0: nothing to do)
    __device__ void classify(uint32_t s_tmem, long long i, int J, unsigned int w) {
        if (w == 0u) return;
        const bool row_ok = i < p.n;
        const float r_i = row_ok ? p.r[i] : 0.0f, e_i = row_ok ? p.e[i] : 0.0f;
        const float *rj = p.r + (size_t)J * 128 + wg * WCOLS;    // this warp's columns
        const float *ej = p.e + (size_t)J * 128 + wg * WCOLS;
        // column terms (the previous tile's are no longer read: same warp, program order)
        {
            const long long j0 = (long long)J * 128 + wg * WCOLS + lane;
#pragma unroll
            for (int cc = 0; cc < CHUNKS; ++cc)
                wCol[32 * cc + lane] =
                    j0 + 32 * cc < p.n ? 0.5f * (__ldg(rj + 32 * cc + lane) - __ldg(ej + 32 * cc + lane)) * s2 : INFINITY;
            __syncwarp();
        }
        const float slack = (r_i + rmax) * 9.5367431640625e-07f;   // 2^-20
        // rows beyond n: lo_i = hi_i = +inf, so every pair is "above" (neither counted nor listed)
        const float lo_i = row_ok ? (0.5f * (r_i - e_i) - 0.5f * whi - 2.0f * slack) * s2 : INFINITY;
        const float hi_i = row_ok ? (0.5f * (r_i + e_i) + emax - 0.5f * wlo + 2.0f * slack) * s2 : INFINITY;
#pragma unroll 1
        for (int cc = 0; cc < CHUNKS; ++cc) {
            const int ch = wg * CHUNKS + cc;
            const long long jbase = (long long)J * 128 + ch * 32;
            uint32_t v[32];
            tmem_ld32(s_tmem + lane_addr + ch * 32, v);
            tmem_wait_ld();
            // per element: t = g - B_j, then the SIGN BITS of (hi_i - t) [set: below the window] and
            // of (t - lo_i) [set: above] are shifted into two masks -- 5 instructions, no
            // predicates.  Element c ends up at bit 31 - c.
            uint32_t mb = 0u, ma = 0u;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 bj = *reinterpret_cast<const float4 *>(wCol + cc * 32 + 4 * c4);
                const float bjs[4] = {bj.x, bj.y, bj.z, bj.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float tt = __uint_as_float(v[4 * c4 + k]) - bjs[k];
                    mb = __funnelshift_l(__float_as_uint(hi_i - tt), mb, 1);
                    ma = __funnelshift_l(__float_as_uint(tt - lo_i), ma, 1);
                }
            }
            const uint32_t hmask = __brev(~(mb | ma));        // bit c <=> column c is listed
            below += w * (unsigned)__popc(mb);
            listed += w * (unsigned)__popc(hmask);
            // rare path (~0.7 % of the pairs, ~0.2 hits per thread and chunk): each thread with
            // hits reserves its slots with one shared-memory atomic, then walks its set bits; the
            // value of column c is picked out of the 32 registers with a 5-level select tree
            if (hmask) {
                unsigned int slot = atomicAdd(wCount, (unsigned)__popc(hmask));
                uint32_t hm = hmask;
                while (hm) {
                    const int c = __ffs(hm) - 1;
                    hm &= hm - 1u;
                    const float g = __uint_as_float(select32(v, c));
                    PairEntry e;
                    e.i = (uint32_t)i;
                    e.jw = (uint32_t)(jbase + c) | (w == 2u ? 0x80000000u : 0u);
                    e.dt = fmaf(-2.0f * inv_s2, g, r_i + __ldg(rj + cc * 32 + c));
                    if (slot < (unsigned)WCAP) {
                        wBuf[slot] = e;
                    } else {   // staging full (degenerate data): straight to global
                        const unsigned long long gi = atomicAdd(p.cnt_len, 1ull);
                        if (gi < p.list_cap) p.list[gi] = e;
                        else *p.overflow = 1;
                        atomicAdd(&sHist[hist_bin(e.dt, hlo, hscale)], w);
                    }
                    ++slot;
                }
            }
        }
    }
    // once per tile, whole warp: move the staged entries to the global list when the buffer is
    // half full (or at the last tile) -- one global reservation per flush
    __device__ void flush(bool last) {
        __syncwarp();
        const unsigned int have = min(*wCount, (unsigned)WCAP);
        if (!(have > (unsigned)WCAP / 2 || (last && have > 0))) return;
        unsigned long long base = 0ull;
        if (lane == 0) base = atomicAdd(p.cnt_len, (unsigned long long)have);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (unsigned int e = lane; e < have; e += 32) {
            const PairEntry pe = wBuf[e];
            if (base + e < p.list_cap) p.list[base + e] = pe;
            else *p.overflow = 1;
            atomicAdd(&sHist[hist_bin(pe.dt, hlo, hscale)], (pe.jw >> 31) ? 2u : 1u);
        }
        __syncwarp();
        if (lane == 0) *wCount = 0u;
        __syncwarp();
    }
    // after the last tile: counts, then (all classification warps together) the histogram
    __device__ void finish(int ew_tid) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            below += __shfl_xor_sync(0xffffffffu, below, o);
            listed += __shfl_xor_sync(0xffffffffu, listed, o);
        }
        if (lane == 0) {
            if (below) atomicAdd(p.cnt_below, (unsigned long long)below);
            if (listed) atomicAdd(p.cnt_listed, (unsigned long long)listed);
        }
        named_bar_sync(3, ETHREADS);
        for (int b = ew_tid; b < SW_HIST_BINS; b += ETHREADS) {
            const unsigned int v = sHist[b];
            if (v) atomicAdd(&p.hist[b], (unsigned long long)v);
        }
    }
};

// The row tile X_I (A operand) lives in TENSOR MEMORY (columns 256..511: FP16 hi and lo,
// two elements per 32-bit column), written by the classification warps when I changes.
// That leaves all of shared memory to a 12-stage ring of column-tile boxes, deep enough to
// hide the L2 latency of the TMA loads; with A resident in shared memory (5 stages) the
// sweep ran at 14 % tensor-pipe utilisation.
__global__ void __launch_bounds__(SW_THREADS, 1)
sweep_tc_kernel(const __grid_constant__ CUtensorMap mapXh, const __grid_constant__ CUtensorMap mapXl,
                const SweepParams p) {
    extern __shared__ uint8_t smem_raw[];
    // (pointer arithmetic on the __shared__ array keeps the shared address space visible to the
    //  compiler: LDS/STS/ATOMS instead of generic LD/ST/ATOM)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sRing = smem;
    uint8_t *tail = sRing + (size_t)SW_STAGES * SW_UNIT_BYTES;
    SweepBarriers *bars = reinterpret_cast<SweepBarriers *>(tail);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail + 256);
    unsigned int *sHist = reinterpret_cast<unsigned int *>(tail + SW_TAIL_HIST);

    const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
    const long long NT = p.t_end - p.t_begin;
    const long long my0 = p.t_begin + NT * blockIdx.x / gridDim.x;
    const long long my1 = p.t_begin + NT * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SW_STAGES; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->a_full, SW_EPI_THREADS);
        mbar_init(&bars->a_empty, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->s_full[b], 1);
            mbar_init(&bars->s_empty[b], SW_EPI_WARPS);
        }
        fence_barrier_init();
        fence_proxy_async();
        tma_prefetch_desc(&mapXh);
        tma_prefetch_desc(&mapXl);
    }
    for (int b = threadIdx.x; b < SW_HIST_BINS; b += SW_THREADS) sHist[b] = 0u;
    if (warp == 1) tmem_alloc(tmem_slot, SW_TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0) {
            // ===================== TMA producer: column tiles only =====================
            // whole warp in uniform control flow; one elected lane issues the copies
            int stage = 0;
            uint32_t phase = 0;
            int I = 0, J = 0;
            if (my0 < my1) tri_tile(my0, p.T, I, J);
            for (long long t = my0; t < my1; ++t, next_tile(I, J, p.T)) {
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    for (int part = 0; part < 2; ++part) {
                        mbar_wait(&bars->empty[stage], phase ^ 1);
                        if (elect_one_sync()) {
                            mbar_expect_tx(&bars->full[stage], SW_UNIT_BYTES);
                            tma_load_2d(sRing + (size_t)stage * SW_UNIT_BYTES, part == 0 ? &mapXh : &mapXl,
                                        &bars->full[stage], kb * 64, J * 128);
                        }
                        __syncwarp();
                        if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            // The whole warp runs the loop (uniform control flow, descriptors in uniform
            // registers); one elected lane issues the tcgen05 instructions.
            const uint32_t idesc = make_idesc(FMT_F16, 128, 128);
            int stage = 0;
            uint32_t phase = 0;
            int prevI = -1, aseg = 0;
            long long jj = 0;
            int I = 0, J = 0;
            if (my0 < my1) tri_tile(my0, p.T, I, J);
            for (long long t = my0; t < my1; ++t, ++jj, next_tile(I, J, p.T)) {
                if (I != prevI) {
                    if (prevI != -1 && elect_one_sync()) tcgen05_commit(&bars->a_empty);   // every MMA on the old A tile is issued
                    __syncwarp();
                    mbar_wait(&bars->a_full, (uint32_t)(aseg & 1));
                    tcgen05_fence_after();
                    ++aseg;
                    prevI = I;
                }
                const int b = (int)(jj & 1);
                if (jj >= 2) {
                    mbar_wait(&bars->s_empty[b], (uint32_t)(((jj >> 1) - 1) & 1));
                    tcgen05_fence_after();
                }
                const uint32_t d_tmem = tmem + b * 128;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    const uint32_t ah = tmem + SW_TMEM_AH + kb * 32, al = tmem + SW_TMEM_AL + kb * 32;
                    mbar_wait(&bars->full[stage], phase);
                    tcgen05_fence_after();
                    uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sRing + (size_t)stage * SW_UNIT_BYTES));
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            umma_f16_ts(d_tmem, ah + 8 * k4, bdesc + 2 * k4, idesc, (kb | k4) != 0);   // hi.hi
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) umma_f16_ts(d_tmem, al + 8 * k4, bdesc + 2 * k4, idesc, 1u);   // lo.hi
                        tcgen05_commit(&bars->empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
                    mbar_wait(&bars->full[stage], phase);
                    tcgen05_fence_after();
                    bdesc = make_kmajor_sw128_desc(smem_u32(sRing + (size_t)stage * SW_UNIT_BYTES));
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) umma_f16_ts(d_tmem, ah + 8 * k4, bdesc + 2 * k4, idesc, 1u);   // hi.lo
                        tcgen05_commit(&bars->empty[stage]);
                        if (kb == p.kblocks - 1) tcgen05_commit(&bars->s_full[b]);
                    }
                    __syncwarp();
                    if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ===================== classification warpgroups =====================
        TileClassifier<2> tc(p, tail, warp, lane);
        const int wg = tc.wg, row = tc.row;
        const uint32_t lane_addr = tc.lane_addr;
        const int wpr = p.kblocks * 32;             // 32-bit words per row of Xh / Xl
        int prevI = -1, aseg = 0;
        long long jj = 0;
        int I = 0, J = 0;
        if (my0 < my1) tri_tile(my0, p.T, I, J);
        for (long long t = my0; t < my1; ++t, ++jj) {
            const long long i = (long long)I * 128 + row;
            if (I != prevI) {
                // new row tile: (re)write the A operand in tensor memory.  Warpgroup 0 writes
                // the hi words of row i, warpgroup 1 the lo words.
                if (aseg > 0) {
                    mbar_wait(&bars->a_empty, (uint32_t)((aseg - 1) & 1));
                    tcgen05_fence_after();
                }
                const uint4 *src = reinterpret_cast<const uint4 *>((wg ? p.Xl : p.Xh) + (size_t)i * wpr);
                const uint32_t dst = tmem + lane_addr + (wg ? SW_TMEM_AL : SW_TMEM_AH);
                for (int ch = 0; ch < p.kblocks; ++ch) {
                    uint32_t v[32];
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const uint4 u = __ldg(src + ch * 8 + q4);
                        v[4 * q4] = u.x; v[4 * q4 + 1] = u.y; v[4 * q4 + 2] = u.z; v[4 * q4 + 3] = u.w;
                    }
                    tmem_st32(dst + ch * 32, v);
                }
                tmem_wait_st();
                tcgen05_fence_before();
                mbar_arrive(&bars->a_full);
                ++aseg;
                prevI = I;
            }
            const int b = (int)(jj & 1);
            mbar_wait(&bars->s_full[b], (uint32_t)((jj >> 1) & 1));
            tcgen05_fence_after();
            tc.classify(tmem + b * 128, i, J, (I == J) ? 1u : 2u);
            // S buffer b may be overwritten by the GEMM of tile jj + 2 (one arrival per warp)
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->s_empty[b]);
            next_tile(I, J, p.T);
            tc.flush(t + 1 == my1);
        }
        tc.finish((warp - 4) * 32 + lane);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem, SW_TMEM_COLS);
    }
}

// =====================================================================================
// CTA-pair sweep (cta_group::2).  The single-CTA kernel above is bound by tensor-memory reads:
// its A operand comes from TMEM (4 KB per 128x128x16 MMA at 64 B/cycle = the whole MMA time)
// next to the 64 KB per tile the classification reads.  Here two CTAs of a cluster own 256
// rows (row tiles 2 I2 and 2 I2 + 1), the leader issues M = 256 MMAs whose B operand is split
// between the two CTAs (each stages 64 of the 128 rows of X_J).  The hi half of a CTA's A tile
// lives in SHARED memory (used by two of the three passes), the lo half in TENSOR memory (used
// by one): per tile that is 288 KB of shared-memory traffic (75 % of the port) and 128 KB of
// tensor-memory reads (67 %) -- with both halves in shared memory the port was the limit
// (117 of 128 B/cycle), with both in tensor memory its read port.
// Tiles: pair rows I2 = 0 .. ceil(T/2)-1, columns J = 2 I2 .. T-1; the half tile below the
// diagonal (row tile 2 I2 + 1 against column tile 2 I2) is computed but not classified.
// =====================================================================================
constexpr int SW2_STAGES = 8;
constexpr int SW2_KB = 4;                       // K blocks of 64 FP16 staged for A (DP <= 256)
constexpr uint32_t SW2_TMEM_COLS = 512;         // two 128-column g buffers + the lo half of the A tile
constexpr uint32_t SW2_TMEM_AL = 256;           // A lo (FP16 pairs, 32 columns per K block of 64)

struct Sweep2Barriers {
    uint64_t full[SW2_STAGES], empty[SW2_STAGES];
    uint64_t a_full, a_empty;
    uint64_t al_full;                // leader: A lo written to tensor memory by both CTAs' classification warps
    uint64_t s_full[2], s_empty[2];
};

// pair tile t (row-major over I2, J >= 2 I2) -> (I2, J)
__host__ __device__ inline void pair_tile(long long t, int T, int &I2, int &J) {
    int i2 = 0;
    while (t >= (long long)(T - 2 * i2)) {
        t -= T - 2 * i2;
        ++i2;
    }
    I2 = i2;
    J = 2 * i2 + (int)t;
}
__device__ __forceinline__ void next_pair_tile(int &I2, int &J, int T) {
    if (++J == T) {
        ++I2;
        J = 2 * I2;
    }
}
static int64_t num_pair_tiles(int64_t T) {
    int64_t nt = 0;
    for (int64_t i2 = 0; 2 * i2 < T; ++i2) nt += T - 2 * i2;
    return nt;
}

// Tile ranges of the clusters of one launch (pair tile indices; n + 1 entries, n = 0: the even split of
// [t_begin, t_end)).  A cluster pays for every pair row it enters (its A operand: 2 x 128 rows x ld, loaded behind
// a drained MMA pipeline) about as much as for a tile, and the rows at the bottom of the triangle are short: an even
// split of the TILES leaves the last cluster with dozens of row changes while all others wait for it (and, sharded,
// the last rank).  sweep_chunk_bounds cuts the tile sequence into world x clusters chunks of equal COST
// (tiles + beta per row entered), the same cut on every rank; rank r owns chunks [r clusters, (r + 1) clusters).
constexpr int SW2_MAX_CLUSTERS = 80;
struct SweepBounds {
    int n;
    int b[SW2_MAX_CLUSTERS + 1];
};
static std::vector<long long> sweep_chunk_bounds(int64_t T, int parts, double beta) {
    const int64_t R = (T + 1) / 2, NT = num_pair_tiles(T);
    auto cut = [&](double C, std::vector<long long> *out) -> int {
        int chunks = 0;
        long long pos = 0;
        double cost = 0.0;
        if (out) out->assign(1, 0);
        for (int64_t i = 0; i < R; ++i) {
            long long rem = T - 2 * i;
            while (rem > 0) {
                if (cost > 0.0 && cost + beta + 1.0 > C) {      // no room for this row's operand and one tile
                    ++chunks;
                    if (out) out->push_back(pos);
                    cost = 0.0;
                }
                cost += beta;
                const long long room = std::max<long long>(1, (long long)std::floor(C - cost));
                const long long take = std::min(rem, room);
                cost += (double)take;
                rem -= take;
                pos += take;
                if (rem > 0) {                                   // chunk full in the middle of the row
                    ++chunks;
                    if (out) out->push_back(pos);
                    cost = 0.0;
                }
            }
        }
        if (cost > 0.0) {
            ++chunks;
            if (out) out->push_back(pos);
        }
        return chunks;
    };
    double lo = 1.0 + beta, hi = (double)NT + beta * (double)R + 1.0;
    for (int it = 0; it < 60 && hi - lo > 1e-3; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (cut(mid, nullptr) <= parts) hi = mid;
        else lo = mid;
    }
    std::vector<long long> b;
    cut(hi, &b);
    while ((int)b.size() < parts + 1) b.push_back(NT);          // (fewer chunks than parts: empty ones at the end)
    return b;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SW2_THREADS, 1)
sweep2_tc_kernel(const __grid_constant__ CUtensorMap mapXh, const __grid_constant__ CUtensorMap mapXl,
                 const __grid_constant__ CUtensorMap mapXh64, const __grid_constant__ CUtensorMap mapXl64,
                 const SweepParams p, const __grid_constant__ SweepBounds bnd) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;                                              // A hi: SW2_KB x 16 KB
    uint8_t *sRing = sA + (size_t)SW2_KB * SW_UNIT_BYTES;            // SW2_STAGES x [64 rows hi | 64 rows lo]
    uint8_t *tail = sRing + (size_t)SW2_STAGES * SW_UNIT_BYTES;
    Sweep2Barriers *bars = reinterpret_cast<Sweep2Barriers *>(tail);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail + 256);
    unsigned int *sHist = reinterpret_cast<unsigned int *>(tail + SW_TAIL_HIST);

    const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const long long NT = p.t_end - p.t_begin;
    const int ncl = (int)gridDim.x / 2, cl = (int)blockIdx.x / 2;
    const long long my0 = bnd.n ? (long long)bnd.b[cl] : p.t_begin + NT * cl / ncl;
    const long long my1 = bnd.n ? (long long)bnd.b[cl + 1] : p.t_begin + NT * (cl + 1) / ncl;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SW2_STAGES; ++s) {
            mbar_init(&bars->full[s], 1);            // leader: one expect_tx arrival, bytes from both CTAs
            mbar_init(&bars->empty[s], 1);           // multicast commit
        }
        mbar_init(&bars->a_full, 1);                 // leader
        mbar_init(&bars->a_empty, 1);                // multicast commit
        mbar_init(&bars->al_full, 2 * 4 * SW2_EPI_WG);   // leader: one arrival per classification warp of both CTAs
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->s_full[b], 1);                        // multicast commit
            mbar_init(&bars->s_empty[b], 2 * 4 * SW2_EPI_WG);      // leader: both CTAs' classification warps
        }
        fence_barrier_init();
        fence_proxy_async();
        tma_prefetch_desc(&mapXh);
        tma_prefetch_desc(&mapXl);
        tma_prefetch_desc(&mapXh64);
        tma_prefetch_desc(&mapXl64);
    }
    for (int b = threadIdx.x; b < SW_HIST_BINS; b += SW2_THREADS) sHist[b] = 0u;
    if (warp == 1) tmem_alloc_pair(tmem_slot, SW2_TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();            // peer barriers are initialised before anyone signals them
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        // 640 threads: 96 registers each at launch; the two control warps give theirs to nobody
        // (128 x 48 + 512 x 104 = 59 392 <= 65 536)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        if (warp == 0) {
            // ===================== TMA producer (both CTAs) =====================
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full0_addr = mapa_shared(smem_u32(&bars->full[0]), 0);
            const uint32_t a_full_addr = mapa_shared(smem_u32(&bars->a_full), 0);
            int prevI2 = -1, aseg = 0;
            int I2 = 0, J = 0;
            if (my0 < my1) pair_tile(my0, p.T, I2, J);
            for (long long t = my0; t < my1; ++t, next_pair_tile(I2, J, p.T)) {
                if (I2 != prevI2) {
                    // this CTA's 128 rows of the pair row, hi and lo
                    if (aseg > 0) mbar_wait(&bars->a_empty, (uint32_t)((aseg - 1) & 1));
                    const int arow = (2 * I2 + (int)rank) * 128;
                    if (elect_one_sync()) {
                        if (leader) mbar_expect_tx(&bars->a_full, 2u * (uint32_t)p.kblocks * SW_UNIT_BYTES);
                        for (int kb = 0; kb < p.kblocks; ++kb)
                            tma_load_2d_pair(sA + (size_t)kb * SW_UNIT_BYTES, &mapXh, a_full_addr, kb * 64, arow);
                    }
                    __syncwarp();
                    ++aseg;
                    prevI2 = I2;
                }
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);      // local: multicast commit of the leader
                    if (elect_one_sync()) {
                        if (leader) mbar_expect_tx(&bars->full[stage], 2u * SW_UNIT_BYTES);
                        uint8_t *dst = sRing + (size_t)stage * SW_UNIT_BYTES;
                        const uint32_t fa = full0_addr + 8u * (uint32_t)stage;
                        tma_load_2d_pair(dst, &mapXh64, fa, kb * 64, J * 128 + (int)rank * 64);
                        tma_load_2d_pair(dst + SW_UNIT_BYTES / 2, &mapXl64, fa, kb * 64, J * 128 + (int)rank * 64);
                    }
                    __syncwarp();
                    if (++stage == SW2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1 && leader) {
            // ===================== MMA issuer (leader CTA only) =====================
            const uint32_t idesc = make_idesc(FMT_F16, 256, 128);
            int stage = 0;
            uint32_t phase = 0;
            int prevI2 = -1, aseg = 0;
            long long jj = 0;
            int I2 = 0, J = 0;
            if (my0 < my1) pair_tile(my0, p.T, I2, J);
            for (long long t = my0; t < my1; ++t, ++jj, next_pair_tile(I2, J, p.T)) {
                if (I2 != prevI2) {
                    if (prevI2 != -1 && elect_one_sync()) tcgen05_commit_pair(&bars->a_empty);   // old A tiles are free
                    __syncwarp();
                    mbar_wait(&bars->a_full, (uint32_t)(aseg & 1));
                    mbar_wait(&bars->al_full, (uint32_t)(aseg & 1));
                    tcgen05_fence_after();
                    ++aseg;
                    prevI2 = I2;
                }
                const int b = (int)(jj & 1);
                if (jj >= 2) {
                    mbar_wait(&bars->s_empty[b], (uint32_t)(((jj >> 1) - 1) & 1));
                    tcgen05_fence_after();
                }
                const uint32_t d_tmem = tmem + b * 128;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    const uint64_t ah = make_kmajor_sw128_desc(smem_u32(sA + (size_t)kb * SW_UNIT_BYTES));
                    const uint32_t al = tmem + SW2_TMEM_AL + kb * 32;
                    mbar_wait(&bars->full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t slot_addr = smem_u32(sRing + (size_t)stage * SW_UNIT_BYTES);
                    const uint64_t bh = make_kmajor_sw128_desc(slot_addr);
                    const uint64_t bl = make_kmajor_sw128_desc(slot_addr + SW_UNIT_BYTES / 2);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) umma2_f16_ss(d_tmem, ah + 2 * k4, bh + 2 * k4, idesc, (kb | k4) != 0);   // hi.hi
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) umma2_f16_ts(d_tmem, al + 8 * k4, bh + 2 * k4, idesc, 1u);             // lo.hi (A from TMEM)
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) umma2_f16_ss(d_tmem, ah + 2 * k4, bl + 2 * k4, idesc, 1u);             // hi.lo
                        tcgen05_commit_pair(&bars->empty[stage]);
                        if (kb == p.kblocks - 1) tcgen05_commit_pair(&bars->s_full[b]);
                    }
                    __syncwarp();
                    if (++stage == SW2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ===================== classification warpgroups (both CTAs, own 128 rows) =====================
        TileClassifier<SW2_EPI_WG> tc(p, tail, warp, lane);
        const uint32_t s_empty_addr0 = mapa_shared(smem_u32(&bars->s_empty[0]), 0);
        const uint32_t s_empty_addr1 = mapa_shared(smem_u32(&bars->s_empty[1]), 0);
        const uint32_t al_full_addr = mapa_shared(smem_u32(&bars->al_full), 0);
        const int wpr = p.kblocks * 32;             // 32-bit words per row of Xl
        long long jj = 0;
        int I2 = 0, J = 0, prevI2 = -1, aseg = 0;
        if (my0 < my1) pair_tile(my0, p.T, I2, J);
        for (long long t = my0; t < my1; ++t, ++jj) {
            const int I = 2 * I2 + (int)rank;
            const long long i = (long long)I * 128 + tc.row;
            if (I2 != prevI2) {
                // new pair row: this CTA's A lo tile goes to tensor memory; warp (lane quarter q,
                // chunk c) writes K block c of its 32 rows (rows beyond the matrix are zero)
                if (aseg > 0) {
                    mbar_wait(&bars->a_empty, (uint32_t)((aseg - 1) & 1));
                    tcgen05_fence_after();
                }
                if (tc.wg < p.kblocks) {
                    const bool in = i < (long long)p.T * 128;
                    const uint4 *src = reinterpret_cast<const uint4 *>(p.Xl + (size_t)(in ? i : 0) * wpr) + tc.wg * 8;
                    uint32_t v[32];
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const uint4 u = in ? __ldg(src + q4) : make_uint4(0u, 0u, 0u, 0u);
                        v[4 * q4] = u.x; v[4 * q4 + 1] = u.y; v[4 * q4 + 2] = u.z; v[4 * q4 + 3] = u.w;
                    }
                    tmem_st32(tmem + tc.lane_addr + SW2_TMEM_AL + tc.wg * 32, v);
                    tmem_wait_st();
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(al_full_addr);
                ++aseg;
                prevI2 = I2;
            }
            const int b = (int)(jj & 1);
            const unsigned int w = J > I ? 2u : (J == I ? 1u : 0u);
            mbar_wait(&bars->s_full[b], (uint32_t)((jj >> 1) & 1));
            tcgen05_fence_after();
            tc.classify(tmem + b * 128, i, J, w);
            // g buffer b may be overwritten by the GEMM of tile jj + 2 (one arrival per warp)
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(b ? s_empty_addr1 : s_empty_addr0);
            next_pair_tile(I2, J, p.T);
            tc.flush(t + 1 == my1);
        }
        tc.finish((warp - 4) * 32 + lane);
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();            // no CTA leaves while its partner may still signal / read it
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_pair(tmem, SW2_TMEM_COLS);
    }
}

// =====================================================================================
// Sweep for more than 256 coordinates (leading dimension 512 / 768 / 1024): the row tile no longer
// fits next to the ring (A hi alone would be 256 KB at d = 1 024), so both operands stream along K
// through the main loop of panel_gemm.cuh -- 256 x 256 tiles (pair rows I2, pair columns J2 >= I2),
// the same three FP16 passes (hi.hi, lo.hi, hi.lo) and the same classification, applied to the two
// 128-column halves of the accumulator.
// =====================================================================================
struct Sweep3Policy {
    static constexpr int STAGES = 6;
    static constexpr size_t TAIL_BYTES = SW_TAIL_BYTES;
    // Tile order: the FP16 hi / lo arrays of 262 144 x 1 024 particles are 1 GB -- far beyond L2 -- so an
    // operand tile only hits when several clusters want it at the same moment.  A wave is therefore a WINDOW of
    // win_a row tiles x win_b column tiles of the upper triangle (9 x 8 on 74 clusters: 17 MB of operands from
    // DRAM per wave instead of up to 148 MB with one private tile pair per cluster, which made the sweep
    // HBM-bound at ~64 % of the tensor rate).  `wins` lists this rank's windows (host-built: those that
    // touch the upper triangle, row-window major).
    struct Params : pg::Core {
        SweepParams sp;
        int T2;                       // 256-row tiles per side
        int win_a, win_b;
        const int2 *wins;             // (row window, column window) of step k
        int nwins;
    };
    __device__ static int tile(const Params &p, long long k, int cl, int, int &ti, int &tj) {
        if (k >= p.nwins || cl >= p.win_a * p.win_b) return 0;
        const int2 w = p.wins[k];
        ti = w.x * p.win_a + cl % p.win_a;
        tj = w.y * p.win_b + cl / p.win_a;
        return (ti < p.T2 && tj < p.T2 && tj >= ti) ? 1 : 2;
    }
    __device__ static void init_shared(uint8_t *tail, int tid) {
        unsigned int *sHist = reinterpret_cast<unsigned int *>(tail + SW_TAIL_HIST);
        for (int b = tid; b < SW_HIST_BINS; b += pg::THREADS) sHist[b] = 0u;
    }
    struct Epilogue {
        TileClassifier<2> tc;
        int I2, J2, ew_tid;
        uint32_t rank;
        __device__ Epilogue(const Params &p, uint8_t *tail, int warp, int lane, uint32_t rank_)
            : tc(p.sp, tail, warp, lane), I2(0), J2(0), ew_tid((warp - 4) * 32 + lane), rank(rank_) {}
        __device__ void tile_begin(int ti, int tj) {
            I2 = ti;
            J2 = tj;
        }
        __device__ void unit(uint32_t acc_tmem, int, bool) {
            const int I = 2 * I2 + (int)rank;
            const long long i = (long long)I * 128 + tc.row;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int J = 2 * J2 + h;
                tc.classify(acc_tmem + (uint32_t)h * 128u, i, J, J > I ? 2u : (J == I ? 1u : 0u));
            }
            tc.flush(false);
        }
        __device__ void finish() {
            tc.flush(true);
            tc.finish(ew_tid);
        }
    };
};

// ---- helpers ---------------------------------------------------------------------------
// Histogram range of the listed D~ (the TileClassifier's formula, same arithmetic), written by EVERY rank before
// its sweep: a rank whose share of the tiles is empty launches no sweep kernel, yet its pick_band_kernel needs
// the same (hlo, hscale) as everybody else's -- rank-dependent band parameters end in ranks that disagree about
// the number of all-reduces that follow.
__global__ void hparams_kernel(const float *__restrict__ window, float wlo, float whi, const float *__restrict__ emax,
                               float *__restrict__ hparams_out) {
    if (window) {
        wlo = window[0];
        whi = window[1];
    }
    const float hpad = 4.0f * emax[0] + 1e-5f * fmaxf(fabsf(wlo), fabsf(whi));
    const float hlo = wlo - hpad;
    hparams_out[0] = hlo;
    hparams_out[1] = (float)SW_HIST_BINS / fmaxf((whi + hpad) - hlo, 1e-30f);
}

// hi = fp16(s x), lo = fp16(s x - hi): 22 bits of s x (entries far below the largest one end in
// the FP16 subnormals; that absolute error is covered by the slack terms, see err_budget_kernel / EPS_ABS)
__global__ void split_f16_kernel(const float *__restrict__ X, int64_t count4, const float *__restrict__ scale,
                                 __half *__restrict__ Xh, __half *__restrict__ Xl) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count4) return;
    const float sc = __ldg(scale);
    const float4 x = reinterpret_cast<const float4 *>(X)[e];
    const float xs[4] = {x.x * sc, x.y * sc, x.z * sc, x.w * sc};
    __half h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        h[k] = __float2half_rn(xs[k]);
        l[k] = __float2half_rn(xs[k] - __half2float(h[k]));
    }
    reinterpret_cast<uint2 *>(Xh)[e] = *reinterpret_cast<uint2 *>(h);
    reinterpret_cast<uint2 *>(Xl)[e] = *reinterpret_cast<uint2 *>(l);
}

// Per-row error budget of the filter: |D~_ij - D_ij| <= e_i + e_j with e_i = sum_m w_m x_im^2.
// The error of g = x_i . x_j is a sum of per-step roundings, each relative to the running sum at
// that step, so a product that enters early is hit by more of them than a late one:
//   contract fma chain (what D is DEFINED by): step k rounds to nearest, <= 2^-24 |s_k| and
//     |s_k| <= (1 + 2^-24)^k sum_{m <= k} |x_m y_m|  =>  weight 2^-24 (ld - m + 1) on |x_m y_m|   [rigorous];
//   tensor-core accumulation: one truncation per MMA (<= 2^-23 of the running sum), 12 MMAs per block of
//     64 coordinates in ascending order in all three sweep kernels  =>  <= 12 (nkb - kb(m)) truncations after
//     product m entered, weight 2^-23 each, doubled as a margin (the datapath is not documented);
//   FP16 split (dropped lo.lo, rounding of lo) 3 * 2^-22, doubled; final roundings of D~ and D: 2^-21.
// err(D) = 2 err(g) <= 2 sum_m w_m |x_m||y_m| <= sum_m w_m x_m^2 + sum_m w_m y_m^2 (Cauchy-Schwarz, AM-GM).
// The flat bound this replaces (weight of the first coordinate for all of them) was 2.6x (d = 256) to
// 3x (d = 1024) larger, and with it the band of pairs whose distance is recomputed exactly.
__global__ void __launch_bounds__(256)
err_budget_kernel(const float *__restrict__ X, int64_t rows, int64_t n, int64_t ld, float *__restrict__ e) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nkb = (int)(ld / 64);
    float acc = 0.0f;
    if (row < n) {
        for (int64_t m = lane; m < ld; m += 32) {
            const float x = X[row * ld + m];
            const float w = 5.9664e-08f * (float)(ld - m + 1)                       // 2^-24 * 1.001
                            + 2.38418579e-07f * 12.0f * (float)(nkb - (int)(m >> 6))    // 2^-22 (= 2 x 2^-23)
                            + 1.90734863e-06f;                                      // 2^-19
            acc = fmaf(w * x, x, acc);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) e[row] = acc * 1.001f;      // the fp32 evaluation of the budget itself
}

// Row norms (contract order: one fp32 fma chain per row, abstract_kernel.py:34 as oracle/svgd_oracle.c defines it)
// AND the error budgets of err_budget_kernel from ONE read of the particles: a block stages 32 rows x 256
// coordinates in shared memory (coalesced), warp 0 walks the 32 chains (lane = row), warps 1..7 take the budgets
// with err_budget_kernel's own lane / order assignment, so both results have the bits of the separate kernels.
constexpr int XS_ROWS = 32, XS_COLS = 256;
__global__ void __launch_bounds__(256)
x_stats_kernel(const float *__restrict__ X, int64_t rows_r, int64_t rows_e, int64_t n, int64_t ld,
               float *__restrict__ r, float *__restrict__ e) {
    __shared__ float tile[XS_ROWS][XS_COLS + 1];
    const int64_t row0 = (int64_t)blockIdx.x * XS_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (int)(ld / 64);
    float chain = 0.0f;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t c0 = 0; c0 < ld; c0 += XS_COLS) {
        const int width = (int)min((int64_t)XS_COLS, ld - c0), w4 = width / 4;
        for (int idx = threadIdx.x; idx < XS_ROWS * w4; idx += 256) {
            const int rr = idx / w4, c4 = idx - rr * w4;
            const int64_t row = row0 + rr;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows_r) v = *reinterpret_cast<const float4 *>(X + row * ld + c0 + 4 * c4);
            tile[rr][4 * c4 + 0] = v.x;
            tile[rr][4 * c4 + 1] = v.y;
            tile[rr][4 * c4 + 2] = v.z;
            tile[rr][4 * c4 + 3] = v.w;
        }
        __syncthreads();
        if (warp == 0) {
            for (int k = 0; k < width; ++k) {
                const float x = tile[lane][k];
                chain = __fmaf_rn(x, x, chain);
            }
        } else {
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                const int rr = warp - 1 + 7 * a;
                if (rr >= XS_ROWS) break;
                float v = acc[a];
                for (int k = lane; k < width; k += 32) {
                    const int64_t m = c0 + k;
                    const float x = tile[rr][k];
                    const float w = 5.9664e-08f * (float)(ld - m + 1)
                                    + 2.38418579e-07f * 12.0f * (float)(nkb - (int)(m >> 6))
                                    + 1.90734863e-06f;
                    v = fmaf(w * x, x, v);
                }
                acc[a] = v;
            }
        }
        __syncthreads();
    }
    if (warp == 0) {
        if (row0 + lane < rows_r) r[row0 + lane] = chain;
    } else {
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            const int rr = warp - 1 + 7 * a;
            if (rr >= XS_ROWS) break;
            float v = acc[a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int64_t row = row0 + rr;
            if (lane == 0 && row < rows_e) e[row] = row < n ? v * 1.001f : 0.0f;
        }
    }
}

__global__ void __launch_bounds__(1024)
max_kernel(const float *__restrict__ r, const float *__restrict__ e, int64_t n, float *__restrict__ out,
           float *__restrict__ emax_out, float *__restrict__ scale_out) {
    __shared__ float red[32];
    __shared__ float red_e[32];
    float m = 0.0f, me = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        m = fmaxf(m, r[i]);
        me = fmaxf(me, e[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) me = fmaxf(me, __shfl_xor_sync(0xffffffffu, me, o));
    if ((threadIdx.x & 31) == 0) red_e[threadIdx.x >> 5] = me;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        me = red_e[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) me = fmaxf(me, __shfl_xor_sync(0xffffffffu, me, o));
        if (threadIdx.x == 0) {
            *out = m;
            *emax_out = me;
            // s = 2^-e with sqrt(rmax) s in [128, 256): no entry of s X exceeds 256
            int e = 0;
            if (m > 0.0f && m < INFINITY) e = ilogbf(sqrtf(m)) - 7;
            e = max(-60, min(60, e));
            scale_out[0] = ldexpf(1.0f, -e);
            scale_out[1] = ldexpf(1.0f, -2 * e);
            scale_out[2] = ldexpf(1.0f, 2 * e);
        }
    }
}

// Pilot sample from the FP16 hi array: APPROXIMATE distances (|error| ~ 2^-11 |x_i||x_j|, about
// 1e-2 at config D against a window of 0.8) are all the pilot needs -- it only places the window,
// and every consumer of the window checks that the target rank is bracketed.  Half the bytes of
// the fp32 chain of pair_chain.cuh and no shared-memory staging: a row of ld halfs is one 16-byte
// load per lane, the 32 per-lane partial sums of 32 pairs are reduced with a 31-shuffle butterfly.
// Same sample as pair_chain_kernel<1>: pair e = splitmix64(seed + e).
__global__ void __launch_bounds__(256)
pilot_h16_kernel(uint32_t *__restrict__ keys, unsigned long long m, const uint4 *__restrict__ Xh,
                 const float *__restrict__ r, const float *__restrict__ scale, int64_t n, int chunks, uint64_t seed) {
    const int lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long ngroups = (m + 31ull) / 32ull;
    const float inv_s2 = __ldg(scale + 2);
    for (unsigned long long g = warp; g < ngroups; g += nwarps) {
        const unsigned long long e = g * 32ull + lane;
        const uint64_t h = splitmix64(seed + (uint64_t)e);
        const uint32_t i = (uint32_t)((h >> 32) % (uint64_t)n), j = (uint32_t)((h & 0xffffffffull) % (uint64_t)n);
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.0f;
        // rows of more than 256 coordinates: 256 at a time (one 16-byte load per lane and row each)
        for (int seg = 0; seg * 32 < chunks; ++seg) {
            const bool active = seg * 32 + lane < chunks;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                uint4 A[8], B[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const uint32_t it = __shfl_sync(0xffffffffu, i, 8 * b + t), jt = __shfl_sync(0xffffffffu, j, 8 * b + t);
                    A[t] = B[t] = make_uint4(0u, 0u, 0u, 0u);
                    if (active) {
                        A[t] = __ldg(Xh + (size_t)it * chunks + seg * 32 + lane);
                        B[t] = __ldg(Xh + (size_t)jt * chunks + seg * 32 + lane);
                    }
                }
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const uint32_t a[4] = {A[t].x, A[t].y, A[t].z, A[t].w}, c[4] = {B[t].x, B[t].y, B[t].z, B[t].w};
                    float acc = 0.0f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 fa = __half22float2(*reinterpret_cast<const __half2 *>(&a[q]));
                        const float2 fb = __half22float2(*reinterpret_cast<const __half2 *>(&c[q]));
                        acc = fmaf(fa.x, fb.x, acc);
                        acc = fmaf(fa.y, fb.y, acc);
                    }
                    v[8 * b + t] += acc;
                }
            }
        }
        // butterfly: after the step with stride s, a lane keeps the half of the pairs whose index
        // bit s equals its own lane bit; at the end v[0] on lane L is the full sum of pair L
#pragma unroll
        for (int st = 16; st >= 1; st >>= 1) {
            const bool up = (lane & st) != 0;
#pragma unroll
            for (int k = 0; k < st; ++k) {
                const float send = up ? v[k] : v[k + st];
                const float keep = up ? v[k + st] : v[k];
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, st);
            }
        }
        if (e < m) keys[e] = float_to_key((r[i] + r[j]) - 2.0f * v[0] * inv_s2);
    }
}

// weighted histogram of the keys that fall into the window [key_lo, key_lo + (nbins << shift)):
// bins[0] += weight of the keys below the window, bins[1 + b] += weight of bin b; keys above the
// window are ignored.  SRC 1: uint2 (key, weight) band entries, SRC 2: plain u32 keys (weight 1).
template <int SRC>
__global__ void __launch_bounds__(256)
window_hist_kernel(const void *__restrict__ src, unsigned long long m, const unsigned long long *__restrict__ m_dev,
                   uint32_t key_lo, uint32_t shift, uint32_t nbins, unsigned long long *__restrict__ bins,
                   const uint32_t *__restrict__ bp = nullptr) {
    extern __shared__ unsigned int wh[];      // nbins + 1
    if (m_dev) m = min(m, *m_dev);            // length only known on the device (m = upper bound)
    if (bp) {                                 // window picked on the device (pick_band_kernel)
        if (bp[BP_STATUS] != 0u) return;
        key_lo = bp[BP_KLO];
        shift = bp[BP_SHIFT];
        nbins = bp[BP_NBINS];
    }
    for (uint32_t b = threadIdx.x; b <= nbins; b += blockDim.x) wh[b] = 0u;
    __syncthreads();
    unsigned int below = 0u;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < m;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t key, w;
        if (SRC == 2) {
            key = reinterpret_cast<const uint32_t *>(src)[e];
            w = 1u;
        } else {
            const uint2 kw = reinterpret_cast<const uint2 *>(src)[e];
            key = kw.x;
            w = kw.y;
        }
        if (key < key_lo) {
            below += w;
        } else {
            const uint32_t b = (key - key_lo) >> shift;
            if (b < nbins) atomicAdd(&wh[1 + b], w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if ((threadIdx.x & 31) == 0 && below) atomicAdd(&wh[0], below);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b <= nbins; b += blockDim.x)
        if (wh[b]) atomicAdd(&bins[b], (unsigned long long)wh[b]);
}

// Device-side pick of the pilot window: bins (1 + nbins u64, counts of a window histogram of the
// pilot keys) -> the keys that bracket rank_lo / rank_hi at bin resolution.  out: [0],[1] =
// window as floats, [2],[3] = as keys, [4] = 1 when both ranks lie inside the histogram window.
__global__ void __launch_bounds__(1024)
pilot_pick_kernel(const unsigned long long *__restrict__ bins, uint32_t key_lo, uint32_t shift, uint32_t nbins,
                  unsigned long long rank_lo, unsigned long long rank_hi, uint32_t *__restrict__ out) {
    __shared__ unsigned long long wsum[32];
    __shared__ long long s_bin[2];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t < 2) s_bin[t] = -1;
    // thread t owns bins [16 t, 16 t + 16) (nbins <= 16384)
    unsigned long long loc[16], tot = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t b = 16u * t + k;
        loc[k] = b < nbins ? bins[1 + b] : 0ull;
        tot += loc[k];
    }
    unsigned long long incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    unsigned long long cum = bins[0] + wsum[warp] + incl - tot;     // weight before this thread's bins
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (rank_lo >= cum && rank_lo < cum + loc[k]) s_bin[0] = 16 * t + k;
        if (rank_hi >= cum && rank_hi < cum + loc[k]) s_bin[1] = 16 * t + k;
        cum += loc[k];
    }
    __syncthreads();
    if (t == 0) {
        const bool ok = s_bin[0] >= 0 && s_bin[1] >= 0 && rank_lo >= bins[0];
        unsigned long long lo = (unsigned long long)key_lo + ((unsigned long long)(ok ? s_bin[0] : 0) << shift);
        unsigned long long hi = (unsigned long long)key_lo + (((unsigned long long)(ok ? s_bin[1] : 0) + 1) << shift) - 1;
        if (hi > 0xffffffffull) hi = 0xffffffffull;
        out[0] = __float_as_uint(key_to_float((uint32_t)lo));
        out[1] = __float_as_uint(key_to_float((uint32_t)hi));
        out[2] = (uint32_t)lo;
        out[3] = (uint32_t)hi;
        out[4] = ok ? 1u : 0u;
    }
}

// Device-side counterpart of the host logic after the sweep: from the (all-reduced) counters and
// the histogram of the listed D~, locate the bin of the target rank and derive the band
// thresholds and the key window of the final select.  cnt = counters block (u64 slots).
__global__ void __launch_bounds__(1024)
pick_band_kernel(const unsigned long long *__restrict__ cnt, int slot_below, int slot_listed, int slot_overflow,
                 int slot_hist, int slot_rmax, int slot_emax, int slot_hparams, unsigned long long rank0,
                 unsigned long long rank1, float eps_abs_coeff, uint32_t max_bins, uint32_t *__restrict__ bp) {
    __shared__ unsigned long long wsum[32];
    __shared__ int s_bin;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_bin = -1;
    unsigned long long loc[4], tot = 0;          // thread t owns bins [4 t, 4 t + 4) of SW_HIST_BINS = 4096
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        loc[k] = cnt[slot_hist + 4 * t + k];
        tot += loc[k];
    }
    unsigned long long incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    const unsigned long long below = cnt[slot_below], listed = cnt[slot_listed];
    unsigned long long cum = below + wsum[warp] + incl - tot;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (rank0 >= cum && rank0 < cum + loc[k]) s_bin = 4 * t + k;
        cum += loc[k];
    }
    __syncthreads();
    if (t != 0) return;
    const int bstar = s_bin;
    const bool ok = cnt[slot_overflow] == 0 && below <= rank0 && rank1 < below + listed && bstar > 0 &&
                    bstar < SW_HIST_BINS - 1;
    const float rmax = __uint_as_float((uint32_t)cnt[slot_rmax]);
    const float emax = __uint_as_float((uint32_t)cnt[slot_emax]);
    const uint32_t hp0 = (uint32_t)cnt[slot_hparams], hp1 = (uint32_t)(cnt[slot_hparams] >> 32);
    const float hlo = __uint_as_float(hp0), hscale = __uint_as_float(hp1);
    // t~ lies in bin bstar; one extra bin on either side covers the rounding of the bin function
    const float bin_w = 1.0f / hscale;
    const float t_lo = hlo + (float)(bstar - 1) * bin_w, t_hi = hlo + (float)(bstar + 2) * bin_w;
    const float eps_abs = eps_abs_coeff * rmax;
    const float delta = 2.0f * emax * 1.0001f + eps_abs + 2.0f * fmaxf(fabsf(t_lo), fabsf(t_hi)) * 1.2e-7f;
    const float tlo = t_lo - delta, thi = t_hi + delta;
    const uint32_t klo = float_to_key(tlo - 2.0f * delta), khi = float_to_key(thi + 2.0f * delta);
    const unsigned long long span = (unsigned long long)khi - klo + 1;
    uint32_t sh = 0;
    while (((span - 1) >> sh) >= (unsigned long long)max_bins) ++sh;
    bp[BP_TLO] = __float_as_uint(tlo);
    bp[BP_THI] = __float_as_uint(thi);
    bp[BP_DELTA] = __float_as_uint(delta);
    bp[BP_EPS_ABS] = __float_as_uint(eps_abs);
    bp[BP_KLO] = klo;
    bp[BP_SHIFT] = sh;
    bp[BP_NBINS] = (uint32_t)(((span - 1) >> sh) + 1);
    bp[BP_STATUS] = (ok && khi >= klo) ? 0u : 1u;
}

// entries certainly below t~ - delta are counted; entries that may fall inside
// [t~ - delta, t~ + delta] are compacted into the band (key slot filled by pair_chain_kernel).
// A block walks a contiguous part of the list and stages its band entries in shared memory, so
// the global cursor is advanced once per ~1000 band entries (one atomic per entry, and even one
// per warp, serialised on that single address: 77 % of the kernel was spent waiting for it).
constexpr int BF_THREADS = 256;
constexpr int BF_UNROLL = 4;               // entries per thread and round: 4 independent load / gather chains in flight
constexpr int BF_STAGE = 4096;
__global__ void __launch_bounds__(BF_THREADS)
band_filter_kernel(const PairEntry *__restrict__ list, unsigned long long m, const float *__restrict__ ebud,
                   float tlo, float thi, float eps_abs, unsigned long long *__restrict__ below_out,
                   unsigned long long *__restrict__ bandw_out, unsigned long long *__restrict__ bandlen_out,
                   uint2 *__restrict__ band_ij, unsigned long long band_cap, int *__restrict__ overflow,
                   const unsigned long long *__restrict__ m_dev = nullptr, const uint32_t *__restrict__ bp = nullptr) {
    __shared__ uint2 stage[BF_STAGE];
    if (m_dev) m = min(m, *m_dev);            // list length / thresholds that only exist on the device
    if (bp) {
        if (bp[BP_STATUS] != 0u) return;
        tlo = __uint_as_float(bp[BP_TLO]);
        thi = __uint_as_float(bp[BP_THI]);
        eps_abs = __uint_as_float(bp[BP_EPS_ABS]);
    }
    __shared__ unsigned int s_count;
    __shared__ unsigned long long s_base;
    if (threadIdx.x == 0) s_count = 0u;
    __syncthreads();
    unsigned int below = 0u, bw = 0u, pending = 0u;
    const int lane = threadIdx.x & 31;
    const unsigned long long per = (m + gridDim.x - 1) / gridDim.x;
    const unsigned long long b0 = per * blockIdx.x, b1 = b0 + per < m ? b0 + per : m;
    auto flush = [&]() {     // whole block
        __syncthreads();
        const unsigned int cnt = s_count;
        if (threadIdx.x == 0 && cnt) s_base = atomicAdd(bandlen_out, (unsigned long long)cnt);
        __syncthreads();
        const unsigned long long base = s_base;
        for (unsigned int e = threadIdx.x; e < cnt; e += BF_THREADS) {
            if (base + e < band_cap) band_ij[base + e] = stage[e];
            else *overflow = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_count = 0u;
        __syncthreads();
    };
    for (unsigned long long e0 = b0; e0 < b1; e0 += BF_UNROLL * BF_THREADS) {
        PairEntry pe[BF_UNROLL];
        bool valid[BF_UNROLL], in_band[BF_UNROLL];
#pragma unroll
        for (int u = 0; u < BF_UNROLL; ++u) {
            const unsigned long long e = e0 + (unsigned long long)u * BF_THREADS + threadIdx.x;
            valid[u] = e < b1;
            pe[u] = PairEntry{};
            if (valid[u]) pe[u] = list[e];
        }
        float eps[BF_UNROLL];
#pragma unroll
        for (int u = 0; u < BF_UNROLL; ++u) {
            eps[u] = 0.0f;
            if (valid[u]) eps[u] = ebud[pe[u].i] + ebud[pe[u].jw & 0x7fffffffu] + eps_abs;
        }
#pragma unroll
        for (int u = 0; u < BF_UNROLL; ++u) {
            in_band[u] = false;
            if (valid[u]) {
                const uint32_t w = (pe[u].jw >> 31) ? 2u : 1u;
                if (pe[u].dt + eps[u] < tlo) {
                    below += w;
                } else if (pe[u].dt - eps[u] <= thi) {
                    bw += w;
                    in_band[u] = true;
                }
            }
            const unsigned int vote = __ballot_sync(0xffffffffu, in_band[u]);
            if (vote) {
                unsigned int base = 0u;
                if (lane == 0) base = atomicAdd(&s_count, (unsigned)__popc(vote));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (in_band[u]) stage[base + (unsigned)__popc(vote & ((1u << lane) - 1u))] = make_uint2(pe[u].i, pe[u].jw);
            }
        }
        // flush while one more round of at most BF_UNROLL * BF_THREADS entries is still guaranteed to fit
        // (decided on the round count, which is uniform without reading the shared counter)
        pending += BF_UNROLL * BF_THREADS;
        if (pending > BF_STAGE - BF_UNROLL * BF_THREADS) {
            flush();
            pending = 0;
        }
    }
    flush();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        below += __shfl_xor_sync(0xffffffffu, below, o);
        bw += __shfl_xor_sync(0xffffffffu, bw, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (below) atomicAdd(below_out, (unsigned long long)below);
        if (bw) atomicAdd(bandw_out, (unsigned long long)bw);
    }
}

// worst-case |D~ - D_contract| <= e_i + e_j + EPS_ABS * rmax: the per-row budgets e_i (err_budget_kernel)
// cover the FP16 split, the tensor-core accumulation, the contract chain and the final roundings; for
// entries that end in the FP16 subnormals the split adds <= 2^-25 absolute per entry of s X,
// i.e. <= 2^-23 sqrt(d) sqrt(rmax) / s <= 2^-26 d^(1/2) rmax / 16 in D: the EPS_ABS term.
constexpr float EPS_ABS = 2.384185791015625e-07f;    // 2^-22 (d <= 256: 2^-26 * 16 / 16 = 2^-26, x16 margin)
// the subnormal term grows with sqrt(d): keep the margin beyond 256 coordinates
static float eps_abs_coeff(int64_t d) { return d > 256 ? EPS_ABS * sqrtf((float)d / 256.0f) : EPS_ABS; }

struct PilotSpec {
    const uint32_t *keys_dev;       // this rank's slice of the pilot keys
    unsigned long long m_local;
    unsigned long long rank_lo, rank_hi;   // ranks in the WHOLE sample that bracket the median
    int direct;                     // 1: no pilot at all -- the window is the previous iteration's, recentred on its
                                    // exact median (median_tc_direct_ok); keys_dev / m_local / rank_* unused
};

constexpr int DSEL_WORDS = 16;      // device select block (MedianArena::dsel)

struct MedianArena {
    __half *Xh = nullptr, *Xl = nullptr;
    float *e = nullptr;                       // per-row error budget (err_budget_kernel), x_rows entries
    int64_t e_rows = 0;
    PairEntry *list = nullptr;
    uint2 *band = nullptr;
    unsigned long long *counters = nullptr;   // CNT_TOTAL u64, layout below
    unsigned long long *bins = nullptr;       // HIST_MAX_BINS + 8 (window counts, then 3 words riding on the last all-reduce)
    unsigned long long *h_pinned = nullptr;   // HIST_MAX_BINS + 8 window counts, then CNT_TOTAL counters
    int64_t x_elems = 0;
    unsigned long long list_cap = 0, band_cap = 0;
    long long direct_hits = 0, direct_misses = 0;      // statistics (stein_debug_median_direct_stats)
    cudaEvent_t ev_tail = nullptr;        // marks the D2H copies of the device-driven tail
    // device-side final select of a deferred median (tail_select_kernel): 4 u32 [status, key0, key1, -] followed by
    // 5 floats [bandwidth, h^2, 1/h^2, log2(e)/h^2, log2(e)/(2 h^2)] that the phi kernels can read without the host
    uint32_t *dsel = nullptr, *h_dsel = nullptr;
    bool fresh = false;                   // median_tc_begin ran and no sweep has used its counters yet
    // deferred tail (stein_ctx::median_defer): everything is enqueued, the host part waits in median_tc_finish_deferred
    struct Deferred {
        bool active = false;
        bool dev_valid = false;           // the device-side select of this median agrees with the host's result
        const void *owner = nullptr;      // stein_ctx::median_owner of the call that enqueued it
        uint64_t ranks[2] = {0, 0};
        int64_t d = 0;
    } deferred;
    bool split_valid = false;             // Xh / Xl / scale / budget belong to the particles of the current call
    const float *begun_X = nullptr;       // median_tc_begin_with_norms ran on these particles (median_tc_take_begun)
};

// counters block (u64 slots).  Slots [0, CNT_G1_END) are global quantities after the sweep (one
// all-reduce), [CNT_BELOW2, CNT_BELOW2 + 3) after the band filter; the rest is rank-local.
constexpr int CNT_BELOW = 0, CNT_LISTED = 1, CNT_OVERFLOW = 2, CNT_HIST = 3;
constexpr int CNT_G1_END = CNT_HIST + SW_HIST_BINS;
constexpr int CNT_LIST_LEN = CNT_G1_END;
constexpr int CNT_BELOW2 = CNT_G1_END + 1, CNT_BANDW = CNT_G1_END + 2, CNT_OVERFLOW2 = CNT_G1_END + 3;
constexpr int CNT_BAND_LEN = CNT_G1_END + 4, CNT_RMAX = CNT_G1_END + 5, CNT_HPARAMS = CNT_G1_END + 6;
constexpr int CNT_SCALE = CNT_G1_END + 7;      // 3 floats: s, s^2, 1/s^2
constexpr int CNT_WINDOW = CNT_G1_END + 9;     // device-picked pilot window: 2 floats, 2 keys, status word
constexpr int CNT_BANDP = CNT_G1_END + 12;     // device-picked band parameters (BP_* words below)
constexpr int CNT_EMAX = CNT_G1_END + 18;      // 1 float: max_i e_i
constexpr int CNT_TOTAL = CNT_G1_END + 20;


// Device-side counterpart of median_tc_tail_host: the final select of the steady-state tail, so that the kernels
// of phi can take the bandwidth from device memory while the host is still on its way to the same result (the
// host verifies: same status, same keys; engine.cu).  compute_median.py:9-15 (middle value / mean of the two
// middle values) and abstract_kernel.py:40 (h = sqrt(med / ln n)), in the host's fp32 operations.
// out: u32 [status (0 = usable), key0, key1, -], then floats [h, h^2, 1/h^2, log2(e)/h^2, log2(e)/(2 h^2)].
__global__ void __launch_bounds__(1024)
tail_select_kernel(const unsigned long long *__restrict__ cnt, const unsigned long long *__restrict__ bins,
                   int distributed, unsigned long long rank0, unsigned long long rank1, unsigned long long band_cap,
                   int even, float ln_n, uint32_t *__restrict__ out) {
    __shared__ unsigned long long wsum[32];
    __shared__ long long s_bin[2];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t < 2) s_bin[t] = -1;
    const uint32_t *wv = reinterpret_cast<const uint32_t *>(cnt + CNT_WINDOW);
    const uint32_t *hbp = reinterpret_cast<const uint32_t *>(cnt + CNT_BANDP);
    const unsigned long long *g3 = distributed ? bins + HIST_MAX_BINS + 1 : cnt + CNT_BELOW2;
    const uint32_t nbins = min(hbp[BP_NBINS], (uint32_t)HIST_MAX_BINS);
    const unsigned long long c1 = cnt[CNT_BELOW] + g3[0], band_w = g3[1];
    const bool pre = wv[4] != 0u && hbp[BP_STATUS] == 0u && g3[2] == 0ull && cnt[CNT_BAND_LEN] <= band_cap &&
                     c1 <= rank0 && rank1 < c1 + band_w && hbp[BP_SHIFT] == 0u;
    const unsigned long long rk0 = rank0 - c1, rk1 = rank1 - c1;
    unsigned long long loc[16], tot = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t b = 16u * t + k;
        loc[k] = (pre && b < nbins) ? bins[1 + b] : 0ull;
        tot += loc[k];
    }
    unsigned long long incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    unsigned long long cum = bins[0] + wsum[warp] + incl - tot;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (rk0 >= cum && rk0 < cum + loc[k]) s_bin[0] = 16 * t + k;
        if (rk1 >= cum && rk1 < cum + loc[k]) s_bin[1] = 16 * t + k;
        cum += loc[k];
    }
    __syncthreads();
    if (t != 0) return;
    uint32_t status = 1u, k0 = 0u, k1 = 0u;
    float h = 0.0f;
    if (pre && s_bin[0] >= 0 && s_bin[1] >= 0 && rk0 >= bins[0]) {
        k0 = hbp[BP_KLO] + (uint32_t)s_bin[0];
        k1 = hbp[BP_KLO] + (uint32_t)s_bin[1];
        const float m0 = key_to_float(k0), m1 = key_to_float(k1);
        const float tlo = __uint_as_float(hbp[BP_TLO]), thi = __uint_as_float(hbp[BP_THI]);
        const float wlo = __uint_as_float(wv[0]), whi = __uint_as_float(wv[1]);
        if (m0 >= tlo && m1 <= thi && m0 >= wlo && m1 <= whi) {
            const float med = even ? __fdiv_rn(__fadd_rn(m0, m1), 2.0f) : m0;
            h = __fsqrt_rn(__fdiv_rn(med, ln_n));
            if (h > 0.0f && h == h && h < INFINITY) status = 0u;
        }
    }
    const float h2 = __fmul_rn(h, h), l2e = 1.4426950408889634f;
    out[0] = status;
    out[1] = k0;
    out[2] = k1;
    out[3] = 0u;
    float *f = reinterpret_cast<float *>(out + 4);
    f[0] = h;
    f[1] = h2;
    f[2] = __fdiv_rn(1.0f, h2);
    f[3] = __fdiv_rn(l2e, h2);
    f[4] = __fdiv_rn(__fmul_rn(0.5f, l2e), h2);
}

static MedianArena g_arena;   // one per process (one GPU per process)

// What one engine's sequence of iterations remembers about its median (stein_ctx::median_owner identifies the
// sequence): several engines of a process each keep their own history, whatever order they step in.
struct HintState {
    // centre of the last pilot window (keys): successive iterations move the median only slightly
    uint32_t last_center = 0u;
    bool have_last = false;
    // pilot-less steady state: half-width (keys) of the last pilot-derived window, the exact median key of the
    // last three iterations, and how many iterations to stay on the pilot after a miss
    uint32_t last_half = 0u, med_key = 0u, prev_med_key = 0u, prev2_med_key = 0u;
    int med_keys_known = 0, direct_cooldown = 0;
};
static std::map<const void *, HintState> g_hints;
static HintState &hint_of(const stein_ctx *ctx) {
    static HintState none;
    if (ctx->median_owner == nullptr) {      // an independent call: no history in, none kept
        none = HintState();
        return none;
    }
    return g_hints[ctx->median_owner];
}
void median_tc_forget_owner(const void *owner) { g_hints.erase(owner); }

// the hint is only used within one engine's sequence of iterations
static bool hint_usable(const stein_ctx *ctx) {
    return ctx->median_owner != nullptr && hint_of(ctx).have_last;
}

static int ensure_arena(stein_ctx *ctx, int64_t rows, int64_t DP, uint64_t pairs) {
    MedianArena &A = g_arena;
    if (!A.counters) {
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&A.counters, CNT_TOTAL * 8));
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&A.bins, (HIST_MAX_BINS + 8) * 8));
        STEIN_CHECK_CUDA(ctx, cudaMallocHost(&A.h_pinned, (HIST_MAX_BINS + 8 + CNT_TOTAL) * 8));
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&A.dsel, DSEL_WORDS * 4));
        STEIN_CHECK_CUDA(ctx, cudaMallocHost(&A.h_dsel, DSEL_WORDS * 4));
        STEIN_CHECK_CUDA(ctx, cudaMemsetAsync(A.dsel, 0xff, DSEL_WORDS * 4, ctx->stream));
    }
    if (rows * DP > A.x_elems && rows * DP > 0) {
        if (A.Xh) cudaFree(A.Xh);
        if (A.Xl) cudaFree(A.Xl);
        A.Xh = A.Xl = nullptr;
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&A.Xh, rows * DP * 2));
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&A.Xl, rows * DP * 2));
        A.x_elems = rows * DP;
    }
    if (rows + 256 > A.e_rows && rows > 0) {
        if (A.e) cudaFree(A.e);
        A.e = nullptr;
        // 256 spare entries: the wide sweep's last 256-row tile may reach past the 128-padded rows (masked reads)
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&A.e, (rows + 256) * 4));
        STEIN_CHECK_CUDA(ctx, cudaMemsetAsync(A.e, 0, (rows + 256) * 4, ctx->stream));
        A.e_rows = rows + 256;
    }
    // the pilot window holds ~0.7 % of the pairs; room for 2 % (upper-triangular count)
    const unsigned long long want = pairs ? std::max<unsigned long long>(1ull << 20, pairs / 50) : 0ull;
    if (want > A.list_cap) {
        if (A.list) cudaFree(A.list);
        if (A.band) cudaFree(A.band);
        A.list = nullptr;
        A.band = nullptr;
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&A.list, want * sizeof(PairEntry)));
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&A.band, want * sizeof(uint2)));
        A.list_cap = A.band_cap = want;
    }
    return STEIN_OK;
}

// One histogram pass over a key window; the counts (nbins + 1 u64) end up in A.h_pinned.
template <int SRC>
static int window_counts(stein_ctx *ctx, const void *src, unsigned long long m_local, uint32_t key_lo,
                         uint32_t shift, uint32_t nbins, bool distributed,
                         const unsigned long long *m_dev = nullptr, const unsigned long long *extra_dev = nullptr,
                         int extra_words = 0, unsigned long long *extra_host = nullptr) {
    MedianArena &A = g_arena;
    static bool attr_set = false;
    if (!attr_set) {
        STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(window_hist_kernel<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (HIST_MAX_BINS + 1) * 4));
        attr_set = true;
    }
    STEIN_CHECK_CUDA(ctx, cudaMemsetAsync(A.bins, 0, (nbins + 1) * 8, ctx->stream));
    if (m_local) {
        const unsigned grid = (unsigned)std::min<unsigned long long>((m_local + 255) / 256, 4ull * ctx->num_sms);
        window_hist_kernel<SRC><<<grid, 256, (nbins + 1) * 4, ctx->stream>>>(src, m_local, m_dev, key_lo, shift, nbins,
                                                                             A.bins);
        STEIN_CHECK_LAUNCH(ctx);
    }
    if (distributed && ctx->has_comm) STEIN_TRY(allreduce_u64(ctx, A.bins, (int64_t)nbins + 1));
    STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(A.h_pinned, A.bins, (nbins + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    // a few more device words can ride on the same round trip
    if (extra_words)
        STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(extra_host, extra_dev, (size_t)extra_words * 8, cudaMemcpyDeviceToHost,
                                              ctx->stream));
    STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return STEIN_OK;
}

struct KeyWindow {
    uint32_t key_lo, shift, nbins;
};
// smallest shift for which [lo, hi] fits into HIST_MAX_BINS bins
static KeyWindow window_over(uint32_t lo, uint32_t hi) {
    const uint64_t span = (uint64_t)hi - lo + 1;
    uint32_t sh = 0;
    while (((span - 1) >> sh) >= (uint64_t)HIST_MAX_BINS) ++sh;
    return {lo, sh, (uint32_t)(((span - 1) >> sh) + 1)};
}

// Exact keys at the 0-based ascending weighted ranks rank[0] <= rank[1] among keys known to lie
// in [key_lo, key_hi] (keys outside: below counts, above is ignored).  One pass when the span is
// at most HIST_MAX_BINS keys (the normal case for the band), otherwise the window is narrowed.
// Returns 1 when a rank is not inside the window.
// `first_sync` (optional) describes device words to fetch with the first pass and a function
// that turns them into the ranks (the band's rank offsets are only known on the device until then).
struct FirstSync {
    const unsigned long long *m_dev = nullptr;       // device-side length of src (m_local = bound)
    const unsigned long long *extra_dev = nullptr;   // words copied to extra_host with the first pass
    unsigned long long *extra_host = nullptr;
    int extra_words = 0;
    std::function<int(uint64_t rank[2])> ranks;      // 0 = ok; anything else aborts with that code
};
template <int SRC>
static int window_select2(stein_ctx *ctx, const void *src, unsigned long long m_local, uint32_t key_lo,
                          uint32_t key_hi, uint64_t rank[2], uint32_t key_out[2], bool distributed,
                          const FirstSync *first_sync = nullptr) {
    MedianArena &A = g_arena;
    KeyWindow win[2];
    win[0] = win[1] = window_over(key_lo, key_hi);
    bool done[2] = {false, false};
    const unsigned long long *m_dev = first_sync ? first_sync->m_dev : nullptr;
    for (int iter = 0; iter < 8 && !(done[0] && done[1]); ++iter) {
        const int q0 = done[0] ? 1 : 0;
        const KeyWindow w = win[q0];
        if (iter == 0 && first_sync) {
            STEIN_TRY(window_counts<SRC>(ctx, src, m_local, w.key_lo, w.shift, w.nbins, distributed, m_dev,
                                         first_sync->extra_dev, first_sync->extra_words, first_sync->extra_host));
            const int frc = first_sync->ranks(rank);
            if (frc != 0) return frc;
        } else {
            STEIN_TRY(window_counts<SRC>(ctx, src, m_local, w.key_lo, w.shift, w.nbins, distributed, m_dev));
        }
        for (int q = q0; q < 2; ++q) {
            if (done[q] || win[q].key_lo != w.key_lo || win[q].shift != w.shift || win[q].nbins != w.nbins) continue;
            KeyWindow nw = w;
            const int rc = stein_median_narrow(reinterpret_cast<const uint64_t *>(A.h_pinned), w.key_lo, w.shift, w.nbins,
                                               rank[q], &key_out[q], &nw.key_lo, &nw.shift, &nw.nbins);
            if (rc == 1) done[q] = true;
            else if (rc == 0) win[q] = nw;
            else return 1;
        }
    }
    return (done[0] && done[1]) ? STEIN_OK : 1;
}

// Window keys from the pilot sample: [lo, hi] brackets the sample quantiles at rank_lo / rank_hi.
// Generic route: two histogram passes (14 bits of the full key range, then the bins of the two
// ranks split 16384 ways).  keys_dev / m are this rank's slice of the sample; the counts are
// all-reduced, so every rank takes the same decisions.
// bins of the two ranks in the counts of the last window_counts call; false when a rank lies outside
static bool locate_two(const KeyWindow &w, uint64_t rank_lo, uint64_t rank_hi, uint64_t *lo, uint64_t *hi) {
    MedianArena &A = g_arena;
    uint64_t cum = A.h_pinned[0];
    if (rank_lo < cum) return false;
    int64_t blo = -1, bhi = -1;
    for (uint32_t b = 0; b < w.nbins; ++b) {
        const uint64_t c = A.h_pinned[1 + b];
        if (blo < 0 && cum + c > rank_lo) blo = b;
        if (bhi < 0 && cum + c > rank_hi) bhi = b;
        cum += c;
    }
    if (blo < 0 || bhi < 0) return false;
    *lo = (uint64_t)w.key_lo + ((uint64_t)blo << w.shift);
    *hi = std::min<uint64_t>((uint64_t)w.key_lo + (((uint64_t)bhi + 1) << w.shift) - 1, 0xffffffffull);
    return true;
}

int pilot_window(stein_ctx *ctx, const uint32_t *keys_dev, int64_t m, uint64_t rank_lo, uint64_t rank_hi,
                 uint32_t *lo_key, uint32_t *hi_key) {
    if (!g_arena.counters) STEIN_TRY(ensure_arena(ctx, 0, 0, 0));
    // Successive SVGD iterations move the median by a fraction of a percent: first try ONE pass
    // over +-2^20 keys (about +-9 % in value, 128 keys per bin) around the last window.  The
    // result is only used when both ranks fall inside; otherwise the two generic passes run.
    HintState &H = hint_of(ctx);
    uint32_t &last_center = H.last_center;
    bool &have_last = H.have_last;
    uint64_t lo = 0, hi = 0;
    bool found = false;
    if (hint_usable(ctx)) {
        const uint32_t c = last_center, half = 1u << 20;
        const KeyWindow w = window_over(c > half ? c - half : 0u, c < 0xffffffffu - half ? c + half : 0xffffffffu);
        STEIN_TRY(window_counts<2>(ctx, keys_dev, (unsigned long long)m, w.key_lo, w.shift, w.nbins, true));
        found = locate_two(w, rank_lo, rank_hi, &lo, &hi);
    }
    if (!found) {
        KeyWindow w = {0u, 18u, (uint32_t)HIST_MAX_BINS};
        for (int pass = 0; pass < 2; ++pass) {
            STEIN_TRY(window_counts<2>(ctx, keys_dev, (unsigned long long)m, w.key_lo, w.shift, w.nbins, true));
            if (!locate_two(w, rank_lo, rank_hi, &lo, &hi)) {
                have_last = false;
                return 1;
            }
            if (w.shift == 0) break;
            w = window_over((uint32_t)lo, (uint32_t)hi);
        }
    }
    *lo_key = (uint32_t)lo;
    *hi_key = (uint32_t)hi;
    last_center = (uint32_t)((lo + hi) / 2);
    have_last = true;
    return STEIN_OK;
}

// ld: leading dimension of the particle matrix (its zero pad columns count as coordinates)
bool median_tc_supported(int64_t n, int64_t ld) {
    return (ld == 128 || ld == 256 || ld == 512 || ld == 768 || ld == 1024) && (uint64_t)n * (uint64_t)n >= (1ull << 24) &&
           n < (1ll << 31);
}

// returns STEIN_OK with keys filled, 1 = "not bracketed, use the FFMA route", <0 = error
bool median_tc_has_hint(const stein_ctx *ctx) { return hint_usable(ctx); }

// Successive SVGD iterations move the median smoothly: by a small fraction of the pilot window (which is +-3.5
// sigma of a 2^20-pair sample quantile wide), or -- early in a run, while the cloud contracts or expands -- by a
// larger but steady amount per step.  The next window is therefore the last one recentred on the EXTRAPOLATED
// exact median, last + (last - previous), as long as that prediction was good for the last step: the second
// difference of the last three exact medians is below an eighth of the window's half-width.  No pilot sample, no
// pilot histogram, no all-reduce of it.  Every consumer of the window still checks that the rank is bracketed;
// a miss falls back to the pilot route and keeps it for a few iterations.
static int64_t direct_step(const HintState &A) { return (int64_t)A.med_key - (int64_t)A.prev_med_key; }
static uint32_t direct_center(const HintState &A) {
    const int64_t c = (int64_t)A.med_key + direct_step(A);
    return (uint32_t)std::min<int64_t>(std::max<int64_t>(c, 0), 0xffffffffll);
}
bool median_tc_direct_ok(const stein_ctx *ctx) {
    if (ctx->median_owner == nullptr) return false;
    const HintState &A = hint_of(ctx);
    if (getenv("STEIN_MEDIAN_DEBUG"))
        fprintf(stderr, "[stein] direct_ok: hint %d known %d half %u cooldown %d keys %u %u %u\n", (int)hint_usable(ctx),
                A.med_keys_known, A.last_half, A.direct_cooldown, A.med_key, A.prev_med_key, A.prev2_med_key);
    if (!hint_usable(ctx) || A.med_keys_known < 3 || A.last_half == 0u || A.direct_cooldown > 0) return false;
    if (const char *e = getenv("STEIN_MEDIAN_PILOTLESS"))
        if (e[0] == '0') return false;
    const int64_t prev_step = (int64_t)A.prev_med_key - (int64_t)A.prev2_med_key;
    const uint64_t second = (uint64_t)std::llabs(direct_step(A) - prev_step);
    // (a step of many window widths is not a slow drift, however steady)
    return second * 8u < A.last_half && (uint64_t)std::llabs(direct_step(A)) < 4ull * A.last_half;
}
void median_tc_count_direct_hit(void) { ++g_arena.direct_hits; }
// bookkeeping after a median call of an engine sequence (keys of the two middle values)
void median_tc_note_result(const stein_ctx *ctx, uint32_t k0, uint32_t k1, bool direct_missed) {
    if (ctx->median_owner == nullptr) return;
    HintState &A = hint_of(ctx);
    A.prev2_med_key = A.prev_med_key;
    A.prev_med_key = A.med_key;
    A.med_key = (uint32_t)(((uint64_t)k0 + k1) / 2);
    A.med_keys_known = std::min(A.med_keys_known + 1, 3);
    if (direct_missed) {
        A.direct_cooldown = 8;
        ++g_arena.direct_misses;
    } else if (A.direct_cooldown > 0) {
        --A.direct_cooldown;
    }
}

// window word block for the sweep: [wlo, whi] floats, [klo, khi] keys, ok flag
__global__ void set_window_kernel(uint32_t klo, uint32_t khi, uint32_t *__restrict__ out) {
    out[0] = __float_as_uint(key_to_float(klo));
    out[1] = __float_as_uint(key_to_float(khi));
    out[2] = klo;
    out[3] = khi;
    out[4] = 1u;
}

static SweepParams g_deferred_params;     // window words of the deferred sweep (MedianArena::deferred)
static int median_tc_tail_host(stein_ctx *ctx, int64_t d, const uint64_t ranks[2], SweepParams *p, uint32_t keys_out[2]);

// Steady-state tail of the tensor-core route (spec != NULL in median_tc): everything after the
// sweep is chained on the device -- band thresholds (pick_band_kernel), band filter, exact
// distances, histogram of the exact keys over a device-picked window -- and the host reads the
// counters and that histogram in ONE round trip.  Same formulas as the host-driven path below.
// Returns like median_tc.
static int median_tc_device_tail(stein_ctx *ctx, const float *X, const float *r, int64_t n, int64_t d, int64_t ld,
                                 const uint64_t ranks[2], SweepParams *p, uint32_t keys_out[2]) {
    MedianArena &A = g_arena;
    const int world = ctx->has_comm ? ctx->comm.world : 1;
    uint32_t *bp = reinterpret_cast<uint32_t *>(A.counters + CNT_BANDP);
    int *d_overflow2 = reinterpret_cast<int *>(A.counters + CNT_OVERFLOW2);
    pick_band_kernel<<<1, 1024, 0, ctx->stream>>>(A.counters, CNT_BELOW, CNT_LISTED, CNT_OVERFLOW, CNT_HIST, CNT_RMAX,
                                                 CNT_EMAX, CNT_HPARAMS, ranks[0], ranks[1], eps_abs_coeff(d),
                                                 (uint32_t)HIST_MAX_BINS, bp);
    STEIN_CHECK_LAUNCH(ctx);
    band_filter_kernel<<<16 * ctx->num_sms, BF_THREADS, 0, ctx->stream>>>(
        A.list, A.list_cap, A.e, 0.f, 0.f, 0.f, A.counters + CNT_BELOW2, A.counters + CNT_BANDW,
        A.counters + CNT_BAND_LEN, A.band, A.band_cap, d_overflow2, A.counters + CNT_LIST_LEN, bp);
    STEIN_CHECK_LAUNCH(ctx);
    trace_mark(ctx, "median:pick band + filter");
    // (the three global words of the filter -- below, band weight, overflow -- are only read by the host:
    //  they ride on the all-reduce of the band histogram below instead of taking one of their own)
    STEIN_TRY(launch_pair_chain<0>(ctx, A.band, A.band_cap, X, r, n, ld, 0, A.counters + CNT_BAND_LEN));
    trace_mark(ctx, "median:exact band");
    static bool attr_set = false;
    if (!attr_set) {
        STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(window_hist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (HIST_MAX_BINS + 1) * 4));
        attr_set = true;
    }
    STEIN_CHECK_CUDA(ctx, cudaMemsetAsync(A.bins, 0, (HIST_MAX_BINS + 1) * 8, ctx->stream));
    window_hist_kernel<1><<<4 * ctx->num_sms, 256, (HIST_MAX_BINS + 1) * 4, ctx->stream>>>(
        A.band, A.band_cap, A.counters + CNT_BAND_LEN, 0u, 0u, (uint32_t)HIST_MAX_BINS, A.bins, bp);
    STEIN_CHECK_LAUNCH(ctx);
    if (world > 1) {
        STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(A.bins + HIST_MAX_BINS + 1, A.counters + CNT_BELOW2, 3 * 8,
                                              cudaMemcpyDeviceToDevice, ctx->stream));
        STEIN_TRY(allreduce_u64(ctx, A.bins, HIST_MAX_BINS + 4));
    }
    unsigned long long *h = A.h_pinned + HIST_MAX_BINS + 8;
    STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(A.h_pinned, A.bins, (HIST_MAX_BINS + 4) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(h, A.counters, CNT_TOTAL * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->median_defer) {
        const uint64_t dim = (uint64_t)n * (uint64_t)n;
        tail_select_kernel<<<1, 1024, 0, ctx->stream>>>(A.counters, A.bins, world > 1 ? 1 : 0, ranks[0], ranks[1],
                                                       A.band_cap, dim % 2 == 0 ? 1 : 0, (float)log((double)n), A.dsel);
        STEIN_CHECK_LAUNCH(ctx);
        STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(A.h_dsel, A.dsel, DSEL_WORDS * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    trace_mark(ctx, "median:band histogram + D2H");
    if (ctx->median_defer) {
        // the caller collects the result later (median_tc_finish_deferred): nothing below needs the host now
        if (!A.ev_tail) STEIN_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&A.ev_tail, cudaEventDisableTiming));
        STEIN_CHECK_CUDA(ctx, cudaEventRecord(A.ev_tail, ctx->stream));
        A.deferred.active = true;
        A.deferred.dev_valid = false;
        A.deferred.owner = ctx->median_owner;
        A.deferred.ranks[0] = ranks[0];
        A.deferred.ranks[1] = ranks[1];
        A.deferred.d = d;
        g_deferred_params = *p;
        return 3;
    }
    if (ctx->presync_fn) {
        // the caller has bandwidth-independent work for this stream: queue it behind the copies and
        // wait for the copies only, so the GPU is busy during the host's part of the round trip
        if (!A.ev_tail) STEIN_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&A.ev_tail, cudaEventDisableTiming));
        STEIN_CHECK_CUDA(ctx, cudaEventRecord(A.ev_tail, ctx->stream));
        const int hrc = ctx->presync_fn(ctx->presync_arg);
        STEIN_CHECK_CUDA(ctx, cudaEventSynchronize(A.ev_tail));
        if (hrc != STEIN_OK) return hrc < 0 ? hrc : STEIN_ERR_INVALID;
    } else {
        STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }

    return median_tc_tail_host(ctx, d, ranks, p, keys_out);
}

// Host half of the device-driven tail: reads the counters / histogram the tail copied to pinned memory (the
// caller has waited for the copies).  Same return values as median_tc.
static int median_tc_tail_host(stein_ctx *ctx, int64_t d, const uint64_t ranks[2], SweepParams *p, uint32_t keys_out[2]) {
    MedianArena &A = g_arena;
    (void)d;
    const int world = ctx->has_comm ? ctx->comm.world : 1;
    unsigned long long *h = A.h_pinned + HIST_MAX_BINS + 8;
    const uint32_t *wv = reinterpret_cast<const uint32_t *>(h + CNT_WINDOW);
    HintState &H = hint_of(ctx);
    if (!wv[4]) {                // the pilot ranks were not inside the histogram around the old window
        H.have_last = false;
        return 2;
    }
    memcpy(&p->wlo, &wv[0], 4);
    memcpy(&p->whi, &wv[1], 4);
    H.last_center = (uint32_t)(((uint64_t)wv[2] + wv[3]) / 2);
    if (!p->direct_window) H.last_half = (wv[3] - wv[2]) / 2u + 1u;
    const uint32_t *hbp = reinterpret_cast<const uint32_t *>(h + CNT_BANDP);
    // global (below2, band weight, overflow2): all-reduced at the tail of the histogram on sharded runs
    const unsigned long long *g3 = world > 1 ? A.h_pinned + HIST_MAX_BINS + 1 : h + CNT_BELOW2;
    if (hbp[BP_STATUS] != 0u || g3[2] || h[CNT_BAND_LEN] > A.band_cap) return 1;
    float tlo, thi;
    memcpy(&tlo, &hbp[BP_TLO], 4);
    memcpy(&thi, &hbp[BP_THI], 4);
    const uint64_t c1 = h[CNT_BELOW] + g3[0], band_w = g3[1];
    if (!(c1 <= ranks[0] && ranks[1] < c1 + band_w)) return 1;
    uint64_t rk[2] = {ranks[0] - c1, ranks[1] - c1};
    uint32_t k01[2] = {0u, 0u};
    bool done = true;
    for (int q = 0; q < 2; ++q) {
        uint32_t lo2, sh2, nb2;
        const int rc = stein_median_narrow(reinterpret_cast<const uint64_t *>(A.h_pinned), hbp[BP_KLO], hbp[BP_SHIFT],
                                           hbp[BP_NBINS], rk[q], &k01[q], &lo2, &sh2, &nb2);
        if (rc < 0) return 1;
        if (rc == 0) done = false;       // window wider than HIST_MAX_BINS keys: refine below
    }
    if (!done) {                 // rare: host-driven passes over the same band
        const uint64_t khi = (uint64_t)hbp[BP_KLO] + ((uint64_t)hbp[BP_NBINS] << hbp[BP_SHIFT]) - 1;
        const int rc = window_select2<1>(ctx, A.band, std::min<unsigned long long>(h[CNT_BAND_LEN], A.band_cap),
                                         hbp[BP_KLO], (uint32_t)std::min<uint64_t>(khi, 0xffffffffull), rk, k01, true);
        if (rc != STEIN_OK) return rc;
    }
    // the selected values must lie where the certainty argument holds
    const float m0 = key_to_float(k01[0]), m1 = key_to_float(k01[1]);
    if (!(m0 >= tlo && m1 <= thi && m0 >= p->wlo && m1 <= p->whi)) return 1;
    keys_out[0] = k01[0];
    keys_out[1] = k01[1];
    return STEIN_OK;
}



// First stage of the route: zero the counters, largest row norm and scale, FP16 split of s X.
int median_tc_begin(stein_ctx *ctx, const float *X, const float *r, int64_t n, int64_t ld) {
    const int64_t rows = stein_rows_padded(n), T = (n + TILE - 1) / TILE;
    const uint64_t pairs = (uint64_t)T * (T + 1) / 2 * TILE * TILE;
    STEIN_TRY(ensure_arena(ctx, rows, ld, pairs));
    MedianArena &A = g_arena;
    A.fresh = false;
    STEIN_CHECK_CUDA(ctx, cudaMemsetAsync(A.counters, 0, CNT_TOTAL * 8, ctx->stream));
    float *d_scale = reinterpret_cast<float *>(A.counters + CNT_SCALE);
    err_budget_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, ctx->stream>>>(X, rows, n, ld, A.e);
    STEIN_CHECK_LAUNCH(ctx);
    max_kernel<<<1, 1024, 0, ctx->stream>>>(r, A.e, n, reinterpret_cast<float *>(A.counters + CNT_RMAX),
                                           reinterpret_cast<float *>(A.counters + CNT_EMAX), d_scale);
    STEIN_CHECK_LAUNCH(ctx);
    const int64_t count4 = rows * ld / 4;
    split_f16_kernel<<<(unsigned)((count4 + 255) / 256), 256, 0, ctx->stream>>>(X, count4, d_scale, A.Xh, A.Xl);
    STEIN_CHECK_LAUNCH(ctx);
    A.fresh = true;
    A.split_valid = true;
    trace_mark(ctx, "median:budget + max + split");
    return STEIN_OK;
}
void median_tc_reset(void) {
    g_arena.fresh = false;
    g_arena.split_valid = false;
    g_arena.begun_X = nullptr;
}
// median_tc_begin that also produces the row norms (stein_row_norms' bits) for `rows_r` rows of X: one read of the
// particles instead of two (engine.cu, start of an iteration).  The next median call on the same X finds its first
// stage done (median_tc_take_begun).
int median_tc_begin_with_norms(stein_ctx *ctx, const float *X, float *r, int64_t rows_r, int64_t n, int64_t ld) {
    const int64_t rows = stein_rows_padded(n), T = (n + TILE - 1) / TILE;
    const uint64_t pairs = (uint64_t)T * (T + 1) / 2 * TILE * TILE;
    STEIN_TRY(ensure_arena(ctx, rows, ld, pairs));
    MedianArena &A = g_arena;
    median_tc_reset();
    A.deferred.active = false;
    STEIN_CHECK_CUDA(ctx, cudaMemsetAsync(A.counters, 0, CNT_TOTAL * 8, ctx->stream));
    float *d_scale = reinterpret_cast<float *>(A.counters + CNT_SCALE);
    const int64_t rows_k = std::max(rows_r, rows);
    x_stats_kernel<<<(unsigned)((rows_k + XS_ROWS - 1) / XS_ROWS), 256, 0, ctx->stream>>>(X, rows_r, rows, n, ld, r, A.e);
    STEIN_CHECK_LAUNCH(ctx);
    max_kernel<<<1, 1024, 0, ctx->stream>>>(r, A.e, n, reinterpret_cast<float *>(A.counters + CNT_RMAX),
                                           reinterpret_cast<float *>(A.counters + CNT_EMAX), d_scale);
    STEIN_CHECK_LAUNCH(ctx);
    const int64_t count4 = rows * ld / 4;
    split_f16_kernel<<<(unsigned)((count4 + 255) / 256), 256, 0, ctx->stream>>>(X, count4, d_scale, A.Xh, A.Xl);
    STEIN_CHECK_LAUNCH(ctx);
    A.fresh = true;
    A.split_valid = true;
    A.begun_X = X;
    trace_mark(ctx, "head:norms + budget + max + split");
    return STEIN_OK;
}
// true (once) when median_tc_begin_with_norms ran on exactly these particles and nothing has used the arena since
bool median_tc_take_begun(const float *X) {
    const bool hit = g_arena.begun_X != nullptr && g_arena.begun_X == X && g_arena.fresh;
    g_arena.begun_X = nullptr;
    return hit;
}
bool median_tc_deferred_pending(const void *owner) { return g_arena.deferred.active && g_arena.deferred.owner == owner; }
void median_tc_cancel_deferred(void) { g_arena.deferred.active = false; }
// Collects a median whose device part was enqueued with stein_ctx::median_defer set: waits for the D2H copies of
// the tail and runs its host half.  Returns like median_tc (0 keys found, 1 / 2 take the other routes, < 0 error).
int median_tc_finish_deferred(stein_ctx *ctx, uint32_t keys_out[2]) {
    MedianArena &A = g_arena;
    if (!A.deferred.active) return fail(ctx, STEIN_ERR_INTERNAL, "no deferred median to collect");
    A.deferred.active = false;
    STEIN_CHECK_CUDA(ctx, cudaEventSynchronize(A.ev_tail));
    const int rc = median_tc_tail_host(ctx, A.deferred.d, A.deferred.ranks, &g_deferred_params, keys_out);
    A.deferred.dev_valid = rc == STEIN_OK && A.h_dsel[0] == 0u && A.h_dsel[1] == keys_out[0] && A.h_dsel[2] == keys_out[1];
    return rc;
}
// The bandwidth block of the device-side select of the deferred median (5 floats, see tail_select_kernel); valid
// for the kernels that follow the median on the stream.  NULL before the first deferred median.
const float *median_tc_device_bandwidth(void) {
    return g_arena.dsel ? reinterpret_cast<const float *>(g_arena.dsel + 4) : nullptr;
}
// after median_tc_finish_deferred: did the device-side select produce the keys the host accepted, and its h
bool median_tc_device_select_valid(float *bandwidth) {
    if (bandwidth) memcpy(bandwidth, g_arena.h_dsel + 4, 4);
    return g_arena.deferred.dev_valid;
}

// Pilot keys of samples [s0, s0 + m) from the FP16 hi array (needs median_tc_begin on this X).
int median_tc_pilot(stein_ctx *ctx, uint32_t *keys_dev, unsigned long long m, const float *r, int64_t n, int64_t ld,
                    uint64_t seed) {
    MedianArena &A = g_arena;
    if (!A.split_valid) return fail(ctx, STEIN_ERR_INTERNAL, "median_tc_pilot without median_tc_begin");
    if (m == 0) return STEIN_OK;
    const unsigned long long ngroups = (m + 31ull) / 32ull;
    const unsigned grid = (unsigned)std::min<unsigned long long>((ngroups + 7) / 8, 8ull * ctx->num_sms);
    pilot_h16_kernel<<<grid, 256, 0, ctx->stream>>>(keys_dev, m, reinterpret_cast<const uint4 *>(A.Xh), r,
                                                   reinterpret_cast<const float *>(A.counters + CNT_SCALE), n,
                                                   (int)(ld / 8), seed);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

// spec != NULL: the window is not given but derived ON THE DEVICE from one histogram pass over the
// pilot keys around the previous window (no host round trip between pilot and sweep).  Returns 2
// when the pilot ranks fell outside that histogram (the caller then takes the generic route).
int median_tc(stein_ctx *ctx, const float *X, const float *r, int64_t n, int64_t d, int64_t ld,
              const uint64_t ranks[2], uint32_t win_lo_key, uint32_t win_hi_key, uint32_t keys_out[2],
              int *sweeps, const PilotSpec *spec) {
    const int64_t rows = stein_rows_padded(n), DP = ld;
    const int64_t T = (n + TILE - 1) / TILE;
    const uint64_t pairs = (uint64_t)T * (T + 1) / 2 * TILE * TILE;
    STEIN_TRY(ensure_arena(ctx, rows, DP, pairs));
    MedianArena &A = g_arena;
    const int world = ctx->has_comm ? ctx->comm.world : 1, rank = ctx->has_comm ? ctx->comm.rank : 0;
    // CTA-pair sweep unless the single-CTA kernel is asked for (STEIN_MEDIAN_TC1, tests / comparison)
    const bool wide = DP > 64 * SW2_KB;                  // more than 256 coordinates: K-streaming sweep (Sweep3Policy)
    const bool pair = ctx->median_impl != STEIN_MEDIAN_TC1 && ctx->num_sms >= 2 && !wide;
    const int64_t T2 = (T + 1) / 2;
    // wide sweep: windows of 256 x 256 tiles over the upper triangle (Sweep3Policy); the ranks split the window list
    std::vector<int2> wins;
    int win_a = 1, win_b = 1;
    if (wide) {
        const int G = std::max(1, ctx->num_sms / 2);
        win_a = std::max(1, (int)std::floor(std::sqrt((double)G)));
        while (win_a > 1 && (G / win_a) * win_a < (G / (win_a + 1)) * (win_a + 1)) ++win_a;      // 74 clusters: 9 x 8
        if ((G / (win_a + 1)) * (win_a + 1) > (G / win_a) * win_a) ++win_a;
        win_b = std::max(1, G / win_a);
        const int64_t nwi = (T2 + win_a - 1) / win_a, nwj = (T2 + win_b - 1) / win_b;
        for (int64_t wi = 0; wi < nwi; ++wi)
            for (int64_t wj = 0; wj < nwj; ++wj)
                if ((wj + 1) * win_b - 1 >= wi * win_a) wins.push_back(make_int2((int)wi, (int)wj));   // touches tj >= ti
    }
    const int64_t ntiles = wide ? (int64_t)wins.size() : (pair ? num_pair_tiles(T) : T * (T + 1) / 2);
    int64_t t0 = ntiles * rank / world, t1 = ntiles * (rank + 1) / world;
    // CTA-pair sweep: cost-balanced chunks (sweep_chunk_bounds), cached per shape
    SweepBounds bnd;
    bnd.n = 0;
    const int ncl_full = ctx->num_sms / 2;
    if (pair && ncl_full <= SW2_MAX_CLUSTERS && ntiles >= 4ll * world * ncl_full && ntiles < (1ll << 31)) {
        static std::vector<long long> cached;
        static int64_t c_T = -1;
        static int c_parts = -1;
        static double c_beta = -1.0;
        double beta = 1.5;      // measured at config D: 0 (even tile split) 2.53-2.59 ms, 1-2: 2.45 ms, 3: 2.49 ms
        if (const char *env = getenv("STEIN_SWEEP_BETA")) beta = atof(env);
        if (beta > 0.0) {
            if (c_T != T || c_parts != world * ncl_full || c_beta != beta) {
                cached = sweep_chunk_bounds(T, world * ncl_full, beta);
                c_T = T;
                c_parts = world * ncl_full;
                c_beta = beta;
            }
            bnd.n = ncl_full;
            for (int c = 0; c <= ncl_full; ++c) bnd.b[c] = (int)cached[(size_t)rank * ncl_full + c];
            t0 = bnd.b[0];
            t1 = bnd.b[ncl_full];
        }
    }

    // counters zeroed, largest row norm, scale, FP16 split: already there when the caller ran
    // median_tc_begin for its pilot and this is the first sweep since
    if (!A.fresh) STEIN_TRY(median_tc_begin(ctx, X, r, n, ld));
    A.fresh = false;
    float *d_rmax = reinterpret_cast<float *>(A.counters + CNT_RMAX);
    float *d_scale = reinterpret_cast<float *>(A.counters + CNT_SCALE);
    int *d_overflow = reinterpret_cast<int *>(A.counters + CNT_OVERFLOW);
    int *d_overflow2 = reinterpret_cast<int *>(A.counters + CNT_OVERFLOW2);

    if (spec && spec->direct) {
        // pilot-less: the last window, recentred on the extrapolated exact median
        const uint32_t c = direct_center(hint_of(ctx)), half = hint_of(ctx).last_half;
        set_window_kernel<<<1, 1, 0, ctx->stream>>>(c > half ? c - half : 0u, c < 0xffffffffu - half ? c + half : 0xffffffffu,
                                                   reinterpret_cast<uint32_t *>(A.counters + CNT_WINDOW));
        STEIN_CHECK_LAUNCH(ctx);
    } else if (spec) {
        // one histogram pass over +-2^20 keys around the last window, then the device picks the bins
        const uint32_t c = hint_of(ctx).last_center, half = 1u << 20;
        const KeyWindow w = window_over(c > half ? c - half : 0u, c < 0xffffffffu - half ? c + half : 0xffffffffu);
        static bool attr2 = false;
        if (!attr2) {
            STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(window_hist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (HIST_MAX_BINS + 1) * 4));
            attr2 = true;
        }
        STEIN_CHECK_CUDA(ctx, cudaMemsetAsync(A.bins, 0, (w.nbins + 1) * 8, ctx->stream));
        if (spec->m_local) {
            const unsigned grid = (unsigned)std::min<unsigned long long>((spec->m_local + 255) / 256, 4ull * ctx->num_sms);
            window_hist_kernel<2><<<grid, 256, (w.nbins + 1) * 4, ctx->stream>>>(spec->keys_dev, spec->m_local, nullptr,
                                                                               w.key_lo, w.shift, w.nbins, A.bins);
            STEIN_CHECK_LAUNCH(ctx);
        }
        if (world > 1) STEIN_TRY(allreduce_u64(ctx, A.bins, (int64_t)w.nbins + 1));
        pilot_pick_kernel<<<1, 1024, 0, ctx->stream>>>(A.bins, w.key_lo, w.shift, w.nbins, spec->rank_lo, spec->rank_hi,
                                                      reinterpret_cast<uint32_t *>(A.counters + CNT_WINDOW));
        STEIN_CHECK_LAUNCH(ctx);
    }

    CUtensorMap mapXh, mapXl;
    STEIN_TRY(make_tensor_map_2d(ctx, &mapXh, A.Xh, 2, (uint64_t)DP, (uint64_t)rows, (uint64_t)DP * 2, 128));
    STEIN_TRY(make_tensor_map_2d(ctx, &mapXl, A.Xl, 2, (uint64_t)DP, (uint64_t)rows, (uint64_t)DP * 2, 128));
    SweepParams p{};
    p.T = (int)T;
    p.t_begin = t0;
    p.t_end = t1;
    p.kblocks = (int)(DP / 64);
    p.n = n;
    p.r = r;
    p.Xh = reinterpret_cast<const uint32_t *>(A.Xh);
    p.Xl = reinterpret_cast<const uint32_t *>(A.Xl);
    p.wlo = key_to_float(win_lo_key);
    p.whi = key_to_float(win_hi_key);
    p.window = spec ? reinterpret_cast<const float *>(A.counters + CNT_WINDOW) : nullptr;
    p.direct_window = (spec && spec->direct) ? 1 : 0;
    p.e = A.e;
    p.emax = reinterpret_cast<const float *>(A.counters + CNT_EMAX);
    p.rmax = d_rmax;
    p.scale = d_scale;
    p.hist = A.counters + CNT_HIST;
    p.cnt_below = A.counters + CNT_BELOW;
    p.cnt_listed = A.counters + CNT_LISTED;
    p.cnt_len = A.counters + CNT_LIST_LEN;
    p.hparams_out = reinterpret_cast<float *>(A.counters + CNT_HPARAMS);
    p.overflow = d_overflow;
    p.list = A.list;
    p.list_cap = A.list_cap;
    const size_t tail_bytes = SW_TAIL_BYTES;
    const size_t smem1 = 1024 + (size_t)SW_STAGES * SW_UNIT_BYTES + tail_bytes;
    const size_t smem2 = 1024 + (size_t)(SW2_KB + SW2_STAGES) * SW_UNIT_BYTES + tail_bytes;
    static bool attr_set = false;
    if (!attr_set) {
        STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(sweep_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)smem1));
        STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(sweep2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)smem2));
        attr_set = true;
    }
    hparams_kernel<<<1, 1, 0, ctx->stream>>>(p.window, p.wlo, p.whi, p.emax, reinterpret_cast<float *>(A.counters + CNT_HPARAMS));
    STEIN_CHECK_LAUNCH(ctx);
    if (t1 > t0 && wide) {
        RegionTimer timer(ctx, STEIN_REGION_SWEEP);
        pg::Maps maps;
        memset(&maps, 0, sizeof(maps));
        // table entries (a, b): (0, 0) hi.hi, (1, 1) lo.hi, (0, 2) hi.lo
        const void *am[3] = {A.Xh, A.Xl, A.Xh}, *bm[3] = {A.Xh, A.Xh, A.Xl};
        for (int m = 0; m < 3; ++m) {
            STEIN_TRY(make_tensor_map_2d(ctx, &maps.a[m], am[m], 2, (uint64_t)DP, (uint64_t)rows, (uint64_t)DP * 2, 128));
            STEIN_TRY(make_tensor_map_2d(ctx, &maps.b[m], bm[m], 2, (uint64_t)DP, (uint64_t)rows, (uint64_t)DP * 2, 128));
        }
        Sweep3Policy::Params sp{};
        sp.n_stage = 6;
        for (int h = 0; h < 2; ++h) {
            sp.table[3 * h + 0] = {0, 0, 64 * h, 0};
            sp.table[3 * h + 1] = {1, 1, 64 * h, 0};
            sp.table[3 * h + 2] = {0, 2, 64 * h, 0};
        }
        sp.groups_per_unit = (int)(DP / 128);
        sp.units_per_tile = 1;
        sp.ka0 = sp.kb0 = 0;
        sp.a_row0 = sp.b_row0 = 0;
        sp.route = nullptr;
        sp.my_route = 0;
        sp.sp = p;
        sp.T2 = (int)T2;
        sp.win_a = win_a;
        sp.win_b = win_b;
        // this rank's windows, in a small library-owned device buffer (re-uploaded: a few KB, pageable source)
        static int2 *d_wins = nullptr;
        static size_t d_wins_cap = 0;
        const size_t nw = (size_t)(t1 - t0);
        if (nw > d_wins_cap) {
            STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (d_wins) cudaFree(d_wins);
            d_wins = nullptr;
            STEIN_CHECK_CUDA(ctx, cudaMalloc(&d_wins, nw * sizeof(int2)));
            d_wins_cap = nw;
        }
        STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(d_wins, wins.data() + t0, nw * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
        sp.wins = d_wins;
        sp.nwins = (int)nw;
        const size_t smem3 = pg::smem_bytes<Sweep3Policy::STAGES>(Sweep3Policy::TAIL_BYTES);
        static bool attr3 = false;
        if (!attr3) {
            STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(pg::panel_gemm_kernel<Sweep3Policy>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            attr3 = true;
        }
        const int clusters = win_a * win_b;
        pg::panel_gemm_kernel<Sweep3Policy><<<2 * clusters, pg::THREADS, smem3, ctx->stream>>>(maps, sp);
        STEIN_CHECK_LAUNCH(ctx);
    } else if (t1 > t0) {
        RegionTimer timer(ctx, STEIN_REGION_SWEEP);
        if (pair) {
            CUtensorMap mapXh64, mapXl64;
            STEIN_TRY(make_tensor_map_2d(ctx, &mapXh64, A.Xh, 2, (uint64_t)DP, (uint64_t)rows, (uint64_t)DP * 2, 64));
            STEIN_TRY(make_tensor_map_2d(ctx, &mapXl64, A.Xl, 2, (uint64_t)DP, (uint64_t)rows, (uint64_t)DP * 2, 64));
            const int clusters = bnd.n ? bnd.n : (int)std::min<int64_t>(ctx->num_sms / 2, t1 - t0);
            sweep2_tc_kernel<<<2 * clusters, SW2_THREADS, smem2, ctx->stream>>>(mapXh, mapXl, mapXh64, mapXl64, p, bnd);
        } else {
            const int grid = (int)std::min<int64_t>(ctx->num_sms, t1 - t0);
            sweep_tc_kernel<<<grid, SW_THREADS, smem1, ctx->stream>>>(mapXh, mapXl, p);
        }
        STEIN_CHECK_LAUNCH(ctx);
    }
    if (sweeps) *sweeps += 1;
    trace_mark(ctx, "median:sweep");
    // below / listed / overflow / histogram of the listed D~ become global quantities with ONE
    // all-reduce; the list itself (and its length) stays rank-local
    if (world > 1) STEIN_TRY(allreduce_u64(ctx, A.counters, CNT_G1_END));
    if (spec) return median_tc_device_tail(ctx, X, r, n, d, ld, ranks, &p, keys_out);
    unsigned long long *h = A.h_pinned + HIST_MAX_BINS + 8;
    STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(h, A.counters, CNT_TOTAL * 8, cudaMemcpyDeviceToHost, ctx->stream));
    STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const unsigned long long list_len = std::min<unsigned long long>(h[CNT_LIST_LEN], A.list_cap);
    const float rmax = *reinterpret_cast<float *>(h + CNT_RMAX), emax = *reinterpret_cast<float *>(h + CNT_EMAX);
    const float hlo = reinterpret_cast<float *>(h + CNT_HPARAMS)[0], hscale = reinterpret_cast<float *>(h + CNT_HPARAMS)[1];
    const uint64_t below = h[CNT_BELOW], listed = h[CNT_LISTED];
    if (h[CNT_OVERFLOW]) return 1;
    if (!(below <= ranks[0] && ranks[1] < below + listed)) return 1;   // pilot window missed

    // bin of the D~ value at the lower target rank (the bins are monotone in D~)
    uint64_t cum = below;
    int bstar = -1;
    for (int b = 0; b < SW_HIST_BINS; ++b) {
        if (cum + h[CNT_HIST + b] > ranks[0]) {
            bstar = b;
            break;
        }
        cum += h[CNT_HIST + b];
    }
    if (bstar <= 0 || bstar >= SW_HIST_BINS - 1) return 1;             // clamped end bins: not usable
    // t~ lies in bin bstar; one extra bin on either side covers the rounding of the bin function
    const float bin_w = 1.0f / hscale;
    const float t_lo = hlo + (float)(bstar - 1) * bin_w, t_hi = hlo + (float)(bstar + 2) * bin_w;
    // the exact value at a rank differs from the D~ value at that rank by at most max eps
    const float eps_abs = eps_abs_coeff(d) * rmax;
    const float delta = 2.0f * emax * 1.0001f + eps_abs + 2.0f * fmaxf(fabsf(t_lo), fabsf(t_hi)) * 1.2e-7f;
    const float tlo = t_lo - delta, thi = t_hi + delta;

    if (list_len) {
        const unsigned grid = (unsigned)std::min<unsigned long long>((list_len + 255) / 256, 16ull * ctx->num_sms);
        band_filter_kernel<<<grid, BF_THREADS, 0, ctx->stream>>>(A.list, list_len, A.e, tlo, thi, eps_abs,
                                                                A.counters + CNT_BELOW2, A.counters + CNT_BANDW,
                                                                A.counters + CNT_BAND_LEN, A.band, A.band_cap,
                                                                d_overflow2);
        STEIN_CHECK_LAUNCH(ctx);
    }
    if (world > 1) STEIN_TRY(allreduce_u64(ctx, A.counters + CNT_BELOW2, 3));
    // No host round trip here: the band length stays on the device (list_len bounds it), the
    // counts of the filter are fetched together with the first histogram of the select below.
    const unsigned long long band_bound = std::min<unsigned long long>(list_len, A.band_cap);
    // contract-arithmetic distance of every band pair: (i, jw) -> (key, weight), in place
    STEIN_TRY(launch_pair_chain<0>(ctx, A.band, band_bound, X, r, n, ld, 0, A.counters + CNT_BAND_LEN));

    // exact keys of the two target ranks among the band.  A band pair has |D - D~| <= eps <= delta
    // and D~ within eps of [tlo, thi], so its key lies in the window below (normally a few
    // thousand keys wide: one histogram pass)
    const uint32_t klo = float_to_key(tlo - 2.0f * delta), khi = float_to_key(thi + 2.0f * delta);
    uint64_t rk[2] = {0, 0};
    uint32_t k01[2] = {0u, 0u};
    FirstSync fs;
    fs.m_dev = A.counters + CNT_BAND_LEN;
    fs.extra_dev = A.counters + CNT_BELOW2;
    fs.extra_host = h + CNT_BELOW2;
    fs.extra_words = 4;                     // below2, band weight, overflow flag, band length
    fs.ranks = [&](uint64_t out[2]) -> int {
        if (h[CNT_OVERFLOW2] || h[CNT_BAND_LEN] > A.band_cap) return 1;
        const uint64_t c1 = below + h[CNT_BELOW2], band_w = h[CNT_BANDW];
        if (!(c1 <= ranks[0] && ranks[1] < c1 + band_w)) return 1;
        out[0] = ranks[0] - c1;
        out[1] = ranks[1] - c1;
        return 0;
    };
    const int rc = window_select2<1>(ctx, A.band, band_bound, klo, khi, rk, k01, true, &fs);
    if (rc != STEIN_OK) return rc;
    // the selected values must lie where the certainty argument holds
    const float m0 = key_to_float(k01[0]), m1 = key_to_float(k01[1]);
    if (!(m0 >= tlo && m1 <= thi && m0 >= p.wlo && m1 <= p.whi)) return 1;
    keys_out[0] = k01[0];
    keys_out[1] = k01[1];
    return STEIN_OK;
}

}  // namespace stein

// Test hook: how many median calls of engine sequences ran pilot-less (hits) / tried and fell back (misses).
extern "C" void stein_debug_median_direct_stats(long long *hits, long long *misses) {
    if (hits) *hits = stein::g_arena.direct_hits;
    if (misses) *misses = stein::g_arena.direct_misses;
}
// Test hooks (pure host functions): enumeration of the CTA-pair sweep's tiles.
extern "C" long long stein_debug_num_pair_tiles(long long T) { return stein::num_pair_tiles(T); }
// Test hook (pure host function): the cost-balanced chunk boundaries of the CTA-pair sweep (parts + 1 entries).
extern "C" void stein_debug_sweep_chunk_bounds(long long T, int parts, double beta, long long *out) {
    const std::vector<long long> b = stein::sweep_chunk_bounds(T, parts, beta);
    for (int k = 0; k <= parts; ++k) out[k] = b[(size_t)k];
}
extern "C" void stein_debug_pair_tile(long long t, int T, int *I2, int *J) { stein::pair_tile(t, T, *I2, *J); }
