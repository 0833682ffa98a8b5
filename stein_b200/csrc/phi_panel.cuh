// phi_panel.cuh -- phi for particle matrices of 512 / 768 / 1024 coordinates (BASELINE.json config E:
// n = 262 144, d = 1 024) on the tensor cores.  Included at the end of phi_tc.cu (namespace stein):
// it shares the centring, operand preparation, guard and finalize code of the flash kernels.
//
// Reference math: stein/kernels/squared_exponential_kernel.py:22-35 and
// stein/samplers/abstract_stein_sampler.py:100-105, in the same algebraic form as phi_tc.cu
//     O_i = sum_j K_ij y_j,  y_j = s_j - x_j / h^2,  ksum_i = sum_j K_ij,  phi_i = (O_i + x_i ksum_i / h^2) / n
// on centred particles, with the same split arithmetic (fast: FP16 + 2 x FP8 passes; precise: 3 x FP16).
//
// Beyond d = 256 the fused flash schedule does not exist: the O accumulator of a 128-row tile
// (128 x d fp32) exceeds tensor memory, so either GEMM1 is recomputed for every 256-column slice of O
// (2.5 x the tensor work at d = 1 024) or K crosses memory once.  It crosses once, but never as an
// n x n object: the local rows are cut into PANELS and the columns into CHUNKS, and one block of
// P = exp(-D / 2h^2) (2 B FP16 + 2 x 1 B FP8 per entry, <= 256 MiB, see p_budget_tiles) lives between the
// two kernels that touch it.  The block is stored BOX-MAJOR -- every [128 rows][128 bytes] box that kernel B's
// TMA loads is one contiguous 16 KB piece -- so that kernel A's stores and kernel B's loads both move whole
// 128-byte lines with DRAM-page locality (in a row-major block the 32-byte row pieces of the FP8 arrays,
// 26 KB apart, held kernel A at 39 % tensor-pipe activity; without its stores it ran at 95 %):
//   kernel A (ExpPolicy)  P[panel, chunk] = exp2(c1 X_panel X_chunk^T + a_i + b_j), row sums per tile
//   kernel B (AccPolicy)  O[panel, :]    += P[panel, chunk] Y[chunk, :]
// (Measured alternative, not kept: ONE cooperative launch per panel that alternates the two phases over
//  L2-sized blocks with grid barriers in between.  P then never reaches DRAM and kernel A's operands are not
//  evicted by its write stream, but each phase is only 1-2 waves of tiles long: the drained pipelines and the
//  exposed last epilogue per phase cost more than the traffic saved -- 18.0 ms against 14.5 ms for 32 768 rows
//  x 65 536 columns, d = 1 024.)
// Both are instances of the K-streaming main loop of panel_gemm.cuh (256 x 256 tiles, CTA pairs).
// Tensor work per pair of kernels = the algorithmic 2 GEMMs; HBM traffic = the operand arrays once per
// panel (L2 serves the reuse inside a launch).
#pragma once
#include <stdlib.h>

#include "panel_gemm.cuh"

namespace stein {
namespace panel {

using namespace pg;

// 256 x 256 x 4 B tiles of P per block (environment STEIN_PANEL_TILES).  Measured on one B200.  n = 32 768,
// d = 1 024: an L2-sized block (256 tiles = 64 MiB) runs the two kernels at 487 TFLOP/s algorithmic -- 190
// launches of ~60 us each, K loops of 1 024 -- while 1 024 tiles (240 MiB, 40 launches) reach 769: the P block
// then streams through HBM (its four column-slice readers run side by side and share it in L2), which the
// tensor-bound kernels hide.  One rank's share of config E (32 768 rows x 262 144 columns), second repetition
// (power-capped clocks): 1 024 tiles 682, 1 536: 705, 2 048: 709, 3 072: 711 TFLOP/s -- 2 048 it is (481 MiB).
static int64_t p_budget_tiles() {
    static int64_t v = 0;
    if (!v) {
        v = 2048;
        if (const char *e = getenv("STEIN_PANEL_TILES")) {
            const long long x = atoll(e);
            if (x >= 2 && x <= (1 << 20)) v = x;
        }
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// kernel A: exponentials of one block of the kernel matrix
// ---------------------------------------------------------------------------------------------
// Tensor maps of the P arrays for the epilogue's TMA stores (boxes of 32 rows x 32 columns, dense rows)
struct PStoreMaps {
    CUtensorMap p16, pl, ph;      // pl: E4M3 residual (fast) or FP16 residual (precise); ph: E4M3 of P (fast only)
};

struct ExpPolicy {
    // 5 ring stages + 32 KB of staging: a warp's 32 x 32 block of P goes to shared memory and leaves through
    // TMA stores.  (Direct stores -- every thread its own row, 16 bytes at a time -- cost 32 L2 transactions
    // per warp instruction and held this kernel at 47 % tensor-pipe activity: the epilogue, not the MMAs,
    // set the pace.)
    static constexpr int STAGES = 5;
    static constexpr size_t STAGING_PER_WARP = 6144;
    static constexpr size_t TAIL_STAGING = 3072;     // offset of the staging area in the tail (128-byte aligned)
    static constexpr size_t TAIL_BYTES = TAIL_STAGING + EPI_WARPS * STAGING_PER_WARP + 64;
    struct Params : Core {
        const PStoreMaps *smaps;     // device copy of the store maps (64-byte aligned)
        int tiles_i, tiles_j;        // tile grid of this launch (256-row x 256-column tiles)
        int win_a, win_b;            // the clusters work on a window of win_a row tiles x win_b column tiles at a time
        const float *nrm;            // -r_j log2(e) / (2 h^2) by global particle index; -inf beyond n
        float c1;                    // log2(e) / h^2
        const float *c1mul;          // device: undoes the power-of-two scaling of X
        int precise;
        int debug_skip;              // 0 in production; 1: no epilogue work, 2: TMEM loads only, 3: no stores
        uint16_t *P16;               // [panel rows][pcols] FP16
        uint8_t *Pl8, *Ph8;          // fast: E4M3 of (P - P16) 2^12 / of P
        uint16_t *Pl16;              // precise: FP16 of (P - P16) 2^12
        long long pcols;             // columns of the P block (capacity)
        int nkb16, nkb8;             // boxes per row tile: pcols / 64 (two-byte arrays), pcols / 128 (one-byte arrays)
        float *ksp;                  // [panel rows][ksp_ld] row sums of P per column tile
        int ksp_ld;
    };
    // Tile order.  The stream of P stores (0.5 GB per launch) pushes the operand lines out of L2 between two uses
    // (measured: with the clusters on 37 row tiles x 2 column tiles per wave, every wave re-read its 39 MB of
    // operands from DRAM -- 1 GB per launch, 57 % tensor-pipe activity against 96 % without the stores).  What
    // DOES hit is a line that several clusters want at the same moment.  So a wave is a WINDOW of win_a row tiles
    // x win_b column tiles (win_a * win_b <= clusters, win_a + win_b as small as possible: 9 x 8 on 74 clusters):
    // each row tile is shared by win_b clusters, each column tile by win_a, and a wave brings win_a + win_b
    // tiles of operands from DRAM instead of one per cluster.
    __device__ static int tile(const Params &p, long long k, int cl, int, int &ti, int &tj) {
        const int a = p.win_a, b = p.win_b;
        const int nwi = (p.tiles_i + a - 1) / a, nwj = (p.tiles_j + b - 1) / b;
        if (cl >= a * b || k >= (long long)nwi * nwj) return 0;
        ti = (int)(k / nwj) * a + cl % a;          // column windows fastest: the row tiles of a window stay
        tj = (int)(k % nwj) * b + cl / a;
        return (ti < p.tiles_i && tj < p.tiles_j) ? 1 : 2;
    }
    __device__ static void init_shared(uint8_t *, int) {}

    struct Epilogue {
        const Params &p;
        float *sB, *sK;
        uint8_t *stg;                // this warp's staging: [P16 32 x 64 B | pl 32 x 64 B | ph 32 x 64 B]; the FP8 arrays
                                     // collect TWO chunks per row before they leave: 64-byte pieces (32-byte ones cost
                                     // 2.4 x more per byte -- partial DRAM write granules)
        int q, wg, row, lane, tid256;
        uint32_t lane_addr, rank;
        float c1, a_i, ksum;
        long long prow;
        int tj, par;
        int b_row0;                  // first column of the current block
        __device__ Epilogue(const Params &p_, uint8_t *tail, int warp, int lane_, uint32_t rank_)
            : p(p_), lane(lane_), rank(rank_), par(0), b_row0(p_.b_row0) {
            sB = reinterpret_cast<float *>(tail);
            sK = sB + 2 * 256;
            stg = tail + TAIL_STAGING + (size_t)(warp - 4) * STAGING_PER_WARP;
            q = warp & 3;
            wg = (warp - 4) >> 2;
            row = q * 32 + lane;
            tid256 = (warp - 4) * 32 + lane;
            lane_addr = (uint32_t)(q * 32) << 16;
            c1 = p.c1 * __ldg(p.c1mul);
        }
        __device__ void tile_begin(int ti, int tj_) {
            tj = tj_;
            par ^= 1;
            prow = (long long)ti * 256 + rank * 128 + row;
            a_i = p.nrm[(size_t)p.a_row0 + (size_t)prow];
            sB[par * 256 + tid256] = p.nrm[(size_t)b_row0 + (size_t)tj * 256 + tid256];
            named_bar_sync(1, EPI_THREADS);
            ksum = 0.0f;
        }
        __device__ void unit(uint32_t acc_tmem, int, bool) {
            const float *bj = sB + par * 256;
            if (p.debug_skip == 1) return;                      // timing experiments only (STEIN_PANEL_DEBUG_SKIP)
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int ch = wg * 4 + cc;                     // 32-column chunk of the 256-column tile
                uint32_t v[32];
                tmem_ld32(acc_tmem + lane_addr + ch * 32, v);
                tmem_wait_ld();
                if (p.debug_skip == 2) continue;
                uint32_t w[32];
#pragma unroll
                for (int c2 = 0; c2 < 16; ++c2) {
                    const float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * c2]), c1, a_i + bj[ch * 32 + 2 * c2]));
                    const float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * c2 + 1]), c1, a_i + bj[ch * 32 + 2 * c2 + 1]));
                    ksum += e0 + e1;
                    const uint32_t wh = pack_f16x2(e0, e1);
                    const float l0 = (e0 - f16_lo_to_f32(wh)) * 4096.0f, l1 = (e1 - f16_hi_to_f32(wh)) * 4096.0f;
                    w[c2] = wh;
                    if (p.precise) {
                        w[16 + c2] = pack_f16x2(l0, l1);
                    } else {
                        const uint32_t pl = pack_e4m3x2(l0, l1), ph = pack_e4m3x2(e0, e1);
                        if (c2 & 1) {
                            w[16 + (c2 >> 1)] |= pl << 16;
                            w[24 + (c2 >> 1)] |= ph << 16;
                        } else {
                            w[16 + (c2 >> 1)] = pl;
                            w[24 + (c2 >> 1)] = ph;
                        }
                    }
                }
                if (p.debug_skip == 3) continue;
                // this warp's 32 rows x 32 columns through shared memory and out by TMA (the previous block's
                // stores must have finished reading the staging area)
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
                uint4 *s16 = reinterpret_cast<uint4 *>(stg + lane * 64);
#pragma unroll
                for (int k = 0; k < 4; ++k) s16[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
                if (p.precise) {
                    uint4 *sl = reinterpret_cast<uint4 *>(stg + 2048 + lane * 64);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        sl[k] = make_uint4(w[16 + 4 * k], w[17 + 4 * k], w[18 + 4 * k], w[19 + 4 * k]);
                } else {
                    uint4 *sl = reinterpret_cast<uint4 *>(stg + 2048 + lane * 64 + (ch & 1) * 32);
                    uint4 *sh = reinterpret_cast<uint4 *>(stg + 4096 + lane * 64 + (ch & 1) * 32);
                    sl[0] = make_uint4(w[16], w[17], w[18], w[19]);
                    sl[1] = make_uint4(w[20], w[21], w[22], w[23]);
                    sh[0] = make_uint4(w[24], w[25], w[26], w[27]);
                    sh[1] = make_uint4(w[28], w[29], w[30], w[31]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    // box-major destination: row tile rt, K block of the 32 columns, 32 rows from q * 32
                    const int rt = (int)((prow - lane) >> 7), r32 = (int)((prow - lane) & 127);
                    const int o16 = (rt * p.nkb16 + tj * 4 + (ch >> 1)) * 128 + r32, i16 = (ch & 1) * 32;
                    const int o8 = (rt * p.nkb8 + tj * 2 + (ch >> 2)) * 128 + r32, i8 = ((ch >> 1) & 1) * 64;
                    // the P block streams through L2 once on its way to kernel B: evict it before the operands
                    tma_store_2d_hint(&p.smaps->p16, stg, i16, o16, L2_EVICT_FIRST);
                    if (p.precise) {
                        tma_store_2d_hint(&p.smaps->pl, stg + 2048, i16, o16, L2_EVICT_FIRST);
                    } else if ((ch & 1) && p.debug_skip != 4) {      // both halves of the 64-byte pieces are there
                        tma_store_2d_hint(&p.smaps->pl, stg + 2048, i8, o8, L2_EVICT_FIRST);
                        tma_store_2d_hint(&p.smaps->ph, stg + 4096, i8, o8, L2_EVICT_FIRST);
                    }
                    tma_store_commit();
                }
            }
            // row sum of this tile: the two warpgroups own 128 columns each
            if (wg == 1) sK[row] = ksum;
            named_bar_sync(2, EPI_THREADS);
            if (wg == 0) p.ksp[(size_t)prow * p.ksp_ld + tj] = ksum + sK[row];
        }
        __device__ void finish() {
            if (lane == 0) tma_store_wait_all();      // the stores of this CTA are complete before it exits
            __syncwarp();
        }
    };
};

// ---------------------------------------------------------------------------------------------
// kernel B: O[panel, 256-column slice] (+)= P[panel, chunk] Y[chunk, slice]
// ---------------------------------------------------------------------------------------------
struct AccPolicy {
    static constexpr int STAGES = 6;
    static constexpr size_t TAIL_BYTES = 64;
    struct Params : Core {
        int tiles_i, tiles_j;        // panel row tiles x (ld / 256) column slices
        float *O;                    // [local rows padded to 256][ldo] fp32
        long long ldo, prow0;        // first local row of the panel
        int first;                   // first chunk of the panel: overwrite O / ksum instead of accumulating
        const float *ksp;            // row sums of P per column tile of this chunk (kernel A)
        int ksp_ld, ksp_n;
        float *ksum;                 // [local rows]
    };
    __device__ static int tile(const Params &p, long long k, int cl, int ncl, int &ti, int &tj) {
        const long long t = (long long)cl + k * ncl;
        if (t >= (long long)p.tiles_i * p.tiles_j) return 0;
        tj = (int)(t % p.tiles_j);       // the slices of one row tile run side by side: its P rows come from L2 once
        ti = (int)(t / p.tiles_j);
        return 1;
    }
    __device__ static void init_shared(uint8_t *, int) {}

    struct Epilogue {
        const Params &p;
        int q, wg, row, lane;
        uint32_t lane_addr, rank;
        long long prow;
        int tj;
        int first, ksp_n;            // of the current block
        float acc[128];              // fp32 round-to-nearest running sum of the drained units (the tensor
                                     // core accumulates with truncation: the chains inside TMEM stay short)
        __device__ Epilogue(const Params &p_, uint8_t *, int warp, int lane_, uint32_t rank_)
            : p(p_), lane(lane_), rank(rank_), first(p_.first), ksp_n(p_.ksp_n) {
            q = warp & 3;
            wg = (warp - 4) >> 2;
            row = q * 32 + lane;
            lane_addr = (uint32_t)(q * 32) << 16;
        }
        __device__ void tile_begin(int ti, int tj_) {
            tj = tj_;
            prow = (long long)ti * 256 + rank * 128 + row;
#pragma unroll
            for (int c = 0; c < 128; ++c) acc[c] = 0.0f;
        }
        __device__ void unit(uint32_t acc_tmem, int, bool last) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t v[32];
                tmem_ld32(acc_tmem + lane_addr + wg * 128 + ch * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[ch * 32 + c] += __uint_as_float(v[c]);
            }
            if (!last) return;
            float4 *dst = reinterpret_cast<float4 *>(p.O + (size_t)(p.prow0 + prow) * (size_t)p.ldo + (size_t)tj * 256 + wg * 128);
#pragma unroll
            for (int c4 = 0; c4 < 32; ++c4) {
                float4 o = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
                if (!first) {
                    const float4 old = dst[c4];
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                dst[c4] = o;
            }
            if (tj == 0 && wg == 0) {        // row sums of this chunk, in column-tile order
                float s = first ? 0.0f : p.ksum[p.prow0 + prow];
                for (int t = 0; t < ksp_n; ++t) s += p.ksp[(size_t)prow * p.ksp_ld + t];
                p.ksum[p.prow0 + prow] = s;
            }
        }
        __device__ void finish() {}
    };
};

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct PanelPlan {
    int64_t rows, rowsP, cols, colsP, ld;      // rowsP: local rows padded to 256; colsP: columns padded to 512
    int64_t rp, cc;                            // row tiles per panel, column tiles per chunk (even)
};

// Window of kernel A's tile order for a block of r x c tiles on G clusters (see ExpPolicy::tile): returns the
// cost in wave units -- waves times a penalty for the operand bytes a wave brings from DRAM, which grow with
// win_a + win_b (measured: 37 x 2 runs 1.53 x slower than the MMAs alone would, 9 x 8 is the minimum).
static double best_window(int64_t r, int64_t c, int64_t G, int *wa, int *wb) {
    double best = 1e300;
    for (int64_t a = 1; a <= G; ++a) {
        const int64_t b = G / a;
        if (b < 1) break;
        const double waves = (double)(((r + a - 1) / a) * ((c + b - 1) / b));
        const double pen = 1.0 + 0.53 * std::max<double>(0.0, (double)(a + b) - 17.0) / 22.0;
        if (waves * pen < best) {
            best = waves * pen;
            if (wa) *wa = (int)a;
            if (wb) *wb = (int)b;
        }
    }
    return best;
}

// Cost model of the panel / chunk loop in units of (tile x 128 K elements) per cluster wave: picks the
// block shape that wastes the fewest cluster slots to wave quantisation within the L2 budget.
static PanelPlan panel_plan(const stein_ctx *ctx, int64_t n_local, int64_t n_total, int64_t ld) {
    // the search below costs milliseconds: remember the last shape (an engine asks for the same one every iteration)
    static int64_t c_key[4] = {-1, -1, -1, -1};
    static PanelPlan c_plan;
    if (c_key[0] == n_local && c_key[1] == n_total && c_key[2] == ld && c_key[3] == ctx->num_sms) return c_plan;
    PanelPlan pl;
    pl.rows = stein_rows_padded(n_local);
    pl.rowsP = round_up(pl.rows, 256);
    pl.cols = stein_rows_padded(n_total);
    pl.colsP = round_up(pl.cols, 512);
    pl.ld = ld;
    const int64_t R = pl.rowsP / 256, C = pl.colsP / 256, G = std::max(1, ctx->num_sms / 2), NS = ld / 256;
    // waves of one (row tiles r, column tiles c) block: kernel A then kernel B, plus launch overheads
    auto block_cost = [&](int64_t r, int64_t c) {
        return best_window(r, c, G, nullptr, nullptr) * (double)(ld / 128) + (double)((r * NS + G - 1) / G) * (double)(c * 2) + 2.0;
    };
    double best = 1e300;
    pl.rp = 1;
    pl.cc = 2;
    for (int64_t rp = 1; rp <= std::min<int64_t>(R, p_budget_tiles() / 2); ++rp) {
        for (int64_t cc = 2; cc <= std::min<int64_t>(C, p_budget_tiles() / rp); cc += 2) {
            const int64_t nr = R / rp, rr = R % rp, nc = C / cc, cr = C % cc;
            double cost = (double)nr * (double)nc * block_cost(rp, cc);
            if (rr) cost += (double)nc * block_cost(rr, cc);
            if (cr) cost += (double)nr * block_cost(rp, cr);
            if (rr && cr) cost += block_cost(rr, cr);
            if (cost < best) {
                best = cost;
                pl.rp = rp;
                pl.cc = cc;
            }
        }
    }
    if (const char *e = getenv("STEIN_PANEL_RP")) pl.rp = std::max<int64_t>(1, std::min<int64_t>(R, atoll(e)));
    if (const char *e = getenv("STEIN_PANEL_CC")) pl.cc = std::max<int64_t>(2, std::min<int64_t>(C, atoll(e) / 2 * 2));
    c_key[0] = n_local;
    c_key[1] = n_total;
    c_key[2] = ld;
    c_key[3] = ctx->num_sms;
    c_plan = pl;
    return pl;
}

bool panel_supported(const stein_ctx *ctx, int64_t n_total, int64_t ld) {
    return (ld == 512 || ld == 768 || ld == 1024) && n_total >= 2 && ctx->num_sms >= 2 && n_total < (1ll << 30);
}

int64_t panel_workspace_bytes(const stein_ctx *ctx, int64_t n_local, int64_t n_total, int64_t ld) {
    const PanelPlan pl = panel_plan(ctx, n_local, n_total, ld);
    int64_t b = 0;
    b += 5 * pl.cols * ld * 2;                                 // X16, XL, B8 (or Yx), YT16, YTL
    b += (pl.colsP + 256) * 4;                                 // nrm
    b += centred_bytes(pl.cols, ld);
    b += ((int64_t)CM_BLOCKS + 2) * ld * 4 + 64;               // column maxima / scales of Y, scale of X
    b += pl.rowsP * ld * 4 + pl.rowsP * 4;                     // O, ksum
    b += pl.rp * 256 * pl.cc * 256 * 4;                        // one block of P
    b += pl.rp * 256 * pl.cc * 4;                              // its row sums per column tile
    b += FINALIZE_MAX_BLOCKS * 8;
    b += 1024;                                                 // tensor maps of the epilogue's TMA stores
    return b + 8192;
}

static void fill_table(Core &c, bool precise) {
    if (!precise) {           // per 128 K elements: X16.X16 (two blocks of 64), a8l.b8h, a8h.b8l
        c.n_stage = 4;
        c.table[0] = {0, 0, 0, 0};
        c.table[1] = {0, 0, 64, 0};
        c.table[2] = {1, 1, 0, 1};
        c.table[3] = {2, 2, 0, 1};
    } else {                  // per 64 K elements: hi.hi, lo.hi, hi.lo (kernel B: P16.Y16, Pl.Yx, P16.Yl)
        c.n_stage = 6;
        for (int h = 0; h < 2; ++h) {
            c.table[3 * h + 0] = {0, 0, 64 * h, 0};
            c.table[3 * h + 1] = {1, 1, 64 * h, 0};
            c.table[3 * h + 2] = {0, 2, 64 * h, 0};
        }
    }
}

int phi_panel(stein_ctx *ctx, const float *X_all, const float *S_all, const float *r_all, int64_t n_total, int64_t d,
              int64_t d_true, int64_t ld, int64_t row_begin, int64_t n_local, float h2, void *ws, int64_t ws_bytes,
              float *phi, double *sumsq, int mode /* 2 fast, 3 precise, 4 guarded */) {
    STEIN_REQUIRE(ctx, panel_supported(ctx, n_total, ld), "panel phi needs a leading dimension of 512, 768 or 1024");
    STEIN_REQUIRE(ctx, ws_bytes >= panel_workspace_bytes(ctx, n_local, n_total, ld), "phi workspace too small");
    STEIN_REQUIRE(ctx, row_begin % TILE == 0, "row_begin must be a multiple of %d", TILE);
    const PanelPlan pl = panel_plan(ctx, n_local, n_total, ld);
    const int64_t cols = pl.cols, rows = pl.rows;
    char *pws = (char *)ws;
    __half *X16 = (__half *)pws;        pws += cols * ld * 2;
    uint8_t *XL = (uint8_t *)pws;       pws += cols * ld * 2;      // a8l | a8h, or the FP16 residual
    uint8_t *B8 = (uint8_t *)pws;       pws += cols * ld * 2;      // b8h | b8l, or Y16 2^-12 (transposed)
    __half *YT16 = (__half *)pws;       pws += cols * ld * 2;
    uint8_t *YTL = (uint8_t *)pws;      pws += cols * ld * 2;      // y8h | y8l, or the FP16 residual
    float *nrm = (float *)pws;          pws += (pl.colsP + 256) * 4;
    Centred cen{};
    STEIN_TRY(make_centred(ctx, X_all, n_total, d, cols, ld, pws, &cen, true));
    const float *Xc = cen.Xc, *rc = cen.rc;
    pws = (char *)(((uintptr_t)pws + 15) & ~(uintptr_t)15);
    float *cmax_part = (float *)pws;    pws += (int64_t)CM_BLOCKS * ld * 4;
    float *cs_down = (float *)pws;      pws += ld * 4;
    float *cs_up = (float *)pws;        pws += ld * 4;
    float *xscale = (float *)pws;       pws += 16;
    float *O = (float *)pws;            pws += pl.rowsP * ld * 4;
    float *ksum = (float *)pws;         pws += pl.rowsP * 4;
    const int64_t prow_cap = pl.rp * 256, pcols = pl.cc * 256;
    uint16_t *P16 = (uint16_t *)pws;    pws += prow_cap * pcols * 2;
    uint8_t *PL = (uint8_t *)pws;       pws += prow_cap * pcols * 2;    // Pl8 | Ph8, or Pl16
    float *ksp = (float *)pws;          pws += prow_cap * pl.cc * 4;
    double *partials = (double *)(((uintptr_t)pws + 7) & ~(uintptr_t)7);
    pws = (char *)partials + FINALIZE_MAX_BLOCKS * 8;
    PStoreMaps *d_smaps = (PStoreMaps *)(((uintptr_t)pws + 127) & ~(uintptr_t)127);

    xscale_kernel<<<1, 1024, 0, ctx->stream>>>(cen.blockmax, cen.nblockmax, xscale);
    STEIN_CHECK_LAUNCH(ctx);
    const float l2e = 1.4426950408889634f;
    if (mode == 4) STEIN_TRY(guard_begin(ctx, rc, n_total, h2));
    colmax_partial_kernel<<<CM_BLOCKS, 256, 0, ctx->stream>>>(Xc, S_all, n_total, ld, 1.0f / h2, cmax_part);
    STEIN_CHECK_LAUNCH(ctx);
    colscale_kernel<<<(unsigned)((ld * 32 + 255) / 256), 256, 0, ctx->stream>>>(cmax_part, ld, cs_down, cs_up);
    STEIN_CHECK_LAUNCH(ctx);
    if (mode == 4) {
        int route = 0;
        STEIN_TRY(guard_end(ctx, d_true, true, true, false, &route));
        if (route == 2)
            return phi_dense(ctx, X_all, S_all, r_all, n_total, d, ld, row_begin, n_local, h2, ws, ws_bytes, phi, sumsq);
        mode = route == 0 ? 2 : 3;
    }
    const bool precise = mode == 3;
    {
        const int64_t tot = std::max<int64_t>(cols * ld / 4, pl.colsP + 256);
        prep_x_route_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(
            Xc, rc, cols, n_total, ld, 0.5f * l2e / h2, xscale, nullptr, precise ? 1 : 0, X16, XL, XL + cols * ld, B8,
            B8 + cols * ld, nrm, pl.colsP + 256);
        STEIN_CHECK_LAUNCH(ctx);
        dim3 g((unsigned)(cols / 32), (unsigned)(ld / 32)), b(32, 8);
        // precise: the 2^-12 copy of Y16 takes the place of b8h + b8l -- which the precise X layout does not use
        prep_yt_route_kernel<<<g, b, 0, ctx->stream>>>(Xc, S_all, cols, ld, 1.0f / h2, cs_down, nullptr, precise ? 1 : 0,
                                                      YT16, YTL, YTL + cols * ld, (__half *)B8);
        STEIN_CHECK_LAUNCH(ctx);
    }

    // ---- tensor maps: X arrays [cols][ld] (K = coordinates), Y^T arrays [ld][cols] and P arrays [panel rows][pcols]
    // (K = particles); boxes of 128 rows x 128 bytes throughout
    Maps mA, mB;
    memset(&mA, 0, sizeof(mA));
    memset(&mB, 0, sizeof(mB));
    auto mapx = [&](CUtensorMap *m, const void *base, int eb) {
        return make_tensor_map_2d(ctx, m, base, eb, (uint64_t)ld, (uint64_t)cols, (uint64_t)ld * eb, 128);
    };
    auto mapy = [&](CUtensorMap *m, const void *base, int eb) {
        return make_tensor_map_2d(ctx, m, base, eb, (uint64_t)cols, (uint64_t)ld, (uint64_t)cols * eb, 128);
    };
    // P arrays, box-major: [prow_cap / 128 row tiles x nkb boxes x 128 rows][128 bytes]
    const int64_t nkb16 = pcols / 64, nkb8 = pcols / 128;
    auto mapp = [&](CUtensorMap *m, const void *base, int eb) {
        return make_tensor_map_2d(ctx, m, base, eb, (uint64_t)(128 / eb), (uint64_t)(prow_cap * (eb == 2 ? nkb16 : nkb8)), 128,
                                  128);
    };
    STEIN_TRY(mapx(&mA.a[0], X16, 2));
    STEIN_TRY(mapx(&mA.b[0], X16, 2));
    STEIN_TRY(mapy(&mB.b[0], YT16, 2));
    STEIN_TRY(mapp(&mB.a[0], P16, 2));
    if (precise) {
        // table entries (a, b): (0, 0) hi.hi, (1, 1) lo.hi, (0, 2) hi.lo -- slot 1 of B is the HI array again
        STEIN_TRY(mapx(&mA.a[1], XL, 2));
        STEIN_TRY(mapx(&mA.a[2], X16, 2));          // unused
        STEIN_TRY(mapx(&mA.b[1], X16, 2));
        STEIN_TRY(mapx(&mA.b[2], XL, 2));
        // kernel B: (0, 0) P16.Y16, (1, 1) (P - P16) 2^12 . Y16 2^-12, (0, 2) P16.(Y - Y16)
        STEIN_TRY(mapp(&mB.a[1], PL, 2));
        STEIN_TRY(mapp(&mB.a[2], P16, 2));          // unused
        STEIN_TRY(mapy(&mB.b[1], B8, 2));
        STEIN_TRY(mapy(&mB.b[2], YTL, 2));
    } else {
        STEIN_TRY(mapx(&mA.a[1], XL, 1));                         // a8l
        STEIN_TRY(mapx(&mA.a[2], XL + cols * ld, 1));             // a8h
        STEIN_TRY(mapx(&mA.b[1], B8, 1));                         // b8h
        STEIN_TRY(mapx(&mA.b[2], B8 + cols * ld, 1));             // b8l
        STEIN_TRY(mapp(&mB.a[1], PL, 1));                         // Pl8
        STEIN_TRY(mapp(&mB.a[2], PL + prow_cap * pcols, 1));      // Ph8
        STEIN_TRY(mapy(&mB.b[1], YTL, 1));                        // y8h
        STEIN_TRY(mapy(&mB.b[2], YTL + cols * ld, 1));            // y8l
    }

    // store maps of kernel A's epilogue: boxes of 32 x 32 elements, dense rows in shared memory
    {
        PStoreMaps hm;
        memset(&hm, 0, sizeof(hm));
        STEIN_TRY(make_tensor_map_2d_box(ctx, &hm.p16, P16, 2, 64, (uint64_t)(prow_cap * nkb16), 128, 32, 32, false));
        if (precise) {
            STEIN_TRY(make_tensor_map_2d_box(ctx, &hm.pl, PL, 2, 64, (uint64_t)(prow_cap * nkb16), 128, 32, 32, false));
            hm.ph = hm.pl;
        } else {
            STEIN_TRY(make_tensor_map_2d_box(ctx, &hm.pl, PL, 1, 128, (uint64_t)(prow_cap * nkb8), 128, 64, 32, false));
            STEIN_TRY(make_tensor_map_2d_box(ctx, &hm.ph, PL + prow_cap * pcols, 1, 128, (uint64_t)(prow_cap * nkb8), 128, 64, 32,
                                             false));
        }
        // pageable source: the runtime stages the bytes before returning
        STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(d_smaps, &hm, sizeof(hm), cudaMemcpyHostToDevice, ctx->stream));
    }
    using KA = ExpPolicy;
    using KB = AccPolicy;
    const size_t smemA = smem_bytes<KA::STAGES>(KA::TAIL_BYTES), smemB = smem_bytes<KB::STAGES>(KB::TAIL_BYTES);
    static bool attr_set = false;
    if (!attr_set) {
        STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(panel_gemm_kernel<KA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
        STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(panel_gemm_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB));
        attr_set = true;
    }
    const int G = std::max(1, ctx->num_sms / 2);
    const int64_t R = pl.rowsP / 256, C = pl.colsP / 256;
    if (getenv("STEIN_PANEL_VERBOSE"))
        fprintf(stderr, "[stein] panel plan: %lld x %lld tiles, blocks of %lld x %lld (%lld launches), P block %.1f MiB\n",
                (long long)R, (long long)C, (long long)pl.rp, (long long)pl.cc,
                (long long)(2 * ((R + pl.rp - 1) / pl.rp) * ((C + pl.cc - 1) / pl.cc)), pl.rp * pl.cc * 0.25);
    {
        RegionTimer timer(ctx, STEIN_REGION_PHI);
        for (int64_t r0 = 0; r0 < R; r0 += pl.rp) {
            const int64_t r = std::min(pl.rp, R - r0);
            for (int64_t c0 = 0; c0 < C; c0 += pl.cc) {
                const int64_t c = std::min(pl.cc, C - c0);       // even: C and cc are
                KA::Params pa{};
                fill_table(pa, precise);
                pa.groups_per_unit = (int)(ld / 128);
                pa.units_per_tile = 1;
                pa.ka0 = pa.kb0 = 0;
                pa.a_row0 = (int)(row_begin + r0 * 256);
                pa.b_row0 = (int)(c0 * 256);
                pa.route = nullptr;
                pa.my_route = 0;
                pa.pol_a = pa.pol_b = L2_EVICT_LAST;           // the X arrays are read once per tile; P passes through once
                if (const char *e = getenv("STEIN_PANEL_NOHINT")) pa.pol_a = pa.pol_b = atoi(e) ? L2_EVICT_NORMAL : pa.pol_a;
                pa.tiles_i = (int)r;
                pa.tiles_j = (int)c;
                const int clustersA = (int)std::min<int64_t>(G, r * c);
                best_window(r, c, clustersA, &pa.win_a, &pa.win_b);
                pa.nrm = nrm;
                pa.c1 = l2e / h2;
                pa.c1mul = xscale + 1;
                pa.precise = precise ? 1 : 0;
                pa.debug_skip = getenv("STEIN_PANEL_DEBUG_SKIP") ? atoi(getenv("STEIN_PANEL_DEBUG_SKIP")) : 0;
                pa.smaps = d_smaps;
                pa.P16 = P16;
                pa.Pl8 = PL;
                pa.Ph8 = PL + prow_cap * pcols;
                pa.Pl16 = (uint16_t *)PL;
                pa.pcols = pcols;
                pa.nkb16 = (int)nkb16;
                pa.nkb8 = (int)nkb8;
                pa.a_blocked = 0;
                pa.ksp = ksp;
                pa.ksp_ld = (int)pl.cc;
                const int gridA = 2 * clustersA;
                panel_gemm_kernel<KA><<<gridA, THREADS, smemA, ctx->stream>>>(mA, pa);
                STEIN_CHECK_LAUNCH(ctx);

                KB::Params pb{};
                fill_table(pb, precise);
                pb.groups_per_unit = 4;                          // 512 particles per accumulation unit
                pb.units_per_tile = (int)(c * 256 / 512);
                pb.ka0 = 0;
                pb.kb0 = (int)(c0 * 256);
                pb.a_row0 = 0;
                pb.b_row0 = 0;
                pb.route = nullptr;
                pb.my_route = 0;
                pb.a_blocked = 1;
                pb.a_nkb16 = (int)nkb16;
                pb.a_nkb8 = (int)nkb8;
                pb.pol_a = L2_EVICT_FIRST;                     // P: consumed here, never needed again
                pb.pol_b = L2_EVICT_LAST;                      // Y^T of this chunk: read by every row tile
                pb.tiles_i = (int)r;
                pb.tiles_j = (int)(ld / 256);
                pb.O = O;
                pb.ldo = ld;
                pb.prow0 = r0 * 256;
                pb.first = c0 == 0 ? 1 : 0;
                pb.ksp = ksp;
                pb.ksp_ld = (int)pl.cc;
                pb.ksp_n = (int)c;
                pb.ksum = ksum;
                const int gridB = 2 * (int)std::min<int64_t>(G, r * (ld / 256));
                panel_gemm_kernel<KB><<<gridB, THREADS, smemB, ctx->stream>>>(mB, pb);
                STEIN_CHECK_LAUNCH(ctx);
            }
        }
    }
    // finalize: one slot per row
    SlotLayout L{};
    L.O0 = O;
    L.Ox = nullptr;
    L.k0 = ksum;
    L.kx = nullptr;
    L.row0 = 0;
    L.xrows = 0;
    L.DP = (int)ld;
    int *d_ones = nullptr;
    STEIN_TRY(plan_upload(ctx, 2, std::vector<int>((size_t)(pl.rowsP / TILE), 1), n_local, n_total, d, &d_ones));
    const int64_t rows_valid = std::max<int64_t>(0, std::min<int64_t>(n_local, n_total - row_begin));
    const int64_t total4 = rows * ld / 4;
    const int blocks = (int)std::min<int64_t>((total4 + 255) / 256, FINALIZE_MAX_BLOCKS);
    finalize_slots_kernel<<<blocks, 256, 0, ctx->stream>>>(L, d_ones, Xc + row_begin * ld, rows_valid, rows, ld, 1.0f / h2,
                                                           1.0f / (float)n_total, cs_up, phi, partials);
    STEIN_CHECK_LAUNCH(ctx);
    reduce_partials2_kernel<<<1, 256, 0, ctx->stream>>>(partials, blocks, sumsq);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

}  // namespace panel
}  // namespace stein
