// median.cu -- kernel (2): exact median of the n x n squared-distance matrix
// without materialising it.
//
// Reference: stein/kernels/abstract_kernel.py:33-40 (D, bandwidth) and
// stein/utilities/compute_median.py:4-16 (top_k median over all n*n entries).
//
// Method: D is symmetric, so only the upper-triangular 128x128 tiles are
// computed (off-diagonal tiles weigh 2).  Each sweep recomputes the tiles with
// the FFMA mainloop in contract arithmetic and histograms the order-preserving
// u32 keys that fall into a window [key_lo, key_lo + nbins << shift); keys
// below the window are only counted.  A cheap pilot (2^20 sampled pairs, pair_chain.cuh) puts
// the window around the median so that one sweep usually resolves individual
// fp32 values (shift == 0); otherwise the window is narrowed and swept again
// (radix select).  Counts are u64 (n*n = 2^32 at n = 65 536).
#include <math.h>

#include <algorithm>
#include <vector>

#include "gemm_simt.cuh"
#include "pair_chain.cuh"

namespace stein {

// ---- r_i = sum_k x_ik^2, fma chain over k ascending ------------------------
__global__ void row_norms_kernel(const float *__restrict__ X, int64_t rows, int64_t ld,
                                 float *__restrict__ r) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const float4 *row = reinterpret_cast<const float4 *>(X + i * ld);
    float acc = 0.0f;
    for (int64_t k4 = 0; k4 < ld / 4; ++k4) {
        const float4 v = row[k4];
        acc = __fmaf_rn(v.x, v.x, acc);
        acc = __fmaf_rn(v.y, v.y, acc);
        acc = __fmaf_rn(v.z, v.z, acc);
        acc = __fmaf_rn(v.w, v.w, acc);
    }
    r[i] = acc;
}

// ---- one histogram sweep ------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 2)
sqdist_hist_kernel(const float *__restrict__ X, const float *__restrict__ r, int64_t n, int64_t ld,
                   int64_t T, int64_t tile_begin, int64_t tile_end, uint32_t key_lo, uint32_t shift,
                   uint32_t nbins, unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GemmSmem &gs = *reinterpret_cast<GemmSmem *>(smem_raw);
    unsigned int *hist = reinterpret_cast<unsigned int *>(smem_raw + sizeof(GemmSmem));
    unsigned int *below_s = hist + nbins;

    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) hist[b] = 0u;
    if (threadIdx.x == 0) *below_s = 0u;
    __syncthreads();

    unsigned int below = 0u;
    float acc[8][8];
    for (int64_t t = tile_begin + blockIdx.x; t < tile_end; t += gridDim.x) {
        int I, J;
        tri_tile(t, T, I, J);
        const int64_t m0 = (int64_t)I * TILE, n0 = (int64_t)J * TILE;
        gemm_tile<true>(X, ld, m0, X, ld, n0, 0, (int)ld, gs, acc);
        const unsigned int w = (I == J) ? 1u : 2u;
        float ri[8], rj[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            ri[q] = r[m0 + acc_row(q)];
            rj[q] = r[n0 + acc_col(q)];
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int64_t i = m0 + acc_row(a);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int64_t j = n0 + acc_col(c);
                if (i < n && j < n) {
                    const float tsum = ri[a] + rj[c];
                    const float dist = tsum - 2.0f * acc[a][c];
                    const uint32_t key = float_to_key(dist);
                    if (key < key_lo) {
                        below += w;
                    } else {
                        const uint32_t b = (key - key_lo) >> shift;
                        if (b < nbins) atomicAdd(&hist[b], w);
                    }
                }
            }
        }
    }
    // per-thread "below" -> warp -> block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if ((threadIdx.x & 31) == 0 && below) atomicAdd(below_s, below);
    __syncthreads();
    if (threadIdx.x == 0 && *below_s) atomicAdd(&counts[0], (unsigned long long)*below_s);
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) {
        const unsigned int v = hist[b];
        if (v) atomicAdd(&counts[1 + b], (unsigned long long)v);
    }
}

// ---- small n: every key of D, once ----------------------------------------------------------
// keys[i * n + j] for all i, j < n from the upper-triangular tiles (the mirror image is written
// too: D is symmetric bit for bit in contract arithmetic).  With n <= 2048 the n*n keys fit
// 16 MB and the device-side radix select below finds both ranks with one host round trip.
__global__ void __launch_bounds__(GEMM_THREADS, 2)
sqdist_keys_kernel(const float *__restrict__ X, const float *__restrict__ r, int64_t n, int64_t ld, int64_t T,
                   uint32_t *__restrict__ keys) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GemmSmem &gs = *reinterpret_cast<GemmSmem *>(smem_raw);
    float acc[8][8];
    const int64_t ntiles = T * (T + 1) / 2;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int I, J;
        tri_tile(t, T, I, J);
        const int64_t m0 = (int64_t)I * TILE, n0 = (int64_t)J * TILE;
        gemm_tile<true>(X, ld, m0, X, ld, n0, 0, (int)ld, gs, acc);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int64_t i = m0 + acc_row(a);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int64_t j = n0 + acc_col(c);
                if (i < n && j < n) {
                    const uint32_t key = float_to_key((r[i] + r[j]) - 2.0f * acc[a][c]);
                    keys[i * n + j] = key;
                    if (I != J) keys[j * n + i] = key;
                }
            }
        }
    }
}

// ---- device-side radix select of two ranks among m unweighted keys ---------------------------
// Three digit passes (11 + 11 + 10 bits).  Each pass is a multi-block histogram of the keys that
// still match the prefix of rank q (q = 0, 1) followed by a one-block kernel that picks the digit
// and narrows the prefix -- the host only reads the two final keys.
struct SelectState {
    uint32_t prefix[2], mask[2];
    unsigned long long rank[2];
    unsigned int bins[2][2048];
};

__global__ void select_init_kernel(SelectState *st, unsigned long long rank0, unsigned long long rank1) {
    for (int b = threadIdx.x; b < 2 * 2048; b += blockDim.x) (&st->bins[0][0])[b] = 0u;
    if (threadIdx.x == 0) {
        st->prefix[0] = st->prefix[1] = 0u;
        st->mask[0] = st->mask[1] = 0u;
        st->rank[0] = rank0;
        st->rank[1] = rank1;
    }
}

__global__ void __launch_bounds__(256)
select_hist_kernel(const uint32_t *__restrict__ keys, int64_t m, SelectState *st, int shift, int bits) {
    __shared__ unsigned int h[2][2048];
    const int nb = 1 << bits;
    for (int b = threadIdx.x; b < 2 * 2048; b += blockDim.x) (&h[0][0])[b] = 0u;
    __syncthreads();
    const uint32_t p0 = st->prefix[0], p1 = st->prefix[1], m0 = st->mask[0], m1 = st->mask[1];
    const bool same = p0 == p1;          // both ranks still in the same bucket: one histogram serves both
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = keys[i];
        const unsigned dg = (k >> shift) & (nb - 1);
        if ((k & m0) == p0) atomicAdd(&h[0][dg], 1u);
        if (!same && (k & m1) == p1) atomicAdd(&h[1][dg], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        if (h[0][b]) atomicAdd(&st->bins[0][b], h[0][b]);
        if (!same && h[1][b]) atomicAdd(&st->bins[1][b], h[1][b]);
    }
}

// one block of 1024 threads: thread t owns bins 2t and 2t+1; a block-wide scan locates the digit
__global__ void __launch_bounds__(1024)
select_pick_kernel(SelectState *st, int shift, int bits, uint32_t *out /* [2] after the last pass */) {
    __shared__ unsigned long long wsum[32];
    __shared__ int s_dg[2];
    __shared__ unsigned long long s_cum[2];
    const int nb = 1 << bits, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool same = st->prefix[0] == st->prefix[1];
    if (t < 2) s_dg[t] = -1;
    __syncthreads();
    for (int q = 0; q < 2; ++q) {
        const unsigned int *bins = st->bins[(q == 1 && same) ? 0 : q];
        const unsigned long long rk = st->rank[q];
        const unsigned long long a = 2 * t < nb ? bins[2 * t] : 0ull, b = 2 * t + 1 < nb ? bins[2 * t + 1] : 0ull;
        unsigned long long incl = a + b;
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
            wsum[lane] = wi - w;        // exclusive prefix of the warp totals
        }
        __syncthreads();
        const unsigned long long excl = wsum[warp] + incl - (a + b);
        if (rk >= excl && rk < excl + a) {
            s_dg[q] = 2 * t;
            s_cum[q] = excl;
        } else if (rk >= excl + a && rk < excl + a + b) {
            s_dg[q] = 2 * t + 1;
            s_cum[q] = excl + a;
        }
        __syncthreads();
    }
    if (t == 0) {
        for (int q = 0; q < 2; ++q) {
            int dg = s_dg[q];
            unsigned long long cum = s_cum[q];
            if (dg < 0) {            // rank beyond the data (cannot happen for valid ranks): last digit
                dg = nb - 1;
                cum = 0;
            }
            st->rank[q] -= cum;
            st->prefix[q] |= (uint32_t)dg << shift;
            st->mask[q] |= (uint32_t)(nb - 1) << shift;
        }
        if (shift == 0) {
            out[0] = st->prefix[0];
            out[1] = st->prefix[1];
        }
    }
    __syncthreads();
    for (int b2 = t; b2 < 2 * 2048; b2 += blockDim.x) (&st->bins[0][0])[b2] = 0u;
}

__global__ void values_to_keys_kernel(const float *__restrict__ v, int64_t m, uint32_t *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) keys[i] = float_to_key(v[i]);
}

static int ensure_scratch(stein_ctx *ctx, int64_t pilot_m);

// keys at ranks rank0 <= rank1 among m device keys -> ctx->d_sel[0..1] (no host round trip)
static int select_two_keys(stein_ctx *ctx, const uint32_t *keys_dev, int64_t m, uint64_t rank0, uint64_t rank1) {
    static SelectState *st = nullptr;       // one per process (one GPU per process)
    if (!st) STEIN_CHECK_CUDA(ctx, cudaMalloc(&st, sizeof(SelectState)));
    select_init_kernel<<<1, 1024, 0, ctx->stream>>>(st, rank0, rank1);
    STEIN_CHECK_LAUNCH(ctx);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((m + 255) / 256, 2 * (int64_t)ctx->num_sms));
    const int bits[3] = {11, 11, 10};
    int shift = 32;
    for (int pass = 0; pass < 3; ++pass) {
        shift -= bits[pass];
        select_hist_kernel<<<grid, 256, 0, ctx->stream>>>(keys_dev, m, st, shift, bits[pass]);
        STEIN_CHECK_LAUNCH(ctx);
        select_pick_kernel<<<1, 1024, 0, ctx->stream>>>(st, shift, bits[pass], ctx->d_sel);
        STEIN_CHECK_LAUNCH(ctx);
    }
    return STEIN_OK;
}

static int ensure_scratch(stein_ctx *ctx, int64_t pilot_m) {
    if (!ctx->d_counts) {
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_counts, sizeof(uint64_t) * (HIST_MAX_BINS + 2)));
        STEIN_CHECK_CUDA(ctx, cudaMallocHost(&ctx->h_counts, sizeof(uint64_t) * (HIST_MAX_BINS + 2)));
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_sel, sizeof(uint32_t) * 2));
        STEIN_CHECK_CUDA(ctx, cudaMallocHost(&ctx->h_sel, sizeof(uint32_t) * 2));
    }
    if (pilot_m > ctx->pilot_cap) {
        if (ctx->d_pilot_keys) cudaFree(ctx->d_pilot_keys);
        ctx->d_pilot_keys = nullptr;
        STEIN_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_pilot_keys, sizeof(uint32_t) * pilot_m));
        ctx->pilot_cap = pilot_m;
    }
    return STEIN_OK;
}

static int check_layout(stein_ctx *ctx, const void *X, int64_t n, int64_t d, int64_t ld) {
    STEIN_REQUIRE(ctx, X != nullptr, "null particle pointer");
    STEIN_REQUIRE(ctx, n >= 1 && d >= 1, "n=%lld d=%lld must be positive", (long long)n, (long long)d);
    STEIN_REQUIRE(ctx, ld >= d && ld % LD_ALIGN == 0, "ld=%lld must be a multiple of %d and >= d=%lld",
                  (long long)ld, LD_ALIGN, (long long)d);
    STEIN_REQUIRE(ctx, ((uintptr_t)X & 15u) == 0, "particle pointer must be 16-byte aligned");
    return STEIN_OK;
}

struct Window {
    uint32_t key_lo, shift, nbins;
    bool operator==(const Window &o) const {
        return key_lo == o.key_lo && shift == o.shift && nbins == o.nbins;
    }
};
static const Window kFullWindow = {0u, 18u, (uint32_t)HIST_MAX_BINS};

// median_tc.cu
bool median_tc_supported(int64_t n, int64_t ld);
struct PilotSpec {
    const uint32_t *keys_dev;
    unsigned long long m_local;
    unsigned long long rank_lo, rank_hi;
    int direct;
};
bool median_tc_has_hint(const stein_ctx *ctx);
bool median_tc_direct_ok(const stein_ctx *ctx);
void median_tc_count_direct_hit(void);
void median_tc_note_result(const stein_ctx *ctx, uint32_t k0, uint32_t k1, bool direct_missed);
bool median_tc_deferred_pending(const void *owner);
void median_tc_cancel_deferred(void);
int median_tc_finish_deferred(stein_ctx *ctx, uint32_t keys_out[2]);
const float *median_tc_device_bandwidth(void);
bool median_tc_device_select_valid(float *bandwidth);
int median_tc_begin(stein_ctx *ctx, const float *X, const float *r, int64_t n, int64_t ld);
void median_tc_reset(void);
bool median_tc_take_begun(const float *X);
int median_tc_pilot(stein_ctx *ctx, uint32_t *keys_dev, unsigned long long m, const float *r, int64_t n, int64_t ld,
                    uint64_t seed);
int median_tc(stein_ctx *ctx, const float *X, const float *r, int64_t n, int64_t d, int64_t ld,
              const uint64_t ranks[2], uint32_t win_lo_key, uint32_t win_hi_key, uint32_t keys_out[2],
              int *sweeps, const PilotSpec *spec);
int pilot_window(stein_ctx *ctx, const uint32_t *keys_dev, int64_t m, uint64_t rank_lo, uint64_t rank_hi,
                 uint32_t *lo_key, uint32_t *hi_key);

}  // namespace stein

using namespace stein;

extern "C" {

int64_t stein_num_tiles(int64_t n) {
    const int64_t T = (n + TILE - 1) / TILE;
    return T * (T + 1) / 2;
}

int stein_tile_coords(int64_t t, int64_t n, int32_t *I, int32_t *J) {
    if (!I || !J || t < 0 || t >= stein_num_tiles(n)) return STEIN_ERR_INVALID;
    int i, j;
    tri_tile(t, (n + TILE - 1) / TILE, i, j);
    *I = i;
    *J = j;
    return STEIN_OK;
}

uint32_t stein_float_to_key(float f) { return float_to_key(f); }
float stein_key_to_float(uint32_t k) { return key_to_float(k); }

float stein_bandwidth(float median, int64_t n_particles) {
    // abstract_kernel.py:40 -- np.log(n) is folded into an fp32 constant
    const float ln_n = (float)log((double)n_particles);
    return sqrtf(median / ln_n);
}

int stein_row_norms(stein_ctx *ctx, const float *X_dev, int64_t n, int64_t d, int64_t ld,
                    float *r_dev) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_TRY(check_layout(ctx, X_dev, n, d, ld));
    const int64_t rows = stein_rows_padded(n);
    const int threads = 128;
    row_norms_kernel<<<(unsigned)((rows + threads - 1) / threads), threads, 0, ctx->stream>>>(
        X_dev, rows, ld, r_dev);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

int stein_sqdist_hist(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n, int64_t d,
                      int64_t ld, int64_t tile_begin, int64_t tile_end, uint32_t key_lo,
                      uint32_t shift, uint32_t nbins, uint64_t *counts_dev) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_TRY(check_layout(ctx, X_dev, n, d, ld));
    STEIN_REQUIRE(ctx, nbins >= 1 && nbins <= (uint32_t)HIST_MAX_BINS, "nbins=%u out of range", nbins);
    STEIN_REQUIRE(ctx, shift < 32, "shift=%u out of range", shift);
    STEIN_REQUIRE(ctx, tile_begin >= 0 && tile_end <= stein_num_tiles(n) && tile_begin <= tile_end,
                  "tile range [%lld,%lld) invalid", (long long)tile_begin, (long long)tile_end);
    if (tile_begin == tile_end) return STEIN_OK;
    const size_t smem = sizeof(GemmSmem) + sizeof(unsigned int) * (nbins + 4);
    static bool attr_set = false;
    if (!attr_set) {
        STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(sqdist_hist_kernel,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(sizeof(GemmSmem) + sizeof(unsigned int) * (HIST_MAX_BINS + 4))));
        attr_set = true;
    }
    const int64_t ntiles = tile_end - tile_begin;
    const int64_t grid = std::min<int64_t>(ntiles, 2 * (int64_t)ctx->num_sms);
    const int64_t T = (n + TILE - 1) / TILE;
    RegionTimer timer(ctx, STEIN_REGION_SWEEP);
    sqdist_hist_kernel<<<(unsigned)grid, GEMM_THREADS, smem, ctx->stream>>>(
        X_dev, r_dev, n, ld, T, tile_begin, tile_end, key_lo, shift, nbins,
        reinterpret_cast<unsigned long long *>(counts_dev));
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

int stein_median_narrow(const uint64_t *counts, uint32_t key_lo, uint32_t shift, uint32_t nbins,
                        uint64_t rank, uint32_t *key_out, uint32_t *key_lo_out,
                        uint32_t *shift_out, uint32_t *nbins_out) {
    if (rank < counts[0]) {
        *key_lo_out = 0u;  // below the window
        return -1;
    }
    uint64_t cum = counts[0];
    uint32_t b = 0;
    for (; b < nbins; ++b) {
        if (cum + counts[1 + b] > rank) break;
        cum += counts[1 + b];
    }
    if (b == nbins) {
        *key_lo_out = 1u;  // above the window
        return -1;
    }
    const uint32_t lo = key_lo + (b << shift);
    if (shift == 0) {
        *key_out = lo;
        return 1;
    }
    // the bin holds 2^shift consecutive keys: split it into up to HIST_MAX_BINS bins
    uint32_t nb_log = 0;
    while ((1u << nb_log) < (uint32_t)HIST_MAX_BINS && nb_log < shift) ++nb_log;
    *key_lo_out = lo;
    *shift_out = shift - nb_log;
    *nbins_out = 1u << nb_log;
    return 0;
}

int stein_median_values(stein_ctx *ctx, const float *V_dev, int64_t m, float *median_host) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, V_dev && median_host && m >= 1, "bad arguments");
    STEIN_TRY(ensure_scratch(ctx, m));
    values_to_keys_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(V_dev, m, ctx->d_pilot_keys);
    STEIN_CHECK_LAUNCH(ctx);
    // compute_median.py:9-15
    const uint64_t r0 = (m % 2 == 0) ? (uint64_t)m / 2 - 1 : (uint64_t)m / 2, r1 = (uint64_t)m / 2;
    STEIN_TRY(select_two_keys(ctx, ctx->d_pilot_keys, m, r0, r1));
    STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_sel, ctx->d_sel, sizeof(uint32_t) * 2,
                                          cudaMemcpyDeviceToHost, ctx->stream));
    STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const float lo = key_to_float(ctx->h_sel[0]), hi = key_to_float(ctx->h_sel[1]);
    *median_host = (m % 2 == 0) ? (lo + hi) / 2.0f : lo;
    return STEIN_OK;
}

// mode MEDIAN_FULL: the whole call.  MEDIAN_BEGIN: only when the pilot-less steady state applies (else returns
// MEDIAN_NOT_DEFERRED with nothing enqueued) -- enqueues the device part of the median and returns
// MEDIAN_DEFERRED without waiting for it.  MEDIAN_RESUME: collects that median; a miss continues on the routes of
// the full call.  (engine.cu: the next iteration's median runs behind the download of the updated particles.)
static int median_sqdist_impl(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n, int64_t d,
                              int64_t ld, float *median_host, float *mid_host, int32_t *sweeps_host, int mode);
}  // extern "C"

namespace stein {
int median_sqdist_begin(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n, int64_t d, int64_t ld) {
    return median_sqdist_impl(ctx, X_dev, r_dev, n, d, ld, nullptr, nullptr, nullptr, MEDIAN_BEGIN);
}
int median_sqdist_resume(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n, int64_t d, int64_t ld,
                         float *median_host, int32_t *sweeps_host) {
    return median_sqdist_impl(ctx, X_dev, r_dev, n, d, ld, median_host, nullptr, sweeps_host, MEDIAN_RESUME);
}
bool median_sqdist_deferred_pending(const void *owner) { return median_tc_deferred_pending(owner); }
const float *median_sqdist_device_bandwidth(void) { return median_tc_device_bandwidth(); }
bool median_sqdist_device_select_valid(float *bandwidth) { return median_tc_device_select_valid(bandwidth); }
int median_tc_begin_with_norms(stein_ctx *ctx, const float *X, float *r, int64_t rows_r, int64_t n, int64_t ld);
// Row norms of `rows_r` rows and, when the tensor-core median route will take these particles, its first stage
// (error budgets, scale, FP16 split) from the same read.  Returns false if only the norms are needed.
bool median_sqdist_wants_fused_begin(const stein_ctx *ctx, int64_t n, int64_t ld) {
    return (uint64_t)n * (uint64_t)n >= (1ull << 24) && ctx->median_impl != STEIN_MEDIAN_FFMA && median_tc_supported(n, ld);
}
int median_sqdist_begin_with_norms(stein_ctx *ctx, const float *X_dev, float *r_dev, int64_t rows_r, int64_t n, int64_t ld) {
    return median_tc_begin_with_norms(ctx, X_dev, r_dev, rows_r, n, ld);
}
// which route a deferred median took: 1 pilot-less (window recentred on the last exact median), 2 pilot sample +
// device-picked window around the last one
static int g_deferred_kind = 0;
bool median_sqdist_can_defer(const stein_ctx *ctx, int64_t n, int64_t ld) {
    return (uint64_t)n * (uint64_t)n >= (1ull << 24) && ctx->median_impl != STEIN_MEDIAN_FFMA &&
           median_tc_supported(n, ld) && (median_tc_direct_ok(ctx) || median_tc_has_hint(ctx));
}
}  // namespace stein

extern "C" {

int stein_median_sqdist(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n, int64_t d,
                        int64_t ld, float *median_host, float *mid_host, int32_t *sweeps_host) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    return median_sqdist_impl(ctx, X_dev, r_dev, n, d, ld, median_host, mid_host, sweeps_host, MEDIAN_FULL);
}

static int median_sqdist_impl(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n, int64_t d,
                              int64_t ld, float *median_host, float *mid_host, int32_t *sweeps_host, int mode) {
    STEIN_TRY(check_layout(ctx, X_dev, n, d, ld));
    STEIN_REQUIRE(ctx, median_host != nullptr || mode == MEDIAN_BEGIN, "null output pointer");
    if (mode == MEDIAN_BEGIN) {
        if (!median_sqdist_can_defer(ctx, n, ld)) return MEDIAN_NOT_DEFERRED;
    } else if (mode == MEDIAN_FULL) {
        median_tc_cancel_deferred();      // a deferred median of another call is void once the arena is reused
    }
    const uint64_t dim = (uint64_t)n * (uint64_t)n;
    // compute_median.py:9-15: 0-based ascending ranks of the middle value(s)
    const uint64_t ranks[2] = {dim % 2 == 0 ? dim / 2 - 1 : dim / 2, dim / 2};
    const int64_t pilot_m = (dim >= (1ull << 24)) ? (1ll << 20) : 0;
    STEIN_TRY(ensure_scratch(ctx, pilot_m));

    // Small problems (the reference's own example sizes): all n*n keys once, both ranks selected
    // on the device, one host round trip.  Not used on sharded runs or when a route is forced.
    if (n <= 2048 && !ctx->has_comm && ctx->median_impl == STEIN_MEDIAN_AUTO) {
        STEIN_TRY(ensure_scratch(ctx, (int64_t)dim));
        const int64_t T = (n + TILE - 1) / TILE;
        static bool attr_set = false;
        if (!attr_set) {
            STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(sqdist_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)sizeof(GemmSmem)));
            attr_set = true;
        }
        {
            RegionTimer timer(ctx, STEIN_REGION_SWEEP);
            sqdist_keys_kernel<<<(unsigned)std::min<int64_t>(T * (T + 1) / 2, 2 * (int64_t)ctx->num_sms), GEMM_THREADS,
                                 sizeof(GemmSmem), ctx->stream>>>(X_dev, r_dev, n, ld, T, ctx->d_pilot_keys);
            STEIN_CHECK_LAUNCH(ctx);
        }
        STEIN_TRY(select_two_keys(ctx, ctx->d_pilot_keys, (int64_t)dim, ranks[0], ranks[1]));
        STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_sel, ctx->d_sel, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost,
                                              ctx->stream));
        STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const float lo = key_to_float(ctx->h_sel[0]), hi = key_to_float(ctx->h_sel[1]);
        *median_host = (dim % 2 == 0) ? (lo + hi) / 2.0f : lo;
        if (mid_host) {
            mid_host[0] = lo;
            mid_host[1] = hi;
        }
        if (sweeps_host) *sweeps_host = 1;
        return STEIN_OK;
    }

    Window win[2] = {kFullWindow, kFullWindow};
    bool done[2] = {false, false};
    uint32_t key[2] = {0u, 0u};
    int sweeps = 0;
    bool direct_missed = false;
    if (pilot_m) {
        // sample ranks m/2 -+ 3.5 sqrt(m): the true median lies between them with
        // probability ~1 - 1e-11; a miss is caught below and falls back to the
        // full-range radix select.
        // every rank draws the same sample; with collective hooks each rank evaluates its slice
        // of it and the window histograms are all-reduced
        const int pw = ctx->has_comm ? ctx->comm.world : 1, pr = ctx->has_comm ? ctx->comm.rank : 0;
        const int64_t s0 = pilot_m * pr / pw, s1 = pilot_m * (pr + 1) / pw;
        const bool tc_ok = ctx->median_impl != STEIN_MEDIAN_FFMA && median_tc_supported(n, ld);
        // (the engine may have run the first stage together with the row norms: median_tc_begin_with_norms)
        const bool begun = mode != MEDIAN_RESUME && tc_ok && median_tc_take_begun(X_dev);
        if (mode != MEDIAN_RESUME && !begun) median_tc_reset();
        const uint64_t delta = (uint64_t)(3.5 * sqrt((double)pilot_m));
        auto run_pilot = [&]() -> int {
            if (tc_ok) {
                // the tensor-core route splits s X into FP16 hi + lo anyway: the pilot only has to
                // PLACE the window (every use of it is checked), so it reads the hi half alone
                return median_tc_pilot(ctx, ctx->d_pilot_keys + s0, (unsigned long long)(s1 - s0), r_dev, n, ld,
                                       0x5eedull + (uint64_t)s0);
            }
            return launch_pair_chain<1>(ctx, ctx->d_pilot_keys + s0, (unsigned long long)(s1 - s0), X_dev, r_dev, n, ld,
                                        0x5eedull + (uint64_t)s0);
        };
        if (tc_ok && mode != MEDIAN_RESUME && !begun) STEIN_TRY(median_tc_begin(ctx, X_dev, r_dev, n, ld));
        // steady state of an engine: no pilot at all while the median drifts slowly (median_tc_direct_ok)
        const bool resume_direct = mode == MEDIAN_RESUME && g_deferred_kind == 1;
        const bool resume_pilot = mode == MEDIAN_RESUME && g_deferred_kind == 2;
        if (resume_direct || (mode != MEDIAN_RESUME && tc_ok && median_tc_direct_ok(ctx))) {
            const PilotSpec spec = {nullptr, 0ull, 0ull, 0ull, 1};
            int rc;
            if (resume_direct) {
                rc = median_tc_finish_deferred(ctx, key);
                sweeps += 1;
            } else {
                ctx->median_defer = mode == MEDIAN_BEGIN ? 1 : 0;
                rc = median_tc(ctx, X_dev, r_dev, n, d, ld, ranks, 0u, 0u, key, &sweeps, &spec);
                ctx->median_defer = 0;
                if (mode == MEDIAN_BEGIN) {
                    g_deferred_kind = 1;
                    return rc == 3 ? MEDIAN_DEFERRED : (rc < 0 ? rc : fail(ctx, STEIN_ERR_INTERNAL, "median not deferred"));
                }
            }
            if (rc < 0) return rc;
            if (rc == STEIN_OK) {
                done[0] = done[1] = true;
                median_tc_count_direct_hit();
            } else {
                direct_missed = true;
            }
        }
        if (!(done[0] && done[1]) && !resume_pilot) STEIN_TRY(run_pilot());
        // steady state with a pilot: pilot histogram, window pick and sweep chained on the device (not after a
        // pilot-less miss: the median jumped, the host-driven route below places the window from scratch)
        if (resume_pilot || (!(done[0] && done[1]) && !direct_missed && tc_ok && median_tc_has_hint(ctx))) {
            const PilotSpec spec = {ctx->d_pilot_keys + s0, (unsigned long long)(s1 - s0), pilot_m / 2 - delta,
                                    pilot_m / 2 + delta, 0};
            int rc;
            if (resume_pilot) {
                rc = median_tc_finish_deferred(ctx, key);
                sweeps += 1;
            } else {
                ctx->median_defer = mode == MEDIAN_BEGIN ? 1 : 0;
                rc = median_tc(ctx, X_dev, r_dev, n, d, ld, ranks, 0u, 0u, key, &sweeps, &spec);
                ctx->median_defer = 0;
                if (mode == MEDIAN_BEGIN) {
                    g_deferred_kind = 2;
                    return rc == 3 ? MEDIAN_DEFERRED : (rc < 0 ? rc : fail(ctx, STEIN_ERR_INTERNAL, "median not deferred"));
                }
            }
            if (rc < 0) return rc;
            if (rc == STEIN_OK) done[0] = done[1] = true;
        }
        uint32_t ka = 1u, kb = 0u;
        if (!(done[0] && done[1])) {
            const int prc = pilot_window(ctx, ctx->d_pilot_keys + s0, s1 - s0, pilot_m / 2 - delta,
                                         pilot_m / 2 + delta, &ka, &kb);
            if (prc < 0) return prc;
            if (prc != STEIN_OK) {   // degenerate sample: no window, full-range FFMA select below
                ka = 1u;
                kb = 0u;
            }
        }
        auto window_of = [](uint32_t a, uint32_t b) {
            const uint64_t span = (uint64_t)b - a + 1;
            uint32_t sh = 0;
            while (((span - 1) >> sh) >= (uint64_t)HIST_MAX_BINS) ++sh;
            return Window{a, sh, (uint32_t)(((span - 1) >> sh) + 1)};
        };
        if (kb >= ka) {
            win[0] = win[1] = window_of(ka, kb);
            // tensor-core route: one tcgen05 sweep + exact recomputation of the few pairs
            // that can matter; falls through to the FFMA sweeps if it cannot bracket the rank
            if (tc_ok) {
                const int rc = median_tc(ctx, X_dev, r_dev, n, d, ld, ranks, ka, kb, key, &sweeps, nullptr);
                if (rc < 0) return rc;
                if (rc == STEIN_OK) done[0] = done[1] = true;
            }
        }
        if (tc_ok && !(done[0] && done[1])) {
            // The tensor-core route could not decide (particles far from the origin relative to
            // their spread: the FP16 split is too coarse, and so was the FP16 pilot).  The FFMA
            // sweeps below get a window from the exact pilot, as on the routes without tensor cores.
            STEIN_TRY(launch_pair_chain<1>(ctx, ctx->d_pilot_keys + s0, (unsigned long long)(s1 - s0), X_dev, r_dev, n,
                                           ld, 0x5eedull + (uint64_t)s0));
            win[0] = win[1] = kFullWindow;
            const int prc = pilot_window(ctx, ctx->d_pilot_keys + s0, s1 - s0, pilot_m / 2 - delta,
                                         pilot_m / 2 + delta, &ka, &kb);
            if (prc < 0) return prc;
            if (prc == STEIN_OK && kb >= ka) win[0] = win[1] = window_of(ka, kb);
        }
    }
    if ((ctx->median_impl == STEIN_MEDIAN_TC || ctx->median_impl == STEIN_MEDIAN_TC1) && !(done[0] && done[1]))
        return fail(ctx, STEIN_ERR_UNSUPPORTED, "tensor-core median route not applicable (n=%lld d=%lld)",
                    (long long)n, (long long)d);

    const int world = ctx->has_comm ? ctx->comm.world : 1;
    const int rank = ctx->has_comm ? ctx->comm.rank : 0;
    const int64_t ntiles = stein_num_tiles(n);
    const int64_t t0 = ntiles * rank / world, t1 = ntiles * (rank + 1) / world;

    for (int iter = 0; iter < 16 && !(done[0] && done[1]); ++iter) {
        const int q0 = done[0] ? 1 : 0;
        const Window w = win[q0];
        const size_t bytes = sizeof(uint64_t) * (w.nbins + 1);
        STEIN_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_counts, 0, bytes, ctx->stream));
        STEIN_TRY(stein_sqdist_hist(ctx, X_dev, r_dev, n, d, ld, t0, t1, w.key_lo, w.shift, w.nbins,
                                    ctx->d_counts));
        if (world > 1) STEIN_TRY(allreduce_u64(ctx, ctx->d_counts, (int64_t)w.nbins + 1));
        STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, bytes,
                                              cudaMemcpyDeviceToHost, ctx->stream));
        STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ++sweeps;
        for (int q = q0; q < 2; ++q) {
            if (done[q] || !(win[q] == w)) continue;
            Window nw = w;
            const int rc = stein_median_narrow(ctx->h_counts, w.key_lo, w.shift, w.nbins, ranks[q],
                                               &key[q], &nw.key_lo, &nw.shift, &nw.nbins);
            if (rc == 1) {
                done[q] = true;
            } else if (rc == 0) {
                win[q] = nw;
            } else {
                if (w == kFullWindow)
                    return fail(ctx, STEIN_ERR_INTERNAL, "rank outside the full key range");
                win[q] = kFullWindow;  // pilot window missed: exact fallback
            }
        }
    }
    if (!(done[0] && done[1])) return fail(ctx, STEIN_ERR_INTERNAL, "median select did not converge");
    if (pilot_m && ctx->median_impl != STEIN_MEDIAN_FFMA && median_tc_supported(n, ld))
        median_tc_note_result(ctx, key[0], key[1], direct_missed);
    const float lo = key_to_float(key[0]), hi = key_to_float(key[1]);
    // compute_median.py:13 -- tf.reduce_mean of two fp32 values
    *median_host = (dim % 2 == 0) ? (lo + hi) / 2.0f : lo;
    if (mid_host) {
        mid_host[0] = lo;
        mid_host[1] = hi;
    }
    if (sweeps_host) *sweeps_host = sweeps;
    return STEIN_OK;
}

}  // extern "C"
