// phi_dense.cu -- kernel (3), dense FP32-FFMA path, plus the pieces every phi
// path shares (Y = S - X/h^2, finalize + sum(phi^2)).
//
// Reference: stein/kernels/squared_exponential_kernel.py:22-35 (K, dK) and
// stein/samplers/abstract_stein_sampler.py:100-105 (phi = (K.S + dK)/n).
//
// Algebra used by all phi paths ("TMEM budget"):
//     dK_i = (x_i sum_j K_ij - sum_j K_ij x_j) / h^2
//     phi_i = ( sum_j K_ij (s_j - x_j/h^2)  +  x_i (sum_j K_ij) / h^2 ) / n
// so one GEMM K.Y with Y = S - X/h^2 plus the row sums of K is enough.
//
// The dense path materialises K for a slab of local rows (it is what
// kernel_and_grad() has to return anyway) and is the fallback for shapes the
// tcgen05 flash kernel does not take.  It is exact-order FP32 (FFMA).
#include <algorithm>

#include "gemm_simt.cuh"
#include "phi_common.cuh"

namespace stein {

// Kernel functions of the squared distance (the plugin point stein/kernels/abstract_kernel.py:45-62):
//   KF_SE   exp(-D / h^2 / 2)                      squared_exponential_kernel.py:22
//   KF_IMQ  (1 + D / h^2)^beta                     inverse multiquadric (beta < 0; not in the reference)
// KF_IMQ with beta - 1 gives the weights of its gradient term.
constexpr int KF_SE = 0, KF_IMQ = 1;

// K[i, j] = k(D_ij) for local rows i (global row row0 + i) x all columns
template <int KF>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gram_fn_kernel(const float *__restrict__ X, const float *__restrict__ r, int64_t n, int64_t ld,
               int64_t row0, float h2, float beta, float *__restrict__ K, int64_t ldk) {
    __shared__ GemmSmem gs;
    const int64_t m0 = row0 + (int64_t)blockIdx.y * TILE;   // global particle row of the tile
    const int64_t n0 = (int64_t)blockIdx.x * TILE;
    float acc[8][8];
    gemm_tile<true>(X, ld, m0, X, ld, n0, 0, (int)ld, gs, acc);
    float ri[8], rj[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        ri[q] = r[m0 + acc_row(q)];
        rj[q] = r[n0 + acc_col(q)];
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t i = m0 + acc_row(a);
        float out[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int64_t j = n0 + acc_col(c);
            const float dist = (ri[a] + rj[c]) - 2.0f * acc[a][c];
            // squared_exponential_kernel.py:22 -- exp(-D / square(bandwidth) / 2.)
            const float kv = KF == KF_SE ? expf(-dist / h2 / 2.0f) : powf(1.0f + fmaxf(dist, 0.0f) / h2, beta);
            out[c] = (i < n && j < n) ? kv : 0.0f;
        }
        float *dst = K + (i - row0) * ldk + n0;
        const int tx = threadIdx.x % 16;
        *reinterpret_cast<float4 *>(dst + tx * 4) = make_float4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<float4 *>(dst + 64 + tx * 4) = make_float4(out[4], out[5], out[6], out[7]);
    }
}

// ksum[i] = sum_j K[i, j]  (one warp per row, fixed order)
__global__ void rowsum_kernel(const float *__restrict__ K, int64_t rows, int64_t cols, int64_t ldk,
                              float *__restrict__ ksum) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float4 *p = reinterpret_cast<const float4 *>(K + row * ldk);
    float s = 0.0f;
    for (int64_t c4 = lane; c4 < cols / 4; c4 += 32) {
        const float4 v = p[c4];
        s += (v.x + v.y) + (v.z + v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) ksum[row] = s;
}

// O[rows x ldo] = K[rows x kdim] * Y[kdim x ldy]
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_nn_kernel(const float *__restrict__ K, int64_t ldk, int64_t kdim, const float *__restrict__ Y,
               int64_t ldy, float *__restrict__ O, int64_t ldo) {
    __shared__ GemmSmem gs;
    const int64_t m0 = (int64_t)blockIdx.y * TILE;
    const int64_t n0 = (int64_t)blockIdx.x * TILE;
    float acc[8][8];
    gemm_tile<false>(K, ldk, m0, Y, ldy, n0, ldy, (int)kdim, gs, acc);
    const int tx = threadIdx.x % 16;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        float *dst = O + (m0 + acc_row(a)) * ldo + n0;
        if (n0 + tx * 4 < ldo)
            *reinterpret_cast<float4 *>(dst + tx * 4) =
                make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
        if (n0 + 64 + tx * 4 < ldo)
            *reinterpret_cast<float4 *>(dst + 64 + tx * 4) =
                make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
    }
}

// Y = S - X / h2 over rows x ld (pad stays zero)
__global__ void make_y_kernel(const float *__restrict__ X, const float *__restrict__ S, int64_t count4,
                              float inv_h2, float *__restrict__ Y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count4) return;
    const float4 x = reinterpret_cast<const float4 *>(X)[i];
    const float4 s = reinterpret_cast<const float4 *>(S)[i];
    reinterpret_cast<float4 *>(Y)[i] = make_float4(s.x - x.x * inv_h2, s.y - x.y * inv_h2,
                                                   s.z - x.z * inv_h2, s.w - x.w * inv_h2);
}

// phi = (sum_slots O + x * ksum / h2) / n ; block partial of sum(phi^2) in double
__global__ void __launch_bounds__(256)
finalize_phi_kernel(const float *__restrict__ O, int64_t slot_stride, int nslots,
                    const float *__restrict__ ksum, int64_t ksum_slot_stride,
                    const float *__restrict__ X_local, int64_t rows_valid, int64_t rows, int64_t ld,
                    float inv_h2, float inv_n, float *__restrict__ phi,
                    double *__restrict__ partials) {
    const int64_t ld4 = ld / 4;
    const int64_t total4 = rows * ld4;
    double local = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total4;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / ld4;
        float4 o = reinterpret_cast<const float4 *>(O)[e];
        float ks = ksum[row];
        for (int s = 1; s < nslots; ++s) {
            const float4 o2 = reinterpret_cast<const float4 *>(O + s * slot_stride)[e];
            o.x += o2.x; o.y += o2.y; o.z += o2.z; o.w += o2.w;
            ks += ksum[row + s * ksum_slot_stride];
        }
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows_valid) {
            const float4 x = reinterpret_cast<const float4 *>(X_local)[e];
            const float w = ks * inv_h2;
            p.x = (o.x + x.x * w) * inv_n;
            p.y = (o.y + x.y * w) * inv_n;
            p.z = (o.z + x.z * w) * inv_n;
            p.w = (o.w + x.w * w) * inv_n;
        }
        reinterpret_cast<float4 *>(phi)[e] = p;
        local += (double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z + (double)p.w * p.w;
    }
    __shared__ double red[256];
    red[threadIdx.x] = local;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256)
reduce_partials_kernel(const double *__restrict__ partials, int count, double *__restrict__ out) {
    __shared__ double red[256];
    double local = 0.0;
    for (int i = threadIdx.x; i < count; i += 256) local += partials[i];
    red[threadIdx.x] = local;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

// dK = (x * ksum - KX) / h2     (squared_exponential_kernel.py:23,32)
__global__ void dk_kernel(const float *__restrict__ X, const float *__restrict__ KX,
                          const float *__restrict__ ksum, int64_t rows_valid, int64_t rows, int64_t ld,
                          float inv_h2, float *__restrict__ dK) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * ld) return;
    const int64_t row = e / ld;
    dK[e] = (row < rows_valid) ? (X[e] * ksum[row] - KX[e]) * inv_h2 : 0.0f;
}

// ---- host helpers shared with phi_tc.cu --------------------------------------
int launch_make_y(stein_ctx *ctx, const float *X, const float *S, int64_t rows, int64_t ld, float h2,
                  float *Y) {
    const int64_t count4 = rows * ld / 4;
    make_y_kernel<<<(unsigned)((count4 + 255) / 256), 256, 0, ctx->stream>>>(X, S, count4, 1.0f / h2, Y);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

int launch_finalize(stein_ctx *ctx, const float *O, int64_t slot_stride, int nslots, const float *ksum,
                    int64_t ksum_slot_stride, const float *X_local, int64_t rows_valid, int64_t rows,
                    int64_t ld, float h2, int64_t n_total, float *phi, double *partials,
                    double *sumsq) {
    const int64_t total4 = rows * ld / 4;
    const int blocks = (int)std::min<int64_t>((total4 + 255) / 256, FINALIZE_MAX_BLOCKS);
    finalize_phi_kernel<<<blocks, 256, 0, ctx->stream>>>(O, slot_stride, nslots, ksum, ksum_slot_stride,
                                                         X_local, rows_valid, rows, ld, 1.0f / h2,
                                                         1.0f / (float)n_total, phi, partials);
    STEIN_CHECK_LAUNCH(ctx);
    reduce_partials_kernel<<<1, 256, 0, ctx->stream>>>(partials, blocks, sumsq);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

// dense path workspace: [Y | O | ksum | partials | K slab]
static int64_t dense_slab_rows(int64_t rows_local, int64_t cols) {
    const int64_t budget = (int64_t)2 << 30;  // 2 GiB of K at a time
    int64_t slab = budget / (cols * 4) / TILE * TILE;
    slab = std::max<int64_t>(slab, TILE);
    return std::min(slab, rows_local);
}

int64_t dense_workspace_bytes(int64_t n_local, int64_t n_total, int64_t d) {
    const int64_t ld = stein_ld(d);
    const int64_t rows = stein_rows_padded(n_local), cols = stein_rows_padded(n_total);
    int64_t b = 0;
    b += cols * ld * 4;                       // Y
    b += rows * ld * 4;                       // O
    b += rows * 4;                            // ksum
    b += FINALIZE_MAX_BLOCKS * 8;             // partials
    b = round_up(b, 256);
    b += dense_slab_rows(rows, cols) * cols * 4;  // K slab
    return b + 1024;
}

int phi_dense(stein_ctx *ctx, const float *X_all, const float *S_all, const float *r_all,
              int64_t n_total, int64_t d, int64_t ld, int64_t row_begin, int64_t n_local, float h2,
              void *ws, int64_t ws_bytes, float *phi, double *sumsq) {
    const int64_t rows = stein_rows_padded(n_local), cols = stein_rows_padded(n_total);
    STEIN_REQUIRE(ctx, ws_bytes >= dense_workspace_bytes(n_local, n_total, d),
                  "phi workspace too small: %lld < %lld", (long long)ws_bytes,
                  (long long)dense_workspace_bytes(n_local, n_total, d));
    STEIN_REQUIRE(ctx, row_begin % TILE == 0, "row_begin must be a multiple of %d", TILE);
    char *p = (char *)ws;
    float *Y = (float *)p;  p += cols * ld * 4;
    float *O = (float *)p;  p += rows * ld * 4;
    float *ksum = (float *)p;  p += rows * 4;
    double *partials = (double *)p;  p += FINALIZE_MAX_BLOCKS * 8;
    p = (char *)ws + round_up(p - (char *)ws, 256);
    float *K = (float *)p;
    const int64_t slab = dense_slab_rows(rows, cols);

    STEIN_TRY(launch_make_y(ctx, X_all, S_all, cols, ld, h2, Y));
    RegionTimer timer(ctx, STEIN_REGION_PHI);   // gram_exp + rowsum + gemm_nn
    for (int64_t s0 = 0; s0 < rows; s0 += slab) {
        const int64_t sr = std::min(slab, rows - s0);
        dim3 g1((unsigned)(cols / TILE), (unsigned)(sr / TILE));
        gram_fn_kernel<KF_SE><<<g1, GEMM_THREADS, 0, ctx->stream>>>(X_all, r_all, n_total, ld, row_begin + s0, h2, 0.0f,
                                                                    K, cols);
        STEIN_CHECK_LAUNCH(ctx);
        rowsum_kernel<<<(unsigned)((sr + 7) / 8), 256, 0, ctx->stream>>>(K, sr, cols, cols, ksum + s0);
        STEIN_CHECK_LAUNCH(ctx);
        dim3 g2((unsigned)((ld + TILE - 1) / TILE), (unsigned)(sr / TILE));
        gemm_nn_kernel<<<g2, GEMM_THREADS, 0, ctx->stream>>>(K, cols, cols, Y, ld, O + s0 * ld, ld);
        STEIN_CHECK_LAUNCH(ctx);
    }
    timer.stop();
    const int64_t rows_valid = std::max<int64_t>(0, std::min<int64_t>(n_local, n_total - row_begin));
    return launch_finalize(ctx, O, 0, 1, ksum, 0, X_all + row_begin * ld, rows_valid, rows, ld, h2,
                           n_total, phi, partials, sumsq);
}

}  // namespace stein

using namespace stein;

extern "C" int stein_kernel_and_grad(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n,
                                     int64_t d, int64_t ld, float bandwidth, float *K_dev, int64_t ldk,
                                     float *dK_dev, void *ws, int64_t ws_bytes) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, X_dev && r_dev && K_dev && dK_dev && ws, "null pointer");
    STEIN_REQUIRE(ctx, ld >= d && ld % LD_ALIGN == 0, "bad ld");
    const int64_t rows = stein_rows_padded(n);
    STEIN_REQUIRE(ctx, ldk >= rows && ldk % 4 == 0, "ldk=%lld must be >= %lld and a multiple of 4",
                  (long long)ldk, (long long)rows);
    STEIN_REQUIRE(ctx, ws_bytes >= rows * ld * 4 + rows * 4, "workspace too small");
    float *KX = (float *)ws;
    float *ksum = KX + rows * ld;
    const float h2 = bandwidth * bandwidth;
    dim3 g1((unsigned)(rows / TILE), (unsigned)(rows / TILE));
    gram_fn_kernel<KF_SE><<<g1, GEMM_THREADS, 0, ctx->stream>>>(X_dev, r_dev, n, ld, 0, h2, 0.0f, K_dev, ldk);
    STEIN_CHECK_LAUNCH(ctx);
    rowsum_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(K_dev, rows, rows, ldk, ksum);
    STEIN_CHECK_LAUNCH(ctx);
    dim3 g2((unsigned)((ld + TILE - 1) / TILE), (unsigned)(rows / TILE));
    gemm_nn_kernel<<<g2, GEMM_THREADS, 0, ctx->stream>>>(K_dev, ldk, rows, X_dev, ld, KX, ld);
    STEIN_CHECK_LAUNCH(ctx);
    const int64_t total = rows * ld;
    dk_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(X_dev, KX, ksum, n, rows, ld,
                                                                       1.0f / h2, dK_dev);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}


// Inverse multiquadric kernel through the same plugin point:
//   K_ij  = (1 + D_ij / h^2)^beta,  beta < 0
//   dK_i  = -0.5 d(sum K)/d theta_i   (the reference's recipe, squared_exponential_kernel.py:23,32)
//         = (-2 beta / h^2) (x_i sum_j G_ij - sum_j G_ij x_j),   G_ij = (1 + D_ij / h^2)^(beta - 1)
// Dense (small n), like stein_kernel_and_grad: G is formed in the K buffer first, then K itself.
extern "C" int stein_imq_kernel_and_grad(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n,
                                         int64_t d, int64_t ld, float bandwidth, float beta, float *K_dev,
                                         int64_t ldk, float *dK_dev, void *ws, int64_t ws_bytes) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, X_dev && r_dev && K_dev && dK_dev && ws, "null pointer");
    STEIN_REQUIRE(ctx, ld >= d && ld % LD_ALIGN == 0, "bad ld");
    STEIN_REQUIRE(ctx, beta < 0.0f && beta > -64.0f, "beta must be negative");
    STEIN_REQUIRE(ctx, bandwidth > 0.0f, "bandwidth must be positive");
    const int64_t rows = stein_rows_padded(n);
    STEIN_REQUIRE(ctx, ldk >= rows && ldk % 4 == 0, "ldk=%lld must be >= %lld and a multiple of 4",
                  (long long)ldk, (long long)rows);
    STEIN_REQUIRE(ctx, ws_bytes >= rows * ld * 4 + rows * 4, "workspace too small");
    float *GX = (float *)ws;
    float *gsum = GX + rows * ld;
    const float h2 = bandwidth * bandwidth;
    dim3 g1((unsigned)(rows / TILE), (unsigned)(rows / TILE));
    gram_fn_kernel<KF_IMQ><<<g1, GEMM_THREADS, 0, ctx->stream>>>(X_dev, r_dev, n, ld, 0, h2, beta - 1.0f, K_dev, ldk);
    STEIN_CHECK_LAUNCH(ctx);
    rowsum_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(K_dev, rows, rows, ldk, gsum);
    STEIN_CHECK_LAUNCH(ctx);
    dim3 g2((unsigned)((ld + TILE - 1) / TILE), (unsigned)(rows / TILE));
    gemm_nn_kernel<<<g2, GEMM_THREADS, 0, ctx->stream>>>(K_dev, ldk, rows, X_dev, ld, GX, ld);
    STEIN_CHECK_LAUNCH(ctx);
    const int64_t total = rows * ld;
    dk_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(X_dev, GX, gsum, n, rows, ld, -2.0f * beta / h2,
                                                                       dK_dev);
    STEIN_CHECK_LAUNCH(ctx);
    gram_fn_kernel<KF_IMQ><<<g1, GEMM_THREADS, 0, ctx->stream>>>(X_dev, r_dev, n, ld, 0, h2, beta, K_dev, ldk);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

namespace stein {
// phi = (K S + dK) / n, block partial of sum(phi^2)
__global__ void __launch_bounds__(256)
phi_from_kernel_finalize(const float *__restrict__ KS, const float *__restrict__ dK, int64_t rows_valid, int64_t rows,
                         int64_t ld, float inv_n, float *__restrict__ phi, double *__restrict__ partials) {
    double local = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * ld; e += (int64_t)gridDim.x * blockDim.x) {
        const float v = (e / ld) < rows_valid ? (KS[e] + dK[e]) * inv_n : 0.0f;
        phi[e] = v;
        local += (double)v * v;
    }
    __shared__ double red[256];
    red[threadIdx.x] = local;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}
}  // namespace stein

// compute_phi for ANY kernel operator (abstract_stein_sampler.py:100-105): K (n x n, device, leading
// dimension ldk, pad rows / columns zero) and dK (n x ld) as an AbstractKernel.kernel_and_grad returned
// them, S the scores: phi = (K S + dK) / n on the device, sum(phi^2) for the clip.
extern "C" int stein_phi_from_kernel(stein_ctx *ctx, const float *K_dev, int64_t ldk, const float *dK_dev,
                                     const float *S_dev, int64_t n, int64_t d, int64_t ld, void *ws,
                                     int64_t ws_bytes, float *phi_dev, double *sumsq_dev) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, K_dev && dK_dev && S_dev && ws && phi_dev && sumsq_dev, "null pointer");
    STEIN_REQUIRE(ctx, n >= 1 && d >= 1 && ld >= d && ld % LD_ALIGN == 0, "bad shape");
    const int64_t rows = stein_rows_padded(n);
    STEIN_REQUIRE(ctx, ldk >= rows && ldk % 4 == 0, "ldk=%lld must be >= %lld and a multiple of 4",
                  (long long)ldk, (long long)rows);
    STEIN_REQUIRE(ctx, ws_bytes >= rows * ld * 4 + FINALIZE_MAX_BLOCKS * 8 + 8, "workspace too small");
    float *KS = (float *)ws;
    double *partials = (double *)(((uintptr_t)(KS + rows * ld) + 7) & ~(uintptr_t)7);
    dim3 g2((unsigned)((ld + TILE - 1) / TILE), (unsigned)(rows / TILE));
    gemm_nn_kernel<<<g2, GEMM_THREADS, 0, ctx->stream>>>(K_dev, ldk, rows, S_dev, ld, KS, ld);
    STEIN_CHECK_LAUNCH(ctx);
    const int blocks = (int)std::min<int64_t>((rows * ld + 255) / 256, FINALIZE_MAX_BLOCKS);
    phi_from_kernel_finalize<<<blocks, 256, 0, ctx->stream>>>(KS, dK_dev, n, rows, ld, 1.0f / (float)n, phi_dev, partials);
    STEIN_CHECK_LAUNCH(ctx);
    reduce_partials_kernel<<<1, 256, 0, ctx->stream>>>(partials, blocks, sumsq_dev);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}
