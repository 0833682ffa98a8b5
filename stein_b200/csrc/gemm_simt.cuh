// gemm_simt.cuh -- FP32 FFMA 128x128x16 tile mainloop shared by the exact
// distance sweep (median) and the dense phi path.
//
// The accumulation order of every output element is k = 0, 1, 2, ... with one
// fmaf per k starting from +0: this is the "contract arithmetic" that
// oracle/svgd_oracle.c restates, so results are bit-reproducible on the CPU.
#pragma once
#include "common.cuh"

namespace stein {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int GEMM_THREADS = 256;
constexpr int SPAD = 4;  // smem row padding (floats): keeps float4 alignment, spreads banks

struct __align__(16) GemmSmem {
    float A[2][BK][BM + SPAD];
    float B[2][BK][BN + SPAD];
};

// global row of accumulator row r (0..7) for this thread, relative to the tile
__device__ __forceinline__ int acc_row(int r) {
    const int ty = threadIdx.x / 16;
    return (r < 4) ? (ty * 4 + r) : (64 + ty * 4 + (r - 4));
}
__device__ __forceinline__ int acc_col(int c) {
    const int tx = threadIdx.x % 16;
    return (c < 4) ? (tx * 4 + c) : (64 + tx * 4 + (c - 4));
}

// acc = A[m0:m0+128, 0:K] * op(B)   (K % 16 == 0)
//   kBNT = true : B is N x K row-major (K contiguous), rows n0..n0+127        -> A * B^T
//   kBNT = false: B is K x N row-major (N contiguous), cols n0..n0+127 (< ncolsB guarded)
// All A rows / B rows touched must be allocated (padding contract).
template <bool kBNT>
__device__ __forceinline__ void gemm_tile(const float *__restrict__ A, int64_t lda, int64_t m0,
                                          const float *__restrict__ B, int64_t ldb, int64_t n0,
                                          int64_t ncolsB, int K, GemmSmem &sm,
                                          float (&acc)[8][8]) {
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;

    // global -> register staging
    const int a_row = tid / 4, a_kq = tid % 4;      // A (and NT-B): 64 rows x 4 float4 per pass
    const int b_krow = tid / 32, b_c4 = tid % 32;   // NN-B: 8 k-rows x 32 float4 per pass
    float4 ra[2], rb[2];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load_global = [&](int k0) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            ra[p] = *reinterpret_cast<const float4 *>(A + (m0 + a_row + 64 * p) * lda + k0 + 4 * a_kq);
            if (kBNT) {
                rb[p] = *reinterpret_cast<const float4 *>(B + (n0 + a_row + 64 * p) * ldb + k0 + 4 * a_kq);
            } else {
                const int64_t col = n0 + 4 * b_c4;
                rb[p] = (col < ncolsB)
                            ? *reinterpret_cast<const float4 *>(B + (int64_t)(k0 + b_krow + 8 * p) * ldb + col)
                            : zero4;
            }
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int row = a_row + 64 * p;
            sm.A[buf][4 * a_kq + 0][row] = ra[p].x;
            sm.A[buf][4 * a_kq + 1][row] = ra[p].y;
            sm.A[buf][4 * a_kq + 2][row] = ra[p].z;
            sm.A[buf][4 * a_kq + 3][row] = ra[p].w;
            if (kBNT) {
                sm.B[buf][4 * a_kq + 0][row] = rb[p].x;
                sm.B[buf][4 * a_kq + 1][row] = rb[p].y;
                sm.B[buf][4 * a_kq + 2][row] = rb[p].z;
                sm.B[buf][4 * a_kq + 3][row] = rb[p].w;
            } else {
                *reinterpret_cast<float4 *>(&sm.B[buf][b_krow + 8 * p][4 * b_c4]) = rb[p];
            }
        }
    };

    const int nkt = K / BK;
    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) load_global((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&sm.A[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&sm.A[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&sm.B[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&sm.B[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = __fmaf_rn(a[r], b[c], acc[r][c]);
        }
        if (kt + 1 < nkt) store_smem(buf ^ 1);
        __syncthreads();
    }
}

}  // namespace stein
