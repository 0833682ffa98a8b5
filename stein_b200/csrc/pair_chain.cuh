// pair_chain.cuh -- contract-arithmetic squared distances of a LIST of (i, j) pairs.
//
// Used by the median (stein/utilities/compute_median.py:4-16 over the D of
// stein/kernels/abstract_kernel.py:33-35) in two places: the pilot sample and the exact
// recomputation of the pairs the tensor-core filter could not decide.  The value is the
// contract one of oracle/svgd_oracle.c:  g = fma chain over k ascending from +0,
// D = fl(fl(r_i + r_j) - 2 g)  ->  order-preserving u32 key.
//
// The chain of one pair is strictly sequential in k, so a thread owns a pair.  What is
// shared is the LOADING: a warp takes 32 pairs and moves their 64 rows through shared
// memory 32 columns at a time, every global load instruction covering four complete
// 128-byte lines (8 lanes x 16 B per row).  The one-thread-one-row form this replaces touched
// 32 different lines per load instruction and was bound by the L1 tag stage (99 % l1tex).
// Shared layout per warp: A[p][k] / B[p][k] (p = pair, k = column in the chunk) with a row
// stride of 36 floats: the 16-byte stores of a quarter warp (one row, 8 consecutive float4)
// and the 16-byte reads of a quarter warp (8 rows, same float4 index) both cover all 32 banks
// exactly once, so every shared-memory access is a conflict-free 128-bit one.
#pragma once
#include "common.cuh"

namespace stein {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

constexpr int PC_WARPS = 4;                       // warps per block
constexpr int PC_CHUNK = 32;                      // columns staged per step (one 128-byte line per row)
constexpr int PC_ROW = PC_CHUNK + 4;              // padded row stride (floats)
constexpr int PC_SMEM_PER_WARP = 2 * 32 * PC_ROW * 4;

// SRC 0: io = uint2 (i, jw) list, rewritten in place as (key, weight)   [band of median_tc.cu]
// SRC 1: pairs drawn from splitmix64(seed + s), io = u32 keys[s]         [pilot of median.cu]
//        (seed already includes the index of the first sample of this launch)
template <int SRC>
__global__ void __launch_bounds__(PC_WARPS * 32, 4)
pair_chain_kernel(void *__restrict__ io, unsigned long long m, const unsigned long long *__restrict__ m_dev,
                  const float *__restrict__ X, const float *__restrict__ r, int64_t n, int64_t ld, uint64_t seed) {
    // the list length may only be known on the device (m is then its upper bound)
    if (m_dev) m = min(m, *m_dev);
    extern __shared__ __align__(16) unsigned char pc_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *sA = reinterpret_cast<float *>(pc_smem + (size_t)warp * PC_SMEM_PER_WARP);
    float *sB = sA + 32 * PC_ROW;
    const unsigned long long ngroups = (m + 31ull) / 32ull;
    // contiguous group range per block (neighbouring list entries share rows: L1 reuse)
    const unsigned long long g0 = ngroups * blockIdx.x / gridDim.x, g1 = ngroups * (blockIdx.x + 1) / gridDim.x;
    const int sub = lane >> 3, f = lane & 7;      // row slot within a load instruction, float4 within the line
    const int nchunks = (int)(ld / PC_CHUNK);
    for (unsigned long long g = g0 + warp; g < g1; g += PC_WARPS) {
        const unsigned long long e = g * 32ull + lane;
        const bool valid = e < m;
        uint32_t i = 0u, j = 0u, w = 1u;
        if (valid) {
            if (SRC == 0) {
                const uint2 ij = reinterpret_cast<const uint2 *>(io)[e];
                i = ij.x;
                j = ij.y & 0x7fffffffu;
                w = (ij.y >> 31) ? 2u : 1u;
            } else {
                const uint64_t h = splitmix64(seed + (uint64_t)e);
                i = (uint32_t)((h >> 32) % (uint64_t)n);
                j = (uint32_t)((h & 0xffffffffull) % (uint64_t)n);
            }
        }
        // row handled by this lane in load instruction t: pair (4 t + sub) & 31, i-row for t < 8
        const float4 *rowp[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const uint32_t idx = __shfl_sync(0xffffffffu, t < 8 ? i : j, (4 * t + sub) & 31);
            rowp[t] = reinterpret_cast<const float4 *>(X + (size_t)idx * ld) + f;
        }
        float4 buf[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) buf[t] = __ldg(rowp[t]);
        float acc = 0.0f;
        for (int c = 0; c < nchunks; ++c) {
            __syncwarp();                          // the previous chunk has been consumed
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                *reinterpret_cast<float4 *>((t < 8 ? sA : sB) + ((4 * t + sub) & 31) * PC_ROW + 4 * f) = buf[t];
            }
            if (c + 1 < nchunks) {
#pragma unroll
                for (int t = 0; t < 16; ++t) buf[t] = __ldg(rowp[t] + (c + 1) * (PC_CHUNK / 4));
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < PC_CHUNK / 4; ++q) {
                const float4 a = *reinterpret_cast<const float4 *>(sA + lane * PC_ROW + 4 * q);
                const float4 b = *reinterpret_cast<const float4 *>(sB + lane * PC_ROW + 4 * q);
                acc = __fmaf_rn(a.x, b.x, acc);
                acc = __fmaf_rn(a.y, b.y, acc);
                acc = __fmaf_rn(a.z, b.z, acc);
                acc = __fmaf_rn(a.w, b.w, acc);
            }
        }
        if (valid) {
            const float tsum = r[i] + r[j];
            const uint32_t key = float_to_key(tsum - 2.0f * acc);
            if (SRC == 0) reinterpret_cast<uint2 *>(io)[e] = make_uint2(key, w);
            else reinterpret_cast<uint32_t *>(io)[e] = key;
        }
    }
}

template <int SRC>
static int launch_pair_chain(stein_ctx *ctx, void *io, unsigned long long m, const float *X, const float *r,
                             int64_t n, int64_t ld, uint64_t seed, const unsigned long long *m_dev = nullptr) {
    if (m == 0) return STEIN_OK;
    const size_t smem = (size_t)PC_WARPS * PC_SMEM_PER_WARP;
    const unsigned long long ngroups = (m + 31ull) / 32ull;
    // 4 blocks of 4 warps per SM are resident (register-limited: 128 per thread; 36 KB of shared memory each)
    const unsigned grid = (unsigned)std::min<unsigned long long>((ngroups + PC_WARPS - 1) / PC_WARPS,
                                                                 4ull * (unsigned long long)ctx->num_sms);
    pair_chain_kernel<SRC><<<grid, PC_WARPS * 32, smem, ctx->stream>>>(io, m, m_dev, X, r, n, ld, seed);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

}  // namespace stein
