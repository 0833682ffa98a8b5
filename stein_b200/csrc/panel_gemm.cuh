// panel_gemm.cuh -- K-streaming tcgen05 GEMM main loop for particle matrices of more than 256
// coordinates (leading dimension 512 / 768 / 1024: BASELINE.json config E), shared by
//   * the phi panel kernels (phi_panel.cuh):  P = exp(-D / 2h^2) for a block of rows x columns,
//     then O += P Y  -- stein/kernels/squared_exponential_kernel.py:22-35 and
//     stein/samplers/abstract_stein_sampler.py:100-105, the same algebra as phi_tc.cu;
//   * the median's filter sweep for d > 256 (median_tc.cu, sweep3) --
//     stein/kernels/abstract_kernel.py:33-35 + stein/utilities/compute_median.py:4-16.
//
// Why not the flash kernel of phi_tc.cu: it keeps a row tile's A operand (256 x d) in shared memory
// and the O accumulator (128 x d fp32 per CTA) in tensor memory; neither fits beyond d = 256
// (O alone would need 1024 of the 512 TMEM columns).  Here BOTH operands stream through a TMA ring
// along K and the accumulator is a plain 256 x 256 tile:
//   cluster of two CTAs (cta_group::2): C tile of 256 rows x 256 columns, each CTA holds its 128 rows
//   x 256 fp32 columns in tensor memory, twice (512 columns: the epilogue of one accumulation unit
//   overlaps the MMAs of the next);
//   ring stage = one 16 KB A box (this CTA's 128 rows x 128 bytes of K) + one 16 KB B box (this
//   CTA's half, 128 of the tile's 256 columns x 128 bytes of K), consumed by four M = 256, N = 256
//   MMAs (K = 16 sixteen-bit or K = 32 eight-bit elements each);
//   the products are split exactly like in phi_tc.cu: a STAGE TABLE lists, per group of 128 K
//   elements, which operand arrays meet in which kind of MMA (fast: X16.X16 twice, a8l.b8h, a8h.b8l;
//   precise / median: hi.hi, lo.hi, hi.lo per 64 elements).
// Per stage an SM takes in 32 KB for 512 tensor-pipe cycles = 64 B/cycle, the operand diet of a
// plain BF16 GEMM with this tile (the L2 -> SM fabric, ~43 B/cycle/SM with all SMs pulling, is what
// bounds it -- as it bounds the cuBLAS figure the roofline is quoted against).
//
// Roles (384 threads, as in phi_tc.cu): warp 0 TMA producer (both CTAs), warp 1 MMA issuer
// (leader CTA), warps 4-11 epilogue (two warpgroups; warp w may touch TMEM lanes 32 (w % 4) ..).
// A Policy supplies the tile enumeration and the epilogue.
#pragma once
#include "tc_common.cuh"

namespace stein {
namespace pg {

using namespace tc;

constexpr int THREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr int EPI_WARPS = 8;
constexpr uint32_t BOX_BYTES = 128 * 128;        // one TMA box: 128 rows x 128 bytes
constexpr uint32_t STAGE_BYTES = 2 * BOX_BYTES;  // A box + B box
constexpr uint32_t TMEM_COLS = 512;
constexpr int MAX_TABLE = 6;
constexpr int MAX_MAPS = 3;

struct Stage {
    int a, b;     // operand arrays (indices into Maps::a / Maps::b)
    int koff;     // element offset inside the group of 128 K elements (0 or 64)
    int f8;       // 1: kind::f8f6f4 (A e4m3, B e5m2, 128 elements per box), 0: kind::f16 (64 elements per box)
};

struct Maps {
    CUtensorMap a[MAX_MAPS], b[MAX_MAPS];
};

// What the main loop needs to know; a Policy's Params derives from it.
struct Core {
    int n_stage;               // entries of `table` per group of 128 K elements
    Stage table[MAX_TABLE];
    int groups_per_unit;       // K groups accumulated in tensor memory before the epilogue sees them
    int units_per_tile;
    int ka0, kb0;              // first K element of unit 0 in the A / B arrays
    int a_row0, b_row0;        // A rows of tile row 0 / B rows of tile column 0 (in the arrays' own row numbering)
    const int *route;          // device-side route word (phi_guard_kernel) or NULL
    int my_route;              // this launch runs only when *route == my_route
    // L2 eviction priority of the A / B operand loads (tc::L2_EVICT_*; 0 = normal): an operand that the launch
    // reads many times should outlive the streams that pass through L2 once
    unsigned long long pol_a, pol_b;
    // A operand stored BOX-MAJOR (the P block of phi_panel.cuh): box (row tile rt of 128 rows, K block kb) is one
    // contiguous 16 KB piece [128 rows][128 bytes] at row (rt * nkb + kb) * 128 of a [rows][128 B] array, with
    // nkb = a_nkb16 K blocks of 64 two-byte elements or a_nkb8 blocks of 128 one-byte elements per row tile.
    // 0: plain row-major A (coordinates = K element, row).
    int a_blocked, a_nkb16, a_nkb8;
};

struct Barriers {
    uint64_t full[8], empty[8];
    uint64_t acc_full[2], acc_empty[2];
};

template <int STAGES>
constexpr size_t smem_bytes(size_t tail) {
    return 1024 + (size_t)STAGES * STAGE_BYTES + 256 + tail;
}

// Policy interface:
//   struct Params : Core { ... };
//   static constexpr int STAGES; static constexpr size_t TAIL_BYTES;
//   __device__ static int tile(const Params &, long long k, int cluster, int nclusters, int &ti, int &tj);
//        step k of this cluster: 0 = no more steps, 1 = tile (ti, tj), 2 = nothing to do at this step (the
//        cluster's cell of the window lies outside the problem); A rows = a_row0 + 256 ti + 128 rank,
//        B rows = b_row0 + 256 tj + 128 rank
//   struct Epilogue { __device__ Epilogue(const Params &, uint8_t *tail, int warp, int lane, uint32_t rank);
//        __device__ void tile_begin(int ti, int tj);
//        __device__ void unit(uint32_t acc_tmem, int u, bool last);     // all 8 epilogue warps
//        __device__ void finish(); };
template <class Policy>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
panel_gemm_kernel(const __grid_constant__ Maps maps, const typename Policy::Params p) {
    if (p.route != nullptr && *p.route != p.my_route) return;     // uniform: nothing has been set up yet
    constexpr int STAGES = Policy::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sRing = smem;                                       // STAGES x [A box | B box]
    uint8_t *tail0 = sRing + (size_t)STAGES * STAGE_BYTES;
    Barriers *bars = reinterpret_cast<Barriers *>(tail0);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail0 + 224);
    uint8_t *tail = tail0 + 256;                                 // the Policy's own shared memory

    const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int ncl = (int)gridDim.x / 2, cl = (int)blockIdx.x / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->full[s], 1);           // leader: one expect_tx arrival, bytes from both CTAs
            mbar_init(&bars->empty[s], 1);          // multicast commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars->acc_full[b], 1);                  // multicast commit
            mbar_init(&bars->acc_empty[b], 2 * EPI_WARPS);     // leader: one arrival per epilogue warp of both CTAs
        }
        fence_barrier_init();
        fence_proxy_async();
        for (int m = 0; m < MAX_MAPS; ++m) {
            tma_prefetch_desc(&maps.a[m]);
            tma_prefetch_desc(&maps.b[m]);
        }
    }
    Policy::init_shared(tail, threadIdx.x);
    if (warp == 1) tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();            // peer barriers are initialised before anyone signals them
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 0) {
            // ===================== TMA producer (both CTAs) =====================
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full0_addr = mapa_shared(smem_u32(&bars->full[0]), 0);
            const uint64_t pol_a = p.pol_a ? p.pol_a : L2_EVICT_NORMAL, pol_b = p.pol_b ? p.pol_b : L2_EVICT_NORMAL;
            int ti, tj;
            for (long long k = 0, rc; (rc = Policy::tile(p, k, cl, ncl, ti, tj)) != 0; ++k) {
                if (rc == 2) continue;
                const int arow = p.a_row0 + ti * 256 + (int)rank * 128;
                const int brow = p.b_row0 + tj * 256 + (int)rank * 128;
                const int ngroups = p.units_per_tile * p.groups_per_unit;
                for (int g = 0; g < ngroups; ++g) {
#pragma unroll 1
                    for (int s = 0; s < p.n_stage; ++s) {
                        const Stage st = p.table[s];
                        mbar_wait(&bars->empty[stage], phase ^ 1, 2);      // local: multicast commit of the leader
                        if (elect_one_sync()) {
                            if (leader) mbar_expect_tx(&bars->full[stage], 2u * STAGE_BYTES);
                            uint8_t *dst = sRing + (size_t)stage * STAGE_BYTES;
                            const uint32_t fa = full0_addr + 8u * (uint32_t)stage;
                            if (p.a_blocked) {
                                const int kk = p.ka0 + g * 128 + st.koff;
                                const int rt = ti * 2 + (int)rank;
                                const int orow = st.f8 ? (rt * p.a_nkb8 + (kk >> 7)) * 128 : (rt * p.a_nkb16 + (kk >> 6)) * 128;
                                tma_load_2d_pair_hint(dst, &maps.a[st.a], fa, 0, orow, pol_a);
                            } else {
                                tma_load_2d_pair_hint(dst, &maps.a[st.a], fa, p.ka0 + g * 128 + st.koff, arow, pol_a);
                            }
                            tma_load_2d_pair_hint(dst + BOX_BYTES, &maps.b[st.b], fa, p.kb0 + g * 128 + st.koff, brow, pol_b);
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else if (warp == 1 && leader) {
            // ===================== MMA issuer (leader CTA only) =====================
            // whole warp in uniform control flow, one elected lane issues (tc_common.cuh elect_one_sync)
            const uint32_t idesc16 = make_idesc(FMT_F16, 256, 256);
            const uint32_t idesc8 = make_idesc_ab(FMT8_E4M3, FMT8_E5M2, 256, 256);
            int stage = 0;
            uint32_t phase = 0;
            long long uc = 0;      // accumulation units issued so far (TMEM buffer = uc & 1)
            int ti, tj;
            for (long long k = 0, rc; (rc = Policy::tile(p, k, cl, ncl, ti, tj)) != 0; ++k) {
                if (rc == 2) continue;
                for (int u = 0; u < p.units_per_tile; ++u, ++uc) {
                    const int b = (int)(uc & 1);
                    if (uc >= 2) {
                        mbar_wait(&bars->acc_empty[b], (uint32_t)(((uc >> 1) - 1) & 1), 7);
                        tcgen05_fence_after();
                    }
                    const uint32_t d_tmem = tmem + (uint32_t)b * 256u;
                    for (int g = 0; g < p.groups_per_unit; ++g) {
#pragma unroll 1
                        for (int s = 0; s < p.n_stage; ++s) {
                            const int f8 = p.table[s].f8;
                            mbar_wait(&bars->full[stage], phase, 4);
                            tcgen05_fence_after();
                            const uint32_t base = smem_u32(sRing + (size_t)stage * STAGE_BYTES);
                            const uint64_t ad = make_kmajor_sw128_desc(base), bd = make_kmajor_sw128_desc(base + BOX_BYTES);
                            const uint32_t first = (g | s) != 0;
                            if (elect_one_sync()) {
                                if (f8) {
#pragma unroll
                                    for (int k4 = 0; k4 < 4; ++k4)
                                        umma2_f8_ss(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc8, (first | k4) != 0);
                                } else {
#pragma unroll
                                    for (int k4 = 0; k4 < 4; ++k4)
                                        umma2_f16_ss(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc16, (first | k4) != 0);
                                }
                                tcgen05_commit_pair(&bars->empty[stage]);
                                if (g == p.groups_per_unit - 1 && s == p.n_stage - 1) tcgen05_commit_pair(&bars->acc_full[b]);
                            }
                            __syncwarp();
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ===================== epilogue warpgroups (both CTAs, own 128 rows) =====================
        typename Policy::Epilogue epi(p, tail, warp, lane, rank);
        const uint32_t acc_empty_addr0 = mapa_shared(smem_u32(&bars->acc_empty[0]), 0);
        const uint32_t acc_empty_addr1 = mapa_shared(smem_u32(&bars->acc_empty[1]), 0);
        long long uc = 0;
        int ti, tj;
        for (long long k = 0, rc; (rc = Policy::tile(p, k, cl, ncl, ti, tj)) != 0; ++k) {
                if (rc == 2) continue;
            epi.tile_begin(ti, tj);
            for (int u = 0; u < p.units_per_tile; ++u, ++uc) {
                const int b = (int)(uc & 1);
                mbar_wait(&bars->acc_full[b], (uint32_t)((uc >> 1) & 1), 8);
                tcgen05_fence_after();
                epi.unit(tmem + (uint32_t)b * 256u, u, u == p.units_per_tile - 1);
                // buffer b may be overwritten by the MMAs of unit uc + 2 (one arrival per warp)
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(b ? acc_empty_addr1 : acc_empty_addr0);
            }
        }
        epi.finish();
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();            // no CTA leaves while its partner may still signal / read it
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_pair(tmem, TMEM_COLS);
    }
}

}  // namespace pg
}  // namespace stein
