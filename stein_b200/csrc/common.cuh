// common.cuh -- context object, error plumbing, shared device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/stein_b200.h"

namespace stein {

constexpr int TILE = 128;          // particle tile edge (rows/cols of one distance tile)
constexpr int LD_ALIGN = 32;       // leading dimension granularity (floats)
constexpr int HIST_MAX_BINS = 16384;
constexpr int NUM_SMS_B200 = 148;

}  // namespace stein

namespace stein {
struct PeerReduce;   // peer_reduce.cu
}

struct stein_ctx {
    int device = 0;
    int num_sms = stein::NUM_SMS_B200;
    cudaStream_t stream = nullptr;
    bool has_comm = false;
    stein_comm comm{};
    void *nccl_state = nullptr;     // built-in NCCL transport (comm_nccl.cu), if initialised
    // installed by an engine whose peers are open: the small all-reduces then run over NVLink peer
    // memory (peer_reduce.cu) instead of the hooks above
    stein::PeerReduce *peer_reduce = nullptr;
    int phi_impl = STEIN_PHI_AUTO;
    int median_impl = STEIN_MEDIAN_AUTO;
    // set by an engine around its median call: identifies the sequence of calls whose medians move
    // slowly, so that the previous window may be used as a hint (NULL: independent call, no hint)
    const void *median_owner = nullptr;
    // Work that does not depend on the bandwidth: the median calls this hook right before its
    // host round trip, so the GPU has kernels queued while the host reads the counters.
    int (*presync_fn)(void *) = nullptr;
    void *presync_arg = nullptr;
    // set by flash_tc2_prepare_x, consumed by the next phi call on the same problem (phi_tc.cu)
    struct XPrep {
        const void *X, *ws;
        int64_t n_total, n_local, d;
        int mode;
    } xprep{nullptr, nullptr, 0, 0, 0, 0};
    // set by flash_tc2_prepare_s: the column maxima of these scores (and of the prepared X) are in the workspace
    const void *sprep_S = nullptr, *sprep_ws = nullptr;
    // phi route guard (phi_tc.cu): device / pinned words [kappa, max centred norm^2] of the last guarded
    // phi call, the event that marks their copy, the route that call took (0 fast, 1 precise, 2 FP32
    // FFMA; -1 none yet) and the predicted error up to which a faster route is taken
    float *d_guard = nullptr, *h_guard = nullptr;       // two slots of 4 floats each, alternating per call
    cudaEvent_t ev_guard[2] = {nullptr, nullptr};
    uint64_t guard_calls = 0;
    // Set by an engine around its phi call: within one engine's iteration sequence the cloud moves slowly,
    // so the route may be decided from the PREVIOUS iteration's kappa (already on the host: no round trip)
    // while this iteration's value travels for the next one.  guard_lag_owner: whose value the other slot
    // holds (NULL: nobody's -- the next guarded call waits for its own value).
    const void *guard_owner = nullptr, *guard_lag_owner = nullptr;
    // set around a median call whose host part is collected later (median_sqdist_begin / _resume)
    int median_defer = 0;
    // set by an engine around a phi call that runs BEFORE the host has collected the deferred median: the kernels
    // take the bandwidth from this device block (median_sqdist_device_bandwidth) instead of the host argument
    const float *dev_bw = nullptr;
    int last_route = -1;
    float last_kappa = 0.0f, last_pred_fast = 0.0f;
    float phi_guard_tol = 5.0e-5f;
    int64_t launches = 0;
    std::string error;
    // pinned host staging + device scratch for the median loop
    uint64_t *h_counts = nullptr;   // pinned, HIST_MAX_BINS + 2
    uint64_t *d_counts = nullptr;   // device, HIST_MAX_BINS + 2
    uint32_t *d_pilot_keys = nullptr;
    uint32_t *d_sel = nullptr;      // device, 2 keys
    uint32_t *h_sel = nullptr;      // pinned
    int64_t pilot_cap = 0;
    // optional region timing (bench.py): event pairs recorded on `stream`
    int profile = 0;                // 0 off, 1 = regions PHI / SWEEP only, 2 = all regions (timeline)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[STEIN_REGION_COUNT];
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pool;
    // optional fine-grained timeline (tools/step_trace.py): labelled events on `stream`, read by stein_ctx_trace_read
    int trace = 0;
    std::vector<std::pair<const char *, cudaEvent_t>> trace_marks;
    std::vector<cudaEvent_t> trace_pool;
};

namespace stein {

extern thread_local std::string g_last_error;

int fail(stein_ctx *ctx, int code, const char *fmt, ...);
void nccl_release(stein_ctx *ctx);   // comm_nccl.cu
// in-place sums across the ranks, ordered on the ctx stream (peer_reduce.cu)
int allreduce_u64(stein_ctx *ctx, void *buf_dev, int64_t count);
int allreduce_f64(stein_ctx *ctx, void *buf_dev, int64_t count);

// Median of the squared distances in two halves (median.cu): _begin enqueues the device part when the pilot-less
// steady state applies (MEDIAN_DEFERRED) or does nothing (MEDIAN_NOT_DEFERRED); _resume collects the result.
constexpr int MEDIAN_FULL = 0, MEDIAN_BEGIN = 1, MEDIAN_RESUME = 2;
constexpr int MEDIAN_DEFERRED = 3, MEDIAN_NOT_DEFERRED = 4;
int median_sqdist_begin(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n, int64_t d, int64_t ld);
int median_sqdist_resume(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n, int64_t d, int64_t ld,
                         float *median_host, int32_t *sweeps_host);
bool median_sqdist_deferred_pending(const void *owner);
// device-side result of a deferred median: [h, h^2, 1/h^2, log2(e)/h^2, log2(e)/(2 h^2)] for the kernels that follow it
// on the stream; after _resume: whether that result is the one the host accepted (and its h)
const float *median_sqdist_device_bandwidth(void);
bool median_sqdist_device_select_valid(float *bandwidth);
void median_tc_forget_owner(const void *owner);     // an engine is destroyed: drop its median history
bool median_sqdist_can_defer(const stein_ctx *ctx, int64_t n, int64_t ld);
bool median_sqdist_wants_fused_begin(const stein_ctx *ctx, int64_t n, int64_t ld);
int median_sqdist_begin_with_norms(stein_ctx *ctx, const float *X_dev, float *r_dev, int64_t rows_r, int64_t n, int64_t ld);

#define STEIN_CHECK_CUDA(ctx, expr)                                                        \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return stein::fail((ctx), STEIN_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, \
                               #expr, cudaGetErrorString(_e));                             \
    } while (0)

#define STEIN_CHECK_LAUNCH(ctx)                                                            \
    do {                                                                                   \
        (ctx)->launches++;                                                                 \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess)                                                             \
            return stein::fail((ctx), STEIN_ERR_CUDA, "%s:%d: kernel launch -> %s", __FILE__, \
                               __LINE__, cudaGetErrorString(_e));                          \
    } while (0)

#define STEIN_REQUIRE(ctx, cond, ...)                                     \
    do {                                                                  \
        if (!(cond)) return stein::fail((ctx), STEIN_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define STEIN_TRY(expr)              \
    do {                             \
        int _rc = (expr);            \
        if (_rc != STEIN_OK) return _rc; \
    } while (0)

// RAII region timer: records an event pair around a region when profiling is on
struct RegionTimer {
    stein_ctx *ctx;
    int region;
    std::pair<cudaEvent_t, cudaEvent_t> ev{};
    bool on;
    RegionTimer(stein_ctx *c, int r) : ctx(c), region(r), on(c->profile >= (r <= STEIN_REGION_SWEEP ? 1 : 2)) {
        if (!on) return;
        if (!ctx->prof_pool.empty()) {
            ev = ctx->prof_pool.back();
            ctx->prof_pool.pop_back();
        } else {
            cudaEventCreate(&ev.first);
            cudaEventCreate(&ev.second);
        }
        cudaEventRecord(ev.first, ctx->stream);
    }
    void stop() {
        if (!on) return;
        cudaEventRecord(ev.second, ctx->stream);
        ctx->prof_events[region].push_back(ev);
        on = false;
    }
    ~RegionTimer() { stop(); }
};

// labelled timeline mark: completes when everything enqueued before it on the ctx stream has finished
inline void trace_mark(stein_ctx *ctx, const char *label) {
    if (!ctx->trace) return;
    cudaEvent_t ev;
    if (!ctx->trace_pool.empty()) {
        ev = ctx->trace_pool.back();
        ctx->trace_pool.pop_back();
    } else {
        cudaEventCreate(&ev);
    }
    cudaEventRecord(ev, ctx->stream);
    ctx->trace_marks.emplace_back(label, ev);
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Where the optimizer kernel additionally stores the updated rows: the same row range in the
// particle buffers of the other ranks of the node (opened through CUDA IPC, engine.cu).
constexpr int MAX_PEERS = 15;
struct PeerTargets {
    float4 *dst[MAX_PEERS];
    int n = 0;
};
// optimizer.cu
int clip_adam_step(stein_ctx *ctx, float *X_dev, const float *phi_dev, float *mu_dev, float *nu_dev,
                   int64_t count, const double *sumsq_dev, double learning_rate, double beta_1, double beta_2,
                   int64_t n_iters, const PeerTargets &peers);
int clip_adagrad_step(stein_ctx *ctx, float *X_dev, const float *phi_dev, float *hist_dev, int64_t count,
                      const double *sumsq_dev, double learning_rate, double alpha, int64_t n_iters,
                      const PeerTargets &peers);

// order-preserving fp32 -> u32 (ascending); -0 folded onto +0
__host__ __device__ inline uint32_t float_to_key(float f) {
    f = f + 0.0f;
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float key_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// map linear upper-triangular tile index t (row-major over I<=J) to (I, J)
__host__ __device__ inline void tri_tile(int64_t t, int64_t T, int &I, int &J) {
    // row I starts at offset I*T - I*(I-1)/2
    double Td = (double)T;
    double disc = (2.0 * Td + 1.0) * (2.0 * Td + 1.0) - 8.0 * (double)t;
    int64_t i = (int64_t)(((2.0 * Td + 1.0) - sqrt(disc)) * 0.5);
    if (i < 0) i = 0;
    if (i >= T) i = T - 1;
    while (i > 0 && i * T - i * (i - 1) / 2 > t) --i;
    while ((i + 1) * T - (i + 1) * i / 2 <= t) ++i;
    I = (int)i;
    J = (int)(i + (t - (i * T - i * (i - 1) / 2)));
}

}  // namespace stein
