// ctx.cu -- context lifetime, error text, collective hooks, phi dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>

#include "phi_common.cuh"

namespace stein {

thread_local std::string g_last_error;

int fail(stein_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->error = buf;
    return code;
}

}  // namespace stein

using namespace stein;

extern "C" {

int stein_abi_version(void) { return STEIN_ABI_VERSION; }

int64_t stein_ld(int64_t d) { return round_up(d < 1 ? 1 : d, LD_ALIGN); }
int64_t stein_rows_padded(int64_t n) { return round_up(n < 1 ? 1 : n, TILE); }

int stein_ctx_create(stein_ctx **out, int device, void *cuda_stream) {
    if (!out) return fail(nullptr, STEIN_ERR_INVALID, "null output pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, STEIN_ERR_CUDA,
                    "no CUDA device available (%s): libstein_b200 has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= count)
        return fail(nullptr, STEIN_ERR_INVALID, "device %d out of range (%d devices)", device, count);
    // One process drives ONE GPU: the library keeps per-process device scratch (median arena, select
    // state, slot plan) and per-device function attributes.  A context on a second device of the
    // same process would use buffers that live on the first one -- refuse it loudly.
    static int process_device = -1;
    if (process_device >= 0 && process_device != device)
        return fail(nullptr, STEIN_ERR_UNSUPPORTED,
                    "this process already drives CUDA device %d: libstein_b200 supports one GPU per process "
                    "(launch one process per GPU, e.g. torchrun); device %d refused", process_device, device);
    process_device = device;
    stein_ctx *ctx = new stein_ctx();
    ctx->device = device;
    e = cudaSetDevice(device);
    cudaDeviceProp prop{};
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete ctx;
        return fail(nullptr, STEIN_ERR_CUDA, "cudaSetDevice/GetDeviceProperties: %s", cudaGetErrorString(e));
    }
    if (prop.major != 10) {
        const int major = prop.major, minor = prop.minor;
        delete ctx;
        return fail(nullptr, STEIN_ERR_UNSUPPORTED,
                    "device is sm_%d%d; libstein_b200 is built for sm_100a (B200) only", major, minor);
    }
    ctx->num_sms = prop.multiProcessorCount;
    ctx->stream = (cudaStream_t)cuda_stream;
    if (const char *tol = getenv("STEIN_PHI_GUARD_TOL")) {
        const float v = (float)atof(tol);
        if (v >= 0.0f) ctx->phi_guard_tol = v;
    }
    *out = ctx;
    return STEIN_OK;
}

int stein_ctx_destroy(stein_ctx *ctx) {
    if (!ctx) return STEIN_OK;
    cudaSetDevice(ctx->device);
    nccl_release(ctx);
    if (ctx->d_counts) cudaFree(ctx->d_counts);
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    if (ctx->d_pilot_keys) cudaFree(ctx->d_pilot_keys);
    if (ctx->d_sel) cudaFree(ctx->d_sel);
    if (ctx->h_sel) cudaFreeHost(ctx->h_sel);
    if (ctx->d_guard) cudaFree(ctx->d_guard);
    if (ctx->h_guard) cudaFreeHost(ctx->h_guard);
    for (int k = 0; k < 2; ++k)
        if (ctx->ev_guard[k]) cudaEventDestroy(ctx->ev_guard[k]);
    for (int r = 0; r < STEIN_REGION_COUNT; ++r)
        for (auto &ev : ctx->prof_events[r]) ctx->prof_pool.push_back(ev);
    for (auto &m : ctx->trace_marks) cudaEventDestroy(m.second);
    for (auto &ev : ctx->trace_pool) cudaEventDestroy(ev);
    for (auto &ev : ctx->prof_pool) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    delete ctx;
    return STEIN_OK;
}

int stein_ctx_set_stream(stein_ctx *ctx, void *cuda_stream) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    ctx->stream = (cudaStream_t)cuda_stream;
    return STEIN_OK;
}

int stein_ctx_set_comm(stein_ctx *ctx, const stein_comm *comm) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    if (!comm || comm->world <= 1) {
        ctx->has_comm = false;
        return STEIN_OK;
    }
    STEIN_REQUIRE(ctx, comm->rank >= 0 && comm->rank < comm->world, "rank %d outside world %d",
                  comm->rank, comm->world);
    STEIN_REQUIRE(ctx, comm->allreduce_sum_u64 && comm->allreduce_sum_f64 && comm->allgather_f32,
                  "all three collective hooks are required when world > 1");
    ctx->comm = *comm;
    ctx->has_comm = true;
    return STEIN_OK;
}

int stein_ctx_set_phi_impl(stein_ctx *ctx, int impl) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, impl >= STEIN_PHI_AUTO && impl <= STEIN_PHI_FLASH_TC5, "unknown phi impl %d", impl);
    ctx->phi_impl = impl;
    return STEIN_OK;
}

int stein_ctx_set_median_impl(stein_ctx *ctx, int impl) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, impl >= STEIN_MEDIAN_AUTO && impl <= STEIN_MEDIAN_TC1, "unknown median impl %d", impl);
    ctx->median_impl = impl;
    return STEIN_OK;
}

const char *stein_last_error(const stein_ctx *ctx) {
    return ctx ? ctx->error.c_str() : g_last_error.c_str();
}

int stein_ctx_profile_enable(stein_ctx *ctx, int enable) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    ctx->profile = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
    return STEIN_OK;
}

int stein_ctx_profile_read(stein_ctx *ctx, int region, double *ms_total, int64_t *launches) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, region >= 0 && region < STEIN_REGION_COUNT, "unknown region");
    STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double total = 0.0;
    for (auto &ev : ctx->prof_events[region]) {
        float ms = 0.f;
        STEIN_CHECK_CUDA(ctx, cudaEventElapsedTime(&ms, ev.first, ev.second));
        total += ms;
        ctx->prof_pool.push_back(ev);
    }
    if (ms_total) *ms_total = total;
    if (launches) *launches = (int64_t)ctx->prof_events[region].size();
    ctx->prof_events[region].clear();
    return STEIN_OK;
}

int stein_ctx_trace_enable(stein_ctx *ctx, int enable) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    ctx->trace = enable != 0;
    return STEIN_OK;
}

// "label<TAB>ms since the previous mark" per line, in enqueue order; clears the marks
int stein_ctx_trace_read(stein_ctx *ctx, char *buf, int64_t cap) {
    STEIN_REQUIRE(ctx, ctx != nullptr && buf != nullptr && cap > 0, "bad arguments");
    STEIN_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    std::string out;
    for (size_t k = 0; k < ctx->trace_marks.size(); ++k) {
        float ms = 0.f;
        if (k > 0) STEIN_CHECK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->trace_marks[k - 1].second, ctx->trace_marks[k].second));
        char line[160];
        snprintf(line, sizeof(line), "%s\t%.6f\n", ctx->trace_marks[k].first, (double)ms);
        out += line;
    }
    for (auto &m : ctx->trace_marks) ctx->trace_pool.push_back(m.second);
    ctx->trace_marks.clear();
    const size_t ncopy = std::min<size_t>(out.size(), (size_t)cap - 1);
    memcpy(buf, out.data(), ncopy);
    buf[ncopy] = 0;
    return STEIN_OK;
}

int64_t stein_ctx_launch_count(const stein_ctx *ctx) { return ctx ? ctx->launches : 0; }

int stein_ctx_phi_route(stein_ctx *ctx, int32_t *route, float *kappa, float *predicted_fast_error) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    if (route) *route = ctx->last_route;
    if (kappa) *kappa = ctx->last_kappa;
    if (predicted_fast_error) *predicted_fast_error = ctx->last_pred_fast;
    return STEIN_OK;
}

int stein_ctx_set_phi_guard_tol(stein_ctx *ctx, float tol) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, tol >= 0.0f, "tolerance must be >= 0");
    ctx->phi_guard_tol = tol;
    return STEIN_OK;
}

// internal: CTA-pair kernel, fast (FLASH_TC4) or precise (FLASH_TC5) route picked on the device by the
// conditioning guard -- what STEIN_PHI_AUTO resolves to for a leading dimension of 256
constexpr int PHI_FLASH_GUARDED = STEIN_PHI_FLASH_TC5 + 1;
static bool is_pair_impl(int impl) { return impl >= STEIN_PHI_FLASH_TC2 && impl <= PHI_FLASH_GUARDED; }
// leading dimensions that make a matrix eligible for the tensor-core kernels whatever d is: the zero pad
// columns are then simply treated as coordinates
static bool is_tc_ld(int64_t ld) { return ld == 128 || ld == 256 || ld == 512 || ld == 768 || ld == 1024; }

static int pick_phi_impl(const stein_ctx *ctx, int64_t n_local, int64_t n_total, int64_t d) {
    if (ctx->phi_impl != STEIN_PHI_AUTO) return ctx->phi_impl;
    if (panel::panel_supported(ctx, n_total, stein_ld(d))) return PHI_FLASH_GUARDED;      // panel kernels, guarded
    if (flash_tc2_supported(ctx, n_local, n_total, d)) return PHI_FLASH_GUARDED;
    return flash_tc_supported(ctx, n_local, n_total, d) ? STEIN_PHI_FLASH_TC : STEIN_PHI_DENSE_SIMT;
}

int64_t stein_phi_workspace_bytes(const stein_ctx *ctx, int64_t n_local, int64_t n_total, int64_t d) {
    if (!ctx) return -1;
    // sized so that either implementation can run (the impl may be switched later)
    int64_t b = dense_workspace_bytes(n_local, n_total, d);
    if (flash_tc_supported(ctx, n_local, n_total, d))
        b = std::max(b, flash_tc_workspace_bytes(ctx, n_local, n_total, d));
    if (panel::panel_supported(ctx, n_total, stein_ld(d)))
        b = std::max(b, panel::panel_workspace_bytes(ctx, n_local, n_total, stein_ld(d)));
    return b;
}

int stein_phi(stein_ctx *ctx, const float *X_all_dev, const float *S_all_dev, const float *r_all_dev,
              int64_t n_total, int64_t d, int64_t ld, int64_t row_begin, int64_t n_local,
              float bandwidth, void *workspace_dev, int64_t workspace_bytes, float *phi_dev,
              double *sumsq_dev) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, X_all_dev && S_all_dev && r_all_dev && workspace_dev && phi_dev && sumsq_dev,
                  "null pointer");
    STEIN_REQUIRE(ctx, n_total >= 1 && d >= 1 && ld >= d && ld % LD_ALIGN == 0, "bad shape");
    STEIN_REQUIRE(ctx, row_begin >= 0 && n_local >= 1 && row_begin % TILE == 0,
                  "row_begin=%lld must be a non-negative multiple of %d", (long long)row_begin, TILE);
    if (ctx->dev_bw) bandwidth = 1.0f;       // (internal: the kernels read the bandwidth from the device block)
    STEIN_REQUIRE(ctx, bandwidth > 0.0f && bandwidth == bandwidth, "bandwidth must be positive and finite");
    const float h2 = bandwidth * bandwidth;  // squared_exponential_kernel.py:22 tf.square(bandwidth)
    // A leading dimension of 128 / 256 makes the matrix eligible for the tensor-core kernels
    // whatever d is: the zero pad columns are then simply treated as coordinates.
    const int64_t d_true = d;
    if (is_tc_ld(ld) && d < ld) d = ld;
    const int impl = pick_phi_impl(ctx, n_local, n_total, d);
    if (ctx->dev_bw && !(is_pair_impl(impl) && !panel::panel_supported(ctx, n_total, ld) &&
                         flash_tc2_supported(ctx, n_local, n_total, d)))
        return PHI_DEV_BW_NA;                // only the CTA-pair flash kernels take the device block
    if (is_pair_impl(impl) && impl >= STEIN_PHI_FLASH_TC4 && panel::panel_supported(ctx, n_total, ld))
        return panel::phi_panel(ctx, X_all_dev, S_all_dev, r_all_dev, n_total, d, d_true, ld, row_begin, n_local, h2,
                                workspace_dev, workspace_bytes, phi_dev, sumsq_dev, impl - STEIN_PHI_FLASH_TC2);
    if (is_pair_impl(impl)) {
        if (!flash_tc2_supported(ctx, n_local, n_total, d))
            return fail(ctx, STEIN_ERR_UNSUPPORTED, "CTA-pair flash phi does not take n=%lld d=%lld",
                        (long long)n_total, (long long)d);
        return phi_flash_tc2(ctx, X_all_dev, S_all_dev, r_all_dev, n_total, d, d_true, ld, row_begin, n_local, h2,
                             workspace_dev, workspace_bytes, phi_dev, sumsq_dev, impl - STEIN_PHI_FLASH_TC2);
    }
    if (impl == STEIN_PHI_FLASH_TC) {
        if (!flash_tc_supported(ctx, n_local, n_total, d))
            return fail(ctx, STEIN_ERR_UNSUPPORTED, "flash tcgen05 phi does not take n=%lld d=%lld",
                        (long long)n_total, (long long)d);
        return phi_flash_tc(ctx, X_all_dev, S_all_dev, r_all_dev, n_total, d, ld, row_begin, n_local, h2,
                            workspace_dev, workspace_bytes, phi_dev, sumsq_dev, ctx->phi_impl == STEIN_PHI_AUTO, d_true);
    }
    return phi_dense(ctx, X_all_dev, S_all_dev, r_all_dev, n_total, d, ld, row_begin, n_local, h2,
                     workspace_dev, workspace_bytes, phi_dev, sumsq_dev);
}

}  // extern "C"

namespace stein {
// Same dispatch as stein_phi; a no-op for the kernels that have nothing to hoist.
int phi_prepare_x(stein_ctx *ctx, const float *X_all, int64_t n_total, int64_t d, int64_t ld, int64_t n_local,
                  void *ws, int64_t ws_bytes) {
    if (is_tc_ld(ld) && d < ld) d = ld;
    const int impl = pick_phi_impl(ctx, n_local, n_total, d);
    if (is_pair_impl(impl) && flash_tc2_supported(ctx, n_local, n_total, d))
        return flash_tc2_prepare_x(ctx, X_all, n_total, d, ld, n_local, ws, ws_bytes, impl - STEIN_PHI_FLASH_TC2);
    return STEIN_OK;
}
// ... and, once the scores are there, the bandwidth-independent column maxima behind the column scales of Y
int phi_prepare_s(stein_ctx *ctx, const float *X_all, const float *S_all, int64_t n_total, int64_t d, int64_t ld,
                  int64_t n_local, void *ws, int64_t ws_bytes) {
    if (is_tc_ld(ld) && d < ld) d = ld;
    const int impl = pick_phi_impl(ctx, n_local, n_total, d);
    if (is_pair_impl(impl) && flash_tc2_supported(ctx, n_local, n_total, d))
        return flash_tc2_prepare_s(ctx, X_all, S_all, n_total, d, ld, n_local, ws, ws_bytes, impl - STEIN_PHI_FLASH_TC2);
    return STEIN_OK;
}
}  // namespace stein
