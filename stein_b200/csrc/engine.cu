// engine.cu -- device-resident particles + optimizer state, one SVGD
// update_particles() per stein_engine_step().
//
// Reference: AbstractSteinSampler.update_particles
//            stein/samplers/abstract_stein_sampler.py:107-127
//   theta_array  <- dict->array                      (:121, converters.py:4-55)
//   phi          <- compute_phi(theta_array, grads)  (:123 -> :100-105)
//   phi          *= 10 / max(10, ||phi||_F)          (:125)
//   theta_array  += gd.update(phi)                   (:126)
// The reference keeps theta as float64 NumPy on the host and round-trips through
// TensorFlow every iteration; here the particles, scores, phi and the optimizer
// moments live in HBM (fp32, padded layout) and only cross PCIe when the caller
// asks (set_/get_ functions, update_particles_host).
//
// Sharding (ctx collective hooks set): rank g owns global rows
// [g*q, g*q + q) with q = rows_padded(ceil(n/world)); X_all/S_all hold all
// world*q rows and are refreshed by an in-place all-gather each step.
#include <algorithm>

#include "phi_common.cuh"
#include "peer_reduce.cuh"

struct stein_engine {
    stein_ctx *ctx = nullptr;
    int64_t n_total = 0, d = 0, ld = 0;
    int world = 1, rank = 0;
    int64_t q = 0;          // padded rows per rank
    int64_t row_begin = 0;  // first global row owned
    int64_t n_local = 0;    // valid rows owned
    int opt = STEIN_OPT_ADAM;
    double lr = 1e-3, decay = 1.0, p1 = 0.9, p2 = 0.999;
    int64_t n_iters = 0;
    float *X_all = nullptr, *S_all = nullptr, *phi = nullptr, *m1 = nullptr, *m2 = nullptr;
    float *r_all = nullptr;
    double *sumsq = nullptr;
    double *h_sumsq = nullptr;
    void *ws = nullptr;
    int64_t ws_bytes = 0;
    void *stage = nullptr;  // device staging for float64 host arrays
    cudaStream_t copy_stream = nullptr;  // score upload overlapped with the median (update_particles_host)
    cudaEvent_t ev_scores = nullptr, ev_ready = nullptr;
    // The part of the phi preparation that needs neither the bandwidth nor the scores (centring, scale of X, the
    // X operand arrays) runs on this stream NEXT TO the median: HBM-bound kernels beside the tensor-bound sweep.
    cudaStream_t prep_stream = nullptr;
    cudaEvent_t ev_x = nullptr, ev_prep = nullptr;
    bool prep_pending = false;
    // update_particles_host: the next iteration's head and median are enqueued behind the optimizer kernel, so they
    // run while the updated particles cross PCIe (the median needs only the particles, not the caller's next scores)
    bool prefetch = true;
    bool bw_pending = false;        // a deferred median (median_sqdist_begin) of the current particles is in flight
    int64_t prefetch_begun = 0, prefetch_used = 0;
    // step / update_particles_host: phi is enqueued BEFORE the host has collected the median, its kernels take the
    // bandwidth from the device-side select of the deferred median (median_sqdist_device_bandwidth); the host
    // verifies afterwards, behind the phi kernel instead of in front of it
    bool device_bw = true;
    int64_t devbw_used = 0, devbw_redone = 0;
    cudaEvent_t ev_update = nullptr;    // after the optimizer kernel: the download and the next score upload wait for it
    // peer push of the updated particles (CUDA IPC views of the other ranks' X_all)
    float *peer_X[stein::MAX_PEERS + 1] = {nullptr};   // indexed by rank; own entry unused
    bool peers_open = false;
    bool x_all_current = false;     // every rank's X_all holds everybody's current rows
    unsigned long long *barrier_word = nullptr;   // 1 u64 all-reduced as the cross-rank barrier after a push
    // sharded engines: the particle buffer is followed by a mailbox (peer_reduce.cuh) so that one
    // IPC handle exports both; peer_reduce is what the context uses while the peers are open
    int64_t mbox_offset = 0;
    stein::PeerReduce *peer_reduce = nullptr;
    unsigned long long mbox_epoch = 0;   // all-reduces issued through the mailbox so far: survives a close / re-open
                                         // of the peers (the flags in the mailboxes keep the old epochs)
    uintptr_t uid = 0;      // unique per engine ever created in the process (median window hint owner)
    float last_med = 0.f, last_bw = 0.f;
    float fixed_bw = 0.f;   // > 0: use this bandwidth instead of the median heuristic
    int32_t last_sweeps = 0;
    float *X_local() const { return X_all + (int64_t)rank * q * ld; }
    float *S_local() const { return S_all + (int64_t)rank * q * ld; }
};

namespace stein {

// padded fp32 [rows x ld]  <->  dense float64 [rows x d]
__global__ void f64_to_padded_kernel(const double *__restrict__ src, int64_t rows, int64_t d, int64_t ld,
                                     float *__restrict__ dst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * d) return;
    const int64_t r = e / d, c = e % d;
    dst[r * ld + c] = (float)src[e];
}
__global__ void padded_to_f64_kernel(const float *__restrict__ src, int64_t rows, int64_t d, int64_t ld,
                                     double *__restrict__ dst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * d) return;
    const int64_t r = e / d, c = e % d;
    dst[e] = (double)src[r * ld + c];
}

static int upload(stein_engine *e, const void *host, int is_f64, float *dst, cudaStream_t stream) {
    stein_ctx *ctx = e->ctx;
    STEIN_REQUIRE(ctx, host != nullptr, "null host pointer");
    if (e->n_local == 0) return STEIN_OK;
    if (!is_f64) {
        if (e->ld == e->d) {
            STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(dst, host, e->n_local * e->d * 4, cudaMemcpyHostToDevice, stream));
        } else {
            STEIN_CHECK_CUDA(ctx, cudaMemcpy2DAsync(dst, e->ld * 4, host, e->d * 4, e->d * 4, e->n_local,
                                                    cudaMemcpyHostToDevice, stream));
        }
    } else {
        STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(e->stage, host, e->n_local * e->d * 8,
                                              cudaMemcpyHostToDevice, stream));
        const int64_t total = e->n_local * e->d;
        f64_to_padded_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
            (const double *)e->stage, e->n_local, e->d, e->ld, dst);
        STEIN_CHECK_LAUNCH(ctx);
    }
    return STEIN_OK;
}
static int upload(stein_engine *e, const void *host, int is_f64, float *dst) {
    return upload(e, host, is_f64, dst, e->ctx->stream);
}

static int download_async(stein_engine *e, const float *src, void *host, int is_f64, cudaStream_t stream) {
    stein_ctx *ctx = e->ctx;
    STEIN_REQUIRE(ctx, host != nullptr, "null host pointer");
    if (e->n_local == 0) return STEIN_OK;
    if (!is_f64) {
        if (e->ld == e->d) {
            STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(host, src, e->n_local * e->d * 4, cudaMemcpyDeviceToHost, stream));
        } else {
            STEIN_CHECK_CUDA(ctx, cudaMemcpy2DAsync(host, e->d * 4, src, e->ld * 4, e->d * 4, e->n_local,
                                                    cudaMemcpyDeviceToHost, stream));
        }
    } else {
        const int64_t total = e->n_local * e->d;
        padded_to_f64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
            src, e->n_local, e->d, e->ld, (double *)e->stage);
        STEIN_CHECK_LAUNCH(ctx);
        STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(host, e->stage, total * 8, cudaMemcpyDeviceToHost, stream));
    }
    return STEIN_OK;
}

static int download(stein_engine *e, const float *src, void *host, int is_f64) {
    stein_ctx *ctx = e->ctx;
    STEIN_TRY(download_async(e, src, host, is_f64, ctx->stream));
    STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return STEIN_OK;
}

}  // namespace stein

using namespace stein;

extern "C" {

int stein_engine_create(stein_engine **out, stein_ctx *ctx, int64_t n_total, int64_t d, int optimizer,
                        double learning_rate, double decay, double p1, double p2) {
    STEIN_REQUIRE(ctx, ctx != nullptr && out != nullptr, "null ctx/out");
    *out = nullptr;
    STEIN_REQUIRE(ctx, n_total >= 2, "n_particles=%lld: the bandwidth needs ln(n) > 0", (long long)n_total);
    STEIN_REQUIRE(ctx, d >= 1, "d must be positive");
    STEIN_REQUIRE(ctx, optimizer == STEIN_OPT_ADAM || optimizer == STEIN_OPT_ADAGRAD, "unknown optimizer");
    STEIN_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    static uintptr_t next_uid = 1;
    stein_engine *e = new stein_engine();
    e->uid = next_uid++;
    e->ctx = ctx;
    e->n_total = n_total;
    e->d = d;
    // Many particles of up to 256 coordinates: pad the rows to 128 / 256 floats so that the
    // tensor-core median and phi kernels apply (pad columns are zero and stay zero).
    e->ld = stein_ld(d);
    if (n_total >= 2048 && d <= 256) e->ld = d <= 128 ? 128 : 256;
    else if (n_total >= 2048 && d <= 1024) e->ld = stein::round_up(d, 256);     // 512 / 768 / 1024: the panel kernels
    e->world = ctx->has_comm ? ctx->comm.world : 1;
    e->rank = ctx->has_comm ? ctx->comm.rank : 0;
    e->q = stein_rows_padded((n_total + e->world - 1) / e->world);
    e->row_begin = (int64_t)e->rank * e->q;
    e->n_local = std::max<int64_t>(0, std::min<int64_t>(e->q, n_total - e->row_begin));
    e->opt = optimizer;
    e->lr = learning_rate;
    e->decay = decay;
    e->p1 = p1;
    e->p2 = p2;
    const int64_t all = e->q * e->world * e->ld * 4, loc = e->q * e->ld * 4;
    e->ws_bytes = stein_phi_workspace_bytes(ctx, std::max<int64_t>(e->n_local, 1), n_total, e->ld);
    cudaError_t err = cudaSuccess;
    auto alloc0 = [&](void **p, int64_t bytes) {
        if (err != cudaSuccess) return;
        err = cudaMalloc(p, bytes);
        if (err == cudaSuccess) err = cudaMemsetAsync(*p, 0, bytes, ctx->stream);
    };
    e->mbox_offset = all;
    alloc0((void **)&e->X_all, all + (e->world > 1 ? (int64_t)MBOX_BYTES : 0));
    alloc0((void **)&e->S_all, all);
    alloc0((void **)&e->phi, loc);
    alloc0((void **)&e->m1, loc);
    alloc0((void **)&e->m2, loc);
    alloc0((void **)&e->r_all, e->q * e->world * 4);
    alloc0((void **)&e->sumsq, 64);
    alloc0(&e->ws, e->ws_bytes);
    alloc0(&e->stage, std::max<int64_t>(e->q * d * 8, 256));
    if (err == cudaSuccess) err = cudaMallocHost(&e->h_sumsq, 64);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_scores, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_ready, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->prep_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_x, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_prep, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_update, cudaEventDisableTiming);
    if (const char *env = getenv("STEIN_PREFETCH")) e->prefetch = env[0] != '0';
    // (on one GPU the host round trip it hides costs 0.017 ms, the device-side select and its copy 0.02-0.05 ms:
    //  default on for sharded engines only, where the round trip's jitter is maximised over the ranks)
    e->device_bw = ctx->has_comm && ctx->comm.world > 1;
    if (const char *env = getenv("STEIN_DEVICE_BW")) e->device_bw = env[0] != '0';
    if (err != cudaSuccess) {
        stein_engine_destroy(e);
        return fail(ctx, err == cudaErrorMemoryAllocation ? STEIN_ERR_NOMEM : STEIN_ERR_CUDA,
                    "engine allocation failed: %s", cudaGetErrorString(err));
    }
    *out = e;
    return STEIN_OK;
}

static void release_peer_reduce(stein_engine *e) {
    if (!e->peer_reduce) return;
    if (e->ctx->peer_reduce == e->peer_reduce) e->ctx->peer_reduce = nullptr;
    e->mbox_epoch = peer_reduce_epoch(e->peer_reduce);
    peer_reduce_destroy(e->peer_reduce);
    e->peer_reduce = nullptr;
}

int stein_engine_destroy(stein_engine *e) {
    if (!e) return STEIN_OK;
    cudaSetDevice(e->ctx->device);
    cudaStreamSynchronize(e->ctx->stream);
    median_tc_forget_owner(reinterpret_cast<const void *>(e->uid));
    release_peer_reduce(e);
    if (e->peers_open)
        for (int r = 0; r < e->world; ++r)
            if (r != e->rank && e->peer_X[r]) cudaIpcCloseMemHandle(e->peer_X[r]);
    if (e->barrier_word) cudaFree(e->barrier_word);
    void *ptrs[] = {e->X_all, e->S_all, e->phi, e->m1, e->m2, e->r_all, e->sumsq, e->ws, e->stage};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (e->h_sumsq) cudaFreeHost(e->h_sumsq);
    if (e->copy_stream) {
        cudaStreamSynchronize(e->copy_stream);
        cudaStreamDestroy(e->copy_stream);
    }
    if (e->prep_stream) {
        cudaStreamSynchronize(e->prep_stream);
        cudaStreamDestroy(e->prep_stream);
    }
    if (e->ev_x) cudaEventDestroy(e->ev_x);
    if (e->ev_prep) cudaEventDestroy(e->ev_prep);
    if (e->ev_update) cudaEventDestroy(e->ev_update);
    if (e->ev_scores) cudaEventDestroy(e->ev_scores);
    if (e->ev_ready) cudaEventDestroy(e->ev_ready);
    delete e;
    return STEIN_OK;
}

int stein_engine_local_rows(const stein_engine *e, int64_t *row_begin, int64_t *n_local) {
    if (!e) return STEIN_ERR_INVALID;
    if (row_begin) *row_begin = e->row_begin;
    if (n_local) *n_local = e->n_local;
    return STEIN_OK;
}

int stein_engine_buffers(stein_engine *e, float **X_local_dev, float **S_local_dev,
                         float **phi_local_dev, int64_t *ld, int64_t *rows_padded_local) {
    if (!e) return STEIN_ERR_INVALID;
    if (X_local_dev) *X_local_dev = e->X_local();
    if (S_local_dev) *S_local_dev = e->S_local();
    if (phi_local_dev) *phi_local_dev = e->phi;
    if (ld) *ld = e->ld;
    if (rows_padded_local) *rows_padded_local = e->q;
    return STEIN_OK;
}

int stein_engine_set_particles(stein_engine *e, const void *X_host, int is_f64) {
    if (!e) return STEIN_ERR_INVALID;
    e->x_all_current = false;      // the other ranks' copies of these rows are stale now
    e->bw_pending = false;         // a prefetched median belongs to the old particles
    if (e->ctx->guard_lag_owner == reinterpret_cast<const void *>(e->uid))
        e->ctx->guard_lag_owner = nullptr;     // a new cloud: the next phi call waits for its own conditioning number
    STEIN_TRY(upload(e, X_host, is_f64, e->X_local()));
    STEIN_CHECK_CUDA(e->ctx, cudaStreamSynchronize(e->ctx->stream));
    return STEIN_OK;
}

int stein_engine_get_particles(stein_engine *e, void *X_host, int is_f64) {
    if (!e) return STEIN_ERR_INVALID;
    return download(e, e->X_local(), X_host, is_f64);
}

int stein_engine_set_scores(stein_engine *e, const void *S_host, int is_f64) {
    if (!e) return STEIN_ERR_INVALID;
    return upload(e, S_host, is_f64, e->S_local());
}

int stein_engine_get_phi(stein_engine *e, void *phi_host, int is_f64) {
    if (!e) return STEIN_ERR_INVALID;
    return download(e, e->phi, phi_host, is_f64);
}

// The part of the phi preparation that needs neither the bandwidth nor the scores (centring, scale of X, the X
// operand arrays in the fast format); step_bandwidth enqueues it on the engine's prep stream beside the median.
static int engine_presync(void *arg) {
    stein_engine *e = static_cast<stein_engine *>(arg);
    return phi_prepare_x(e->ctx, e->X_all, e->n_total, e->d, e->ld, std::max<int64_t>(e->n_local, 1), e->ws,
                         e->ws_bytes);
}

// Start of an iteration, needs only the particles: all-gather X (or the barrier behind the peer push), row norms,
// and the bandwidth-independent part of the phi preparation on its own stream.
static int step_head(stein_engine *e) {
    stein_ctx *ctx = e->ctx;
    const int64_t rows_all = e->q * e->world;
    ctx->xprep.X = nullptr;
    ctx->sprep_S = nullptr;
    RegionTimer head(ctx, STEIN_REGION_HEAD);
    trace_mark(ctx, "step:begin");
    if (e->world > 1) {
        if (e->peers_open && e->x_all_current) {
            // every rank pushed its updated rows into this buffer during its last optimizer
            // step; a 1-word all-reduce is the barrier that orders those stores (it completes
            // only after every rank has enqueued it, i.e. after its step kernel)
            STEIN_TRY(allreduce_u64(ctx, e->barrier_word, 1));
        } else {
            const int64_t cnt = e->q * e->ld;
            if (ctx->comm.allgather_f32(ctx->comm.user, e->X_local(), e->X_all, cnt) != 0)
                return fail(ctx, STEIN_ERR_COMM, "allgather_f32 hook failed");
        }
    }
    // abstract_kernel.py:34 -- r = sum(T*T, 1), contract order; when the tensor-core median will run on these
    // particles, its first stage (error budgets, scale, FP16 split) comes out of the same read
    if (e->fixed_bw <= 0.0f && median_sqdist_wants_fused_begin(ctx, e->n_total, e->ld))
        STEIN_TRY(median_sqdist_begin_with_norms(ctx, e->X_all, e->r_all, rows_all, e->n_total, e->ld));
    else
        STEIN_TRY(stein_row_norms(ctx, e->X_all, rows_all, e->d, e->ld, e->r_all));
    head.stop();
    trace_mark(ctx, "head:row norms");
    if (e->fixed_bw > 0.0f) return STEIN_OK;
    // the bandwidth-independent part of the phi preparation, on its own stream beside the median
    STEIN_CHECK_CUDA(ctx, cudaEventRecord(e->ev_x, ctx->stream));
    STEIN_CHECK_CUDA(ctx, cudaStreamWaitEvent(e->prep_stream, e->ev_x, 0));
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = e->prep_stream;
    const int prc = engine_presync(e);
    ctx->stream = main_stream;
    if (prc != STEIN_OK) return prc;
    STEIN_CHECK_CUDA(ctx, cudaEventRecord(e->ev_prep, e->prep_stream));
    e->prep_pending = true;
    return STEIN_OK;
}

// The part of the phi preparation that needs the scores but not the bandwidth (column maxima of S and of the
// centred X), on the prep stream behind the X-side preparation, beside the median.  scores_ready: the event behind
// which S_all is complete (NULL: the scores were written on the ctx stream).
static int step_prepare_s(stein_engine *e, cudaEvent_t scores_ready) {
    stein_ctx *ctx = e->ctx;
    if (!e->prep_pending) return STEIN_OK;
    if (scores_ready) {
        STEIN_CHECK_CUDA(ctx, cudaStreamWaitEvent(e->prep_stream, scores_ready, 0));
    } else {
        STEIN_CHECK_CUDA(ctx, cudaEventRecord(e->ev_x, ctx->stream));
        STEIN_CHECK_CUDA(ctx, cudaStreamWaitEvent(e->prep_stream, e->ev_x, 0));
    }
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = e->prep_stream;
    const int prc = phi_prepare_s(ctx, e->X_all, e->S_all, e->n_total, e->d, e->ld, std::max<int64_t>(e->n_local, 1),
                                  e->ws, e->ws_bytes);
    ctx->stream = main_stream;
    if (prc != STEIN_OK) return prc;
    STEIN_CHECK_CUDA(ctx, cudaEventRecord(e->ev_prep, e->prep_stream));
    return STEIN_OK;
}

// Phase 1: head, exact median, bandwidth.  When the previous update_particles_host call has already enqueued the
// head and the device part of the median for these particles (step_prefetch), only its result is collected.
// scores_in_place: S_all is complete behind `scores_ready` (or, NULL, in ctx stream order).
static int step_bandwidth(stein_engine *e, float *bw_out, bool scores_in_place, cudaEvent_t scores_ready) {
    stein_ctx *ctx = e->ctx;
    const bool resume = e->bw_pending && median_sqdist_deferred_pending(reinterpret_cast<const void *>(e->uid));
    e->bw_pending = false;
    if (resume) e->prefetch_used += 1;
    if (!resume) STEIN_TRY(step_head(e));
    if (scores_in_place) STEIN_TRY(step_prepare_s(e, scores_ready));
    if (e->fixed_bw > 0.0f) {
        e->last_med = nanf("");
        e->last_bw = e->fixed_bw;
        e->last_sweeps = 0;
        *bw_out = e->fixed_bw;
        return STEIN_OK;
    }
    // compute_median.py + abstract_kernel.py:40
    float med = 0.f;
    // successive medians of one engine move slowly: window hint allowed (owner = the engine's uid)
    ctx->median_owner = reinterpret_cast<const void *>(e->uid);
    RegionTimer mtimer(ctx, STEIN_REGION_MEDIAN);
    e->last_sweeps = 0;
    const int mrc = resume ? median_sqdist_resume(ctx, e->X_all, e->r_all, e->n_total, e->d, e->ld, &med, &e->last_sweeps)
                           : stein_median_sqdist(ctx, e->X_all, e->r_all, e->n_total, e->d, e->ld, &med, nullptr,
                                                 &e->last_sweeps);
    mtimer.stop();
    ctx->median_owner = nullptr;
    if (mrc != STEIN_OK) return mrc;
    const float bw = stein_bandwidth(med, e->n_total);
    e->last_med = med;
    e->last_bw = bw;
    if (!(bw > 0.0f) || bw != bw)
        return fail(ctx, STEIN_ERR_INVALID,
                    "median squared distance is %g: bandwidth undefined (all particles equal?)", (double)med);
    *bw_out = bw;
    return STEIN_OK;
}

// The next iteration's head and (in the pilot-less steady state of the median) the device part of its median,
// enqueued right behind the optimizer kernel.  No host wait: step_bandwidth of the next call collects the result.
static int step_prefetch(stein_engine *e) {
    stein_ctx *ctx = e->ctx;
    if (!e->prefetch || e->fixed_bw > 0.0f || e->bw_pending) return STEIN_OK;
    if (e->world > 1 && !(e->peers_open && e->x_all_current)) return STEIN_OK;   // the all-gather hook may block
    ctx->median_owner = reinterpret_cast<const void *>(e->uid);
    // (asked before the head is enqueued: a head without a deferred median would only be repeated)
    const bool can = median_sqdist_can_defer(ctx, e->n_total, e->ld);
    int rc = STEIN_OK;
    if (can) {
        rc = step_head(e);
        if (rc == STEIN_OK) rc = median_sqdist_begin(ctx, e->X_all, e->r_all, e->n_total, e->d, e->ld);
    }
    ctx->median_owner = nullptr;
    if (rc < 0) return rc;
    e->bw_pending = can && rc == MEDIAN_DEFERRED;
    if (e->bw_pending) e->prefetch_begun += 1;
    return STEIN_OK;
}

// Phase 2a needs the scores: (all-gather S,) phi for the local rows and the global sum(phi^2).
static int step_phi(stein_engine *e, float bw, bool scores_gathered) {
    stein_ctx *ctx = e->ctx;
    if (e->world > 1 && !scores_gathered) {
        const int64_t cnt = e->q * e->ld;
        if (ctx->comm.allgather_f32(ctx->comm.user, e->S_local(), e->S_all, cnt) != 0)
            return fail(ctx, STEIN_ERR_COMM, "allgather_f32 hook failed");
    }
    if (e->prep_pending) {      // the side-stream preparation of this iteration's X
        STEIN_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, e->ev_prep, 0));
        e->prep_pending = false;
    }
    // abstract_stein_sampler.py:100-105.  The conditioning guard of the phi call may decide from the previous
    // iteration's kappa (guard_owner): successive clouds of one engine differ by one small step.
    ctx->guard_owner = reinterpret_cast<const void *>(e->uid);
    const int prc = stein_phi(ctx, e->X_all, e->S_all, e->r_all, e->n_total, e->d, e->ld, e->row_begin,
                              std::max<int64_t>(e->n_local, 1), bw, e->ws, e->ws_bytes, e->phi, e->sumsq);
    ctx->guard_owner = nullptr;
    if (prc == PHI_DEV_BW_NA) return prc;
    STEIN_TRY(prc);
    if (e->world > 1 && !ctx->dev_bw) STEIN_TRY(allreduce_f64(ctx, e->sumsq, 1));
    return STEIN_OK;
}

static int step_bandwidth(stein_engine *e, float *bw_out, bool scores_in_place, cudaEvent_t scores_ready);
static int step_optimizer(stein_engine *e);

// One iteration with phi AHEAD of the host: the head and the device part of the median are enqueued (or already
// are: step_prefetch), phi follows on the stream reading the bandwidth the device-side select left in device
// memory, and only then the host collects the median.  It finds the same keys (same logic, same inputs) -- if not,
// or if the median had to take another route, phi is simply run again with the host's value.  The optimizer is
// launched after that check, so nothing irreversible ever uses an unverified bandwidth.  The host round trip of
// the median thus sits behind the phi kernel instead of in front of it (on 8 GPUs its wake-up jitter, maximised
// over the ranks by the next all-reduce, cost ~0.15 ms of a 1.65 ms step).
// Returns STEIN_OK with *done = false when this form does not apply (the caller takes the plain sequence).
static int step_device_bandwidth(stein_engine *e, bool scores_in_place, cudaEvent_t scores_ready, bool *done) {
    stein_ctx *ctx = e->ctx;
    *done = false;
    if (!e->device_bw || e->fixed_bw > 0.0f || !scores_in_place) return STEIN_OK;
    const void *owner = reinterpret_cast<const void *>(e->uid);
    bool pending = e->bw_pending && median_sqdist_deferred_pending(owner);
    const bool was_prefetched = pending;
    bool prepared_s = false;
    if (!pending) {
        e->bw_pending = false;
        ctx->median_owner = owner;
        const bool can = median_sqdist_can_defer(ctx, e->n_total, e->ld);
        ctx->median_owner = nullptr;
        if (!can) return STEIN_OK;
        STEIN_TRY(step_head(e));
        STEIN_TRY(step_prepare_s(e, scores_ready));      // (ahead of the median's kernels, as in the plain sequence)
        prepared_s = true;
        ctx->median_owner = owner;
        RegionTimer mtimer(ctx, STEIN_REGION_MEDIAN);
        const int rc = median_sqdist_begin(ctx, e->X_all, e->r_all, e->n_total, e->d, e->ld);
        mtimer.stop();
        ctx->median_owner = nullptr;
        if (rc < 0) return rc;
        if (rc != MEDIAN_DEFERRED) return fail(ctx, STEIN_ERR_INTERNAL, "median not deferred");
        e->bw_pending = true;
    }
    if (!prepared_s) STEIN_TRY(step_prepare_s(e, scores_ready));
    if (scores_ready) STEIN_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, scores_ready, 0));
    ctx->dev_bw = median_sqdist_device_bandwidth();
    const int prc = ctx->dev_bw ? step_phi(e, 1.0f, true) : PHI_DEV_BW_NA;
    ctx->dev_bw = nullptr;
    if (prc != STEIN_OK && prc != PHI_DEV_BW_NA) return prc;
    // the host's median (waits for the copies behind the median's tail -- long done unless phi was not enqueued)
    float bw = 0.f;
    STEIN_TRY(step_bandwidth(e, &bw, false, nullptr));
    if (!was_prefetched) e->prefetch_used -= 1;      // (collected a median this very call enqueued)
    float dev_h = 0.f;
    const bool valid = prc == STEIN_OK && median_sqdist_device_select_valid(&dev_h) && memcmp(&dev_h, &bw, 4) == 0;
    if (valid) {
        e->devbw_used += 1;
        if (e->world > 1) STEIN_TRY(allreduce_f64(ctx, e->sumsq, 1));
    } else {
        if (prc == STEIN_OK) e->devbw_redone += 1;
        STEIN_TRY(step_phi(e, bw, true));
    }
    STEIN_TRY(step_optimizer(e));
    *done = true;
    return STEIN_OK;
}

// Phase 2: phi, clip, optimizer step.
static int step_update(stein_engine *e, float bw, bool scores_gathered) {
    STEIN_TRY(step_phi(e, bw, scores_gathered));
    return step_optimizer(e);
}

// Phase 2b: clip + optimizer on the phi / sum(phi^2) in the engine's buffers.
static int step_optimizer(stein_engine *e) {
    stein_ctx *ctx = e->ctx;
    // abstract_stein_sampler.py:125-126
    const int64_t count = e->q * e->ld;
    // peers: the same rows inside the other ranks' X_all.  Safe to overwrite now: the all-reduce
    // of sum(phi^2) above means every rank is past its phi kernel, and nothing after it reads
    // the uncentred X_all of other ranks' rows.
    PeerTargets peers{};
    if (e->world > 1 && e->peers_open) {
        for (int r = 0; r < e->world; ++r)
            if (r != e->rank) peers.dst[peers.n++] = reinterpret_cast<float4 *>(e->peer_X[r] + (int64_t)e->rank * e->q * e->ld);
    }
    e->bw_pending = false;      // (a median prefetched for the particles this step replaces would be stale)
    RegionTimer otimer(ctx, STEIN_REGION_OPT);
    if (e->opt == STEIN_OPT_ADAM) {
        STEIN_TRY(clip_adam_step(ctx, e->X_local(), e->phi, e->m1, e->m2, count, e->sumsq, e->lr, e->p1, e->p2,
                                 e->n_iters, peers));
        e->lr *= e->decay;  // adam_gradient_descent.py:56
    } else {
        // adagrad_gradient_descent.py never applies `decay`
        STEIN_TRY(clip_adagrad_step(ctx, e->X_local(), e->phi, e->m1, count, e->sumsq, e->lr, e->p1, e->n_iters,
                                    peers));
    }
    trace_mark(ctx, "step:optimizer + push");
    if (peers.n) e->x_all_current = true;
    e->n_iters += 1;
    return STEIN_OK;
}

// Sharded runs with a side-stream all-gather hook: the score shards travel on the copy stream
// (own communicator) while the median runs; `after` = event the gather must wait for.
static int gather_scores_async(stein_engine *e, bool *done) {
    stein_ctx *ctx = e->ctx;
    *done = false;
    if (e->world <= 1 || !ctx->comm.allgather_f32_on) return STEIN_OK;
    const int64_t cnt = e->q * e->ld;
    if (ctx->comm.allgather_f32_on(ctx->comm.user, e->S_local(), e->S_all, cnt, (void *)e->copy_stream) != 0)
        return fail(ctx, STEIN_ERR_COMM, "allgather_f32_on hook failed");
    *done = true;
    return STEIN_OK;
}

int stein_engine_step(stein_engine *e) {
    if (!e) return STEIN_ERR_INVALID;
    stein_ctx *ctx = e->ctx;
    STEIN_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    // the scores were written on the ctx stream; their all-gather may overlap the median
    bool gathered = false;
    if (e->world > 1 && ctx->comm.allgather_f32_on) {
        STEIN_CHECK_CUDA(ctx, cudaEventRecord(e->ev_ready, ctx->stream));
        STEIN_CHECK_CUDA(ctx, cudaStreamWaitEvent(e->copy_stream, e->ev_ready, 0));
        STEIN_TRY(gather_scores_async(e, &gathered));
        STEIN_CHECK_CUDA(ctx, cudaEventRecord(e->ev_scores, e->copy_stream));
    }
    bool done = false;
    STEIN_TRY(step_device_bandwidth(e, gathered || e->world == 1, gathered ? e->ev_scores : nullptr, &done));
    if (done) return STEIN_OK;
    float bw = 0.f;
    const int rc = step_bandwidth(e, &bw, gathered || e->world == 1, gathered ? e->ev_scores : nullptr);
    if (gathered) STEIN_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, e->ev_scores, 0));
    if (rc != STEIN_OK) return rc;
    return step_update(e, bw, gathered);
}

int stein_engine_update_particles_host(stein_engine *e, const void *S_host, void *X_host_out,
                                       int is_f64) {
    if (!e) return STEIN_ERR_INVALID;
    stein_ctx *ctx = e->ctx;
    STEIN_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    // The median needs the particles only: the scores travel on a second stream while it runs.  The S buffer's
    // last reader is the previous phi: ev_update (recorded behind the previous optimizer kernel) orders the copy
    // after it when a prefetched median is already queued on the ctx stream, else an event recorded now.
    if (!e->bw_pending) STEIN_CHECK_CUDA(ctx, cudaEventRecord(e->ev_update, ctx->stream));
    STEIN_CHECK_CUDA(ctx, cudaStreamWaitEvent(e->copy_stream, e->ev_update, 0));
    STEIN_TRY(upload(e, S_host, is_f64, e->S_local(), e->copy_stream));
    bool gathered = false;
    STEIN_TRY(gather_scores_async(e, &gathered));
    STEIN_CHECK_CUDA(ctx, cudaEventRecord(e->ev_scores, e->copy_stream));
    bool done = false;
    STEIN_TRY(step_device_bandwidth(e, gathered || e->world == 1, e->ev_scores, &done));
    if (!done) {
        float bw = 0.f;
        const int rc = step_bandwidth(e, &bw, gathered || e->world == 1, e->ev_scores);
        STEIN_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, e->ev_scores, 0));
        if (rc != STEIN_OK) return rc;
        STEIN_TRY(step_update(e, bw, gathered));
    }
    if (!X_host_out) return STEIN_OK;
    // The updated particles cross PCIe on the copy stream while the ctx stream already runs the next iteration's
    // head and median (they need only the particles): the caller holds the new particles when this returns, and
    // the next call finds its bandwidth (nearly) ready.
    STEIN_CHECK_CUDA(ctx, cudaEventRecord(e->ev_update, ctx->stream));
    STEIN_CHECK_CUDA(ctx, cudaStreamWaitEvent(e->copy_stream, e->ev_update, 0));
    STEIN_TRY(download_async(e, e->X_local(), X_host_out, is_f64, e->copy_stream));
    const int prc = step_prefetch(e);
    STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(e->copy_stream));
    return prc;
}

int stein_engine_set_prefetch(stein_engine *e, int on) {
    if (!e) return STEIN_ERR_INVALID;
    e->prefetch = on != 0;
    return STEIN_OK;
}

int stein_engine_set_device_bandwidth(stein_engine *e, int on) {
    if (!e) return STEIN_ERR_INVALID;
    e->device_bw = on != 0;
    return STEIN_OK;
}

int stein_engine_device_bandwidth_stats(const stein_engine *e, int64_t *used, int64_t *redone) {
    if (!e) return STEIN_ERR_INVALID;
    if (used) *used = e->devbw_used;
    if (redone) *redone = e->devbw_redone;
    return STEIN_OK;
}

int stein_engine_prefetch_stats(const stein_engine *e, int64_t *begun, int64_t *used) {
    if (!e) return STEIN_ERR_INVALID;
    if (begun) *begun = e->prefetch_begun;
    if (used) *used = e->prefetch_used;
    return STEIN_OK;
}

int stein_engine_particles_changed(stein_engine *e) {
    if (!e) return STEIN_ERR_INVALID;
    e->x_all_current = false;
    e->bw_pending = false;
    if (e->ctx->guard_lag_owner == reinterpret_cast<const void *>(e->uid)) e->ctx->guard_lag_owner = nullptr;
    return STEIN_OK;
}

int stein_engine_phi_only(stein_engine *e) {
    if (!e) return STEIN_ERR_INVALID;
    stein_ctx *ctx = e->ctx;
    STEIN_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    float bw = 0.f;
    STEIN_TRY(step_bandwidth(e, &bw, e->world == 1, nullptr));
    return step_phi(e, bw, false);
}

int stein_engine_apply_phi(stein_engine *e) {
    if (!e) return STEIN_ERR_INVALID;
    stein_ctx *ctx = e->ctx;
    STEIN_REQUIRE(ctx, e->world == 1, "stein_engine_apply_phi: single-GPU engines only");
    STEIN_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    return step_optimizer(e);
}

int stein_engine_sumsq_dev(stein_engine *e, double **sumsq_dev) {
    if (!e || !sumsq_dev) return STEIN_ERR_INVALID;
    *sumsq_dev = e->sumsq;
    return STEIN_OK;
}

int stein_engine_set_hyper(stein_engine *e, double learning_rate, double decay, double p1, double p2) {
    if (!e) return STEIN_ERR_INVALID;
    e->lr = learning_rate;
    e->decay = decay;
    e->p1 = p1;
    e->p2 = p2;
    return STEIN_OK;
}

int stein_engine_set_bandwidth(stein_engine *e, float bandwidth) {
    if (!e) return STEIN_ERR_INVALID;
    if (!(bandwidth >= 0.0f) || bandwidth > 3.0e38f)
        return fail(e->ctx, STEIN_ERR_INVALID, "bandwidth must be finite and >= 0 (0 = median heuristic)");
    e->fixed_bw = bandwidth;
    return STEIN_OK;
}

int stein_engine_ipc_handle(stein_engine *e, void *handle_out) {
    if (!e || !handle_out) return STEIN_ERR_INVALID;
    stein_ctx *ctx = e->ctx;
    STEIN_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    STEIN_CHECK_CUDA(ctx, cudaIpcGetMemHandle(&h, e->X_all));
    static_assert(sizeof(h) == STEIN_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(handle_out, &h, sizeof(h));
    return STEIN_OK;
}

int stein_engine_set_peer_handles(stein_engine *e, const void *handles) {
    if (!e) return STEIN_ERR_INVALID;
    stein_ctx *ctx = e->ctx;
    if (!handles) {     // back to the all-gather hook
        STEIN_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
        STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int r = 0; r < e->world && r <= MAX_PEERS; ++r)
            if (r != e->rank && e->peer_X[r]) {
                cudaIpcCloseMemHandle(e->peer_X[r]);
                e->peer_X[r] = nullptr;
            }
        release_peer_reduce(e);
        e->peers_open = false;
        e->x_all_current = false;
        return STEIN_OK;
    }
    STEIN_REQUIRE(ctx, e->world > 1 && e->world <= MAX_PEERS + 1, "peer push needs 2..%d ranks", MAX_PEERS + 1);
    STEIN_REQUIRE(ctx, !e->peers_open, "peer handles already set");
    STEIN_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int r = 0; r < e->world; ++r) {
        if (r == e->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)r * STEIN_IPC_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        const cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) {
            for (int q = 0; q < r; ++q)
                if (q != e->rank && e->peer_X[q]) {
                    cudaIpcCloseMemHandle(e->peer_X[q]);
                    e->peer_X[q] = nullptr;
                }
            cudaGetLastError();
            return fail(ctx, STEIN_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(err));
        }
        e->peer_X[r] = (float *)p;
    }
    if (!e->barrier_word) STEIN_CHECK_CUDA(ctx, cudaMalloc(&e->barrier_word, 8));
    STEIN_CHECK_CUDA(ctx, cudaMemsetAsync(e->barrier_word, 0, 8, ctx->stream));
    // The small all-reduces of the iteration go through the mailboxes behind the particle buffers
    // from now on (STEIN_PEER_REDUCE=0 keeps them on the stein_comm hooks; same value on every rank).
    const char *env = getenv("STEIN_PEER_REDUCE");
    if (!(env && env[0] == '0')) {
        void *boxes[MAX_PEERS + 1];
        for (int r = 0; r < e->world; ++r)
            boxes[r] = reinterpret_cast<char *>(r == e->rank ? e->X_all : e->peer_X[r]) + e->mbox_offset;
        STEIN_TRY(peer_reduce_create(ctx, e->rank, e->world, boxes, e->mbox_epoch, &e->peer_reduce));
        ctx->peer_reduce = e->peer_reduce;
    }
    // the zero-fill of this rank's mailbox (engine creation) must be over before a peer writes
    // into it; the caller's exchange of the outcome is the barrier between the ranks
    STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    e->peers_open = true;
    e->x_all_current = false;
    return STEIN_OK;
}

int stein_engine_last(const stein_engine *e, float *median, float *bandwidth, double *phi_norm,
                      int32_t *sweeps) {
    if (!e) return STEIN_ERR_INVALID;
    stein_ctx *ctx = e->ctx;
    if (median) *median = e->last_med;
    if (bandwidth) *bandwidth = e->last_bw;
    if (sweeps) *sweeps = e->last_sweeps;
    if (phi_norm) {
        STEIN_CHECK_CUDA(ctx, cudaMemcpyAsync(e->h_sumsq, e->sumsq, 8, cudaMemcpyDeviceToHost, ctx->stream));
        STEIN_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *phi_norm = sqrt(*e->h_sumsq);
    }
    return STEIN_OK;
}

int stein_engine_get_state(stein_engine *e, int64_t *n_iters, double *learning_rate, void *m1_host,
                           void *m2_host, int is_f64) {
    if (!e) return STEIN_ERR_INVALID;
    if (n_iters) *n_iters = e->n_iters;
    if (learning_rate) *learning_rate = e->lr;
    if (m1_host) STEIN_TRY(download(e, e->m1, m1_host, is_f64));
    if (m2_host) STEIN_TRY(download(e, e->m2, m2_host, is_f64));
    return STEIN_OK;
}

int stein_engine_set_state(stein_engine *e, int64_t n_iters, double learning_rate, const void *m1_host,
                           const void *m2_host, int is_f64) {
    if (!e) return STEIN_ERR_INVALID;
    e->n_iters = n_iters;
    e->lr = learning_rate;
    if (m1_host) STEIN_TRY(upload(e, m1_host, is_f64, e->m1));
    if (m2_host) STEIN_TRY(upload(e, m2_host, is_f64, e->m2));
    STEIN_CHECK_CUDA(e->ctx, cudaStreamSynchronize(e->ctx->stream));
    return STEIN_OK;
}

}  // extern "C"
