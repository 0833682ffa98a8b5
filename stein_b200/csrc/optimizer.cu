// optimizer.cu -- kernel (4): Frobenius-norm clip + optimizer step, fused.
//
// Reference: stein/samplers/abstract_stein_sampler.py:125-126
//                phi *= 10. / max(10., np.linalg.norm(phi));  theta += gd.update(phi)
//            stein/optimizers/adam_gradient_descent.py:41-58
//            stein/optimizers/adagrad_gradient_descent.py:34-44
// HBM-bound: Adam reads phi, X, mu, nu and writes X, mu, nu (7 x 4 B/element).
//
// Particle-sharded runs: the updated rows are what every other rank needs
// at the start of the next iteration.  When the engines of a node have exchanged CUDA-IPC handles
// of their particle buffers, the step kernel stores each new row into the peers' buffers over
// NVLink as it produces it (PeerTargets), so the all-gather of X disappears behind the update.
#include "common.cuh"

namespace stein {

__device__ __forceinline__ float clip_scale(const double *sumsq) {
    // 10. / max(10., ||phi||_F)
    const double nrm = sqrt(*sumsq);
    return (float)(10.0 / fmax(10.0, nrm));
}

// first == 1: mu = phi, nu = phi^2 (adam_gradient_descent.py:45-46, NOT (1-beta) phi)
__global__ void __launch_bounds__(256)
clip_adam_kernel(float4 *__restrict__ X, const float4 *__restrict__ phi, float4 *__restrict__ mu,
                 float4 *__restrict__ nu, int64_t count4, const double *__restrict__ sumsq, float lr,
                 float b1, float b2, float c1, float c2, int first, const PeerTargets peers) {
    const float scale = clip_scale(sumsq);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p4 = phi[i];
        float4 x = X[i];
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f), v = m;
        if (!first) {
            m = mu[i];
            v = nu[i];
        }
        const float p[4] = {p4.x * scale, p4.y * scale, p4.z * scale, p4.w * scale};
        float mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
        float xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (first) {
                mm[k] = p[k];
                vv[k] = p[k] * p[k];
            } else {
                mm[k] = b1 * mm[k] + (1.0f - b1) * p[k];
                vv[k] = b2 * vv[k] + (1.0f - b2) * (p[k] * p[k]);
            }
            // mup / (1e-8 + sqrt(nup)) * lr, eps outside the sqrt (:53-55)
            const float mup = mm[k] * c1, nup = vv[k] * c2;
            xx[k] += mup / (1e-8f + sqrtf(nup)) * lr;
        }
        const float4 xn = make_float4(xx[0], xx[1], xx[2], xx[3]);
        X[i] = xn;
        for (int q = 0; q < peers.n; ++q) peers.dst[q][i] = xn;      // same rows in the peers' X_all
        mu[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        nu[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
}

__global__ void __launch_bounds__(256)
clip_adagrad_kernel(float4 *__restrict__ X, const float4 *__restrict__ phi, float4 *__restrict__ hist,
                    int64_t count4, const double *__restrict__ sumsq, float lr, float alpha,
                    int first, const PeerTargets peers) {
    const float scale = clip_scale(sumsq);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p4 = phi[i];
        float4 x = X[i];
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!first) h = hist[i];
        const float p[4] = {p4.x * scale, p4.y * scale, p4.z * scale, p4.w * scale};
        float hh[4] = {h.x, h.y, h.z, h.w};
        float xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hh[k] = first ? p[k] * p[k] : alpha * hh[k] + (1.0f - alpha) * (p[k] * p[k]);
            // phi / (1e-6 + sqrt(hist)) * lr  (:44)
            xx[k] += p[k] / (1e-6f + sqrtf(hh[k])) * lr;
        }
        const float4 xn = make_float4(xx[0], xx[1], xx[2], xx[3]);
        X[i] = xn;
        for (int q = 0; q < peers.n; ++q) peers.dst[q][i] = xn;
        hist[i] = make_float4(hh[0], hh[1], hh[2], hh[3]);
    }
}

static unsigned step_grid(const stein_ctx *ctx, int64_t count4) {
    const int64_t want = (count4 + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

int clip_adam_step(stein_ctx *ctx, float *X_dev, const float *phi_dev, float *mu_dev, float *nu_dev,
                   int64_t count, const double *sumsq_dev, double learning_rate, double beta_1, double beta_2,
                   int64_t n_iters, const PeerTargets &peers) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, X_dev && phi_dev && mu_dev && nu_dev && sumsq_dev, "null pointer");
    STEIN_REQUIRE(ctx, count > 0 && count % 4 == 0, "count=%lld must be a positive multiple of 4",
                  (long long)count);
    // adam_gradient_descent.py:52-54: n_iters is incremented before the bias correction
    const double t = (double)(n_iters + 1);
    const float c1 = (float)(1.0 / (1.0 - pow(beta_1, t)));
    const float c2 = (float)(1.0 / (1.0 - pow(beta_2, t)));
    clip_adam_kernel<<<step_grid(ctx, count / 4), 256, 0, ctx->stream>>>(
        (float4 *)X_dev, (const float4 *)phi_dev, (float4 *)mu_dev, (float4 *)nu_dev, count / 4,
        sumsq_dev, (float)learning_rate, (float)beta_1, (float)beta_2, c1, c2, n_iters == 0 ? 1 : 0, peers);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

int clip_adagrad_step(stein_ctx *ctx, float *X_dev, const float *phi_dev, float *hist_dev, int64_t count,
                      const double *sumsq_dev, double learning_rate, double alpha, int64_t n_iters,
                      const PeerTargets &peers) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, X_dev && phi_dev && hist_dev && sumsq_dev, "null pointer");
    STEIN_REQUIRE(ctx, count > 0 && count % 4 == 0, "count=%lld must be a positive multiple of 4",
                  (long long)count);
    clip_adagrad_kernel<<<step_grid(ctx, count / 4), 256, 0, ctx->stream>>>(
        (float4 *)X_dev, (const float4 *)phi_dev, (float4 *)hist_dev, count / 4, sumsq_dev,
        (float)learning_rate, (float)alpha, n_iters == 0 ? 1 : 0, peers);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

}  // namespace stein

using namespace stein;

extern "C" {

int stein_clip_adam_step(stein_ctx *ctx, float *X_dev, const float *phi_dev, float *mu_dev,
                         float *nu_dev, int64_t count, const double *sumsq_dev, double learning_rate,
                         double beta_1, double beta_2, int64_t n_iters) {
    return clip_adam_step(ctx, X_dev, phi_dev, mu_dev, nu_dev, count, sumsq_dev, learning_rate, beta_1, beta_2,
                          n_iters, PeerTargets{});
}

int stein_clip_adagrad_step(stein_ctx *ctx, float *X_dev, const float *phi_dev, float *hist_dev,
                            int64_t count, const double *sumsq_dev, double learning_rate, double alpha,
                            int64_t n_iters) {
    return clip_adagrad_step(ctx, X_dev, phi_dev, hist_dev, count, sumsq_dev, learning_rate, alpha, n_iters,
                             PeerTargets{});
}

}  // extern "C"
