// scores.cu -- kernel (1): batched scores grad_theta log p for every particle of
// the three built-in likelihoods, and the matching function_posterior kernels.
//
// Reference: the per-particle loop stein/samplers/stein_sampler.py:59-68 (n
// sess.run calls of tf.gradients(log_p, model_vars), abstract_stein_sampler.py:55)
// over the log_p graphs of
//   examples/linear_regression/main.py:25-31
//   examples/logistic_regression/main.py:28-49
//   examples/regression_neural_network/main.py:35-85
// Closed forms: SURVEY.md appendix A.4 (autograd-verified in oracle/).
//
// These are skinny (minibatch 50-1000 rows, 10-90 features): latency/HBM bound,
// not tensor-core work.  One CTA per particle, parameters and per-row
// intermediates in shared memory, fixed summation order (deterministic).
#include <algorithm>

#include "common.cuh"

namespace stein {

constexpr int SCORE_THREADS = 256;
constexpr int GLM_CHUNK = 1024;  // data rows staged per pass

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result valid in every thread; `red` has 32 floats
__device__ __forceinline__ float block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
    if (warp == 0) {
        t = warp_sum(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    t = red[0];
    __syncthreads();
    return t;
}

// MODEL 0: linear    e_b = y_b - x_b.w          grad_w = sum_b e_b x_b - w
// MODEL 1: logistic  e_b = y_b - sigmoid(x_b.w) grad_w = scale sum_b e_b x_b - alpha w
//                    grad_la = F/2 - alpha |w|^2 / 2 + (a-1) - b alpha
template <int MODEL>
__global__ void __launch_bounds__(SCORE_THREADS)
glm_score_kernel(const float *__restrict__ theta, int64_t F, int64_t ld, const float *__restrict__ Xd,
                 const float *__restrict__ y, int64_t N, float scale, float pa, float pb,
                 int fstride, int nparts, float *__restrict__ S) {
    extern __shared__ float sm[];
    float *w = sm;                    // F
    float *e = w + F;                 // GLM_CHUNK
    float *gacc = e + GLM_CHUNK;      // nparts * F
    __shared__ float red[32];
    const float *th = theta + (int64_t)blockIdx.x * ld;
    float *out = S + (int64_t)blockIdx.x * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = SCORE_THREADS / 32;

    for (int64_t f = tid; f < F; f += SCORE_THREADS) w[f] = th[f];
    for (int64_t q = tid; q < (int64_t)nparts * F; q += SCORE_THREADS) gacc[q] = 0.0f;
    __syncthreads();

    const int part = tid / fstride, fl = tid % fstride;
    for (int64_t c0 = 0; c0 < N; c0 += GLM_CHUNK) {
        const int cn = (int)min((int64_t)GLM_CHUNK, N - c0);
        // phase 1: residuals.  Few features (the linear example has 10): one THREAD per data row -- a warp per row
        // would leave most lanes idle and spend its time in the shuffle reduction (99 us -> see profiles/).
        if (F <= 16) {
            for (int b = tid; b < cn; b += SCORE_THREADS) {
                const float *xr = Xd + (c0 + b) * F;
                float z = 0.0f;
                for (int64_t f = 0; f < F; ++f) z = fmaf(xr[f], w[f], z);
                const float yy = y[c0 + b];
                e[b] = (MODEL == 0) ? (yy - z) : (yy - 1.0f / (1.0f + expf(-z)));
            }
        }
        // otherwise one warp per data row
        for (int b = warp; b < cn && F > 16; b += nwarps) {
            const float *xr = Xd + (c0 + b) * F;
            float z = 0.0f;
            for (int64_t f = lane; f < F; f += 32) z = fmaf(xr[f], w[f], z);
            z = warp_sum(z);
            if (lane == 0) {
                const float yy = y[c0 + b];
                e[b] = (MODEL == 0) ? (yy - z) : (yy - 1.0f / (1.0f + expf(-z)));
            }
        }
        __syncthreads();
        // phase 2: X^T e, thread = (row-part, feature)
        if (part < nparts) {
            for (int64_t f = fl; f < F; f += fstride) {
                float g = 0.0f;
                for (int b = part; b < cn; b += nparts) g = fmaf(e[b], Xd[(c0 + b) * F + f], g);
                gacc[(int64_t)part * F + f] += g;
            }
        }
        __syncthreads();
    }
    float wsq = 0.0f;
    for (int64_t f = tid; f < F; f += SCORE_THREADS) {
        float g = 0.0f;
        for (int p = 0; p < nparts; ++p) g += gacc[(int64_t)p * F + f];
        if (MODEL == 0) {
            out[f] = g - w[f];
        } else {
            const float alpha = expf(th[F]);
            out[f] = scale * g - alpha * w[f];
            wsq += w[f] * w[f];
        }
    }
    if (MODEL == 1) {
        wsq = block_sum(wsq, red);
        if (tid == 0) {
            const float alpha = expf(th[F]);
            out[F] = 0.5f * (float)F - 0.5f * alpha * wsq + (pa - 1.0f) - pb * alpha;
        }
    }
}

// theta = [log_lambda, log_gamma, w1 (F*H, row-major F x H), b1 (H), w2 (H), b2]
__global__ void __launch_bounds__(SCORE_THREADS)
bnn_score_kernel(const float *__restrict__ theta, int64_t F, int64_t H, int64_t ld,
                 const float *__restrict__ Xb, const float *__restrict__ yb, int64_t B, float n_train,
                 float pa, float pb, float *__restrict__ S) {
    extern __shared__ float sm[];
    float *w1 = sm;              // F*H
    float *b1 = w1 + F * H;      // H
    float *w2 = b1 + H;          // H
    float *Z = w2 + H;           // B*H
    float *delta = Z + B * H;    // B
    __shared__ float red[32];
    const float *th = theta + (int64_t)blockIdx.x * ld;
    float *out = S + (int64_t)blockIdx.x * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = SCORE_THREADS / 32;
    const int64_t o_w1 = 2, o_b1 = 2 + F * H, o_w2 = o_b1 + H, o_b2 = o_w2 + H;
    const int64_t Pw = F * H + 2 * H + 1;

    float wsq = 0.0f;
    for (int64_t q = tid; q < Pw; q += SCORE_THREADS) {
        const float v = th[o_w1 + q];
        wsq += v * v;
        if (q < F * H) w1[q] = v;
        else if (q < F * H + H) b1[q - F * H] = v;
        else if (q < F * H + 2 * H) w2[q - F * H - H] = v;
    }
    const float lam = expf(th[0]), gam = expf(th[1]), b2 = th[o_b2];
    wsq = block_sum(wsq, red);  // also the barrier that publishes the parameters

    // hidden pre-activations
    for (int64_t idx = tid; idx < B * H; idx += SCORE_THREADS) {
        const int64_t b = idx / H, h = idx % H;
        float z = b1[h];
        const float *xr = Xb + b * F;
        for (int64_t f = 0; f < F; ++f) z = fmaf(xr[f], w1[f * H + h], z);
        Z[idx] = z;
    }
    __syncthreads();
    // residuals
    const float nb = n_train / (float)B;
    float ressq = 0.0f, dsum = 0.0f;
    for (int64_t b = warp; b < B; b += nwarps) {
        float s = 0.0f;
        for (int64_t h = lane; h < H; h += 32) s = fmaf(fmaxf(Z[b * H + h], 0.0f), w2[h], s);
        s = warp_sum(s);
        if (lane == 0) {
            const float res = yb[b] - (s + b2);
            const float dl = nb * gam * res;
            delta[b] = dl;
            ressq += res * res;
            dsum += dl;
        }
    }
    ressq = block_sum(ressq, red);
    dsum = block_sum(dsum, red);
    const float inv_n = 1.0f / n_train;
    // output layer
    for (int64_t h = tid; h < H; h += SCORE_THREADS) {
        float s = 0.0f;
        for (int64_t b = 0; b < B; ++b) s = fmaf(fmaxf(Z[b * H + h], 0.0f), delta[b], s);
        out[o_w2 + h] = (s - lam * w2[h]) * inv_n;
    }
    __syncthreads();
    // back through the relu, in place
    for (int64_t idx = tid; idx < B * H; idx += SCORE_THREADS) {
        const int64_t b = idx / H, h = idx % H;
        Z[idx] = (Z[idx] > 0.0f) ? delta[b] * w2[h] : 0.0f;
    }
    __syncthreads();
    for (int64_t idx = tid; idx < F * H; idx += SCORE_THREADS) {
        const int64_t f = idx / H, h = idx % H;
        float s = 0.0f;
        for (int64_t b = 0; b < B; ++b) s = fmaf(Xb[b * F + f], Z[b * H + h], s);
        out[o_w1 + idx] = (s - lam * w1[idx]) * inv_n;
    }
    for (int64_t h = tid; h < H; h += SCORE_THREADS) {
        float s = 0.0f;
        for (int64_t b = 0; b < B; ++b) s += Z[b * H + h];
        out[o_b1 + h] = (s - lam * b1[h]) * inv_n;
    }
    if (tid == 0) {
        out[o_b2] = (dsum - lam * b2) * inv_n;
        out[1] = (nb * (0.5f * (float)B - 0.5f * gam * ressq) + (pa - 1.0f) - pb * gam) * inv_n;
        out[0] = (0.5f * (float)Pw - 0.5f * lam * wsq + (pa - 1.0f) - pb * lam) * inv_n;
    }
}

// out[i, t] = x_t . w_i     (logits / linear predictions; theta's first F columns)
__global__ void __launch_bounds__(256)
predict_linear_kernel(const float *__restrict__ theta, int64_t F, int64_t ld,
                      const float *__restrict__ Xt, int64_t N, float *__restrict__ out) {
    extern __shared__ float w[];
    const float *th = theta + (int64_t)blockIdx.y * ld;
    for (int64_t f = threadIdx.x; f < F; f += blockDim.x) w[f] = th[f];
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const float *xr = Xt + t * F;
    float z = 0.0f;
    for (int64_t f = 0; f < F; ++f) z = fmaf(xr[f], w[f], z);
    out[(int64_t)blockIdx.y * N + t] = z;
}

__global__ void __launch_bounds__(256)
predict_bnn_kernel(const float *__restrict__ theta, int64_t F, int64_t H, int64_t ld,
                   const float *__restrict__ Xt, int64_t N, float *__restrict__ out) {
    extern __shared__ float sm[];
    const int64_t Pw = F * H + 2 * H + 1;
    const float *th = theta + (int64_t)blockIdx.y * ld + 2;
    for (int64_t q = threadIdx.x; q < Pw; q += blockDim.x) sm[q] = th[q];
    __syncthreads();
    const float *w1 = sm, *b1 = sm + F * H, *w2 = b1 + H;
    const float b2 = sm[Pw - 1];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const float *xr = Xt + t * F;
    float pred = b2;
    for (int64_t h = 0; h < H; ++h) {
        float z = b1[h];
        for (int64_t f = 0; f < F; ++f) z = fmaf(xr[f], w1[f * H + h], z);
        pred = fmaf(fmaxf(z, 0.0f), w2[h], pred);
    }
    out[(int64_t)blockIdx.y * N + t] = pred;
}

// S_i = sum_k r_ik (mu_k - x_i) / sigma2 ; one warp per particle
__global__ void __launch_bounds__(256)
gmm_score_kernel(const float *__restrict__ theta, int64_t n, int64_t d, int64_t ld,
                 const float *__restrict__ mu, int ncomp, float inv_s2, float *__restrict__ S) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const float *x = theta + i * ld;
    float *out = S + i * ld;
    if (ncomp == 1 && mu == nullptr) {
        for (int64_t c = lane; c < d; c += 32) out[c] = -x[c] * inv_s2;
        return;
    }
    // responsibilities from squared distances to the means (max-shifted softmax)
    float logit[8];
    float mx = -3.4e38f;
    for (int k = 0; k < ncomp; ++k) {
        float s = 0.0f;
        for (int64_t c = lane; c < d; c += 32) {
            const float df = x[c] - mu[(int64_t)k * d + c];
            s = fmaf(df, df, s);
        }
        s = warp_sum(s);
        logit[k] = -0.5f * s * inv_s2;
        mx = fmaxf(mx, logit[k]);
    }
    float den = 0.0f;
    for (int k = 0; k < ncomp; ++k) {
        logit[k] = expf(logit[k] - mx);
        den += logit[k];
    }
    for (int64_t c = lane; c < d; c += 32) {
        float g = 0.0f;
        for (int k = 0; k < ncomp; ++k) g = fmaf(logit[k], mu[(int64_t)k * d + c] - x[c], g);
        out[c] = g / den * inv_s2;
    }
}

static int glm_launch(stein_ctx *ctx, int model, const float *theta, int64_t n, int64_t F, int64_t ld,
                      const float *Xd, const float *y, int64_t N, float scale, float pa, float pb,
                      float *S) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, theta && Xd && y && S, "null pointer");
    const int64_t dparam = F + (model == 1 ? 1 : 0);
    STEIN_REQUIRE(ctx, n >= 1 && F >= 1 && N >= 1 && ld >= dparam, "bad shape n=%lld F=%lld N=%lld ld=%lld",
                  (long long)n, (long long)F, (long long)N, (long long)ld);
    int fstride = 8;            // smallest power of two >= F (threads of one row-part), at most the block
    while (fstride < F && fstride < SCORE_THREADS) fstride *= 2;
    const int nparts = SCORE_THREADS / fstride;
    const size_t smem = sizeof(float) * (size_t)(F + GLM_CHUNK + (int64_t)nparts * F);
    STEIN_REQUIRE(ctx, smem <= 200 * 1024, "F=%lld too large for the score kernel", (long long)F);
    if (model == 0) {
        STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(glm_score_kernel<0>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        glm_score_kernel<0><<<(unsigned)n, SCORE_THREADS, smem, ctx->stream>>>(
            theta, F, ld, Xd, y, N, scale, pa, pb, fstride, nparts, S);
    } else {
        STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(glm_score_kernel<1>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        glm_score_kernel<1><<<(unsigned)n, SCORE_THREADS, smem, ctx->stream>>>(
            theta, F, ld, Xd, y, N, scale, pa, pb, fstride, nparts, S);
    }
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

}  // namespace stein

using namespace stein;

extern "C" {

int stein_score_linear(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t F, int64_t ld,
                       const float *Xd_dev, const float *y_dev, int64_t N, float *S_dev) {
    return glm_launch(ctx, 0, theta_dev, n, F, ld, Xd_dev, y_dev, N, 1.0f, 0.f, 0.f, S_dev);
}

int stein_score_logistic(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t F, int64_t ld,
                         const float *Xb_dev, const float *yb_dev, int64_t B, double n_train,
                         double prior_a, double prior_b, float *S_dev) {
    // logistic main.py:46 -- log_l * (n_train / n_batch)
    return glm_launch(ctx, 1, theta_dev, n, F, ld, Xb_dev, yb_dev, B, (float)(n_train / (double)B),
                      (float)prior_a, (float)prior_b, S_dev);
}

int stein_score_bnn(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t F, int64_t H, int64_t ld,
                    const float *Xb_dev, const float *yb_dev, int64_t B, double n_train, double prior_a,
                    double prior_b, float *S_dev) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, theta_dev && Xb_dev && yb_dev && S_dev, "null pointer");
    const int64_t dparam = 2 + F * H + 2 * H + 1;
    STEIN_REQUIRE(ctx, n >= 1 && F >= 1 && H >= 1 && B >= 1 && ld >= dparam,
                  "bad shape n=%lld F=%lld H=%lld B=%lld ld=%lld", (long long)n, (long long)F,
                  (long long)H, (long long)B, (long long)ld);
    const size_t smem = sizeof(float) * (size_t)(F * H + 2 * H + B * H + B);
    if (smem > 200 * 1024)
        return fail(ctx, STEIN_ERR_UNSUPPORTED,
                    "BNN score kernel needs %zu B of shared memory (F*H + B*H too large)", smem);
    STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(bnn_score_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bnn_score_kernel<<<(unsigned)n, SCORE_THREADS, smem, ctx->stream>>>(
        theta_dev, F, H, ld, Xb_dev, yb_dev, B, (float)n_train, (float)prior_a, (float)prior_b, S_dev);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

int stein_score_gaussian_mixture(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t d, int64_t ld,
                                 const float *mu_dev, int64_t ncomp, double sigma2, float *S_dev) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, theta_dev && S_dev, "null pointer");
    STEIN_REQUIRE(ctx, n >= 1 && d >= 1 && ld >= d, "bad shape");
    STEIN_REQUIRE(ctx, ncomp >= 1 && ncomp <= 8 && (mu_dev != nullptr || ncomp == 1),
                  "1..8 components, means required when ncomp > 1");
    STEIN_REQUIRE(ctx, sigma2 > 0, "sigma2 must be positive");
    gmm_score_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(theta_dev, n, d, ld, mu_dev, (int)ncomp,
                                                                     (float)(1.0 / sigma2), S_dev);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

int stein_predict_linear(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t F, int64_t ld,
                         const float *Xt_dev, int64_t N, float *out_dev) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, theta_dev && Xt_dev && out_dev, "null pointer");
    STEIN_REQUIRE(ctx, n >= 1 && n <= 65535 && F >= 1 && N >= 1 && ld >= F, "bad shape");
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)n);
    predict_linear_kernel<<<grid, 256, sizeof(float) * F, ctx->stream>>>(theta_dev, F, ld, Xt_dev, N,
                                                                         out_dev);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

int stein_predict_bnn(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t F, int64_t H, int64_t ld,
                      const float *Xt_dev, int64_t N, float *out_dev) {
    STEIN_REQUIRE(ctx, ctx != nullptr, "null ctx");
    STEIN_REQUIRE(ctx, theta_dev && Xt_dev && out_dev, "null pointer");
    const int64_t Pw = F * H + 2 * H + 1;
    STEIN_REQUIRE(ctx, n >= 1 && n <= 65535 && F >= 1 && H >= 1 && N >= 1 && ld >= Pw + 2, "bad shape");
    const size_t smem = sizeof(float) * (size_t)Pw;
    STEIN_REQUIRE(ctx, smem <= 200 * 1024, "F*H too large");
    STEIN_CHECK_CUDA(ctx, cudaFuncSetAttribute(predict_bnn_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)n);
    predict_bnn_kernel<<<grid, 256, smem, ctx->stream>>>(theta_dev, F, H, ld, Xt_dev, N, out_dev);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

}  // extern "C"
