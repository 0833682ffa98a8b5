/*
 * stein_b200.h -- C ABI of libstein_b200.so, the B200 (sm_100a) implementation
 * of one SVGD iteration of JamesBrofos/Stein.
 *
 * The reference has no native boundary (it is pure Python on TensorFlow 1.12);
 * each entry point below names the reference Python interface it replaces
 * (file:line relative to the reference tree).  INTEGRATION.md shows the ctypes
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - one process drives one GPU: the library keeps per-process scratch (median arena, slot plan,
 *     window hint), so contexts of different devices must live in different processes;
 *   - every function returns 0 (STEIN_OK) or a negative STEIN_ERR_* code;
 *     stein_last_error() returns the text of the last failure on that context;
 *   - pointers named *_dev are DEVICE pointers, *_host are HOST pointers;
 *   - particle matrices are row-major fp32 with an explicit leading dimension
 *     `ld` (in floats).  PADDING CONTRACT for device matrices handed to the
 *     stein_* device-pointer functions: ld % 32 == 0, ld >= d, the row count
 *     allocated is a multiple of 128, and every pad element (columns >= d,
 *     rows >= n) is ZERO.  stein_ld()/stein_rows_padded() compute the sizes.
 *     The stein_engine_* functions take plain, unpadded HOST arrays and do the
 *     padding themselves;
 *   - all device work is enqueued on the context's stream; functions that
 *     return a value to the host synchronise that stream.
 */
#ifndef STEIN_B200_H
#define STEIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STEIN_OK 0
#define STEIN_ERR_INVALID (-1)     /* bad argument / padding contract violated   */
#define STEIN_ERR_CUDA (-2)        /* CUDA runtime/driver error                   */
#define STEIN_ERR_UNSUPPORTED (-3) /* shape not supported by the requested path   */
#define STEIN_ERR_NOMEM (-4)
#define STEIN_ERR_COMM (-5)        /* a collective hook failed                    */
#define STEIN_ERR_INTERNAL (-6)

#define STEIN_ABI_VERSION 2

/* phi-kernel implementations (stein_ctx_set_phi_impl).  AUTO: leading dimension 256 -> the CTA-pair
 * kernel, FLASH_TC4 (fast) or FLASH_TC5 (precise) or the FFMA path picked per call from the conditioning
 * of the cloud (stein_ctx_phi_route); 128 -> FLASH_TC (same guard); 512 / 768 / 1024 -> the panel kernels
 * (same guard); anything else -> DENSE_SIMT (the engine pads the rows of >= 2048 particles of up
 * to 256 coordinates to 128 / 256 floats). */
#define STEIN_PHI_AUTO 0
#define STEIN_PHI_DENSE_SIMT 1 /* materialises K for the local row block; FP32 FFMA */
#define STEIN_PHI_FLASH_TC 2   /* tcgen05/TMEM/TMA fused kernel; never stores K   */
#define STEIN_PHI_FLASH_TC2 3  /* same, CTA pairs (cta_group::2, M = 256); d <= 256 padded to 256 */
#define STEIN_PHI_FLASH_TC3 4  /* CTA pairs, GEMM2 as one FP16 pass + two FP8 passes instead of 3 BF16 */
#define STEIN_PHI_FLASH_TC4 5  /* CTA pairs, both GEMMs as FP16 + 2 x FP8 ("fast" route) */
#define STEIN_PHI_FLASH_TC5 6  /* CTA pairs, both GEMMs as 3 x FP16 on a 2-term split ("precise" route, fp32-Gram accuracy) */

/* median implementations (stein_ctx_set_median_impl).  AUTO: n <= 2048 -> all n*n keys once and a
 * device-side select; leading dimension 128 / 256 and n*n >= 2^24 -> TC; otherwise FFMA. */
#define STEIN_MEDIAN_AUTO 0
#define STEIN_MEDIAN_FFMA 1 /* every sweep in contract arithmetic on the FP32 pipe          */
#define STEIN_MEDIAN_TC 2   /* tcgen05 filter sweep + contract recomputation of candidates  */
#define STEIN_MEDIAN_TC1 3  /* same, single-CTA sweep kernel (the default sweeps with CTA pairs) */

/* optimizer kinds (stein_engine_create) */
#define STEIN_OPT_ADAM 0    /* stein/optimizers/adam_gradient_descent.py    */
#define STEIN_OPT_ADAGRAD 1 /* stein/optimizers/adagrad_gradient_descent.py */

typedef struct stein_ctx stein_ctx;
typedef struct stein_engine stein_engine;

/* Collective hooks for particle-sharded runs (one process per GPU).  The host
 * program supplies them (torch.distributed/NCCL in the Python package); all
 * operate in place on device memory and must be ordered on the ctx stream.   */
typedef struct stein_comm {
    int32_t rank;
    int32_t world;
    void *user;
    int (*allreduce_sum_u64)(void *user, void *buf_dev, int64_t count);
    int (*allreduce_sum_f64)(void *user, void *buf_dev, int64_t count);
    /* gathers `count` floats from every rank into recv_dev[rank*count ...] */
    int (*allgather_f32)(void *user, const void *send_dev, void *recv_dev, int64_t count);
    /* OPTIONAL (may be NULL): the same gather enqueued on the given CUDA stream instead of the ctx
     * stream, through resources that may run concurrently with the other hooks (the built-in NCCL
     * transport uses a second communicator).  Lets the score shards travel while the median runs. */
    int (*allgather_f32_on)(void *user, const void *send_dev, void *recv_dev, int64_t count, void *cuda_stream);
} stein_comm;

/* ---- context ------------------------------------------------------------- */
int stein_abi_version(void);
int stein_ctx_create(stein_ctx **out, int device, void *cuda_stream /* may be NULL */);
int stein_ctx_destroy(stein_ctx *ctx);
int stein_ctx_set_stream(stein_ctx *ctx, void *cuda_stream);
int stein_ctx_set_comm(stein_ctx *ctx, const stein_comm *comm /* NULL = single GPU */);
/* Built-in hooks on NCCL (resolved at run time from the libnccl.so.2 loaded in the process):
 * rank 0 calls stein_nccl_unique_id and hands the 128 bytes to every rank by any means; every
 * rank then calls stein_ctx_init_nccl, which creates the communicator (collective) and installs
 * hooks that enqueue ncclAllGather / ncclAllReduce on the ctx stream.  A second id (optional,
 * id_side) creates a second communicator for the allgather_f32_on hook.  world == 1 clears them. */
#define STEIN_NCCL_ID_BYTES 128
int stein_nccl_unique_id(void *id_out);
int stein_ctx_init_nccl(stein_ctx *ctx, int rank, int world, const void *id, const void *id_side /* or NULL */);
int stein_ctx_set_phi_impl(stein_ctx *ctx, int impl);
int stein_ctx_set_median_impl(stein_ctx *ctx, int impl);
const char *stein_last_error(const stein_ctx *ctx /* NULL = last error of any ctx */);
/* number of kernels of this library launched on ctx since creation */
int64_t stein_ctx_launch_count(const stein_ctx *ctx);

/* Which route the last guarded (STEIN_PHI_AUTO) phi call took on this context:
 * *route = 0 fast (FLASH_TC4 arithmetic), 1 precise (FLASH_TC5), 2 the FP32 FFMA path (DENSE_SIMT:
 * the reference's own arithmetic, for clouds no tensor-core route serves to 1e-4), -1 no guarded call yet;
 * *kappa = max_i |x_i - mean|^2 / h^2 of that call's cloud; *predicted_fast_error = the guard's
 * estimate of the relative error of phi on the fast route.  A faster route is taken while its
 * predicted error stays below the tolerance (default 5e-5; stein_ctx_set_phi_guard_tol / environment
 * STEIN_PHI_GUARD_TOL).  The decision is taken on the host from one 8-byte device value per call; every
 * rank of a sharded run sees the same particles and takes the same route.
 * No reference counterpart: the reference evaluates stein/kernels/abstract_kernel.py:33-35 in fp32. */
int stein_ctx_phi_route(stein_ctx *ctx, int32_t *route, float *kappa, float *predicted_fast_error);
int stein_ctx_set_phi_guard_tol(stein_ctx *ctx, float tol);

/* Optional per-region device timing (CUDA events on the ctx stream), used by
 * bench.py for the live roofline figure and the per-phase timeline of an iteration.
 * read() synchronises the stream, returns the accumulated milliseconds and the number
 * of timed intervals of that region since the last reset, and resets. */
#define STEIN_REGION_PHI 0        /* phi main kernel(s)                                          */
#define STEIN_REGION_SWEEP 1      /* median distance sweep(s)                                    */
#define STEIN_REGION_MEDIAN 2     /* whole median call: split, pilot, sweep, band, select, host round trip (contains 1 and part of 6) */
#define STEIN_REGION_PHI_PREP 3   /* phi call up to its main kernel: centring, guard, operand arrays */
#define STEIN_REGION_PHI_TAIL 4   /* finalize + sum(phi^2)                                        */
#define STEIN_REGION_OPT 5        /* clip + optimizer kernel (+ peer push)                        */
#define STEIN_REGION_COLL 6       /* collectives of the iteration (all-reduces, barrier word, all-gathers on the ctx stream) */
#define STEIN_REGION_HEAD 7       /* start of the iteration: barrier / all-gather of the particles, row norms (+ error budgets, scale and FP16 split of the median when its tensor-core route applies: one read of the particles) */
#define STEIN_REGION_COUNT 8
/* enable: 0 off; 1 = regions PHI and SWEEP only (two event pairs per iteration: what bench.py keeps on inside
 * its timed region); 2 = all regions (the per-phase timeline, measured in a separate pass) */
int stein_ctx_profile_enable(stein_ctx *ctx, int enable);
int stein_ctx_profile_read(stein_ctx *ctx, int region, double *ms_total, int64_t *launches);
/* Fine-grained timeline of the ctx stream (tools/step_trace.py): while enabled the library records a
 * labelled event after each stage of an iteration; _read waits for the device and writes
 * "label<TAB>milliseconds since the previous mark" lines (in enqueue order) into buf.             */
int stein_ctx_trace_enable(stein_ctx *ctx, int enable);
int stein_ctx_trace_read(stein_ctx *ctx, char *buf, int64_t cap);

int64_t stein_ld(int64_t d);           /* leading dimension for d columns     */
int64_t stein_rows_padded(int64_t n);  /* rows to allocate for n particles    */

/* ---- kernel (2): squared distances, exact median, bandwidth --------------
 * replaces  AbstractKernel.__init__ graph  stein/kernels/abstract_kernel.py:30-40
 *           compute_median                  stein/utilities/compute_median.py:4-16 */

/* r_i = sum_k x_ik^2 in the contract order (fma chain over k ascending). */
int stein_row_norms(stein_ctx *ctx, const float *X_dev, int64_t n, int64_t d, int64_t ld,
                    float *r_dev);

/* One histogram sweep over the upper-triangular 128x128 tiles
 * [tile_begin, tile_end) of D (tile t <-> (I,J), I<=J, row-major order; off-
 * diagonal tiles count twice).  Keys are the order-preserving u32 image of the
 * fp32 distance.  counts_dev[0] += #(key < key_lo); counts_dev[1+b] += #keys
 * with (key-key_lo)>>shift == b, b < nbins (nbins <= 16384).  counts are u64
 * and are ACCUMULATED (zero them first).  D_ij = fl(fl(r_i+r_j) - 2 g_ij),
 * g_ij an fma chain over k ascending -- bit-identical to oracle/svgd_oracle.c. */
int stein_sqdist_hist(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n,
                      int64_t d, int64_t ld, int64_t tile_begin, int64_t tile_end,
                      uint32_t key_lo, uint32_t shift, uint32_t nbins, uint64_t *counts_dev);

/* number of upper-triangular tiles for n particles, and the (I, J) tile
 * coordinates of linear tile index t (pure host functions) */
int64_t stein_num_tiles(int64_t n);
int stein_tile_coords(int64_t t, int64_t n, int32_t *I, int32_t *J);

/* Exact median of all n*n entries of D (even count: fp32 mean of the two middle
 * values), compute_median.py:9-15.  Uses the ctx collective hooks when set
 * (tiles are dealt round-robin to ranks, histograms all-reduced), so every rank
 * gets the same bits.  mid_host[2] (optional) receives the middle value(s);
 * sweeps_host (optional) the number of full distance sweeps that were needed. */
int stein_median_sqdist(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n,
                        int64_t d, int64_t ld, float *median_host, float *mid_host,
                        int32_t *sweeps_host);

/* compute_median() of an arbitrary device array of m fp32 values (the public
 * stein/utilities/compute_median.py:4-16 entry point; m = n*n for a dense D). */
int stein_median_values(stein_ctx *ctx, const float *V_dev, int64_t m, float *median_host);

/* h = sqrt(med / ln n) in fp32, abstract_kernel.py:40.  Pure host function. */
float stein_bandwidth(float median, int64_t n_particles);

/* Pure host helper used by stein_median_sqdist (exposed for tests): given the
 * counts of one sweep, locate 0-based ascending `rank`.  Returns 1 and sets
 * *key_out when the key is determined (shift == 0), 0 when another sweep over
 * the narrowed window (*key_lo_out, *shift_out, *nbins_out) is needed, and -1
 * when the rank lies outside the window (below: *key_lo_out = 0, above: 1). */
int stein_median_narrow(const uint64_t *counts_host, uint32_t key_lo, uint32_t shift,
                        uint32_t nbins, uint64_t rank, uint32_t *key_out,
                        uint32_t *key_lo_out, uint32_t *shift_out, uint32_t *nbins_out);

uint32_t stein_float_to_key(float f);
float stein_key_to_float(uint32_t key);

/* ---- kernel (3): phi ------------------------------------------------------
 * replaces  SquaredExponentialKernel.kernel_and_grad
 *                 stein/kernels/squared_exponential_kernel.py:22-35
 *           AbstractSteinSampler.compute_phi
 *                 stein/samplers/abstract_stein_sampler.py:100-105
 * phi_i = ( sum_j K_ij s_j + (x_i sum_j K_ij - sum_j K_ij x_j)/h^2 ) / n
 * for the local rows [row_begin, row_begin+n_local) against all n_total
 * columns; K_ij = exp(-D_ij / h^2 / 2), h^2 = bandwidth^2.  phi_dev is
 * n_local(padded) x ld.  sumsq_dev (double, device) receives sum(phi^2) over
 * the local rows (the Frobenius-norm partial of abstract_stein_sampler.py:125).
 * workspace: stein_phi_workspace_bytes() bytes of device memory.              */
int64_t stein_phi_workspace_bytes(const stein_ctx *ctx, int64_t n_local, int64_t n_total,
                                  int64_t d);
int stein_phi(stein_ctx *ctx, const float *X_all_dev, const float *S_all_dev,
              const float *r_all_dev, int64_t n_total, int64_t d, int64_t ld, int64_t row_begin,
              int64_t n_local, float bandwidth, void *workspace_dev, int64_t workspace_bytes,
              float *phi_dev, double *sumsq_dev);

/* Small-n compatibility path with the reference's return values: dense K
 * (n x n, leading dimension ldk >= n) and dK (n x ld), given the bandwidth.   */
int stein_kernel_and_grad(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n,
                          int64_t d, int64_t ld, float bandwidth, float *K_dev, int64_t ldk,
                          float *dK_dev, void *workspace_dev, int64_t workspace_bytes);

/* ---- other kernel operators through the plugin point ----------------
 * replaces  AbstractKernel.kernel_and_grad  stein/kernels/abstract_kernel.py:45-62  for an inverse
 * multiquadric kernel K_ij = (1 + D_ij / h^2)^beta (beta < 0), with the reference's gradient recipe
 * dK = -0.5 d(sum K)/d theta (squared_exponential_kernel.py:23,32).  Dense, small n, same layout as
 * stein_kernel_and_grad.                                                                     */
int stein_imq_kernel_and_grad(stein_ctx *ctx, const float *X_dev, const float *r_dev, int64_t n,
                              int64_t d, int64_t ld, float bandwidth, float beta, float *K_dev,
                              int64_t ldk, float *dK_dev, void *workspace_dev, int64_t workspace_bytes);
/* replaces  AbstractSteinSampler.compute_phi  abstract_stein_sampler.py:100-105  for ANY kernel
 * operator: phi = (K S + dK) / n from a dense K (rows_padded(n) x ldk, pads zero) and dK (n x ld) as a
 * kernel_and_grad returned them; sum(phi^2) -> *sumsq_dev.  workspace: rows_padded(n)*ld*4 + 9 480 bytes. */
int stein_phi_from_kernel(stein_ctx *ctx, const float *K_dev, int64_t ldk, const float *dK_dev,
                          const float *S_dev, int64_t n, int64_t d, int64_t ld, void *workspace_dev,
                          int64_t workspace_bytes, float *phi_dev, double *sumsq_dev);

/* ---- kernel (4): norm clip + optimizer step --------------------------------
 * replaces  AbstractSteinSampler.update_particles  abstract_stein_sampler.py:125-126
 *           AdamGradientDescent.update             adam_gradient_descent.py:41-58
 *           AdagradGradientDescent.update          adagrad_gradient_descent.py:34-44
 * phi is scaled by 10/max(10, sqrt(*sumsq_dev)) (sumsq already all-reduced),
 * the moments are updated (first call: mu=phi, nu=phi^2 / hist=phi^2) and the
 * step is ADDED to X.  n_iters is the optimizer's counter BEFORE this call.
 * count = rows_padded * ld elements (pad elements stay zero).                 */
int stein_clip_adam_step(stein_ctx *ctx, float *X_dev, const float *phi_dev, float *mu_dev,
                         float *nu_dev, int64_t count, const double *sumsq_dev,
                         double learning_rate, double beta_1, double beta_2, int64_t n_iters);
int stein_clip_adagrad_step(stein_ctx *ctx, float *X_dev, const float *phi_dev, float *hist_dev,
                            int64_t count, const double *sumsq_dev, double learning_rate,
                            double alpha, int64_t n_iters);

/* ---- kernel (1): batched scores of the built-in likelihoods ----------------
 * replaces the per-particle loop  stein/samplers/stein_sampler.py:59-68  (n
 * sess.run calls of tf.gradients(log_p, model_vars)) for the three example
 * models.  theta_dev / S_dev: n(padded) x ld in the flat layout of
 * stein/utilities/converters.py:40-53.  Data are plain dense device arrays.   */
/* examples/linear_regression/main.py:25-31.   theta = [w (F)]                 */
int stein_score_linear(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t F, int64_t ld,
                       const float *Xd_dev, const float *y_dev, int64_t N, float *S_dev);
/* examples/logistic_regression/main.py:28-49. theta = [w (F), log_alpha]      */
int stein_score_logistic(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t F, int64_t ld,
                         const float *Xb_dev, const float *yb_dev, int64_t B, double n_train,
                         double prior_a, double prior_b, float *S_dev);
/* examples/regression_neural_network/main.py:35-85.
 * theta = [log_lambda, log_gamma, w1 (F*H), b1 (H), w2 (H), b2]                */
int stein_score_bnn(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t F, int64_t H,
                    int64_t ld, const float *Xb_dev, const float *yb_dev, int64_t B,
                    double n_train, double prior_a, double prior_b, float *S_dev);

/* Synthetic targets of BASELINE.json configs D/E (no data): isotropic Gaussian
 * mixture with `ncomp` equal-weight components, means mu_dev (ncomp x d, dense)
 * and common variance sigma2: S_i = sum_k r_ik (mu_k - x_i) / sigma2, r_ik the
 * responsibilities.  ncomp = 1 with mu_dev = NULL is the standard normal target
 * S = -X / sigma2.                                                             */
int stein_score_gaussian_mixture(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t d,
                                 int64_t ld, const float *mu_dev, int64_t ncomp, double sigma2,
                                 float *S_dev);

/* ---- function_posterior for the built-in models ----------------------------
 * replaces  AbstractSteinSampler.function_posterior  abstract_stein_sampler.py:157-159
 * out_dev is n x N row-major (particle-major), fp32.                           */
int stein_predict_linear(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t F, int64_t ld,
                         const float *Xt_dev, int64_t N, float *out_dev);
int stein_predict_bnn(stein_ctx *ctx, const float *theta_dev, int64_t n, int64_t F, int64_t H,
                      int64_t ld, const float *Xt_dev, int64_t N, float *out_dev);

/* ---- engine: device-resident particles + optimizer state -------------------
 * replaces  AbstractSteinSampler.update_particles(grads_array)
 *                 stein/samplers/abstract_stein_sampler.py:107-127
 * with HOST buffers at the boundary (what the reference's NumPy caller holds).
 * With collective hooks set on ctx the engine owns rows
 * [row_begin, row_begin+n_local) of the n_total particles.                     */
int stein_engine_create(stein_engine **out, stein_ctx *ctx, int64_t n_total, int64_t d,
                        int optimizer, double learning_rate, double decay, double p1, double p2);
int stein_engine_destroy(stein_engine *eng);
int stein_engine_local_rows(const stein_engine *eng, int64_t *row_begin, int64_t *n_local);
/* device views (padded layout) for callers that fill scores on the device     */
int stein_engine_buffers(stein_engine *eng, float **X_local_dev, float **S_local_dev,
                         float **phi_local_dev, int64_t *ld, int64_t *rows_padded_local);
/* host <-> device: plain row-major n_local x d arrays (float32 or float64).  On a sharded engine
 * set_particles is COLLECTIVE (every rank replaces its rows): it also tells the engine that the
 * other ranks' copies of these rows are stale, so the next step all-gathers instead of relying on
 * the peer push.                                                                                   */
int stein_engine_set_particles(stein_engine *eng, const void *X_host, int is_f64);
int stein_engine_get_particles(stein_engine *eng, void *X_host, int is_f64);
int stein_engine_set_scores(stein_engine *eng, const void *S_host, int is_f64);
int stein_engine_get_phi(stein_engine *eng, void *phi_host, int is_f64);
/* one update_particles() on the scores currently in the engine's S buffer     */
int stein_engine_step(stein_engine *eng);
/* compute_phi() only (abstract_stein_sampler.py:100-105) on the scores in the S buffer: all-gathers,
 * median / bandwidth, phi into the engine's phi buffer (stein_engine_get_phi), sum(phi^2) all-reduced
 * -- no clip, no optimizer step.  For callers that bring their own AbstractGradientDescent
 * (stein/optimizers/abstract_gradient_descent.py:32-52). */
int stein_engine_phi_only(stein_engine *eng);
/* clip + optimizer step (abstract_stein_sampler.py:125-126) on the phi and sum(phi^2) that are already
 * in the engine's buffers (stein_engine_buffers / stein_engine_sumsq_dev), e.g. written by
 * stein_phi_from_kernel for a user-defined kernel operator.  Single-GPU engines. */
int stein_engine_apply_phi(stein_engine *eng);
int stein_engine_sumsq_dev(stein_engine *eng, double **sumsq_dev);
/* optimizer hyper-parameters for the following steps (the reference reads learning_rate / decay /
 * betas from the gd object at every update(), adam_gradient_descent.py:41-58) */
int stein_engine_set_hyper(stein_engine *eng, double learning_rate, double decay, double p1, double p2);
/* the host-buffer drop-in: H2D scores -> step -> D2H particles (X_host_out may
 * be NULL to leave the particles on the device).  Replaces update_particles(grads_array),
 * stein/samplers/abstract_stein_sampler.py:107-127.  Synchronous contract: the caller holds the
 * updated particles when the call returns.  With X_host_out given the call also ENQUEUES the start
 * of the next iteration (row norms, operand preparation and, in the steady state of the median,
 * the device part of the exact median -- they need only the particles, not the caller's next
 * scores) behind the optimizer kernel, so that this work runs while the particles cross PCIe and
 * while the caller evaluates its scores; the next call (or stein_engine_step / _phi_only) collects
 * the result.  Results are the same bits with and without this prefetch (environment
 * STEIN_PREFETCH=0 or stein_engine_set_prefetch(eng, 0) turn it off).                          */
int stein_engine_update_particles_host(stein_engine *eng, const void *S_host, void *X_host_out,
                                       int is_f64);
int stein_engine_set_prefetch(stein_engine *eng, int on);
/* stein_engine_step / _update_particles_host enqueue phi BEFORE the host has collected the median: in the steady
 * state of the median its final select also runs on the device and leaves the bandwidth in device memory for the
 * kernels of phi; the host verifies (same keys, same bits of h) behind the phi kernel and launches the optimizer
 * only after that -- on a mismatch, or when the median had to take another route, phi is run again with the
 * host's value.  Same results; default on for sharded engines (where the round trip's wake-up jitter is
 * maximised over the ranks by the next all-reduce), off on one GPU; environment STEIN_DEVICE_BW=0 / 1.  _stats: iterations that took
 * this form / that had to repeat phi.                                                                         */
int stein_engine_set_device_bandwidth(stein_engine *eng, int on);
int stein_engine_device_bandwidth_stats(const stein_engine *eng, int64_t *used, int64_t *redone);
/* how many next-iteration medians were enqueued ahead / later collected by a step */
int stein_engine_prefetch_stats(const stein_engine *eng, int64_t *begun, int64_t *used);
/* A caller that writes the particle buffer of stein_engine_buffers itself (instead of
 * stein_engine_set_particles) says so here before its next step: anything derived from the old
 * particles (a prefetched median, the peers' copies of the rows) is dropped.                  */
int stein_engine_particles_changed(stein_engine *eng);
/* Peer push (optional, ranks on one node): each rank exports the CUDA-IPC handle of its particle
 * buffer (STEIN_IPC_HANDLE_BYTES), the handles of all ranks are concatenated in rank order by any
 * means, and every rank imports them.  From then on the optimizer kernel stores the updated rows
 * straight into the other ranks' buffers over NVLink and the all-gather of the particles is replaced
 * by a one-word all-reduce.  Returns STEIN_ERR_UNSUPPORTED when a handle cannot be opened (the
 * engine then keeps using the all-gather hook).  handles == NULL closes the peers again.  Either
 * every rank pushes or none: the caller must agree on the outcome across ranks, and that exchange
 * is also the barrier after which a rank may be written to by its peers.
 * The exported buffer carries a mailbox behind the particles; while the peers are open the small
 * all-reduces of an iteration (histograms, counters, sum(phi^2), the barrier word) run as one
 * kernel each over those mappings instead of the allreduce hooks (environment STEIN_PEER_REDUCE=0,
 * same on every rank, keeps them on the hooks).  Destroying an engine with open peers is
 * collective: every rank must have drained its stream and met the others first. */
#define STEIN_IPC_HANDLE_BYTES 64
int stein_engine_ipc_handle(stein_engine *eng, void *handle_out);
int stein_engine_set_peer_handles(stein_engine *eng, const void *handles /* world x 64 bytes */);
/* Fixed-bandwidth squared-exponential kernel (the plugin point is
 * stein/kernels/abstract_kernel.py:40, where the reference defines `bandwidth` by the median
 * heuristic).  bandwidth > 0: every following step uses exactly this h and skips the median;
 * 0 (the default): the median heuristic of the reference.  Collective state: sharded engines must
 * be given the same value on every rank. */
int stein_engine_set_bandwidth(stein_engine *eng, float bandwidth);
/* diagnostics of the last step (median is NaN for a step with a fixed bandwidth) */
int stein_engine_last(const stein_engine *eng, float *median, float *bandwidth,
                      double *phi_norm, int32_t *sweeps);
/* optimizer state access (checkpoint-lite)               */
int stein_engine_get_state(stein_engine *eng, int64_t *n_iters, double *learning_rate,
                           void *m1_host, void *m2_host, int is_f64);
int stein_engine_set_state(stein_engine *eng, int64_t n_iters, double learning_rate,
                           const void *m1_host, const void *m2_host, int is_f64);

#ifdef __cplusplus
}
#endif
#endif /* STEIN_B200_H */
