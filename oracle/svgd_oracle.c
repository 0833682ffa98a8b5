/*
 * oracle/svgd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the arithmetic on the reference's SVGD hot path
 * (JamesBrofos/Stein).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object.
 * The product (stein_b200/) never links, imports or calls it.
 *
 * Pinning: the reference ships no tests and no golden vectors, and TensorFlow
 * 1.12 cannot be installed here.  The formulas below are checked (through
 * oracle/svgd_oracle.py, tests/test_reference_run.py) against the reference's own
 * library code executed on the TF1 graph-API stand-in of compat/
 * (tests/golden/make_golden_reference_run.py) to float32 rounding.  PARITY UNPINNED
 * at the bit level: TensorFlow's own SGEMM / exp kernels cannot be run, so the bits
 * of D -- and with them "bit-exact median" -- are defined by the contract
 * arithmetic stated below.  The restated algorithm, reference line by line:
 *
 *   D   = r + r^T - 2 T T^T            stein/kernels/abstract_kernel.py:33-35
 *   med = median of all n*n entries    stein/utilities/compute_median.py:4-16
 *   h   = sqrt(med / ln n)             stein/kernels/abstract_kernel.py:40
 *   K   = exp(-D / h^2 / 2)            stein/kernels/squared_exponential_kernel.py:22
 *   dK  = -0.5 * d(sum K)/d theta      stein/kernels/squared_exponential_kernel.py:23,32
 *   phi = (K.dot(S) + dK) / n          stein/samplers/abstract_stein_sampler.py:105
 *
 * "Contract arithmetic" for D (what makes the median bit-reproducible between
 * this oracle and the CUDA path): fp32 throughout,
 *     g_ij = fma-chain over k = 0..d-1 starting from +0  (acc = fmaf(x_ik, x_jk, acc))
 *     r_i  = g_ii
 *     D_ij = fl( fl(r_i + r_j) - 2*g_ij )
 * TensorFlow/Eigen's own SGEMM accumulation order is not specified (and differs
 * per build), so a fixed order has to be chosen by whoever wants reproducible
 * bits; this is that choice.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- helpers ------------------------------------------------------------ */

static inline uint32_t f2key(float f) {
    /* order-preserving map fp32 -> u32 (ascending); -0 is folded onto +0 */
    f = f + 0.0f;
    uint32_t b;
    memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static inline float key2f(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* fma-chain dot product in the contract order */
static inline float chain_dot(const float *a, const float *b, int64_t d) {
    float acc = 0.0f;
    for (int64_t k = 0; k < d; ++k) acc = fmaf(a[k], b[k], acc);
    return acc;
}

/* r_i = g_ii   (abstract_kernel.py:34, contract order) */
void oracle_row_norms(const float *X, int64_t n, int64_t d, float *r) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) r[i] = chain_dot(X + i * d, X + i * d, d);
}

/* XT = X^T (d x n) so that the j index is contiguous */
static float *transpose(const float *X, int64_t n, int64_t d) {
    float *XT = (float *)malloc(sizeof(float) * (size_t)n * (size_t)d);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < d; ++k)
        for (int64_t j = 0; j < n; ++j) XT[k * n + j] = X[j * d + k];
    return XT;
}

/* g[j] = chain_dot(x_i, x_j) for all j.  j is the vector dimension (so the
 * compiler can use SIMD fma); the k order of every chain is the contract order. */
#define JB 64
static void gram_row(const float *xi, const float *XT, int64_t n, int64_t d, float *g) {
    for (int64_t j0 = 0; j0 < n; j0 += JB) {
        int64_t w = (n - j0 < JB) ? n - j0 : JB;
        float acc[JB];
        for (int jj = 0; jj < JB; ++jj) acc[jj] = 0.0f;
        if (w == JB) {
            for (int64_t k = 0; k < d; ++k) {
                const float a = xi[k];
                const float *row = XT + k * n + j0;
                for (int jj = 0; jj < JB; ++jj) acc[jj] = fmaf(a, row[jj], acc[jj]);
            }
        } else {
            for (int64_t k = 0; k < d; ++k) {
                const float a = xi[k];
                const float *row = XT + k * n + j0;
                for (int jj = 0; jj < w; ++jj) acc[jj] = fmaf(a, row[jj], acc[jj]);
            }
        }
        for (int jj = 0; jj < w; ++jj) g[j0 + jj] = acc[jj];
    }
}

/* one block of rows of D in contract arithmetic: rows [i0,i1) x all n columns */
static void sqdist_rows(const float *X, const float *XT, const float *r, int64_t n, int64_t d,
                        int64_t i0, int64_t i1, float *D /* (i1-i0) x n */) {
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; ++i) {
        float *Di = D + (i - i0) * n;
        gram_row(X + i * d, XT, n, d, Di);
        for (int64_t j = 0; j < n; ++j) {
            float t = r[i] + r[j];
            Di[j] = (t - 2.0f * Di[j]) + 0.0f;
        }
    }
}

/* Full D (abstract_kernel.py:35) in contract arithmetic; D is n x n row-major. */
void oracle_sqdist_chain(const float *X, int64_t n, int64_t d, float *D) {
    float *r = (float *)malloc(sizeof(float) * (size_t)n);
    float *XT = transpose(X, n, d);
    oracle_row_norms(X, n, d, r);
    sqdist_rows(X, XT, r, n, d, 0, n, D);
    free(XT);
    free(r);
}

/* ---- exact median of all n*n entries (compute_median.py:4-16) ----------- */

/* Radix select over a histogram pass: counts keys of D whose top bits match
 * `prefix` (under `mask`) into 2^bits bins of the next digit.  D is recomputed
 * block by block so nothing n x n is ever stored (same idea as the CUDA path,
 * different code). */
#define OR_BITS 11
static void hist_pass(const float *X, const float *XT, const float *r, int64_t n, int64_t d,
                      uint32_t prefix, uint32_t mask, int shift, int bits, uint64_t *bins) {
    const int nb = 1 << bits;
    int nt = oracle_num_threads();
    uint64_t *tb = (uint64_t *)calloc((size_t)nt * nb, sizeof(uint64_t));
#pragma omp parallel
    {
#ifdef _OPENMP
        uint64_t *my = tb + (size_t)omp_get_thread_num() * nb;
#else
        uint64_t *my = tb;
#endif
        float *g = (float *)malloc(sizeof(float) * (size_t)n);
#pragma omp for schedule(dynamic, 8)
        for (int64_t i = 0; i < n; ++i) {
            gram_row(X + i * d, XT, n, d, g);
            for (int64_t j = 0; j < n; ++j) {
                float t = r[i] + r[j];
                uint32_t key = f2key(t - 2.0f * g[j]);
                if ((key & mask) == prefix) my[(key >> shift) & (nb - 1)]++;
            }
        }
        free(g);
    }
    for (int b = 0; b < nb; ++b) {
        uint64_t s = 0;
        for (int t = 0; t < nt; ++t) s += tb[(size_t)t * nb + b];
        bins[b] = s;
    }
    free(tb);
}

/* value at 0-based ascending rank `rank` among the n*n entries of D */
static float select_rank(const float *X, const float *XT, const float *r, int64_t n, int64_t d,
                         uint64_t rank) {
    const int nb = 1 << OR_BITS;
    uint64_t *bins = (uint64_t *)malloc(sizeof(uint64_t) * nb);
    uint32_t prefix = 0, mask = 0;
    int consumed = 0;
    while (consumed < 32) {
        int bits = (32 - consumed < OR_BITS) ? 32 - consumed : OR_BITS;
        int shift = 32 - consumed - bits;
        hist_pass(X, XT, r, n, d, prefix, mask, shift, bits, bins);
        uint64_t acc = 0;
        int b = 0, nbins = 1 << bits;
        for (; b < nbins; ++b) {
            if (acc + bins[b] > rank) break;
            acc += bins[b];
        }
        rank -= acc;
        prefix |= ((uint32_t)b) << shift;
        mask |= ((uint32_t)(nbins - 1)) << shift;
        consumed += bits;
    }
    free(bins);
    return key2f(prefix);
}

/* k-th smallest (0-based) of v[0..m) by Hoare quickselect; permutes v */
static float quickselect(float *v, uint64_t m, uint64_t k) {
    uint64_t lo = 0, hi = m - 1;
    while (lo < hi) {
        float p = v[lo + (hi - lo) / 2];
        uint64_t i = lo, j = hi;
        while (i <= j) {
            while (v[i] < p) ++i;
            while (v[j] > p) --j;
            if (i <= j) {
                float t = v[i]; v[i] = v[j]; v[j] = t;
                ++i;
                if (j == 0) break;
                --j;
            }
        }
        if (k <= j) hi = j;
        else if (k >= i) lo = i;
        else break;
    }
    return v[k];
}

static float median_of_two(uint64_t dim, float lo, float hi) {
    /* compute_median.py:13 -- tf.reduce_mean of two fp32 values: fp32 sum, / 2 */
    return (dim % 2 == 0) ? (lo + hi) / 2.0f : lo;
}

/* compute_median.py:9-15: top_k(V, dim//2+1); even -> mean of the last two
 * (= the two middle values), odd -> the middle value.  `mid[0..1]` receive the
 * middle value(s) (equal when n*n is odd).  Returns the fp32 median.
 * This variant never stores n x n: radix select with D recomputed per pass. */
float oracle_median_chain_radix(const float *X, int64_t n, int64_t d, float *mid) {
    float *r = (float *)malloc(sizeof(float) * (size_t)n);
    float *XT = transpose(X, n, d);
    oracle_row_norms(X, n, d, r);
    uint64_t dim = (uint64_t)n * (uint64_t)n;
    float lo, hi;
    if (dim % 2 == 0) {
        lo = select_rank(X, XT, r, n, d, dim / 2 - 1);
        hi = select_rank(X, XT, r, n, d, dim / 2);
    } else {
        lo = hi = select_rank(X, XT, r, n, d, dim / 2);
    }
    free(XT);
    free(r);
    if (mid) { mid[0] = lo; mid[1] = hi; }
    return median_of_two(dim, lo, hi);
}

/* Same result by the literal route: materialise D (abstract_kernel.py:35),
 * flatten, pick the middle order statistic(s).  n*n*4 bytes of memory. */
float oracle_median_chain(const float *X, int64_t n, int64_t d, float *mid) {
    uint64_t dim = (uint64_t)n * (uint64_t)n;
    float *D = (float *)malloc(sizeof(float) * (size_t)dim);
    if (!D) return oracle_median_chain_radix(X, n, d, mid);
    oracle_sqdist_chain(X, n, d, D);
    float lo, hi;
    if (dim % 2 == 0) {
        hi = quickselect(D, dim, dim / 2);
        /* after the partition every element left of dim/2 is <= hi: the lower
         * middle is their maximum */
        lo = D[0];
        for (uint64_t t = 1; t < dim / 2; ++t) if (D[t] > lo) lo = D[t];
    } else {
        lo = hi = quickselect(D, dim, dim / 2);
    }
    free(D);
    if (mid) { mid[0] = lo; mid[1] = hi; }
    return median_of_two(dim, lo, hi);
}

/* abstract_kernel.py:40 -- np.log(n) is a Python double, folded by TF into an
 * fp32 constant because `m` is an fp32 tensor; sqrt in fp32. */
float oracle_bandwidth(float med, int64_t n) {
    float ln_n = (float)log((double)n);
    return sqrtf(med / ln_n);
}

/* ---- phi for a block of rows (abstract_stein_sampler.py:100-105) ---------
 * fp32 D (contract arithmetic), fp32 K = expf(-D / h^2 / 2) with h^2 =
 * square(bandwidth) (squared_exponential_kernel.py:22), fp32 dK, and the final
 * (K.dot(S) + dK) / n in float64 exactly as the NumPy line does.
 * X, S: n x d fp32 (S holds fp32 score values).  phi: (i1-i0) x d float64.
 * dK_i = (x_i * sum_j K_ij - sum_j K_ij x_j) / h^2   (the -0.5 * tf.gradients
 * result).  ksum_out (optional): sum_j K_ij per row. */
void oracle_phi_rows(const float *X, const float *S, int64_t n, int64_t d, float bandwidth,
                     int64_t i0, int64_t i1, double *phi, float *ksum_out) {
    float *r = (float *)malloc(sizeof(float) * (size_t)n);
    float *XT = transpose(X, n, d);
    oracle_row_norms(X, n, d, r);
    const float h2 = bandwidth * bandwidth;
#pragma omp parallel
    {
        double *ks = (double *)malloc(sizeof(double) * (size_t)d); /* sum_j K_ij S_j (f64) */
        float *kx = (float *)malloc(sizeof(float) * (size_t)d);    /* sum_j K_ij x_j (fp32) */
        float *g = (float *)malloc(sizeof(float) * (size_t)n);
#pragma omp for schedule(dynamic, 4)
        for (int64_t i = i0; i < i1; ++i) {
            const float *xi = X + i * d;
            gram_row(xi, XT, n, d, g);
            for (int64_t c = 0; c < d; ++c) { ks[c] = 0.0; kx[c] = 0.0f; }
            float ksum = 0.0f;
            for (int64_t j = 0; j < n; ++j) {
                const float *xj = X + j * d;
                float dij = (r[i] + r[j]) - 2.0f * g[j];
                float kij = expf(-dij / h2 / 2.0f);
                ksum += kij;
                const float *sj = S + j * d;
                for (int64_t c = 0; c < d; ++c) {
                    ks[c] += (double)kij * (double)sj[c];
                    kx[c] += kij * xj[c];
                }
            }
            double *out = phi + (i - i0) * d;
            for (int64_t c = 0; c < d; ++c) {
                float dk = (xi[c] * ksum - kx[c]) / h2;
                out[c] = (ks[c] + (double)dk) / (double)n;
            }
            if (ksum_out) ksum_out[i - i0] = ksum;
        }
        free(ks);
        free(kx);
        free(g);
    }
    free(XT);
    free(r);
}
