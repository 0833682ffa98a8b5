// phi_common.cuh -- declarations shared by the phi paths (dense FFMA, tcgen05 flash).
#pragma once
#include "common.cuh"

namespace stein {

constexpr int FINALIZE_MAX_BLOCKS = 1184;  // 8 x 148 SMs

int launch_make_y(stein_ctx *ctx, const float *X, const float *S, int64_t rows, int64_t ld, float h2,
                  float *Y);
// phi = (sum over nslots partial O buffers + x * (sum of ksum slots) / h2) / n_total,
// sum(phi^2) -> *sumsq (double, device) through `partials` (FINALIZE_MAX_BLOCKS doubles)
int launch_finalize(stein_ctx *ctx, const float *O, int64_t slot_stride, int nslots, const float *ksum,
                    int64_t ksum_slot_stride, const float *X_local, int64_t rows_valid, int64_t rows,
                    int64_t ld, float h2, int64_t n_total, float *phi, double *partials,
                    double *sumsq);

int64_t dense_workspace_bytes(int64_t n_local, int64_t n_total, int64_t d);
int phi_dense(stein_ctx *ctx, const float *X_all, const float *S_all, const float *r_all,
              int64_t n_total, int64_t d, int64_t ld, int64_t row_begin, int64_t n_local, float h2,
              void *ws, int64_t ws_bytes, float *phi, double *sumsq);

// tcgen05 flash path (phi_tc.cu)
bool flash_tc_supported(const stein_ctx *ctx, int64_t n_local, int64_t n_total, int64_t d);
int64_t flash_tc_workspace_bytes(const stein_ctx *ctx, int64_t n_local, int64_t n_total, int64_t d);
// guarded: conditioning guard on (STEIN_PHI_AUTO): clouds beyond the range of the kernel's three BF16
// passes go to phi_dense; d_true = number of real coordinates
int phi_flash_tc(stein_ctx *ctx, const float *X_all, const float *S_all, const float *r_all,
                 int64_t n_total, int64_t d, int64_t ld, int64_t row_begin, int64_t n_local, float h2,
                 void *ws, int64_t ws_bytes, float *phi, double *sumsq, bool guarded, int64_t d_true);

// CTA-pair (cta_group::2) variant, d padded to 256 only
bool flash_tc2_supported(const stein_ctx *ctx, int64_t n_local, int64_t n_total, int64_t d);
// mode 0: BF16x3; 1: mixed-precision GEMM2; 2: mixed precision for both GEMMs (fast); 3: FP16x3 for both
// (precise); 4: fast or precise, picked on the device from the conditioning of the cloud.  d_true = number
// of real coordinates (d may have been promoted to the leading dimension).
int phi_flash_tc2(stein_ctx *ctx, const float *X_all, const float *S_all, const float *r_all,
                  int64_t n_total, int64_t d, int64_t d_true, int64_t ld, int64_t row_begin, int64_t n_local,
                  float h2, void *ws, int64_t ws_bytes, float *phi, double *sumsq, int mode);
// the bandwidth-independent kernels of phi_flash_tc2, enqueued ahead of the call (ctx->xprep)
int flash_tc2_prepare_x(stein_ctx *ctx, const float *X_all, int64_t n_total, int64_t d, int64_t ld,
                        int64_t n_local, void *ws, int64_t ws_bytes, int mode);
int flash_tc2_prepare_s(stein_ctx *ctx, const float *X_all, const float *S_all, int64_t n_total, int64_t d, int64_t ld,
                        int64_t n_local, void *ws, int64_t ws_bytes, int mode);
// panel kernels (phi_panel.cuh): leading dimension 512 / 768 / 1024.  mode 2 fast, 3 precise, 4 guarded
// returned by a phi call that was asked to take the bandwidth from the device (ctx->dev_bw) but cannot
constexpr int PHI_DEV_BW_NA = 5;
namespace panel {
bool panel_supported(const stein_ctx *ctx, int64_t n_total, int64_t ld);
int64_t panel_workspace_bytes(const stein_ctx *ctx, int64_t n_local, int64_t n_total, int64_t ld);
int phi_panel(stein_ctx *ctx, const float *X_all, const float *S_all, const float *r_all, int64_t n_total, int64_t d,
              int64_t d_true, int64_t ld, int64_t row_begin, int64_t n_local, float h2, void *ws, int64_t ws_bytes,
              float *phi, double *sumsq, int mode);
}  // namespace panel
// what stein_phi would run for this problem: prepares its X side if that kernel supports it (ctx.cu)
int phi_prepare_x(stein_ctx *ctx, const float *X_all, int64_t n_total, int64_t d, int64_t ld, int64_t n_local,
                  void *ws, int64_t ws_bytes);
int phi_prepare_s(stein_ctx *ctx, const float *X_all, const float *S_all, int64_t n_total, int64_t d, int64_t ld,
                  int64_t n_local, void *ws, int64_t ws_bytes);

}  // namespace stein
