// phi_tc.cu -- kernel (3), tcgen05/TMEM/TMA flash path (placeholder until the
// tensor-core kernel lands; AUTO dispatch uses the dense FFMA path meanwhile).
#include "phi_common.cuh"

namespace stein {

bool flash_tc_supported(const stein_ctx *, int64_t, int64_t, int64_t) { return false; }
int64_t flash_tc_workspace_bytes(const stein_ctx *, int64_t, int64_t, int64_t) { return 0; }
int phi_flash_tc(stein_ctx *ctx, const float *, const float *, const float *, int64_t, int64_t, int64_t,
                 int64_t, int64_t, float, void *, int64_t, float *, double *) {
    return fail(ctx, STEIN_ERR_UNSUPPORTED, "flash tcgen05 phi kernel not built");
}

}  // namespace stein
