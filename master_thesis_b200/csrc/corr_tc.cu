// corr_tc.cu - tcgen05 path of K2 (placeholder until the tensor-core kernel lands:
// reports "unsupported" so that mt_corr4d_fwd serves every shape with the SIMT path).
#include "mt_common.cuh"

namespace mt {
int corr4d_tc_supported(int, int) { return 0; }
int64_t corr4d_tc_workspace_bytes(int, int, int, int) { return 0; }
int corr4d_tc_launch(const float *, const float *, const float *, const float *, float *, void *,
                     int64_t, int, int, int, int, cudaStream_t) {
    set_error("corr4d tcgen05 path not built");
    return MT_ERR_INVALID;
}
}  // namespace mt
