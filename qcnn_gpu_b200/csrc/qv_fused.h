// Host-visible interface of the fused tcgen05 path (qv_fused.cu).
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "qv_internal.h"

namespace qv {

struct FusedModel;   // device-resident weights / descriptors in the layout the fused kernel wants

// Builds the device image of the model for the fused kernel. Returns nullptr + error on failure.
FusedModel *fused_upload(const ModelHost &m, cudaStream_t st);
void fused_free(FusedModel *fm);
// Whole net on n frames of HxW luma resident in device memory, one launch.
cudaError_t fused_forward(const FusedModel *fm, const uint8_t *d_in, uint8_t *d_out, int n, int H, int W,
                          cudaStream_t st, long long *launches);
// After the stream the kernel ran on has been synchronised: 0, or what a CTA reported (1 = an mbarrier wait timed
// out, 2 = shared-memory / TMEM bases other than the operand table was built for); the report is cleared.
int fused_take_failure(const FusedModel *fm);

// Test hook (no GPU): weight image, per-phase MMA operand table (bases 0) and constants as the host builds them.
void fused_debug_tables(const ModelHost &m, std::vector<uint8_t> &wimg, std::vector<uint32_t> &ops, std::vector<int32_t> &consts);

}  // namespace qv
