// Host-visible interface of the fused tcgen05 path (qv_fused.cu).
#pragma once
#include <cstdint>
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

}  // namespace qv
