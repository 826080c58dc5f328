// Host-visible interface of the per-layer CUDA-core path (qv_layered.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "qv_device.cuh"
#include "qv_internal.h"

namespace qv {

struct LayeredLayer {
    int32_t *d_wpk = nullptr;   // packed weights (layout documented at each kernel)
    int32_t *d_bias = nullptr;  // int32 bias[cout]
    QParam q{};
    int cout = 0;
};

struct LayeredModel {
    LayeredLayer L[QV_NLAYER];
    int c4_bias = 0;
};

// Whole net, n frames of HxW luma resident in device memory; a1/a2/a3 are NHWC int8 scratch
// tensors of n*H*W*{64,48,48} bytes.
cudaError_t layered_forward(const LayeredModel &M, const uint8_t *d_in, uint8_t *d_out, int n, int H, int W,
                            int8_t *a1, int8_t *a2, int8_t *a3, cudaStream_t st, long long *launches);
cudaError_t nhwc_to_planar(const int8_t *d_in, int8_t *d_out, int C, size_t HW, cudaStream_t st);
cudaError_t sse_accumulate(const uint8_t *a, const uint8_t *b, size_t n, int64_t *d_accum, cudaStream_t st);

}  // namespace qv
