// Device-side helpers shared by the layered and the fused kernels: the two requantisation
// formulas of the reference, bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qv {

// Static per-layer scale triple plus the derived rounding bias of inference/mat.cu:268.
struct QParam {
    int blu;      // BLU bound in accumulator units (compared BEFORE scaling, strict >)
    int mul;      // multiplier
    int shift;    // right shift
    int rbias;    // (1 << (shift-1)) / mul      (hidden layers, mat.cu:268)
};

// Hidden-layer epilogue: CHW2CHW_VECT_C_QUANT_BLU, inference/mat.cu:286-291.
// `u` = int32 accumulator + int32 bias (equal to the reference's fp32 u inside the exact-integer
// envelope that qv_load_static_para enforces).  The store to `xwtype` (char) keeps the low byte.
__device__ __forceinline__ int blu_requant(int u, const QParam &q)
{
    if (u > q.blu) return 127;
    if (u < 0) return 0;
    return (int)(int8_t)((int)((unsigned)(u + q.rbias) * (unsigned)q.mul) >> q.shift);
}

// Output-layer requantisation + residual add + clamp: applyRes_GPU_y, inference/cnn.cu:507-523.
// bias = 1 << (shift-1) is added AFTER the multiply (cnn.cu:512,516); the sum with x goes through
// a `short` (cnn.cu:510,517) before the [0,255] clamp.
__device__ __forceinline__ int residual_apply(int u4, int x, int mul, int shift)
{
    int t = (int)((unsigned)u4 * (unsigned)mul + (1u << (shift - 1))) >> shift;
    int r = (int)(short)(x + t);
    return r > 255 ? 255 : (r < 0 ? 0 : r);
}

}  // namespace qv
