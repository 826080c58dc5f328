// Per-layer CUDA-core path (QV_IMPL_LAYERED): one kernel per conv with the reference's
// surrounding glue folded in -- the -128 preprocess (inference/cnn.cu:445-453) into C1's tile
// load, bias add + BLU requantisation + channel concat (inference/cnn.cu:155, mat.cu:262-303,
// cnn.cu:390-391) into every conv's epilogue, and the output requantisation + residual add +
// clamp (inference/cnn.cu:507-523) into C4.  Activations live in HBM as NHWC int8.
// This path is the simple, independently written second CUDA implementation the fused tcgen05
// kernel is cross-checked against at sizes the CPU oracle cannot reach.
#include <cstdint>
#include <cuda_runtime.h>

#include "qv_device.cuh"
#include "qv_layered.h"

namespace qv {

constexpr int TX = 32, TY = 8, NT = TX * TY;

__device__ __forceinline__ int pack4(int b0, int b1, int b2, int b3)
{
    return (b0 & 0xff) | ((b1 & 0xff) << 8) | ((b2 & 0xff) << 16) | ((b3 & 0xff) << 24);
}

// ---- C1: 1 -> 64, 5x5, input u8 luma ---------------------------------------------------
// wpk[r][k][2]: word0 = taps s=0..3 of row r, word1 = tap s=4 (upper lanes 0).
__global__ void __launch_bounds__(NT) k_c1(const uint8_t *__restrict__ x, int8_t *__restrict__ a1,
                                           const int32_t *__restrict__ wpk, const int32_t *__restrict__ bias,
                                           QParam q, int H, int W)
{
    constexpr int TH = TY + 4, TW = TX + 4, PITCH = TW + 4;
    __shared__ int8_t s_x[TH * PITCH];
    __shared__ int2 s_w[5 * 64];
    __shared__ int s_b[64];
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const size_t f = blockIdx.z;
    const uint8_t *xf = x + f * (size_t)H * W;
    for (int i = tid; i < 5 * 64; i += NT) s_w[i] = reinterpret_cast<const int2 *>(wpk)[i];
    if (tid < 64) s_b[tid] = bias[tid];
    for (int i = tid; i < TH * TW; i += NT) {
        const int ty = i / TW, tx = i % TW, gy = y0 + ty - 2, gx = x0 + tx - 2;
        int v = 0;                                                  // zero pad in the x-128 domain
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = (int)xf[(size_t)gy * W + gx] - 128;   // cnn.cu:450
        s_x[ty * PITCH + tx] = (int8_t)v;
    }
    __syncthreads();
    int acc[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) acc[k] = 0;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const int8_t *row = &s_x[(threadIdx.y + r) * PITCH + threadIdx.x];
        const int w0 = pack4(row[0], row[1], row[2], row[3]);
        const int w1 = row[4] & 0xff;
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const int2 wv = s_w[r * 64 + k];
            acc[k] = __dp4a(w0, wv.x, acc[k]);
            acc[k] = __dp4a(w1, wv.y, acc[k]);
        }
    }
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx < W && gy < H) {
        int4 *dst = reinterpret_cast<int4 *>(a1 + ((f * H + gy) * (size_t)W + gx) * 64);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            int o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = v * 16 + j * 4;
                o[j] = pack4(blu_requant(acc[k] + s_b[k], q), blu_requant(acc[k + 1] + s_b[k + 1], q),
                             blu_requant(acc[k + 2] + s_b[k + 2], q), blu_requant(acc[k + 3] + s_b[k + 3], q));
            }
            dst[v] = make_int4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---- dense hidden convs: CIN -> 16 channels per block, KSxKS, NHWC int8 in/out ----------
// wpk[group][tap][c4][16] words (4 input channels per word); output written at channel
// offset out_coff + 16*group of an NHWC tensor with out_ch channels (the concat).
template <int CIN, int KS>
__global__ void __launch_bounds__(NT) k_conv(const int8_t *__restrict__ in, int8_t *__restrict__ out,
                                             const int32_t *__restrict__ wpk, const int32_t *__restrict__ bias,
                                             QParam q, int H, int W, int out_ch, int out_coff, int ngroups)
{
    constexpr int P = (KS - 1) / 2, TH = TY + KS - 1, TW = TX + KS - 1, C4 = CIN / 4;
    extern __shared__ int32_t sm[];
    int32_t *s_in = sm;                       // [C4][TH][TW]
    int32_t *s_w = sm + C4 * TH * TW;         // [KS*KS][C4][16]
    __shared__ int s_b[16];
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int g = blockIdx.z % ngroups;
    const size_t f = blockIdx.z / ngroups;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int32_t *wg = wpk + (size_t)g * KS * KS * C4 * 16;
    for (int i = tid; i < KS * KS * C4 * 16; i += NT) s_w[i] = wg[i];
    if (tid < 16) s_b[tid] = bias[g * 16 + tid];
    const int8_t *inf = in + f * (size_t)H * W * CIN;
    for (int i = tid; i < TH * TW * (CIN / 16); i += NT) {
        const int chunk = i % (CIN / 16), p = i / (CIN / 16);
        const int ty = p / TW, tx = p % TW, gy = y0 + ty - P, gx = x0 + tx - P;
        int4 v = make_int4(0, 0, 0, 0);                              // SAME zero padding, cnn.cu:44-49
        if (gy >= 0 && gy < H && gx >= 0 && gx < W)
            v = *reinterpret_cast<const int4 *>(inf + ((size_t)gy * W + gx) * CIN + chunk * 16);
        int32_t *d = s_in + (chunk * 4) * TH * TW + ty * TW + tx;
        d[0] = v.x; d[TH * TW] = v.y; d[2 * TH * TW] = v.z; d[3 * TH * TW] = v.w;
    }
    __syncthreads();
    int acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0;
    for (int r = 0; r < KS; ++r)
        for (int s = 0; s < KS; ++s) {
            const int32_t *ap = s_in + (threadIdx.y + r) * TW + threadIdx.x + s;
            const int4 *wp = reinterpret_cast<const int4 *>(s_w + (r * KS + s) * C4 * 16);
#pragma unroll 4
            for (int c4 = 0; c4 < C4; ++c4) {
                const int a = ap[c4 * TH * TW];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int4 w = wp[c4 * 4 + v];
                    acc[v * 4 + 0] = __dp4a(a, w.x, acc[v * 4 + 0]);
                    acc[v * 4 + 1] = __dp4a(a, w.y, acc[v * 4 + 1]);
                    acc[v * 4 + 2] = __dp4a(a, w.z, acc[v * 4 + 2]);
                    acc[v * 4 + 3] = __dp4a(a, w.w, acc[v * 4 + 3]);
                }
            }
        }
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx < W && gy < H) {
        int o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o[j] = pack4(blu_requant(acc[4 * j] + s_b[4 * j], q), blu_requant(acc[4 * j + 1] + s_b[4 * j + 1], q),
                         blu_requant(acc[4 * j + 2] + s_b[4 * j + 2], q), blu_requant(acc[4 * j + 3] + s_b[4 * j + 3], q));
        *reinterpret_cast<int4 *>(out + ((f * H + gy) * (size_t)W + gx) * out_ch + out_coff + g * 16) =
            make_int4(o[0], o[1], o[2], o[3]);
    }
}

// ---- C4 48 -> 1, 3x3, fused with applyRes_y ----------------------------------------------
__global__ void __launch_bounds__(NT) k_c4_res(const int8_t *__restrict__ a3, const uint8_t *__restrict__ x,
                                               uint8_t *__restrict__ rec, const int32_t *__restrict__ wpk, int bias,
                                               int mul, int shift, int H, int W)
{
    constexpr int TH = TY + 2, TW = TX + 2, C4 = 12;
    __shared__ int32_t s_in[C4 * TH * TW];
    __shared__ int32_t s_w[9 * C4];
    const int tid = threadIdx.y * TX + threadIdx.x;
    const size_t f = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    if (tid < 9 * C4) s_w[tid] = wpk[tid];
    const int8_t *inf = a3 + f * (size_t)H * W * 48;
    for (int i = tid; i < TH * TW * 3; i += NT) {
        const int chunk = i % 3, p = i / 3;
        const int ty = p / TW, tx = p % TW, gy = y0 + ty - 1, gx = x0 + tx - 1;
        int4 v = make_int4(0, 0, 0, 0);
        if (gy >= 0 && gy < H && gx >= 0 && gx < W)
            v = *reinterpret_cast<const int4 *>(inf + ((size_t)gy * W + gx) * 48 + chunk * 16);
        int32_t *d = s_in + (chunk * 4) * TH * TW + ty * TW + tx;
        d[0] = v.x; d[TH * TW] = v.y; d[2 * TH * TW] = v.z; d[3 * TH * TW] = v.w;
    }
    __syncthreads();
    int acc = 0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int32_t *ap = s_in + (threadIdx.y + t / 3) * TW + threadIdx.x + t % 3;
#pragma unroll
        for (int c4 = 0; c4 < C4; ++c4) acc = __dp4a(ap[c4 * TH * TW], s_w[t * C4 + c4], acc);
    }
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx < W && gy < H) {
        const size_t i = (f * H + gy) * (size_t)W + gx;
        rec[i] = (uint8_t)residual_apply(acc + bias, (int)x[i], mul, shift);
    }
}

// ---- NHWC int8 -> planar [C][H][W] (debug taps only) --------------------------------------
__global__ void k_nhwc_to_planar(const int8_t *__restrict__ in, int8_t *__restrict__ out, int C, int HW)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)C * HW) return;
    const int c = (int)(i / HW);
    const size_t p = i % HW;
    out[i] = in[p * C + c];
}

// ---- exact int64 SSE (PSNR core, inference/yuv_data.cpp:92-93) -----------------------------
__global__ void __launch_bounds__(256) k_sse(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, size_t n,
                                             unsigned long long *__restrict__ accum)
{
    unsigned long long s = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n16 = n / 16;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    size_t done = 0;
    if (aligned) {
        const uint4 *a4 = reinterpret_cast<const uint4 *>(a), *b4 = reinterpret_cast<const uint4 *>(b);
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 va = a4[i], vb = b4[i];
            const unsigned wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
            unsigned t = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int d = (int)((wa[j] >> (8 * k)) & 0xff) - (int)((wb[j] >> (8 * k)) & 0xff);
                    t += (unsigned)(d * d);
                }
            s += t;
        }
        done = n16 * 16;
    }
    for (size_t i = done + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const int d = (int)a[i] - (int)b[i];
        s += (unsigned)(d * d);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ unsigned long long ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += ws[w];
        atomicAdd(accum, t);
    }
}

// ---- host launchers ---------------------------------------------------------------------
template <int CIN, int KS>
static cudaError_t launch_conv(const int8_t *in, int8_t *out, const LayeredLayer &L, int n, int H, int W, int out_ch,
                               int out_coff, cudaStream_t st)
{
    constexpr int TH = TY + KS - 1, TW = TX + KS - 1, C4 = CIN / 4;
    const size_t smem = sizeof(int32_t) * (size_t)(C4 * TH * TW + KS * KS * C4 * 16);
    // function attributes are per device: set it on every launch (a few hundred ns) so that handles on
    // several GPUs of one process all get it
    cudaError_t e = cudaFuncSetAttribute(k_conv<CIN, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int ngroups = L.cout / 16;
    dim3 grid((W + TX - 1) / TX, (H + TY - 1) / TY, n * ngroups), block(TX, TY);
    k_conv<CIN, KS><<<grid, block, smem, st>>>(in, out, L.d_wpk, L.d_bias, L.q, H, W, out_ch, out_coff, ngroups);
    return cudaGetLastError();
}

cudaError_t layered_forward(const LayeredModel &M, const uint8_t *d_in, uint8_t *d_out, int n, int H, int W,
                            int8_t *a1, int8_t *a2, int8_t *a3, cudaStream_t st, long long *launches)
{
    cudaError_t e;
    dim3 block(TX, TY), grid((W + TX - 1) / TX, (H + TY - 1) / TY, n);
    k_c1<<<grid, block, 0, st>>>(d_in, a1, M.L[QV_C1].d_wpk, M.L[QV_C1].d_bias, M.L[QV_C1].q, H, W);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = launch_conv<64, 3>(a1, a2, M.L[QV_C2_1], n, H, W, 48, 0, st)) != cudaSuccess) return e;
    if ((e = launch_conv<64, 5>(a1, a2, M.L[QV_C2_2], n, H, W, 48, 32, st)) != cudaSuccess) return e;
    if ((e = launch_conv<48, 3>(a2, a3, M.L[QV_C3_1], n, H, W, 48, 0, st)) != cudaSuccess) return e;
    if ((e = launch_conv<48, 1>(a2, a3, M.L[QV_C3_2], n, H, W, 48, 16, st)) != cudaSuccess) return e;
    k_c4_res<<<grid, block, 0, st>>>(a3, d_in, d_out, M.L[QV_C4].d_wpk, M.c4_bias, M.L[QV_C4].q.mul,
                                     M.L[QV_C4].q.shift, H, W);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) *launches += 6;
    return cudaSuccess;
}

cudaError_t nhwc_to_planar(const int8_t *d_in, int8_t *d_out, int C, size_t HW, cudaStream_t st)
{
    const size_t n = (size_t)C * HW;
    k_nhwc_to_planar<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, d_out, C, (int)HW);
    return cudaGetLastError();
}

cudaError_t sse_accumulate(const uint8_t *a, const uint8_t *b, size_t n, int64_t *d_accum, cudaStream_t st)
{
    int blocks = (int)((n / 16 + 255) / 256);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_sse<<<blocks, 256, 0, st>>>(a, b, n, reinterpret_cast<unsigned long long *>(d_accum));
    return cudaGetLastError();
}

}  // namespace qv
