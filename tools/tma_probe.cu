// tma_probe -- which forms of a 2-D byte-tensor TMA load does sm_100a accept?  (One variant per process: a fault kills the context.)
// usage: tma_probe <variant> <c0>
//   variant 0: CUtensorMap as a top-level __grid_constant__ parameter      1: as the first member of a __grid_constant__ struct
//   c0      : first column of the 144-byte box (may be negative / not a multiple of 16)
// Not part of the product.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

struct alignas(64) Params { CUtensorMap map; int c0, c1, pad[14]; uint8_t *out; };

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ void body(const CUtensorMap *map, int c0, int c1, uint8_t *out)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 1024);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = 0xEE;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(144u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(sm)), "l"(map), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
    }
    uint32_t ok = 0;
    for (int n = 0; n < 2000000 && !ok; ++n)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < 160; i += blockDim.x) out[i] = sm[i];
    if (threadIdx.x == 0) out[160] = (uint8_t)ok;
}
__global__ void k_direct(const __grid_constant__ CUtensorMap map, int c0, int c1, uint8_t *out) { body(&map, c0, c1, out); }
__global__ void k_member(const __grid_constant__ Params P) { body(&P.map, P.c0, P.c1, P.out); }

int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0, c0 = argc > 2 ? atoi(argv[2]) : 0;
    const int W = 416, H = 240;
    std::vector<uint8_t> h((size_t)W * H);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) h[(size_t)y * W + x] = (uint8_t)(1 + (x + 3 * y) % 250);
    uint8_t *d_in, *d_out;
    CK(cudaMalloc(&d_in, h.size())); CK(cudaMalloc(&d_out, 256));
    CK(cudaMemcpy(d_in, h.data(), h.size(), cudaMemcpyHostToDevice));
    void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    Params P{};
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W};
    const cuuint32_t box[2] = {144, 1}, estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fn)(&P.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d c0 %d: encode rc=%d\n", variant, c0, (int)r);
    if (r != CUDA_SUCCESS) return 1;
    const int c1 = 7;
    P.c0 = c0; P.c1 = c1; P.out = d_out;
    if (variant == 0) k_direct<<<1, 128, 2048>>>(P.map, c0, c1, d_out); else k_member<<<1, 128, 2048>>>(P);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d c0 %d: kernel: %s\n", variant, c0, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    uint8_t o[161];
    CK(cudaMemcpy(o, d_out, 161, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < 144; ++i) { const int x = c0 + i; const uint8_t want = (x >= 0 && x < W) ? h[(size_t)c1 * W + x] : 0; bad += o[i] != want; }
    printf("variant %d c0 %d: barrier completed=%d, %d of 144 bytes wrong, bytes 144..147 (untouched = 0xEE): %02x %02x %02x %02x, first bytes: %02x %02x %02x %02x %02x %02x %02x %02x %02x %02x\n",
           variant, c0, o[160], bad, o[144], o[145], o[146], o[147], o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9]);
    return 0;
}
