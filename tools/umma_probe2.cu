// umma_probe2 -- does tcgen05.mma.cta_group::2 (CTA pair, M = 2 x 128) work with the fused kernel's
// operand layouts, what does an instruction cost, and can cta_group::1 MMAs be mixed in on the same TMEM?
// Not part of the product.   usage: umma_probe2 num | thr [flags] | thrb | lat
//   thr flags (sum): 2 a different B tile for every MMA (= thrb), 4 six accumulator regions in rotation, 8 the other
//   warps of both CTAs drain TMEM (tcgen05.ld) meanwhile, 16 the other warps hammer shared memory (ld.shared) meanwhile
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../qcnn_gpu_b200/csrc/qv_tcgen05.cuh"

using namespace qv::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int PX = 192, PLANE_B = PX * 16;          // A: [2 planes][192 px][16 B]
constexpr int A_BYTES = 2 * PLANE_B;
constexpr int NMAX = 256;
constexpr int B_BYTES = 2 * NMAX * 16;              // per CTA: [2 K-chunks][N/2 rows][16 B]
constexpr int NBT = 8;                              // thrb: B tiles to rotate through

__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst, uint32_t n)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(n) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t t, uint32_t n)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(n) : "memory");
}
__device__ __forceinline__ void mma2_i8_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit2(uint64_t *bar, uint32_t mask)
{
    asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\t"
                 "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], lo;\n\t}" ::"r"(
                     smem_u32(bar)),
                 "r"(mask)
                 : "memory");
}

// mode 0: numerics (one MMA, optional second cta_group::1 MMA on columns 128.. of each CTA)
// mode 1: throughput (64 x 32 MMAs)
template <int N>
__global__ void __cluster_dims__(2, 1, 1) k2(const int8_t *gA, const int8_t *gB, int *out, int mode, int shift, long long *cycles, int *status, int flags)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm, *sB = sm + A_BYTES;
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cta_rank();
    // A: each CTA its own 128 rows (CTA r uses global rows r*PX..); B: CTA r holds rows [r*N/2, (r+1)*N/2)
    for (int i = tid; i < A_BYTES / 16; i += blockDim.x) {
        const int plane = i / PX, px = i % PX;
        reinterpret_cast<int4 *>(sA)[i] = reinterpret_cast<const int4 *>(gA)[(rank * 2 + plane) * PX + px];
    }
    for (int t = 0; t < ((flags & 2) ? NBT : 1); ++t)
        for (int i = tid; i < N; i += blockDim.x) {      // i = kc * (N/2) + n
            const int kc = i / (N / 2), n = i % (N / 2);
            reinterpret_cast<int4 *>(sB + t * N * 16)[kc * (N / 2) + n] = reinterpret_cast<const int4 *>(gB)[(rank * (N / 2) + n) * 2 + kc];
        }
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); s_stop = 0; }
    if (warp == 0) tmem_alloc2(&s_tmem, 512);
    fence_before_sync();
    __syncthreads();
    cluster_sync();
    fence_after_sync();
    const uint32_t tm = s_tmem;
    long long t0 = 0, t1 = 0;
    if (rank == 0 && warp == 0) {
        const bool leader = elect_one();
        const uint64_t ad = smem_desc(smem_u32(sA) + shift * 16, PLANE_B, 128);
        const uint64_t bd = smem_desc(smem_u32(sB), (N / 2) * 16, 128);
        const uint32_t id = idesc_i8(256, N, 1, 1);
        t0 = clock64();
        if (mode == 0) {
            if (leader) mma2_i8_ss(tm, ad, bd, id, 0);
        } else {
            for (int o = 0; o < 64; ++o)
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (leader) mma2_i8_ss(tm + ((flags & 4) ? (uint32_t)((j % 6) * 64) : 0u), ad + (uint64_t)((j * 5) % 60),
                                           bd + (uint64_t)((flags & 2) ? ((j * 3) % NBT) * N : 0), id, (o | (j / 6)) != 0);
        }
        if (leader) commit2(&bar, 3);
        __syncwarp();
    }
    long long side = 0;
    if (mode == 1 && warp != 0 && (flags & 24)) {
        uint32_t r[16], acc = 0;
        while (!s_stop) {
            if (flags & 8) { tmem_ld_x16(tm + ((uint32_t)(warp * 32) << 16) + 384 + (side & 3) * 16, r); tmem_ld_wait(); acc += r[0]; }
            if (flags & 16) acc += reinterpret_cast<volatile uint32_t *>(sA)[(tid * 4 + side * 128) & (A_BYTES / 4 - 1)];
            ++side;
            if (mbar_try_wait(&bar, 0)) break;
        }
        if (acc == 0x12345678u) out[0] = (int)acc;
    }
    const bool ok = mbar_wait(&bar, 0);
    t1 = clock64();
    if (tid == 0) s_stop = 1;
    fence_after_sync();
    if (!ok && tid == 0) atomicOr(status, 1 << rank);
    if (rank == 0 && tid == 0) { cycles[0] = t1 - t0; }
    if (mode == 0 && ok) {
        // optional: a cta_group::1 MMA by each CTA into its columns 128.. (N=16, shifted A, its own B half rows 0..15)
        __shared__ uint64_t bar1;
        if (tid == 0) { mbar_init(&bar1, 1); mbar_fence_init(); }
        __syncthreads();
        if (warp == 0) {
            const bool leader = elect_one();
            if (leader) {
                mma_i8_ss(tm + 128, smem_desc(smem_u32(sA) + shift * 16, PLANE_B, 128), smem_desc(smem_u32(sB), (N / 2) * 16, 128),
                          idesc_i8(128, 16, 1, 1), 0);
                mma_commit(&bar1);
            }
            __syncwarp();
        }
        const bool ok1 = mbar_wait(&bar1, 0);
        fence_after_sync();
        if (!ok1 && tid == 0) atomicOr(status, 4 << rank);
        for (int j = 0; j < (N + 16) / 8; ++j) {
            uint32_t r[8];
            const int col = j < N / 8 ? j * 8 : 128 + (j - N / 8) * 8;
            tmem_ld_x8(tm + ((uint32_t)(warp * 32) << 16) + col, r);
            tmem_ld_wait();
            for (int i = 0; i < 8; ++i) out[(rank * 128 + warp * 32 + lane) * (N + 16) + j * 8 + i] = (int)r[i];
        }
    }
    fence_before_sync();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc2(tm, 512);
}

// lat: what one cross-CTA handshake costs.  Per round: CTA 0 issues one M = 256 MMA and a multicast commit; in each CTA one
// lane spins on the local mbarrier; then every thread of the cluster goes through barrier.cluster arrive + wait.
//   cycles[0] = issue -> CTA 0's own waiter sees the commit      (MMA execution + commit latency, same SM clock)
//   cycles[1] = issue -> CTA 0 leaves the cluster barrier        (+ the partner's waiter seeing it + barrier latency)
//   cycles[2] = barrier.cluster arrive + wait alone, all threads arriving together
__global__ void __cluster_dims__(2, 1, 1) k_lat(const int8_t *gA, long long *cycles, int *status, int rounds)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm, *sB = sm + A_BYTES;
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    __shared__ long long s_seen;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cta_rank();
    for (int i = tid; i < (A_BYTES + B_BYTES) / 16; i += blockDim.x) reinterpret_cast<int4 *>(sm)[i] = reinterpret_cast<const int4 *>(gA)[i % 512];
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc2(&s_tmem, 512);
    fence_before_sync();
    __syncthreads();
    cluster_sync();
    fence_after_sync();
    const uint32_t tm = s_tmem;
    const uint64_t ad = smem_desc(smem_u32(sA), PLANE_B, 128), bd = smem_desc(smem_u32(sB), 8 * 16, 128);
    const uint32_t id = idesc_i8(256, 16, 1, 1);
    long long a0 = 0, a1 = 0, a2 = 0;
    for (int r = 0; r < rounds; ++r) {
        long long t0 = 0;
        if (rank == 0 && warp == 0) {
            const bool leader = elect_one();
            t0 = clock64();
            if (leader) { mma2_i8_ss(tm, ad, bd, id, 0); commit2(&bar, 3); }
            __syncwarp();
        }
        if (warp == 1) {
            if ((tid & 31) == 0) { if (!mbar_wait(&bar, r & 1)) atomicOr(status, 1 << rank); s_seen = clock64(); }
            __syncwarp();
        }
        __syncthreads();
        cluster_sync();
        const long long t1 = clock64();
        if (rank == 0 && tid == 0) { a0 += s_seen - t0; a1 += t1 - t0; }
        __syncthreads();
    }
    for (int r = 0; r < rounds; ++r) {
        const long long t0 = clock64();
        cluster_sync();
        a2 += clock64() - t0;
    }
    if (rank == 0 && tid == 0) { cycles[0] = a0 / rounds; cycles[1] = a1 / rounds; cycles[2] = a2 / rounds; }
    fence_before_sync();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc2(tm, 512);
}

static int run_lat()
{
    std::vector<int8_t> hA(8192, 1);
    int8_t *dA; int *dSt; long long *dC;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dSt, 4)); CK(cudaMalloc(&dC, 32));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dSt, 0, 4));
    CK(cudaFuncSetAttribute(k_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, A_BYTES + B_BYTES));
    for (int rep = 0; rep < 2; ++rep) {
        k_lat<<<2, 128, A_BYTES + B_BYTES>>>(dA, dC, dSt, 200);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lat: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    }
    int st; long long c[3];
    CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(c, dC, 24, cudaMemcpyDeviceToHost));
    printf("lat cta_group::2: issue -> own waiter sees the multicast commit %lld cyc; issue -> out of the cluster barrier %lld cyc; "
           "barrier.cluster arrive+wait alone %lld cyc (status %d)\n", c[0], c[1], c[2], st);
    return 0;
}

template <int N>
static int run(int mode, int flags = 0)
{
    std::vector<int8_t> hA(2 * A_BYTES), hB(NMAX * 32);
    uint32_t s = 4242;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int8_t)(s >> 24); };
    for (auto &v : hA) v = rnd();
    for (auto &v : hB) v = rnd();
    int8_t *dA, *dB; int *dOut, *dSt; long long *dC;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dOut, 256 * (N + 16) * 4)); CK(cudaMalloc(&dSt, 4)); CK(cudaMalloc(&dC, 16));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dSt, 0, 4)); CK(cudaMemset(dOut, 0xEE, 256 * (N + 16) * 4));
    CK(cudaFuncSetAttribute(k2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, A_BYTES + NBT * B_BYTES));
    const int shift = 3;
    for (int rep = 0; rep < (mode ? 2 : 1); ++rep) {
        k2<N><<<2, 128, A_BYTES + NBT * B_BYTES>>>(dA, dB, dOut, mode, shift, dC, dSt, flags);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d mode=%d: CUDA error %s\n", N, mode, cudaGetErrorString(e)); return 2; }
    }
    int st; long long c[2];
    CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
    if (mode >= 1) {
        printf("thr flags=%2d cta_group::2 M=256 N=%3d : %.1f cyc/mma (status %d)\n", flags, N, (double)c[0] / (64 * 32), st);
        return 0;
    }
    std::vector<int> out(256 * (N + 16));
    CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
    long bad2 = 0, bad1 = 0;
    for (int r = 0; r < 2; ++r)
        for (int m = 0; m < 128; ++m) {
            for (int n = 0; n < N; ++n) {
                long acc = 0;
                for (int k = 0; k < 32; ++k) acc += (long)hA[((r * 2 + (k >> 4)) * PX + m + shift) * 16 + (k & 15)] * hB[n * 32 + k];
                bad2 += out[(r * 128 + m) * (N + 16) + n] != (int)acc;
            }
            for (int n = 0; n < 16; ++n) {          // the cta_group::1 MMA used this CTA's own B half: global rows r*N/2 + n
                long acc = 0;
                for (int k = 0; k < 32; ++k) acc += (long)hA[((r * 2 + (k >> 4)) * PX + m + shift) * 16 + (k & 15)] * hB[(r * (N / 2) + n) * 32 + k];
                bad1 += out[(r * 128 + m) * (N + 16) + N + n] != (int)acc;
            }
        }
    printf("num cta_group::2 M=256 N=%3d : status=%d  2-CTA product mismatches=%ld  mixed cta_group::1 N=16 mismatches=%ld -> %s\n", N, st, bad2,
           bad1, (!st && !bad2 && !bad1) ? "PASS" : "FAIL");
    return 0;
}

int main(int argc, char **argv)
{
    const char *t = argc > 1 ? argv[1] : "num";
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("# %s  test=%s\n", p.name, t);
    if (!strcmp(t, "lat")) return run_lat();
    const int mode = (!strcmp(t, "thr") || !strcmp(t, "thrb")) ? 1 : 0;
    const int flags = !strcmp(t, "thrb") ? 2 : (argc > 2 ? atoi(argv[2]) : 0);
    if (mode) run<16>(mode, flags);
    run<32>(mode, flags); run<64>(mode, flags); run<96>(mode, flags); run<128>(mode, flags);
    return 0;
}
