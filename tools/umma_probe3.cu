// umma_probe3 -- does the tcgen05.mma A-collector (collector::a::fill / use / lastuse) save the
// shared-memory read of the A operand for kind::i8 SS MMAs, and what are its semantics?
// Not part of the product; run on the B200 box, results summarised in profiles/.
//   numerics : MMA1 = A1*B1 (fill) ; MMA2 = "A2"*B2 (lastuse)  -> is D2 = A1*B2 (collector honoured) or A2*B2?
//   timing   : 32-MMA loops, N = 32..128, all-discard vs (fill,lastuse) pairs vs fill + use chain,
//              with and without 128 extra threads hammering shared memory (ld.shared) to expose the
//              smem-bandwidth share the tensor core needs.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../qcnn_gpu_b200/csrc/qv_tcgen05.cuh"

using namespace qv::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int COL>   // 0 discard, 1 fill, 2 use, 3 lastuse
__device__ __forceinline__ void mma_col(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    if (COL == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (COL == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (COL == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (COL == 3)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

constexpr int A_BYTES = 2 * 160 * 16;     // [2 K-planes][160 px][16 B]
constexpr int B_BYTES = 2 * 128 * 16;     // [2 K-chunks][128 rows][16 B]

// out[0]: D after MMA1 (N1 cols at col 0) and MMA2 (N2 cols at col 128)
// mode: 0 = MMA2 plain (discard) with A2 ; 1 = MMA1 fill, MMA2 lastuse with desc of A2 ; 2 = MMA1 fill, MMA2 lastuse with desc of A1
//       3 = MMA1 fill, <an unrelated plain MMA with A3 in between>, MMA2 lastuse with desc of A2
__global__ void k_num(const int8_t *gA, const int8_t *gB, int *out, int mode, int N1, int N2, int *status)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm, *sB = sm + A_BYTES;
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < A_BYTES / 16; i += blockDim.x) reinterpret_cast<int4 *>(sA)[i] = reinterpret_cast<const int4 *>(gA)[i];
    for (int i = tid; i < B_BYTES / 16; i += blockDim.x) reinterpret_cast<int4 *>(sB)[i] = reinterpret_cast<const int4 *>(gB)[i];
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = s_tmem;
    if (tid == 0) {
        const uint64_t a1 = smem_desc(smem_u32(sA), 160 * 16, 128);                 // pixels 0..127
        const uint64_t a2 = smem_desc(smem_u32(sA) + 9 * 16, 160 * 16, 128);        // pixels 9..136
        const uint64_t a3 = smem_desc(smem_u32(sA) + 20 * 16, 160 * 16, 128);       // pixels 20..147
        const uint64_t b1 = smem_desc(smem_u32(sB), 128 * 16, 128);                 // rows 0..N1
        const uint64_t b2 = smem_desc(smem_u32(sB) + 32 * 16, 128 * 16, 128);       // rows 32..32+N2
        if (mode == 0) {
            mma_col<0>(tm, a1, b1, idesc_i8(128, N1), 0);
            mma_col<0>(tm + 128, a2, b2, idesc_i8(128, N2), 0);
        } else if (mode == 1) {
            mma_col<1>(tm, a1, b1, idesc_i8(128, N1), 0);
            mma_col<3>(tm + 128, a2, b2, idesc_i8(128, N2), 0);
        } else if (mode == 2) {
            mma_col<1>(tm, a1, b1, idesc_i8(128, N1), 0);
            mma_col<3>(tm + 128, a1, b2, idesc_i8(128, N2), 0);
        } else {
            mma_col<1>(tm, a1, b1, idesc_i8(128, N1), 0);
            mma_col<0>(tm + 256, a3, b1, idesc_i8(128, N1), 0);
            mma_col<3>(tm + 128, a2, b2, idesc_i8(128, N2), 0);
        }
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    fence_after_sync();
    if (!ok) { if (tid == 0) status[0] = 1; }
    else {
        for (int j = 0; j < 32; ++j) {
            uint32_t r[8];
            tmem_ld_x8(tm + ((uint32_t)(warp * 32) << 16) + j * 8, r);
            tmem_ld_wait();
            for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * 256 + j * 8 + i] = (int)r[i];
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

static int run_num()
{
    std::vector<int8_t> hA(A_BYTES), hB(B_BYTES);
    uint32_t s = 4242;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int8_t)(s >> 24); };
    for (auto &v : hA) v = rnd();
    for (auto &v : hB) v = rnd();
    int8_t *dA, *dB; int *dOut, *dSt;
    CK(cudaMalloc(&dA, A_BYTES)); CK(cudaMalloc(&dB, B_BYTES)); CK(cudaMalloc(&dOut, 128 * 256 * 4)); CK(cudaMalloc(&dSt, 4));
    CK(cudaMemcpy(dA, hA.data(), A_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), B_BYTES, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_num, cudaFuncAttributeMaxDynamicSharedMemorySize, A_BYTES + B_BYTES));
    auto ref = [&](int m, int shift, int row0, int n) {
        long acc = 0;
        for (int k = 0; k < 32; ++k)
            acc += (long)hA[(k >> 4) * 160 * 16 + (m + shift) * 16 + (k & 15)] * hB[(k >> 4) * 128 * 16 + (row0 + n) * 16 + (k & 15)];
        return (int)acc;
    };
    const int shapes[][2] = {{96, 96}, {96, 64}, {64, 32}, {32, 32}, {96, 16}};
    for (auto &sh : shapes)
        for (int mode = 0; mode < 4; ++mode) {
            const int N1 = sh[0], N2 = sh[1];
            CK(cudaMemset(dSt, 0, 4));
            CK(cudaMemset(dOut, 0xEE, 128 * 256 * 4));
            k_num<<<1, 128, A_BYTES + B_BYTES>>>(dA, dB, dOut, mode, N1, N2, dSt);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("num mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
            int st; std::vector<int> out(128 * 256);
            CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
            long bad1 = 0, eqA1 = 0, eqA2 = 0, eqA3 = 0;
            for (int m = 0; m < 128; ++m) {
                for (int n = 0; n < N1; ++n) bad1 += out[m * 256 + n] != ref(m, 0, 0, n);
                for (int n = 0; n < N2; ++n) {
                    const int v = out[m * 256 + 128 + n];
                    eqA1 += v == ref(m, 0, 32, n);
                    eqA2 += v == ref(m, 9, 32, n);
                    eqA3 += v == ref(m, 20, 32, n);
                }
            }
            printf("num N1=%3d N2=%3d mode %d (%s): status=%d  D1 mismatches=%ld | D2 == A1*B2: %ld  == A2*B2: %ld  == A3*B2: %ld  of %d\n", N1, N2, mode,
                   mode == 0 ? "plain, desc A2" : mode == 1 ? "fill A1 ; lastuse, desc A2" : mode == 2 ? "fill A1 ; lastuse, desc A1" : "fill A1 ; plain A3 ; lastuse, desc A2",
                   st, bad1, eqA1, eqA2, eqA3, 128 * N2);
        }
    return 0;
}

// ---- timing -----------------------------------------------------------------------------------------
// PAT 0: all discard, every MMA its own A.  PAT 1: pairs (fill A_j, lastuse) with N1 then N2.
// PAT 2: pairs without collector hints (same A address twice) -- control for "same address" effects.
// PAT 3: one fill then 31 x use.
template <int N1, int N2, int PAT>
__global__ void k_thr(int outer, long long *cycles, int *status, int hammer)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm;                            // 2 planes x 1024 px x 16 B = 32 KB
    uint8_t *sB = sm + 32768;                    // 8 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (32768 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0x01010101u * (i & 3);
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); s_stop = 0; }
    if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = s_tmem;
    if (tid == 0) {
        constexpr uint32_t id1 = idesc_i8(128, N1, 1, 1), id2 = idesc_i8(128, N2, 1, 1);
        const uint64_t bd1 = smem_desc(smem_u32(sB), 128 * 16, 128), bd2 = smem_desc(smem_u32(sB) + 16 * 16, 128 * 16, 128);
        const uint64_t ad0 = smem_desc(smem_u32(sA), 16384, 128);
        long long t0 = clock64();
        for (int o = 0; o < outer; ++o) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint64_t a = ad0 + (uint64_t)((j * 37) % 800), a2 = ad0 + (uint64_t)((j * 37 + 19) % 800);
                if (PAT == 0) { mma_col<0>(tm, a, bd1, id1, 1); mma_col<0>(tm + 128, a2, bd2, id2, 1); }
                if (PAT == 1) { mma_col<1>(tm, a, bd1, id1, 1); mma_col<3>(tm + 128, a, bd2, id2, 1); }
                if (PAT == 2) { mma_col<0>(tm, a, bd1, id1, 1); mma_col<0>(tm + 128, a, bd2, id2, 1); }
                if (PAT == 3) {
                    if (j == 0) mma_col<1>(tm, a, bd1, id1, 1); else mma_col<2>(tm, ad0, bd1, id1, 1);
                    mma_col<2>(tm + 128, ad0, bd2, id2, 1);
                }
            }
        }
        long long t1 = clock64();
        mma_commit(&bar);
        const bool ok = mbar_wait(&bar, 0);
        long long t2 = clock64();
        cycles[0] = t1 - t0;
        cycles[1] = t2 - t0;
        status[0] = ok ? 0 : 1;
        s_stop = 1;
    } else if (hammer && tid >= 128) {
        // conflict-free 128-byte ld.shared wavefronts, as fast as 4 warps can issue them
        const volatile uint32_t *p = reinterpret_cast<const volatile uint32_t *>(sm) + (tid & 31);
        uint32_t acc = 0;
        long long n = 0;
        while (!s_stop) {
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += p[u * 32];
            n += 8;
        }
        if (acc == 0x12345678u) cycles[7] = acc;
        if ((tid & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&cycles[2]), (unsigned long long)n);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N1, int N2, int PAT>
static void thr_case(long long *dC, int *dSt)
{
    static const char *names[] = {"discard, distinct A", "fill/lastuse pairs", "same A twice, no hint", "fill + use chain"};
    CK(cudaFuncSetAttribute(k_thr<N1, N2, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 8192));
    const int outer = 64;
    for (int hammer = 0; hammer < 2; ++hammer) {
        long long c[4]; int st = 0;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaMemset(dSt, 0, 4));
            CK(cudaMemset(dC, 0, 64));
            k_thr<N1, N2, PAT><<<1, 256, 32768 + 8192>>>(outer, dC, dSt, hammer);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("thr: CUDA error %s\n", cudaGetErrorString(e)); exit(2); }
        }
        CK(cudaMemcpy(c, dC, 32, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
        const double per = (double)c[1] / (outer * 16);
        printf("thr N1=%3d N2=%3d %-24s %s: %.1f cyc per pair (smem wavefronts if A read each time %d, if A reused %d; math %d)  lsu wavefronts/clk %.2f timeout=%d\n",
               N1, N2, names[PAT], hammer ? "+ld.shared hammer" : "alone            ", per, (128 + N1) / 4 + (128 + N2) / 4, (128 + N1) / 4 + N2 / 4,
               N1 / 2 + N2 / 2, (double)c[2] / (double)c[1], st);
    }
}


// ---- contention: what slows a back-to-back MMA stream down? ------------------------------------------
// Thread 0 (warp 0, SM sub-partition 0) issues pairs (N=96 fill, N=128 lastuse) as in the fused kernel.  `nh`
// hammer warps, placed on sub-partition `sp` (warp % 4 == sp) or spread over all four (sp < 0), run one of:
//   1 ld.shared 128 B wavefronts   2 st.shared.v4 (512 B per warp instruction)   3 tcgen05.ld 32x32b.x16
//   4 integer ALU only (IMAD / VIADDMNMX-like chains, 8 independent)             5 mbarrier.try_wait polling
template <int MODE>
__global__ void k_cont(int outer, long long *cycles, int *status, int nh, int sp)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm, *sB = sm + 32768;
    uint8_t *scratch = sm + 32768 + 8192;        // 16 KB for the st.shared hammer
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (32768 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0x01010101u * (i & 3);
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_fence_init(); s_stop = 0; }
    if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = s_tmem;
    // which warps hammer: warps 1.. ; on sub-partition sp only, or all
    bool hammer = false;
    if (warp >= 1) {
        if (sp < 0) hammer = warp <= nh;
        else hammer = (warp & 3) == sp && (warp >> 2) >= (sp == 0 ? 1 : 0) && ((warp >> 2) - (sp == 0 ? 1 : 0)) < nh;
    }
    if (tid == 0) {
        constexpr uint32_t id1 = idesc_i8(128, 96, 1, 1), id2 = idesc_i8(128, 128, 1, 1);
        const uint64_t bd1 = smem_desc(smem_u32(sB), 128 * 16, 128), bd2 = smem_desc(smem_u32(sB) + 16 * 16, 128 * 16, 128);
        const uint64_t ad0 = smem_desc(smem_u32(sA), 16384, 128);
        long long t0 = clock64();
        for (int o = 0; o < outer; ++o) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint64_t a = ad0 + (uint64_t)((j * 37) % 800);
                mma_col<1>(tm, a, bd1, id1, 1);
                mma_col<3>(tm + 128, a, bd2, id2, 1);
            }
        }
        long long t1 = clock64();
        mma_commit(&bar);
        const bool ok = mbar_wait(&bar, 0);
        long long t2 = clock64();
        cycles[0] = t1 - t0;
        cycles[1] = t2 - t0;
        status[0] = ok ? 0 : 1;
        s_stop = 1;
    } else if (hammer) {
        long long n = 0;
        uint32_t acc = lane;
        if (MODE == 1) {
            const volatile uint32_t *p = reinterpret_cast<const volatile uint32_t *>(sm) + lane;
            while (!s_stop) {
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += p[u * 32];
                n += 8;
            }
        } else if (MODE == 2) {
            uint4 *p = reinterpret_cast<uint4 *>(scratch) + (warp & 7) * 64 + lane;
            while (!s_stop) {
#pragma unroll
                for (int u = 0; u < 2; ++u) { p[u * 32] = make_uint4(acc, n, u, 1); }
                asm volatile("" ::: "memory");
                n += 8;           // 2 x 4 wavefronts
            }
        } else if (MODE == 3) {
            const uint32_t ta = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256;
            while (!s_stop) {
                uint32_t r[16];
                tmem_ld_x16(ta + ((n & 7) * 16), r);
                tmem_ld_wait();
                acc += r[0] + r[15];
                n += 1;
            }
        } else if (MODE == 4) {
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = acc + u;
            while (!s_stop) {
#pragma unroll
                for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = v[u] * 1664525u + (v[(u + 1) & 7] >> 3);
                n += 32;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        } else if (MODE == 6) {          // IADD3 / LOP3 only (alu pipe)
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = acc + u;
            while (!s_stop) {
#pragma unroll
                for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = (v[u] + v[(u + 1) & 7]) ^ (v[(u + 3) & 7] | 0x55u);
                n += 32;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        } else if (MODE == 7) {          // FFMA only
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (float)(acc + u);
            while (!s_stop) {
#pragma unroll
                for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = fmaf(v[u], 1.0001f, v[(u + 1) & 7]);
                n += 32;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += (uint32_t)v[u];
        } else if (MODE == 8) {          // IDP.4A only
            int v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (int)(acc + u);
            while (!s_stop) {
#pragma unroll
                for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = __dp4a(v[(u + 1) & 7], 0x01020304, v[u]);
                n += 32;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += (uint32_t)v[u];
        } else if (MODE == 9) {          // the requantiser mix: VIADDMNMX.RELU + IMAD + PRMT
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = acc + u;
            while (!s_stop) {
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = (uint32_t)__viaddmin_s32_relu((int)v[u], 37 + u, 19000) * 6431u;
#pragma unroll
                    for (int u = 0; u < 8; u += 2) v[u] = __byte_perm(v[u], v[u + 1], 0x0073);
                }
                n += 40;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        } else if (MODE == 5) {
            while (!s_stop) {
                acc += mbar_try_wait(&bar2, 0) ? 1u : 0u;      // never completes: every call is a (suspended) poll
                n += 1;
            }
        }
        if (acc == 0x12345678u) cycles[7] = acc;
        if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&cycles[2]), (unsigned long long)n);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

template <int MODE>
static void cont_case(long long *dC, int *dSt, int nh, int sp)
{
    static const char *names[] = {"", "ld.shared", "st.shared.v4", "tcgen05.ld.x16", "IMAD chain", "mbarrier poll", "IADD3/LOP3", "FFMA", "IDP.4A", "requant mix"};
    const int smem = 32768 + 8192 + 16384;
    CK(cudaFuncSetAttribute(k_cont<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long c[4]; int st = 0;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(dSt, 0, 4));
        CK(cudaMemset(dC, 0, 64));
        k_cont<MODE><<<1, 17 * 32, smem>>>(64, dC, dSt, nh, sp);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cont: CUDA error %s\n", cudaGetErrorString(e)); exit(2); }
    }
    CK(cudaMemcpy(c, dC, 32, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
    printf("cont %-15s x%2d warps on %-18s: %.1f cyc per (N=96,N=128) pair [112 alone]; hammer ops/clk %.3f timeout=%d\n", names[MODE], nh,
           sp < 0 ? "all sub-partitions" : sp == 0 ? "the MMA warp's SMSP" : "another SMSP", (double)c[1] / (64 * 16), (double)c[2] / (double)c[1], st);
}

// ---- queue depth: how long does the issuing thread take to hand k MMAs to an idle tensor pipe? --------
template <int K, int N>
__global__ void k_burst(long long *cycles, int *status)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm, *sB = sm + 32768;
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (32768 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0x01010101u * (i & 3);
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = s_tmem;
    if (tid == 0) {
        constexpr uint32_t id = idesc_i8(128, N, 1, 1);
        const uint64_t bd = smem_desc(smem_u32(sB), 128 * 16, 128);
        const uint64_t ad0 = smem_desc(smem_u32(sA), 16384, 128);
        long long t0 = clock64();
#pragma unroll
        for (int j = 0; j < K; ++j) mma_col<0>(tm, ad0 + (uint64_t)((j * 37) % 800), bd, id, 1);
        long long t1 = clock64();
        mma_commit(&bar);
        const bool ok = mbar_wait(&bar, 0);
        long long t2 = clock64();
        cycles[0] = t1 - t0;
        cycles[1] = t2 - t0;
        status[0] = ok ? 0 : 1;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}
template <int K, int N>
static void burst_case(long long *dC, int *dSt)
{
    CK(cudaFuncSetAttribute(k_burst<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 8192));
    long long c[2]; int st = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(dSt, 0, 4));
        k_burst<K, N><<<1, 128, 32768 + 8192>>>(dC, dSt);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
    printf("burst N=%3d k=%2d MMAs: issued after %lld cycles, complete (commit seen) after %lld  timeout=%d\n", N, K, c[0], c[1], st);
}


// ---- two issuing threads: is the ~46-cycle issue cost of an MMA a property of the issuing THREAD (then two warps could feed
// ---- the tensor pipe twice as fast with small MMAs) or of the pipe's front end?  NW warps (on different SM sub-partitions)
// ---- each issue K MMAs of width N into their own TMEM region, own B tile, own A tile; REUSE = every second MMA takes A from
// ---- the collector (fill / lastuse pairs).  Reported: cycles until ALL commits are seen, per MMA.
template <int NW, int K, int N, int REUSE>
__global__ void k_dual(long long *cycles, int *status)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm, *sB = sm + 32768;
    __shared__ uint64_t bar[4];
    __shared__ uint32_t s_tmem;
    __shared__ long long s_t[4][2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (32768 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0x01010101u * (i & 3);
    fence_proxy_async_smem();
    if (tid == 0) { for (int w = 0; w < 4; ++w) mbar_init(&bar[w], 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = s_tmem;
    if (warp < NW && lane == 0) {
        constexpr uint32_t id = idesc_i8(128, N, 1, 1);
        const uint64_t bd = smem_desc(smem_u32(sB + warp * 4096), 128 * 16, 128);
        const uint64_t ad0 = smem_desc(smem_u32(sA + warp * 8192), 4096, 128);
        const uint32_t d = tm + warp * 128;
        long long t0 = clock64();
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (REUSE) { if (j & 1) mma_col<3>(d, ad0, bd, id, 1); else mma_col<1>(d, ad0, bd, id, 1); }
            else mma_col<0>(d, ad0 + (uint64_t)((j * 37) % 200), bd, id, 1);
        }
        long long t1 = clock64();
        mma_commit(&bar[warp]);
        const bool ok = mbar_wait(&bar[warp], 0);
        long long t2 = clock64();
        s_t[warp][0] = t1 - t0; s_t[warp][1] = t2 - t0;
        if (!ok) status[0] = 1;
    }
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        long long a = 0, b = 0;
        for (int w = 0; w < NW; ++w) { a = max(a, s_t[w][0]); b = max(b, s_t[w][1]); }
        cycles[0] = a; cycles[1] = b;
    }
    if (warp == 0) tmem_dealloc(tm, 512);
}
template <int NW, int K, int N, int REUSE>
static void dual_case(long long *dC, int *dSt)
{
    CK(cudaFuncSetAttribute(k_dual<NW, K, N, REUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 16384));
    long long c[2]; int st = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(dSt, 0, 4));
        k_dual<NW, K, N, REUSE><<<1, 128, 32768 + 16384>>>(dC, dSt);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
    printf("dual  %d issuing warp(s) x %2d MMAs N=%3d %-22s: issue %.1f cyc/MMA/warp, all complete after %lld = %.1f cyc per MMA overall  timeout=%d\n", NW, K, N,
           REUSE ? "(A fill/lastuse pairs)" : "(A from smem each)", (double)c[0] / K, c[1], (double)c[1] / (K * NW), st);
}

// ---- handshake latencies: what one leg of the MMA-warp / worker ping-pong costs ---------------------------------------
// mode 0: pure mbarrier ping-pong between warp 0 and warp 1 (arrive -> try_wait sees it), round trip / 2
// mode 1: warp 0 issues one small MMA (N = 16) + tcgen05.commit, warp 1 waits for the commit and arrives back
// mode 2: as mode 1 with `k` wide MMAs (N = 128) queued before the commit (latency from LAST issue to commit seen = total - issue)
template <int W>
__device__ __forceinline__ bool pp_wait(uint64_t *bar, uint32_t parity)
{
    if (W == 0) return mbar_wait(bar, parity);                       // the product's wait: try_wait + suspend-time hint
    if (W == 1) {                                                    // try_wait without a hint, spinning
        for (int n = 0; n < 4000000; ++n) if (mbar_try_wait(bar, parity)) return true;
        return false;
    }
    for (int n = 0; n < 40000000; ++n) {                             // test_wait: non-blocking poll
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
        if (W == 3) __nanosleep(20);
    }
    return false;
}
template <int W>
__global__ void k_pingpong(int mode, int kmma, int rounds, long long *cycles, int *status)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar_a[2], bar_b[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (32768 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0x01010101u;
    fence_proxy_async_smem();
    if (tid == 0) { for (int i = 0; i < 2; ++i) { mbar_init(&bar_a[i], 1); mbar_init(&bar_b[i], 1); } mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = s_tmem;
    bool ok = true;
    if (warp == 0) {
        const uint64_t ad = smem_desc(smem_u32(sm), 16384, 128), bd = smem_desc(smem_u32(sm) + 32768, 128 * 16, 128);
        long long t_issue = 0;
        const long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            if (lane == 0) {
                if (mode == 0) mbar_arrive(&bar_a[r & 1]);
                else {
                    const long long ti = clock64();
                    for (int j = 0; j < kmma; ++j) mma_col<0>(tm, ad, bd, idesc_i8(128, mode == 1 ? 16 : 128), 1);
                    mma_commit(&bar_a[r & 1]);
                    t_issue += clock64() - ti;
                }
                ok &= pp_wait<W>(&bar_b[r & 1], (r >> 1) & 1);
            }
            __syncwarp();
            fence_after_sync();
        }
        if (lane == 0) { cycles[0] = clock64() - t0; cycles[1] = t_issue; status[0] = ok ? 0 : 1; }
    } else if (warp == 1) {
        for (int r = 0; r < rounds; ++r) {
            if (lane == 0) ok &= pp_wait<W>(&bar_a[r & 1], (r >> 1) & 1);
            __syncwarp();
            fence_after_sync();
            fence_before_sync();
            if (lane == 0) mbar_arrive(&bar_b[r & 1]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- the same ping-pong through plain shared-memory counters (red.shared + ld.volatile.shared spin) instead of mbarriers ----
__global__ void k_pingpong_flag(int rounds, int nwait_warps, long long *cycles)
{
    __shared__ unsigned int ca, cb;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { ca = 0; cb = 0; }
    __syncthreads();
    volatile unsigned int *va = &ca, *vb = &cb;
    if (warp == 0) {
        const long long t0 = clock64();
        for (int r = 1; r <= rounds; ++r) {
            if (lane == 0) {
                atomicAdd(&ca, 1u);
                while (*vb < (unsigned)(r * nwait_warps)) { }
            }
            __syncwarp();
        }
        if (lane == 0) cycles[0] = clock64() - t0;
    } else if (warp <= nwait_warps) {
        for (int r = 1; r <= rounds; ++r) {
            if (lane == 0) {
                while (*va < (unsigned)r) { }
                __threadfence_block();
                atomicAdd(&cb, 1u);
            }
            __syncwarp();
        }
    }
}

int main(int argc, char **argv)
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("# %s  sm_%d%d  %d SMs\n", p.name, p.major, p.minor, p.multiProcessorCount);
    const char *t = argc > 1 ? argv[1] : "all";
    if (!strcmp(t, "num") || !strcmp(t, "all"))
        if (int rc = run_num()) return rc;
    if (!strcmp(t, "thr") || !strcmp(t, "all")) {
        long long *dC; int *dSt;
        CK(cudaMalloc(&dC, 64)); CK(cudaMalloc(&dSt, 4));
        thr_case<96, 128, 0>(dC, dSt); thr_case<96, 128, 1>(dC, dSt); thr_case<96, 128, 2>(dC, dSt); thr_case<96, 128, 3>(dC, dSt);
        thr_case<64, 32, 0>(dC, dSt); thr_case<64, 32, 1>(dC, dSt); thr_case<64, 32, 3>(dC, dSt);
        thr_case<16, 32, 0>(dC, dSt); thr_case<16, 32, 1>(dC, dSt);
        thr_case<96, 96, 0>(dC, dSt); thr_case<96, 96, 1>(dC, dSt);
    }
    if (!strcmp(t, "cont") || !strcmp(t, "all")) {
        long long *dC; int *dSt;
        CK(cudaMalloc(&dC, 64)); CK(cudaMalloc(&dSt, 4));
        cont_case<4>(dC, dSt, 0, 0);
        cont_case<1>(dC, dSt, 8, -1); cont_case<1>(dC, dSt, 2, 0); cont_case<1>(dC, dSt, 2, 1);
        cont_case<2>(dC, dSt, 8, -1); cont_case<2>(dC, dSt, 2, 0); cont_case<2>(dC, dSt, 2, 1);
        cont_case<3>(dC, dSt, 8, -1); cont_case<3>(dC, dSt, 16, -1); cont_case<3>(dC, dSt, 2, 0); cont_case<3>(dC, dSt, 2, 1);
        cont_case<4>(dC, dSt, 8, -1); cont_case<4>(dC, dSt, 16, -1); cont_case<4>(dC, dSt, 1, 0); cont_case<4>(dC, dSt, 2, 0); cont_case<4>(dC, dSt, 3, 0); cont_case<4>(dC, dSt, 2, 1);
        cont_case<5>(dC, dSt, 8, -1); cont_case<5>(dC, dSt, 2, 0);
    }
    if (!strcmp(t, "burst") || !strcmp(t, "all")) {
        long long *dC; int *dSt;
        CK(cudaMalloc(&dC, 64)); CK(cudaMalloc(&dSt, 4));
        burst_case<1, 128>(dC, dSt); burst_case<2, 128>(dC, dSt); burst_case<4, 128>(dC, dSt); burst_case<8, 128>(dC, dSt);
        burst_case<16, 128>(dC, dSt); burst_case<32, 128>(dC, dSt); burst_case<64, 128>(dC, dSt);
        burst_case<1, 32>(dC, dSt); burst_case<4, 32>(dC, dSt); burst_case<8, 32>(dC, dSt); burst_case<16, 32>(dC, dSt); burst_case<32, 32>(dC, dSt);
    }
    if (!strcmp(t, "dual")) {
        long long *dC; int *dSt;
        CK(cudaMalloc(&dC, 64)); CK(cudaMalloc(&dSt, 4));
        dual_case<1, 32, 16, 0>(dC, dSt); dual_case<2, 16, 16, 0>(dC, dSt); dual_case<2, 32, 16, 0>(dC, dSt); dual_case<4, 16, 16, 0>(dC, dSt);
        dual_case<1, 32, 16, 1>(dC, dSt); dual_case<2, 16, 16, 1>(dC, dSt); dual_case<2, 32, 16, 1>(dC, dSt);
        dual_case<1, 32, 64, 0>(dC, dSt); dual_case<2, 16, 64, 0>(dC, dSt); dual_case<2, 32, 64, 0>(dC, dSt);
        dual_case<1, 32, 64, 1>(dC, dSt); dual_case<2, 32, 64, 1>(dC, dSt);
        dual_case<1, 32, 128, 0>(dC, dSt); dual_case<2, 32, 128, 0>(dC, dSt); dual_case<2, 32, 128, 1>(dC, dSt);
    }
    if (!strcmp(t, "cont2")) {
        long long *dC; int *dSt;
        CK(cudaMalloc(&dC, 64)); CK(cudaMalloc(&dSt, 4));
        cont_case<4>(dC, dSt, 2, 0); cont_case<6>(dC, dSt, 2, 0); cont_case<7>(dC, dSt, 2, 0); cont_case<8>(dC, dSt, 2, 0); cont_case<9>(dC, dSt, 2, 0);
        cont_case<9>(dC, dSt, 1, 0); cont_case<9>(dC, dSt, 4, 0); cont_case<9>(dC, dSt, 2, 1);
    }
    if (!strcmp(t, "pingpong") || !strcmp(t, "all")) {
        long long *dC; int *dSt;
        CK(cudaMalloc(&dC, 64)); CK(cudaMalloc(&dSt, 4));
        const int rounds = 2000;
        const int cfg[][2] = {{0, 0}, {1, 1}, {2, 1}, {2, 8}};
        auto run = [&](auto kern, const char *wname) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 8192));
            for (auto &c : cfg) {
                long long h[2]; int st = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaMemset(dSt, 0, 4));
                    kern<<<1, 64, 32768 + 8192>>>(c[0], c[1], rounds, dC, dSt);
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(h, dC, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
                if (c[0] == 0) printf("pingpong [%s] mbarrier only: %.0f cycles per round trip (two legs)  timeout=%d\n", wname, (double)h[0] / rounds, st);
                else printf("pingpong [%s] %d x MMA N=%d + commit -> seen -> arrive -> seen: %.0f cycles per round, of which issuing %.0f  timeout=%d\n",
                            wname, c[1], c[0] == 1 ? 16 : 128, (double)h[0] / rounds, (double)h[1] / rounds, st);
            }
        };
        run(k_pingpong<0>, "try_wait + 10 us hint");
        run(k_pingpong<1>, "try_wait, no hint");
        run(k_pingpong<2>, "test_wait spin");
        run(k_pingpong<3>, "test_wait + nanosleep(20)");
        for (int nw : {1, 8, 12}) {
            long long h[1];
            for (int rep = 0; rep < 2; ++rep) { k_pingpong_flag<<<1, 32 * (nw + 1)>>>(rounds, nw, dC); CK(cudaDeviceSynchronize()); }
            CK(cudaMemcpy(h, dC, 8, cudaMemcpyDeviceToHost));
            printf("pingpong [smem counters, atomicAdd + volatile spin] 1 signaller, %d waiting warps that all answer: %.0f cycles per round trip\n", nw, (double)h[0] / rounds);
        }
    }
    return 0;
}
