// umma_probe -- hardware probe for the tcgen05 facts the fused QVRCNN kernel is designed on.
// Not part of the product; run on the B200 box, results summarised in profiles/.
//   umma_probe num     numerics of kind::i8 SS MMA on the [plane][pixel][16B] no-swizzle K-major
//                      layout with tap-shifted start addresses and arbitrary LBO
//   umma_probe thr     cycles per MMA for M=128, N in {8..256}, SS (A from smem) and TS (A from TMEM)
//   umma_probe ts      numerics of A-from-TMEM (tcgen05.st -> mma.ts) and of tcgen05.cp 128x256b
//   umma_probe ldst    tcgen05.ld / tcgen05.st throughput
//   umma_probe shift   what tcgen05.shift.down does
// Every wait is bounded, so a wrong guess produces a report, not a hung GPU.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../qcnn_gpu_b200/csrc/qv_tcgen05.cuh"

using namespace qv::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int PLANE_PX = 320;                 // pixels per 16-channel plane
constexpr int PLANE_B = PLANE_PX * 16;
constexpr int NPLANE = 3;
constexpr int A_BYTES = NPLANE * PLANE_B;     // 15360
constexpr int B_MAXN = 256;
constexpr int B_BYTES = 2 * B_MAXN * 16;      // [kchunk][n][16B]

struct NumCase { int plane0, shift0, plane1, shift1, N, a_signed, two; };

// ---- numerics -------------------------------------------------------------------------
__global__ void k_num(const int8_t *gA, const int8_t *gB, int *out, NumCase c, int *status)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm, *sB = sm + A_BYTES;
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < A_BYTES / 16; i += blockDim.x) reinterpret_cast<int4 *>(sA)[i] = reinterpret_cast<const int4 *>(gA)[i];
    // B global is plain [N][32]; smem is [kchunk][n][16]
    for (int i = tid; i < c.N * 2; i += blockDim.x) {
        const int n = i >> 1, kc = i & 1;
        reinterpret_cast<int4 *>(sB)[kc * c.N + n] = reinterpret_cast<const int4 *>(gB)[n * 2 + kc];
    }
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&s_tmem, 256); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = s_tmem;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(sA) + c.plane0 * PLANE_B + c.shift0 * 16;
        const uint32_t a1 = smem_u32(sA) + c.plane1 * PLANE_B + c.shift1 * 16;
        const uint64_t ad = smem_desc(a0, (a1 - a0) & 0x3FFFFu, 128);
        const uint64_t bd = smem_desc(smem_u32(sB), c.N * 16, 128);
        const uint32_t id = idesc_i8(128, c.N, c.a_signed, 1);
        mma_i8_ss(tm, ad, bd, id, 0);
        if (c.two) mma_i8_ss(tm, ad, bd, id, 1);    // accumulate the same product again -> 2x
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    fence_after_sync();
    if (!ok) { if (tid == 0) status[0] = 1; }
    else {
        for (int j = 0; j < c.N / 8; ++j) {
            uint32_t r[8];
            tmem_ld_x8(tm + ((uint32_t)(warp * 32) << 16) + j * 8, r);
            tmem_ld_wait();
            for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * c.N + j * 8 + i] = (int)r[i];
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

static int run_num()
{
    std::vector<int8_t> hA(A_BYTES), hB(B_MAXN * 32);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int8_t)(s >> 24); };
    for (auto &v : hA) v = rnd();
    for (auto &v : hB) v = rnd();
    int8_t *dA, *dB; int *dOut, *dSt;
    CK(cudaMalloc(&dA, A_BYTES)); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dOut, 128 * 256 * 4)); CK(cudaMalloc(&dSt, 4));
    CK(cudaMemcpy(dA, hA.data(), A_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_num, cudaFuncAttributeMaxDynamicSharedMemorySize, A_BYTES + B_BYTES));
    const NumCase cases[] = {
        {0, 0, 1, 0, 48, 1, 0},      // aligned, planes 0/1 (the plain layout)
        {0, 1, 1, 1, 48, 1, 0},      // tap shift of 1 pixel (16 B): start not 128B-aligned
        {0, 37, 1, 37, 48, 1, 0},    // arbitrary shift
        {0, 139, 1, 139, 16, 1, 0},  // N=16, shift = one row of a 136-px pitch + 3
        {0, 5, 2, 9, 48, 1, 0},      // different taps in the two K halves (layer-3 style pairing)
        {2, 9, 0, 5, 48, 1, 0},      // NEGATIVE LBO (second half below the first): wraps?
        {0, 0, 1, 0, 64, 1, 1},      // accumulate flag
        {0, 3, 1, 3, 48, 0, 0},      // A declared unsigned
        {0, 2, 1, 2, 8, 1, 0},       // N=8
        {0, 2, 1, 2, 256, 1, 0},     // N=256
    };
    int fails = 0;
    for (const NumCase &c : cases) {
        CK(cudaMemset(dSt, 0, 4));
        CK(cudaMemset(dOut, 0xEE, 128 * 256 * 4));
        k_num<<<1, 128, A_BYTES + B_BYTES>>>(dA, dB, dOut, c, dSt);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("num case p0=%d s0=%d p1=%d s1=%d N=%d: CUDA error %s\n", c.plane0, c.shift0, c.plane1, c.shift1, c.N, cudaGetErrorString(e)); return 2; }
        int st; std::vector<int> out(128 * c.N);
        CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
        long bad = 0; int first = -1;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < c.N; ++n) {
                long acc = 0;
                for (int k = 0; k < 32; ++k) {
                    const int plane = k < 16 ? c.plane0 : c.plane1, shift = k < 16 ? c.shift0 : c.shift1;
                    int a = hA[plane * PLANE_B + (m + shift) * 16 + (k & 15)];
                    if (!c.a_signed) a &= 0xff;
                    acc += (long)a * hB[n * 32 + k];
                }
                if (c.two) acc *= 2;
                if (out[m * c.N + n] != (int)acc) { if (first < 0) first = m * c.N + n; ++bad; }
            }
        printf("num p0=%d s0=%d p1=%d s1=%d N=%d a_signed=%d two=%d : %s  timeout=%d mismatches=%ld", c.plane0, c.shift0, c.plane1,
               c.shift1, c.N, c.a_signed, c.two, bad == 0 && !st ? "PASS" : "FAIL", st, bad);
        if (bad) printf(" first@(m=%d,n=%d) got=%d", first / c.N, first % c.N, out[first]);
        printf("\n");
        fails += (bad != 0 || st);
    }
    return fails ? 1 : 0;
}

// ---- throughput -----------------------------------------------------------------------
// mode 0: SS, A address cycles over tap-like offsets.  mode 1: TS (A from TMEM columns 256..).
// mode 2: SS, two alternating accumulators.
__global__ void k_thr(int N, int mode, int nmma, long long *cycles, int *status)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm;                            // 2 planes x 1024 px x 16 B = 32 KB
    uint8_t *sB = sm + 32768;                    // 8 KB
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
        const uint32_t id = idesc_i8(128, N, 1, 1);
        const uint64_t bd = smem_desc(smem_u32(sB), N * 16, 128);
        const uint32_t abase = smem_u32(sA);
        long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            if (mode == 1) {
                mma_i8_ts(tm, tm + 256 + (i & 7) * 8, bd, id, i > 0);
            } else {
                const uint32_t off = (uint32_t)((i * 37) % 800) * 16;
                const uint64_t ad = smem_desc(abase + off, 16384, 128);
                mma_i8_ss(mode == 2 ? tm + (i & 1) * 256 : tm, ad, bd, id, i > 1);
            }
        }
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

static int run_thr()
{
    long long *dC; int *dSt;
    CK(cudaMalloc(&dC, 16)); CK(cudaMalloc(&dSt, 4));
    CK(cudaFuncSetAttribute(k_thr, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 8192));
    const int Ns[] = {8, 16, 32, 48, 64, 96, 128, 192, 256};
    const char *names[] = {"SS", "TS", "SS-2acc"};
    for (int mode = 0; mode < 3; ++mode)
        for (int N : Ns) {
            if (mode == 2 && N > 128) continue;
            const int nmma = 2000;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaMemset(dSt, 0, 4));
                k_thr<<<1, 128, 32768 + 8192>>>(N, mode, nmma, dC, dSt);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("thr %s N=%d: CUDA error %s\n", names[mode], N, cudaGetErrorString(e)); return 2; }
            }
            long long c[2]; int st;
            CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
            const double per = (double)c[1] / nmma;
            printf("thr %-7s M=128 N=%3d : issue %.1f cyc/mma, complete %.1f cyc/mma -> %.0f MAC/clk/SM (math floor %.1f cyc) timeout=%d\n",
                   names[mode], N, (double)c[0] / nmma, per, 128.0 * N * 32 / per, N / 2.0, st);
        }
    return 0;
}


// ---- throughput v2: unrolled issue loop, descriptors formed by one add of a constant -------
// MODE 0: SS.  MODE 1: TS.  MODE 2: tcgen05.cp.128x256b only.  MODE 3: L2-like mix for one output
// row (18 x N=48 + 32 x N=16, SS).  MODE 4: TMEM-A mix: 10 cp + (18 x N=48 + 32 x N=16) TS.
template <int N, int MODE>
__global__ void k_thr2(int outer, long long *cycles, int *status)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm;                            // 2 planes x 1024 px x 16 B = 32 KB
    uint8_t *sB = sm + 32768;                    // 8 KB
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
        constexpr uint32_t id = idesc_i8(128, N, 1, 1), id48 = idesc_i8(128, 48, 1, 1), id16 = idesc_i8(128, 16, 1, 1);
        const uint64_t bd = smem_desc(smem_u32(sB), N * 16, 128);
        const uint64_t bd48 = smem_desc(smem_u32(sB), 48 * 16, 128), bd16 = smem_desc(smem_u32(sB), 16 * 16, 128);
        const uint64_t ad0 = smem_desc(smem_u32(sA), 16384, 128);
        long long t0 = clock64();
        for (int o = 0; o < outer; ++o) {
            if (MODE == 0 || MODE == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t acc = (j > 0) ? 1u : (o > 0);
                    if (MODE == 0) mma_i8_ss(tm, ad0 + (uint64_t)((j * 37) % 800), bd, id, acc);
                    else mma_i8_ts(tm, tm + 256 + (j & 7) * 8, bd, id, acc);
                }
            } else if (MODE == 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tm + 256 + (j & 15) * 8), "l"(ad0 + (uint64_t)((j * 37) % 800)) : "memory");
            } else {
                if (MODE == 4) {
#pragma unroll
                    for (int j = 0; j < 10; ++j)
                        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tm + 256 + j * 8), "l"(ad0 + (uint64_t)((j * 37) % 800)) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 50; ++j) {
                    const bool inner = j < 18;
                    const uint32_t acc = (j > 0) ? 1u : (o > 0);
                    const uint32_t d = inner ? tm : tm + 0;     // outer taps target the first 16 columns
                    if (MODE == 3) mma_i8_ss(d, ad0 + (uint64_t)((j * 37) % 800), inner ? bd48 : bd16, inner ? id48 : id16, acc);
                    else mma_i8_ts(d, tm + 256 + (j & 31) * 8, inner ? bd48 : bd16, inner ? id48 : id16, acc);
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
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int MODE>
static void thr2_case(const char *name, long long *dC, int *dSt)
{
    CK(cudaFuncSetAttribute(k_thr2<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 8192));
    const int outer = 64;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(dSt, 0, 4));
        k_thr2<N, MODE><<<1, 128, 32768 + 8192>>>(outer, dC, dSt);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("thr2 %s N=%d: CUDA error %s\n", name, N, cudaGetErrorString(e)); exit(2); }
    }
    long long c[2]; int st;
    CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
    if (MODE <= 1) {
        const double per = (double)c[1] / (outer * 32);
        printf("thr2 %-3s M=128 N=%3d : issue %.1f, complete %.1f cyc/mma -> %.0f MAC/clk/SM (math floor %.1f, smem bytes/clk %.0f) timeout=%d\n", name, N,
               (double)c[0] / (outer * 32), per, 128.0 * N * 32 / per, N / 2.0, (MODE == 0 ? 4096.0 + 32.0 * N : 32.0 * N) / per, st);
    } else if (MODE == 2) {
        const double per = (double)c[1] / (outer * 32);
        printf("thr2 cp.128x256b : issue %.1f, complete %.1f cyc/copy -> %.0f B/clk timeout=%d\n", (double)c[0] / (outer * 32), per, 4096.0 / per, st);
    } else {
        const double per = (double)c[1] / outer;
        printf("thr2 %s : %.0f cyc per layer-2 row (50 MMAs%s); ideal math 688; issue %.0f  timeout=%d\n", name, per, MODE == 4 ? " + 10 cp" : "", (double)c[0] / outer, st);
    }
}

static int run_thr2()
{
    long long *dC; int *dSt;
    CK(cudaMalloc(&dC, 16)); CK(cudaMalloc(&dSt, 4));
    thr2_case<8, 0>("SS", dC, dSt);   thr2_case<16, 0>("SS", dC, dSt);  thr2_case<32, 0>("SS", dC, dSt);  thr2_case<48, 0>("SS", dC, dSt);
    thr2_case<64, 0>("SS", dC, dSt);  thr2_case<96, 0>("SS", dC, dSt);  thr2_case<128, 0>("SS", dC, dSt); thr2_case<256, 0>("SS", dC, dSt);
    thr2_case<8, 1>("TS", dC, dSt);   thr2_case<16, 1>("TS", dC, dSt);  thr2_case<32, 1>("TS", dC, dSt);  thr2_case<48, 1>("TS", dC, dSt);
    thr2_case<64, 1>("TS", dC, dSt);  thr2_case<96, 1>("TS", dC, dSt);  thr2_case<128, 1>("TS", dC, dSt); thr2_case<256, 1>("TS", dC, dSt);
    thr2_case<16, 2>("cp", dC, dSt);
    thr2_case<16, 3>("L2-mix SS", dC, dSt);
    thr2_case<16, 4>("L2-mix TS+cp", dC, dSt);
    return 0;
}

// ---- A from TMEM: tcgen05.st layout hypothesis and tcgen05.cp ---------------------------
// variant 0: A written by tcgen05.st (lane m, column j = bytes k=4j..4j+3 little endian)
// variant 1: A copied by tcgen05.cp.128x256b from the smem K-major no-swizzle layout
__global__ void k_ts(const int8_t *gA, const int8_t *gB, int *out, int *araw, int variant, int N, int *status)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sA = sm, *sB = sm + A_BYTES;
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < A_BYTES / 16; i += blockDim.x) reinterpret_cast<int4 *>(sA)[i] = reinterpret_cast<const int4 *>(gA)[i];
    for (int i = tid; i < N * 2; i += blockDim.x) {
        const int n = i >> 1, kc = i & 1;
        reinterpret_cast<int4 *>(sB)[kc * N + n] = reinterpret_cast<const int4 *>(gB)[n * 2 + kc];
    }
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&s_tmem, 128); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = s_tmem, ta = tm + 64;       // A operand at columns 64..71
    const int m = warp * 32 + lane;
    if (variant == 0) {
        uint32_t r[8];
        for (int j = 0; j < 8; ++j) {               // k = 4j..4j+3 ; k<16 in plane 0, else plane 1; pixel m (shift 0)
            const int k = 4 * j, plane = k < 16 ? 0 : 1;
            r[j] = *reinterpret_cast<const uint32_t *>(sA + plane * PLANE_B + m * 16 + (k & 15));
        }
        tmem_st_x8(ta + ((uint32_t)(warp * 32) << 16), r);
        tmem_st_wait();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    uint32_t phase = 0;
    if (variant == 1) {
        if (tid == 0) {
            const uint64_t ad = smem_desc(smem_u32(sA), PLANE_B, 128);
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(ta), "l"(ad) : "memory");
            mma_commit(&bar);
        }
        const bool ok = mbar_wait(&bar, phase);
        phase ^= 1;
        fence_after_sync();
        if (!ok && tid == 0) status[0] |= 2;
    }
    {   // dump the raw A columns as the LSU sees them
        uint32_t r[8];
        tmem_ld_x8(ta + ((uint32_t)(warp * 32) << 16), r);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) araw[m * 8 + j] = (int)r[j];
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
        const uint64_t bd = smem_desc(smem_u32(sB), N * 16, 128);
        mma_i8_ts(tm, ta, bd, idesc_i8(128, N, 1, 1), 0);
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, phase);
    fence_after_sync();
    if (!ok) { if (tid == 0) status[0] |= 1; }
    else
        for (int j = 0; j < N / 8; ++j) {
            uint32_t r[8];
            tmem_ld_x8(tm + ((uint32_t)(warp * 32) << 16) + j * 8, r);
            tmem_ld_wait();
            for (int i = 0; i < 8; ++i) out[m * N + j * 8 + i] = (int)r[i];
        }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 128);
}

static int run_ts()
{
    const int N = 48;
    std::vector<int8_t> hA(A_BYTES), hB(B_MAXN * 32);
    uint32_t s = 777;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int8_t)(s >> 24); };
    for (auto &v : hA) v = rnd();
    for (auto &v : hB) v = rnd();
    int8_t *dA, *dB; int *dOut, *dRaw, *dSt;
    CK(cudaMalloc(&dA, A_BYTES)); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dOut, 128 * N * 4)); CK(cudaMalloc(&dRaw, 128 * 8 * 4)); CK(cudaMalloc(&dSt, 4));
    CK(cudaMemcpy(dA, hA.data(), A_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, A_BYTES + B_BYTES));
    for (int variant = 0; variant < 2; ++variant) {
        CK(cudaMemset(dSt, 0, 4)); CK(cudaMemset(dOut, 0xEE, 128 * N * 4)); CK(cudaMemset(dRaw, 0xEE, 128 * 8 * 4));
        k_ts<<<1, 128, A_BYTES + B_BYTES>>>(dA, dB, dOut, dRaw, variant, N, dSt);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ts variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 2; }
        int st; std::vector<int> out(128 * N), raw(128 * 8);
        CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(raw.data(), dRaw, raw.size() * 4, cudaMemcpyDeviceToHost));
        long bad = 0, rawbad = 0;
        for (int m = 0; m < 128; ++m) {
            for (int j = 0; j < 8; ++j) {
                const int k = 4 * j, plane = k < 16 ? 0 : 1;
                uint32_t want; memcpy(&want, &hA[plane * PLANE_B + m * 16 + (k & 15)], 4);
                rawbad += ((uint32_t)raw[m * 8 + j] != want);
            }
            for (int n = 0; n < N; ++n) {
                long acc = 0;
                for (int k = 0; k < 32; ++k) acc += (long)hA[(k < 16 ? 0 : 1) * PLANE_B + m * 16 + (k & 15)] * hB[n * 32 + k];
                bad += out[m * N + n] != (int)acc;
            }
        }
        printf("ts variant %d (%s): status=%d  A-columns-as-expected mismatches=%ld  product mismatches=%ld -> %s\n", variant,
               variant ? "tcgen05.cp.128x256b" : "tcgen05.st 32x32b", st, rawbad, bad, (!st && !bad) ? "PASS" : "FAIL");
        if (rawbad) {
            printf("   raw A lanes 0,1,8,9 columns 0..7:\n");
            for (int m : {0, 1, 8, 9}) { printf("   m=%d:", m); for (int j = 0; j < 8; ++j) printf(" %08x", raw[m * 8 + j]); printf("\n"); }
            printf("   expected:\n");
            for (int m : {0, 1, 8, 9}) { printf("   m=%d:", m); for (int j = 0; j < 8; ++j) { uint32_t w; memcpy(&w, &hA[(j < 4 ? 0 : 1) * PLANE_B + m * 16 + ((4 * j) & 15)], 4); printf(" %08x", w); } printf("\n"); }
        }
    }
    return 0;
}

// ---- tcgen05.ld / st throughput ----------------------------------------------------------
__global__ void k_ldst(int iters, long long *cycles, int *sink)
{
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[16];
        tmem_ld_x16(tm + ((i * 16) & 255) + (warp >> 2) * 256, r);
        tmem_ld_wait();
        acc += r[0] ^ r[15];
    }
    __syncthreads();
    long long t1 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[8];
        for (int j = 0; j < 8; ++j) r[j] = acc + j + i;
        tmem_st_x8(tm + ((i * 8) & 255) + (warp >> 2) * 256, r);
    }
    tmem_st_wait();
    __syncthreads();
    long long t2 = clock64();
    if (tid == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t1; }
    sink[tid] = (int)acc;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_tmem, 512);
}

static int run_ldst()
{
    long long *dC; int *dS;
    CK(cudaMalloc(&dC, 16)); CK(cudaMalloc(&dS, 4096));
    for (int nthreads : {128, 256}) {
        const int iters = 4096;
        for (int rep = 0; rep < 2; ++rep) { k_ldst<<<1, nthreads>>>(iters, dC, dS); CK(cudaDeviceSynchronize()); }
        long long c[2];
        CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
        const double ldB = (double)iters * nthreads * 16 * 4, stB = (double)iters * nthreads * 8 * 4;
        printf("ldst %d threads: tcgen05.ld.x16 %.1f B/clk/SM (%.1f cyc per warp-instr), tcgen05.st.x8 %.1f B/clk/SM\n", nthreads,
               ldB / c[0], (double)c[0] / iters, stB / c[1]);
    }
    return 0;
}

// ---- tcgen05.shift -----------------------------------------------------------------------
__global__ void k_shift(int *out, int *status)
{
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&s_tmem, 32); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = s_tmem;
    uint32_t r[8];
    for (int rep = 0; rep < 2; ++rep) {             // columns 0..15 : value = lane*256 + column
        for (int j = 0; j < 8; ++j) r[j] = (uint32_t)((warp * 32 + lane) * 256 + rep * 8 + j);
        tmem_st_x8(tm + ((uint32_t)(warp * 32) << 16) + rep * 8, r);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
        asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tm) : "memory");
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    fence_after_sync();
    if (!ok && tid == 0) status[0] = 1;
    for (int rep = 0; rep < 2; ++rep) {
        tmem_ld_x8(tm + ((uint32_t)(warp * 32) << 16) + rep * 8, r);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * 16 + rep * 8 + j] = (int)r[j];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 32);
}

static int run_shift()
{
    int *dOut, *dSt;
    CK(cudaMalloc(&dOut, 128 * 16 * 4)); CK(cudaMalloc(&dSt, 4)); CK(cudaMemset(dSt, 0, 4));
    k_shift<<<1, 128>>>(dOut, dSt);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<int> out(128 * 16); int st;
    CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost));
    printf("shift: timeout=%d. After tcgen05.shift.down at column 0, value = srclane*256+col; printing (src lane, col) per lane for cols 0,7,8,15\n", st);
    for (int m : {0, 1, 2, 3, 30, 31, 32, 33, 63, 64, 65, 126, 127}) {
        printf("  lane %3d:", m);
        for (int c : {0, 7, 8, 15}) printf("  (%3d,%2d)", out[m * 16 + c] >> 8, out[m * 16 + c] & 255);
        printf("\n");
    }
    return 0;
}

int main(int argc, char **argv)
{
    const char *t = argc > 1 ? argv[1] : "num";
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("# %s  sm_%d%d  %d SMs  clock %d MHz   test=%s\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate / 1000, t);
    if (!strcmp(t, "num")) return run_num();
    if (!strcmp(t, "thr")) return run_thr();
    if (!strcmp(t, "thr2")) return run_thr2();
    if (!strcmp(t, "ts")) return run_ts();
    if (!strcmp(t, "ldst")) return run_ldst();
    if (!strcmp(t, "shift")) return run_shift();
    printf("unknown test\n");
    return 2;
}
