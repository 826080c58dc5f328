// Thin inline-PTX wrappers for the sm_100a features the fused kernel uses: mbarrier, the
// generic->async proxy fence, TMEM allocation, tcgen05.mma kind::i8 (SS and TS forms),
// tcgen05.commit, tcgen05.ld/st.  Hand-written; no CUTLASS/CuTe dependency.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qv {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp (elect.sync).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Same with a suspend-time hint (ns): the thread sleeps in hardware until the phase completes or the hint expires,
// instead of coming back to re-issue the poll -- fewer instructions of a waiting warp in everybody else's way.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t *bar, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// Bounded wait: returns false (instead of hanging the GPU) if the phase does not complete within ~0.2 s (20 000 polls of
// up to 10 us each).  (Both a longer bound and a non-unrolled poll loop were measured: the code the compiler emits for this
// cold loop shifts the schedule of the hot code around it, 6.06 against 5.95 ms -- profiles/r2_kernel_ab_wait_loop.log.)  The waiter parks in hardware with a suspend-time hint.  A parked waiter sees an arrive later than one
// that spins on try_wait (398 against 194 cycles, tools/umma_probe3.cu pingpong, profiles/r1_probe3_pingpong.log), and yet
// the fused kernel is 10 % FASTER with parked waiters (6.60 against 7.25 ms, same box, same call:
// profiles/experiments/r1_wait_spin_vs_hint_and_three_workers_ab.log).  It is the MMA warp's own spinning that costs: with
// only that warp spinning 6.97 ms, with only the workers spinning 6.62 ms (r1_split_commit_v2_and_hybrid_wait_ab.log).
// QV_WAIT_SPIN builds the spinning form.
#ifdef QV_WAIT_SPIN
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int max_polls = 4000000)
{
    for (int n = 0; n < max_polls; ++n)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
#else
#ifndef QV_WAIT_HINT_NS
#define QV_WAIT_HINT_NS 10000
#endif
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int max_polls = 200000000 / QV_WAIT_HINT_NS)
{
    if (mbar_try_wait(bar, parity)) return true;
#ifdef QV_WAIT_UNROLL1
#pragma unroll 1
#endif
    for (int n = 0; n < max_polls; ++n)
        if (mbar_try_wait_hint(bar, parity, (uint32_t)QV_WAIT_HINT_NS)) return true;
    return false;
}
#endif

// Generic-proxy smem writes (st.shared) -> visible to the async proxy (tcgen05.mma / tcgen05.cp reads).
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -----------------------------------------------------------------------------
// Whole-warp (sync.aligned).  ncols: power of two in [32, 512].  The base address lands in *smem_dst.
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------
// Shared-memory matrix descriptor, no swizzle ("interleave"), K-major: the matrix is a grid of
// 8-row x 16-byte core matrices, each 128 contiguous bytes.  Core matrix (i, j) (i = 8-row group
// along M/N, j = 16-byte step along K) starts at  start + i*sbo + j*lbo.   version field = 1.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::i8: D = S32, A/B signed (1) or unsigned (0) 8-bit, both K-major.
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N, int a_signed = 1, int b_signed = 1)
{
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// ---- MMA ------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues on behalf of the CTA.
__device__ __forceinline__ void mma_i8_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, with an A-collector hint (measured: tools/umma_probe3.cu, profiles/r1_probe3_collector.log).  FILL keeps
// the A tile in the tensor core's collector after use; USE / LASTUSE take A from the collector instead of
// reading shared memory again (the descriptor must still name the same tile: if the collector was
// invalidated by an intervening MMA, the hardware re-reads it from there).
enum Collector { COL_DISCARD = 0, COL_FILL = 1, COL_USE = 2, COL_LASTUSE = 3 };
template <int COL>
__device__ __forceinline__ void mma_i8_ss_col(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    if (COL == COL_DISCARD) mma_i8_ss(d_tmem, adesc, bdesc, idesc, accumulate);
    if (COL == COL_FILL)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    if (COL == COL_USE)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    if (COL == COL_LASTUSE)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from TMEM (128 lanes x K/4 columns).
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05 async ops of this thread complete.
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM <-> registers ---------------------------------------------------------------
// 32x32b: thread t of warp w reads lane 32*(w%4)+t; .xN = N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld_x1(uint32_t taddr)
{
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    return r;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace tc
}  // namespace qv
