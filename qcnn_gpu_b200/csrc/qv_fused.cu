// Fused QVRCNN forward for sm_100a: the whole network per column strip, activations resident in
// shared memory, all five dense convolutions (and C1, via an in-smem im2col) as tcgen05.mma
// kind::i8 implicit GEMMs with int32 accumulators in TMEM; HBM sees one luma byte in and one
// reconstructed byte out per pixel.  Replaces, for one frame batch, the whole of
// qvrcnn::forward_blu (inference/qvrcnn.cu:168-242): ppro, 6 x (cudnnConvolutionForward +
// cudnnAddTensor), quantize_out_blu / concat_blu, applyRes_y.
//
// Geometry.  A CTA owns a work unit = (frame, column strip of WT=120 output pixels, row segment)
// and rolls down the rows.  Every activation row lives in smem as [16-channel plane][pixel][16 B]
// with a pixel pitch of PW=136: that is the canonical no-swizzle K-major UMMA operand layout
// (8-pixel x 16-byte core matrices, SBO = 128 B, LBO = plane stride), so a convolution tap (r, s)
// is nothing but a different start address in the A descriptor: row slot r, pixel offset s.
// Buffer pixel p of a strip whose first output column is X0 maps to image column X0 - 8 + p.
//   input  p in [2,134)   a1: p = 4+m   a2: p = 6+m   a3: p = 7+m   out: p = 8+m  (m = MMA row)
//
// Pipeline.  9 warps: warps 0-7 are workers (TMEM->BLU->smem epilogues, im2col for C1, C4 + the
// residual on CUDA cores, global I/O), warp 8 lane 0 issues every MMA.  Iteration i handles
//   MMA side   : L1 for a1 row R1=y0-4+i, L2 for a2 row R1-4, L3 for a3 row R1-7  -> TMEM D*[i&1]
//   worker side: epilogues of iteration i-1's accumulators (a1 row R1-1, a2 row R1-5, a3 row
//                R1-8), im2col of input rows for a1 row R1+1, then C4 for output row R1-9.
// The two sides meet at two mbarriers per iteration (work_done -> MMA may issue i+1,
// mma_done -> workers may read D*[i&1] and overwrite the smem rows iteration i read).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "qv_device.cuh"
#include "qv_fused.h"
#include "qv_tcgen05.cuh"

namespace qv {
namespace {
using namespace tc;

constexpr int WT = 120;                    // output columns per strip
constexpr int PW = 136;                    // pixel pitch of every activation row buffer
constexpr int PLANE = PW * 16;             // bytes of one 16-channel plane of one row
constexpr int PW16 = PW;                   // the same in 16-byte units
constexpr int A1_SLOTS = 6, A2_SLOTS = 4, A3_SLOTS = 4, IN_SLOTS = 16, IN_PITCH = 144;
constexpr int A1_ROW = 4 * PLANE, A2_ROW = 3 * PLANE, A3_ROW = 3 * PLANE;
constexpr int IM_BYTES = 2 * 128 * 16;     // one im2col A operand for C1: [2 K-planes][128 px][16 B]

constexpr int W1_BYTES = 2 * 64 * 16;                  // [2][64][16]
constexpr int W2_INNER = 2 * 48 * 16, W2_OUTER = 2 * 16 * 16;
constexpr int W2_BYTES = 18 * W2_INNER + 32 * W2_OUTER;   // 44032
constexpr int W3_BIG = 2 * 48 * 16, W3_SMALL = 2 * 16 * 16;
constexpr int W3_BYTES = 2 * W3_BIG + 13 * W3_SMALL;      // 9728
constexpr int BIAS_INTS = 64 + 48 + 48;

constexpr int OFF_W1 = 0;
constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
constexpr int OFF_W3 = OFF_W2 + W2_BYTES;
constexpr int OFF_BIAS = OFF_W3 + W3_BYTES;
constexpr int WIMG_BYTES = OFF_BIAS + BIAS_INTS * 4;        // what lives in global memory per model
constexpr int OFF_A1 = (WIMG_BYTES + 127) / 128 * 128;
constexpr int OFF_A2 = OFF_A1 + A1_SLOTS * A1_ROW;
constexpr int OFF_A3 = OFF_A2 + A2_SLOTS * A2_ROW;
constexpr int OFF_IM = OFF_A3 + A3_SLOTS * A3_ROW;
constexpr int OFF_IN = OFF_IM + 2 * IM_BYTES;
constexpr int OFF_PART = OFF_IN + IN_SLOTS * IN_PITCH;     // C4 partial sums, 128 ints
constexpr int OFF_CTRL = OFF_PART + 512;
constexpr int SMEM_BYTES = OFF_CTRL + 64;

constexpr int NWORKER = 256, NTHREADS = NWORKER + 32;
constexpr int PIPE = 13;                   // pipeline depth in rows (first output row appears at i = 13)

// TMEM columns (int32 accumulators), double-buffered by iteration parity
constexpr int TM_D1 = 0, TM_D2 = 128, TM_D3 = 224, TM_COLS = 512;

// Per 16-column accumulator group: everything the requantiser needs.
struct GroupQ {
    int hi;          // FAST: blu + rbias          (upper clamp of acc + bias')
    unsigned M;      // FAST: mul << (24 - shift)  (q = byte 3 of t * M)
    int blu, mul, shift, rbias;   // generic path: the reference formula verbatim
};

struct FusedParams {
    const uint8_t *in;
    uint8_t *out;
    const uint8_t *wimg;
    int n_frames, H, W, nstrips, nseg, seg_rows, n_units;
    GroupQ q1, q22, q21, q31, q32;
    int c4_bias, c4_mul, c4_shift;
    int c4_w[108];                     // [tap][plane][4 words], 4 channels per word
    long long *dbg;                    // optional per-block phase timers (QV_FUSED_PROFILE=1), else null
};

__device__ __forceinline__ int mod_pos(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

// ---- requantise 16 accumulator columns of this thread's pixel and store them as one 16-byte
// ---- channel group of an activation row (mat.cu:262-303 folded into the TMEM epilogue)
template <bool FAST>
__device__ __forceinline__ void requant_store(const uint32_t (&r)[16], const int *bias16, const GroupQ &g, bool valid,
                                              uint8_t *dst)
{
    uint32_t o[4];
    const unsigned Mz = valid ? g.M : 0u;           // FAST: an out-of-image pixel multiplies by 0
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int4 b = reinterpret_cast<const int4 *>(bias16)[v];
        const int bb[4] = {b.x, b.y, b.z, b.w};
        uint32_t q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (FAST) {
                // t = clamp(acc + b + rbias, 0, blu + rbias);  t * (mul << (24-shift)) < 2^31 and its top
                // byte is (t * mul) >> shift  -- see DESIGN.md "epilogue arithmetic"
                const unsigned t = (unsigned)__viaddmin_s32_relu((int)r[4 * v + j], bb[j], g.hi);
                q[j] = t * Mz;
            } else {
                QParam qp{g.blu, g.mul, g.shift, g.rbias};
                q[j] = valid ? ((unsigned)blu_requant((int)r[4 * v + j] + bb[j], qp) << 24) : 0u;
            }
        }
        // gather the four top bytes into one word
        o[v] = __byte_perm(__byte_perm(q[0], q[1], 0x0073), __byte_perm(q[2], q[3], 0x0073), 0x5410);
    }
    *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
}

// Partial C4 dot product of one pixel over (tap, plane) units [U0, U0+NU): 16 channels per unit.
template <int U0, int NU>
__device__ __forceinline__ int c4_partial(const uint8_t *px, const int (&slot3)[3], const FusedParams &P)
{
    int acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
#pragma unroll
    for (int u = U0; u < U0 + NU; ++u) {
        const int t = u / 3, pl = u % 3, r = t / 3, sft = t % 3;
        const uint4 v = *reinterpret_cast<const uint4 *>(px + slot3[r] + pl * PLANE + sft * 16);
        acc0 = __dp4a((int)v.x, P.c4_w[u * 4 + 0], acc0);
        acc1 = __dp4a((int)v.y, P.c4_w[u * 4 + 1], acc1);
        acc2 = __dp4a((int)v.z, P.c4_w[u * 4 + 2], acc2);
        acc3 = __dp4a((int)v.w, P.c4_w[u * 4 + 3], acc3);
    }
    return (acc0 + acc1) + (acc2 + acc3);
}

template <bool FAST>
__global__ void __launch_bounds__(NTHREADS, 1) k_fused(const __grid_constant__ FusedParams P)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    // Two mbarriers per direction, used alternately (event e -> barrier e&1, parity (e>>1)&1): a
    // waiter can then never be lapped, because the second-next completion of the SAME barrier
    // needs the waiter's own arrival in between.
    uint64_t *bar_work = reinterpret_cast<uint64_t *>(sm + OFF_CTRL);        // [2] workers -> MMA
    uint64_t *bar_mma = bar_work + 2;                                        // [2] MMA (tcgen05.commit) -> workers
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(sm + OFF_CTRL + 32);
    int *s_fail = reinterpret_cast<int *>(sm + OFF_CTRL + 36);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = P.H, W = P.W;

    // ---- one-time setup: weights -> smem, barriers, TMEM ---------------------------------
    for (int i = tid; i < WIMG_BYTES / 16; i += NTHREADS)
        reinterpret_cast<uint4 *>(sm)[i] = reinterpret_cast<const uint4 *>(P.wimg)[i];
    for (int i = tid; i < (OFF_CTRL - OFF_A1) / 16; i += NTHREADS)      // finite data everywhere the MMAs may read
        reinterpret_cast<uint4 *>(sm + OFF_A1)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(&bar_work[0], 8);
        mbar_init(&bar_work[1], 8);
        mbar_init(&bar_mma[0], 1);
        mbar_init(&bar_mma[1], 1);
        *s_fail = 0;
        mbar_fence_init();
    }
    if (warp == 8) { tmem_alloc(s_tmem, TM_COLS); tmem_relinquish(); }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = *s_tmem;
    const uint32_t sbase = smem_u32(sm);

    if (warp == 8) {
        // =============================== MMA issuer ========================================
        // The whole warp runs the control flow (so that descriptors stay in uniform registers);
        // one elected lane issues the tcgen05 instructions.
        const bool leader = elect_one();
        {
            uint32_t ev_work = 0, ev_mma = 0;
            long long t_wait = 0, t_issue = 0, tc0 = clock64();
            constexpr uint32_t ID64 = idesc_i8(128, 64), ID48 = idesc_i8(128, 48), ID16 = idesc_i8(128, 16);
            constexpr uint32_t HI_A = (128u >> 4) | (1u << 14);                 // SBO = 128 B, version 1
            const uint32_t a_lo_plane = (uint32_t)PW16 << 16;                    // LBO = one plane
            const uint32_t w1_lo = ((sbase + OFF_W1) >> 4) | ((64u * 16 >> 4) << 16);
            const uint32_t w2_lo48 = ((sbase + OFF_W2) >> 4) | ((48u * 16 >> 4) << 16);
            const uint32_t w2_lo16 = ((sbase + OFF_W2 + 18 * W2_INNER) >> 4) | ((16u * 16 >> 4) << 16);
            const uint32_t w3_lo48 = ((sbase + OFF_W3) >> 4) | ((48u * 16 >> 4) << 16);
            const uint32_t w3_lo16 = ((sbase + OFF_W3 + 2 * W3_BIG) >> 4) | ((16u * 16 >> 4) << 16);
            auto desc = [](uint32_t lo) { return ((uint64_t)HI_A << 32) | lo; };
            auto MMA = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
                if (leader) mma_i8_ss(d, a, b, idesc, acc);
            };
            for (int unit = blockIdx.x; unit < P.n_units; unit += gridDim.x) {
                const int seg = unit % P.nseg;
                const int y0 = seg * P.seg_rows, y1 = min(H, y0 + P.seg_rows);
                const int niter = y1 - y0 + PIPE;
                for (int i = 0; i < niter; ++i) {
                    const int R1 = y0 - 4 + i;
                    if (!mbar_wait(&bar_work[ev_work & 1], (ev_work >> 1) & 1)) { *s_fail = 1; }
                    ++ev_work;
                    fence_after_sync();
                    { const long long t = clock64(); t_wait += t - tc0; tc0 = t; }
                    const uint32_t par = i & 1;
                    // ---- L1: a1 row R1 = im2col[par] x W1 -------------------------------------
                    MMA(tm + TM_D1 + par * 64, desc((((sbase + OFF_IM + par * IM_BYTES) >> 4)) | ((128u * 16 >> 4) << 16)),
                              desc(w1_lo), ID64, 0);
                    // ---- L2: a2 row R2 = R1-4 from a1 rows R2-2..R2+2 ---------------------------
                    {
                        uint32_t row16[5];
#pragma unroll
                        for (int r = 0; r < 5; ++r)
                            row16[r] = ((sbase + OFF_A1 + mod_pos(R1 - 6 + r, A1_SLOTS) * A1_ROW) >> 4) | a_lo_plane;
                        const uint32_t d2 = tm + TM_D2 + par * 48;
                        // inner 3x3 taps: N = 48 ([C2_2 | C2_1]); the first one initialises all 48 columns
#pragma unroll
                        for (int t = 0; t < 9; ++t)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int r = 1 + t / 3, s = 1 + t % 3;
                                MMA(d2, desc(row16[r] + h * 2 * PW16 + 4 + s),
                                          desc(w2_lo48 + ((t * 2 + h) * W2_INNER >> 4)), ID48, (t | h) != 0);
                            }
                        // outer ring of the 5x5: N = 16 (C2_2 only, columns 0..15)
                        int oi = 0;
#pragma unroll
                        for (int t = 0; t < 25; ++t) {
                            const int r = t / 5, s = t % 5;
                            if (r >= 1 && r <= 3 && s >= 1 && s <= 3) continue;
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                MMA(d2, desc(row16[r] + h * 2 * PW16 + 4 + s),
                                          desc(w2_lo16 + ((oi * 2 + h) * W2_OUTER >> 4)), ID16, 1);
                            ++oi;
                        }
                    }
                    // ---- L3: a3 row R3 = R1-7 from a2 rows R3-1..R3+1 ---------------------------
                    {
                        uint32_t row16[3];
#pragma unroll
                        for (int r = 0; r < 3; ++r)
                            row16[r] = (sbase + OFF_A2 + mod_pos(R1 - 8 + r, A2_SLOTS) * A2_ROW) >> 4;
                        const uint32_t d3 = tm + TM_D3 + par * 48;
                        const uint32_t lbo_px = 1u << 16;                       // LBO = 16 B: next pixel, same plane
                        // the two K-steps that contain the centre tap carry C3_2 as well: N = 48
                        MMA(d3, desc((row16[1] + 6 + 1) | a_lo_plane), desc(w3_lo48), ID48, 0);
                        MMA(d3, desc((row16[1] + 2 * PW16 + 6 + 0) | lbo_px), desc(w3_lo48 + (W3_BIG >> 4)), ID48, 1);
                        int bi = 0;
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
#pragma unroll
                            for (int k = 0; k < 5; ++k) {
                                if (r == 1 && (k == 1 || k == 3)) continue;
                                uint32_t a;
                                if (k < 3) a = (row16[r] + 6 + k) | a_lo_plane;                  // (s=k : planes 0,1)
                                else if (k == 3) a = (row16[r] + 2 * PW16 + 6 + 0) | lbo_px;     // (s0 plane2 | s1 plane2)
                                else a = (row16[r] + 2 * PW16 + 6 + 2) | lbo_px;                 // (s2 plane2 | zero weights)
                                MMA(d3, desc(a), desc(w3_lo16 + (bi * W3_SMALL >> 4)), ID16, 1);
                                ++bi;
                            }
                        }
                    }
                    if (leader) mma_commit(&bar_mma[ev_mma & 1]);
                    ++ev_mma;
                    __syncwarp();
                    { const long long t = clock64(); t_issue += t - tc0; tc0 = t; }
                }
            }
            if (P.dbg && leader) { P.dbg[blockIdx.x * 16 + 0] = t_wait; P.dbg[blockIdx.x * 16 + 1] = t_issue; }
        }
    } else {
        // ================================= workers =========================================
        const int q = warp & 3, hh = warp >> 2;
        const int m = q * 32 + lane;                              // this thread's MMA row / pixel
        const uint32_t tm_lane = tm + ((uint32_t)(q * 32) << 16);
        const int *s_bias = reinterpret_cast<const int *>(sm + OFF_BIAS);
        uint32_t ev_work = 0, ev_mma = 0;
        long long tw[6] = {0, 0, 0, 0, 0, 0}, tc0 = clock64();
        auto lap = [&](int k) { const long long t = clock64(); tw[k] += t - tc0; tc0 = t; };
        auto worker_bar = []() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
        for (int unit = blockIdx.x; unit < P.n_units; unit += gridDim.x) {
            const int seg = unit % P.nseg, strip = (unit / P.nseg) % P.nstrips, f = unit / (P.nseg * P.nstrips);
            const int X0 = strip * WT;
            const int y0 = seg * P.seg_rows, y1 = min(H, y0 + P.seg_rows);
            const int niter = y1 - y0 + PIPE;
            const uint8_t *inf = P.in + (size_t)f * H * W;
            uint8_t *outf = P.out + (size_t)f * H * W;
            const int col_in = X0 - 8 + tid;                      // input-ring byte this thread loads (tid < PW)
            const bool in_col_ok = tid < PW && col_in >= 0 && col_in < W;
            auto load_in = [&](int row) -> unsigned {
                return (in_col_ok && row >= 0 && row < H) ? (unsigned)inf[(size_t)row * W + col_in] : 128u;
            };
            auto store_in = [&](int row, unsigned v) {
                if (tid < PW) sm[OFF_IN + mod_pos(row, IN_SLOTS) * IN_PITCH + tid] = (uint8_t)v;
            };
            // im2col of input rows R-2..R+2 for a1 row R (threads of warps 4-7: pixel m)
            auto im2col = [&](int R, int buf) {
                uint32_t A[5], bcol = 0, b4 = 0;
                const int p = m + 2, o8 = (p & 3) * 8;
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    const uint32_t *rowp = reinterpret_cast<const uint32_t *>(sm + OFF_IN + mod_pos(R - 2 + r, IN_SLOTS) * IN_PITCH) + (p >> 2);
                    const uint32_t w0 = rowp[0], w1 = rowp[1];
                    A[r] = __funnelshift_r(w0, w1, o8);           // bytes p..p+3   (taps s = 0..3)
                    const uint32_t b = (w1 >> o8) & 0xffu;        // byte  p+4      (tap  s = 4)
                    if (r < 4) bcol |= b << (8 * r); else b4 = b;
                }
                // x - 128 as int8 == x ^ 0x80 (cnn.cu:450); K order: k = 4r+s (s<4), 20+r (s=4), 25..31 zero weights
                uint8_t *dst = sm + OFF_IM + buf * IM_BYTES + m * 16;
                *reinterpret_cast<uint4 *>(dst) = make_uint4(A[0] ^ 0x80808080u, A[1] ^ 0x80808080u, A[2] ^ 0x80808080u, A[3] ^ 0x80808080u);
                *reinterpret_cast<uint4 *>(dst + 128 * 16) = make_uint4(A[4] ^ 0x80808080u, bcol ^ 0x80808080u, b4 ^ 0x80u, 0u);
            };

            // ---- prologue: input rows for a1 rows y0-4 and y0-3, im2col of the first ------------
            {
                const int R1 = y0 - 4;
                for (int r = R1 - 2; r <= R1 + 3; ++r) store_in(r, load_in(r));
                worker_bar();
                if (hh == 1) im2col(R1, 0);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_work[ev_work & 1]);
                ++ev_work;
            }
            for (int i = 0; i < niter; ++i) {
                const int R1 = y0 - 4 + i;
                const unsigned in_next = load_in(R1 + 4);         // prefetch; stored at the end of the iteration
                if (i >= 1) {
                    if (!mbar_wait(&bar_mma[ev_mma & 1], (ev_mma >> 1) & 1)) { *s_fail = 1; }
                    ++ev_mma;
                    fence_after_sync();
                    lap(0);
                    const uint32_t par = (i - 1) & 1;
                    // ---- epilogues: a1 row R1-1 (D1, 4 groups), a2 row R1-5 (D2: group 0 = C2_2 -> plane 2,
                    //      groups 1,2 = C2_1 -> planes 0,1), a3 row R1-8 (D3: group 0 = C3_1, 1,2 = C3_2).
                    //      Each warp half (hh) takes 5 of the 10 sixteen-column groups; all TMEM loads first.
                    {
                        const int row1 = R1 - 1, row2 = R1 - 5, row3 = R1 - 8;
                        const bool v1 = row1 >= 0 && row1 < H && X0 - 4 + m >= 0 && X0 - 4 + m < W;
                        const bool v2 = row2 >= 0 && row2 < H && X0 - 2 + m >= 0 && X0 - 2 + m < W;
                        const bool v3 = row3 >= 0 && row3 < H && X0 - 1 + m >= 0 && X0 - 1 + m < W;
                        uint8_t *dst1 = sm + OFF_A1 + mod_pos(row1, A1_SLOTS) * A1_ROW + (4 + m) * 16;
                        uint8_t *dst2 = sm + OFF_A2 + mod_pos(row2, A2_SLOTS) * A2_ROW + (6 + m) * 16;
                        uint8_t *dst3 = sm + OFF_A3 + mod_pos(row3, A3_SLOTS) * A3_ROW + (7 + m) * 16;
                        const uint32_t d1 = tm_lane + TM_D1 + par * 64, d2 = tm_lane + TM_D2 + par * 48, d3 = tm_lane + TM_D3 + par * 48;
                        uint32_t ra[16], rb[16], rc[16], rd[16], re[16];
                        if (hh == 0) {
                            tmem_ld_x16(d1 + 0, ra); tmem_ld_x16(d1 + 16, rb);
                            tmem_ld_x16(d2 + 0, rc); tmem_ld_x16(d2 + 16, rd);
                            tmem_ld_x16(d3 + 0, re);
                            tmem_ld_wait();
                            requant_store<FAST>(ra, s_bias + 0, P.q1, v1, dst1 + 0 * PLANE);
                            requant_store<FAST>(rb, s_bias + 16, P.q1, v1, dst1 + 1 * PLANE);
                            requant_store<FAST>(rc, s_bias + 64 + 0, P.q22, v2, dst2 + 2 * PLANE);
                            requant_store<FAST>(rd, s_bias + 64 + 16, P.q21, v2, dst2 + 0 * PLANE);
                            requant_store<FAST>(re, s_bias + 112 + 0, P.q31, v3, dst3 + 0 * PLANE);
                        } else {
                            tmem_ld_x16(d1 + 32, ra); tmem_ld_x16(d1 + 48, rb);
                            tmem_ld_x16(d2 + 32, rc);
                            tmem_ld_x16(d3 + 16, rd); tmem_ld_x16(d3 + 32, re);
                            tmem_ld_wait();
                            requant_store<FAST>(ra, s_bias + 32, P.q1, v1, dst1 + 2 * PLANE);
                            requant_store<FAST>(rb, s_bias + 48, P.q1, v1, dst1 + 3 * PLANE);
                            requant_store<FAST>(rc, s_bias + 64 + 32, P.q21, v2, dst2 + 1 * PLANE);
                            requant_store<FAST>(rd, s_bias + 112 + 16, P.q32, v3, dst3 + 1 * PLANE);
                            requant_store<FAST>(re, s_bias + 112 + 32, P.q32, v3, dst3 + 2 * PLANE);
                        }
                    }
                }
                lap(1);
                if (i + 1 < niter) {
                    if (hh == 1) im2col(R1 + 1, (i + 1) & 1);
                    fence_proxy_async_smem();                     // st.shared above -> visible to the tensor core
                    fence_before_sync();                          // tcgen05.ld above ordered before the MMAs that overwrite D
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_work[ev_work & 1]);
                    ++ev_work;
                }
                lap(2);
                store_in(R1 + 4, in_next);
                worker_bar();                                     // a3 row + input ring visible to every worker
                lap(3);
                // ---- C4 (48 -> 1, 3x3) + applyRes_y for output row R1-9 (cnn.cu:507-523).  The 27 (tap, plane)
                //      units of a pixel are split 13 / 14 between the two warp halves; partial sums meet in smem.
                const int R4 = R1 - 9;
                if (R4 >= y0 && R4 < y1) {
                    const uint8_t *px = sm + OFF_A3 + (7 + m) * 16;
                    int slot3[3];
#pragma unroll
                    for (int r = 0; r < 3; ++r) slot3[r] = mod_pos(R4 - 1 + r, A3_SLOTS) * A3_ROW;
                    const int part = hh == 0 ? c4_partial<0, 13>(px, slot3, P) : c4_partial<13, 14>(px, slot3, P);
                    int *s_part = reinterpret_cast<int *>(sm + OFF_PART);
                    if (hh == 1) s_part[m] = part;
                    worker_bar();
                    if (hh == 0 && m < WT && X0 + m < W) {
                        const int x = sm[OFF_IN + mod_pos(R4, IN_SLOTS) * IN_PITCH + 8 + m];
                        outf[(size_t)R4 * W + X0 + m] =
                            (uint8_t)residual_apply(part + s_part[m] + P.c4_bias, x, P.c4_mul, P.c4_shift);
                    }
                }
                lap(4);
            }
            // drain: the MMAs of the last iteration still read smem / write TMEM
            if (!mbar_wait(&bar_mma[ev_mma & 1], (ev_mma >> 1) & 1)) { *s_fail = 1; }
            ++ev_mma;
            fence_after_sync();
            worker_bar();
            lap(5);
        }
        if (P.dbg && (tid == 0 || tid == 128))
            for (int k = 0; k < 6; ++k) P.dbg[blockIdx.x * 16 + 2 + (tid >> 7) * 6 + k] = tw[k];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tm, TM_COLS);
    if (tid == 0 && *s_fail) printf("qv fused kernel: mbarrier wait timed out in block %d\n", blockIdx.x);
}

}  // namespace

// =================================== host side ==========================================
struct FusedModel {
    uint8_t *d_wimg = nullptr;
    FusedParams proto{};
    bool fast = true;
    int sm_count = 148;
};

static void put_chunk(uint8_t *blk, int N, int kchunk, int n, const int8_t *src16) { memcpy(blk + ((size_t)kchunk * N + n) * 16, src16, 16); }

FusedModel *fused_upload(const ModelHost &m, cudaStream_t st)
{
    std::vector<uint8_t> img(WIMG_BYTES, 0);
    auto W = [&](int l, int k, int c, int r, int s) -> int8_t {
        const LayerShape &sh = kLayers[l];
        return m.L[l].w[(((size_t)k * sh.cin + c) * sh.k + r) * sh.k + s];
    };
    // ---- W1: B[n][k], k = 4r+s (s<4) | 20+r (s=4) | zero --------------------------------------
    for (int n = 0; n < 64; ++n) {
        int8_t k32[32] = {0};
        for (int r = 0; r < 5; ++r) {
            for (int s = 0; s < 4; ++s) k32[4 * r + s] = W(QV_C1, n, 0, r, s);
            k32[20 + r] = W(QV_C1, n, 0, r, 4);
        }
        put_chunk(img.data() + OFF_W1, 64, 0, n, k32);
        put_chunk(img.data() + OFF_W1, 64, 1, n, k32 + 16);
    }
    // ---- W2: 18 inner blocks (N=48: rows 0..15 = C2_2, 16..47 = C2_1) then 32 outer blocks (N=16: C2_2)
    {
        int oi = 0;
        for (int t = 0; t < 25; ++t) {
            const int r = t / 5, s = t % 5;
            const bool inner = r >= 1 && r <= 3 && s >= 1 && s <= 3;
            for (int h = 0; h < 2; ++h) {
                uint8_t *blk;
                int N;
                if (inner) { const int ti = (r - 1) * 3 + (s - 1); blk = img.data() + OFF_W2 + (ti * 2 + h) * W2_INNER; N = 48; }
                else { blk = img.data() + OFF_W2 + 18 * W2_INNER + (oi * 2 + h) * W2_OUTER; N = 16; }
                for (int j = 0; j < 2; ++j) {
                    int8_t c16[16];
                    for (int n = 0; n < 16; ++n) {
                        for (int b = 0; b < 16; ++b) c16[b] = W(QV_C2_2, n, 32 * h + 16 * j + b, r, s);
                        put_chunk(blk, N, j, n, c16);
                    }
                    if (inner)
                        for (int n = 0; n < 32; ++n) {
                            for (int b = 0; b < 16; ++b) c16[b] = W(QV_C2_1, n, 32 * h + 16 * j + b, r - 1, s - 1);
                            put_chunk(blk, N, j, 16 + n, c16);
                        }
                }
            }
            if (!inner) ++oi;
        }
    }
    // ---- W3: K-steps are pairs of (tap, 16-channel plane) units of a2 ----------------------------
    {
        // unit (r, s, pl) -> 16 weights per output row; C3_1 rows 0..15, C3_2 rows 16..47 (centre tap only)
        auto fill = [&](uint8_t *blk, int N, int j, int r, int s, int pl) {
            int8_t c16[16];
            for (int n = 0; n < 16; ++n) {
                for (int b = 0; b < 16; ++b) c16[b] = W(QV_C3_1, n, 16 * pl + b, r, s);
                put_chunk(blk, N, j, n, c16);
            }
            if (N == 48 && r == 1 && s == 1)
                for (int n = 0; n < 32; ++n) {
                    for (int b = 0; b < 16; ++b) c16[b] = W(QV_C3_2, n, 16 * pl + b, 0, 0);
                    put_chunk(blk, N, j, 16 + n, c16);
                }
        };
        uint8_t *big = img.data() + OFF_W3;
        fill(big, 48, 0, 1, 1, 0); fill(big, 48, 1, 1, 1, 1);                       // (r1,s1: planes 0,1)
        fill(big + W3_BIG, 48, 0, 1, 0, 2); fill(big + W3_BIG, 48, 1, 1, 1, 2);     // (r1,s0 plane 2 | r1,s1 plane 2)
        uint8_t *small = img.data() + OFF_W3 + 2 * W3_BIG;
        int bi = 0;
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 5; ++k) {
                if (r == 1 && (k == 1 || k == 3)) continue;
                uint8_t *blk = small + bi * W3_SMALL;
                if (k < 3) { fill(blk, 16, 0, r, k, 0); fill(blk, 16, 1, r, k, 1); }
                else if (k == 3) { fill(blk, 16, 0, r, 0, 2); fill(blk, 16, 1, r, 1, 2); }
                else { fill(blk, 16, 0, r, 2, 2); /* second half stays zero */ }
                ++bi;
            }
    }
    // ---- requantiser constants ---------------------------------------------------------------------
    FusedModel *fm = new FusedModel();
    auto mkq = [&](int l, GroupQ &g) -> bool {
        const LayerHost &L = m.L[l];
        g.blu = L.blu; g.mul = L.mul; g.shift = L.shift; g.rbias = (1 << (L.shift - 1)) / L.mul;
        g.hi = L.blu + g.rbias;
        const long long top = ((long long)L.blu + g.rbias) * L.mul;
        const bool ok = L.shift <= 24 && L.mul < (1ll << L.shift) && top < (1ll << 31) && (top >> L.shift) == 127 && g.hi < (1 << 30);
        g.M = ok ? (unsigned)((unsigned long long)L.mul << (24 - L.shift)) : 0u;
        return ok;
    };
    FusedParams &P = fm->proto;
    bool fast = true;
    fast &= mkq(QV_C1, P.q1); fast &= mkq(QV_C2_2, P.q22); fast &= mkq(QV_C2_1, P.q21);
    fast &= mkq(QV_C3_1, P.q31); fast &= mkq(QV_C3_2, P.q32);
    fm->fast = fast;
    int *bias = reinterpret_cast<int *>(img.data() + OFF_BIAS);
    auto addb = [&](int l, int dst0, const GroupQ &g) {
        for (int k = 0; k < kLayers[l].cout; ++k) bias[dst0 + k] = m.L[l].b[k] + (fast ? g.rbias : 0);
    };
    addb(QV_C1, 0, P.q1);
    addb(QV_C2_2, 64, P.q22); addb(QV_C2_1, 64 + 16, P.q21);       // D2 order: [C2_2 | C2_1]
    addb(QV_C3_1, 112, P.q31); addb(QV_C3_2, 112 + 16, P.q32);     // D3 order: [C3_1 | C3_2]
    P.c4_bias = m.L[QV_C4].b[0]; P.c4_mul = m.L[QV_C4].mul; P.c4_shift = m.L[QV_C4].shift;
    for (int t = 0; t < 9; ++t)
        for (int pl = 0; pl < 3; ++pl)
            for (int j = 0; j < 4; ++j) {
                unsigned v = 0;
                for (int b = 0; b < 4; ++b) v |= (unsigned)(uint8_t)W(QV_C4, 0, 16 * pl + 4 * j + b, t / 3, t % 3) << (8 * b);
                P.c4_w[(t * 3 + pl) * 4 + j] = (int)v;
            }
    cudaError_t e = cudaMalloc(&fm->d_wimg, WIMG_BYTES);
    if (e == cudaSuccess) e = cudaMemcpyAsync(fm->d_wimg, img.data(), WIMG_BYTES, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    int dev = 0, sms = 148;
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
        set_error("fused_upload: %s", cudaGetErrorString(e));
        cudaFree(fm->d_wimg);
        delete fm;
        return nullptr;
    }
    fm->sm_count = sms;
    P.wimg = fm->d_wimg;
    return fm;
}

void fused_free(FusedModel *fm)
{
    if (!fm) return;
    cudaFree(fm->d_wimg);
    delete fm;
}

cudaError_t fused_forward(const FusedModel *fm, const uint8_t *d_in, uint8_t *d_out, int n, int H, int W, cudaStream_t st,
                          long long *launches)
{
    if (n <= 0) return cudaSuccess;
    FusedParams P = fm->proto;
    P.in = d_in; P.out = d_out; P.n_frames = n; P.H = H; P.W = W;
    P.nstrips = (W + WT - 1) / WT;
    // Row segments: only when whole-column strips alone cannot fill the SMs about twice over; every
    // segment pays PIPE extra iterations, so never cut below 32 rows.
    const long long cols = (long long)n * P.nstrips;
    int nseg = 1;
    if (cols < 2ll * fm->sm_count) {
        nseg = (int)((2ll * fm->sm_count + cols - 1) / cols);
        nseg = std::max(1, std::min(nseg, (H + 31) / 32));
    }
    P.seg_rows = (H + nseg - 1) / nseg;
    P.nseg = (H + P.seg_rows - 1) / P.seg_rows;
    const long long units = cols * P.nseg;
    if (units > 0x7fffffffll) return cudaErrorInvalidValue;
    P.n_units = (int)units;
    const int grid = (int)std::min<long long>(units, fm->sm_count);
    const bool prof = getenv("QV_FUSED_PROFILE") != nullptr;
    P.dbg = nullptr;
    if (prof && cudaMalloc(&P.dbg, (size_t)grid * 16 * sizeof(long long)) != cudaSuccess) P.dbg = nullptr;
    if (P.dbg) cudaMemsetAsync(P.dbg, 0, (size_t)grid * 16 * sizeof(long long), st);
    if (fm->fast) k_fused<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(P);
    else k_fused<false><<<grid, NTHREADS, SMEM_BYTES, st>>>(P);
    if (launches) *launches += 1;
    cudaError_t e = cudaGetLastError();
    if (P.dbg) {
        std::vector<long long> h((size_t)grid * 16);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), P.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(P.dbg);
        double a[14] = {0};
        for (int b = 0; b < grid; ++b) for (int k = 0; k < 14; ++k) a[k] += (double)h[(size_t)b * 16 + k] / grid;
        const double iters = (double)P.n_units / grid * (P.seg_rows + PIPE);
        fprintf(stderr, "[qv fused profile] units=%d grid=%d iters/block~%.0f | cycles per iteration: MMA warp wait=%.0f issue=%.0f | "
                "worker w0 (C4 side): wait_mma=%.0f epi=%.0f im2col+arrive=%.0f bar=%.0f c4=%.0f drain=%.0f | "
                "worker w4 (im2col side): wait_mma=%.0f epi=%.0f im2col+arrive=%.0f bar=%.0f c4=%.0f drain=%.0f\n",
                P.n_units, grid, iters, a[0] / iters, a[1] / iters, a[2] / iters, a[3] / iters, a[4] / iters, a[5] / iters, a[6] / iters,
                a[7] / iters, a[8] / iters, a[9] / iters, a[10] / iters, a[11] / iters, a[12] / iters, a[13] / iters);
    }
    return e;
}

}  // namespace qv
