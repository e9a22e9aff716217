// gabor_tc.cu — filter bank with the ROW pass on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Reference: none (segmenter slot, BSD_metrics/script.py:30; spec in DESIGN.md section 3).
//
// Why.  ncu on the FP32-pipe kernel (gabor.cu) shows the stage is compute bound and that its row pass, whose
// sliding window runs along the lane-private direction, only reaches ~40 % of the FMA pipe.
// The row pass is a small dense GEMM per 32-column strip,
//     T[r][(c, re|im)] = sum_k  I[r][x0 - hmax + k] * Gt[(c, re|im)][k],      Gt[(c,p)][k] = g_p[c + h + hmax - k]
// (a Toeplitz matrix of the job's complex row taps, band waste (32 + 2h) / (2h + 1) <= 1.3 for the wide scales),
// and for the rgb colour space its data operand is EXACT in bf16: the planes hold the u8 pixel value (0..255,
// 8 significant bits) and the 1/255 is folded into the taps.  Only the taps need splitting: g/255 = g1 + g2 + g3
// with three bf16 terms (24 significant bits), i.e. three accumulating MMAs per K step with fp32 accumulation
// in tensor memory.  The column pass keeps running on the FP32 pipe (it sits at ~76 % of its FMA bound and its
// operand T would need a hi/lo split that does not fit in shared memory, profiles/r02_gabor_tc.md), but it now
// has the SM's CUDA cores to itself: the tensor cores produce the next orientation job's T while the column
// pass of the current one runs.
//
// One CTA (512 threads, 1 CTA/SM) = one 32-column strip of one (image, channel, scale), all orientation jobs:
//   warp 14, one lane   TMA producer: per (job, 128-row block, 64-column K atom) one box of the bf16 plane
//                       (A operand, 128 rows x 128 B, SWIZZLE_128B) and three boxes of the tap table (B operand,
//                       64 n x 128 B each) into a 3-stage ring; completion on mbarriers.  The plane tile is the
//                       same for every job of the scale, so it is staged from L2, never re-read from HBM.
//   warp 15, one lane   issues tcgen05.mma.cta_group::1.kind::f16 (M = 128 rows, N = 64 = 32 columns x re|im,
//                       K = 16 per instruction) into one of two TMEM accumulators (3 row blocks x 64 columns
//                       each); tcgen05.commit releases ring stages and publishes finished accumulators.
//   warps 0..13         wait for the accumulator, move it TMEM -> registers -> shared memory as the complex
//                       intermediate T[row][33] (tcgen05.ld 32x32b), hand the accumulator back, then run the
//                       register-blocked FFMA2 column pass of gabor_dev.cuh and the magnitude epilogue.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "async.cuh"
#include "gabor_dev.cuh"

namespace gcis {

using namespace gbdev;

namespace {

#ifndef TC_COLW_N
#define TC_COLW_N 16
#endif
constexpr int TC_COLW = TC_COLW_N;             // column-pass warps: 4 per SM sub-partition.  With the complex column taps in uniform
                                               // registers (ConstTaps) the kernel needs 62 registers per thread, so the warp count
                                               // is free: 12 / 16 / 20 / 24 / 28 warps -> 12.9 / 11.5 / 11.9 / 12.3 / 12.4 ms per 200
                                               // images (with 12 warps the compiler keeps the taps in ordinary registers: 116).
                                               // Before, with the taps in ordinary registers (128 per thread): 13.5 ms at 12 warps.
constexpr int TC_COLT = TC_COLW * 32;
constexpr int TC_TMA_WARP = TC_COLW, TC_MMA_WARP = TC_COLW + 1;
constexpr int TC_THREADS = TC_COLT + 64;
constexpr int TC_STAGES = 3;                   // deepest TMA ring; a plan whose T needs the room runs with 2 (TcParams::stages)
constexpr int TC_MROWS = 128;                  // rows per MMA (M)
constexpr int TC_N = 2 * GB_TW;                // 64: (column, re|im)
constexpr int TC_KATOM = 64;                   // bf16 elements per 128-byte swizzle atom row
constexpr int TC_A_BYTES = TC_MROWS * 128;     // 16 KB
constexpr int TC_B_BYTES = TC_N * 128;         // 8 KB per split term
constexpr int TC_SPLIT = 3;
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_SPLIT * TC_B_BYTES;   // 40 KB
constexpr int TC_MAX_RB = 4;                   // row blocks per accumulator: T holds up to 512 rows (a 481-row portrait image in one tile)
constexpr int TC_ACC_COLS = TC_MAX_RB * TC_N;  // 256 TMEM columns per accumulator
constexpr int TC_TMEM_COLS = 512;              // allocation: two accumulators
constexpr int TC_XFER_RB = TC_COLW / 4;        // row blocks moved TMEM -> shared memory at a time (one warp per lane quadrant and block)
static_assert(TC_COLW % 4 == 0 && TC_COLW >= 4, "a column warp reads the TMEM lane quadrant warp % 4");

constexpr int TC_CTAP_CAP = 6144;              // complex column taps the constant table holds (48 KB; the 4 x 6 bank needs 1440)
constexpr int TC_CTAP_JOBS = 8;                // jobs per scale the constant table covers

struct TcParams {
    GaborParams g;                // shapes, feature tensor, FP32 tap table (column taps), scales
    int ksteps[GB_MAX_SCALES];    // K steps of 16 per scale: ceil((32 + 2 hmax_s + kshift_s) / 16)
    int kshift[GB_MAX_SCALES];    // (P - hmax_s) mod 8: TMA boxes must start on a 16-byte boundary of the plane row, so the K
                                  // origin of a strip is moved left to the previous multiple of 8 columns
    int table_row0[GB_MAX_SCALES];// first row of scale s in the B-operand table
    int hmax[GB_MAX_SCALES], n_jobs[GB_MAX_SCALES];   // copies of the scale table for the single-thread roles
    int max_jobs;                 // jobs of the scale with the most jobs (column-tap slots in shared memory)
    int plane_rows;               // rows of the bf16 plane tensor = B * C * H
    int stages;                   // depth of the TMA ring (2 or 3)
    // Complex column taps live in CONSTANT memory (c_ctaps below), in the layout stage_taps() gives them in shared memory
    // (slot of scale s, job ji at complex index ctap_base[s] + ji * ctap_slot[s]; tap j of a job at 1 + GB_TAP_PAD + j).
    int ctap_base[GB_MAX_SCALES], ctap_slot[GB_MAX_SCALES];
    int ctap_w0[GB_MAX_SCALES * TC_CTAP_JOBS];   // complex index of the block-0 window of (scale, job); -1: real column taps
    // half-width and kind of every job once more, here, so that the column warps derive their loop bounds and tap
    // indices from parameters, block indices and loop counters only: the compiler then keeps them in uniform registers
    int job_h[GB_MAX_SCALES * TC_CTAP_JOBS], job_kind[GB_MAX_SCALES * TC_CTAP_JOBS];   // kind: bit 0 complex row taps, bit 1 complex column taps
};

// The complex column taps of the bank the kernel runs (uploaded by gabor_tc_launch when the bank changes).  The sweeps read
// them with uniform constant loads (LDCU.64) into UNIFORM registers and FFMA2 takes the pair as a uniform operand: per
// packed FMA the register file delivers one accumulator pair and one scalar instead of two pairs (a stream of FFMA2 with
// two register-pair operands sustains 0.75 of the FP32 peak, benchmarks/ffma2_forms.cu), the taps cost no shared-memory
// loads and no ordinary registers (128 -> 62 registers per thread).  (From the kernel's parameter space instead the
// compiler kept the taps in ordinary registers.)
__constant__ float2 c_ctaps[TC_CTAP_CAP];

// instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = 64 (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_MROWS >> 4) << 24);

// shared-memory matrix descriptor, K-major operand in the 128-byte swizzle layout the TMA boxes land in:
// 8 rows x 128 B per swizzle atom, 1024 B between 8-row groups (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);          // start address
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: lane i of the warp receives TMEM lane (quadrant base + i)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}

// u8 pixel value -> bf16 planes with horizontal reflect padding: [B][3][H][Wp16]
__global__ void colour_pad16_kernel(const uint8_t *__restrict__ img, __nv_bfloat16 *__restrict__ planes, int B, int H,
                                    int W, int P, int Wp)
{
    const long long total = (long long)B * H * Wp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cp = (int)(i % Wp);
        const long long br = i / Wp;
        const int r = (int)(br % H);
        const int b = (int)(br / H);
        const int c = reflect_index(cp - P, W);
        const uint8_t *px = img + (((size_t)b * H + r) * W + c) * 3;
        const size_t plane = (size_t)H * Wp;
        __nv_bfloat16 *dst = planes + (size_t)b * 3 * plane + (size_t)r * Wp + cp;
        dst[0] = __ushort2bfloat16_rn(px[0]);
        dst[plane] = __ushort2bfloat16_rn(px[1]);
        dst[2 * plane] = __ushort2bfloat16_rn(px[2]);
    }
}

#ifdef TC_TRACE   // timing experiment only: per-CTA cycles per phase of the column warps (thread 0)
constexpr int TC_TR_CTAS = 8192;
__device__ long long tc_trace_buf[TC_TR_CTAS][8];
#define TC_TR_DECL long long tr_t = clock64(), tr_acc[6] = {0, 0, 0, 0, 0, 0}; const long long tr_t0 = tr_t
#define TC_TR_ADD(slot) do { const long long n_ = clock64(); tr_acc[slot] += n_ - tr_t; tr_t = n_; } while (0)
#else
#define TC_TR_DECL do { } while (0)
#define TC_TR_ADD(slot) do { } while (0)
#endif

// One work item = one 32-column strip of one (image, channel, scale): decoded identically by every role.
struct TcItem {
    int s, b, c, x0, y0, th, hmax, n_jobs, lo, nsrc, n_rb, ksteps, n_atoms;
};

__device__ __forceinline__ TcItem tc_decode(const TcParams &Q, int item)
{
    const GaborParams &P = Q.g;
    TcItem it;
    int range = 0;
    while (range + 1 < P.S && item >= P.first_block[range + 1]) ++range;   // ranges are ordered widest scale first
    it.s = P.order[range];
    int rem = item - P.first_block[range];
    const int nvt = P.n_vt[it.s];
    // A last strip of a few columns (481 = 15 x 32 + 1) costs the column warps ~5 % of a full one (col_pass's lanes-on-
    // rows path): those items come LAST in the scale's range, so the round-robin deal gives every CTA its share of them.
    const int thin = (P.n_strips > 1 && P.W - (P.n_strips - 1) * GB_TW <= gbdev::GB_THIN_COLS) ? 1 : 0;
    const int nfat = P.n_strips - thin, n_fat_items = P.B * P.C * nfat * nvt;
    int strip;
    const int vt = rem % nvt;
    if (rem < n_fat_items) {
        rem /= nvt;
        strip = rem % nfat; rem /= nfat;
    } else {
        rem = (rem - n_fat_items) / nvt;
        strip = P.n_strips - 1;
    }
    it.c = rem % P.C;
    it.b = rem / P.C;
    it.x0 = strip * GB_TW;
    it.y0 = vt * P.TH[it.s];
    it.th = min(P.TH[it.s], P.H - it.y0);
    it.hmax = Q.hmax[it.s];
    it.n_jobs = Q.n_jobs[it.s];
    // image rows the column passes of this tile touch after reflect folding: [lo, lo + nsrc)
    const int a = it.y0 - it.hmax, e = it.y0 + it.th + it.hmax;
    int lo = a < 0 ? 0 : a, hi = e > P.H ? P.H : e;
    if (e > P.H) lo = min(lo, max(0, 2 * P.H - e));
    if (a < 0) hi = max(hi, min(P.H, -a));
    if (-a > P.H || e - P.H > P.H) { lo = 0; hi = P.H; }
    it.lo = lo;
    it.nsrc = hi - lo;                                   // <= nsrc_cap <= 384 (host)
    it.n_rb = (it.nsrc + TC_MROWS - 1) / TC_MROWS;
    it.ksteps = Q.ksteps[it.s];
    it.n_atoms = (it.ksteps + 3) / 4;
    return it;
}

// Tap source of the column sweeps (gabor_dev.cuh: sweep()): the complex taps of one job from the constant table.
struct ConstTaps {
    int w0c;   // complex index of the block-0 window (warp uniform: derived from the work item and the job counter only)
    static constexpr bool DIRECT = true;   // a tap is loaded where it is used: one or two taps are live at a time
    template <int R, bool CT>
    __device__ __forceinline__ void half(int, u64 (&)[CT ? R : 1], float (&)[CT ? 1 : R]) const {}
    template <int R>
    __device__ __forceinline__ u64 tap(int m, int t) const   // tap t (0 .. 2R - 2) of the window of block m
    {
        const float2 v = c_ctaps[w0c - m * R + t];
        u64 w;
        asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(v.x), "f"(v.y));
        return w;
    }
};

// Persistent kernel: one CTA per SM walks the work items i = blockIdx.x, blockIdx.x + gridDim.x, ...; the
// orientation jobs of consecutive items form one stream through the TMA ring and the two TMEM accumulators,
// so the tensor cores already work on the next item while the column warps finish the current one.
template <bool STATS>
__global__ void __launch_bounds__(TC_THREADS, 1)
gabor_tc_kernel(const __grid_constant__ TcParams Q, const __grid_constant__ CUtensorMap map_plane,
                const __grid_constant__ CUtensorMap map_table)
{
    extern __shared__ unsigned char tc_smem_raw[];
    __shared__ __align__(8) unsigned long long s_full[TC_STAGES], s_empty[TC_STAGES], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem;
    const GaborParams &P = Q.g;
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // (warp uniform for the compiler)
    const int n_items = P.first_block[P.S];

    // ---- carve shared memory: [ring of stages, 1024-byte aligned][T][column taps of every job][row table] ----
    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t ring = (raw + 1023u) & ~1023u;
    unsigned char *base = tc_smem_raw + (ring - raw);
    const int n_st = Q.stages;
    float2 *T = reinterpret_cast<float2 *>(base + n_st * TC_STAGE_BYTES);            // [nsrc_cap][GB_TWP]
    float *tap_col = reinterpret_cast<float *>(T + (((size_t)P.nsrc_cap * GB_TWP + 1) & ~(size_t)1));   // 16-byte aligned: 128-bit tap loads
    int *rowtab = reinterpret_cast<int *>(tap_col + (size_t)Q.max_jobs * P.tap_slot);

    TC_TR_DECL;
    // ---- one-time set-up ----
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(smem_u32(&s_full[i]), 1); mbar_init(smem_u32(&s_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_tfull[i]), 1); mbar_init(smem_u32(&s_tempty[i]), TC_COLW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_MMA_WARP) {   // tensor memory: the whole warp allocates, the address lands in shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == TC_TMA_WARP) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const TcItem w = tc_decode(Q, item);
                const int gcol0 = w.x0 - w.hmax + P.P - Q.kshift[w.s];       // first K column in the padded plane (multiple of 8)
                const int prow0 = (w.b * P.C + w.c) * P.H + w.lo;            // first T row in the plane tensor
                for (int ji = 0; ji < w.n_jobs; ++ji)
                    for (int rb = 0; rb < w.n_rb; ++rb)
                        for (int a = 0; a < w.n_atoms; ++a, ++it) {
                            const int st = it % n_st;
                            mbar_wait_parked(smem_u32(&s_empty[st]), ((it / n_st) & 1) ^ 1);
                            const uint32_t fb = smem_u32(&s_full[st]);
                            const uint32_t dst = ring + st * TC_STAGE_BYTES;
                            mbar_expect_tx(fb, TC_STAGE_BYTES);
                            tma_box_2d(dst, &map_plane, gcol0 + TC_KATOM * a, prow0 + TC_MROWS * rb, fb);
#pragma unroll
                            for (int sp = 0; sp < TC_SPLIT; ++sp)
                                tma_box_2d(dst + TC_A_BYTES + sp * TC_B_BYTES, &map_table, TC_KATOM * a,
                                           Q.table_row0[w.s] + (ji * TC_SPLIT + sp) * TC_N, fb);
                        }
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            int it = 0, jg = 0;   // ring step and job counter over the whole item stream
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const TcItem w = tc_decode(Q, item);
                for (int ji = 0; ji < w.n_jobs; ++ji, ++jg) {
                    const int buf = jg & 1;
                    mbar_wait_parked(smem_u32(&s_tempty[buf]), ((jg >> 1) & 1) ^ 1);    // accumulator drained by the column warps
                    tc_fence_after();
                    for (int rb = 0; rb < w.n_rb; ++rb) {
                        const uint32_t d = tmem + (uint32_t)(buf * TC_ACC_COLS + rb * TC_N);
                        for (int a = 0; a < w.n_atoms; ++a, ++it) {
                            const int st = it % n_st;
                            mbar_wait_parked(smem_u32(&s_full[st]), (it / n_st) & 1);
                            tc_fence_after();
                            const uint32_t sa = ring + st * TC_STAGE_BYTES;
                            const uint64_t da = tc_smem_desc(sa);
                            const int nk = min(4, w.ksteps - 4 * a);
                            for (int kk = 0; kk < nk; ++kk)
#pragma unroll
                                for (int sp = 0; sp < TC_SPLIT; ++sp) {
                                    const uint64_t db = tc_smem_desc(sa + TC_A_BYTES + sp * TC_B_BYTES);
                                    // K advance inside the swizzle atom: 16 bf16 = 32 bytes = 2 descriptor units
                                    tc_mma(d, da + 2u * kk, db + 2u * kk, (a | kk | sp) ? 1u : 0u);
                                }
                            tc_commit(smem_u32(&s_empty[st]));                   // stage free once these MMAs have read it
                        }
                    }
                    tc_commit(smem_u32(&s_tfull[buf]));                          // accumulator complete
                }
            }
        }
    } else if (warp < TC_COLW) {
        // =============================== column-pass warps ===============================
        const int D = P.C * P.S * P.O;
        int jg = 0, cur_s = -1, cur_y0 = -1, cur_h = -1;
        TC_TR_ADD(0);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const TcItem w = tc_decode(Q, item);
            const GaborScale &sc = P.scales[w.s];
            float *featb = P.feat + (size_t)w.b * D * P.feat_plane_stride;
            if (w.s != cur_s) {
                // (the trailing barrier of the previous job guarantees nobody still reads the tables)
                // column taps of every job of the scale
                for (int ji = 0; ji < w.n_jobs; ++ji)
                    stage_taps<GB_RC>(tap_col + (size_t)ji * P.tap_slot, P.taps, sc.jobs[ji].col_re, sc.jobs[ji].col_im, sc.jobs[ji].h, TC_COLT);
                cur_s = w.s; cur_h = -1;
            }
            TC_TR_ADD(4);
            for (int ji = 0; ji < w.n_jobs; ++ji, ++jg) {
                const GaborJob job = sc.jobs[ji];
                const int h = Q.job_h[w.s * TC_CTAP_JOBS + ji], buf = jg & 1;
                const float *w_col = stage_taps_window<GB_RC>(tap_col + (size_t)ji * P.tap_slot, job.col_im, h);
                const int nblk_col = (2 * h + GB_RC + GB_RC - 1) / GB_RC;
                if (h != cur_h || w.y0 != cur_y0) {
                    // row table of the job: BYTE offset into T of every input row of the column pass (reflect-folded); entry 0
                    // is row y0 - h, so a block's eight entries are two aligned 128-bit loads (col_pass)
                    const int ne = (w.th + GB_RC - 1) / GB_RC * GB_RC + 2 * h + 2 * GB_RC;
                    for (int e = threadIdx.x; e < ne; e += TC_COLT)
                        rowtab[e] = (reflect_index(w.y0 - h + min(e, w.th + 2 * h - 1), P.H) - w.lo) * (GB_TWP * 8);
                    cur_h = h; cur_y0 = w.y0;
                }
                mbar_wait_parked(smem_u32(&s_tfull[buf]), (jg >> 1) & 1);
                tc_fence_after();
                TC_TR_ADD(1);
                {   // every column warp moves its lane quadrant of the row blocks warp / 4, warp / 4 + TC_XFER_RB, ...
                    const int q = warp & 3;
                    for (int rb = warp >> 2; rb < w.n_rb; rb += TC_XFER_RB) {
                        const int r = rb * TC_MROWS + q * 32 + lane;                 // T row of this thread
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            uint32_t v[32];
                            tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_ACC_COLS + rb * TC_N + half * 32), v);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            if (r < w.nsrc) {
                                float2 *dst = T + (size_t)r * GB_TWP + half * 16;
#pragma unroll
                                for (int j = 0; j < 16; ++j) dst[j] = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&s_tempty[buf]));            // the tensor cores may refill it
                }
                named_bar_sync(1, TC_COLT);                                          // T (and the tables) are in place
                TC_TR_ADD(2);
                const int d0 = (w.c * P.S + w.s) * P.O;
                float *f0 = featb + (size_t)(d0 + job.out0) * P.feat_plane_stride;
                float *f1 = job.out1 >= 0 ? featb + (size_t)(d0 + job.out1) * P.feat_plane_stride : nullptr;
                const int *rt = rowtab;
                const int kind = Q.job_kind[w.s * TC_CTAP_JOBS + ji];
                const bool cx = kind & 1, ct = kind & 2;
                long long *st0 = STATS ? P.stats + ((size_t)w.b * D + d0 + job.out0) * GB_STAT_SLOTS : nullptr;
                long long *st1 = STATS && job.out1 >= 0 ? P.stats + ((size_t)w.b * D + d0 + job.out1) * GB_STAT_SLOTS : nullptr;
                const ConstTaps ctaps{Q.ctap_w0[w.s * TC_CTAP_JOBS + ji]};
                col_pass_dispatch<STATS>(cx, ct, P, T, rt, w_col, nblk_col, w.y0, w.th, w.x0, f0, f1, TC_COLW, st0, st1, h, ctaps);
                TC_TR_ADD(3);
                named_bar_sync(1, TC_COLT);                                          // T may be overwritten
                TC_TR_ADD(5);
            }
        }
#ifdef TC_TRACE
        if (threadIdx.x == 0 && blockIdx.x < TC_TR_CTAS) {
            long long *o = tc_trace_buf[blockIdx.x];
            for (int i = 0; i < 6; ++i) o[i] = tr_acc[i];
            o[6] = clock64() - tr_t0; o[7] = 0;
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
    }
}

typedef CUresult (*tc_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tc_encode_2d(CUtensorMap *map, const void *ptr, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows)
{
    static tc_encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        GCIS_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) return set_error(GCIS_E_CUDA, "gabor: cuTensorMapEncodeTiled not available");
        encode = reinterpret_cast<tc_encode_fn>(fn);
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2u};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(GCIS_E_CUDA, "gabor: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return GCIS_OK;
}

uint16_t bf16_rne(double v)
{
    const float f = (float)v;
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(r >> 16);
}
double bf16_value(uint16_t h)
{
    const uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return (double)f;
}

}  // namespace

struct GaborTcPlan {
    TcParams q;
    std::vector<float2> ctaps;                 // complex column taps as c_ctaps holds them
    unsigned long long ctaps_id = 0;           // hash of ctaps: plans with the same bank share the upload
    size_t smem = 0;
    int kt = 0, table_rows = 0;                // B-operand table: [table_rows][kt] bf16
    __nv_bfloat16 *d_table = nullptr;
    CUtensorMap map_table;
};

size_t gabor_tc_smem_bytes(int nsrc, int hmax, int th_max, int max_jobs, int stages)
{
    const int tap_slot = round_up(2 * (2 * hmax + 1 + 2 * GB_TAP_PAD + 2) + 8, 4);
    const int rowtab = round_up((th_max + GB_RC - 1) / GB_RC * GB_RC + 2 * hmax + 2 * GB_RC, 4);
    return 1024 + (size_t)stages * TC_STAGE_BYTES + sizeof(float2) * (((size_t)nsrc * GB_TWP + 1) & ~(size_t)1) + sizeof(float) * (size_t)tap_slot * max_jobs +
           sizeof(int) * (size_t)rowtab;
}

void gabor_tc_plan_delete(GaborTcPlan *tp)
{
    if (!tp) return;
    cudaFree(tp->d_table);
    delete tp;
}

// Plan for the tensor-core path, or nullptr (with no error set) when the configuration is not covered:
// it needs planes that are exact in bf16 (rgb: the u8 pixel value).
GaborTcPlan *gabor_tc_plan_new(const GaborBankHost &bank, int H, int W, int C, int P, int Wp16, int feature, int colour_space)
{
    if (colour_space != GCIS_COLOUR_RGB) return nullptr;
    const size_t budget = 227 * 1024;
    const int hmax = bank.hmax;
    if (hmax > 96) return nullptr;
    GaborTcPlan *tp = new GaborTcPlan();
    GaborParams &p = tp->q.g;
    memset(&tp->q, 0, sizeof(tp->q));
    p.C = C; p.H = H; p.W = W; p.Wp = Wp16; p.P = P; p.S = bank.S; p.O = bank.O; p.feature = feature;
    p.n_strips = ceil_div(W, GB_TW);
    const int cap = TC_MAX_RB * TC_MROWS;
    int max_jobs = 1;
    for (int s = 0; s < bank.S; ++s) {
        max_jobs = std::max(max_jobs, bank.scales[s].n_jobs);
        tp->q.hmax[s] = bank.scales[s].hmax;
        tp->q.n_jobs[s] = bank.scales[s].n_jobs;
    }
    tp->q.max_jobs = max_jobs;
    // Tile heights for a ring of `stages` stages: the whole image height if T fits beside the ring, else vertical tiles
    // (each pays the 2 hmax halo rows of the row pass again).  Returns the number of tiles of the widest scale, 0 = no fit.
    struct Layout { int nsrc_cap = 0, TH[GB_MAX_SCALES] = {0}, n_vt[GB_MAX_SCALES] = {0}, tiles = 0; };
    auto layout = [&](int stages) {
        Layout L;
        if (H <= cap && gabor_tc_smem_bytes(H, hmax, H, max_jobs, stages) <= budget) {
            L.nsrc_cap = H;
            for (int s = 0; s < bank.S; ++s) { L.TH[s] = H; L.n_vt[s] = 1; }
            L.tiles = 1;
            return L;
        }
        for (int rows = 2 * hmax + GB_RC; rows <= cap && gabor_tc_smem_bytes(rows, hmax, rows, max_jobs, stages) <= budget; rows += GB_RC) L.nsrc_cap = rows;
        if (L.nsrc_cap == 0) return L;
        for (int s = 0; s < bank.S; ++s) {
            const int hs = bank.scales[s].hmax;
            int th = (L.nsrc_cap - 2 * hs) / GB_RC * GB_RC;
            if (th < GB_RC) { L.tiles = 0; return L; }
            if (th > H) th = H;
            L.n_vt[s] = ceil_div(H, th);
            L.TH[s] = round_up(ceil_div(H, L.n_vt[s]), GB_RC);
            if (L.TH[s] > th) L.TH[s] = th;
            L.n_vt[s] = ceil_div(H, L.TH[s]);
            L.tiles = std::max(L.tiles, L.n_vt[s]);
        }
        return L;
    };
    // three stages unless two stages leave room for a taller T that saves vertical tiles (a 481-row portrait image:
    // one tile of 60 row blocks, five per column warp, instead of two tiles with their halo)
    Layout L = layout(TC_STAGES);
    tp->q.stages = TC_STAGES;
    if (L.tiles != 1) {
        const Layout L2 = layout(2);
        if (L2.tiles > 0 && (L.tiles == 0 || L2.tiles < L.tiles)) { L = L2; tp->q.stages = 2; }
    }
    if (L.tiles == 0) { delete tp; return nullptr; }
    const int nsrc_cap = L.nsrc_cap;
    for (int s = 0; s < bank.S; ++s) { p.TH[s] = L.TH[s]; p.n_vt[s] = L.n_vt[s]; }
    int th_max = 0;
    for (int s = 0; s < bank.S; ++s) th_max = std::max(th_max, p.TH[s]);
    p.nsrc_cap = nsrc_cap;
    p.tap_slot = round_up(2 * (2 * hmax + 1 + 2 * GB_TAP_PAD + 2) + 8, 4);
    p.rowtab_cap = round_up((th_max + GB_RC - 1) / GB_RC * GB_RC + 2 * hmax + 2 * GB_RC, 4);
    tp->smem = gabor_tc_smem_bytes(nsrc_cap, hmax, th_max, max_jobs, tp->q.stages);
    std::vector<int> order(bank.S);
    for (int s = 0; s < bank.S; ++s) order[s] = s;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bank.scales[a].hmax > bank.scales[b].hmax; });
    for (int i = 0; i < bank.S; ++i) p.order[i] = order[i];

    // ---- complex column taps for the constant table (ConstTaps); a bank that does not fit runs on the FP32 kernel ----
    {
        int used = 0;
        tp->ctaps.assign(TC_CTAP_CAP, make_float2(0.f, 0.f));
        for (int s = 0; s < bank.S; ++s) {
            const GaborScale &sc = bank.scales[s];
            const int slot = 2 * sc.hmax + 1 + 2 * GB_TAP_PAD + 2 + 4;
            if (sc.n_jobs > TC_CTAP_JOBS || used + sc.n_jobs * slot > TC_CTAP_CAP) { delete tp; return nullptr; }
            tp->q.ctap_base[s] = used; tp->q.ctap_slot[s] = slot;
            for (int ji = 0; ji < sc.n_jobs; ++ji) {
                const GaborJob &job = sc.jobs[ji];
                tp->q.ctap_w0[s * TC_CTAP_JOBS + ji] = -1;
                tp->q.job_h[s * TC_CTAP_JOBS + ji] = job.h;
                tp->q.job_kind[s * TC_CTAP_JOBS + ji] = (job.row_im >= 0 ? 1 : 0) | (job.col_im >= 0 ? 2 : 0);
                if (job.col_im < 0) continue;
                const int ntap = 2 * job.h + 1 + 2 * GB_TAP_PAD;
                float2 *dst = tp->ctaps.data() + used + ji * slot;
                for (int i = 1; i <= ntap; ++i) dst[i] = make_float2(bank.taps[job.col_re + i - 1], bank.taps[job.col_im + i - 1]);
                tp->q.ctap_w0[s * TC_CTAP_JOBS + ji] = used + ji * slot + GB_TAP_PAD + 2 * job.h + 2 - GB_RC;
            }
            used += sc.n_jobs * slot;
        }
        tp->ctaps.resize(used);
        unsigned long long hsh = 1469598103934665603ull;   // FNV-1a over the table
        const unsigned char *bytes = reinterpret_cast<const unsigned char *>(tp->ctaps.data());
        for (size_t i = 0; i < tp->ctaps.size() * sizeof(float2); ++i) hsh = (hsh ^ bytes[i]) * 1099511628211ull;
        tp->ctaps_id = hsh | 1ull;
    }

    // ---- B-operand table: per (scale, job, split term) a [64 n][kt] bf16 matrix, n = 2 c + (re|im) ----
    int kmax = 0, rows = 0;
    for (int s = 0; s < bank.S; ++s) {
        const int hs = bank.scales[s].hmax;
        tp->q.kshift[s] = (P - hs) & 7;
        tp->q.ksteps[s] = ceil_div(GB_TW + 2 * hs + tp->q.kshift[s], 16);
        kmax = std::max(kmax, tp->q.ksteps[s] * 16);
        tp->q.table_row0[s] = rows;
        rows += bank.scales[s].n_jobs * TC_SPLIT * TC_N;
    }
    tp->kt = round_up(kmax, TC_KATOM);
    tp->table_rows = rows;
    std::vector<uint16_t> table((size_t)rows * tp->kt, 0);
    for (int s = 0; s < bank.S; ++s) {
        const GaborScale &sc = bank.scales[s];
        for (int ji = 0; ji < sc.n_jobs; ++ji) {
            const GaborJob &job = sc.jobs[ji];
            for (int n = 0; n < TC_N; ++n) {
                const int cc = n >> 1, part = n & 1;
                const int off = part ? job.row_im : job.row_re;
                if (off < 0) continue;                          // real row filter: the imaginary half stays zero
                for (int k = 0; k < tp->q.ksteps[s] * 16; ++k) {
                    // tap index of plane column x0 - hmax - kshift + k for output column cc
                    const int t = cc + job.h + sc.hmax + tp->q.kshift[s] - k;
                    if (t < 0 || t > 2 * job.h) continue;
                    double v = (double)bank.taps[off + GB_TAP_PAD + t] / 255.0;
                    for (int sp = 0; sp < TC_SPLIT; ++sp) {
                        const uint16_t hb = bf16_rne(v);
                        table[((size_t)tp->q.table_row0[s] + (ji * TC_SPLIT + sp) * TC_N + n) * tp->kt + k] = hb;
                        v -= bf16_value(hb);
                    }
                }
            }
        }
    }
    if (cudaMalloc(reinterpret_cast<void **>(&tp->d_table), table.size() * 2) != cudaSuccess ||
        cudaMemcpy(tp->d_table, table.data(), table.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
        tc_encode_2d(&tp->map_table, tp->d_table, tp->kt, rows, TC_KATOM, TC_N) != GCIS_OK) {
        cudaGetLastError();
        gabor_tc_plan_delete(tp);
        return nullptr;
    }
    return tp;
}

size_t gabor_tc_plan_bytes(const GaborTcPlan *tp) { return tp ? (size_t)tp->table_rows * tp->kt * 2 : 0; }

int colour_planes16_launch(const uint8_t *d_img, void *d_planes16, int B, int H, int W, int P, int Wp16, cudaStream_t st)
{
    const long long total = (long long)B * H * Wp16;
    const int threads = 256;
    const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148 * 16);
    colour_pad16_kernel<<<blocks, threads, 0, st>>>(d_img, static_cast<__nv_bfloat16 *>(d_planes16), B, H, W, P, Wp16);
    GCIS_LAUNCH_CHECK();
    return GCIS_OK;
}

int gabor_tc_launch(GaborTcPlan &tp, const void *d_planes16, float *d_feat, const float *d_taps, const GaborScale *d_scales,
                    int B, int feat_plane_stride, cudaStream_t st, long long *d_stats, float stat_scale)
{
    GaborParams &p = tp.q.g;
    p.stats = d_stats; p.stat_scale = stat_scale;
    p.planes = nullptr; p.feat = d_feat; p.taps = d_taps; p.scales = d_scales; p.B = B;
    p.feat_plane_stride = feat_plane_stride;
    tp.q.plane_rows = B * p.C * p.H;
    int acc = 0;
    for (int i = 0; i < p.S; ++i) {
        p.first_block[i] = acc;
        acc += B * p.C * p.n_strips * p.n_vt[p.order[i]];
    }
    p.first_block[p.S] = acc;
    alignas(64) CUtensorMap map_plane;
    const int rc = tc_encode_2d(&map_plane, d_planes16, (uint64_t)p.Wp, (uint64_t)tp.q.plane_rows, TC_KATOM, TC_MROWS);
    if (rc) return rc;
    static SmemAttrCache attr_cache;
    size_t &attr_smem = attr_cache.cur();
    if (tp.smem > attr_smem) {
        GCIS_CUDA_TRY(cudaFuncSetAttribute(gabor_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp.smem));
        GCIS_CUDA_TRY(cudaFuncSetAttribute(gabor_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp.smem));
        attr_smem = tp.smem;
    }
    // constant table: one bank per device at a time.  A change of bank waits for the kernels that still read the old one;
    // the lock is held until this launch is in its stream, so that no other host thread swaps the table in between.
    static std::mutex mu;
    static unsigned long long owner[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (owner[dev & 63] != tp.ctaps_id) {
        if (owner[dev & 63]) GCIS_CUDA_TRY(cudaDeviceSynchronize());
        owner[dev & 63] = 0;
        GCIS_CUDA_TRY(cudaMemcpyToSymbol(c_ctaps, tp.ctaps.data(), tp.ctaps.size() * sizeof(float2)));
        owner[dev & 63] = tp.ctaps_id;
    }
    static const int n_sm = [] {
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n > 0 ? n : 148;
    }();
    // GCIS_GABOR_SMS=n: run the persistent kernel on n SMs only, leaving the others to the HBM-bound k-means passes of
    // the previous image group that the second lane runs concurrently (plan.cu)
    static const int sm_cap = [] { const char *e = getenv("GCIS_GABOR_SMS"); return e ? atoi(e) : 0; }();
    const int grid = std::min(acc, sm_cap > 0 ? std::min(sm_cap, n_sm) : n_sm);
    if (d_stats) gabor_tc_kernel<true><<<grid, TC_THREADS, tp.smem, st>>>(tp.q, map_plane, tp.map_table);
    else gabor_tc_kernel<false><<<grid, TC_THREADS, tp.smem, st>>>(tp.q, map_plane, tp.map_table);
    GCIS_LAUNCH_CHECK();
    return GCIS_OK;
}

}  // namespace gcis

#ifdef TC_TRACE
extern "C" __attribute__((visibility("default"))) int gcis_tc_trace_read(long long *out, size_t bytes)
{
    return (int)cudaMemcpyFromSymbol(out, gcis::tc_trace_buf, bytes);
}
#endif
