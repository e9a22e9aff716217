// gabor_dev.cuh — device code shared by the two filter-bank kernels (gabor.cu: both passes on the FP32 pipe;
// gabor_tc.cu: row pass on the tcgen05 tensor cores, column pass on the FP32 pipe).
// Reference: none (segmenter slot, BSD_metrics/script.py:30; spec in DESIGN.md section 3).
#pragma once
#include "async.cuh"
#include "gabor.cuh"

namespace gcis {
namespace gbdev {

constexpr int GB_TW = 32;        // strip width = one warp of columns
constexpr int GB_TWP = 33;       // odd stride: row-pass writes (lane = row) and column-pass reads conflict-free
constexpr int GB_THREADS = 256;
constexpr int GB_WARPS = GB_THREADS / 32;
constexpr int GB_RC = 8;         // column pass: output rows per thread
constexpr int GB_RR = 4;         // row pass: output columns per thread (32 / 4 = 8 column blocks = 8 warps)
constexpr int GB_CHUNK = 32;     // input rows staged per row-pass step (lane = row)
// Narrow filters (staged row <= 64 columns, i.e. half-width <= 14): 64-row chunks, two rows per lane.  The tap
// loads are shared by the two rows and the per-chunk staging / barrier cost is paid half as often.
constexpr int GB_CHUNK2 = 64;
constexpr int GB_CW2 = 64;       // staged columns per row on this path
constexpr int GB_ISTR2 = 68;     // chunk row stride: 32 n + 4 floats
__host__ __device__ constexpr int gb_chunk_floats(int istr)
{
    return GB_CHUNK * istr > GB_CHUNK2 * GB_ISTR2 ? GB_CHUNK * istr : GB_CHUNK2 * GB_ISTR2;
}

struct GaborParams {
    const float *planes;   // [B][C][H][Wp]
    float *feat;           // [B][C*S*O][H][W]
    const float *taps;
    const GaborScale *scales;
    int B, C, H, W, Wp, P, S, O, feature;
    int feat_plane_stride;       // floats between feature planes (>= H*W)
    int n_strips;
    int TH[GB_MAX_SCALES];       // output rows per CTA at scale s
    int n_vt[GB_MAX_SCALES];     // vertical tiles at scale s
    int first_block[GB_MAX_SCALES + 1];  // block ranges ordered from the widest scale to the narrowest
    int order[GB_MAX_SCALES];    // scale handled by range i
    int nsrc_cap;                // rows of T the shared buffer holds
    int istr;                    // chunk row stride: 32 n + 4 floats (aligned, conflict-free 128-bit row loads)
    int tap_slot;                // floats reserved per staged filter (complex, interleaved) in shared memory
    int rowtab_cap;
    // per-plane moments for the optional normalisation, accumulated in the epilogue:
    // [B][D][GB_STAT_SLOTS] (common.cuh)
    long long *stats;            // null = not requested
    float stat_scale;            // 2^fix_shift
};

typedef unsigned long long u64;

// acc.{lo,hi} += a.{lo,hi} * s: one packed FP32 FMA (fma.rn.f32x2, sm_100+), half the issue slots
__device__ __forceinline__ void fma2_vs(u64 &acc, u64 a, float s)
{
    u64 b;
    asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(s));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// 64-bit shared load at a 32-bit shared-window address (one IADD per address instead of generic pointer arithmetic)
__device__ __forceinline__ u64 lds64(uint32_t addr)
{
    u64 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}

// Register-blocked sliding window shared by both passes:
//     out[i] += sum_u g[i + 2h - u] * x[u],   i < R, u < nblk*R
// The 2R-1 taps a block of R inputs needs are re-read each block with aligned 128-bit shared
// loads (broadcast, one wavefront each).  Complex taps are stored interleaved (re, im), so one
// packed FMA updates the (re*x, im*x) pair of an output; with real taps and complex inputs the
// pair is (w*xr, w*xi).
//   CT: complex taps     CX: complex input
//   CT &&  CX: P[i] = (A, D) += W*xr,  Q[i] = (C, B) += W*xi
//   CT && !CX: P[i] = (A, D) += W*x
//  !CT &&  CX: P[i] = (A, C) += w*X
//  !CT && !CX: S[i] = A      += w*x
// w0 points at the window of block 0; the window moves down by R taps per block.
#ifndef GB_SWEEP_UNROLL
#define GB_SWEEP_UNROLL 1
#endif
constexpr int GB_SWEEP_UNR = GB_SWEEP_UNROLL;   // pairs of blocks per unrolled step of the sweeps

//
// The first and the last block are triangular: block 0 only reaches taps <= 2h for outputs i <= u, the last block only
// taps >= 0 for outputs i >= u + d with d = (nblk - 1) R - 2h; the other products multiply the zero padding around the
// taps.  With `d` >= 0 (caller guarantees 2h >= R, so no block is cut on both sides) those products are not issued:
// 16 % fewer FMAs for the 15/29/55/109-tap bank, same sums (x * 0 added to an accumulator leaves its value).
template <int N> struct IntC { static constexpr int value = N; };

// Tap source of sweep(): half(m, c, r) delivers the R new taps of block m (complex pairs in c, real taps in r).
// SmemTaps reads the window stage_taps() laid out in shared memory (warp-uniform 128-bit loads into ordinary registers);
// a kernel may pass a source of its own (gabor_tc.cu: a constant-memory table read tap by tap (DIRECT), so that the taps
// live in UNIFORM registers and the packed FMA reads only its accumulator pair and one scalar from the register file).
struct SmemTaps {
    static constexpr bool DIRECT = false;   // the taps are held in registers half a window at a time (half())
    const float *w0;   // window of block 0; the window moves down by R taps per block
    template <int R> __device__ __forceinline__ u64 tap(int, int) const { return 0ull; }
    template <int R, bool CT>
    __device__ __forceinline__ void half(int m, u64 (&c)[CT ? R : 1], float (&r)[CT ? 1 : R]) const
    {
        if constexpr (CT) {
            const ulonglong2 *wp = reinterpret_cast<const ulonglong2 *>(w0 - (ptrdiff_t)m * 2 * R);
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
                const ulonglong2 v = wp[q];
                c[2 * q] = v.x; c[2 * q + 1] = v.y;
            }
        } else {
            const float4 *wp = reinterpret_cast<const float4 *>(w0 - (ptrdiff_t)m * R);
#pragma unroll
            for (int q = 0; q < R / 4; ++q) {
                const float4 v = wp[q];
                r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
            }
        }
    }
};

template <int R, bool CT, bool CX, class XLoad, class Taps>
__device__ __forceinline__ void sweep(XLoad xload, const Taps &taps, int nblk, u64 (&P)[R], u64 (&Q)[R], float (&S)[R],
                                      const float *xvec = nullptr, int d = -1)
{
    // The window of block m is taps [base - m R, base - m R + 2R): its upper half is the lower half of
    // block m - 1, so each block loads only its R new taps (the tap loads are warp-uniform shared loads
    // and the load pipe, not the FMA pipe, bounds the row pass) and two register halves swap roles.
    u64 ca[CT ? R : 1], cb[CT ? R : 1];
    float ra[CT ? 1 : R], rb[CT ? 1 : R];
    auto load_half = [&](int m, u64 (&c)[CT ? R : 1], float (&r)[CT ? 1 : R]) {
        if constexpr (!Taps::DIRECT) taps.template half<R, CT>(m, c, r);
    };
    // MODE 0: all R x R products; 1: first block (outputs i <= u); 2: last block (outputs i >= u + DD)
    auto block = [&](auto mode_c, auto dd_c, int m, const u64 (&clo)[CT ? R : 1], const u64 (&chi)[CT ? R : 1],
                     const float (&rlo)[CT ? 1 : R], const float (&rhi)[CT ? 1 : R]) {
        constexpr int MODE = decltype(mode_c)::value, DD = decltype(dd_c)::value;
        constexpr int NIN = MODE == 2 ? R - DD : R;   // inputs that reach at least one output
        float xr[R], xi[R];
        u64 xp[R];
        if (R == 4 && !CX && xvec) {   // row pass: the R inputs of a block are one aligned 128-bit load
            const float4 v = *reinterpret_cast<const float4 *>(xvec + m * 4);
            xr[0] = v.x; xr[1 % R] = v.y; xr[2 % R] = v.z; xr[3 % R] = v.w;
        } else {
            xload(m, IntC<NIN>{}, xp);   // the block's first NIN inputs
#pragma unroll
            for (int uu = 0; uu < NIN; ++uu) unpack2(xp[uu], xr[uu], xi[uu]);
        }
        if constexpr (Taps::DIRECT && CT) {
            // A DIRECT source hands out tap t of the block's window where it is used (tap(m, t)); the products are issued
            // tap by tap, t descending: one or two taps are live at a time (they fit the uniform registers), and every
            // output still adds its products in the order uu = 0, 1, ... of the loop below: the same bits.
#pragma unroll
            for (int t = 2 * R - 2; t >= 0; --t) {
                bool any = false;
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const int uu = i - t + R - 1;
                    if (uu < 0 || uu >= NIN || (MODE == 1 && i > uu) || (MODE == 2 && i < uu + DD)) continue;
                    any = true;
                }
                if (!any) continue;
                const u64 w = taps.template tap<R>(m, t);
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const int uu = i - t + R - 1;
                    if (uu < 0 || uu >= NIN || (MODE == 1 && i > uu) || (MODE == 2 && i < uu + DD)) continue;
                    fma2_vs(P[i], w, xr[uu]);
                    if constexpr (CX) fma2_vs(Q[i], w, xi[uu]);
                }
            }
        } else {
#pragma unroll
            for (int uu = 0; uu < NIN; ++uu) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    if (MODE == 1 && i > uu) continue;
                    if (MODE == 2 && i < uu + DD) continue;
                    const int t = i - uu + R - 1;
                    if constexpr (CT) {
                        const u64 w = t < R ? clo[t % R] : chi[t % R];
                        fma2_vs(P[i], w, xr[uu]);
                        if constexpr (CX) fma2_vs(Q[i], w, xi[uu]);
                    } else {
                        const float w = t < R ? rlo[t % R] : rhi[t % R];
                        if constexpr (CX) fma2_vs(P[i], xp[uu], w);
                        else S[i] = fmaf(w, xr[uu], S[i]);
                    }
                }
            }
        }
    };
    if (d >= 0) {
        const int last = nblk - 1;
        load_half(0, ca, ra);
        block(IntC<1>{}, IntC<0>{}, 0, ca, ca, ra, ra);            // the upper half (padding) is never read
        int m = 1;
#pragma unroll 1
        while (m < last) {
            load_half(m, cb, rb);
            block(IntC<0>{}, IntC<0>{}, m, cb, ca, rb, ra);
            ++m;
            if (m < last) {
                load_half(m, ca, ra);
                block(IntC<0>{}, IntC<0>{}, m, ca, cb, ra, rb);
                ++m;
            }
        }
        if (last & 1) {   // the half shared with block last - 1 sits in ca: the last block reads it from cb
#pragma unroll
            for (int q = 0; q < (CT ? R : 1); ++q) cb[q] = ca[q];
#pragma unroll
            for (int q = 0; q < (CT ? 1 : R); ++q) rb[q] = ra[q];
        }
        switch (d) {
        case 2: block(IntC<2>{}, IntC<2>{}, last, ca, cb, ra, rb); break;   // d >= 1: only the shared half is reached
        case 4: block(IntC<2>{}, IntC<4>{}, last, ca, cb, ra, rb); break;
        case 6: block(IntC<2>{}, IntC<6 < R ? 6 : 0>{}, last, ca, cb, ra, rb); break;
        default:                                                            // d = 0 (tap 0 of the new half) or any other
            load_half(last, ca, ra);
            if (d == 0) block(IntC<2>{}, IntC<0>{}, last, ca, cb, ra, rb);
            else block(IntC<0>{}, IntC<0>{}, last, ca, cb, ra, rb);
        }
        return;
    }
    load_half(-1, cb, rb);   // upper half of block 0
#pragma unroll GB_SWEEP_UNR
    for (int m = 0; m < nblk; m += 2) {
        load_half(m, ca, ra);
        block(IntC<0>{}, IntC<0>{}, m, ca, cb, ra, rb);
        if (m + 1 < nblk) {
            load_half(m + 1, cb, rb);
            block(IntC<0>{}, IntC<0>{}, m + 1, cb, ca, rb, ra);
        }
    }
}

// Row pass of one staged chunk: lane = image row, warp = block of GB_RR output columns.
// One tap half-window (R taps) of the row pass: R complex pairs or R real taps.
template <bool CT>
struct RowTaps {
    u64 c[CT ? 4 : 1];
    float r[CT ? 1 : 4];
    __device__ __forceinline__ void load(const float *w0, int m)
    {
        if constexpr (CT) {
            const ulonglong2 *wp = reinterpret_cast<const ulonglong2 *>(w0 - (ptrdiff_t)m * 8);
            const ulonglong2 v0 = wp[0], v1 = wp[1];
            c[0] = v0.x; c[1] = v0.y; c[2] = v1.x; c[3] = v1.y;
        } else {
            const float4 v = *reinterpret_cast<const float4 *>(w0 - (ptrdiff_t)m * 4);
            r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
        }
    }
};

template <bool CT>
__device__ __forceinline__ void row_pass_chunk(const float *chunk, int istr, const float *w0, int nblk, float2 *T,
                                               int trow, bool active)
{
    // Software-pipelined form of sweep<4, CT, false>: the inputs and the new tap half of block m + 1 are
    // loaded before the 16 FMAs of block m (three tap buffers rotate, two input buffers alternate), so the
    // shared-memory latency of a block hides behind the previous block's math.  Per output the taps are
    // applied in the same order as in sweep().
    constexpr int R = 4;
    static_assert(GB_RR == R, "row pass is written for 4 output columns per thread");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xb = warp * R;
    const float *src = chunk + lane * istr + xb;
    u64 Pv[R];
    float Sv[R];
#pragma unroll
    for (int i = 0; i < R; ++i) { Pv[i] = 0ull; Sv[i] = 0.f; }
    auto block = [&](const float4 &x4, const RowTaps<CT> &lo, const RowTaps<CT> &hi) {
        const float x[R] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int uu = 0; uu < R; ++uu)
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int t = i - uu + R - 1;
                if constexpr (CT) fma2_vs(Pv[i], t < R ? lo.c[t % R] : hi.c[t % R], x[uu]);
                else Sv[i] = fmaf(t < R ? lo.r[t % R] : hi.r[t % R], x[uu], Sv[i]);
            }
    };
    auto xload = [&](int m) { return *reinterpret_cast<const float4 *>(src + 4 * min(m, nblk - 1)); };
    RowTaps<CT> A, B, C;
    A.load(w0, -1);
    B.load(w0, 0);
    float4 x0 = xload(0), x1;
    int m = 0;
    // block m uses lo = taps(m), hi = taps(m - 1); the loads for block m + 1 are issued first
#define GB_ROW_STEP(XC, XN, LO, HI, NX)            \
    NX.load(w0, min(m + 1, nblk - 1));             \
    XN = xload(m + 1);                             \
    block(XC, LO, HI);                             \
    if (++m >= nblk) break;
#pragma unroll 1
    for (;;) {
        GB_ROW_STEP(x0, x1, B, A, C)
        GB_ROW_STEP(x1, x0, C, B, A)
        GB_ROW_STEP(x0, x1, A, C, B)
        GB_ROW_STEP(x1, x0, B, A, C)
        GB_ROW_STEP(x0, x1, C, B, A)
        GB_ROW_STEP(x1, x0, A, C, B)
    }
#undef GB_ROW_STEP
    if (active) {
        u64 *dst = reinterpret_cast<u64 *>(T + (size_t)trow * GB_TWP + xb);
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if constexpr (CT) dst[i] = Pv[i];                      // (Tr, Ti)
            else T[(size_t)trow * GB_TWP + xb + i] = make_float2(Sv[i], 0.f);
        }
    }
}

// Row pass of a 64-row chunk: lane = image rows `lane` and `lane + 32` of the chunk, warp = block of 4 output
// columns.  Same arithmetic per output as row_pass_chunk (one FMA per tap, taps in the same order); real taps
// update the two rows with one packed FMA.
template <bool CT>
__device__ __forceinline__ void row_pass_chunk2(const float *chunk, const float *w0, int nblk, float2 *T, int trow0,
                                                int n_rows)
{
    constexpr int R = 4;
    static_assert(GB_RR == R, "two-row path is written for 4 output columns per thread");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xb = warp * R;
    const float *s0 = chunk + lane * GB_ISTR2 + xb, *s1 = s0 + 32 * GB_ISTR2;
    u64 P0[R], P1[R];   // CT: (Tr, Ti) of row 0 / row 1;  !CT: P0[i] = (T of row 0, T of row 1)
#pragma unroll
    for (int i = 0; i < R; ++i) { P0[i] = 0ull; P1[i] = 0ull; }
    u64 ca[CT ? R : 1], cb[CT ? R : 1];
    float ra[CT ? 1 : R], rb[CT ? 1 : R];
    auto load_half = [&](int m, u64 (&c)[CT ? R : 1], float (&r)[CT ? 1 : R]) {
        if constexpr (CT) {
            const ulonglong2 *wp = reinterpret_cast<const ulonglong2 *>(w0 - (ptrdiff_t)m * 2 * R);
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
                const ulonglong2 v = wp[q];
                c[2 * q] = v.x; c[2 * q + 1] = v.y;
            }
        } else {
            const float4 v = *reinterpret_cast<const float4 *>(w0 - (ptrdiff_t)m * R);
            r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
        }
    };
    auto block = [&](int m, const u64 (&clo)[CT ? R : 1], const u64 (&chi)[CT ? R : 1], const float (&rlo)[CT ? 1 : R],
                     const float (&rhi)[CT ? 1 : R]) {
        const float4 v0 = *reinterpret_cast<const float4 *>(s0 + m * 4), v1 = *reinterpret_cast<const float4 *>(s1 + m * 4);
        const float x0[R] = {v0.x, v0.y, v0.z, v0.w}, x1[R] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int uu = 0; uu < R; ++uu) {
            u64 xp = 0ull;
            if constexpr (!CT) asm("mov.b64 %0, {%1, %2};" : "=l"(xp) : "f"(x0[uu]), "f"(x1[uu]));
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int t = i - uu + R - 1;
                if constexpr (CT) {
                    const u64 w = t < R ? clo[t % R] : chi[t % R];
                    fma2_vs(P0[i], w, x0[uu]);
                    fma2_vs(P1[i], w, x1[uu]);
                } else {
                    fma2_vs(P0[i], xp, t < R ? rlo[t % R] : rhi[t % R]);
                }
            }
        }
    };
    load_half(-1, cb, rb);
#pragma unroll GB_SWEEP_UNR
    for (int m = 0; m < nblk; m += 2) {
        load_half(m, ca, ra);
        block(m, ca, cb, ra, rb);
        if (m + 1 < nblk) {
            load_half(m + 1, cb, rb);
            block(m + 1, cb, ca, rb, ra);
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (lane + 32 * h < n_rows) {
            float2 *dst = T + (size_t)(trow0 + lane + 32 * h) * GB_TWP + xb;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                if constexpr (CT) {
                    reinterpret_cast<u64 *>(dst)[i] = h ? P1[i] : P0[i];
                } else {
                    float a, b;
                    unpack2(P0[i], a, b);
                    dst[i] = make_float2(h ? b : a, 0.f);
                }
            }
        }
    }
}

// sqrt.approx (MUFU): relative error <= 2^-22, far inside the feature tolerance; the IEEE
// sqrtf costs a Newton fix-up and a slow-path branch per output, which was 13 % of the kernel's stall samples.
__device__ __forceinline__ float fast_sqrt(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Column pass: lane = column of the strip, warp = blocks of GB_RC output rows.
constexpr int GB_TAIL_MAX = 2;   // leftover rows (th mod GB_RC) up to this many get the thin single-row path
#ifndef GB_TRI
#define GB_TRI 1                 // triangular first/last sweep blocks (needs the registers of a <= 16-warp CTA)
#endif
constexpr int GB_THIN_COLS = 4;  // strips with at most this many image columns get the lanes-on-rows path

template <bool CX, bool CT, bool STATS, class Taps>
__device__ __forceinline__ void col_pass(const GaborParams &P, const float2 *T, const int *rowtab, const float *w0,
                                         int nblk, int y0, int th, int x0, float *feat0, float *feat1, int nwarps,
                                         long long *st0, long long *st1, int h, const Taps &taps)
{
    // (the shuffle tells the compiler that the warp index is warp uniform: loop counters and addresses derived from it may
    // then live in uniform registers)
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    // BSDS images are 321 or 481 rows = whole blocks of GB_RC rows + ONE row: as a masked block that row costs a full
    // block and unbalances the warps (41 blocks over 16 warps: one SM sub-partition gets 11, the others 10).  Up to
    // GB_TAIL_MAX leftover rows are instead computed one at a time (same taps in the same order: same bits) by the
    // warps that own the fewest blocks.  (Dealing the rows in units of 4 so that every warp gets the same number, with
    // one 4-row block per warp, was measured slower: 36.7 vs 34.3 us per image at K = 109.)
    const int nfull = th / GB_RC, tail = th - nfull * GB_RC;
    const bool thin_tail = h >= 0 && tail > 0 && tail <= GB_TAIL_MAX && nfull >= nwarps && P.W - x0 > GB_THIN_COLS;
    const int nrb = thin_tail ? nfull : (th + GB_RC - 1) / GB_RC;
    const uint32_t tl = smem_u32(T) + 8u * lane;   // this lane's column of T in the shared window; the row table holds BYTE offsets
    // triangular first/last block of the sweep when the taps span at least one block (always for the BSDS bank)
    const int tri_d = (GB_TRI && h >= 0 && 2 * h >= GB_RC && nblk == (2 * h + 2 * GB_RC - 1) / GB_RC) ? (nblk - 1) * GB_RC - 2 * h : -1;
    int m1[2] = {0, 0};                      // exact integer moments of what this thread writes (normalisation)
    unsigned long long m2[2] = {0, 0};
    const bool mag = P.feature == GCIS_FEATURE_MAGNITUDE, pair = feat1 != nullptr;
    // accumulators of one output -> the feature value(s): theta: (A - B) + i(C + D);  pi - theta: (A + B) + i(D - C)
    auto values = [&](u64 Pa, u64 Qa, float Sa, float &v0, float &v1) {
        float A, Bv = 0.f, Cv = 0.f, Dv = 0.f;
        if constexpr (CT) {
            unpack2(Pa, A, Dv);
            if constexpr (CX) unpack2(Qa, Cv, Bv);
        } else if constexpr (CX) {
            unpack2(Pa, A, Cv);
        } else {
            A = Sa;
        }
        const float re0 = A - Bv, im0 = Cv + Dv;
        const float e0 = fmaf(re0, re0, im0 * im0);
        v0 = mag ? fast_sqrt(e0) : e0;
        v1 = 0.f;
        if (pair) {
            const float re1 = A + Bv, im1 = Dv - Cv;
            const float e1 = fmaf(re1, re1, im1 * im1);
            v1 = mag ? fast_sqrt(e1) : e1;
        }
    };
    auto store = [&](float *o0, float *o1, float v0, float v1) {
        *o0 = v0;
        if constexpr (STATS) stat_add(v0, m1[0], m2[0]);
        if (pair) {
            *o1 = v1;
            if constexpr (STATS) if (st1) stat_add(v1, m1[1], m2[1]);
        }
    };
    auto emit = [&](int r, int col, u64 Pa, u64 Qa, float Sa) {   // one output of the thin paths
        if (r < th && x0 + col < P.W) {
            float v0, v1;
            values(Pa, Qa, Sa, v0, v1);
            const size_t o = (size_t)(y0 + r) * P.W + x0 + col;
            store(feat0 + o, feat1 + o, v0, v1);
        }
    };
    // the GB_RC rows of a block: one pointer per plane stepping by the image width, no per-row index arithmetic; whole
    // blocks (all but a masked last one) without per-row tests (the column pass is issue bound: see profiles/)
    auto emit_block = [&](int r0, const u64 (&Pv)[GB_RC], const u64 (&Qv)[GB_RC], const float (&Sv)[GB_RC]) {
        if (x0 + lane >= P.W) return;
        const size_t o = (size_t)(y0 + r0) * P.W + x0 + lane;
        float *o0 = feat0 + o, *o1 = feat1 + o;
        const int nrow = th - r0;
        if (nrow >= GB_RC) {
#pragma unroll
            for (int i = 0; i < GB_RC; ++i, o0 += P.W, o1 += P.W) {
                float v0, v1;
                values(Pv[i], Qv[i], Sv[i], v0, v1);
                store(o0, o1, v0, v1);
            }
        } else {
#pragma unroll
            for (int i = 0; i < GB_RC; ++i, o0 += P.W, o1 += P.W)
                if (i < nrow) {
                    float v0, v1;
                    values(Pv[i], Qv[i], Sv[i], v0, v1);
                    store(o0, o1, v0, v1);
                }
        }
    };
    // One output (row r, strip column col) as a plain tap loop: the taps in the order the sweep applies them, so the
    // same bits.  r and col may differ from lane to lane.
    auto one_output = [&](int r, int col) {
        const int *rt = rowtab + r;
        const uint32_t tc = smem_u32(T) + 8u * col;
        u64 Pa = 0ull, Qa = 0ull;
        float Sa = 0.f;
        if constexpr (CT) {
            const u64 *tp = reinterpret_cast<const u64 *>(w0 - 2 * (2 * h + 1 - GB_RC));   // complex tap 0
#pragma unroll 4
            for (int u = 0; u <= 2 * h; ++u) {
                float xr, xi;
                unpack2(lds64(tc + (uint32_t)rt[u]), xr, xi);
                const u64 w = tp[2 * h - u];
                fma2_vs(Pa, w, xr);
                if constexpr (CX) fma2_vs(Qa, w, xi);
            }
        } else {
            const float *tp = w0 - (2 * h - GB_RC + 1);                                      // real tap 0
#pragma unroll 4
            for (int u = 0; u <= 2 * h; ++u) {
                const u64 xp = lds64(tc + (uint32_t)rt[u]);
                const float w = tp[2 * h - u];
                if constexpr (CX) {
                    fma2_vs(Pa, xp, w);
                } else {
                    float xr, xi;
                    unpack2(xp, xr, xi);
                    Sa = fmaf(w, xr, Sa);
                }
            }
        }
        emit(r, col, Pa, Qa, Sa);
    };
    // 481 and 321 columns are whole 32-column strips + ONE column: the last strip would run the full register-blocked
    // sweep for 1 useful lane in 32 (6 % / 9 % of the whole column pass).  A strip of at most GB_THIN_COLS columns
    // instead puts the lanes on ROWS, one output per lane and pass.
    const int ncols = min(GB_TW, P.W - x0);
    const bool thin_strip = h >= 0 && ncols <= GB_THIN_COLS;
    if (thin_strip) {
        for (int c = 0; c < ncols; ++c)
            for (int r = warp * 32 + lane; r < th; r += nwarps * 32) one_output(r, c);
    }
    for (int rb = warp; rb < (thin_strip ? 0 : nrb); rb += nwarps) {
        u64 Pv[GB_RC], Qv[GB_RC];
        float Sv[GB_RC];
#pragma unroll
        for (int i = 0; i < GB_RC; ++i) { Pv[i] = 0ull; Qv[i] = 0ull; Sv[i] = 0.f; }
        const int *rt = rowtab + rb * GB_RC;
        sweep<GB_RC, CT, CX>(
            [&](int m, auto nin_c, u64 (&xp)[GB_RC]) {
                // the block's eight row-table entries (byte offsets into T) as two 128-bit loads, then one add per row
                constexpr int NIN = decltype(nin_c)::value;
                const int4 *rp = reinterpret_cast<const int4 *>(rt + m * GB_RC);
                const int4 a = rp[0];
                int4 b = a;
                if constexpr (NIN > 4) b = rp[1];
                const int off[GB_RC] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int uu = 0; uu < NIN; ++uu) xp[uu] = lds64(tl + (uint32_t)off[uu]);
            },
            taps, nblk, Pv, Qv, Sv, nullptr, tri_d);
        emit_block(rb * GB_RC, Pv, Qv, Sv);
    }
    if (thin_tail)
        for (int tr = 0; tr < tail; ++tr)
            if (warp == nwarps - 1 - tr) one_output(nfull * GB_RC + tr, lane);
    if constexpr (STATS) {   // warp-shuffle reduction, then one atomic per warp, plane and moment (integers: order-independent)
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            long long *dst = pl ? st1 : st0;
            if (!dst) continue;
            stat_flush(dst, m1[pl], m2[pl], lane);
        }
    }
}

// run-time (CX, CT) -> compile-time instantiation; STATS (the moments of the normalisation) is a property of the whole
// kernel, so that the kernel without it carries none of its registers
// `ctaps`: the tap source of the sweeps with COMPLEX column taps (real taps always come from the staged window w0)
template <bool STATS, class CTaps>
__device__ __forceinline__ void col_pass_dispatch(bool cx, bool ct, const GaborParams &P, const float2 *T, const int *rowtab,
                                                  const float *w0, int nblk, int y0, int th, int x0, float *f0, float *f1,
                                                  int nwarps, long long *st0, long long *st1, int h, const CTaps &ctaps)
{
#define GB_COL(CXV, CTV, TAPS) col_pass<CXV, CTV, STATS>(P, T, rowtab, w0, nblk, y0, th, x0, f0, f1, nwarps, st0, st1, h, TAPS)
    if (cx && ct) GB_COL(true, true, ctaps);
    else if (cx) GB_COL(true, false, SmemTaps{w0});
    else if (ct) GB_COL(false, true, ctaps);
    else GB_COL(false, false, SmemTaps{w0});
#undef GB_COL
}

// Block-0 window of a filter that stage_taps() already copied to `dst` (same arithmetic as its return value).
template <int R>
__device__ __forceinline__ const float *stage_taps_window(const float *dst, int off_im, int h)
{
    if (off_im >= 0) return dst + 2 * (GB_TAP_PAD + 2 * h + 2 - R);
    const int sr = (4 - ((2 * h + 1) & 3)) & 3;
    return dst + sr + GB_TAP_PAD + 2 * h - R + 1;
}

// Copy one filter's taps from the global table into shared memory in the layout sweep() reads:
// complex -> interleaved (re, im) with tap j at complex index 1 + PAD + j (window starts land on
// 16-byte boundaries); real -> tap j at float index sr + PAD + j.  Returns the block-0 window.
template <int R>
__device__ __forceinline__ const float *stage_taps(float *dst, const float *taps, int off_re, int off_im, int h, int nthr = GB_THREADS)
{
    const int ntap = 2 * h + 1 + 2 * GB_TAP_PAD;
    if (off_im >= 0) {
        for (int i = threadIdx.x; i < ntap + 2; i += nthr) {
            const bool in = i >= 1 && i <= ntap;
            dst[2 * i] = in ? taps[off_re + i - 1] : 0.f;
            dst[2 * i + 1] = in ? taps[off_im + i - 1] : 0.f;
        }
        return dst + 2 * (GB_TAP_PAD + 2 * h + 2 - R);
    }
    const int sr = (4 - ((2 * h + 1) & 3)) & 3;
    for (int i = threadIdx.x; i < ntap + sr + 4; i += nthr) {
        const int j = i - sr;
        dst[i] = (j >= 0 && j < ntap) ? taps[off_re + j] : 0.f;
    }
    return dst + sr + GB_TAP_PAD + 2 * h - R + 1;
}


}  // namespace gbdev
}  // namespace gcis
