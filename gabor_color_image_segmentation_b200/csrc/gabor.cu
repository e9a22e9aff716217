// gabor.cu — colour planes + multi-scale, multi-orientation Gabor bank -> feature planes.
//
// Reference: none.  BSD_metrics/script.py:30 is the segmenter slot; the reference fills it
// with third-party SLIC and contains no filter-bank code, so this stage follows the spec in
// DESIGN.md §3 (scikit-image gabor_kernel definition, sigma_x = sigma_y, reflect borders).
//
// Design.  With sigma_x = sigma_y the complex Gabor kernel is exactly
// rank one: g[y][x] = gy[y] * gx[x] with complex 1-D factors, so one CTA computes a
// 32-column strip of one (image, channel, scale) as a row pass into shared memory followed
// by a column pass out of it, and never writes the intermediate to HBM.  Orientations
// theta and pi - theta share the row pass (conjugate) and the four real column sums.
// Both passes are register-blocked sliding-window FMA loops on the FP32 pipe (the stage is
// FP32-bound, not HBM-bound: 45 MB vs 3.7 GFLOP per 321x481 image).
#include <math.h>

#include <algorithm>

#include "gabor.cuh"

namespace gcis {

// ------------------------------------------------------------------------------------
// Host: bank description
// ------------------------------------------------------------------------------------

double gabor_sigma(double frequency, double bandwidth)
{
    const double b = bandwidth;
    return std::sqrt(std::log(2.0) / 2.0) / M_PI * (std::pow(2.0, b) + 1) / (std::pow(2.0, b) - 1) / frequency;
}

int gabor_half_width(double frequency, double theta, double bandwidth, double n_stds)
{
    const double s = gabor_sigma(frequency, bandwidth);
    const double a = std::fabs(n_stds * s * std::cos(theta)), b = std::fabs(n_stds * s * std::sin(theta));
    return (int)std::ceil(std::max(std::max(a, b), 1.0));
}

int gabor_separable(double frequency, double theta, double bandwidth, double n_stds, std::vector<double> &gx_re,
                    std::vector<double> &gx_im, std::vector<double> &gy_re, std::vector<double> &gy_im)
{
    const double s = gabor_sigma(frequency, bandwidth);
    const int h = gabor_half_width(frequency, theta, bandwidth, n_stds);
    const double ct = std::cos(theta), st = std::sin(theta);
    const double norm = 1.0 / (2.0 * M_PI * s * s);
    gx_re.resize(2 * h + 1); gx_im.resize(2 * h + 1); gy_re.resize(2 * h + 1); gy_im.resize(2 * h + 1);
    for (int t = 0; t <= 2 * h; ++t) {
        const double x = t - h;
        const double env = std::exp(-0.5 * x * x / (s * s));
        const double px = 2.0 * M_PI * frequency * ct * x, py = 2.0 * M_PI * frequency * st * x;
        gx_re[t] = env * std::cos(px);
        gx_im[t] = env * std::sin(px);
        gy_re[t] = env * std::cos(py) * norm;
        gy_im[t] = env * std::sin(py) * norm;
    }
    return h;
}

static int push_taps(std::vector<float> &table, const std::vector<double> &g, bool allow_drop, double ref_max)
{
    double m = 0;
    for (double v : g) m = std::max(m, std::fabs(v));
    if (allow_drop && m <= 1e-12 * ref_max) return -1;
    while (table.size() % 4) table.push_back(0.f);
    const int off = (int)table.size();
    table.insert(table.end(), GB_TAP_PAD, 0.f);
    for (double v : g) table.push_back((float)v);
    table.insert(table.end(), GB_TAP_PAD, 0.f);
    return off;
}

int build_bank(const double *freqs, int S, const double *thetas, int O, double bandwidth, double n_stds,
               GaborBankHost &out)
{
    if (S < 1 || S > GB_MAX_SCALES || O < 1 || O > 2 * GB_MAX_JOBS)
        return set_error(GCIS_E_INVALID, "bank: n_scales=%d n_orient=%d unsupported", S, O);
    out = GaborBankHost();
    out.S = S; out.O = O;
    out.scales.resize(S);
    for (int s = 0; s < S; ++s) {
        if (!(freqs[s] > 0)) return set_error(GCIS_E_INVALID, "bank: frequency[%d] <= 0", s);
        GaborScale &sc = out.scales[s];
        sc.n_jobs = 0; sc.hmax = 0;
        std::vector<int> used(O, 0);
        for (int o = 0; o < O; ++o) {
            if (used[o]) continue;
            used[o] = 1;
            const int h = gabor_half_width(freqs[s], thetas[o], bandwidth, n_stds);
            int partner = -1;
            for (int o2 = o + 1; o2 < O; ++o2)
                if (!used[o2] && std::fabs(thetas[o] + thetas[o2] - M_PI) < 1e-9 &&
                    gabor_half_width(freqs[s], thetas[o2], bandwidth, n_stds) == h) {
                    partner = o2;
                    break;
                }
            if (partner >= 0) used[partner] = 1;
            if (sc.n_jobs >= GB_MAX_JOBS) return set_error(GCIS_E_INVALID, "bank: too many jobs per scale");
            std::vector<double> xr, xi, yr, yi;
            gabor_separable(freqs[s], thetas[o], bandwidth, n_stds, xr, xi, yr, yi);
            double mx = 0, my = 0;
            for (double v : xr) mx = std::max(mx, std::fabs(v));
            for (double v : yr) my = std::max(my, std::fabs(v));
            GaborJob &j = sc.jobs[sc.n_jobs++];
            j.h = h;
            j.row_re = push_taps(out.taps, xr, false, mx);
            j.row_im = push_taps(out.taps, xi, true, mx);
            j.col_re = push_taps(out.taps, yr, false, my);
            j.col_im = push_taps(out.taps, yi, true, my);
            j.out0 = o; j.out1 = partner;
            sc.hmax = std::max(sc.hmax, h);
            const double K = 2 * h + 1;
            const double row = K * (1 + (j.row_im >= 0));
            const double col = K * (1 + (j.row_im >= 0)) * (1 + (j.col_im >= 0));
            out.flops_per_pixel_channel += 2.0 * (row + col);
        }
        out.hmax = std::max(out.hmax, sc.hmax);
    }
    return GCIS_OK;
}

// ------------------------------------------------------------------------------------
// Device: colour planes with horizontal reflect padding
// ------------------------------------------------------------------------------------

namespace {

__device__ __forceinline__ float srgb_to_linear(float v)
{
    return v <= 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);
}
__device__ __forceinline__ float lab_f(float t)
{
    const float d = 6.0f / 29.0f;
    return t > d * d * d ? cbrtf(t) : t / (3.0f * d * d) + 4.0f / 29.0f;
}

// img [B][H][W][3] u8 -> planes [B][3][H][Wp] f32; plane column cp holds image column
// reflect(cp - P): the Gabor row pass then reads its halo without any border logic.
__global__ void colour_pad_kernel(const uint8_t *__restrict__ img, float *__restrict__ planes, int B, int H, int W,
                                  int P, int Wp, int space)
{
    const long long total = (long long)B * H * Wp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cp = (int)(i % Wp);
        const long long br = i / Wp;
        const int r = (int)(br % H);
        const int b = (int)(br / H);
        const int c = reflect_index(cp - P, W);
        const uint8_t *px = img + (((size_t)b * H + r) * W + c) * 3;
        float R = px[0] * (1.0f / 255.0f), G = px[1] * (1.0f / 255.0f), Bl = px[2] * (1.0f / 255.0f);
        float o0, o1, o2;
        if (space == GCIS_COLOUR_RGB) {
            o0 = R; o1 = G; o2 = Bl;
        } else if (space == GCIS_COLOUR_OPPONENT) {
            o0 = (R - G) * 0.70710678118654752f;
            o1 = (R + G - 2.0f * Bl) * 0.40824829046386302f;
            o2 = (R + G + Bl) * 0.57735026918962576f;
        } else {
            const float lr = srgb_to_linear(R), lg = srgb_to_linear(G), lb = srgb_to_linear(Bl);
            const float X = (0.4124564f * lr + 0.3575761f * lg + 0.1804375f * lb) / 0.95047f;
            const float Y = (0.2126729f * lr + 0.7151522f * lg + 0.0721750f * lb);
            const float Z = (0.0193339f * lr + 0.1191920f * lg + 0.9503041f * lb) / 1.08883f;
            const float fx = lab_f(X), fy = lab_f(Y), fz = lab_f(Z);
            o0 = (116.0f * fy - 16.0f) / 100.0f;
            o1 = 500.0f * (fx - fy) / 100.0f;
            o2 = 200.0f * (fy - fz) / 100.0f;
        }
        const size_t plane = (size_t)H * Wp;
        float *dst = planes + (size_t)b * 3 * plane + (size_t)r * Wp + cp;
        dst[0] = o0; dst[plane] = o1; dst[2 * plane] = o2;
    }
}

// ------------------------------------------------------------------------------------
// Device: the bank
// ------------------------------------------------------------------------------------

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

template <int R, bool CT, bool CX, class XLoad>
__device__ __forceinline__ void sweep(XLoad xload, const float *w0, int nblk, u64 (&P)[R], u64 (&Q)[R], float (&S)[R],
                                      const float *xvec = nullptr)
{
    // The window of block m is taps [base - m R, base - m R + 2R): its upper half is the lower half of
    // block m - 1, so each block loads only its R new taps (the tap loads are warp-uniform shared loads
    // and the load pipe, not the FMA pipe, bounds the row pass) and two register halves swap roles.
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
            const float4 *wp = reinterpret_cast<const float4 *>(w0 - (ptrdiff_t)m * R);
#pragma unroll
            for (int q = 0; q < R / 4; ++q) {
                const float4 v = wp[q];
                r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
            }
        }
    };
    auto block = [&](int m, const u64 (&clo)[CT ? R : 1], const u64 (&chi)[CT ? R : 1], const float (&rlo)[CT ? 1 : R],
                     const float (&rhi)[CT ? 1 : R]) {
        float xr[R], xi[R];
        u64 xp[R];
        if (R == 4 && !CX && xvec) {   // row pass: the R inputs of a block are one aligned 128-bit load
            const float4 v = *reinterpret_cast<const float4 *>(xvec + m * 4);
            xr[0] = v.x; xr[1 % R] = v.y; xr[2 % R] = v.z; xr[3 % R] = v.w;
        } else {
#pragma unroll
            for (int uu = 0; uu < R; ++uu) xload(m * R + uu, xr[uu], xi[uu], xp[uu]);
        }
#pragma unroll
        for (int uu = 0; uu < R; ++uu) {
#pragma unroll
            for (int i = 0; i < R; ++i) {
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
    };
    load_half(-1, cb, rb);   // upper half of block 0
#pragma unroll GB_SWEEP_UNR
    for (int m = 0; m < nblk; m += 2) {
        load_half(m, ca, ra);
        block(m, ca, cb, ra, rb);
        if (m + 1 < nblk) {
            load_half(m + 1, cb, rb);
            block(m + 1, cb, ca, rb, ra);
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
template <bool CX, bool CT>
__device__ __forceinline__ void col_pass(const GaborParams &P, const float2 *T, const int *rowtab, const float *w0,
                                         int nblk, int y0, int th, int x0, float *feat0, float *feat1)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nrb = (th + GB_RC - 1) / GB_RC;
    const bool col_ok = x0 + lane < P.W;
    const u64 *Tl = reinterpret_cast<const u64 *>(T) + lane;
    for (int rb = warp; rb < nrb; rb += GB_WARPS) {
        u64 Pv[GB_RC], Qv[GB_RC];
        float Sv[GB_RC];
#pragma unroll
        for (int i = 0; i < GB_RC; ++i) { Pv[i] = 0ull; Qv[i] = 0ull; Sv[i] = 0.f; }
        const int *rt = rowtab + rb * GB_RC;
        sweep<GB_RC, CT, CX>(
            [&](int u, float &xr, float &xi, u64 &xp) {
                xp = Tl[rt[u]];
                unpack2(xp, xr, xi);
            },
            w0, nblk, Pv, Qv, Sv);
#pragma unroll
        for (int i = 0; i < GB_RC; ++i) {
            const int r = rb * GB_RC + i;
            if (r < th && col_ok) {
                float A, Bv = 0.f, Cv = 0.f, Dv = 0.f;
                if constexpr (CT) {
                    unpack2(Pv[i], A, Dv);
                    if constexpr (CX) unpack2(Qv[i], Cv, Bv);
                } else if constexpr (CX) {
                    unpack2(Pv[i], A, Cv);
                } else {
                    A = Sv[i];
                }
                // theta: (A - B) + i(C + D);  pi - theta: (A + B) + i(D - C)
                const float re0 = A - Bv, im0 = Cv + Dv;
                const float e0 = fmaf(re0, re0, im0 * im0);
                const size_t o = (size_t)(y0 + r) * P.W + x0 + lane;
                feat0[o] = P.feature == GCIS_FEATURE_MAGNITUDE ? fast_sqrt(e0) : e0;
                if (feat1) {
                    const float re1 = A + Bv, im1 = Dv - Cv;
                    const float e1 = fmaf(re1, re1, im1 * im1);
                    feat1[o] = P.feature == GCIS_FEATURE_MAGNITUDE ? fast_sqrt(e1) : e1;
                }
            }
        }
    }
}

// Copy one filter's taps from the global table into shared memory in the layout sweep() reads:
// complex -> interleaved (re, im) with tap j at complex index 1 + PAD + j (window starts land on
// 16-byte boundaries); real -> tap j at float index sr + PAD + j.  Returns the block-0 window.
template <int R>
__device__ __forceinline__ const float *stage_taps(float *dst, const float *taps, int off_re, int off_im, int h)
{
    const int ntap = 2 * h + 1 + 2 * GB_TAP_PAD;
    if (off_im >= 0) {
        for (int i = threadIdx.x; i < ntap + 2; i += GB_THREADS) {
            const bool in = i >= 1 && i <= ntap;
            dst[2 * i] = in ? taps[off_re + i - 1] : 0.f;
            dst[2 * i + 1] = in ? taps[off_im + i - 1] : 0.f;
        }
        return dst + 2 * (GB_TAP_PAD + 2 * h + 2 - R);
    }
    const int sr = (4 - ((2 * h + 1) & 3)) & 3;
    for (int i = threadIdx.x; i < ntap + sr + 4; i += GB_THREADS) {
        const int j = i - sr;
        dst[i] = (j >= 0 && j < ntap) ? taps[off_re + j] : 0.f;
    }
    return dst + sr + GB_TAP_PAD + 2 * h - R + 1;
}

#ifdef GB_TRACE   // timing experiment only: per-CTA cycles per phase (thread 0)
constexpr int GB_TR_CTAS = 4096;
__device__ long long gb_trace_buf[GB_TR_CTAS][8];
#define GB_TR_DECL long long tr_t = clock64(), tr_acc[5] = {0, 0, 0, 0, 0}; const long long tr_t0 = tr_t
#define GB_TR_ADD(slot) do { const long long n_ = clock64(); tr_acc[slot] += n_ - tr_t; tr_t = n_; } while (0)
#else
#define GB_TR_DECL do { } while (0)
#define GB_TR_ADD(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(GB_THREADS, 2) gabor_bank_kernel(const __grid_constant__ GaborParams P)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_lo, s_hi;

    // ---- decode the work item: ranges are ordered widest scale first ----
    int range = 0;
    while (range + 1 < P.S && (int)blockIdx.x >= P.first_block[range + 1]) ++range;
    const int s = P.order[range];
    int rem = blockIdx.x - P.first_block[range];
    const int nvt = P.n_vt[s];
    const int vt = rem % nvt; rem /= nvt;
    const int strip = rem % P.n_strips; rem /= P.n_strips;
    const int c = rem % P.C;
    const int b = rem / P.C;
    const int x0 = strip * GB_TW;
    const int y0 = vt * P.TH[s];
    const int th = min(P.TH[s], P.H - y0);

    float *tap_row = smem;                       // first, so the 128-bit tap loads stay 16-byte aligned
    float *tap_col = tap_row + P.tap_slot;
    int *rowtab = reinterpret_cast<int *>(tap_col + P.tap_slot);
    float *chunk = reinterpret_cast<float *>(rowtab + P.rowtab_cap);   // [GB_CHUNK][istr], 16-byte aligned rows
    float2 *T = reinterpret_cast<float2 *>(chunk + gb_chunk_floats(P.istr)); // [nsrc_cap][GB_TWP] complex row-pass output

    const GaborScale &sc = P.scales[s];
    const float *plane = P.planes + ((size_t)b * P.C + c) * P.H * P.Wp;
    const int D = P.C * P.S * P.O;
    float *featb = P.feat + (size_t)b * D * P.feat_plane_stride;
    const int lane = threadIdx.x & 31;

    GB_TR_DECL;
    for (int ji = 0; ji < sc.n_jobs; ++ji) {
        const GaborJob job = sc.jobs[ji];
        const int h = job.h;
        __syncthreads();  // previous job's column pass is done with T, taps and rowtab
        GB_TR_ADD(3);     // (tail of the previous column pass)
        if (threadIdx.x == 0) { s_lo = P.H; s_hi = 0; }
        const float *w_row = stage_taps<GB_RR>(tap_row, P.taps, job.row_re, job.row_im, h);
        const float *w_col = stage_taps<GB_RC>(tap_col, P.taps, job.col_re, job.col_im, h);
        const int nblk_row = (2 * h + GB_RR + GB_RR - 1) / GB_RR, nblk_col = (2 * h + GB_RC + GB_RC - 1) / GB_RC;
        __syncthreads();
        // rows of the image the column pass will touch (reflect-folded), and their span
        const int ne = (th + GB_RC - 1) / GB_RC * GB_RC + 2 * h + 2 * GB_RC;
        {
            int lo = P.H, hi = 0;
            for (int e = threadIdx.x; e < ne; e += GB_THREADS) {
                const int r = reflect_index(y0 - h + min(e, th + 2 * h - 1), P.H);
                rowtab[e] = r;
                lo = min(lo, r); hi = max(hi, r + 1);
            }
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if (lane == 0) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); }
        }
        __syncthreads();
        const int lo = s_lo, hi = s_hi;
        for (int e = threadIdx.x; e < ne; e += GB_THREADS) rowtab[e] = (rowtab[e] - lo) * GB_TWP;

        // ---- row pass: image rows [lo, hi) -> T (complex) in shared memory ----
        const int cw = GB_TW + 2 * h + GB_RR;            // staged columns per row
        const int gcol0 = x0 - h + P.P;                  // first staged column in the padded plane
        const int warp = threadIdx.x >> 5;
        if (cw <= GB_CW2) {
            // narrow filter: 64-row chunks, two rows per lane (warp w stages chunk rows w, w+8, ..., w+56)
            constexpr int SROWS2 = GB_CHUNK2 / GB_WARPS, SCOLS2 = GB_CW2 / 32;
            float stage2[SROWS2][SCOLS2];
            auto fetch2 = [&](int ch0) {
#pragma unroll
                for (int a = 0; a < SROWS2; ++a) {
                    const float *src = plane + (size_t)min(ch0 + warp + a * GB_WARPS, P.H - 1) * P.Wp + gcol0 + lane;
#pragma unroll
                    for (int j = 0; j < SCOLS2; ++j) stage2[a][j] = __ldg(src + 32 * j);
                }
            };
            fetch2(lo);
            GB_TR_ADD(0);
            for (int ch0 = lo; ch0 < hi; ch0 += GB_CHUNK2) {
                __syncthreads();                         // chunk buffer free (and rowtab complete on 1st pass)
                GB_TR_ADD(2);
#pragma unroll
                for (int a = 0; a < SROWS2; ++a)
#pragma unroll
                    for (int j = 0; j < SCOLS2; ++j) chunk[(warp + a * GB_WARPS) * GB_ISTR2 + lane + 32 * j] = stage2[a][j];
                __syncthreads();
                GB_TR_ADD(1);
                if (ch0 + GB_CHUNK2 < hi) fetch2(ch0 + GB_CHUNK2);
                if (job.row_im >= 0) row_pass_chunk2<true>(chunk, w_row, nblk_row, T, ch0 - lo, hi - ch0);
                else row_pass_chunk2<false>(chunk, w_row, nblk_row, T, ch0 - lo, hi - ch0);
                GB_TR_ADD(2);
            }
        } else {
            // Staging: warp w owns chunk rows w, w+8, w+16, w+24; a lane covers columns lane + 32 j.
            // The next chunk is fetched into registers while the current one is convolved, so the
            // global-load latency hides behind the row pass.
            constexpr int SROWS = GB_CHUNK / GB_WARPS;                       // 4 rows per warp
            constexpr int SCOLS = (GB_TW + 2 * 96 + GB_RR + 31) / 32;        // lane columns for the widest supported row
            const int ncol = (cw + 31) / 32;                                 // <= SCOLS (h <= 96 checked on the host)
            float stage[SROWS][SCOLS];
            // The planes are padded so that every staged address is in bounds: no per-lane guards.
            auto fetch = [&](int ch0) {
    #pragma unroll
                for (int a = 0; a < SROWS; ++a) {
                    const float *src = plane + (size_t)min(ch0 + warp + a * GB_WARPS, P.H - 1) * P.Wp + gcol0 + lane;
    #pragma unroll
                    for (int j = 0; j < SCOLS; ++j)
                        if (j < ncol) stage[a][j] = __ldg(src + 32 * j);
                }
            };
            fetch(lo);
            GB_TR_ADD(0);
            for (int ch0 = lo; ch0 < hi; ch0 += GB_CHUNK) {
                __syncthreads();                             // chunk buffer free (and rowtab complete on 1st pass)
                GB_TR_ADD(2);                                // (waiting for the slowest warp of the row pass)
    #pragma unroll
                for (int a = 0; a < SROWS; ++a)
    #pragma unroll
                    for (int j = 0; j < SCOLS; ++j)
                        if (j < ncol) chunk[(warp + a * GB_WARPS) * P.istr + lane + 32 * j] = stage[a][j];
                __syncthreads();
                GB_TR_ADD(1);
                if (ch0 + GB_CHUNK < hi) fetch(ch0 + GB_CHUNK);
                const int trow = ch0 - lo + lane;
                const bool active = ch0 + lane < hi;
                if (job.row_im >= 0) row_pass_chunk<true>(chunk, P.istr, w_row, nblk_row, T, trow, active);
                else row_pass_chunk<false>(chunk, P.istr, w_row, nblk_row, T, trow, active);
                GB_TR_ADD(2);
            }
        }
        __syncthreads();
        GB_TR_ADD(2);

        // ---- column pass: T -> |response| for theta (and pi - theta) ----
        const int d0 = (c * P.S + s) * P.O;
        float *f0 = featb + (size_t)(d0 + job.out0) * P.feat_plane_stride;
        float *f1 = job.out1 >= 0 ? featb + (size_t)(d0 + job.out1) * P.feat_plane_stride : nullptr;
        const bool cx = job.row_im >= 0, ct = job.col_im >= 0;
        if (cx && ct) col_pass<true, true>(P, T, rowtab, w_col, nblk_col, y0, th, x0, f0, f1);
        else if (cx) col_pass<true, false>(P, T, rowtab, w_col, nblk_col, y0, th, x0, f0, f1);
        else if (ct) col_pass<false, true>(P, T, rowtab, w_col, nblk_col, y0, th, x0, f0, f1);
        else col_pass<false, false>(P, T, rowtab, w_col, nblk_col, y0, th, x0, f0, f1);
        GB_TR_ADD(3);
    }
#ifdef GB_TRACE
    if (threadIdx.x == 0 && blockIdx.x < GB_TR_CTAS) {
        long long *o = gb_trace_buf[blockIdx.x];
        o[0] = tr_acc[0]; o[1] = tr_acc[1]; o[2] = tr_acc[2]; o[3] = tr_acc[3]; o[4] = clock64() - tr_t0; o[5] = s;
    }
#endif
}

}  // namespace

int colour_planes_launch(const uint8_t *d_img, float *d_planes, int B, int H, int W, int P, int Wp, int space,
                         cudaStream_t st)
{
    const long long total = (long long)B * H * Wp;
    const int threads = 256;
    const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148 * 16);
    colour_pad_kernel<<<blocks, threads, 0, st>>>(d_img, d_planes, B, H, W, P, Wp, space);
    GCIS_LAUNCH_CHECK();
    return GCIS_OK;
}

// Shared-memory plan for the bank kernel on an H x W image.
struct GaborLaunchPlan {
    GaborParams p;
    size_t smem = 0;
    int blocks = 0;
};

static size_t gabor_smem_bytes(int nsrc, int hmax, int th_max)
{
    const int istr = (GB_TW + 2 * hmax + GB_RR + 31) / 32 * 32 + 4;
    const int tap_slot = round_up(2 * (2 * hmax + 1 + 2 * GB_TAP_PAD + 2) + 8, 4);
    const int rowtab = round_up((th_max + GB_RC - 1) / GB_RC * GB_RC + 2 * hmax + 2 * GB_RC, 4);
    return sizeof(float) * ((size_t)2 * nsrc * GB_TWP + (size_t)gb_chunk_floats(istr) + 2 * (size_t)tap_slot) +
           sizeof(int) * (size_t)rowtab;
}

int gabor_plan(const GaborBankHost &bank, int H, int W, int C, int P, int Wp, int feature, GaborLaunchPlan &lp)
{
    GaborParams &p = lp.p;
    memset(&p, 0, sizeof(p));
    p.C = C; p.H = H; p.W = W; p.Wp = Wp; p.P = P; p.S = bank.S; p.O = bank.O; p.feature = feature;
    p.n_strips = ceil_div(W, GB_TW);
    const int hmax = bank.hmax;
    if (hmax > 96) return set_error(GCIS_E_INVALID, "gabor: kernel half-width %d > 96 is not supported", hmax);
    const size_t two_per_sm = 112 * 1024, one_per_sm = 226 * 1024;
    size_t budget;
    int nsrc_cap;
    if (gabor_smem_bytes(H, hmax, H) <= one_per_sm) {
        // whole image height per CTA: the row pass is never recomputed for a vertical halo
        nsrc_cap = H;
        for (int s = 0; s < bank.S; ++s) { p.TH[s] = H; p.n_vt[s] = 1; }
        budget = gabor_smem_bytes(H, hmax, H);
    } else {
        budget = two_per_sm;
        nsrc_cap = 0;
        for (int rows = 2 * hmax + GB_RC; gabor_smem_bytes(rows, hmax, rows) <= budget; rows += GB_RC) nsrc_cap = rows;
        if (nsrc_cap == 0) {
            budget = one_per_sm;
            for (int rows = 2 * hmax + GB_RC; gabor_smem_bytes(rows, hmax, rows) <= budget; rows += GB_RC) nsrc_cap = rows;
        }
        if (nsrc_cap == 0) return set_error(GCIS_E_INVALID, "gabor: kernel half-width %d too large for shared memory", hmax);
        for (int s = 0; s < bank.S; ++s) {
            const int hs = bank.scales[s].hmax;
            int th = (nsrc_cap - 2 * hs) / GB_RC * GB_RC;
            if (th < GB_RC) th = GB_RC;
            if (th > H) th = H;
            p.n_vt[s] = ceil_div(H, th);
            p.TH[s] = round_up(ceil_div(H, p.n_vt[s]), GB_RC);  // balance the tiles
            if (p.TH[s] > th) p.TH[s] = th;
            p.n_vt[s] = ceil_div(H, p.TH[s]);
        }
    }
    int th_max = 0;
    for (int s = 0; s < bank.S; ++s) th_max = std::max(th_max, p.TH[s]);
    p.nsrc_cap = nsrc_cap;
    p.istr = (GB_TW + 2 * hmax + GB_RR + 31) / 32 * 32 + 4;
    p.tap_slot = round_up(2 * (2 * hmax + 1 + 2 * GB_TAP_PAD + 2) + 8, 4);
    p.rowtab_cap = round_up((th_max + GB_RC - 1) / GB_RC * GB_RC + 2 * hmax + 2 * GB_RC, 4);
    lp.smem = gabor_smem_bytes(nsrc_cap, hmax, th_max);
    // widest scale first so the long CTAs are not left for the tail
    std::vector<int> order(bank.S);
    for (int s = 0; s < bank.S; ++s) order[s] = s;
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return bank.scales[a].hmax > bank.scales[b].hmax; });
    for (int i = 0; i < bank.S; ++i) p.order[i] = order[i];
    return GCIS_OK;
}

GaborLaunchPlan *gabor_plan_new(const GaborBankHost &bank, int H, int W, int C, int P, int Wp, int feature, size_t *smem)
{
    GaborLaunchPlan *lp = new GaborLaunchPlan();
    if (gabor_plan(bank, H, W, C, P, Wp, feature, *lp)) {
        delete lp;
        return nullptr;
    }
    if (smem) *smem = lp->smem;
    return lp;
}

void gabor_plan_delete(GaborLaunchPlan *lp) { delete lp; }

int gabor_launch(GaborLaunchPlan &lp, const float *d_planes, float *d_feat, const float *d_taps,
                 const GaborScale *d_scales, int B, int feat_plane_stride, cudaStream_t st)
{
    GaborParams &p = lp.p;
    p.planes = d_planes; p.feat = d_feat; p.taps = d_taps; p.scales = d_scales; p.B = B;
    p.feat_plane_stride = feat_plane_stride;
    int acc = 0;
    for (int i = 0; i < p.S; ++i) {
        p.first_block[i] = acc;
        acc += B * p.C * p.n_strips * p.n_vt[p.order[i]];
    }
    p.first_block[p.S] = acc;
    lp.blocks = acc;
    static SmemAttrCache attr_cache;
    size_t &attr_smem = attr_cache.cur();
    if (lp.smem > attr_smem) {
        GCIS_CUDA_TRY(cudaFuncSetAttribute(gabor_bank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.smem));
        attr_smem = lp.smem;
    }
    gabor_bank_kernel<<<acc, GB_THREADS, lp.smem, st>>>(p);
    GCIS_LAUNCH_CHECK();
    return GCIS_OK;
}

}  // namespace gcis

#ifdef GB_TRACE
extern "C" __attribute__((visibility("default"))) int gcis_gb_trace_read(long long *out, size_t bytes)
{
    return (int)cudaMemcpyFromSymbol(out, gcis::gb_trace_buf, bytes);
}
#endif
