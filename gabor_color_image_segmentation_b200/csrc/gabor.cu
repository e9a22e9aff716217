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

#include "gabor_dev.cuh"

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

using namespace gbdev;

#ifdef GB_TRACE   // timing experiment only: per-CTA cycles per phase (thread 0)
constexpr int GB_TR_CTAS = 4096;
__device__ long long gb_trace_buf[GB_TR_CTAS][8];
#define GB_TR_DECL long long tr_t = clock64(), tr_acc[5] = {0, 0, 0, 0, 0}; const long long tr_t0 = tr_t
#define GB_TR_ADD(slot) do { const long long n_ = clock64(); tr_acc[slot] += n_ - tr_t; tr_t = n_; } while (0)
#else
#define GB_TR_DECL do { } while (0)
#define GB_TR_ADD(slot) do { } while (0)
#endif

template <bool STATS>
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
        for (int e = threadIdx.x; e < ne; e += GB_THREADS) rowtab[e] = (rowtab[e] - lo) * (GB_TWP * 8);   // byte offsets into T

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
        long long *st0 = STATS ? P.stats + ((size_t)b * D + d0 + job.out0) * GB_STAT_SLOTS : nullptr;
        long long *st1 = STATS && job.out1 >= 0 ? P.stats + ((size_t)b * D + d0 + job.out1) * GB_STAT_SLOTS : nullptr;
        col_pass_dispatch<STATS>(cx, ct, P, T, rowtab, w_col, nblk_col, y0, th, x0, f0, f1, GB_WARPS, st0, st1, h, SmemTaps{w_col});
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
                 const GaborScale *d_scales, int B, int feat_plane_stride, cudaStream_t st, long long *d_stats, float stat_scale)
{
    GaborParams &p = lp.p;
    p.stats = d_stats; p.stat_scale = stat_scale;
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
        GCIS_CUDA_TRY(cudaFuncSetAttribute(gabor_bank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.smem));
        GCIS_CUDA_TRY(cudaFuncSetAttribute(gabor_bank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.smem));
        attr_smem = lp.smem;
    }
    if (d_stats) gabor_bank_kernel<true><<<acc, GB_THREADS, lp.smem, st>>>(p);
    else gabor_bank_kernel<false><<<acc, GB_THREADS, lp.smem, st>>>(p);
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
