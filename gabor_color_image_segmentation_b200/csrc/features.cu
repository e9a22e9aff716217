// features.cu — feature assembly after the filter bank: optional Gaussian smoothing of the magnitude planes and
// the per-feature normalisation (z-score) statistics.
//
// Reference: none (segmenter slot, BSD_metrics/script.py:30).  north_star names "per-pixel feature-vector assembly
// (magnitude/energy, optional smoothing and normalisation)"; the spec is DESIGN.md 3.5-3.6:
//   smoothing      plane (c, s, o) is convolved with a Gaussian of sigma = smooth * sigma_s (sigma_s = the Gabor
//                  envelope of scale s), truncated at ceil(3 sigma), taps normalised to sum 1, 'reflect' borders;
//                  separable: row pass then column pass, fp32.
//   normalisation  z_d = (x_d - mean_d) / std_d over the image, with mean and std from EXACT integer moments
//                  S1 = sum r, S2 = sum r^2, r = rint(x 2^16) (order-independent, so bit-identical for any
//                  grid and on the CPU checker); the map z = a x + b is never applied to the feature tensor: the
//                  k-means kernels fold it into their score table (kmeans.cu), so it costs no HBM traffic.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace gcis {

namespace {

constexpr int SM_THREADS = 256;
constexpr int SM_ROWS = 4;        // rows kernel: rows per CTA
constexpr int SM_COLS = 256;      // rows kernel: output columns per CTA (one per thread)
constexpr int SC_ROWS = 64;       // cols kernel: output rows per CTA (8 per thread)
constexpr int SC_COLS = 32;       // cols kernel: columns per CTA (lane = column)

struct SmoothParams {
    const float *in;
    float *out;
    const float *taps;            // per scale: normalised Gaussian taps [2 r + 1] at tap_off[s]
    int tap_off[16], radius[16];
    int B, D, H, W, S, O;
    size_t in_img_stride, out_img_stride;   // floats between images
    int in_plane_stride, out_plane_stride;  // floats between planes
    long long *stats;             // cols kernel only: per-plane moments of what it writes, or null
    float stat_scale;
};

// row pass: out[y][x] = sum_t g[t] in[y][reflect(x + t - r)]
__global__ void __launch_bounds__(SM_THREADS) smooth_rows_kernel(const __grid_constant__ SmoothParams P)
{
    extern __shared__ float sm_s[];
    const int d = blockIdx.y, b = blockIdx.z;
    const int s = (d / P.O) % P.S, r = P.radius[s];
    const int tiles_x = (P.W + SM_COLS - 1) / SM_COLS;
    const int x0 = (blockIdx.x % tiles_x) * SM_COLS, y0 = (blockIdx.x / tiles_x) * SM_ROWS;
    const int span = SM_COLS + 2 * r;
    float *tap = sm_s;                       // [2r + 1]
    float *tile = sm_s + ((2 * r + 1 + 3) & ~3);   // [SM_ROWS][span]
    const float *src = P.in + (size_t)b * P.in_img_stride + (size_t)d * P.in_plane_stride;
    for (int i = threadIdx.x; i <= 2 * r; i += SM_THREADS) tap[i] = P.taps[P.tap_off[s] + i];
    for (int i = threadIdx.x; i < SM_ROWS * span; i += SM_THREADS) {
        const int ry = i / span, cx = i - ry * span;
        const int y = min(y0 + ry, P.H - 1);
        tile[i] = src[(size_t)y * P.W + reflect_index(x0 - r + cx, P.W)];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    float acc[SM_ROWS];
#pragma unroll
    for (int ry = 0; ry < SM_ROWS; ++ry) acc[ry] = 0.f;
    for (int t = 0; t <= 2 * r; ++t) {
        const float g = tap[t];
#pragma unroll
        for (int ry = 0; ry < SM_ROWS; ++ry) acc[ry] = fmaf(g, tile[ry * span + threadIdx.x + t], acc[ry]);
    }
    if (x < P.W) {
        float *dst = P.out + (size_t)b * P.out_img_stride + (size_t)d * P.out_plane_stride;
#pragma unroll
        for (int ry = 0; ry < SM_ROWS; ++ry)
            if (y0 + ry < P.H) dst[(size_t)(y0 + ry) * P.W + x] = acc[ry];
    }
}

// column pass: out[y][x] = sum_t g[t] in[reflect(y + t - r)][x]; optionally accumulates the plane's moments
__global__ void __launch_bounds__(SM_THREADS) smooth_cols_kernel(const __grid_constant__ SmoothParams P)
{
    extern __shared__ float sm_s[];
    const int d = blockIdx.y, b = blockIdx.z;
    const int s = (d / P.O) % P.S, r = P.radius[s];
    const int tiles_x = (P.W + SC_COLS - 1) / SC_COLS;
    const int x0 = (blockIdx.x % tiles_x) * SC_COLS, y0 = (blockIdx.x / tiles_x) * SC_ROWS;
    const int span = SC_ROWS + 2 * r;
    float *tap = sm_s;
    float *tile = sm_s + ((2 * r + 1 + 3) & ~3);   // [span][SC_COLS]
    const float *src = P.in + (size_t)b * P.in_img_stride + (size_t)d * P.in_plane_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i <= 2 * r; i += SM_THREADS) tap[i] = P.taps[P.tap_off[s] + i];
    const int xs = min(x0 + lane, P.W - 1);
    for (int ry = warp; ry < span; ry += SM_THREADS / 32)
        tile[ry * SC_COLS + lane] = src[(size_t)reflect_index(y0 - r + ry, P.H) * P.W + xs];
    __syncthreads();
    constexpr int RPT = SC_ROWS / (SM_THREADS / 32);   // 8 output rows per thread
    float acc[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) acc[i] = 0.f;
    const float *col = tile + (warp * RPT) * SC_COLS + lane;
    for (int t = 0; t <= 2 * r; ++t) {
        const float g = tap[t];
#pragma unroll
        for (int i = 0; i < RPT; ++i) acc[i] = fmaf(g, col[(i + t) * SC_COLS], acc[i]);
    }
    int m1 = 0;
    unsigned long long m2 = 0;
    float *dst = P.out + (size_t)b * P.out_img_stride + (size_t)d * P.out_plane_stride;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int y = y0 + warp * RPT + i;
        if (y < P.H && x0 + lane < P.W) {
            dst[(size_t)y * P.W + x0 + lane] = acc[i];
            stat_add(acc[i], m1, m2);
        }
    }
    if (P.stats) stat_flush(P.stats + ((size_t)b * P.D + d) * GB_STAT_SLOTS, m1, m2, lane);
}

// moments of finished planes (the default normalisation path, and caller-supplied feature tensors): a streaming
// read of the features, CTAs = (plane segment, plane, image); 128-bit loads when the planes are 16-byte aligned
constexpr int FM_SEG = 8192;     // pixels per CTA
__global__ void __launch_bounds__(256) feature_moments_kernel(const float *__restrict__ feat, size_t img_stride, int plane_stride,
                                                              int D, int N, float stat_scale, long long *stats, int vec4)
{
    const int d = blockIdx.y, b = blockIdx.z;
    const float *x = feat + (size_t)b * img_stride + (size_t)d * plane_stride;
    const int p0 = blockIdx.x * FM_SEG, p1 = min(N, p0 + FM_SEG);
    int m1 = 0;                   // <= 32 values per thread: the 32-bit / 64-bit partial sums are safe (common.cuh)
    unsigned long long m2 = 0;
    if (vec4) {
        for (int i = p0 + 4 * threadIdx.x; i < p1; i += 4 * 256) {
            if (i + 4 <= p1) {
                const float4 v = *reinterpret_cast<const float4 *>(x + i);
                stat_add(v.x, m1, m2); stat_add(v.y, m1, m2); stat_add(v.z, m1, m2); stat_add(v.w, m1, m2);
            } else {
                for (int j = i; j < p1; ++j) stat_add(x[j], m1, m2);
            }
        }
    } else {
        for (int i = p0 + threadIdx.x; i < p1; i += 256) stat_add(x[i], m1, m2);
    }
    stat_flush(stats + ((size_t)b * D + d) * GB_STAT_SLOTS, m1, m2, threadIdx.x & 31);
}

// moments -> the affine map of the z-score, z = a x + b (a = b = 0 for a constant plane)
__global__ void feature_affine_kernel(const long long *__restrict__ stats, float *__restrict__ affine, int n, int N, float stat_scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dn = (double)N;
    const long long *m = stats + (size_t)i * GB_STAT_SLOTS;
    const double mean = __ddiv_rn((double)m[0], __dmul_rn(dn, 65536.0));
    const double s2 = __dadd_rn(__dmul_rn((double)m[2], 4294967296.0), (double)m[1]);   // sum r^2, r = rint(x 2^16)
    const double ex2 = __ddiv_rn(s2, __dmul_rn(dn, 4294967296.0));
    double var = __dsub_rn(ex2, __dmul_rn(mean, mean));
    if (var < 0.0) var = 0.0;
    const double sd = __dsqrt_rn(var);
    float a = 0.f, b = 0.f;
    if (sd > 1e-12) {
        a = __double2float_rn(__ddiv_rn(1.0, sd));
        b = __double2float_rn(__ddiv_rn(-mean, sd));
    }
    affine[2 * i] = a;
    affine[2 * i + 1] = b;
}

}  // namespace

// Host side of the smoothing: normalised Gaussian taps per scale.
struct SmoothPlan {
    SmoothParams p;
    float *d_taps = nullptr;
    int rmax = 0;
};

void smooth_plan_delete(SmoothPlan *sp)
{
    if (!sp) return;
    cudaFree(sp->d_taps);
    delete sp;
}

SmoothPlan *smooth_plan_new(const double *sigmas, int S, int O, int D, int H, int W, double factor)
{
    if (!(factor > 0) || S > 16) return nullptr;
    SmoothPlan *sp = new SmoothPlan();
    memset(&sp->p, 0, sizeof(sp->p));
    std::vector<float> taps;
    for (int s = 0; s < S; ++s) {
        const double sig = factor * sigmas[s];
        const int r = std::max(1, (int)std::ceil(3.0 * sig));
        std::vector<double> g(2 * r + 1);
        double sum = 0;
        for (int t = 0; t <= 2 * r; ++t) { g[t] = std::exp(-0.5 * (t - r) * (double)(t - r) / (sig * sig)); sum += g[t]; }
        sp->p.tap_off[s] = (int)taps.size();
        sp->p.radius[s] = r;
        sp->rmax = std::max(sp->rmax, r);
        for (double v : g) taps.push_back((float)(v / sum));
    }
    if (sp->rmax > 512 || cudaMalloc(reinterpret_cast<void **>(&sp->d_taps), taps.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(sp->d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        smooth_plan_delete(sp);
        return nullptr;
    }
    sp->p.taps = sp->d_taps; sp->p.D = D; sp->p.H = H; sp->p.W = W; sp->p.S = S; sp->p.O = O;
    return sp;
}

// feat (in place through tmp): rows feat -> tmp (dense planes), cols tmp -> feat; moments of the result when d_stats != null
int smooth_launch(SmoothPlan &sp, float *d_feat, size_t img_stride, int plane_stride, float *d_tmp, int B, long long *d_stats,
                  float stat_scale, cudaStream_t st)
{
    SmoothParams p = sp.p;
    const int N = p.H * p.W;
    p.B = B; p.stats = nullptr; p.stat_scale = stat_scale;
    p.in = d_feat; p.in_img_stride = img_stride; p.in_plane_stride = plane_stride;
    p.out = d_tmp; p.out_img_stride = (size_t)p.D * N; p.out_plane_stride = N;
    static SmemAttrCache attr_rows, attr_cols;
    const size_t tapf = (2 * sp.rmax + 1 + 3) & ~3;
    const size_t smem_r = sizeof(float) * (tapf + (size_t)SM_ROWS * (SM_COLS + 2 * sp.rmax));
    const size_t smem_c = sizeof(float) * (tapf + (size_t)(SC_ROWS + 2 * sp.rmax) * SC_COLS);
    if (smem_r > attr_rows.cur()) {
        GCIS_CUDA_TRY(cudaFuncSetAttribute(smooth_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
        attr_rows.cur() = smem_r;
    }
    if (smem_c > attr_cols.cur()) {
        GCIS_CUDA_TRY(cudaFuncSetAttribute(smooth_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        attr_cols.cur() = smem_c;
    }
    dim3 gr(ceil_div(p.W, SM_COLS) * ceil_div(p.H, SM_ROWS), p.D, B);
    smooth_rows_kernel<<<gr, SM_THREADS, smem_r, st>>>(p);
    GCIS_LAUNCH_CHECK();
    p.in = d_tmp; p.in_img_stride = (size_t)p.D * N; p.in_plane_stride = N;
    p.out = d_feat; p.out_img_stride = img_stride; p.out_plane_stride = plane_stride;
    p.stats = d_stats;
    dim3 gc(ceil_div(p.W, SC_COLS) * ceil_div(p.H, SC_ROWS), p.D, B);
    smooth_cols_kernel<<<gc, SM_THREADS, smem_c, st>>>(p);
    GCIS_LAUNCH_CHECK();
    return GCIS_OK;
}

int feature_moments_launch(const float *d_feat, size_t img_stride, int plane_stride, int B, int D, int N, float stat_scale,
                           long long *d_stats, cudaStream_t st)
{
    const int vec4 = plane_stride % 4 == 0 && img_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(d_feat) & 15) == 0;
    feature_moments_kernel<<<dim3(ceil_div(N, FM_SEG), D, B), 256, 0, st>>>(d_feat, img_stride, plane_stride, D, N, stat_scale, d_stats, vec4);
    GCIS_LAUNCH_CHECK();
    return GCIS_OK;
}

int feature_affine_launch(const long long *d_stats, float *d_affine, int B, int D, int N, float stat_scale, cudaStream_t st)
{
    const int n = B * D;
    feature_affine_kernel<<<ceil_div(n, 256), 256, 0, st>>>(d_stats, d_affine, n, N, stat_scale);
    GCIS_LAUNCH_CHECK();
    return GCIS_OK;
}

}  // namespace gcis
