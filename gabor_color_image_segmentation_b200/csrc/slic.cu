// slic.cu — SLIC superpixels, the segmenter the reference actually calls:
//     labels = slic(img, n_segments=300, compactness=10.0)          BSD_metrics/script.py:11,30
// scikit-image is third-party, un-vendored, unpinned (BSD_metrics/README.md:23) and absent from this image, so
// there is nothing to run against: PARITY UNPINNED.  This follows the algorithm scikit-image publishes
// (slic_superpixels.py / _slic.pyx, 0.19 line): float64 CIELAB scaled by 1 / compactness,
// regular_grid seeds with zero colour, max_num_iter rounds of {window-limited assignment with strict '>' in
// centroid order, centroid = mean of its pixels accumulated in raster order}, then the sequential connectivity pass.
//   slic_lab_kernel      u8 RGB -> Lab * ratio (double), sRGB table from the host, IEEE-only cube root
//   slic_assign_kernel   thread = pixel: candidate centroids in index order, windows as scikit-image casts them
//   slic_update_kernel   thread = centroid: raster-order sums over its window (the order the sums are defined in:
//                        fp64 addition is not associative, and the checker must be reproduced bit for bit)
//   host                 _enforce_label_connectivity (breadth-first relabelling in raster order: inherently serial,
//                        ~1 ms per image, as in scikit-image's own Cython)
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace gcis {

namespace {

__host__ __device__ inline double det_cbrt(double t)
{
    // IEEE operations only (no fused multiply-add, no libm): the same bits on the host checker and on the GPU
    if (t <= 0.0) return 0.0;
    int e = 0;
    double m = t;
    while (m >= 1.0) { m *= 0.125; e += 1; }
    while (m < 0.125) { m *= 8.0; e -= 1; }
#ifdef __CUDA_ARCH__
    double y = __dadd_rn(0.4928, __dmul_rn(m, __dsub_rn(0.8203, __dmul_rn(m, 0.3131))));
    for (int i = 0; i < 5; ++i) {
        const double y2 = __dmul_rn(y, y);
        const double num = __dsub_rn(__dmul_rn(y2, y), m);
        const double den = __dmul_rn(3.0, y2);
        y = __dsub_rn(y, __ddiv_rn(num, den));
    }
#else
    double y = 0.4928 + m * (0.8203 - m * 0.3131);
    for (int i = 0; i < 5; ++i) {
        const double y2 = y * y;
        y = y - (y2 * y - m) / (3.0 * y2);
    }
#endif
    while (e > 0) { y *= 2.0; e -= 1; }
    while (e < 0) { y *= 0.5; e += 1; }
    return y;
}

__device__ __forceinline__ double lab_f(double t)
{
    return t > 0.008856 ? det_cbrt(t) : __dadd_rn(__dmul_rn(7.787, t), 16.0 / 116.0);
}

__global__ void slic_lab_kernel(const uint8_t *__restrict__ img, const double *__restrict__ lin, double ratio, int N,
                                double *__restrict__ lab)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const double r = lin[img[3 * p]], g = lin[img[3 * p + 1]], b = lin[img[3 * p + 2]];
    double x = __dadd_rn(__dadd_rn(__dmul_rn(0.412453, r), __dmul_rn(0.357580, g)), __dmul_rn(0.180423, b));
    const double y = __dadd_rn(__dadd_rn(__dmul_rn(0.212671, r), __dmul_rn(0.715160, g)), __dmul_rn(0.072169, b));
    double z = __dadd_rn(__dadd_rn(__dmul_rn(0.019334, r), __dmul_rn(0.119193, g)), __dmul_rn(0.950227, b));
    x = __ddiv_rn(x, 0.95047); z = __ddiv_rn(z, 1.08883);
    const double fx = lab_f(x), fy = lab_f(y), fz = lab_f(z);
    lab[3 * p] = __dmul_rn(__dsub_rn(__dmul_rn(116.0, fy), 16.0), ratio);
    lab[3 * p + 1] = __dmul_rn(__dmul_rn(500.0, __dsub_rn(fx, fy)), ratio);
    lab[3 * p + 2] = __dmul_rn(__dmul_rn(200.0, __dsub_rn(fy, fz)), ratio);
}

// windows of every centroid as scikit-image computes them: [y_min, y_max) x [x_min, x_max), casts truncate
__global__ void slic_window_kernel(const double *__restrict__ segs, int K, int H, int W, double sy, double sx, int4 *__restrict__ win)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const double cy = segs[5 * k], cx = segs[5 * k + 1];
    double t;
    int4 w;
    t = __dsub_rn(cy, __dmul_rn(2.0, sy)); w.x = (int)(long long)(t > 0.0 ? t : 0.0);
    t = __dadd_rn(__dadd_rn(cy, __dmul_rn(2.0, sy)), 1.0); w.y = (int)(long long)(t < (double)H ? t : (double)H);
    t = __dsub_rn(cx, __dmul_rn(2.0, sx)); w.z = (int)(long long)(t > 0.0 ? t : 0.0);
    t = __dadd_rn(__dadd_rn(cx, __dmul_rn(2.0, sx)), 1.0); w.w = (int)(long long)(t < (double)W ? t : (double)W);
    win[k] = w;
}

__global__ void __launch_bounds__(256) slic_assign_kernel(const double *__restrict__ lab, const double *__restrict__ segs,
                                                          const int4 *__restrict__ win, int K, int H, int W, double spatial_weight,
                                                          int32_t *__restrict__ nearest, int *__restrict__ uncovered)
{
    extern __shared__ double s_seg[];          // [K][5] + windows behind it
    int4 *s_win = reinterpret_cast<int4 *>(s_seg + (((size_t)K * 5 + 1) & ~(size_t)1));   // 16-byte aligned
    for (int i = threadIdx.x; i < K * 5; i += blockDim.x) s_seg[i] = segs[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_win[i] = win[i];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    const int y = p / W, x = p - y * W;
    const double l0 = lab[3 * p], l1 = lab[3 * p + 1], l2 = lab[3 * p + 2];
    double best = 1.7976931348623157e308;
    int arg = -1;
    for (int k = 0; k < K; ++k) {
        const int4 w = s_win[k];
        if (y < w.x || y >= w.y || x < w.z || x >= w.w) continue;
        const double *s = s_seg + 5 * k;
        const double ey = __dsub_rn(s[0], (double)y), ex = __dsub_rn(s[1], (double)x);
        double d = __dmul_rn(__dadd_rn(__dmul_rn(ey, ey), __dmul_rn(ex, ex)), spatial_weight);
        const double e0 = __dsub_rn(l0, s[2]), e1 = __dsub_rn(l1, s[3]), e2 = __dsub_rn(l2, s[4]);
        double dc = __dadd_rn(0.0, __dmul_rn(e0, e0));
        dc = __dadd_rn(dc, __dmul_rn(e1, e1));
        dc = __dadd_rn(dc, __dmul_rn(e2, e2));
        d = __dadd_rn(d, dc);
        if (best > d) { best = d; arg = k; }      // strict '>': the lowest centroid index wins ties; NaN never wins
    }
    if (arg >= 0) nearest[p] = arg;
    else atomicAdd(uncovered, 1);                  // keeps its previous label (scikit-image does the same)
}

// thread = centroid: sums in raster order.  Pixels of centroid k lie inside its window unless some pixel was
// covered by no window at all (`uncovered` > 0): then the scan takes the whole image.
__global__ void __launch_bounds__(64) slic_update_kernel(const double *__restrict__ lab, const int32_t *__restrict__ nearest,
                                                         const int4 *__restrict__ win, const int *__restrict__ uncovered, int K,
                                                         int H, int W, double *__restrict__ segs)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    int4 w = win[k];
    if (*uncovered) w = make_int4(0, H, 0, W);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
    long long cnt = 0;
    for (int y = w.x; y < w.y; ++y)
        for (int x = w.z; x < w.w; ++x) {
            const int p = y * W + x;
            if (nearest[p] != k) continue;
            cnt += 1;
            s0 = __dadd_rn(s0, (double)y); s1 = __dadd_rn(s1, (double)x);
            s2 = __dadd_rn(s2, lab[3 * p]); s3 = __dadd_rn(s3, lab[3 * p + 1]); s4 = __dadd_rn(s4, lab[3 * p + 2]);
        }
    const double n = (double)cnt;      // 0 / 0 = NaN: the centroid dies, as in scikit-image
    segs[5 * k] = __ddiv_rn(s0, n); segs[5 * k + 1] = __ddiv_rn(s1, n);
    segs[5 * k + 2] = __ddiv_rn(s2, n); segs[5 * k + 3] = __ddiv_rn(s3, n); segs[5 * k + 4] = __ddiv_rn(s4, n);
}

// skimage.util.regular_grid for a (1, H, W) volume
void slic_grid(int H, int W, int n_points, int &y0, int &sy, int &x0, int &sx)
{
    const double space = (double)H * (double)W;
    if (space <= (double)n_points) { y0 = x0 = 0; sy = sx = 1; return; }
    const double s = std::sqrt(space / (double)n_points);
    const int lo = std::min(H, W), hi = std::max(H, W);
    double s_lo = s, s_hi = s;
    if ((double)lo < s) { s_lo = (double)lo; s_hi = (double)hi / (double)n_points; }
    const double uy = H == W ? s : (H < W ? s_lo : s_hi), ux = H == W ? s : (H < W ? s_hi : s_lo);
    y0 = (int)std::floor(uy / 2.0); x0 = (int)std::floor(ux / 2.0);
    sy = std::max(1, (int)std::nearbyint(uy)); sx = std::max(1, (int)std::nearbyint(ux));
}

// _enforce_label_connectivity_cython (host, serial)
void slic_connectivity(const int32_t *seg, int H, int W, int min_size, int max_size, int start_label, int32_t *out)
{
    static const int ddx[4] = {1, -1, 0, 0}, ddy[4] = {0, 0, 1, -1};
    std::vector<int32_t> cy(std::max(max_size, 1)), cx(std::max(max_size, 1));
    std::fill(out, out + (size_t)H * W, -1);
    int32_t cur = start_label;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (out[(size_t)y * W + x] >= 0) continue;
            int32_t adjacent = 0;
            const int32_t label = seg[(size_t)y * W + x];
            out[(size_t)y * W + x] = cur;
            int size = 1, visited = 0;
            cy[0] = y; cx[0] = x;
            while (visited < size && size < max_size) {
                for (int i = 0; i < 4; ++i) {
                    const int yy = cy[visited] + ddy[i], xx = cx[visited] + ddx[i];
                    if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                    const size_t q = (size_t)yy * W + xx;
                    if (seg[q] == label && out[q] == -1) {
                        out[q] = cur;
                        cy[size] = yy; cx[size] = xx;
                        if (++size >= max_size) break;
                    } else if (out[q] >= 0 && out[q] != cur) {
                        adjacent = out[q];
                    }
                }
                ++visited;
            }
            if (size < min_size) {
                for (int i = 0; i < size; ++i) out[(size_t)cy[i] * W + cx[i]] = adjacent;
            } else {
                ++cur;
            }
        }
}

}  // namespace

}  // namespace gcis

using namespace gcis;

extern "C" {

// labels = slic(img, n_segments, compactness, max_num_iter, enforce_connectivity, start_label) for one H x W x 3
// uint8 image in host memory; h_labels [H][W] int32.  Returns the number of seeds (>= 1) or a negative error.
int32_t gcis_slic_host(const uint8_t *h_img, int32_t H, int32_t W, int32_t n_segments, double compactness, int32_t max_iter,
                       int32_t enforce_connectivity, int32_t start_label, int32_t *h_labels)
{
    if (!h_img || !h_labels || H < 1 || W < 1 || n_segments < 1 || !(compactness > 0) || max_iter < 0)
        return set_error(GCIS_E_INVALID, "slic: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return set_error(GCIS_E_CUDA, "slic: no CUDA device (this library has no CPU path)");
    const int N = H * W;
    int y0, sy, x0, sx;
    slic_grid(H, W, n_segments, y0, sy, x0, sx);
    const int ny = ceil_div(H - y0, sy), nx = ceil_div(W - x0, sx), K = ny * nx;
    if (K < 1) return set_error(GCIS_E_INVALID, "slic: empty seed grid");
    double lin[256];
    for (int i = 0; i < 256; ++i) {
        const double v = i / 255.0;
        lin[i] = v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4);
    }
    std::vector<double> segs((size_t)K * 5, 0.0);
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) { segs[(size_t)(j * nx + i) * 5] = y0 + j * sy; segs[(size_t)(j * nx + i) * 5 + 1] = x0 + i * sx; }
    const size_t smem = sizeof(double) * (((size_t)K * 5 + 1) & ~(size_t)1) + sizeof(int4) * K;
    if (smem > 200 * 1024) return set_error(GCIS_E_INVALID, "slic: %d seeds exceed the shared-memory table", K);
    uint8_t *d_img = nullptr;
    double *d_lin = nullptr, *d_lab = nullptr, *d_segs = nullptr;
    int32_t *d_near = nullptr;
    int4 *d_win = nullptr;
    int *d_unc = nullptr;
    int rc = GCIS_OK;
    auto cleanup = [&]() { cudaFree(d_img); cudaFree(d_lin); cudaFree(d_lab); cudaFree(d_segs); cudaFree(d_near); cudaFree(d_win); cudaFree(d_unc); };
    auto A = [&](void **p, size_t n) { if (!rc && cudaMalloc(p, n) != cudaSuccess) rc = set_error(GCIS_E_NOMEM, "slic: cudaMalloc(%zu) failed", n); };
    A((void **)&d_img, (size_t)N * 3); A((void **)&d_lin, sizeof(lin)); A((void **)&d_lab, sizeof(double) * N * 3);
    A((void **)&d_segs, sizeof(double) * K * 5); A((void **)&d_near, sizeof(int32_t) * N); A((void **)&d_win, sizeof(int4) * K); A((void **)&d_unc, sizeof(int));
    if (rc) { cudaGetLastError(); cleanup(); return rc; }
    cudaStream_t st = nullptr;
    static SmemAttrCache attr;
    if (smem > attr.cur()) {
        if (cudaFuncSetAttribute(slic_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cleanup();
            return set_error(GCIS_E_CUDA, "slic: cudaFuncSetAttribute failed");
        }
        attr.cur() = smem;
    }
    cudaMemcpyAsync(d_img, h_img, (size_t)N * 3, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_lin, lin, sizeof(lin), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_segs, segs.data(), sizeof(double) * K * 5, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(d_near, 0, sizeof(int32_t) * N, st);
    slic_lab_kernel<<<ceil_div(N, 256), 256, 0, st>>>(d_img, d_lin, 1.0 / compactness, N, d_lab);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const int step = std::max(sy, sx);
    const double spatial_weight = 1.0 / ((double)step * (double)step);
    for (int it = 0; it < max_iter; ++it) {
        cudaMemsetAsync(d_unc, 0, sizeof(int), st);
        slic_window_kernel<<<ceil_div(K, 128), 128, 0, st>>>(d_segs, K, H, W, (double)sy, (double)sx, d_win);
        slic_assign_kernel<<<ceil_div(N, 256), 256, smem, st>>>(d_lab, d_segs, d_win, K, H, W, spatial_weight, d_near, d_unc);
        slic_update_kernel<<<ceil_div(K, 64), 64, 0, st>>>(d_lab, d_near, d_win, d_unc, K, H, W, d_segs);
        g_launches.fetch_add(3, std::memory_order_relaxed);
    }
    std::vector<int32_t> near(N);
    if (cudaMemcpyAsync(near.data(), d_near, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        rc = set_error(GCIS_E_CUDA, "slic: %s", cudaGetErrorString(cudaGetLastError()));
    cleanup();
    if (rc) return rc;
    if (enforce_connectivity) {
        const double segment_size = (double)N / (double)K;
        slic_connectivity(near.data(), H, W, (int)(0.5 * segment_size), (int)(3.0 * segment_size), start_label, h_labels);
    } else {
        for (int p = 0; p < N; ++p) h_labels[p] = near[p] + start_label;
    }
    return K;
}

}  // extern "C"
