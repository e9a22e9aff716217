/*
 * gcis_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the segmentation + BSD-evaluation hot path, used only
 * as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs. Nothing under gabor_color_image_segmentation_b200/
 * may import, link or execute this file.
 *
 * Parity status
 *   - Metrics stage (orc_label_metrics & helpers): PINNED. It follows the
 *     reference BSD_metrics/metrics.py line by line (citations on each function)
 *     and is checked against golden vectors produced by running the reference's
 *     own metrics.py in the build container (oracle/make_golden.py ->
 *     tests/golden/metrics_*.npz).
 *   - Gabor bank / feature assembly / k-means (orc_conv*, orc_kmeans): PARITY
 *     UNPINNED. The reference snapshot holds no code for these stages (the
 *     segmenter slot at BSD_metrics/script.py:30 calls third-party SLIC); they
 *     follow this repo's own spec in DESIGN.md §3.  The
 *     convolution primitive is additionally cross-checked against
 *     scipy.ndimage.convolve(mode='reflect') in tests/test_oracle.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 * -ffp-contract=off matters: the k-means oracle restates an fp32 FMA chain
 * bit-for-bit and must not let the compiler fuse or unfuse anything.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#if defined(__x86_64__)
#define ORC_CLONES __attribute__((target_clones("arch=x86-64-v3", "default")))
#else
#define ORC_CLONES
#endif

/* ------------------------------------------------------------------------- */
/* Metrics stage                                                              */
/* ------------------------------------------------------------------------- */

/* skimage.segmentation.find_boundaries(x) with its defaults (connectivity=1,
 * mode='thick'), as called at metrics.py:49,69,88,157.  A pixel is a boundary
 * iff some in-bounds 4-neighbour carries a different label. */
ORC_API void orc_find_boundaries(const int32_t *x, int H, int W, uint8_t *bd)
{
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            int32_t v = x[(size_t)r * W + c];
            int b = 0;
            if (r > 0 && x[(size_t)(r - 1) * W + c] != v) b = 1;
            if (r + 1 < H && x[(size_t)(r + 1) * W + c] != v) b = 1;
            if (c > 0 && x[(size_t)r * W + c - 1] != v) b = 1;
            if (c + 1 < W && x[(size_t)r * W + c + 1] != v) b = 1;
            bd[(size_t)r * W + c] = (uint8_t)b;
        }
}

/* Window offsets of dilation(b, rectangle(size,size)) as used at
 * metrics.py:69,93.  Odd size s: [-(s-1)/2, +(s-1)/2].  Even size (never used
 * by the reference's defaults; follows the scipy grey_dilation convention of
 * the SURVEY Appendix-C stand-in): [-(s/2-1), +s/2]. */
ORC_API void orc_dilate_window(int size, int *lo, int *hi)
{
    if (size & 1) { *lo = -(size - 1) / 2; *hi = (size - 1) / 2; }
    else { *lo = -(size / 2 - 1); *hi = size / 2; }
}

/* dilation(b, rectangle(size,size)): OR over the window, out-of-bounds
 * neighbours ignored. */
ORC_API void orc_dilate_square(const uint8_t *b, int H, int W, int size, uint8_t *out)
{
    int lo, hi;
    orc_dilate_window(size, &lo, &hi);
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            int v = 0;
            for (int dr = lo; dr <= hi && !v; ++dr) {
                int rr = r + dr;
                if (rr < 0 || rr >= H) continue;
                for (int dc = lo; dc <= hi; ++dc) {
                    int cc = c + dc;
                    if (cc < 0 || cc >= W) continue;
                    if (b[(size_t)rr * W + cc]) { v = 1; break; }
                }
            }
            out[(size_t)r * W + c] = (uint8_t)v;
        }
}

/*
 * All integer counts behind metrics.set_metrics() (metrics.py:208-217) for one
 * image.  Outputs (SURVEY.md A.3–A.7):
 *   bd_count            |bd(lb)|                                 metrics.py:88-90,157
 *   den_r[g], tp_r[g]   |bd(gt_g)|, |dil_size(bd(lb)) & bd(gt_g)| metrics.py:69-72
 *   tp_p[g]             |bd(lb) & dil_5(bd(gt_g))| (5 hard-coded) metrics.py:91-94
 *   hist[g][i][j]       contingency table, row stride n_lab_cap   metrics.py:115-126
 *   U[g], V[g]          Van den Bergh / Neubert-Protzel numerators metrics.py:129-142
 *   area[i], perim[i]   metrics.py:166-180,197
 * n_seg = max(lb)+1 (metrics.py:51); n_lab[g] = max(gt_g)+1 (metrics.py:115).
 * Returns 0, or -1 if a label is negative or exceeds the caps.
 */
ORC_API int orc_label_metrics(const int32_t *lb, const int32_t *gt, int H, int W, int G,
                              int n_seg_cap, int n_lab_cap, int size_recall,
                              int64_t *bd_count, int64_t *den_r, int64_t *tp_r,
                              int64_t *tp_p, int64_t *U, int64_t *V,
                              int64_t *area, int64_t *perim, int64_t *hist,
                              int32_t *n_seg_out, int32_t *n_lab_out)
{
    size_t N = (size_t)H * W;
    uint8_t *bd = malloc(N), *dil = malloc(N), *tbd = malloc(N), *tdil = malloc(N);
    int rc = 0;
    int32_t mx = -1;
    for (size_t p = 0; p < N; ++p) {
        if (lb[p] < 0) rc = -1;
        if (lb[p] > mx) mx = lb[p];
    }
    int n_seg = mx + 1;
    *n_seg_out = n_seg;
    if (n_seg > n_seg_cap) rc = -1;
    if (rc) goto done;

    orc_find_boundaries(lb, H, W, bd);
    orc_dilate_square(bd, H, W, size_recall, dil);
    int64_t nb = 0;
    for (size_t p = 0; p < N; ++p) nb += bd[p];
    *bd_count = nb;

    memset(area, 0, sizeof(int64_t) * n_seg_cap);
    memset(perim, 0, sizeof(int64_t) * n_seg_cap);
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            size_t p = (size_t)r * W + c;
            area[lb[p]] += 1;
            /* metrics.py:172-180: image-border pixel, else 4-neighbour test */
            if (r == 0 || r == H - 1 || c == 0 || c == W - 1 || bd[p]) perim[lb[p]] += 1;
        }

    for (int g = 0; g < G; ++g) {
        const int32_t *t = gt + (size_t)g * N;
        int64_t *h = hist + (size_t)g * n_seg_cap * n_lab_cap;
        int32_t tm = -1;
        for (size_t p = 0; p < N; ++p) {
            if (t[p] < 0) rc = -1;
            if (t[p] > tm) tm = t[p];
        }
        int n_lab = tm + 1;
        n_lab_out[g] = n_lab;
        if (n_lab > n_lab_cap) rc = -1;
        if (rc) goto done;

        orc_find_boundaries(t, H, W, tbd);
        orc_dilate_square(tbd, H, W, 5, tdil); /* metrics.py:93 ignores `size` */
        int64_t d = 0, a = 0, b = 0;
        for (size_t p = 0; p < N; ++p) {
            d += tbd[p];
            a += dil[p] & tbd[p];
            b += bd[p] & tdil[p];
        }
        den_r[g] = d; tp_r[g] = a; tp_p[g] = b;

        memset(h, 0, sizeof(int64_t) * (size_t)n_seg_cap * n_lab_cap);
        for (size_t p = 0; p < N; ++p) h[(size_t)lb[p] * n_lab_cap + t[p]] += 1;

        int64_t u = 0, v = 0;
        for (int i = 0; i < n_seg; ++i) {
            int64_t rowsum = 0, rowmax = 0;
            for (int j = 0; j < n_lab; ++j) {
                int64_t x = h[(size_t)i * n_lab_cap + j];
                rowsum += x;
                if (x > rowmax) rowmax = x;
            }
            u += rowsum - rowmax;                         /* metrics.py:130-131 */
            for (int j = 0; j < n_lab; ++j) {             /* metrics.py:138-140 */
                int64_t x = h[(size_t)i * n_lab_cap + j];
                int64_t o = rowsum - x;
                v += x < o ? x : o;
            }
        }
        U[g] = u; V[g] = v;
    }
done:
    free(bd); free(dil); free(tbd); free(tdil);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* Gabor stage (builder-defined spec; parity unpinned upstream)  */
/* ------------------------------------------------------------------------- */

/* scipy.ndimage 'reflect' index folding: (d c b a | a b c d | d c b a). */
static inline int reflect_idx(int i, int n)
{
    if (n == 1) return 0;
    int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

/* Definition: out = scipy.ndimage.convolve(img, ker, mode='reflect') for an
 * odd-sized kernel: out[y,x] = sum_{dy,dx} ker[cy+dy, cx+dx] * img[y-dy, x-dx]. */
ORC_CLONES
ORC_API void orc_conv2d_reflect(const double *img, int H, int W, const double *ker,
                                int kh, int kw, double *out)
{
    int cy = kh / 2, cx = kw / 2;
    int *ry = malloc(sizeof(int) * (size_t)(H + 2 * cy)), *rx = malloc(sizeof(int) * (size_t)(W + 2 * cx));
    for (int i = 0; i < H + 2 * cy; ++i) ry[i] = reflect_idx(i - cy, H);
    for (int i = 0; i < W + 2 * cx; ++i) rx[i] = reflect_idx(i - cx, W);
    /* padded copy so the inner loop is a plain dot product */
    int PW = W + 2 * cx, PH = H + 2 * cy;
    double *pad = malloc(sizeof(double) * (size_t)PW * PH);
    for (int y = 0; y < PH; ++y)
        for (int x = 0; x < PW; ++x) pad[(size_t)y * PW + x] = img[(size_t)ry[y] * W + rx[x]];
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double acc = 0.0;
            for (int j = 0; j < kh; ++j) {
                /* ker row j pairs with image row y - (j - cy) -> padded row y + 2cy - j */
                const double *prow = pad + (size_t)(y + 2 * cy - j) * PW + x + 2 * cx;
                const double *krow = ker + (size_t)j * kw;
                for (int i = 0; i < kw; ++i) acc += krow[i] * prow[-i];
            }
            out[(size_t)y * W + x] = acc;
        }
    free(pad); free(ry); free(rx);
}

/* Complex separable form of the same convolution: ker = gy (x) gx with
 * gy[j] = gy_re[j] + i gy_im[j], j = 0..2hy, and likewise gx.  Row pass on
 * in-bounds rows with reflected columns, then column pass with reflected rows —
 * identical to the 2-D reflect convolution. fp64 throughout. */
ORC_CLONES
ORC_API void orc_conv_sep_complex_reflect(const double *img, int H, int W,
                                          const double *gx_re, const double *gx_im, int hx,
                                          const double *gy_re, const double *gy_im, int hy,
                                          double *out_re, double *out_im)
{
    size_t N = (size_t)H * W;
    double *tr = malloc(sizeof(double) * N), *ti = malloc(sizeof(double) * N);
    int PW = W + 2 * hx;
    double *prow = malloc(sizeof(double) * (size_t)PW);
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < PW; ++x) prow[x] = img[(size_t)y * W + reflect_idx(x - hx, W)];
        for (int x = 0; x < W; ++x) {
            double ar = 0.0, ai = 0.0;
            const double *p = prow + x + 2 * hx; /* img[y, x - (i - hx)] = prow[x + 2hx - i] */
            for (int i = 0; i <= 2 * hx; ++i) {
                ar += gx_re[i] * p[-i];
                ai += gx_im[i] * p[-i];
            }
            tr[(size_t)y * W + x] = ar;
            ti[(size_t)y * W + x] = ai;
        }
    }
    int *ry = malloc(sizeof(int) * (size_t)(H + 2 * hy));
    for (int i = 0; i < H + 2 * hy; ++i) ry[i] = reflect_idx(i - hy, H);
    for (int y = 0; y < H; ++y) {
        double *orow = out_re + (size_t)y * W, *oiw = out_im + (size_t)y * W;
        for (int x = 0; x < W; ++x) { orow[x] = 0.0; oiw[x] = 0.0; }
        for (int j = 0; j <= 2 * hy; ++j) {
            size_t src = (size_t)ry[y + 2 * hy - j] * W; /* row y - (j - hy) */
            double gr = gy_re[j], gi = gy_im[j];
            const double *a = tr + src, *b = ti + src;
            for (int x = 0; x < W; ++x) {
                orow[x] += gr * a[x] - gi * b[x];
                oiw[x] += gr * b[x] + gi * a[x];
            }
        }
    }
    free(tr); free(ti); free(prow); free(ry);
}

/* ------------------------------------------------------------------------- */
/* k-means stage (builder-defined spec; parity unpinned)       */
/* ------------------------------------------------------------------------- */

#define KM_BLK 64

/*
 * Lloyd iterations with scipy.cluster.vq.kmeans2 ordering: for t < T:
 *   labels = assign(c_t); c_{t+1} = update(labels).  Returns the labels of the
 *   last assignment and c_T.
 * Arithmetic contract (restated bit-for-bit by the CUDA kernel):
 *   m_jd = -2 * c_jd (exact);  cn_j = (float) sum_d (double)c_jd*(double)c_jd, d ascending
 *   score_j(x) = fmaf(x_{D-1}, m_{j,D-1}, ... fmaf(x_0, m_{j0}, cn_j))   (fp32 FMA chain)
 *   label = lowest j attaining the minimum score (strict '<' scan, j ascending)
 *   q_d = lrintf(x_d * 2^fix_shift)  (round-to-nearest-even; |x_d| < 2^(31-fix_shift));
 *   sum_jd = exact int64 sum of q_d
 *   c_jd <- (float)((double)sum_jd / ((double)count_j * 2^fix_shift)); empty cluster keeps c_jd
 * feat is planar [D][N]; centroids [k][D]; k <= 64.
 */
/* Per-feature normalisation map: z_d = a_d x_d + b_d = (x_d - mean_d) / std_d over the image, from
 * EXACT integer moments so that the result does not depend on any summation order:
 *   r = lrintf(x * 2^16);  S1 = sum_p r,  S2 = sum_p r^2 (exact integers; S2 taken as (double)hi * 2^32 + (double)lo of
 *   its high and low 32-bit parts, which is how the kernels carry it)
 *   mean = S1 / (N 2^16),  E[x^2] = S2 / (N 2^32),  var = max(E[x^2] - mean^2, 0)   (double)
 *   a = (float)(1 / sqrt(var)), b = (float)(-mean / sqrt(var));  a = b = 0 when sqrt(var) <= 1e-12.
 * affine is [D][2]. */
ORC_API void orc_feature_affine(const float *feat, int D, int64_t N, int fix_shift, float *affine)
{
    (void)fix_shift;
    for (int d = 0; d < D; ++d) {
        const float *x = feat + (size_t)d * N;
        int64_t s1 = 0;
        unsigned __int128 s2 = 0;
        for (int64_t p = 0; p < N; ++p) {
            int64_t r = (int64_t)__builtin_lrintf(x[p] * 65536.0f);
            s1 += r;
            s2 += (uint64_t)(r * r);
        }
        /* the kernels carry S2 as two int64 sums of 32-bit halves of partial sums: hi * 2^32 + lo with lo possibly
         * above 2^32; any such split represents the same integer, but (double)hi * 2^32 + (double)lo rounds once per
         * term, so the canonical value is defined on the exact integer: round-to-nearest double of S2 */
        double mean = (double)s1 / ((double)N * 65536.0);
        double ex2 = (double)s2 / ((double)N * 4294967296.0);
        double var = ex2 - mean * mean;
        if (var < 0.0) var = 0.0;
        double sd = sqrt(var);
        float a = 0.f, b = 0.f;
        if (sd > 1e-12) { a = (float)(1.0 / sd); b = (float)(-mean / sd); }
        affine[2 * d] = a; affine[2 * d + 1] = b;
    }
}

/* centroid in the clustered space from its x-space value: (float)fma((double)a, (double)cx, (double)b) */
static inline float km_affine(const float *affine, int d, float cx)
{
    if (!affine) return cx;
    return (float)__builtin_fma((double)affine[2 * d], (double)cx, (double)affine[2 * d + 1]);
}

/* orc_kmeans with the optional normalisation folded into the score table (affine may be NULL):
 *   c_jd (z space) = km_affine(cx_jd);  m_jd = a_d * (-2 c_jd) (fp32);
 *   cn_j = (float) sum_d [ (double)c_jd^2 + (double)b_d * (double)(-2 c_jd) ]  (d ascending, c^2 first)
 *   score and update as in orc_kmeans, on the RAW features. */
ORC_CLONES
ORC_API void orc_kmeans_affine(const float *feat, int D, int64_t N, int k, int T, int fix_shift,
                               const int32_t *init_idx, const float *affine, int32_t *labels, float *centroids,
                               int64_t *counts_out)
{
    const float fix_scale = (float)(1u << fix_shift);
    float *m = malloc(sizeof(float) * (size_t)k * D);
    float *cn = malloc(sizeof(float) * (size_t)k);
    int64_t *sums = malloc(sizeof(int64_t) * (size_t)k * D);
    int64_t *cnt = malloc(sizeof(int64_t) * (size_t)k);
    float *s = malloc(sizeof(float) * (size_t)k * KM_BLK);
    for (int j = 0; j < k; ++j)
        for (int d = 0; d < D; ++d) centroids[(size_t)j * D + d] = km_affine(affine, d, feat[(size_t)d * N + init_idx[j]]);

    for (int t = 0; t < T; ++t) {
        for (int j = 0; j < k; ++j) {
            double acc = 0.0;
            for (int d = 0; d < D; ++d) {
                float c = centroids[(size_t)j * D + d];
                float mm = -2.0f * c;
                acc += (double)c * (double)c;
                if (affine) {
                    acc += (double)affine[2 * d + 1] * (-2.0 * (double)c);
                    mm = affine[2 * d] * mm;
                }
                m[(size_t)j * D + d] = mm;
            }
            cn[j] = (float)acc;
        }
        memset(sums, 0, sizeof(int64_t) * (size_t)k * D);
        memset(cnt, 0, sizeof(int64_t) * (size_t)k);
        for (int64_t p0 = 0; p0 < N; p0 += KM_BLK) {
            const int nb = (int)((N - p0) < KM_BLK ? (N - p0) : KM_BLK);
            float xb[KM_BLK];
            for (int j = 0; j < k; ++j)
                for (int i = 0; i < KM_BLK; ++i) s[j * KM_BLK + i] = cn[j];
            for (int d = 0; d < D; ++d) {
                const float *x = feat + (size_t)d * N + p0;
                for (int i = 0; i < KM_BLK; ++i) xb[i] = i < nb ? x[i] : 0.0f;
                for (int j = 0; j < k; ++j) {
                    const float mj = m[(size_t)j * D + d];
                    float *restrict sj = s + j * KM_BLK;
                    for (int i = 0; i < KM_BLK; ++i) sj[i] = __builtin_fmaf(xb[i], mj, sj[i]);
                }
            }
            for (int i = 0; i < nb; ++i) {
                int best = 0;
                float bs = s[i];
                for (int j = 1; j < k; ++j)
                    if (s[j * KM_BLK + i] < bs) { bs = s[j * KM_BLK + i]; best = j; }
                labels[p0 + i] = best;
                cnt[best] += 1;
            }
            for (int d = 0; d < D; ++d) {
                const float *x = feat + (size_t)d * N + p0;
                int64_t *sd = sums + d;
                for (int i = 0; i < nb; ++i)
                    sd[(size_t)labels[p0 + i] * D] += (int64_t)__builtin_lrintf(x[i] * fix_scale);
            }
        }
        for (int j = 0; j < k; ++j) {
            if (cnt[j] == 0) continue;
            double den = (double)cnt[j] * (double)fix_scale;
            for (int d = 0; d < D; ++d)
                centroids[(size_t)j * D + d] = km_affine(affine, d, (float)((double)sums[(size_t)j * D + d] / den));
        }
    }
    if (counts_out) memcpy(counts_out, cnt, sizeof(int64_t) * (size_t)k);
    free(m); free(cn); free(sums); free(cnt); free(s);
}

ORC_API void orc_kmeans(const float *feat, int D, int64_t N, int k, int T, int fix_shift,
                        const int32_t *init_idx, int32_t *labels, float *centroids,
                        int64_t *counts_out)
{
    orc_kmeans_affine(feat, D, N, k, T, fix_shift, init_idx, NULL, labels, centroids, counts_out);
}

/* One assignment pass only (teacher-forced tests): labels + best/second-best
 * fp64 squared distances so tests can classify near-ties. */
ORC_API void orc_kmeans_assign_f64(const float *feat, int D, int64_t N, int k,
                                   const float *centroids, int32_t *labels,
                                   double *best_d2, double *second_d2)
{
    for (int64_t p = 0; p < N; ++p) {
        double b1 = INFINITY, b2 = INFINITY;
        int bj = 0;
        for (int j = 0; j < k; ++j) {
            double acc = 0.0;
            for (int d = 0; d < D; ++d) {
                double df = (double)feat[(size_t)d * N + p] - (double)centroids[(size_t)j * D + d];
                acc += df * df;
            }
            if (acc < b1) { b2 = b1; b1 = acc; bj = j; }
            else if (acc < b2) b2 = acc;
        }
        labels[p] = bj; best_d2[p] = b1; second_d2[p] = b2;
    }
}

/* ------------------------------------------------------------------------- */
/* SLIC superpixels (the segmenter the reference actually calls,               */
/* BSD_metrics/script.py:11,30: skimage.segmentation.slic(img, n_segments=300, */
/* compactness=10.0)).  scikit-image is third-party, un-vendored, unpinned and */
/* absent from this image, so this follows its published algorithm             */
/* (slic_superpixels.py / _slic.pyx, 0.19 line) as restated in DESIGN.md 3.8:  */
/* PARITY UNPINNED.  float64 throughout, like scikit-image.                    */
/* ------------------------------------------------------------------------- */

/* cube root with IEEE operations only (bit-identical on the CPU and in the CUDA kernel; not correctly rounded):
 * frexp-style range reduction to [1/8, 1), a quadratic seed and five Newton steps y <- y - (y^3 - t) / (3 y^2) */
ORC_API double orc_det_cbrt(double t)
{
    if (t <= 0.0) return 0.0;
    int e = 0;
    double m = t;
    while (m >= 1.0) { m *= 0.125; e += 1; }
    while (m < 0.125) { m *= 8.0; e -= 1; }
    double y = 0.4928 + m * (0.8203 - m * 0.3131);         /* rough fit of cbrt on [1/8, 1) */
    for (int i = 0; i < 5; ++i) {
        double y2 = y * y;
        double num = y2 * y - m;
        double den = 3.0 * y2;
        y = y - num / den;
    }
    while (e > 0) { y *= 2.0; e -= 1; }
    while (e < 0) { y *= 0.5; e += 1; }
    return y;
}

/* skimage.color.rgb2lab (D65, 2 degree observer) of one 8-bit pixel, channels scaled by `ratio` = 1 / compactness */
static void slic_lab(const double *lin, const uint8_t *px, double ratio, double *out)
{
    const double r = lin[px[0]], g = lin[px[1]], b = lin[px[2]];
    double x = (0.412453 * r + 0.357580 * g) + 0.180423 * b;
    double y = (0.212671 * r + 0.715160 * g) + 0.072169 * b;
    double z = (0.019334 * r + 0.119193 * g) + 0.950227 * b;
    x = x / 0.95047; z = z / 1.08883;
    const double fx = x > 0.008856 ? orc_det_cbrt(x) : 7.787 * x + 16.0 / 116.0;
    const double fy = y > 0.008856 ? orc_det_cbrt(y) : 7.787 * y + 16.0 / 116.0;
    const double fz = z > 0.008856 ? orc_det_cbrt(z) : 7.787 * z + 16.0 / 116.0;
    out[0] = (116.0 * fy - 16.0) * ratio;
    out[1] = (500.0 * (fx - fy)) * ratio;
    out[2] = (200.0 * (fy - fz)) * ratio;
}

/* skimage.util.regular_grid for a (1, H, W) volume: start and step along y and x */
ORC_API void orc_slic_grid(int H, int W, int n_points, int *start_y, int *step_y, int *start_x, int *step_x)
{
    const double space = (double)H * (double)W;
    if (space <= (double)n_points) { *start_y = *start_x = 0; *step_y = *step_x = 1; return; }
    /* sorted dims (1, min, max): the unit depth axis is absorbed first, the other two share sqrt(space / n) */
    double s = sqrt(space / (double)n_points);
    const int lo = H < W ? H : W, hi = H < W ? W : H;
    double s_lo = s, s_hi = s;
    if ((double)lo < s) { s_lo = (double)lo; s_hi = (double)hi / (double)n_points; }
    const double sy = H < W ? s_lo : s_hi, sx = H < W ? s_hi : s_lo;
    const double use_y = (H == W) ? s : sy, use_x = (H == W) ? s : sx;
    *start_y = (int)floor(use_y / 2.0); *start_x = (int)floor(use_x / 2.0);
    *step_y = (int)nearbyint(use_y); *step_x = (int)nearbyint(use_x);
    if (*step_y < 1) *step_y = 1;
    if (*step_x < 1) *step_x = 1;
}

/* _enforce_label_connectivity_cython: raster scan, breadth-first growth of each 4-connected component up to
 * max_size pixels; components smaller than min_size take the label of the last adjacent, already relabelled
 * component seen during the search; the others get consecutive new labels from start_label. */
ORC_API void orc_slic_connectivity(const int32_t *seg, int H, int W, int min_size, int max_size, int start_label, int32_t *out)
{
    const int ddx[4] = {1, -1, 0, 0}, ddy[4] = {0, 0, 1, -1};
    int32_t *cy = malloc(sizeof(int32_t) * (size_t)(max_size > 0 ? max_size : 1)), *cx = malloc(sizeof(int32_t) * (size_t)(max_size > 0 ? max_size : 1));
    for (size_t i = 0; i < (size_t)H * W; ++i) out[i] = -1;
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
                        size += 1;
                        if (size >= max_size) break;
                    } else if (out[q] >= 0 && out[q] != cur) {
                        adjacent = out[q];
                    }
                }
                visited += 1;
            }
            if (size < min_size) {
                for (int i = 0; i < size; ++i) out[(size_t)cy[i] * W + cx[i]] = adjacent;
            } else {
                cur += 1;
            }
        }
    free(cy); free(cx);
}

/* slic(img, n_segments, compactness, max_num_iter, sigma=0, convert2lab=True, enforce_connectivity, min_size_factor=0.5,
 * max_size_factor=3, start_label): labels [H][W] int32.  lin256 = the sRGB -> linear table of the 256 channel values
 * (computed by the caller with libm pow, so that checker and product use the very same numbers). */
ORC_API int orc_slic(const uint8_t *img, int H, int W, int n_segments, double compactness, int max_iter,
                     int enforce, int start_label, const double *lin256, int32_t *labels)
{
    const size_t N = (size_t)H * W;
    int sy0, sy, sx0, sx;
    orc_slic_grid(H, W, n_segments, &sy0, &sy, &sx0, &sx);
    const int ny = (H - sy0 + sy - 1) / sy, nx = (W - sx0 + sx - 1) / sx, K = ny * nx;
    if (K < 1) return -1;
    const double ratio = 1.0 / compactness;
    const int step = sy > sx ? sy : sx;
    const double spatial_weight = 1.0 / ((double)step * (double)step);
    double *lab = malloc(sizeof(double) * N * 3), *segs = malloc(sizeof(double) * (size_t)K * 5), *dist = malloc(sizeof(double) * N);
    int32_t *near = calloc(N, sizeof(int32_t));
    int64_t *cnt = malloc(sizeof(int64_t) * (size_t)K);
    for (size_t p = 0; p < N; ++p) slic_lab(lin256, img + 3 * p, ratio, lab + 3 * p);
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            double *s = segs + (size_t)(j * nx + i) * 5;
            s[0] = (double)(sy0 + j * sy); s[1] = (double)(sx0 + i * sx); s[2] = s[3] = s[4] = 0.0;   /* colours start at zero */
        }
    for (int it = 0; it < max_iter; ++it) {
        for (size_t p = 0; p < N; ++p) dist[p] = 1.7976931348623157e308;
        for (int k = 0; k < K; ++k) {
            const double *s = segs + (size_t)k * 5;
            const double c_y = s[0], c_x = s[1];
            double t;
            t = c_y - 2.0 * sy; const long y_min = (long)(t > 0.0 ? t : 0.0);
            t = c_y + 2.0 * sy + 1.0; const long y_max = (long)(t < (double)H ? t : (double)H);
            t = c_x - 2.0 * sx; const long x_min = (long)(t > 0.0 ? t : 0.0);
            t = c_x + 2.0 * sx + 1.0; const long x_max = (long)(t < (double)W ? t : (double)W);
            for (long y = y_min; y < y_max; ++y) {
                const double dy = (c_y - (double)y) * (c_y - (double)y);
                for (long x = x_min; x < x_max; ++x) {
                    const double dx = (c_x - (double)x) * (c_x - (double)x);
                    double d = (dy + dx) * spatial_weight;
                    const double *l = lab + 3 * ((size_t)y * W + x);
                    double dc = 0.0;
                    for (int c = 0; c < 3; ++c) dc += (l[c] - s[2 + c]) * (l[c] - s[2 + c]);
                    d += dc;
                    if (dist[(size_t)y * W + x] > d) { near[(size_t)y * W + x] = k; dist[(size_t)y * W + x] = d; }
                }
            }
        }
        memset(cnt, 0, sizeof(int64_t) * (size_t)K);
        for (size_t i = 0; i < (size_t)K * 5; ++i) segs[i] = 0.0;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t p = (size_t)y * W + x;
                double *s = segs + (size_t)near[p] * 5;
                cnt[near[p]] += 1;
                s[0] += (double)y; s[1] += (double)x;
                for (int c = 0; c < 3; ++c) s[2 + c] += lab[3 * p + c];
            }
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < 5; ++c) segs[(size_t)k * 5 + c] /= (double)cnt[k];   /* 0/0 -> NaN, like scikit-image: the segment dies */
    }
    if (enforce) {
        const double segment_size = (double)N / (double)K;
        const int min_size = (int)(0.5 * segment_size), max_size = (int)(3.0 * segment_size);
        orc_slic_connectivity(near, H, W, min_size, max_size, start_label, labels);
    } else {
        for (size_t p = 0; p < N; ++p) labels[p] = near[p] + start_label;
    }
    free(lab); free(segs); free(dist); free(near); free(cnt);
    return K;
}
