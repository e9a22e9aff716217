/*
 * gcis.h — C ABI of libgcis.so, the B200 (sm_100a) implementation of the
 * gabor_color_image_segmentation hot path: Gabor bank -> per-pixel features ->
 * k-means labels -> BSD_metrics integer counts.
 *
 * The reference (pure Python) has no FFI layer; its boundary is a Python call
 * surface.  Each entry point below names the reference interface it stands
 * behind (paths relative to the reference checkout):
 *
 *   gcis_label_metrics_*    BSD_metrics/metrics.py:25-51   metrics.__init__  (GT boundary maps, n_segments)
 *                           BSD_metrics/metrics.py:58-74   set_boundary_recall
 *                           BSD_metrics/metrics.py:77-96   set_boundary_precision
 *                           BSD_metrics/metrics.py:102-146 set_undersegmentation
 *                           BSD_metrics/metrics.py:152-157 set_density
 *                           BSD_metrics/metrics.py:160-180 perimeter
 *                           BSD_metrics/metrics.py:182-201 set_compactness (integer part; floats are
 *                                                          finished by the host in reference order)
 *   gcis_segment_*          BSD_metrics/script.py:30       the segmenter slot (labels = f(img)); the
 *                                                          reference fills it with third-party SLIC and
 *                                                          holds no Gabor/k-means code, so these follow
 *                                                          DESIGN.md §3 (builder-defined spec)
 *   gcis_pipeline_host      BSD_metrics/script.py:22-38    one pass of the driver loop over a batch:
 *                                                          segment -> metrics(img, labels, segments)
 *
 * Conventions: every function returns 0 on success or a negative GCIS_E_* code;
 * gcis_last_error() gives the message for the calling thread.  Pointers named
 * d_* are device pointers, h_* host pointers.  `stream` is a cudaStream_t passed
 * as void* (NULL = legacy default stream).  Device entry points never
 * synchronise the host; host entry points return with results in host memory.
 * No function falls back to the CPU: without a CUDA device every compute entry
 * point fails with GCIS_E_CUDA.
 */
#ifndef GCIS_H
#define GCIS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GCIS_API __attribute__((visibility("default")))
#else
#define GCIS_API
#endif

#define GCIS_VERSION 101

#define GCIS_OK 0
#define GCIS_E_INVALID (-1) /* bad argument                                   */
#define GCIS_E_CUDA (-2)    /* CUDA runtime/driver error (message has details) */
#define GCIS_E_LABEL (-3)   /* negative label or label >= capacity             */
#define GCIS_E_NOMEM (-4)

/* colour spaces for the filter-bank input */
#define GCIS_COLOUR_RGB 0
#define GCIS_COLOUR_OPPONENT 1
#define GCIS_COLOUR_LAB 2
/* per-filter feature */
#define GCIS_FEATURE_MAGNITUDE 0
#define GCIS_FEATURE_ENERGY 1

/* Number of int64 slots per (image, ground truth) in `gt_counts`:
 *   [0] den_r = |bd(gt_g)|                      metrics.py:72
 *   [1] tp_r  = |dil_size(bd(lb)) & bd(gt_g)|   metrics.py:69-72
 *   [2] tp_p  = |bd(lb) & dil_5(bd(gt_g))|      metrics.py:91-94
 *   [3] U_g   = sum_i (area_i - max_j hist_ij)  metrics.py:129-131
 *   [4] V_g   = sum_ij min(hist_ij, area_i - hist_ij)  metrics.py:137-140
 *   [5] sum_ij hist_ij^2                         (PRI extension)
 *   [6] sum_j  colsum_j^2                        (PRI extension)
 *   [7] reserved                                                             */
#define GCIS_GT_SLOTS 8

/* status bits written per image by the label-metrics kernels */
#define GCIS_ST_NEG_LABEL 1
#define GCIS_ST_SEG_OVER 2 /* lb  >= n_seg_cap */
#define GCIS_ST_LAB_OVER 4 /* gt  >= n_lab_cap */

typedef struct gcis_config {
    int32_t height;        /* image rows    (reference nx, metrics.py:44) */
    int32_t width;         /* image columns (reference ny)                */
    int32_t max_batch;     /* images per call (capacity of the workspaces) */
    int32_t colour_space;  /* GCIS_COLOUR_*                                */
    int32_t n_scales;
    int32_t n_orient;
    const double *frequencies; /* [n_scales] cycles/pixel                  */
    const double *thetas;      /* [n_orient] radians                       */
    double bandwidth;      /* octaves (scikit-image gabor_kernel default 1) */
    double n_stds;         /* truncation (default 3)                        */
    int32_t feature;       /* GCIS_FEATURE_*                                */
    int32_t k;             /* clusters, 1..32                               */
    int32_t iters;         /* Lloyd iterations T (fixed, no early stop)     */
    int32_t fix_shift;     /* centroid sums are exact int64 sums of rint(x * 2^fix_shift); default 24 */
    int32_t max_gt;        /* ground truths per image (capacity G)          */
    int32_t n_lab_cap;     /* capacity for max(gt)+1                        */
    int32_t dil_recall;    /* `size` of set_boundary_recall (default 5)     */
    int32_t group;         /* images per launch group, 0 = auto (64; sized for occupancy, not for L2 residency) */
    /* feature assembly options (north_star "optional smoothing and normalisation") */
    int32_t normalise;     /* 1: cluster on per-feature z-scores (folded into the k-means score table) */
    double smooth;         /* > 0: Gaussian smoothing of every magnitude plane, sigma = smooth * sigma_s  */
} gcis_config;

typedef struct gcis_plan gcis_plan;

/* ---- library ---- */
GCIS_API int32_t gcis_version(void);
GCIS_API const char *gcis_last_error(void);
/* kernels launched by this library in this process so far (for bench.py's gpu_launches) */
GCIS_API int64_t gcis_launch_count(void);
/* number of visible CUDA devices, or a negative error */
GCIS_API int32_t gcis_device_count(void);

/* ---- filter bank description (host only; no device needed) ----
 * Half-width of the (2h+1)x(2h+1) kernel of (frequency, theta), the scikit-image
 * gabor_kernel extent with sigma_x = sigma_y. */
GCIS_API int32_t gcis_gabor_half_width(double frequency, double theta, double bandwidth, double n_stds);
/* Complex 1-D factors gx, gy (each 2h+1 taps, index t <-> offset t-h) with
 * kernel[y][x] = gy[y]*gx[x]; returns h or a negative error. */
GCIS_API int32_t gcis_gabor_separable(double frequency, double theta, double bandwidth, double n_stds,
                             double *gx_re, double *gx_im, double *gy_re, double *gy_im,
                             int32_t cap_taps);

/* ---- plan: bank tables + device workspaces for one (H, W, config) ---- */
GCIS_API int32_t gcis_plan_create(const gcis_config *cfg, gcis_plan **out);
GCIS_API void gcis_plan_destroy(gcis_plan *plan);
GCIS_API int32_t gcis_plan_feature_dim(const gcis_plan *plan); /* D = 3 * n_scales * n_orient */
GCIS_API int64_t gcis_plan_workspace_bytes(const gcis_plan *plan);
/* images per kernel launch when a batch of B images runs through the plan (the batch is cut into equal groups) */
GCIS_API int32_t gcis_plan_launch_group(const gcis_plan *plan, int32_t B);
/* 1 when the filter bank's row pass runs on the tcgen05 tensor cores (rgb planes, exact in bf16), 0 when both
 * passes run on the FP32 pipe (opponent / Lab planes, GCIS_GABOR_TC=0, or a bank beyond the kernel's constant tap
 * table: more than 8 orientation jobs per scale or more than 6144 complex column taps) */
GCIS_API int32_t gcis_plan_uses_tensor_cores(const gcis_plan *plan);

/* ---- segmenter slot (script.py:30), device pointers ----
 * d_img      [B][H][W][3] uint8 (interleaved RGB, as skimage.io.imread returns it)
 * d_feat     [B][D][H][W] float32, d = (c*S + s)*O + o
 * d_init_idx [B][k] int32 pixel indices (y*W + x) of the initial centroids
 * d_labels   [B][H][W] int32 in 0..k-1
 * d_centroids[B][k][D] float32 (may be NULL) */
GCIS_API int32_t gcis_gabor_features(gcis_plan *plan, const uint8_t *d_img, int32_t B, float *d_feat, void *stream);
GCIS_API int32_t gcis_kmeans(gcis_plan *plan, const float *d_feat, int32_t B, const int32_t *d_init_idx,
                    int32_t *d_labels, float *d_centroids, void *stream);
/* Per-feature normalisation map of a feature tensor (plans created with normalise = 1):
 * d_affine [B][D][2] float32 {a_d, b_d} with z_d = a_d * x_d + b_d = (x_d - mean_d) / std_d over the image,
 * mean and std from exact integer moments; a = b = 0 for a constant plane.  With normalise = 1
 * gcis_kmeans / gcis_segment_device cluster on z and d_centroids are centroids in z space. */
GCIS_API int32_t gcis_feature_affine(gcis_plan *plan, const float *d_feat, int32_t B, float *d_affine, void *stream);
GCIS_API int32_t gcis_segment_device(gcis_plan *plan, const uint8_t *d_img, int32_t B, const int32_t *d_init_idx,
                            int32_t *d_labels, void *stream);

/* ---- BSD_metrics integer counts (metrics.py:25-201), device pointers ----
 * d_lb        [B][H][W] int32 labels (>= 0)
 * d_gt        [B][G][H][W] uint16 ground-truth label maps (groundtruth.py:26 dtype)
 * d_n_gt      [B] int32 number of valid ground truths per image (NULL = G for all)
 * outputs (all zeroed by the call):
 * d_bd_count  [B] int64              |bd(lb)|
 * d_gt_counts [B][G][GCIS_GT_SLOTS] int64
 * d_area      [B][n_seg_cap] int32,  d_perim [B][n_seg_cap] int32
 * d_hist      [B][G][n_seg_cap][n_lab_cap] int32 contingency tables
 * d_n_seg     [B] int32 max(lb)+1,   d_n_lab [B][G] int32 max(gt)+1
 * d_status    [B] int32 GCIS_ST_* bits (0 = ok)
 */
GCIS_API int32_t gcis_label_metrics_device(const int32_t *d_lb, const uint16_t *d_gt, const int32_t *d_n_gt,
                                  int32_t B, int32_t H, int32_t W, int32_t G,
                                  int32_t n_seg_cap, int32_t n_lab_cap, int32_t dil_recall,
                                  int64_t *d_bd_count, int64_t *d_gt_counts,
                                  int32_t *d_area, int32_t *d_perim, int32_t *d_hist,
                                  int32_t *d_n_seg, int32_t *d_n_lab, int32_t *d_status, void *stream);

/* Same with host buffers: allocates scratch, copies in, runs, copies out, frees.
 * h_hist may be NULL.  Returns GCIS_E_LABEL if any image's status is non-zero
 * (outputs are still copied back). */
GCIS_API int32_t gcis_label_metrics_host(const int32_t *h_lb, const uint16_t *h_gt, const int32_t *h_n_gt,
                                int32_t B, int32_t H, int32_t W, int32_t G,
                                int32_t n_seg_cap, int32_t n_lab_cap, int32_t dil_recall,
                                int64_t *h_bd_count, int64_t *h_gt_counts,
                                int32_t *h_area, int32_t *h_perim, int32_t *h_hist,
                                int32_t *h_n_seg, int32_t *h_n_lab, int32_t *h_status);

/* find_boundaries(x) of whole label maps (metrics.py:47-49, the `img_truth` attribute):
 * h_x [B][H][W] int32 -> h_out [B][H][W] uint8 (1 = boundary). */
GCIS_API int32_t gcis_find_boundaries_host(const int32_t *h_x, int32_t B, int32_t H, int32_t W, uint8_t *h_out);

/* ---- whole path, device-resident inputs (bench `value`) ----
 * Runs segmenter + metrics for B <= max_batch images whose inputs are already in
 * HBM; results stay in the plan's device buffers until gcis_pipeline_fetch. */
GCIS_API int32_t gcis_pipeline_device(gcis_plan *plan, const uint8_t *d_img, const uint16_t *d_gt,
                             const int32_t *d_n_gt, const int32_t *d_init_idx, int32_t B, void *stream);
/* Copies the last gcis_pipeline_device results to the host (synchronises `stream`).
 * h_area/h_perim are [B][k]; h_labels [B][H][W] may be NULL. */
GCIS_API int32_t gcis_pipeline_fetch(gcis_plan *plan, int32_t B, int64_t *h_bd_count, int64_t *h_gt_counts,
                            int32_t *h_area, int32_t *h_perim, int32_t *h_n_lab, int32_t *h_status,
                            int32_t *h_labels, void *stream);

/* Contingency tables of the last pipeline call (metrics.py:115-126), [B][G][k][n_lab_cap] int32:
 * the input of the region scores that go beyond the reference (PRI, VoI, covering). */
GCIS_API int32_t gcis_pipeline_fetch_hist(gcis_plan *plan, int32_t B, int32_t *h_hist, void *stream);

/* ---- whole path, host buffers (bench `e2e`; script.py:22-38 for a batch) ----
 * h_img [B][H][W][3] uint8, h_gt [B][G][H][W] uint16, h_n_gt [B] or NULL,
 * h_init_idx [B][k].  B may exceed max_batch; the call streams chunks of
 * max_batch images through pinned staging buffers, overlapping copies and
 * kernels.  Outputs as in gcis_pipeline_fetch. */
GCIS_API int32_t gcis_pipeline_host(gcis_plan *plan, const uint8_t *h_img, const uint16_t *h_gt,
                           const int32_t *h_n_gt, const int32_t *h_init_idx, int32_t B,
                           int64_t *h_bd_count, int64_t *h_gt_counts, int32_t *h_area,
                           int32_t *h_perim, int32_t *h_n_lab, int32_t *h_status, int32_t *h_labels);

/* ---- image decode (BSD_metrics/script.py:25, `img = imread(...)`) ----
 * Baseline / extended-sequential 8-bit Huffman JPEG with 1 or 3 components (luma sampling 1x1, 2x1 or 2x2) ->
 * the pixels libjpeg's defaults produce (islow inverse DCT, fancy upsampling, fixed-point YCbCr -> RGB), bit for
 * bit.  Entropy decoding runs on host threads, everything after it on the GPU.  Other JPEG flavours are
 * rejected with GCIS_E_INVALID (decode them on the host as before). */
GCIS_API int32_t gcis_jpeg_info(const uint8_t *data, int64_t size, int32_t *h, int32_t *w, int32_t *components);
/* quantised coefficients of one file, per component [blocks_h][blocks_w][64] in natural order (host only; used by the
 * tests to check the device stages in isolation); returns the count or a negative error */
GCIS_API int64_t gcis_jpeg_coefficients(const uint8_t *data, int64_t size, int16_t *coef, int64_t cap);
/* files[i], sizes[i]: B files of one frame size H x W in host memory; d_rgb [B][H][W][3] uint8 on the device, the
 * layout gcis_segment_device / gcis_pipeline_device take.  n_threads <= 0: one host thread per hardware thread.
 * Returns after `stream` has finished. */
GCIS_API int32_t gcis_jpeg_decode_batch(const uint8_t *const *files, const int64_t *sizes, int32_t B, int32_t H,
                                        int32_t W, uint8_t *d_rgb, int32_t n_threads, void *stream);

/* ---- SLIC superpixels: the segmenter the reference itself calls (BSD_metrics/script.py:11,30:
 * skimage.segmentation.slic(img, n_segments=300, compactness=10.0)).  scikit-image's published algorithm in float64
 * (third-party and unpinned upstream: parity unpinned).  One H x W x 3 uint8 image in host memory ->
 * h_labels [H][W] int32; assignment and centroid update run on the GPU, the connectivity pass on the host.
 * Returns the number of seeds or a negative error. */
GCIS_API int32_t gcis_slic_host(const uint8_t *h_img, int32_t H, int32_t W, int32_t n_segments, double compactness,
                                int32_t max_iter, int32_t enforce_connectivity, int32_t start_label, int32_t *h_labels);

/* Stage timings (ms, CUDA events on the plan's stream) of the last
 * gcis_pipeline_device call when profiling is enabled: [colour, gabor, kmeans, metrics]. */
GCIS_API int32_t gcis_plan_set_profiling(gcis_plan *plan, int32_t on);
GCIS_API int32_t gcis_plan_last_stage_ms(gcis_plan *plan, float *ms4);

#ifdef __cplusplus
}
#endif
#endif /* GCIS_H */
