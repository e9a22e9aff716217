// plan.cu — the C ABI (include/gcis.h): plan object, stage entry points and the batch driver.
//
// The batch driver stands behind the reference's driver loop (BSD_metrics/script.py:22-38):
// segment every image, then score it against its ground truths.  Images are processed in launch
// groups of up to 64 (cut to equal sizes).  Measured on B200: the 126 MB L2 cannot hold even two
// images' features (44.5 MB per 321x481 image) across a k-means pass, so the passes are HBM streams
// by design and the group size is chosen for occupancy (more CTAs per launch), not for L2 residency.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "gabor.cuh"

namespace gcis {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};

int set_error(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// implemented in the stage files
struct GaborLaunchPlan;
int colour_planes_launch(const uint8_t *, float *, int, int, int, int, int, int, cudaStream_t);
int label_metrics_launch(const int32_t *, const uint16_t *, const int32_t *, int, int, int, int, int, int, int,
                         int64_t *, int64_t *, int32_t *, int32_t *, int32_t *, int32_t *, int32_t *, int32_t *,
                         cudaStream_t);
int find_boundaries_launch(const int32_t *, uint8_t *, int, int, int, cudaStream_t);
size_t kmeans_workspace_bytes(int B, int D, int N, int k);
int kmeans_launch(const float *, size_t, int, int, int, int, int, int, int, const int32_t *, int32_t *, float *,
                  void *, cudaStream_t, const float *d_affine);
// feature assembly (features.cu): optional smoothing, normalisation statistics
struct SmoothPlan;
SmoothPlan *smooth_plan_new(const double *sigmas, int S, int O, int D, int H, int W, double factor);
void smooth_plan_delete(SmoothPlan *);
int smooth_launch(SmoothPlan &, float *, size_t, int, float *, int, long long *, float, cudaStream_t);
int feature_moments_launch(const float *, size_t, int, int, int, int, float, long long *, cudaStream_t);
int feature_affine_launch(const long long *, float *, int, int, int, float, cudaStream_t);
GaborLaunchPlan *gabor_plan_new(const GaborBankHost &, int H, int W, int C, int P, int Wp, int feature, size_t *smem);
void gabor_plan_delete(GaborLaunchPlan *);
int gabor_launch(GaborLaunchPlan &, const float *, float *, const float *, const GaborScale *, int, int, cudaStream_t,
                 long long *d_stats, float stat_scale);
// tensor-core row pass (gabor_tc.cu); plan_new returns nullptr when the configuration is not covered
struct GaborTcPlan;
GaborTcPlan *gabor_tc_plan_new(const GaborBankHost &, int H, int W, int C, int P, int Wp16, int feature, int colour_space);
void gabor_tc_plan_delete(GaborTcPlan *);
size_t gabor_tc_plan_bytes(const GaborTcPlan *);
int colour_planes16_launch(const uint8_t *, void *, int, int, int, int, int, cudaStream_t);
int gabor_tc_launch(GaborTcPlan &, const void *, float *, const float *, const GaborScale *, int, int, cudaStream_t,
                    long long *d_stats, float stat_scale);

}  // namespace gcis

using namespace gcis;

struct gcis_plan {
    gcis_config cfg;
    std::vector<double> freqs, thetas;
    GaborBankHost bank;
    GaborLaunchPlan *glp = nullptr;
    GaborTcPlan *gtc = nullptr;  // non-null: the filter bank runs its row pass on the tensor cores
    SmoothPlan *smooth = nullptr;  // non-null: Gaussian smoothing of the magnitude planes (cfg.smooth > 0)
    float *d_tmp[2] = {nullptr, nullptr};        // [group][D][N] scratch of the smoothing's row pass
    long long *d_stats[2] = {nullptr, nullptr};  // [group][D][GB_STAT_SLOTS] integer moments (cfg.normalise)
    float *d_affine[2] = {nullptr, nullptr};     // [group][D][2] z-score map a, b
    int D = 0, N = 0, Np = 0, P = 0, Wp = 0, Wp16 = 0, group = 1;
    size_t bytes = 0;
    // device workspaces
    float *d_taps = nullptr;
    GaborScale *d_scales = nullptr;
    // Two "lanes" of per-group workspaces: group g runs on lane g & 1, so the FP32-bound Gabor
    // kernel of one group overlaps the HBM-bound k-means passes of the previous one.
    float *d_planes[2] = {nullptr, nullptr};   // [group][3][H][Wp] f32, or [group][3][H][Wp16] bf16 on the tensor-core path
    float *d_feat[2] = {nullptr, nullptr};     // [group][D][Np]
    void *d_km_ws[2] = {nullptr, nullptr};
    cudaStream_t lane_stream[2] = {nullptr, nullptr};
    cudaEvent_t lane_done[2] = {nullptr, nullptr}, ev_fork = nullptr;
    int n_lanes = 1;
    int32_t *d_labels = nullptr; // [max_batch][N]
    int64_t *d_bd_count = nullptr, *d_gt_counts = nullptr;
    int32_t *d_area = nullptr, *d_perim = nullptr, *d_hist = nullptr, *d_n_seg = nullptr, *d_n_lab = nullptr,
            *d_status = nullptr;
    // device input staging for the host entry point
    uint8_t *d_img = nullptr;
    uint16_t *d_gt = nullptr;
    int32_t *d_n_gt = nullptr, *d_init = nullptr;
    bool host_ready = false;     // staging buffers, copy stream and events of gcis_pipeline_host exist
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_gt[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    // profiling
    bool profiling = false;
    std::vector<cudaEvent_t> events;  // 5 per group: colour | gabor | kmeans | (metrics: 2 at the end)
    float stage_ms[4] = {0, 0, 0, 0};
    int n_groups_last = 0;
};

namespace {

template <typename T>
int dev_alloc(T **p, size_t n, size_t *total)
{
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), bytes);
    if (e != cudaSuccess) {
        *p = nullptr;
        return set_error(e == cudaErrorMemoryAllocation ? GCIS_E_NOMEM : GCIS_E_CUDA, "cudaMalloc(%zu) -> %s", bytes,
                         cudaGetErrorString(e));
    }
    *total += bytes;
    return GCIS_OK;
}

#define TRY(x)                 \
    do {                       \
        int _rc = (x);         \
        if (_rc) return _rc;   \
    } while (0)

int plan_check_batch(const gcis_plan *p, int B)
{
    if (!p) return set_error(GCIS_E_INVALID, "null plan");
    if (B < 1 || B > p->cfg.max_batch) return set_error(GCIS_E_INVALID, "B=%d outside 1..max_batch=%d", B, p->cfg.max_batch);
    return GCIS_OK;
}

cudaEvent_t plan_event(gcis_plan *p, size_t i)
{
    while (p->events.size() <= i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        p->events.push_back(e);
    }
    return p->events[i];
}

// Images per launch group for a batch of B: the batch is cut into ceil(B / capacity) groups of equal size
// rather than full groups plus a small straggler (200 images at capacity 64: 4 x 50, not 3 x 64 + 8).
static int balanced_group(const gcis_plan *p, int B)
{
    const int n = ceil_div(std::max(B, 1), p->group);
    return ceil_div(std::max(B, 1), n);
}

// colour -> Gabor -> k-means for one group of images whose features live in plan->d_feat.
int segment_group(gcis_plan *p, int lane, const uint8_t *d_img, int nb, const int32_t *d_init, int32_t *d_labels,
                  float *d_feat_out, cudaStream_t st, int group_index)
{
    const gcis_config &c = p->cfg;
    // the plan's own feature buffer pads every plane to Np floats so the k-means pass can use
    // aligned 128-bit loads; a caller-supplied tensor is dense [D][H][W]
    float *feat = d_feat_out ? d_feat_out : p->d_feat[lane];
    const int pstride = d_feat_out ? p->N : p->Np;
    const bool prof = p->profiling && group_index >= 0;
    if (prof) cudaEventRecord(plan_event(p, 4 * group_index + 0), st);
    if (p->gtc) TRY(colour_planes16_launch(d_img, p->d_planes[lane], nb, c.height, c.width, p->P, p->Wp16, st));
    else TRY(colour_planes_launch(d_img, p->d_planes[lane], nb, c.height, c.width, p->P, p->Wp, c.colour_space, st));
    if (prof) cudaEventRecord(plan_event(p, 4 * group_index + 1), st);
    const float stat_scale = (float)(1u << c.fix_shift);
    const bool norm = c.normalise != 0 && d_labels != nullptr;
    long long *stats = norm ? p->d_stats[lane] : nullptr;
    if (norm) GCIS_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(long long) * (size_t)nb * p->D * GB_STAT_SLOTS, st));
    // The moments are taken from what the clustering will read: in the smoothing's epilogue, or in the filter bank's
    // (no extra HBM traffic; measured +2.7 ms per 200 images), or with GCIS_STATS_FUSED=0 by a streaming kernel over the
    // finished features (measured +3.7 ms: slower, kept for A/B runs and for caller-supplied feature tensors).
    static const bool fused = [] { const char *e = getenv("GCIS_STATS_FUSED"); return !e || atoi(e) != 0; }();
    long long *gabor_stats = (p->smooth || !fused) ? nullptr : stats;
    if (p->gtc) TRY(gabor_tc_launch(*p->gtc, p->d_planes[lane], feat, p->d_taps, p->d_scales, nb, pstride, st, gabor_stats, stat_scale));
    else TRY(gabor_launch(*p->glp, p->d_planes[lane], feat, p->d_taps, p->d_scales, nb, pstride, st, gabor_stats, stat_scale));
    if (p->smooth) TRY(smooth_launch(*p->smooth, feat, (size_t)p->D * pstride, pstride, p->d_tmp[lane], nb, stats, stat_scale, st));
    else if (norm && !fused) TRY(feature_moments_launch(feat, (size_t)p->D * pstride, pstride, nb, p->D, p->N, stat_scale, stats, st));
    if (prof) cudaEventRecord(plan_event(p, 4 * group_index + 2), st);
    if (d_labels) {
        if (norm) TRY(feature_affine_launch(stats, p->d_affine[lane], nb, p->D, p->N, stat_scale, st));
        TRY(kmeans_launch(feat, (size_t)p->D * pstride, pstride, nb, p->D, p->N, c.k, c.iters, c.fix_shift, d_init,
                          d_labels, nullptr, p->d_km_ws[lane], st, norm ? p->d_affine[lane] : nullptr));
        if (prof) cudaEventRecord(plan_event(p, 4 * group_index + 3), st);
    }
    return GCIS_OK;
}

// Segmenter + metrics for `nb` images whose results land at image offset `off` of the plan's
// result buffers (used by the host entry point to pipeline sub-chunks).
// `gt_ready`, when given, is waited for just before the metrics kernels: the ground truths of a
// host sub-chunk are uploaded behind its images, while the segmenter already runs.
int pipeline_range(gcis_plan *p, int lane, const uint8_t *d_img, const uint16_t *d_gt, const int32_t *d_n_gt,
                   const int32_t *d_init, int off, int nb, cudaStream_t st, cudaEvent_t gt_ready = nullptr)
{
    const gcis_config &c = p->cfg;
    const size_t N = p->N, G = std::max(c.max_gt, 1);
    int32_t *labels = p->d_labels + (size_t)off * N;
    const int grp = balanced_group(p, nb);
    for (int b0 = 0; b0 < nb; b0 += grp) {
        const int n = std::min(grp, nb - b0);
        TRY(segment_group(p, lane, d_img + (size_t)b0 * N * 3, n, d_init + (size_t)b0 * c.k, labels + (size_t)b0 * N, nullptr, st, -1));
    }
    if (gt_ready) GCIS_CUDA_TRY(cudaStreamWaitEvent(st, gt_ready, 0));
    return label_metrics_launch(labels, d_gt, d_n_gt, nb, c.height, c.width, c.max_gt, c.k, c.n_lab_cap, c.dil_recall,
                                p->d_bd_count + off, p->d_gt_counts + (size_t)off * G * GCIS_GT_SLOTS,
                                p->d_area + (size_t)off * c.k, p->d_perim + (size_t)off * c.k,
                                p->d_hist + (size_t)off * G * c.k * c.n_lab_cap, p->d_n_seg + off,
                                p->d_n_lab + (size_t)off * G, p->d_status + off, st);
}

}  // namespace

// Grow-only device scratch of the calling thread for the host-buffer entry points below: one allocation that is
// reused by later calls (a drop-in `metrics(...)` call used to pay 11 cudaMalloc + 11 cudaFree, each of which
// synchronises the device).  Freed when the thread exits.
namespace {
struct ScratchArena {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    int device = -1;
    ~ScratchArena() { if (base) cudaFree(base); }
    int reserve(size_t bytes)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        if (base && (dev != device || bytes > cap)) { cudaFree(base); base = nullptr; cap = 0; }
        if (!base) {
            const size_t want = std::max<size_t>(bytes, 1 << 20);
            cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&base), want);
            if (e != cudaSuccess) {
                base = nullptr;
                return set_error(e == cudaErrorMemoryAllocation ? GCIS_E_NOMEM : GCIS_E_CUDA, "cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
            }
            cap = want; device = dev;
        }
        used = 0;
        return GCIS_OK;
    }
    template <typename T>
    T *take(size_t n)
    {
        used = (used + 255) & ~(size_t)255;
        T *p = reinterpret_cast<T *>(base + used);
        used += std::max<size_t>(n, 1) * sizeof(T);
        return p;
    }
};
thread_local ScratchArena g_arena;
inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }
}  // namespace

extern "C" {

int32_t gcis_version(void) { return GCIS_VERSION; }
const char *gcis_last_error(void) { return g_last_error.c_str(); }
int64_t gcis_launch_count(void) { return g_launches.load(); }

int32_t gcis_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return set_error(GCIS_E_CUDA, "cudaGetDeviceCount -> %s", cudaGetErrorString(e));
    return n;
}

int32_t gcis_gabor_half_width(double frequency, double theta, double bandwidth, double n_stds)
{
    if (!(frequency > 0) || !(bandwidth > 0) || !(n_stds > 0)) return set_error(GCIS_E_INVALID, "gabor: bad parameters");
    return gabor_half_width(frequency, theta, bandwidth, n_stds);
}

int32_t gcis_gabor_separable(double frequency, double theta, double bandwidth, double n_stds, double *gx_re,
                             double *gx_im, double *gy_re, double *gy_im, int32_t cap_taps)
{
    if (!(frequency > 0) || !(bandwidth > 0) || !(n_stds > 0)) return set_error(GCIS_E_INVALID, "gabor: bad parameters");
    std::vector<double> a, b, c, d;
    const int h = gabor_separable(frequency, theta, bandwidth, n_stds, a, b, c, d);
    if (2 * h + 1 > cap_taps) return set_error(GCIS_E_INVALID, "gabor: %d taps needed, capacity %d", 2 * h + 1, cap_taps);
    memcpy(gx_re, a.data(), sizeof(double) * a.size());
    memcpy(gx_im, b.data(), sizeof(double) * b.size());
    memcpy(gy_re, c.data(), sizeof(double) * c.size());
    memcpy(gy_im, d.data(), sizeof(double) * d.size());
    return h;
}

int32_t gcis_plan_create(const gcis_config *cfg, gcis_plan **out)
{
    if (!cfg || !out) return set_error(GCIS_E_INVALID, "plan: null argument");
    *out = nullptr;
    if (cfg->n_scales < 1 || cfg->n_orient < 1)
        return set_error(GCIS_E_INVALID, "plan: n_scales=%d n_orient=%d must be >= 1", cfg->n_scales, cfg->n_orient);
    if (cfg->height < 1 || cfg->width < 1 || cfg->max_batch < 1)
        return set_error(GCIS_E_INVALID, "plan: bad shape H=%d W=%d max_batch=%d", cfg->height, cfg->width, cfg->max_batch);
    if ((int64_t)cfg->height * cfg->width > (1 << 30)) return set_error(GCIS_E_INVALID, "plan: image too large");
    if (!cfg->frequencies || !cfg->thetas) return set_error(GCIS_E_INVALID, "plan: null bank");
    if (cfg->colour_space < 0 || cfg->colour_space > 2) return set_error(GCIS_E_INVALID, "plan: colour_space=%d", cfg->colour_space);
    if (cfg->feature < 0 || cfg->feature > 1) return set_error(GCIS_E_INVALID, "plan: feature=%d", cfg->feature);
    if (cfg->k < 1 || cfg->k > 32) return set_error(GCIS_E_INVALID, "plan: k=%d outside 1..32", cfg->k);
    if (cfg->iters < 1) return set_error(GCIS_E_INVALID, "plan: iters=%d", cfg->iters);
    if (cfg->max_gt < 0 || cfg->n_lab_cap < 1) return set_error(GCIS_E_INVALID, "plan: max_gt=%d n_lab_cap=%d", cfg->max_gt, cfg->n_lab_cap);
    if (!(cfg->bandwidth > 0) || !(cfg->n_stds > 0)) return set_error(GCIS_E_INVALID, "plan: bandwidth/n_stds");
    if (!(cfg->smooth >= 0) || cfg->smooth > 8) return set_error(GCIS_E_INVALID, "plan: smooth=%g outside 0..8", cfg->smooth);
    if (cfg->normalise && (int64_t)cfg->height * cfg->width > (1 << 25))
        return set_error(GCIS_E_INVALID, "plan: normalisation supports up to 2^25 pixels per image");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return set_error(GCIS_E_CUDA, "plan: no CUDA device (this library has no CPU path)");

    gcis_plan *p = new gcis_plan();
    p->cfg = *cfg;
    p->freqs.assign(cfg->frequencies, cfg->frequencies + cfg->n_scales);
    p->thetas.assign(cfg->thetas, cfg->thetas + cfg->n_orient);
    p->cfg.frequencies = p->freqs.data();
    p->cfg.thetas = p->thetas.data();
    if (p->cfg.fix_shift <= 0) p->cfg.fix_shift = 24;
    if (p->cfg.dil_recall <= 0) p->cfg.dil_recall = 5;
    int rc = build_bank(p->freqs.data(), cfg->n_scales, p->thetas.data(), cfg->n_orient, cfg->bandwidth, cfg->n_stds, p->bank);
    if (rc) { delete p; return rc; }
    const int H = cfg->height, W = cfg->width;
    p->N = H * W;
    p->D = 3 * cfg->n_scales * cfg->n_orient;
    p->P = p->bank.hmax;
    // wide enough that the Gabor staging never needs a column guard (gabor.cu: fetch)
    p->Wp = round_up(round_up(W, 32) + 2 * p->P + 40, 4);
    p->Np = round_up(p->N, 32);
    int group = cfg->group;
    if (const char *e = getenv("GCIS_GROUP")) group = atoi(e);
    // Measured on B200 (profiles/): larger groups fill the machine better and the 126 MB L2 cannot
    // hold even two images' features across a pass, so the default favours occupancy.
    if (group <= 0) group = 64;
    p->group = std::min(group, cfg->max_batch);
    size_t smem = 0;
    p->glp = gabor_plan_new(p->bank, H, W, 3, p->P, p->Wp, cfg->feature, &smem);
    if (!p->glp) { delete p; return GCIS_E_INVALID; }
    p->Wp16 = round_up(p->Wp, 8);
    {   // GCIS_GABOR_TC=0 keeps both passes on the FP32 pipe (A/B measurements)
        const char *e = getenv("GCIS_GABOR_TC");
        if (!e || atoi(e) != 0) {
            p->gtc = gabor_tc_plan_new(p->bank, H, W, 3, p->P, p->Wp16, cfg->feature, cfg->colour_space);
            p->bytes += gabor_tc_plan_bytes(p->gtc);
        }
    }

    const size_t MB = cfg->max_batch, G = std::max(cfg->max_gt, 1), k = cfg->k;
    auto fail = [&](int code) { gcis_plan_destroy(p); return code; };
    if (cfg->smooth > 0) {
        std::vector<double> sig(cfg->n_scales);
        for (int si = 0; si < cfg->n_scales; ++si) sig[si] = gabor_sigma(p->freqs[si], cfg->bandwidth);
        p->smooth = smooth_plan_new(sig.data(), cfg->n_scales, cfg->n_orient, p->D, H, W, cfg->smooth);
        if (!p->smooth) { set_error(GCIS_E_INVALID, "plan: smoothing set-up failed (radius too large or no memory)"); return fail(GCIS_E_INVALID); }
    }
#define PA(ptr, n)                                        \
    do {                                                  \
        int _rc = dev_alloc(&(ptr), (n), &p->bytes);      \
        if (_rc) return fail(_rc);                        \
    } while (0)
    PA(p->d_taps, p->bank.taps.size());
    PA(p->d_scales, p->bank.scales.size());
    p->n_lanes = 2;
    if (const char *e = getenv("GCIS_LANES")) p->n_lanes = atoi(e) >= 2 ? 2 : 1;
    if (cfg->max_batch <= p->group) p->n_lanes = 1;
    for (int l = 0; l < p->n_lanes; ++l) {
        // f32 planes, or bf16 planes (half the bytes) when the tensor cores take the row pass
        PA(p->d_planes[l], p->gtc ? ((size_t)p->group * 3 * H * p->Wp16 + 1) / 2 : (size_t)p->group * 3 * H * p->Wp);
        PA(p->d_feat[l], (size_t)p->group * p->D * p->Np);
        if (p->smooth) PA(p->d_tmp[l], (size_t)p->group * p->D * p->N);
        if (cfg->normalise) {
            PA(p->d_stats[l], (size_t)p->group * p->D * GB_STAT_SLOTS);
            PA(p->d_affine[l], (size_t)p->group * p->D * 2);
        }
        char *ws = nullptr;
        int rc2 = dev_alloc(&ws, kmeans_workspace_bytes(p->group, p->D, p->N, cfg->k), &p->bytes);
        if (rc2) return fail(rc2);
        p->d_km_ws[l] = ws;
    }
    PA(p->d_labels, MB * p->N);
    PA(p->d_bd_count, MB);
    PA(p->d_gt_counts, MB * G * GCIS_GT_SLOTS);
    PA(p->d_area, MB * k);
    PA(p->d_perim, MB * k);
    PA(p->d_hist, MB * G * k * cfg->n_lab_cap);
    PA(p->d_n_seg, MB);
    PA(p->d_n_lab, MB * G);
    PA(p->d_status, MB);
#undef PA
    if (cudaMemcpy(p->d_taps, p->bank.taps.data(), sizeof(float) * p->bank.taps.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p->d_scales, p->bank.scales.data(), sizeof(GaborScale) * p->bank.scales.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error(GCIS_E_CUDA, "plan: uploading the bank failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(GCIS_E_CUDA);
    }
    for (int l = 0; l < p->n_lanes; ++l)
        if (cudaStreamCreateWithFlags(&p->lane_stream[l], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->lane_done[l], cudaEventDisableTiming) != cudaSuccess) {
            set_error(GCIS_E_CUDA, "plan: lane stream/event creation failed");
            return fail(GCIS_E_CUDA);
        }
    if (cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) != cudaSuccess) {
        set_error(GCIS_E_CUDA, "plan: cudaEventCreate failed");
        return fail(GCIS_E_CUDA);
    }
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error(GCIS_E_CUDA, "plan: cudaStreamCreate failed");
        return fail(GCIS_E_CUDA);
    }
    *out = p;
    return GCIS_OK;
}

void gcis_plan_destroy(gcis_plan *p)
{
    if (!p) return;
    cudaFree(p->d_taps); cudaFree(p->d_scales);
    for (int l = 0; l < 2; ++l) {
        cudaFree(p->d_planes[l]); cudaFree(p->d_feat[l]); cudaFree(p->d_km_ws[l]);
        cudaFree(p->d_tmp[l]); cudaFree(p->d_stats[l]); cudaFree(p->d_affine[l]);
        if (p->lane_done[l]) cudaEventDestroy(p->lane_done[l]);
        if (p->lane_stream[l]) cudaStreamDestroy(p->lane_stream[l]);
    }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    cudaFree(p->d_labels); cudaFree(p->d_bd_count); cudaFree(p->d_gt_counts); cudaFree(p->d_area); cudaFree(p->d_perim);
    cudaFree(p->d_hist); cudaFree(p->d_n_seg); cudaFree(p->d_n_lab); cudaFree(p->d_status);
    cudaFree(p->d_img); cudaFree(p->d_gt); cudaFree(p->d_n_gt); cudaFree(p->d_init);
    for (cudaEvent_t e : p->events) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (p->ev_copied[i]) cudaEventDestroy(p->ev_copied[i]);
        if (p->ev_gt[i]) cudaEventDestroy(p->ev_gt[i]);
        if (p->ev_free[i]) cudaEventDestroy(p->ev_free[i]);
    }
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->stream) cudaStreamDestroy(p->stream);
    gabor_plan_delete(p->glp);
    gabor_tc_plan_delete(p->gtc);
    smooth_plan_delete(p->smooth);
    delete p;
}

int32_t gcis_plan_feature_dim(const gcis_plan *p) { return p ? p->D : set_error(GCIS_E_INVALID, "null plan"); }
int32_t gcis_plan_launch_group(const gcis_plan *p, int32_t B) { return p ? balanced_group(p, B) : set_error(GCIS_E_INVALID, "null plan"); }
int32_t gcis_plan_uses_tensor_cores(const gcis_plan *p) { return p ? (p->gtc != nullptr) : set_error(GCIS_E_INVALID, "null plan"); }
int64_t gcis_plan_workspace_bytes(const gcis_plan *p) { return p ? (int64_t)p->bytes : 0; }

int32_t gcis_plan_set_profiling(gcis_plan *p, int32_t on)
{
    if (!p) return set_error(GCIS_E_INVALID, "null plan");
    p->profiling = on != 0;
    return GCIS_OK;
}

int32_t gcis_plan_last_stage_ms(gcis_plan *p, float *ms4)
{
    if (!p || !ms4) return set_error(GCIS_E_INVALID, "null argument");
    memcpy(ms4, p->stage_ms, sizeof(p->stage_ms));
    return GCIS_OK;
}

int32_t gcis_gabor_features(gcis_plan *p, const uint8_t *d_img, int32_t B, float *d_feat, void *stream)
{
    TRY(plan_check_batch(p, B));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t img_stride = (size_t)p->N * 3, feat_stride = (size_t)p->D * p->N;
    for (int b0 = 0; b0 < B; b0 += p->group) {
        const int nb = std::min(p->group, B - b0);
        TRY(segment_group(p, 0, d_img + b0 * img_stride, nb, nullptr, nullptr, d_feat + b0 * feat_stride, st, -1));
    }
    return GCIS_OK;
}

int32_t gcis_feature_affine(gcis_plan *p, const float *d_feat, int32_t B, float *d_affine, void *stream)
{
    TRY(plan_check_batch(p, B));
    if (!d_feat || !d_affine) return set_error(GCIS_E_INVALID, "feature_affine: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t feat_stride = (size_t)p->D * p->N;
    const float stat_scale = (float)(1u << p->cfg.fix_shift);
    for (int b0 = 0; b0 < B; b0 += p->group) {
        const int nb = std::min(p->group, B - b0);
        long long *stats = p->d_stats[0];
        if (!stats) return set_error(GCIS_E_INVALID, "feature_affine: the plan was created without normalise");
        GCIS_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(long long) * (size_t)nb * p->D * GB_STAT_SLOTS, st));
        TRY(feature_moments_launch(d_feat + b0 * feat_stride, feat_stride, p->N, nb, p->D, p->N, stat_scale, stats, st));
        TRY(feature_affine_launch(stats, d_affine + (size_t)b0 * p->D * 2, nb, p->D, p->N, stat_scale, st));
    }
    return GCIS_OK;
}

int32_t gcis_kmeans(gcis_plan *p, const float *d_feat, int32_t B, const int32_t *d_init_idx, int32_t *d_labels,
                    float *d_centroids, void *stream)
{
    TRY(plan_check_batch(p, B));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const gcis_config &c = p->cfg;
    const size_t feat_stride = (size_t)p->D * p->N;
    for (int b0 = 0; b0 < B; b0 += p->group) {
        const int nb = std::min(p->group, B - b0);
        const float *affine = nullptr;
        if (c.normalise) {   // caller-supplied features: their moments are taken here, then the same folded map
            TRY(gcis_feature_affine(p, d_feat + b0 * feat_stride, nb, p->d_affine[0], stream));
            affine = p->d_affine[0];
        }
        TRY(kmeans_launch(d_feat + b0 * feat_stride, feat_stride, p->N, nb, p->D, p->N, c.k, c.iters, c.fix_shift, d_init_idx + (size_t)b0 * c.k,
                          d_labels + (size_t)b0 * p->N, d_centroids ? d_centroids + (size_t)b0 * c.k * p->D : nullptr,
                          p->d_km_ws[0], st, affine));
    }
    return GCIS_OK;
}

int32_t gcis_segment_device(gcis_plan *p, const uint8_t *d_img, int32_t B, const int32_t *d_init_idx,
                            int32_t *d_labels, void *stream)
{
    TRY(plan_check_batch(p, B));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t img_stride = (size_t)p->N * 3;
    const bool lanes = p->n_lanes == 2 && B > p->group;
    if (lanes) {   // fork: both lanes start after everything already queued on the caller's stream
        GCIS_CUDA_TRY(cudaEventRecord(p->ev_fork, st));
        for (int l = 0; l < 2; ++l) GCIS_CUDA_TRY(cudaStreamWaitEvent(p->lane_stream[l], p->ev_fork, 0));
    }
    int gi = 0;
    const int grp = balanced_group(p, B);
    for (int b0 = 0; b0 < B; b0 += grp, ++gi) {
        const int nb = std::min(grp, B - b0);
        const int lane = lanes ? (gi & 1) : 0;
        TRY(segment_group(p, lane, d_img + b0 * img_stride, nb, d_init_idx + (size_t)b0 * p->cfg.k,
                          d_labels + (size_t)b0 * p->N, nullptr, lanes ? p->lane_stream[lane] : st, gi));
    }
    if (lanes)     // join
        for (int l = 0; l < 2; ++l) {
            GCIS_CUDA_TRY(cudaEventRecord(p->lane_done[l], p->lane_stream[l]));
            GCIS_CUDA_TRY(cudaStreamWaitEvent(st, p->lane_done[l], 0));
        }
    p->n_groups_last = gi;
    return GCIS_OK;
}

int32_t gcis_label_metrics_device(const int32_t *d_lb, const uint16_t *d_gt, const int32_t *d_n_gt, int32_t B,
                                  int32_t H, int32_t W, int32_t G, int32_t n_seg_cap, int32_t n_lab_cap,
                                  int32_t dil_recall, int64_t *d_bd_count, int64_t *d_gt_counts, int32_t *d_area,
                                  int32_t *d_perim, int32_t *d_hist, int32_t *d_n_seg, int32_t *d_n_lab,
                                  int32_t *d_status, void *stream)
{
    return label_metrics_launch(d_lb, d_gt, d_n_gt, B, H, W, G, n_seg_cap, n_lab_cap, dil_recall, d_bd_count,
                                d_gt_counts, d_area, d_perim, d_hist, d_n_seg, d_n_lab, d_status,
                                static_cast<cudaStream_t>(stream));
}

int32_t gcis_label_metrics_host(const int32_t *h_lb, const uint16_t *h_gt, const int32_t *h_n_gt, int32_t B, int32_t H,
                                int32_t W, int32_t G, int32_t n_seg_cap, int32_t n_lab_cap, int32_t dil_recall,
                                int64_t *h_bd_count, int64_t *h_gt_counts, int32_t *h_area, int32_t *h_perim,
                                int32_t *h_hist, int32_t *h_n_seg, int32_t *h_n_lab, int32_t *h_status)
{
    if (B < 1 || H < 1 || W < 1 || G < 0 || n_seg_cap < 1 || n_lab_cap < 1)
        return set_error(GCIS_E_INVALID, "label_metrics_host: bad shape");
    const size_t N = (size_t)H * W, Gs = std::max(G, 1), Bs = B;
    const size_t n_hist = Bs * Gs * n_seg_cap * n_lab_cap;
    const size_t total = pad256(4 * Bs * N) + pad256(2 * Bs * Gs * N) + pad256(4 * Bs) + 2 * pad256(4 * Bs * n_seg_cap) +
                         pad256(4 * n_hist) + pad256(4 * Bs) + pad256(4 * Bs * Gs) + pad256(4 * Bs) + pad256(8 * Bs) +
                         pad256(8 * Bs * Gs * GCIS_GT_SLOTS) + 4096;
    TRY(g_arena.reserve(total));
    int32_t *d_lb = g_arena.take<int32_t>(Bs * N);
    uint16_t *d_gt = g_arena.take<uint16_t>(Bs * Gs * N);
    int32_t *d_n_gt = g_arena.take<int32_t>(Bs), *d_area = g_arena.take<int32_t>(Bs * n_seg_cap), *d_perim = g_arena.take<int32_t>(Bs * n_seg_cap);
    int32_t *d_hist = g_arena.take<int32_t>(n_hist), *d_n_seg = g_arena.take<int32_t>(Bs), *d_n_lab = g_arena.take<int32_t>(Bs * Gs);
    int32_t *d_status = g_arena.take<int32_t>(Bs);
    int64_t *d_bd = g_arena.take<int64_t>(Bs), *d_gc = g_arena.take<int64_t>(Bs * Gs * GCIS_GT_SLOTS);
    int rc = GCIS_OK;
    // pageable host memory: the copies are staged by the driver; everything is ordered on the legacy default stream
    auto cp = [&](void *dst, const void *src, size_t n, cudaMemcpyKind kind) {
        if (rc || n == 0) return;
        cudaError_t e = cudaMemcpyAsync(dst, src, n, kind, nullptr);
        if (e != cudaSuccess) rc = set_error(GCIS_E_CUDA, "label_metrics_host: cudaMemcpy -> %s", cudaGetErrorString(e));
    };
    cp(d_lb, h_lb, sizeof(int32_t) * Bs * N, cudaMemcpyHostToDevice);
    if (G > 0) cp(d_gt, h_gt, sizeof(uint16_t) * Bs * Gs * N, cudaMemcpyHostToDevice);
    if (h_n_gt) cp(d_n_gt, h_n_gt, sizeof(int32_t) * Bs, cudaMemcpyHostToDevice);
    if (!rc)
        rc = label_metrics_launch(d_lb, d_gt, h_n_gt ? d_n_gt : nullptr, B, H, W, G, n_seg_cap, n_lab_cap, dil_recall, d_bd,
                                  d_gc, d_area, d_perim, d_hist, d_n_seg, d_n_lab, d_status, nullptr);
    cp(h_bd_count, d_bd, sizeof(int64_t) * Bs, cudaMemcpyDeviceToHost);
    cp(h_gt_counts, d_gc, sizeof(int64_t) * Bs * Gs * GCIS_GT_SLOTS, cudaMemcpyDeviceToHost);
    cp(h_area, d_area, sizeof(int32_t) * Bs * n_seg_cap, cudaMemcpyDeviceToHost);
    cp(h_perim, d_perim, sizeof(int32_t) * Bs * n_seg_cap, cudaMemcpyDeviceToHost);
    if (h_hist) cp(h_hist, d_hist, sizeof(int32_t) * n_hist, cudaMemcpyDeviceToHost);
    cp(h_n_seg, d_n_seg, sizeof(int32_t) * Bs, cudaMemcpyDeviceToHost);
    cp(h_n_lab, d_n_lab, sizeof(int32_t) * Bs * Gs, cudaMemcpyDeviceToHost);
    cp(h_status, d_status, sizeof(int32_t) * Bs, cudaMemcpyDeviceToHost);
    if (!rc) {
        cudaError_t e = cudaStreamSynchronize(nullptr);
        if (e != cudaSuccess) rc = set_error(GCIS_E_CUDA, "label_metrics_host: kernel -> %s", cudaGetErrorString(e));
    }
    if (!rc)
        for (int b = 0; b < B; ++b)
            if (h_status[b]) return set_error(GCIS_E_LABEL, "label_metrics: image %d has status %d (negative label or label >= capacity)", b, h_status[b]);
    return rc;
}

int32_t gcis_find_boundaries_host(const int32_t *h_x, int32_t B, int32_t H, int32_t W, uint8_t *h_out)
{
    if (!h_x || !h_out || B < 1 || H < 1 || W < 1) return set_error(GCIS_E_INVALID, "find_boundaries: bad argument");
    const size_t n = (size_t)B * H * W;
    TRY(g_arena.reserve(pad256(n * 4) + pad256(n) + 1024));
    int32_t *d_x = g_arena.take<int32_t>(n);
    uint8_t *d_o = g_arena.take<uint8_t>(n);
    GCIS_CUDA_TRY(cudaMemcpyAsync(d_x, h_x, n * sizeof(int32_t), cudaMemcpyHostToDevice, nullptr));
    TRY(find_boundaries_launch(d_x, d_o, B, H, W, nullptr));
    GCIS_CUDA_TRY(cudaMemcpyAsync(h_out, d_o, n, cudaMemcpyDeviceToHost, nullptr));
    GCIS_CUDA_TRY(cudaStreamSynchronize(nullptr));
    return GCIS_OK;
}

int32_t gcis_pipeline_device(gcis_plan *p, const uint8_t *d_img, const uint16_t *d_gt, const int32_t *d_n_gt,
                             const int32_t *d_init_idx, int32_t B, void *stream)
{
    TRY(plan_check_batch(p, B));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const gcis_config &c = p->cfg;
    TRY(gcis_segment_device(p, d_img, B, d_init_idx, p->d_labels, stream));
    const size_t me = 4 * (size_t)p->n_groups_last;
    if (p->profiling) cudaEventRecord(plan_event(p, me), st);
    TRY(label_metrics_launch(p->d_labels, d_gt, d_n_gt, B, c.height, c.width, c.max_gt, c.k, c.n_lab_cap, c.dil_recall,
                             p->d_bd_count, p->d_gt_counts, p->d_area, p->d_perim, p->d_hist, p->d_n_seg, p->d_n_lab,
                             p->d_status, st));
    if (p->profiling) {
        cudaEventRecord(plan_event(p, me + 1), st);
        GCIS_CUDA_TRY(cudaEventSynchronize(p->events[me + 1]));
        float acc[4] = {0, 0, 0, 0}, ms = 0;
        for (int g = 0; g < p->n_groups_last; ++g) {
            cudaEventElapsedTime(&ms, p->events[4 * g], p->events[4 * g + 1]); acc[0] += ms;
            cudaEventElapsedTime(&ms, p->events[4 * g + 1], p->events[4 * g + 2]); acc[1] += ms;
            cudaEventElapsedTime(&ms, p->events[4 * g + 2], p->events[4 * g + 3]); acc[2] += ms;
        }
        cudaEventElapsedTime(&ms, p->events[me], p->events[me + 1]); acc[3] = ms;
        memcpy(p->stage_ms, acc, sizeof(acc));
    }
    return GCIS_OK;
}

int32_t gcis_pipeline_fetch(gcis_plan *p, int32_t B, int64_t *h_bd_count, int64_t *h_gt_counts, int32_t *h_area,
                            int32_t *h_perim, int32_t *h_n_lab, int32_t *h_status, int32_t *h_labels, void *stream)
{
    TRY(plan_check_batch(p, B));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const gcis_config &c = p->cfg;
    const size_t G = std::max(c.max_gt, 1);
    GCIS_CUDA_TRY(cudaMemcpyAsync(h_bd_count, p->d_bd_count, sizeof(int64_t) * B, cudaMemcpyDeviceToHost, st));
    GCIS_CUDA_TRY(cudaMemcpyAsync(h_gt_counts, p->d_gt_counts, sizeof(int64_t) * B * G * GCIS_GT_SLOTS, cudaMemcpyDeviceToHost, st));
    GCIS_CUDA_TRY(cudaMemcpyAsync(h_area, p->d_area, sizeof(int32_t) * (size_t)B * c.k, cudaMemcpyDeviceToHost, st));
    GCIS_CUDA_TRY(cudaMemcpyAsync(h_perim, p->d_perim, sizeof(int32_t) * (size_t)B * c.k, cudaMemcpyDeviceToHost, st));
    GCIS_CUDA_TRY(cudaMemcpyAsync(h_n_lab, p->d_n_lab, sizeof(int32_t) * B * G, cudaMemcpyDeviceToHost, st));
    GCIS_CUDA_TRY(cudaMemcpyAsync(h_status, p->d_status, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
    if (h_labels)
        GCIS_CUDA_TRY(cudaMemcpyAsync(h_labels, p->d_labels, sizeof(int32_t) * (size_t)B * p->N, cudaMemcpyDeviceToHost, st));
    GCIS_CUDA_TRY(cudaStreamSynchronize(st));
    return GCIS_OK;
}

int32_t gcis_pipeline_fetch_hist(gcis_plan *p, int32_t B, int32_t *h_hist, void *stream)
{
    TRY(plan_check_batch(p, B));
    if (!h_hist) return set_error(GCIS_E_INVALID, "null h_hist");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const gcis_config &c = p->cfg;
    const size_t n = (size_t)B * std::max(c.max_gt, 1) * c.k * c.n_lab_cap;
    GCIS_CUDA_TRY(cudaMemcpyAsync(h_hist, p->d_hist, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    GCIS_CUDA_TRY(cudaStreamSynchronize(st));
    return GCIS_OK;
}

int32_t gcis_pipeline_host(gcis_plan *p, const uint8_t *h_img, const uint16_t *h_gt, const int32_t *h_n_gt,
                           const int32_t *h_init_idx, int32_t B, int64_t *h_bd_count, int64_t *h_gt_counts,
                           int32_t *h_area, int32_t *h_perim, int32_t *h_n_lab, int32_t *h_status, int32_t *h_labels)
{
    if (!p) return set_error(GCIS_E_INVALID, "null plan");
    if (B < 1) return set_error(GCIS_E_INVALID, "B=%d", B);
    const gcis_config &c = p->cfg;
    const size_t G = std::max(c.max_gt, 1), N = p->N;
    // Sub-chunks of one group each travel through two device input buffers: while the compute
    // stream works on sub-chunk i, the copy stream uploads sub-chunk i+1 (pinned host memory
    // makes the copies truly asynchronous).
    const size_t hc = p->group;
    if (!p->host_ready) {
        // Everything is created into locals and committed to the plan only when all of it exists, so a
        // failure half-way leaves the plan exactly as it was (and the next call tries again).
        uint8_t *di = nullptr; uint16_t *dg = nullptr; int32_t *dn = nullptr, *dx = nullptr;
        cudaStream_t cs = nullptr;
        cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        size_t bytes = 0;
        int rc = dev_alloc(&di, 2 * hc * N * 3, &bytes);
        if (!rc) rc = dev_alloc(&dg, 2 * hc * G * N, &bytes);
        if (!rc) rc = dev_alloc(&dn, 2 * hc, &bytes);
        if (!rc) rc = dev_alloc(&dx, 2 * hc * c.k, &bytes);
        if (!rc && cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess)
            rc = set_error(GCIS_E_CUDA, "pipeline_host: cudaStreamCreate failed");
        for (int i = 0; i < 6 && !rc; ++i)
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess)
                rc = set_error(GCIS_E_CUDA, "pipeline_host: cudaEventCreate failed");
        if (rc) {
            cudaFree(di); cudaFree(dg); cudaFree(dn); cudaFree(dx);
            if (cs) cudaStreamDestroy(cs);
            for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
            return rc;
        }
        p->d_img = di; p->d_gt = dg; p->d_n_gt = dn; p->d_init = dx; p->copy_stream = cs; p->bytes += bytes;
        for (int i = 0; i < 2; ++i) { p->ev_copied[i] = ev[3 * i]; p->ev_gt[i] = ev[3 * i + 1]; p->ev_free[i] = ev[3 * i + 2]; }
        p->host_ready = true;
    }
    cudaStream_t st = p->stream, cs = p->copy_stream;
    const bool lanes = p->n_lanes == 2;
    for (int m0 = 0; m0 < B; m0 += c.max_batch) {          // the result buffers hold max_batch images
        const int mb = std::min<int>(c.max_batch, B - m0);
        int i = 0;
        const int sub = balanced_group(p, mb);               // equal sub-chunks, each at most hc images
        for (int s0 = 0; s0 < mb; s0 += sub, ++i) {
            const int nb = std::min<int>(sub, mb - s0), buf = i & 1, b0 = m0 + s0;
            uint8_t *di = p->d_img + (size_t)buf * hc * N * 3;
            uint16_t *dg = p->d_gt + (size_t)buf * hc * G * N;
            int32_t *dn = p->d_n_gt + (size_t)buf * hc, *dx = p->d_init + (size_t)buf * hc * c.k;
            if (i >= 2) GCIS_CUDA_TRY(cudaStreamWaitEvent(cs, p->ev_free[buf], 0));
            // images first: the segmenter starts as soon as they have landed; the ground truths (3.3x the
            // bytes) follow and are only awaited by the metrics kernels
            GCIS_CUDA_TRY(cudaMemcpyAsync(di, h_img + (size_t)b0 * N * 3, (size_t)nb * N * 3, cudaMemcpyHostToDevice, cs));
            GCIS_CUDA_TRY(cudaMemcpyAsync(dx, h_init_idx + (size_t)b0 * c.k, sizeof(int32_t) * nb * c.k, cudaMemcpyHostToDevice, cs));
            GCIS_CUDA_TRY(cudaEventRecord(p->ev_copied[buf], cs));
            if (c.max_gt > 0)
                GCIS_CUDA_TRY(cudaMemcpyAsync(dg, h_gt + (size_t)b0 * G * N, sizeof(uint16_t) * nb * G * N, cudaMemcpyHostToDevice, cs));
            if (h_n_gt) GCIS_CUDA_TRY(cudaMemcpyAsync(dn, h_n_gt + b0, sizeof(int32_t) * nb, cudaMemcpyHostToDevice, cs));
            GCIS_CUDA_TRY(cudaEventRecord(p->ev_gt[buf], cs));
            cudaStream_t ls = lanes ? p->lane_stream[buf] : st;
            GCIS_CUDA_TRY(cudaStreamWaitEvent(ls, p->ev_copied[buf], 0));
            TRY(pipeline_range(p, lanes ? buf : 0, di, dg, h_n_gt ? dn : nullptr, dx, s0, nb, ls, p->ev_gt[buf]));
            GCIS_CUDA_TRY(cudaEventRecord(p->ev_free[buf], ls));
        }
        if (lanes)
            for (int l = 0; l < 2; ++l) {
                GCIS_CUDA_TRY(cudaEventRecord(p->lane_done[l], p->lane_stream[l]));
                GCIS_CUDA_TRY(cudaStreamWaitEvent(st, p->lane_done[l], 0));
            }
        TRY(gcis_pipeline_fetch(p, mb, h_bd_count + m0, h_gt_counts + (size_t)m0 * G * GCIS_GT_SLOTS, h_area + (size_t)m0 * c.k,
                                h_perim + (size_t)m0 * c.k, h_n_lab + (size_t)m0 * G, h_status + m0,
                                h_labels ? h_labels + (size_t)m0 * N : nullptr, st));
        // both input buffers are idle again before the next block reuses them from index 0
        GCIS_CUDA_TRY(cudaStreamSynchronize(cs));
    }
    return GCIS_OK;
}

}  // extern "C"
