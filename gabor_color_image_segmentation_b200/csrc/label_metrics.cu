// label_metrics.cu — BSD_metrics integer counts as label-comparison kernels.
//
// Reference: BSD_metrics/metrics.py (paths relative to the reference checkout)
//   :47-49   GT boundary maps          find_boundaries(gt_g)
//   :69-72   recall numerators         |dil_size(bd(lb)) & bd(gt_g)|, |bd(gt_g)|
//   :88-94   precision numerators      |bd(lb) & dil_5(bd(gt_g))|, |bd(lb)|
//   :115-126 contingency table         hist[lb][gt] (+ area)
//   :129-142 undersegmentation sums    U_g, V_g
//   :166-180 perimeter                 border-or-boundary pixels per label
//
// Layout in HBM: lb [B][H][W] int32, gt [B][G][H][W] uint16, both read exactly once
// (2.16 MB per 321x481 image with G=5).  One CTA owns a 36x64 pixel tile of one image:
// it stages the label tile plus halo in shared memory, derives the boundary map and
// its separable square dilation there, and reuses the same staging buffers for each
// ground truth.  Contingency counts go to a shared-memory histogram (when it fits)
// with warp-aggregated atomics, flushed once per (tile, ground truth).
#include "common.cuh"

namespace gcis {

namespace {

constexpr int LM_TW = 64;
#ifndef LM_TH_N
#define LM_TH_N 36   // 321 rows = 9 tiles of 36 (32: 11 tiles, the last with one row): 1.76 -> 1.68 ms per 200 images; 40: 1.83, 48: 2.79
#endif
constexpr int LM_TH = LM_TH_N;
constexpr int LM_THREADS = 256;
constexpr int LM_PIX_PER_THREAD = LM_TW * LM_TH / LM_THREADS;
constexpr int LM_SMEM_HIST_MAX = 4096;  // entries (16 KB)
constexpr int LM_SENTINEL = INT32_MIN;  // "outside the image"

struct LmParams {
    const int32_t *lb;
    const uint16_t *gt;
    const int32_t *n_gt;
    int B, H, W, G, n_seg_cap, n_lab_cap;
    int rlo, rhi;  // recall dilation window offsets (metrics.py:69)
    int halo;      // 1 (boundary) + widest dilation reach
    int64_t *bd_count;
    int64_t *gt_counts;
    int32_t *area, *perim, *hist, *n_seg, *n_lab, *status;
    int use_smem_hist;
};

__device__ __forceinline__ int warp_sum(int v) { return (int)__reduce_add_sync(0xffffffffu, (unsigned)v); }

// One atomic per distinct key per warp.  key < 0 = lane has nothing to add.
__device__ __forceinline__ void warp_agg_inc(int32_t *base, int key)
{
    unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(base + key, __popc(peers));
}

// The window loops below walk (row = warp + 8 i, column = lane + 32 j): no integer division by the
// run-time window width, which dominated the instruction count of the first version of this kernel.
constexpr int LM_WARPS = LM_THREADS / 32;

// Stage a (LM_TH+2*halo) x (LM_TW+2*halo) label window; outside the image -> sentinel.
template <typename T>
__device__ __forceinline__ void stage_labels(const T *__restrict__ src, int H, int W, int r0, int c0, int halo,
                                             int32_t *sl, int &vmax, int &neg)
{
    const int SW = LM_TW + 2 * halo, SH = LM_TH + 2 * halo;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int rr = warp; rr < SH; rr += LM_WARPS) {
        const int r = r0 - halo + rr;
        const bool row_in = r >= 0 && r < H;
        const bool row_own = rr >= halo && rr < halo + LM_TH;
        const T *srow = src + (size_t)(row_in ? r : 0) * W;
        int32_t *drow = sl + rr * SW;
        for (int cc = lane; cc < SW; cc += 32) {
            const int c = c0 - halo + cc;
            int v = LM_SENTINEL;
            if (row_in && c >= 0 && c < W) {
                v = (int)srow[c];
                // only own pixels decide max / sign so every pixel is judged exactly once
                if (row_own && cc >= halo && cc < halo + LM_TW) {
                    vmax = max(vmax, v);
                    neg |= (v < 0);
                }
            }
            drow[cc] = v;
        }
    }
}

// Boundary map (metrics.py:49 / SURVEY A.1) on the window shrunk by one pixel.
__device__ __forceinline__ void boundary_map(const int32_t *sl, int halo, uint8_t *sb)
{
    const int SW = LM_TW + 2 * halo;
    const int e = halo - 1, EW = LM_TW + 2 * e, EH = LM_TH + 2 * e;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int rr = warp; rr < EH; rr += LM_WARPS) {
        const int32_t *prow = sl + (rr + 1) * SW + 1;
        uint8_t *drow = sb + rr * EW;
        for (int cc = lane; cc < EW; cc += 32) {
            const int32_t *p = prow + cc;
            const int v = p[0];
            int b = 0;
            if (v != LM_SENTINEL) {
                int n;
                n = p[-SW]; b |= (n != LM_SENTINEL) & (n != v);
                n = p[SW];  b |= (n != LM_SENTINEL) & (n != v);
                n = p[-1];  b |= (n != LM_SENTINEL) & (n != v);
                n = p[1];   b |= (n != LM_SENTINEL) & (n != v);
            }
            drow[cc] = (uint8_t)b;
        }
    }
}

// Separable square dilation (metrics.py:69,93 / SURVEY A.2) evaluated on the tile's own
// pixels.  sb is the boundary map on the e-expanded window; sh is scratch.
__device__ __forceinline__ void dilate_own(const uint8_t *sb, int halo, int lo, int hi, uint8_t *sh, uint8_t *out)
{
    static_assert(LM_TW == 64, "two columns per lane");
    const int e = halo - 1, EW = LM_TW + 2 * e, EH = LM_TH + 2 * e;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int rr = warp; rr < EH; rr += LM_WARPS) {
        const uint8_t *p = sb + rr * EW + e + lane;
        int v0 = 0, v1 = 0;
        for (int d = lo; d <= hi; ++d) { v0 |= p[d]; v1 |= p[d + 32]; }
        sh[rr * LM_TW + lane] = (uint8_t)v0;
        sh[rr * LM_TW + lane + 32] = (uint8_t)v1;
    }
    __syncthreads();
    for (int r = warp; r < LM_TH; r += LM_WARPS) {
        const uint8_t *p = sh + (r + e) * LM_TW + lane;
        int v0 = 0, v1 = 0;
        for (int d = lo; d <= hi; ++d) { v0 |= p[d * LM_TW]; v1 |= p[d * LM_TW + 32]; }
        out[r * LM_TW + lane] = (uint8_t)v0;
        out[r * LM_TW + lane + 32] = (uint8_t)v1;
    }
}

__global__ void __launch_bounds__(LM_THREADS) lm_tile_kernel(LmParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int halo = P.halo, e = halo - 1;
    const int SW = LM_TW + 2 * halo, SH = LM_TH + 2 * halo;
    const int EW = LM_TW + 2 * e, EH = LM_TH + 2 * e;
    int32_t *sl = reinterpret_cast<int32_t *>(smem_raw);
    int32_t *shist = sl + SW * SH;
    uint8_t *sb = reinterpret_cast<uint8_t *>(shist + (P.use_smem_hist ? P.n_seg_cap * P.n_lab_cap : 0));
    uint8_t *sh = sb + EW * EH;
    uint8_t *bd_own = sh + EH * LM_TW;
    uint8_t *dil_own = bd_own + LM_TH * LM_TW;
    uint8_t *t_own = dil_own + LM_TH * LM_TW;   // bd(gt) on own pixels
    uint8_t *tdil_own = t_own + LM_TH * LM_TW;  // dil_5(bd(gt)) on own pixels
    __shared__ int s_cnt[4];                    // bd, den_r, tp_r, tp_p
    __shared__ int s_max, s_flag;

    const int b = blockIdx.z;
    const int r0 = blockIdx.y * LM_TH, c0 = blockIdx.x * LM_TW;
    const int H = P.H, W = P.W;
    const size_t N = (size_t)H * W;
    const int32_t *lb = P.lb + (size_t)b * N;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_max = -1; s_flag = 0; }
    __syncthreads();

    // ---- segmentation labels: boundary, dilation, area, perimeter ----
    int vmax = -1, neg = 0;
    stage_labels(lb, H, W, r0, c0, halo, sl, vmax, neg);
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    neg = __any_sync(0xffffffffu, neg);
    if (lane == 0) {
        atomicMax(&s_max, vmax);
        if (neg) atomicOr(&s_flag, GCIS_ST_NEG_LABEL);
    }
    __syncthreads();
    boundary_map(sl, halo, sb);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_max >= 0) atomicMax(P.n_seg + b, s_max + 1);
        if (s_max >= P.n_seg_cap) s_flag |= GCIS_ST_SEG_OVER;
    }
    dilate_own(sb, halo, P.rlo, P.rhi, sh, dil_own);
    int own_lb[LM_PIX_PER_THREAD];
    {
        int nbd = 0;
#pragma unroll
        for (int k = 0; k < LM_PIX_PER_THREAD; ++k) {
            int i = threadIdx.x + k * LM_THREADS;
            int r = i / LM_TW, c = i - r * LM_TW;
            int gr = r0 + r, gc = c0 + c;
            bool in = gr < H && gc < W;
            int v = sl[(r + halo) * SW + c + halo];
            bool ok = in && v >= 0 && v < P.n_seg_cap;
            own_lb[k] = ok ? v : -1;
            int bd = sb[(r + e) * EW + c + e];
            bd_own[i] = (uint8_t)bd;
            nbd += in ? bd : 0;
            warp_agg_inc(P.area + (size_t)b * P.n_seg_cap, ok ? v : -1);
            bool per = ok && (gr == 0 || gr == H - 1 || gc == 0 || gc == W - 1 || bd);  // metrics.py:172-180
            warp_agg_inc(P.perim + (size_t)b * P.n_seg_cap, per ? v : -1);
        }
        nbd = warp_sum(nbd);
        if (lane == 0 && nbd) atomicAdd(&s_cnt[0], nbd);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(reinterpret_cast<unsigned long long *>(P.bd_count + b), (unsigned long long)s_cnt[0]);
        if (s_flag) atomicOr(P.status + b, s_flag);
    }
    const bool lb_ok = !(s_flag & (GCIS_ST_NEG_LABEL | GCIS_ST_SEG_OVER));

    // ---- each ground truth: boundary, dil_5, overlaps, contingency ----
    const int ng = P.n_gt ? min(P.n_gt[b], P.G) : P.G;
    for (int g = 0; g < ng; ++g) {
        __syncthreads();
        if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
        if (threadIdx.x == 0) { s_max = -1; s_flag = 0; }
        if (P.use_smem_hist)
            for (int i = threadIdx.x; i < P.n_seg_cap * P.n_lab_cap; i += LM_THREADS) shist[i] = 0;
        __syncthreads();
        const uint16_t *gt = P.gt + ((size_t)b * P.G + g) * N;
        int tmax = -1, tneg = 0;
        stage_labels(gt, H, W, r0, c0, halo, sl, tmax, tneg);
        tmax = __reduce_max_sync(0xffffffffu, tmax);
        if (lane == 0) atomicMax(&s_max, tmax);
        __syncthreads();
        boundary_map(sl, halo, sb);
        __syncthreads();
        for (int i = threadIdx.x; i < LM_TH * LM_TW; i += LM_THREADS) {
            int r = i / LM_TW, c = i - r * LM_TW;
            t_own[i] = sb[(r + e) * EW + c + e];
        }
        dilate_own(sb, halo, -2, 2, sh, tdil_own);  // metrics.py:93: 5 hard-coded
        __syncthreads();
        const bool over = s_max >= P.n_lab_cap;
        int32_t *ghist = P.hist + ((size_t)b * P.G + g) * P.n_seg_cap * P.n_lab_cap;
        int den = 0, tpr = 0, tpp = 0;
#pragma unroll
        for (int k = 0; k < LM_PIX_PER_THREAD; ++k) {
            int i = threadIdx.x + k * LM_THREADS;
            int r = i / LM_TW, c = i - r * LM_TW;
            bool in = (r0 + r) < H && (c0 + c) < W;
            int t = in ? t_own[i] : 0;
            den += t;
            tpr += t & dil_own[i];
            tpp += in ? (bd_own[i] & tdil_own[i]) : 0;
            int gv = sl[(r + halo) * SW + c + halo];
            int key = (in && lb_ok && !over && own_lb[k] >= 0) ? own_lb[k] * P.n_lab_cap + gv : -1;
            warp_agg_inc(P.use_smem_hist ? shist : ghist, key);
        }
        den = warp_sum(den); tpr = warp_sum(tpr); tpp = warp_sum(tpp);
        if (lane == 0) {
            if (den) atomicAdd(&s_cnt[1], den);
            if (tpr) atomicAdd(&s_cnt[2], tpr);
            if (tpp) atomicAdd(&s_cnt[3], tpp);
        }
        __syncthreads();
        if (P.use_smem_hist)
            for (int i = threadIdx.x; i < P.n_seg_cap * P.n_lab_cap; i += LM_THREADS) {
                int v = shist[i];
                if (v) atomicAdd(ghist + i, v);
            }
        if (threadIdx.x == 0) {
            unsigned long long *gc = reinterpret_cast<unsigned long long *>(P.gt_counts + ((size_t)b * P.G + g) * GCIS_GT_SLOTS);
            if (s_cnt[1]) atomicAdd(gc + 0, (unsigned long long)s_cnt[1]);
            if (s_cnt[2]) atomicAdd(gc + 1, (unsigned long long)s_cnt[2]);
            if (s_cnt[3]) atomicAdd(gc + 2, (unsigned long long)s_cnt[3]);
            if (s_max >= 0) atomicMax(P.n_lab + (size_t)b * P.G + g, s_max + 1);
            if (over) atomicOr(P.status + b, GCIS_ST_LAB_OVER);
        }
    }
}

// U_g, V_g (metrics.py:129-142) and the PRI sums from one contingency table per CTA.
__global__ void __launch_bounds__(128) lm_finalize_kernel(LmParams P)
{
    const int b = blockIdx.x / P.G, g = blockIdx.x % P.G;
    const int ng = P.n_gt ? min(P.n_gt[b], P.G) : P.G;
    if (g >= ng) return;
    const int32_t *h = P.hist + ((size_t)b * P.G + g) * P.n_seg_cap * P.n_lab_cap;
    const int nS = P.n_seg_cap, nL = P.n_lab_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    long long u = 0, v = 0, sq = 0, csq = 0;
    for (int i = warp; i < nS; i += nwarp) {
        const int32_t *row = h + (size_t)i * nL;
        long long rs = 0;
        int rm = 0;
        for (int j = lane; j < nL; j += 32) {
            int x = row[j];
            rs += x;
            rm = max(rm, x);
            sq += (long long)x * x;
        }
        for (int o = 16; o; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
        rm = __reduce_max_sync(0xffffffffu, rm);
        if (lane == 0) u += rs - rm;
        for (int j = lane; j < nL; j += 32) {
            long long x = row[j];
            v += min(x, rs - x);
        }
    }
    for (int j = threadIdx.x; j < nL; j += blockDim.x) {
        long long cs = 0;
        for (int i = 0; i < nS; ++i) cs += h[(size_t)i * nL + j];
        csq += cs * cs;
    }
    __shared__ long long red[4][4];
    for (int o = 16; o; o >>= 1) {
        u += __shfl_xor_sync(0xffffffffu, u, o);
        v += __shfl_xor_sync(0xffffffffu, v, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        csq += __shfl_xor_sync(0xffffffffu, csq, o);
    }
    if (lane == 0) { red[warp][0] = u; red[warp][1] = v; red[warp][2] = sq; red[warp][3] = csq; }
    __syncthreads();
    if (threadIdx.x < 4) {
        long long s = 0;
        for (int w = 0; w < nwarp; ++w) s += red[w][threadIdx.x];
        P.gt_counts[((size_t)b * P.G + g) * GCIS_GT_SLOTS + 3 + threadIdx.x] = s;
    }
}

// find_boundaries(x) for whole maps (metrics.py:47-49: the img_truth attribute).
__global__ void find_boundaries_kernel(const int32_t *__restrict__ x, uint8_t *__restrict__ out, int B, int H, int W)
{
    const long long total = (long long)B * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % W);
        const int r = (int)((i / W) % H);
        const int v = x[i];
        int b = 0;
        if (r > 0) b |= x[i - W] != v;
        if (r + 1 < H) b |= x[i + W] != v;
        if (c > 0) b |= x[i - 1] != v;
        if (c + 1 < W) b |= x[i + 1] != v;
        out[i] = (uint8_t)b;
    }
}

}  // namespace

int find_boundaries_launch(const int32_t *d_x, uint8_t *d_out, int B, int H, int W, cudaStream_t st)
{
    const long long total = (long long)B * H * W;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    find_boundaries_kernel<<<blocks, 256, 0, st>>>(d_x, d_out, B, H, W);
    GCIS_LAUNCH_CHECK();
    return GCIS_OK;
}

int label_metrics_launch(const int32_t *d_lb, const uint16_t *d_gt, const int32_t *d_n_gt, int B, int H, int W,
                         int G, int n_seg_cap, int n_lab_cap, int dil_recall, int64_t *d_bd_count,
                         int64_t *d_gt_counts, int32_t *d_area, int32_t *d_perim, int32_t *d_hist,
                         int32_t *d_n_seg, int32_t *d_n_lab, int32_t *d_status, cudaStream_t st)
{
    if (B <= 0 || H <= 0 || W <= 0 || G < 0 || n_seg_cap <= 0 || n_lab_cap <= 0)
        return set_error(GCIS_E_INVALID, "label_metrics: bad shape B=%d H=%d W=%d G=%d caps=%d,%d", B, H, W, G,
                         n_seg_cap, n_lab_cap);
    if (dil_recall < 1 || dil_recall > 31)
        return set_error(GCIS_E_INVALID, "label_metrics: dil_recall=%d outside 1..31", dil_recall);
    if (B > 65535) return set_error(GCIS_E_INVALID, "label_metrics: B=%d > 65535 per call", B);
    LmParams P;
    P.lb = d_lb; P.gt = d_gt; P.n_gt = d_n_gt;
    P.B = B; P.H = H; P.W = W; P.G = G > 0 ? G : 1; P.n_seg_cap = n_seg_cap; P.n_lab_cap = n_lab_cap;
    if (dil_recall & 1) { P.rlo = -(dil_recall - 1) / 2; P.rhi = (dil_recall - 1) / 2; }
    else { P.rlo = -(dil_recall / 2 - 1); P.rhi = dil_recall / 2; }
    int reach = 2;
    if (-P.rlo > reach) reach = -P.rlo;
    if (P.rhi > reach) reach = P.rhi;
    P.halo = 1 + reach;
    P.bd_count = d_bd_count; P.gt_counts = d_gt_counts; P.area = d_area; P.perim = d_perim; P.hist = d_hist;
    P.n_seg = d_n_seg; P.n_lab = d_n_lab; P.status = d_status;
    P.use_smem_hist = ((int64_t)n_seg_cap * n_lab_cap <= LM_SMEM_HIST_MAX) ? 1 : 0;
    if (G == 0) P.n_gt = nullptr;

    const size_t Gs = (size_t)G;
    GCIS_CUDA_TRY(cudaMemsetAsync(d_bd_count, 0, sizeof(int64_t) * B, st));
    if (G > 0) GCIS_CUDA_TRY(cudaMemsetAsync(d_gt_counts, 0, sizeof(int64_t) * B * Gs * GCIS_GT_SLOTS, st));
    GCIS_CUDA_TRY(cudaMemsetAsync(d_area, 0, sizeof(int32_t) * (size_t)B * n_seg_cap, st));
    GCIS_CUDA_TRY(cudaMemsetAsync(d_perim, 0, sizeof(int32_t) * (size_t)B * n_seg_cap, st));
    if (G > 0) GCIS_CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof(int32_t) * (size_t)B * Gs * n_seg_cap * n_lab_cap, st));
    GCIS_CUDA_TRY(cudaMemsetAsync(d_n_seg, 0, sizeof(int32_t) * B, st));
    if (G > 0) GCIS_CUDA_TRY(cudaMemsetAsync(d_n_lab, 0, sizeof(int32_t) * B * Gs, st));
    GCIS_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t) * B, st));

    const int halo = P.halo, e = halo - 1;
    size_t smem = sizeof(int32_t) * (size_t)(LM_TW + 2 * halo) * (LM_TH + 2 * halo);
    if (P.use_smem_hist) smem += sizeof(int32_t) * (size_t)n_seg_cap * n_lab_cap;
    smem += (size_t)(LM_TW + 2 * e) * (LM_TH + 2 * e) + (size_t)(LM_TH + 2 * e) * LM_TW + 4 * (size_t)LM_TH * LM_TW;
    smem = (smem + 15) & ~(size_t)15;
    if (smem > 48 * 1024)
        GCIS_CUDA_TRY(cudaFuncSetAttribute(lm_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(W, LM_TW), ceil_div(H, LM_TH), B);
    if (G == 0) P.G = 0;
    lm_tile_kernel<<<grid, LM_THREADS, smem, st>>>(P);
    GCIS_LAUNCH_CHECK();
    if (G > 0) {
        lm_finalize_kernel<<<B * G, 128, 0, st>>>(P);
        GCIS_LAUNCH_CHECK();
    }
    return GCIS_OK;
}

}  // namespace gcis
