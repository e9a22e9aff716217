// kmeans.cu — Lloyd iterations as one fused assign + centroid-update kernel per iteration.
//
// Reference: none (segmenter slot, BSD_metrics/script.py:30; spec in DESIGN.md §3.4).
// Arithmetic contract, restated bit-for-bit by oracle/gcis_oracle.c:orc_kmeans:
//   m_jd = -2 c_jd;  cn_j = (float) sum_d (double)c_jd^2 (d ascending)
//   score_j(x) = fmaf(x_{D-1}, m_{j,D-1}, ... fmaf(x_0, m_j0, cn_j))        fp32 FMA chain
//   label = lowest j with the minimum score
//   centroid sums are exact int64 sums of q = rint(x * 2^fix_shift) -> the result does not
//   depend on the grid, the tile order or the number of GPUs
//   c_jd <- (float)((double)sum_jd / ((double)count_j * 2^fix_shift)); empty clusters keep c_jd
//
// Layout: feat [B][D][N] planar fp32 (every warp load is one 128-byte line of one feature
// plane); centroids [B][k][D]; sums [B][k][D] int64; counts [B][k] int32.
// One CTA owns KM_TILES consecutive 256-pixel tiles of one image.  Phase A (thread = pixel)
// streams the D planes once from HBM and keeps only the scores; phase B re-reads the CTA's own
// pixels from L2 in 32-feature slabs, transposes them through shared memory (lane = feature)
// and accumulates into warp-private int64 bins, so no atomics are contended.  The last CTA
// of an image to finish (ticket counter) turns sums into the next centroids.
#include "common.cuh"

namespace gcis {

namespace {

constexpr int KM_THREADS = 256;
constexpr int KM_WARPS = KM_THREADS / 32;
constexpr int KM_TILES = 4;                       // tiles per CTA
constexpr int KM_PX = KM_THREADS * KM_TILES;      // pixels per CTA
constexpr int KM_QSTR = KM_THREADS + 1;           // odd stride of the transposed slab

struct KmParams {
    const float *feat;
    const float *cent_in;
    float *cent_out;
    long long *sums;
    int *counts;
    int *done;
    int32_t *labels;   // written when non-null
    int D, N, k, chunks;
    float fix_scale;
};

__global__ void km_init_kernel(const float *__restrict__ feat, const int32_t *__restrict__ init_idx, float *cent,
                               long long *sums, int *counts, int *done, int D, int N, int k)
{
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < k * D; i += blockDim.x) {
        const int j = i / D, d = i - j * D;
        int p = init_idx[b * k + j];
        p = min(max(p, 0), N - 1);
        cent[(size_t)b * k * D + i] = feat[((size_t)b * D + d) * N + p];
        sums[(size_t)b * k * D + i] = 0;
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) counts[b * k + i] = 0;
    if (threadIdx.x == 0) done[b] = 0;
}

template <int K>
__global__ void __launch_bounds__(KM_THREADS) km_pass_kernel(const __grid_constant__ KmParams P)
{
    extern __shared__ __align__(16) unsigned char km_smem[];
    // [D][K] m (transposed so the K values of one feature are one 128-bit broadcast load)
    float *s_m = reinterpret_cast<float *>(km_smem);
    long long *s_acc = reinterpret_cast<long long *>(s_m + (size_t)((P.D * K + 3) & ~3));  // [warps][K][32]
    int *s_q = reinterpret_cast<int *>(s_acc + KM_WARPS * K * 32);                          // [32][KM_QSTR]
    __shared__ float s_cn[K];
    __shared__ int s_cnt[K];
    __shared__ unsigned char s_lab[KM_PX];
    __shared__ int s_last;

    const int b = blockIdx.y;
    const int D = P.D, N = P.N, k = P.k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *feat = P.feat + (size_t)b * D * N;
    const float *cin = P.cent_in + (size_t)b * k * D;
    const int p_base = blockIdx.x * KM_PX;

    for (int i = threadIdx.x; i < D * K; i += KM_THREADS) {
        const int d = i / K, j = i - d * K;
        s_m[i] = j < k ? -2.0f * cin[j * D + d] : 0.f;
    }
    if (threadIdx.x < K) {
        const int j = threadIdx.x;
        float cn = __int_as_float(0x7f800000);  // +inf: padded clusters never win
        if (j < k) {
            double acc = 0.0;
            for (int d = 0; d < D; ++d) {
                const double c = (double)cin[j * D + d];
                acc = __dadd_rn(acc, __dmul_rn(c, c));
            }
            cn = (float)acc;
        }
        s_cn[j] = cn;
        s_cnt[j] = 0;
    }
    __syncthreads();

    // ---- phase A: scores and labels, one pixel per thread ----
    int my_cnt = 0;  // lane j of every warp counts cluster j (and j + 32k for K > 32: not used, K <= 32)
    for (int t = 0; t < KM_TILES; ++t) {
        const int p = p_base + t * KM_THREADS + threadIdx.x;
        const bool valid = p < N;
        const int pc = valid ? p : N - 1;
        float s[K];
#pragma unroll
        for (int j = 0; j < K; ++j) s[j] = s_cn[j];
        const float *xp = feat + pc;
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
            const float x = __ldg(xp + (size_t)d * N);
            const float4 *mrow = reinterpret_cast<const float4 *>(s_m + d * K);
#pragma unroll
            for (int j4 = 0; j4 < K / 4; ++j4) {
                const float4 m = mrow[j4];
                s[4 * j4 + 0] = fmaf(x, m.x, s[4 * j4 + 0]);
                s[4 * j4 + 1] = fmaf(x, m.y, s[4 * j4 + 1]);
                s[4 * j4 + 2] = fmaf(x, m.z, s[4 * j4 + 2]);
                s[4 * j4 + 3] = fmaf(x, m.w, s[4 * j4 + 3]);
            }
        }
        int best = 0;
        float bs = s[0];
#pragma unroll
        for (int j = 1; j < K; ++j)
            if (s[j] < bs) { bs = s[j]; best = j; }
        s_lab[t * KM_THREADS + threadIdx.x] = valid ? (unsigned char)best : (unsigned char)255;
        if (valid && P.labels) P.labels[(size_t)b * N + p] = best;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const unsigned m = __ballot_sync(0xffffffffu, valid && best == j);
            if (lane == j) my_cnt += __popc(m);
        }
    }
    if (lane < K && my_cnt) atomicAdd(&s_cnt[lane], my_cnt);

    // ---- phase B: exact fixed-point centroid sums, 32 features at a time ----
    const int n_tiles = min(KM_TILES, (N - p_base + KM_THREADS - 1) / KM_THREADS);
    long long *my_acc = s_acc + (size_t)warp * K * 32 + lane;
    for (int d0 = 0; d0 < D; d0 += 32) {
        const int nd = min(32, D - d0);
        for (int j = 0; j < K; ++j) my_acc[j * 32] = 0;
        for (int t = 0; t < n_tiles; ++t) {
            __syncthreads();  // slab free
            const int p = p_base + t * KM_THREADS + threadIdx.x;
            const int pc = p < N ? p : N - 1;
            const float *xp = feat + (size_t)d0 * N + pc;
#pragma unroll 8
            for (int dd = 0; dd < nd; ++dd)
                s_q[dd * KM_QSTR + threadIdx.x] = __float2int_rn(__ldg(xp + (size_t)dd * N) * P.fix_scale);
            __syncthreads();
            if (lane < nd) {
                const unsigned char *lab = s_lab + t * KM_THREADS + warp * 32;
                const int *q = s_q + lane * KM_QSTR + warp * 32;
#pragma unroll 4
                for (int pp = 0; pp < 32; ++pp) {
                    const int l = lab[pp];
                    if (l != 255) my_acc[l * 32] += (long long)q[pp];
                }
            }
        }
        __syncthreads();
        // reduce the warp-private bins and publish: one global atomic per (cluster, feature) per CTA
        for (int i = threadIdx.x; i < K * 32; i += KM_THREADS) {
            const int j = i >> 5, dl = i & 31;
            if (j < k && dl < nd) {
                long long v = 0;
#pragma unroll
                for (int w = 0; w < KM_WARPS; ++w) v += s_acc[(size_t)w * K * 32 + i];
                if (v) atomicAdd(reinterpret_cast<unsigned long long *>(P.sums + ((size_t)b * k + j) * D + d0 + dl),
                                 (unsigned long long)v);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < k && s_cnt[threadIdx.x]) atomicAdd(P.counts + b * k + threadIdx.x, s_cnt[threadIdx.x]);

    // ---- last CTA of this image: sums -> next centroids, and reset the accumulators ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(P.done + b, 1) == P.chunks - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float *cout = P.cent_out + (size_t)b * k * D;
    long long *sums = P.sums + (size_t)b * k * D;
    for (int i = threadIdx.x; i < k * D; i += KM_THREADS) {
        const int j = i / D;
        const int cnt = __ldcg(P.counts + b * k + j);
        float c = cin[i];
        if (cnt > 0) {
            const double den = __dmul_rn((double)cnt, (double)P.fix_scale);
            c = __double2float_rn(__ddiv_rn((double)__ldcg(sums + i), den));
        }
        cout[i] = c;
        sums[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x < k) P.counts[b * k + threadIdx.x] = 0;
    if (threadIdx.x == 0) P.done[b] = 0;
}

template <int K>
int launch_pass(const KmParams &P, int B, cudaStream_t st)
{
    const size_t smem = sizeof(float) * (size_t)((P.D * K + 3) & ~3) + sizeof(long long) * (size_t)KM_WARPS * K * 32 +
                        sizeof(int) * (size_t)32 * KM_QSTR;
    static bool attr_set = false;
    if (!attr_set) {
        GCIS_CUDA_TRY(cudaFuncSetAttribute(km_pass_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (smem > 200 * 1024) return set_error(GCIS_E_INVALID, "kmeans: D=%d too large for shared memory", P.D);
    km_pass_kernel<K><<<dim3(P.chunks, B), KM_THREADS, smem, st>>>(P);
    GCIS_LAUNCH_CHECK();
    return GCIS_OK;
}

}  // namespace

size_t kmeans_workspace_bytes(int B, int D, int k)
{
    return sizeof(float) * (size_t)2 * B * k * D + sizeof(long long) * (size_t)B * k * D + sizeof(int) * (size_t)B * (k + 1) + 64;
}

// d_ws: workspace of kmeans_workspace_bytes(B, D, k), 16-byte aligned.
int kmeans_launch(const float *d_feat, int B, int D, int N, int k, int iters, int fix_shift,
                  const int32_t *d_init_idx, int32_t *d_labels, float *d_centroids, void *d_ws, cudaStream_t st)
{
    if (k < 1 || k > 32) return set_error(GCIS_E_INVALID, "kmeans: k=%d outside 1..32", k);
    if (iters < 1) return set_error(GCIS_E_INVALID, "kmeans: iters=%d < 1", iters);
    if (fix_shift < 0 || fix_shift > 30) return set_error(GCIS_E_INVALID, "kmeans: fix_shift=%d outside 0..30", fix_shift);
    if (B > 65535) return set_error(GCIS_E_INVALID, "kmeans: B=%d > 65535 per call", B);
    char *w = static_cast<char *>(d_ws);
    long long *sums = reinterpret_cast<long long *>(w);
    w += sizeof(long long) * (size_t)B * k * D;
    float *cent[2];
    cent[0] = reinterpret_cast<float *>(w); w += sizeof(float) * (size_t)B * k * D;
    cent[1] = reinterpret_cast<float *>(w); w += sizeof(float) * (size_t)B * k * D;
    int *counts = reinterpret_cast<int *>(w); w += sizeof(int) * (size_t)B * k;
    int *done = reinterpret_cast<int *>(w);

    km_init_kernel<<<B, 256, 0, st>>>(d_feat, d_init_idx, cent[0], sums, counts, done, D, N, k);
    GCIS_LAUNCH_CHECK();
    KmParams P;
    P.feat = d_feat; P.sums = sums; P.counts = counts; P.done = done;
    P.D = D; P.N = N; P.k = k; P.chunks = ceil_div(N, KM_PX);
    P.fix_scale = (float)(1u << fix_shift);
    for (int t = 0; t < iters; ++t) {
        P.cent_in = cent[t & 1];
        P.cent_out = cent[(t + 1) & 1];
        P.labels = (t == iters - 1) ? d_labels : nullptr;
        int rc;
        if (k <= 8) rc = launch_pass<8>(P, B, st);
        else if (k <= 16) rc = launch_pass<16>(P, B, st);
        else rc = launch_pass<32>(P, B, st);
        if (rc) return rc;
    }
    if (d_centroids)
        GCIS_CUDA_TRY(cudaMemcpyAsync(d_centroids, cent[iters & 1], sizeof(float) * (size_t)B * k * D,
                                      cudaMemcpyDeviceToDevice, st));
    return GCIS_OK;
}

}  // namespace gcis
