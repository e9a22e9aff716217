// kmeans.cu — Lloyd iterations: one fused assign + centroid-update kernel per iteration.
//
// Reference: none (segmenter slot, BSD_metrics/script.py:30; spec in DESIGN.md §3.4).
// Arithmetic contract, restated bit-for-bit by oracle/gcis_oracle.c:orc_kmeans:
//   m_jd = -2 c_jd;  cn_j = (float) sum_d (double)c_jd^2 (d ascending)
//   score_j(x) = fmaf(x_{D-1}, m_{j,D-1}, ... fmaf(x_0, m_j0, cn_j))        fp32 FMA chain
//   label = lowest j with the minimum score
//   centroid sums are exact int64 sums of q = rint(x * 2^fix_shift) -> the result does not
//   depend on the grid, the tile order or the number of GPUs
//   c_jd <- (float)((double)sum_jd / ((double)count_j * 2^fix_shift)); empty clusters keep c_jd
// With per-feature normalisation (z_d = a_d x_d + b_d, `affine` = [B][D]{a, b}) the clustering runs on z
// without ever materialising it - no extra HBM traffic: the affine map is folded into the score table,
//   c_jd (a centroid in z space) <- (float)((double)a_d * (double)cx_jd + (double)b_d),  cx = the x-space mean above,
//   m'_jd = a_d * (-2 c_jd) (fp32);  cn'_j = (float) sum_d [ (double)c_jd^2 + (double)b_d * (double)(-2 c_jd) ]  (d ascending),
//   score_j(x) = the same fp32 FMA chain over the RAW features with m' and cn'.
//
// Layout: feat [B][D][plane_stride] planar fp32 (a warp load is 512 contiguous bytes of one
// feature plane); centroids [B][k][D]; per-image score table prep = {m [D][K], cn [K]};
// sums [B][k][D] int64; counts [B][k] int32; labels kept between iterations as one byte/pixel.
//
// Two pass kernels share this contract (kmeans_launch picks one; both update the sums INCREMENTALLY:
// the sums are exact integers, so sum_new = sum_old + q(pixels that joined) - q(pixels that left) is
// bit-identical to a full recomputation, and after the first iterations few pixels change label):
//
//   km_tile_kernel<K,TP,V>  (default when a TP-pixel tile of all D planes fits in shared memory twice
//     per SM and the planes are padded/aligned): the tile is pulled in with TMA tensor boxes, scored in
//     arrival order and the label changes are folded into the sums straight from the resident tile by an
//     exact int8 tensor-core GEMM; km_finalize_kernel turns sums into the next centroids.  See the
//     comment block above the kernel.
//
//   km_pass_kernel<K,VEC>  (fallback: large D*K, or caller-supplied unpadded [D][H][W] tensors):
//     one CTA owns 256*VEC consecutive pixels.
//     Phase A (thread = VEC pixels) streams the D planes exactly once through a shared-memory ring
//     filled by the TMA engine (cp.async.bulk of one 4 KB row per plane, completion on mbarriers;
//     VEC == 4) and keeps only the scores; the FMA chain runs as packed fma.rn.f32x2 over cluster
//     pairs (one IEEE fp32 FMA per lane, same result as scalar FFMA).
//     Phase B compacts the changed pixels in the CTA and re-reads their features from L2/HBM:
//       sparse path (<= 32 changed pixels in the tile): lane = feature, so each (cluster, feature)
//         delta lives in exactly one lane and is published with one global atomic; no shared
//         staging and no barrier;
//       dense path: the per-cluster sums are a small exact integer GEMM on the tensor cores
//         (mma.sync m16n8k32 s8 x u8: one-hot label differences times the byte digits of q + 2^31),
//         128 changed pixels per round, each warp staging and consuming its own digit columns.
//     The last CTA of an image to finish (ticket counter) turns sums into the next centroids and
//     the next score table; only its first warp stays for that serial tail.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "async.cuh"

namespace gcis {

namespace {

constexpr int KM_THREADS = 256;
constexpr int KM_WARPS = KM_THREADS / 32;
constexpr int KM_ROUND = 128;               // changed pixels per tensor-core round (dense path)
#ifndef KM_SPARSE_N
#define KM_SPARSE_N 32
#endif
constexpr int KM_SPARSE = KM_SPARSE_N;      // sparse path handles up to this many changed pixels
#ifndef KM_MINB
#define KM_MINB 4
#endif
constexpr int KM_NONE = 255;                // "no previous label"

struct KmParams {
    const float *feat;
    size_t img_stride;      // floats between images
    int plane_stride;       // floats between feature planes
    float *cent;            // [B][k][D]
    float *prep;            // [B][D*K + K]: m (transposed, zero-padded to K) then cn (+inf padded)
    long long *sums;
    int *counts;
    int *done;
    unsigned char *lab8;    // [B][lab_stride] labels of the previous iteration
    int lab_stride;
    int32_t *labels_out;    // [B][N], written when non-null (last iteration)
    int D, N, k, chunks, first;
    float fix_scale;
    int pass;
    const float *affine;    // [B][D][2] {a_d, b_d} of the per-feature normalisation, or null
};

#ifdef KM_TRACE   // timing experiment only: per-CTA phase timestamps of the first CTAs of each pass
constexpr int KM_TR_PASSES = 20, KM_TR_CTAS = 1024, KM_TR_SLOTS = 8;
__device__ long long km_trace_buf[KM_TR_PASSES][KM_TR_CTAS][KM_TR_SLOTS];
__device__ __forceinline__ long long km_gtime()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define KM_TR(slot)                                                                                          \
    do {                                                                                                     \
        const int cta_ = blockIdx.y * gridDim.x + blockIdx.x;                                                \
        if (threadIdx.x == 0 && cta_ < KM_TR_CTAS && P.pass < KM_TR_PASSES) {                                \
            km_trace_buf[P.pass][cta_][slot] = clock64();                                                    \
            if (slot == 0) km_trace_buf[P.pass][cta_][7] = km_gtime();                                       \
        }                                                                                                    \
    } while (0)
#else
#define KM_TR(slot) do { } while (0)
#endif

// Score table of one image from its centroids (all threads of the CTA; cent must be visible).
__device__ __forceinline__ void km_write_prep(const float *cent, float *prep, int D, int k, int K, int tid, int nthr,
                                              const float *affine = nullptr)
{
    for (int i = tid; i < D * K; i += nthr) {
        const int d = i / K, j = i - d * K;
        float m = j < k ? -2.0f * cent[j * D + d] : 0.f;
        if (affine) m = __fmul_rn(affine[2 * d], m);
        prep[i] = m;
    }
    if (tid < K) {
        const int j = tid;
        float cn = __int_as_float(0x7f800000);  // +inf: padded clusters never win
        if (j < k) {
            double acc = 0.0;
            for (int d = 0; d < D; ++d) {
                const double c = (double)cent[j * D + d];
                acc = __dadd_rn(acc, __dmul_rn(c, c));
                if (affine) acc = __dadd_rn(acc, __dmul_rn((double)affine[2 * d + 1], __dmul_rn(-2.0, c)));
            }
            cn = (float)acc;
        }
        prep[D * K + j] = cn;
    }
}

// centroid in the clustered space from the x-space value (identity without normalisation)
__device__ __forceinline__ float km_affine(const float *affine, int d, float cx)
{
    if (!affine) return cx;
    return __double2float_rn(__fma_rn((double)affine[2 * d], (double)cx, (double)affine[2 * d + 1]));
}

__global__ void km_init_kernel(const float *__restrict__ feat, size_t img_stride, int plane_stride,
                               const int32_t *__restrict__ init_idx, float *cent, float *prep, long long *sums,
                               int *counts, int *done, int D, int N, int k, int K, const float *affine_all)
{
    const int b = blockIdx.x;
    float *c = cent + (size_t)b * k * D;
    const float *affine = affine_all ? affine_all + (size_t)b * D * 2 : nullptr;
    for (int i = threadIdx.x; i < k * D; i += blockDim.x) {
        const int j = i / D, d = i - j * D;
        int p = init_idx[b * k + j];
        p = min(max(p, 0), N - 1);
        c[i] = km_affine(affine, d, feat[(size_t)b * img_stride + (size_t)d * plane_stride + p]);
        sums[(size_t)b * k * D + i] = 0;
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) counts[b * k + i] = 0;
    if (threadIdx.x == 0) done[b] = 0;
    __syncthreads();
    km_write_prep(c, prep + (size_t)b * (D * K + K), D, k, K, threadIdx.x, blockDim.x, affine);
}

// acc.{lo,hi} = a.{lo,hi} * b + acc.{lo,hi}, each lane one IEEE fp32 FMA (round to nearest even)
__device__ __forceinline__ void ffma2(unsigned long long &acc, unsigned long long a, float b)
{
    unsigned long long bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(bb));
}

// acc[ln] += v; acc[lo] -= v with warp-uniform labels: registers for K <= 8, private bins otherwise
template <int K>
__device__ __forceinline__ void km_move(long long (&acc)[K <= 8 ? K : 1], long long *bins, int ln, int lo, long long v)
{
    if constexpr (K <= 8) {
        switch (ln) {
            case 0: acc[0] += v; break;
            case 1: acc[1 % K] += v; break;
            case 2: acc[2 % K] += v; break;
            case 3: acc[3 % K] += v; break;
            case 4: acc[4 % K] += v; break;
            case 5: acc[5 % K] += v; break;
            case 6: acc[6 % K] += v; break;
            default: acc[7 % K] += v; break;
        }
        switch (lo) {
            case 0: acc[0] -= v; break;
            case 1: acc[1 % K] -= v; break;
            case 2: acc[2 % K] -= v; break;
            case 3: acc[3 % K] -= v; break;
            case 4: acc[4 % K] -= v; break;
            case 5: acc[5 % K] -= v; break;
            case 6: acc[6 % K] -= v; break;
            case 7: acc[7 % K] -= v; break;
            default: break;  // KM_NONE
        }
    } else {
        bins[ln * 32] += v;
        if (lo != KM_NONE) bins[lo * 32] -= v;
    }
}

// ---- last CTA of an image: sums -> next centroids and next score table ----
// Called by all threads after their atomics.  The barrier orders every atomic of this CTA before
// lane 0's fence + ticket (fences are cumulative); only warp 0 stays for the ticket, so the other
// warps never wait on its round trip.  `s_c` is k*D floats of shared scratch (the score table is
// no longer needed: every other warp of this CTA has left).
template <int K>
__device__ __forceinline__ void km_ticket_finalize(const KmParams &P, int b, float *s_c)
{
    const int D = P.D, k = P.k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    KM_TR(4);
    if (warp != 0) return;
    int last = 0;
    if (lane == 0) {
        __threadfence();
        last = atomicAdd(P.done + b, 1) == P.chunks - 1;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    KM_TR(5);
    if (!last) return;
    __threadfence();
    // This warp is the image's serial tail of the pass, so keep it short: independent loads are
    // batched, and the new centroids are staged in shared memory for the table rebuild.
    float *cent = P.cent + (size_t)b * k * D;
    const long long *sums = P.sums + (size_t)b * k * D;
    const float *affine = P.affine ? P.affine + (size_t)b * D * 2 : nullptr;
    const int my_cnt = lane < k ? __ldcg(P.counts + b * k + lane) : 0;
#pragma unroll 4
    for (int i0 = 0; i0 < k * D; i0 += 32) {
        const int i = i0 + lane;
        const bool ok = i < k * D;
        const int cnt = __shfl_sync(0xffffffffu, my_cnt, ok ? i / D : 0);
        if (ok) {
            const long long s = __ldcg(sums + i);
            float c;
            if (cnt > 0) c = km_affine(affine, i % D, __double2float_rn(__ddiv_rn((double)s, __dmul_rn((double)cnt, (double)P.fix_scale))));
            else c = cent[i];
            cent[i] = c;
            s_c[i] = c;
        }
    }
    if (lane == 0) P.done[b] = 0;
    __syncwarp();
    km_write_prep(s_c, P.prep + (size_t)b * (D * K + K), D, k, K, lane, 32, affine);
    KM_TR(6);
}

#ifndef KM_STAGES_N
#define KM_STAGES_N 3
#endif
#ifndef KM_DS_N
#define KM_DS_N 4
#endif
constexpr int KM_STAGES = KM_STAGES_N;   // ring depth of the phase-A feature pipeline (VEC == 4)
constexpr int KM_DS = KM_DS_N;           // feature planes per stage

constexpr int KM_BCOLS = 128;                // dense path: 32 features x 4 byte digits per slab
constexpr int KM_BSTR = KM_ROUND / 4 + 4;    // words per digit column (4 entries per word), padded: conflict-free

// shared bytes of the phase-B structures (warp bins of the sparse path for K > 8, digit slab of
// the dense path); they alias the ring
__host__ __device__ constexpr size_t km_phase_b_bytes(int K)
{
    return sizeof(long long) * KM_WARPS * K * 32 + sizeof(int) * KM_BCOLS * KM_BSTR;
}
__host__ __device__ constexpr size_t km_smem_bytes(int K, int vec, int D)
{
    const size_t ring = vec == 4 ? sizeof(float) * KM_STAGES * KM_DS * KM_THREADS * 4 : 0;
    const size_t pb = km_phase_b_bytes(K);
    const size_t front = ((ring > pb ? ring : pb) + 15) & ~(size_t)15;
    return front + (size_t)4 * KM_THREADS * vec + sizeof(float) * (size_t)((D * K + K + 3) & ~3);
}

template <int K, int VEC>
__global__ void __launch_bounds__(KM_THREADS, K <= 8 ? KM_MINB : (K <= 16 ? 2 : 1)) km_pass_kernel(const __grid_constant__ KmParams P)
{
    constexpr int TILE = KM_THREADS * VEC;
    constexpr int NACC = K <= 8 ? K : 1;
    constexpr int RING_FLOATS = VEC == 4 ? KM_STAGES * KM_DS * TILE : 0;
    extern __shared__ __align__(128) unsigned char km_smem[];
    // [ring | (aliased by the dense phase-B path) s_acc, s_q] [changed-pixel list] [s_m: m [D][K], cn [K]]
    float *s_ring = reinterpret_cast<float *>(km_smem);
    long long *s_acc = reinterpret_cast<long long *>(km_smem);                                // [warps][K][32]
    int *s_q = reinterpret_cast<int *>(s_acc + KM_WARPS * K * 32);                            // digit slab [KM_BCOLS][KM_BSTR]
    constexpr size_t PHASE_B_BYTES = km_phase_b_bytes(K);
    constexpr size_t FRONT_BYTES =
        ((RING_FLOATS * sizeof(float) > PHASE_B_BYTES ? RING_FLOATS * sizeof(float) : PHASE_B_BYTES) + 15) & ~(size_t)15;
    unsigned short *s_ent = reinterpret_cast<unsigned short *>(km_smem + FRONT_BYTES);        // [TILE] changed pixels
    unsigned char *s_new = reinterpret_cast<unsigned char *>(s_ent + TILE);                   // [TILE]
    unsigned char *s_old = s_new + TILE;                                                      // [TILE]
    float *s_m = reinterpret_cast<float *>(km_smem + FRONT_BYTES + 4 * TILE);
    __shared__ __align__(8) unsigned long long s_full[KM_STAGES], s_empty[KM_STAGES];
    __shared__ int s_cnt[K];
    __shared__ int s_nchg;

    const int b = blockIdx.y;
    const int D = P.D, N = P.N, k = P.k, stride = P.plane_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *feat = P.feat + (size_t)b * P.img_stride;
    const int tile0 = blockIdx.x * TILE;
    KM_TR(0);
    float *s_cn = s_m + D * K;

    // Feature planes arrive through a KM_STAGES-deep shared-memory ring filled by the TMA engine
    // (cp.async.bulk, one 4 KB row per plane, completion on an mbarrier): loads of the next
    // stages are in flight while this one is consumed, at no register cost (VEC == 4 only).
    const int n_it = (D + KM_DS - 1) / KM_DS;
    const uint32_t row_bytes = (uint32_t)min(TILE, stride - tile0) * 4u;   // multiple of 16: planes are padded
    const float *src0 = feat + tile0;
    auto issue = [&](int it) {   // one elected thread
        const int s = it % KM_STAGES;
        const int nd = min(KM_DS, D - it * KM_DS);
        const uint32_t bar = smem_u32(&s_full[s]);
        mbar_expect_tx(bar, row_bytes * nd);
        for (int dd = 0; dd < nd; ++dd)
            bulk_g2s(smem_u32(s_ring + ((size_t)s * KM_DS + dd) * TILE), src0 + (size_t)(it * KM_DS + dd) * stride,
                     row_bytes, bar);
    };
    if (threadIdx.x == 0) {
        s_nchg = 0;
        if constexpr (VEC == 4) {
            for (int s = 0; s < KM_STAGES; ++s) {
                mbar_init(smem_u32(&s_full[s]), 1);
                mbar_init(smem_u32(&s_empty[s]), KM_WARPS);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // the first stages start flowing before the score table is even loaded
            for (int it = 0; it < KM_STAGES - 1 && it < n_it; ++it) issue(it);
        }
    }
    const int p0 = tile0 + threadIdx.x * VEC;
    unsigned prev = 0xffffffffu;   // labels of the previous iteration (fetched early: needed only after the stream)
    unsigned char *lab = P.lab8 + (size_t)b * P.lab_stride;
    if (!P.first) {
        if (VEC == 4) prev = *reinterpret_cast<const unsigned *>(lab + min(p0, P.lab_stride - 4));
        else prev = lab[min(p0, N - 1)];
    }
    {
        const float4 *src = reinterpret_cast<const float4 *>(P.prep + (size_t)b * (D * K + K));
        float4 *dst = reinterpret_cast<float4 *>(s_m);
        for (int i = threadIdx.x; i < (D * K + K) / 4; i += KM_THREADS) dst[i] = src[i];
    }
    if (threadIdx.x < K) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    KM_TR(1);

    // ---- phase A: scores and labels for VEC consecutive pixels per thread ----
    // planes are padded (VEC == 4) so a clamped vector load never leaves the plane
    const int pl = VEC == 4 ? min(p0, stride - VEC) : min(p0, N - 1);
    unsigned long long s2[VEC][K / 2];
#pragma unroll
    for (int i = 0; i < K / 2; ++i) {
        unsigned long long c2;
        asm("mov.b64 %0, {%1, %2};" : "=l"(c2) : "f"(s_cn[2 * i]), "f"(s_cn[2 * i + 1]));
#pragma unroll
        for (int v = 0; v < VEC; ++v) s2[v][i] = c2;
    }
    if constexpr (VEC == 4) {
        for (int it = 0; it < n_it; ++it) {
            const int s = it % KM_STAGES;
            // refill the stage that was consumed in the previous iteration
            if (threadIdx.x == 0 && it + KM_STAGES - 1 < n_it) {
                const int nxt = it + KM_STAGES - 1;
                if (nxt >= KM_STAGES) mbar_wait(smem_u32(&s_empty[nxt % KM_STAGES]), ((nxt / KM_STAGES) - 1) & 1);
                issue(nxt);
            }
            mbar_wait(smem_u32(&s_full[s]), (it / KM_STAGES) & 1);
            const float4 *xrow = reinterpret_cast<const float4 *>(s_ring + (size_t)s * KM_DS * TILE) + threadIdx.x;
            const ulonglong2 *mrow = reinterpret_cast<const ulonglong2 *>(s_m + (size_t)it * KM_DS * K);
            auto plane = [&](int dd) {   // one feature plane: 4 pixels x K clusters
                const float4 t = xrow[dd * (TILE / 4)];
#pragma unroll
                for (int q = 0; q < K / 4; ++q) {
                    const ulonglong2 mm = mrow[dd * (K / 4) + q];
                    ffma2(s2[0][2 * q], mm.x, t.x); ffma2(s2[0][2 * q + 1], mm.y, t.x);
                    ffma2(s2[1 % VEC][2 * q], mm.x, t.y); ffma2(s2[1 % VEC][2 * q + 1], mm.y, t.y);
                    ffma2(s2[2 % VEC][2 * q], mm.x, t.z); ffma2(s2[2 % VEC][2 * q + 1], mm.y, t.z);
                    ffma2(s2[3 % VEC][2 * q], mm.x, t.w); ffma2(s2[3 % VEC][2 * q + 1], mm.y, t.w);
                }
            };
            if (it * KM_DS + KM_DS <= D) {   // full stage: no per-plane branches
#pragma unroll
                for (int dd = 0; dd < KM_DS; ++dd) plane(dd);
            } else {
                for (int dd = 0; dd < D - it * KM_DS; ++dd) plane(dd);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_empty[s]));
        }
    } else {
        const float *xp = feat + pl;
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
            const float x = __ldg(xp + (size_t)d * stride);
            const ulonglong2 *mrow = reinterpret_cast<const ulonglong2 *>(s_m + d * K);
#pragma unroll
            for (int q = 0; q < K / 4; ++q) {
                const ulonglong2 mm = mrow[q];
                ffma2(s2[0][2 * q], mm.x, x);
                ffma2(s2[0][2 * q + 1], mm.y, x);
            }
        }
    }
    KM_TR(2);
    {
        int newl[VEC], oldl[VEC];
        bool chg[VEC];
        unsigned packed = 0;
        int nchg = 0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float bs, sj[2];
            int best = 0;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(sj[0]), "=f"(sj[1]) : "l"(s2[v][0]));
            bs = sj[0];
            if (sj[1] < bs) { bs = sj[1]; best = 1; }
#pragma unroll
            for (int i = 1; i < K / 2; ++i) {
                asm("mov.b64 {%0, %1}, %2;" : "=f"(sj[0]), "=f"(sj[1]) : "l"(s2[v][i]));
                if (sj[0] < bs) { bs = sj[0]; best = 2 * i; }
                if (sj[1] < bs) { bs = sj[1]; best = 2 * i + 1; }
            }
            const bool valid = p0 + v < N;
            newl[v] = best;
            oldl[v] = P.first ? KM_NONE : (int)((prev >> (8 * v)) & 0xffu);
            chg[v] = valid && best != oldl[v];
            nchg += chg[v];
            packed |= (unsigned)best << (8 * v);
            if (valid && P.labels_out) P.labels_out[(size_t)b * N + p0 + v] = best;
        }
        if (VEC == 4) {
            if (p0 + 3 < P.lab_stride) *reinterpret_cast<unsigned *>(lab + p0) = packed;
        } else if (p0 < N) {
            lab[p0] = (unsigned char)packed;
        }
        // everything below only concerns changed pixels: skip whole warps without any
        if (__any_sync(0xffffffffu, nchg != 0)) {
            // cluster population deltas: lane j of every warp owns cluster j
            int dcnt = 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const unsigned in = __ballot_sync(0xffffffffu, chg[v] && newl[v] == j);
                    const unsigned out = __ballot_sync(0xffffffffu, chg[v] && oldl[v] == j);
                    if (lane == j) dcnt += __popc(in) - __popc(out);
                }
            if (lane < K && dcnt) atomicAdd(&s_cnt[lane], dcnt);
            // compact the changed pixels in pixel order
            int incl = nchg;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            int base = 0;
            if (lane == 31) base = atomicAdd(&s_nchg, incl);
            base = __shfl_sync(0xffffffffu, base, 31);
            int pos = base + incl - nchg;
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (chg[v]) {
                    s_ent[pos] = (unsigned short)(threadIdx.x * VEC + v);
                    s_new[pos] = (unsigned char)newl[v];
                    s_old[pos] = (unsigned char)oldl[v];
                    ++pos;
                }
        }
    }
    __syncthreads();
    KM_TR(3);

    // ---- phase B: exact fixed-point centroid deltas from the changed pixels ----
#ifdef KM_SKIP_B   // timing experiment only: wrong results
    const int n_chg = 0;
#else
    const int n_chg = s_nchg;
#endif
#ifdef KM_SKIP_SPARSE   // timing experiment only: wrong results
    if (n_chg <= KM_SPARSE) {
    } else
#endif
    if (n_chg > 0 && n_chg <= KM_SPARSE) {
        // sparse: lane = feature, so every (cluster, feature) delta lives in exactly one lane and goes
        // straight from L1/L2 to one global atomic: no shared staging, no barrier, and the warps
        // beyond ceil(D/32) are already done
        for (int d = warp * 32 + lane; d < D; d += KM_THREADS) {
            long long acc[NACC];
            long long *bins = s_acc + (size_t)warp * K * 32 + lane;
            if constexpr (K <= 8) {
#pragma unroll
                for (int j = 0; j < NACC; ++j) acc[j] = 0;
            } else {
                for (int j = 0; j < K; ++j) bins[j * 32] = 0;
            }
            const float *xd = feat + (size_t)d * stride + tile0;
            for (int e0 = 0; e0 < n_chg; e0 += 8) {
                float xv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) xv[u] = e0 + u < n_chg ? __ldg(xd + s_ent[e0 + u]) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (e0 + u < n_chg)
                        km_move<K>(acc, bins, s_new[e0 + u], s_old[e0 + u], (long long)__float2int_rn(xv[u] * P.fix_scale));
            }
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(P.sums + (size_t)b * k * D + d);
            if constexpr (K <= 8) {
#pragma unroll
                for (int j = 0; j < NACC; ++j)
                    if (j < k && acc[j]) atomicAdd(dst + (size_t)j * D, (unsigned long long)acc[j]);
            } else {
                for (int j = 0; j < k; ++j)
                    if (bins[j * 32]) atomicAdd(dst + (size_t)j * D, (unsigned long long)bins[j * 32]);
            }
        }
    } else if (n_chg > 0) {
        // Dense: the per-cluster sums are a small integer GEMM, so they run on the tensor cores and stay
        // exact:  S[j][d] += sum_e A[j][e] * q'[e][d]  with A[j][e] = [new_e = j] - [old_e = j] (int8) and
        // q' = q + 2^31 split into four unsigned byte digits (mma.sync m16n8k32 s8 x u8 -> s32).  The
        // 2^31 offset is removed with the cluster's population delta: sum_e A[j][e] = s_cnt[j].
        const int n_pad = (n_chg + KM_ROUND - 1) / KM_ROUND * KM_ROUND;
        for (int i = n_chg + threadIdx.x; i < n_pad; i += KM_THREADS) { s_new[i] = KM_NONE; s_old[i] = KM_NONE; }
        __syncthreads();   // padding visible; from here on every warp only touches its own 16 digit columns
        unsigned *s_b = reinterpret_cast<unsigned *>(s_q);   // [KM_BCOLS][KM_BSTR]: column (feature, digit), word = 4 entries
        constexpr int MT = (K + 15) / 16;
        const int g = lane >> 2, kq = lane & 3;
        for (int d0 = 0; d0 < D; d0 += 32) {
            const int nd = min(32, D - d0);
            int c[MT][2][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) c[mt][nt][i] = 0;
            for (int r0 = 0; r0 < n_pad; r0 += KM_ROUND) {
                __syncwarp();  // this warp's columns are free again
                {   // stage: thread = 4 consecutive entries x 4 features (the warp's own); bytes transposed in registers
                    const int eq = threadIdx.x & 31, fg = threadIdx.x >> 5;
                    const float *xp[4];
                    bool ok[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = r0 + eq * 4 + u;
                        ok[u] = e < n_chg;
                        xp[u] = feat + (size_t)d0 * stride + tile0 + (ok[u] ? (int)s_ent[e] : 0);
                    }
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        const int dd = fg * 4 + f;
                        unsigned q[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            q[u] = (dd < nd && ok[u])
                                       ? ((unsigned)__float2int_rn(__ldg(xp[u] + (size_t)dd * stride) * P.fix_scale) ^ 0x80000000u)
                                       : 0u;
                        const unsigned t0 = __byte_perm(q[0], q[1], 0x5140), t1 = __byte_perm(q[2], q[3], 0x5140);
                        const unsigned t2 = __byte_perm(q[0], q[1], 0x7362), t3 = __byte_perm(q[2], q[3], 0x7362);
                        unsigned *dst = s_b + (dd * 4) * KM_BSTR + eq;
                        dst[0 * KM_BSTR] = __byte_perm(t0, t1, 0x5410);
                        dst[1 * KM_BSTR] = __byte_perm(t0, t1, 0x7632);
                        dst[2 * KM_BSTR] = __byte_perm(t2, t3, 0x5410);
                        dst[3 * KM_BSTR] = __byte_perm(t2, t3, 0x7632);
                    }
                }
                __syncwarp();
                // warp w owns digit columns 16w .. 16w+15 (features 4w .. 4w+3 of the slab)
                const unsigned *newp = reinterpret_cast<const unsigned *>(s_new + r0);
                const unsigned *oldp = reinterpret_cast<const unsigned *>(s_old + r0);
#pragma unroll
                for (int ks = 0; ks < KM_ROUND / 32; ++ks) {
                    const unsigned n0 = newp[ks * 8 + kq], n1 = newp[ks * 8 + 4 + kq];
                    const unsigned o0 = oldp[ks * 8 + kq], o1 = oldp[ks * 8 + 4 + kq];
                    unsigned bf[2][2];
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const unsigned *col = s_b + ((2 * warp + nt) * 8 + g) * KM_BSTR + ks * 8 + kq;
                        bf[nt][0] = col[0];
                        bf[nt][1] = col[4];
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const unsigned r0w = (unsigned)(mt * 16 + g) * 0x01010101u, r1w = r0w + 0x08080808u;
                        const unsigned a0 = (__vcmpeq4(n0, r0w) & 0x01010101u) | __vcmpeq4(o0, r0w);
                        const unsigned a1 = (__vcmpeq4(n0, r1w) & 0x01010101u) | __vcmpeq4(o0, r1w);
                        const unsigned a2 = (__vcmpeq4(n1, r0w) & 0x01010101u) | __vcmpeq4(o1, r0w);
                        const unsigned a3 = (__vcmpeq4(n1, r1w) & 0x01010101u) | __vcmpeq4(o1, r1w);
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
                            asm volatile(
                                "mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                : "+r"(c[mt][nt][0]), "+r"(c[mt][nt][1]), "+r"(c[mt][nt][2]), "+r"(c[mt][nt][3])
                                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bf[nt][0]), "r"(bf[nt][1]));
                    }
                }
            }
            // digits -> int64, remove the offset, publish: one global atomic per (cluster, feature) per CTA
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int hrow = 0; hrow < 2; ++hrow) {
                        long long part = (long long)c[mt][nt][2 * hrow] + ((long long)c[mt][nt][2 * hrow + 1] << 8);
                        if (kq & 1) part <<= 16;
                        part += __shfl_xor_sync(0xffffffffu, part, 1);
                        const int j = mt * 16 + hrow * 8 + g;
                        const int d = d0 + (2 * warp + nt) * 2 + (kq >> 1);
                        if (!(kq & 1) && j < k && d < D) {
                            const long long v = part - ((long long)s_cnt[j] << 31);
                            if (v) atomicAdd(reinterpret_cast<unsigned long long *>(P.sums + ((size_t)b * k + j) * D + d),
                                             (unsigned long long)v);
                        }
                    }
        }
    }
    if (threadIdx.x < k && s_cnt[threadIdx.x]) atomicAdd(P.counts + b * k + threadIdx.x, s_cnt[threadIdx.x]);

    km_ticket_finalize<K>(P, b, s_m);
}

// =================================================================================================
// Tile-resident pass (the default when a TP-pixel tile of all D planes fits twice in an SM's shared
// memory): one CTA owns TP consecutive pixels and pulls ALL of its D plane rows into shared memory
// with KT_GROUPS TMA tensor boxes (ppg planes x TP pixels each, D x TP x 4 bytes in flight per CTA,
// completion on one mbarrier per box), scores them in arrival order (thread = V consecutive
// pixels, same FMA chain), and then updates the centroid sums straight from the
// resident tile, so no feature is read from L2/HBM twice in any pass, including the first one where
// every pixel "changes".  DRAM traffic per pass = the algorithmic N * D * 4 bytes.
//   update = exact integer GEMM on the tensor cores (mma.sync m16n8k32 s8 x u8 -> s32):
//     S[j][d] += sum_p A[j][p] * digit_b(q'[p][d]),  A[j][p] = [new_p = j] - [old_p = j],
//     q' = rint(x * 2^shift) + 2^31 as four unsigned byte digits (offset removed with the cluster's
//     population delta, which one more MMA against a column of ones yields).  The B fragments are converted in registers from two 128-bit loads of a
//     plane row (thread (g, kq) of the warp: feature 8*fg + g, pixels 4kq..4kq+3 and 16+4kq..),
//     the four digit planes of one conversion feed four MMAs.  Only quads (4 aligned pixels) that
//     hold a changed pixel are listed and fed to the MMAs, eight quads per k-block.
//   Tiles with <= KT_SPARSE changed pixels (most tiles after the first passes; K <= 8) skip the digits
//   and the MMAs: thread = feature keeps its deltas in registers and publishes <= k atomics.
//   The score table m [D][K] is warp-uniform, so every plane costs each warp two 128-bit shared
//   loads of it on top of the pixel values: V pixels per thread amortise that load-pipe cost.
// =================================================================================================
constexpr int KT_GROUPS = 3;   // arrival groups (one TMA box + one mbarrier each) per tile; measured: 1: 31.1, 2: 30.9,
                               // 3: 30.7, 4: 31.2, 8: 32.1, 12: 34.1 ms per 200 images (each TMA issue costs the elected thread ~100 cycles)
#ifndef KT_SPARSE_N
#define KT_SPARSE_N 12
#endif
#ifndef KT_V
#define KT_V 2        // pixels per thread
#endif
#ifndef KT_UNROLL
#define KT_UNROLL 8   // planes per unrolled step of the score loop (2: 31.1, 4: 30.8, 8: 30.4, 24: 30.5 ms per 200 images)
#endif
constexpr int KT_UNR = KT_UNROLL;
constexpr int KT_SPARSE = KT_SPARSE_N;  // up to this many changed pixels per tile skip the tensor-core path (K <= 8)

template <int K, int TP>
struct __align__(16) KtState {
    unsigned long long bar[KT_GROUPS];
    unsigned char nw[TP], ol[TP];   // new / old label of changed pixels, KM_NONE otherwise
    unsigned char quads[TP / 4];    // quads (4 aligned pixels) with at least one changed pixel
    unsigned char plist[KT_SPARSE]; // the first changed pixels (sparse update)
    int nq, npx;
};

__host__ __device__ constexpr size_t kt_smem_bytes(int K, int TP, int D)
{
    const size_t state = (KT_GROUPS * 8 + 2 * TP + TP / 4 + KT_SPARSE + 8 + 15) & ~(size_t)15;
    return sizeof(float) * ((size_t)((D + KT_GROUPS - 1) / KT_GROUPS) * KT_GROUPS * TP + (size_t)((D * K + K + 3) & ~3)) + state;
}
constexpr size_t KT_SMEM_BUDGET = 111 * 1024;   // two CTAs per SM below this

// Programmatic dependent launch (the pass and finalize kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): a kernel lets its successor become resident early and
// the successor blocks at griddep_wait() until everything before it in the stream has completed and flushed.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// one TMA box: TP pixels x ppg planes of image b, landing densely as [ppg][TP]; rows or pixels
// outside the tensor are zero-filled and still counted in the transaction bytes
__device__ __forceinline__ void tma_box_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}


template <int K, int TP, int V>
__global__ void __launch_bounds__(TP / V, 2) km_tile_kernel(const __grid_constant__ KmParams P, const __grid_constant__ CUtensorMap tmap)
{
    constexpr int THREADS = TP / V, WARPS = THREADS / 32, MT = (K + 15) / 16;
    extern __shared__ __align__(128) unsigned char km_smem[];
    const int D = P.D, N = P.N, k = P.k;
    const int ppg = (D + KT_GROUPS - 1) / KT_GROUPS;   // planes per arrival group
    float *s_x = reinterpret_cast<float *>(km_smem);   // [KT_GROUPS * ppg][TP]
    float *s_m = s_x + (size_t)KT_GROUPS * ppg * TP;   // m [D][K], cn [K]
    // small state packed behind the tile (no separately rounded static segment): with D = 72, K = 8
    // three CTAs fit in an SM's 228 KB with no room to spare
    KtState<K, TP> &ss = *reinterpret_cast<KtState<K, TP> *>(s_m + ((D * K + K + 3) & ~3));
    unsigned long long *s_bar = ss.bar;
    unsigned char *s_new = ss.nw, *s_old = ss.ol, *s_quads = ss.quads, *s_plist = ss.plist;
    int &s_nq = ss.nq, &s_npx = ss.npx;

    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile0 = blockIdx.x * TP;
    KM_TR(0);

    griddep_launch_dependents();   // the finalize kernel may become resident behind the last wave of this pass
    if (threadIdx.x == 0) {
        for (int g = 0; g < KT_GROUPS; ++g) mbar_init(smem_u32(&s_bar[g]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // one TMA box of ppg planes x TP pixels per arrival group; the features are read-only for the whole
        // clustering, so these loads may start before the previous pass's finalize kernel has finished
        const uint32_t prep_bytes = (uint32_t)(D * K + K) * 4u, box_bytes = (uint32_t)(ppg * TP) * 4u;
        for (int g = 0; g < KT_GROUPS && g * ppg < D; ++g) {
            mbar_expect_tx(smem_u32(&s_bar[g]), box_bytes + (g == 0 ? prep_bytes : 0u));
            tma_box_3d(smem_u32(s_x + (size_t)g * ppg * TP), &tmap, tile0, g * ppg, b, smem_u32(&s_bar[g]));
        }
        griddep_wait();                // score table, labels, sums and counts of the previous pass are final
        bulk_g2s(smem_u32(s_m), P.prep + (size_t)b * (D * K + K), prep_bytes, smem_u32(&s_bar[0]));
    }
    griddep_wait();
    const int p0 = tile0 + threadIdx.x * V;
    unsigned char *lab = P.lab8 + (size_t)b * P.lab_stride;
    unsigned prev = 0xffffffffu;   // labels of the previous iteration, in flight while the tile arrives
    if (!P.first) {                // (lab_stride is a multiple of 16: a vector load never leaves the row)
        if constexpr (V == 4) prev = *reinterpret_cast<const unsigned *>(lab + min(p0, P.lab_stride - 4));
        else if constexpr (V == 2) prev = *reinterpret_cast<const unsigned short *>(lab + min(p0, P.lab_stride - 2));
        else prev = lab[min(p0, P.lab_stride - 1)];
    }
    if (threadIdx.x == 0) { s_nq = 0; s_npx = 0; }
    __syncthreads();   // barriers initialised for every waiter
    KM_TR(1);

    // ---- scores: planes in arrival order ----
    mbar_wait(smem_u32(&s_bar[0]), 0);
    unsigned long long s2[V][K / 2];
    {
        const float *s_cn = s_m + D * K;
#pragma unroll
        for (int i = 0; i < K / 2; ++i) {
            unsigned long long c2;
            asm("mov.b64 %0, {%1, %2};" : "=l"(c2) : "f"(s_cn[2 * i]), "f"(s_cn[2 * i + 1]));
#pragma unroll
            for (int v = 0; v < V; ++v) s2[v][i] = c2;
        }
    }
    for (int g = 0; g < KT_GROUPS; ++g) {
        const int d_lo = g * ppg, d_hi = min(D, d_lo + ppg);
        if (d_lo >= d_hi) break;
        if (g) mbar_wait(smem_u32(&s_bar[g]), 0);
        const float *xp = s_x + (size_t)d_lo * TP + threadIdx.x * V;
        const ulonglong2 *mrow = reinterpret_cast<const ulonglong2 *>(s_m + (size_t)d_lo * K);
#pragma unroll KT_UNR
        for (int d = d_lo; d < d_hi; ++d, xp += TP, mrow += K / 4) {
            float x[V];
            if constexpr (V == 4) {
                const float4 t = *reinterpret_cast<const float4 *>(xp);
                x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
            } else if constexpr (V == 2) {
                const float2 t = *reinterpret_cast<const float2 *>(xp);
                x[0] = t.x; x[1] = t.y;
            } else {
                x[0] = *xp;
            }
#pragma unroll
            for (int q = 0; q < K / 4; ++q) {
                const ulonglong2 mm = mrow[q];
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    ffma2(s2[v][2 * q], mm.x, x[v]);
                    ffma2(s2[v][2 * q + 1], mm.y, x[v]);
                }
            }
        }
    }
    KM_TR(2);
    // ---- labels, population deltas, per-block change flags ----
    {
        int newl[V], oldl[V];
        bool chg[V], anyc = false;
        unsigned packed = 0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            float bs, sj[2];
            int best = 0;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(sj[0]), "=f"(sj[1]) : "l"(s2[v][0]));
            bs = sj[0];
            if (sj[1] < bs) { bs = sj[1]; best = 1; }
#pragma unroll
            for (int i = 1; i < K / 2; ++i) {
                asm("mov.b64 {%0, %1}, %2;" : "=f"(sj[0]), "=f"(sj[1]) : "l"(s2[v][i]));
                if (sj[0] < bs) { bs = sj[0]; best = 2 * i; }
                if (sj[1] < bs) { bs = sj[1]; best = 2 * i + 1; }
            }
            const bool valid = p0 + v < N;
            newl[v] = best;
            oldl[v] = P.first ? KM_NONE : (int)((prev >> (8 * v)) & 0xffu);
            chg[v] = valid && best != oldl[v];
            anyc |= chg[v];
            packed |= (unsigned)best << (8 * v);
            if (valid && P.labels_out) P.labels_out[(size_t)b * N + p0 + v] = best;
            s_new[threadIdx.x * V + v] = chg[v] ? (unsigned char)best : (unsigned char)KM_NONE;
            s_old[threadIdx.x * V + v] = chg[v] ? (unsigned char)oldl[v] : (unsigned char)KM_NONE;
        }
        if (anyc && p0 + V <= P.lab_stride) {
            if constexpr (V == 4) *reinterpret_cast<unsigned *>(lab + p0) = packed;
            else if constexpr (V == 2) *reinterpret_cast<unsigned short *>(lab + p0) = (unsigned short)packed;
            else lab[p0] = (unsigned char)packed;
        }
        const unsigned any = __ballot_sync(0xffffffffu, anyc);
        if (any) {
            // list the quads that hold a changed pixel (4 / V lanes per quad; order is irrelevant: the sums are exact)
            constexpr int LQ = V >= 4 ? 1 : 4 / V;
            bool qf = anyc;
#pragma unroll
            for (int o = 1; o < LQ; o <<= 1) qf |= __shfl_xor_sync(0xffffffffu, qf, o) != 0;
            const unsigned qm = __ballot_sync(0xffffffffu, qf && (lane % LQ) == 0);
            int qbase = 0;
            if (lane == 0) qbase = atomicAdd(&s_nq, __popc(qm));
            qbase = __shfl_sync(0xffffffffu, qbase, 0);
            if (qf && (lane % LQ) == 0)
                s_quads[qbase + __popc(qm & ((1u << lane) - 1u))] = (unsigned char)((threadIdx.x * V) >> 2);
            if constexpr (K <= 8) {   // ... and the first KT_SPARSE changed pixels themselves
                unsigned pm[V];
                int pc = 0;
#pragma unroll
                for (int v = 0; v < V; ++v) { pm[v] = __ballot_sync(0xffffffffu, chg[v]); pc += __popc(pm[v]); }
                int pbase = 0;
                if (lane == 0) pbase = atomicAdd(&s_npx, pc);
                pbase = __shfl_sync(0xffffffffu, pbase, 0);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int pos = pbase + __popc(pm[v] & ((1u << lane) - 1u));
                    if (chg[v] && pos < KT_SPARSE) s_plist[pos] = (unsigned char)(threadIdx.x * V + v);
                    pbase += __popc(pm[v]);
                }
            }
        }
    }
    __syncthreads();
    KM_TR(3);

    // ---- centroid-sum deltas from the resident tile (tensor cores) ----
    int nq = s_nq;
#ifdef KM_SKIP_B   // timing experiment only: wrong results
    nq = 0;
#endif
    bool sparse = false;
    if constexpr (K <= 8) sparse = nq && s_npx <= KT_SPARSE;
    if constexpr (K <= 8) if (sparse) {
        // A handful of changed pixels (the common case after the first passes): thread = feature, the deltas
        // of a feature stay in registers and leave as at most K global atomics; no conversion to digits, no MMA.
        // (All lanes of a warp read one pixel of 32 different planes: a 32-way bank conflict, on <= 12 loads.)
        const int npx = s_npx;
        for (int d = threadIdx.x; d < D; d += THREADS) {
            long long acc[K];
#pragma unroll
            for (int j = 0; j < K; ++j) acc[j] = 0;
            const float *xd = s_x + (size_t)d * TP;
            for (int e = 0; e < npx; ++e) {
                const int px = s_plist[e];
                km_move<K>(acc, nullptr, s_new[px], s_old[px], (long long)__float2int_rn(xd[px] * P.fix_scale));
            }
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(P.sums + (size_t)b * k * D + d);
#pragma unroll
            for (int j = 0; j < K; ++j)
                if (j < k && acc[j]) atomicAdd(dst + (size_t)j * D, (unsigned long long)acc[j]);
        }
        if (threadIdx.x >= THREADS - K) {   // population deltas: one (otherwise idle) thread per cluster
            const int j = threadIdx.x - (THREADS - K);
            int dcnt = 0;
            for (int e = 0; e < npx; ++e) {
                const int px = s_plist[e];
                dcnt += (s_new[px] == j) - (s_old[px] == j);
            }
            if (j < k && dcnt) atomicAdd(P.counts + b * k + j, dcnt);
        }
    }
    if (nq && !sparse) {
        // A warp works on FGU feature groups (of 8 planes) at once: their load -> convert -> MMA chains are
        // independent, which hides the chain latency when a tile has only a few changed quads.
        constexpr int FGU = 3;
        const int g = lane >> 2, kq = lane & 3;
        const int n_fg = (D + 7) / 8;
        const unsigned *newp = reinterpret_cast<const unsigned *>(s_new);
        const unsigned *oldp = reinterpret_cast<const unsigned *>(s_old);
        for (int fgb = warp; fgb < n_fg; fgb += WARPS * FGU) {
            int c[FGU][MT][4][4], cc[MT][4];   // cc: A x ones = the clusters' population deltas (every column alike)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int i = 0; i < 4; ++i) cc[mt][i] = 0;
            const float *row[FGU];
            bool d_ok[FGU];
#pragma unroll
            for (int u = 0; u < FGU; ++u) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int dg = 0; dg < 4; ++dg)
#pragma unroll
                        for (int i = 0; i < 4; ++i) c[u][mt][dg][i] = 0;
                const int d = (fgb + u * WARPS) * 8 + g;
                d_ok[u] = d < D;
                row[u] = s_x + (size_t)(d_ok[u] ? d : 0) * TP;
            }
            // eight listed quads (32 pixels) per MMA k-block: thread (g, kq) converts quads kq and 4 + kq of the block
#pragma unroll 2
            for (int q0 = 0; q0 < nq; q0 += 8) {
                int qi[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int e = q0 + h * 4 + kq;
                    qi[h] = e < nq ? (int)s_quads[e] : -1;
                }
                unsigned bw[FGU][2][4];
#pragma unroll
                for (int u = 0; u < FGU; ++u)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 x = *reinterpret_cast<const float4 *>(row[u] + max(qi[h], 0) * 4);
                        unsigned w0 = (unsigned)__float2int_rn(x.x * P.fix_scale), w1 = (unsigned)__float2int_rn(x.y * P.fix_scale);
                        unsigned w2 = (unsigned)__float2int_rn(x.z * P.fix_scale), w3 = (unsigned)__float2int_rn(x.w * P.fix_scale);
                        if (!d_ok[u]) w0 = w1 = w2 = w3 = 0x80000000u;   // digit 0 after the offset
                        const unsigned t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w2, w3, 0x5140);
                        const unsigned t2 = __byte_perm(w0, w1, 0x7362), t3 = __byte_perm(w2, w3, 0x7362);
                        bw[u][h][0] = __byte_perm(t0, t1, 0x5410);
                        bw[u][h][1] = __byte_perm(t0, t1, 0x7632);
                        bw[u][h][2] = __byte_perm(t2, t3, 0x5410);
                        bw[u][h][3] = __byte_perm(t2, t3, 0x7632) ^ 0x80808080u;   // + 2^31
                    }
                const unsigned n0 = qi[0] >= 0 ? newp[qi[0]] : 0xffffffffu, n1 = qi[1] >= 0 ? newp[qi[1]] : 0xffffffffu;
                const unsigned o0 = qi[0] >= 0 ? oldp[qi[0]] : 0xffffffffu, o1 = qi[1] >= 0 ? oldp[qi[1]] : 0xffffffffu;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const unsigned r0w = (unsigned)(mt * 16 + g) * 0x01010101u, r1w = r0w + 0x08080808u;
                    const unsigned a0 = (__vcmpeq4(n0, r0w) & 0x01010101u) | __vcmpeq4(o0, r0w);
                    const unsigned a2 = (__vcmpeq4(n1, r0w) & 0x01010101u) | __vcmpeq4(o1, r0w);
                    unsigned a1 = 0, a3 = 0;
                    if (K > 8) {
                        a1 = (__vcmpeq4(n0, r1w) & 0x01010101u) | __vcmpeq4(o0, r1w);
                        a3 = (__vcmpeq4(n1, r1w) & 0x01010101u) | __vcmpeq4(o1, r1w);
                    }
                    asm(
                        "mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                        : "+r"(cc[mt][0]), "+r"(cc[mt][1]), "+r"(cc[mt][2]), "+r"(cc[mt][3])
                        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(0x01010101u), "r"(0x01010101u));
#pragma unroll
                    for (int u = 0; u < FGU; ++u)
#pragma unroll
                        for (int dg = 0; dg < 4; ++dg)
                            asm(
                                "mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                : "+r"(c[u][mt][dg][0]), "+r"(c[u][mt][dg][1]), "+r"(c[u][mt][dg][2]), "+r"(c[u][mt][dg][3])
                                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bw[u][0][dg]), "r"(bw[u][1][dg]));
                }
            }
            if (fgb == 0 && kq == 0) {   // warp 0's first round also publishes the population deltas
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int hrow = 0; hrow < (K > 8 ? 2 : 1); ++hrow) {
                        const int j = mt * 16 + hrow * 8 + g;
                        if (j < k && cc[mt][2 * hrow]) atomicAdd(P.counts + b * k + j, cc[mt][2 * hrow]);
                    }
            }
            // digits -> int64, remove the offset, publish: one global atomic per (cluster, feature) per tile
#pragma unroll
            for (int u = 0; u < FGU; ++u)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int hrow = 0; hrow < (K > 8 ? 2 : 1); ++hrow)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int j = mt * 16 + hrow * 8 + g;
                            const int dd = (fgb + u * WARPS) * 8 + 2 * kq + i;
                            long long v = (long long)c[u][mt][0][2 * hrow + i] + ((long long)c[u][mt][1][2 * hrow + i] << 8) +
                                          ((long long)c[u][mt][2][2 * hrow + i] << 16) + ((long long)c[u][mt][3][2 * hrow + i] << 24);
                            if (j < k && dd < D) {
                                v -= (long long)cc[mt][2 * hrow] << 31;
                                if (v) atomicAdd(reinterpret_cast<unsigned long long *>(P.sums + ((size_t)b * k + j) * D + dd),
                                                 (unsigned long long)v);
                            }
                        }
        }
    }
    KM_TR(4);
}

// sums -> next centroids and next score table of every image, after a tile-resident pass
// (same arithmetic as km_ticket_finalize; one CTA per image)
__global__ void km_finalize_kernel(const __grid_constant__ KmParams P, int K)
{
    extern __shared__ __align__(128) unsigned char km_smem[];
    float *s_c = reinterpret_cast<float *>(km_smem);   // [k][D]
    const int b = blockIdx.x, D = P.D, k = P.k;
    griddep_launch_dependents();   // the next pass may start loading its first tiles
    griddep_wait();                // every atomic of the pass has landed
    float *cent = P.cent + (size_t)b * k * D;
    const long long *sums = P.sums + (size_t)b * k * D;
    const float *affine = P.affine ? P.affine + (size_t)b * D * 2 : nullptr;
    for (int i = threadIdx.x; i < k * D; i += blockDim.x) {
        const int cnt = P.counts[b * k + i / D];
        float c;
        if (cnt > 0) c = km_affine(affine, i % D, __double2float_rn(__ddiv_rn((double)sums[i], __dmul_rn((double)cnt, (double)P.fix_scale))));
        else c = cent[i];
        cent[i] = c;
        s_c[i] = c;
    }
    __syncthreads();
    km_write_prep(s_c, P.prep + (size_t)b * (D * K + K), D, k, K, threadIdx.x, blockDim.x, affine);
}

template <int K, int TP, int V>
int launch_tile(const KmParams &P, const CUtensorMap &tmap, int B, cudaStream_t st)
{
    static_assert(sizeof(KtState<K, TP>) == ((KT_GROUPS * 8 + 2 * TP + TP / 4 + KT_SPARSE + 8 + 15) & ~15), "kt_smem_bytes out of sync");
    const size_t smem = kt_smem_bytes(K, TP, P.D);
    static SmemAttrCache attr_cache;
    size_t &attr_smem = attr_cache.cur();
    if (smem > attr_smem) {
        GCIS_CUDA_TRY((cudaFuncSetAttribute(km_tile_kernel<K, TP, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        attr_smem = smem;
        if (getenv("GCIS_DEBUG")) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, km_tile_kernel<K, TP, V>, TP / V, smem);
            fprintf(stderr, "[gcis] km_tile_kernel<%d,%d,%d>: %zu bytes of shared memory, %d CTAs per SM\n", K, TP, V, smem, nb);
        }
    }
    cudaLaunchAttribute pdl;
    pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl.val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P.chunks, B); cfg.blockDim = dim3(TP / V); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = &pdl; cfg.numAttrs = 1;
    GCIS_CUDA_TRY(cudaLaunchKernelEx(&cfg, km_tile_kernel<K, TP, V>, P, tmap));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cfg.gridDim = dim3(B); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = sizeof(float) * P.k * P.D;
    GCIS_CUDA_TRY(cudaLaunchKernelEx(&cfg, km_finalize_kernel, P, K));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GCIS_OK;
}

// feature tensor as the TMA engine sees it: (pixel, plane, image), boxes of TP pixels x ppg planes
static int kt_make_tensor_map(CUtensorMap *map, const float *d_feat, size_t img_stride, int plane_stride, int B, int D, int tp)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        GCIS_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) return set_error(GCIS_E_CUDA, "kmeans: cuTensorMapEncodeTiled not available");
        encode = reinterpret_cast<encode_fn>(fn);
    }
    const int ppg = (D + KT_GROUPS - 1) / KT_GROUPS;
    const cuuint64_t dims[3] = {(cuuint64_t)plane_stride, (cuuint64_t)D, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)plane_stride * 4u, (cuuint64_t)img_stride * 4u};
    const cuuint32_t box[3] = {(cuuint32_t)tp, (cuuint32_t)ppg, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(d_feat), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(GCIS_E_CUDA, "kmeans: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return GCIS_OK;
}

// tile size of the tile-resident pass (pixels): the largest of {256, 128} of which two CTAs fit in
// one SM's shared memory; 0 = none (streaming-ring pass over planar features)
static int kt_tile_pixels(int K, int D)
{
    static const int force = [] { const char *e = getenv("GCIS_KM_TP"); return e ? atoi(e) : 0; }();
    static const bool force_ring = [] { const char *e = getenv("GCIS_KM_RING"); return e && atoi(e) != 0; }();
    if (force_ring) return 0;
    // k = 16: 128-pixel tiles measured 0.64 of the HBM peak against 0.59 with 256 (more CTAs per SM hide the longer score
    // loop); k = 8 and k = 32 are faster with 256 (k = 32: 0.54 against 0.51 of the FP32 bound)
    if (force == 0 && K == 16 && kt_smem_bytes(K, 128, D) <= KT_SMEM_BUDGET) return 128;
    if (force != 128 && kt_smem_bytes(K, 256, D) <= KT_SMEM_BUDGET) return 256;
    if (kt_smem_bytes(K, 128, D) <= KT_SMEM_BUDGET) return 128;
    return 0;
}

template <int K>
int launch_tile_tp(const KmParams &P, const CUtensorMap &tmap, int B, int tp, cudaStream_t st)
{
    // two pixels per thread: measured best on B200 (1: 36.9, 2: 34.4, 4: 39.0 ms per 200 images)
    return tp == 256 ? launch_tile<K, 256, KT_V>(P, tmap, B, st) : launch_tile<K, 128, (KT_V > 2 ? 2 : KT_V)>(P, tmap, B, st);
}

template <int K, int VEC>
int launch_pass(const KmParams &P, int B, cudaStream_t st)
{
    const size_t smem = km_smem_bytes(K, VEC, P.D);
    if (smem > 200 * 1024) return set_error(GCIS_E_INVALID, "kmeans: D=%d too large for shared memory", P.D);
    static SmemAttrCache attr_cache;
    size_t &attr_smem = attr_cache.cur();
    if (smem > attr_smem) {
        GCIS_CUDA_TRY(cudaFuncSetAttribute(km_pass_kernel<K, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    km_pass_kernel<K, VEC><<<dim3(P.chunks, B), KM_THREADS, smem, st>>>(P);
    GCIS_LAUNCH_CHECK();
    return GCIS_OK;
}

template <int VEC>
int launch_pass_k(const KmParams &P, int B, cudaStream_t st)
{
    if (P.k <= 8) return launch_pass<8, VEC>(P, B, st);
    if (P.k <= 16) return launch_pass<16, VEC>(P, B, st);
    return launch_pass<32, VEC>(P, B, st);
}

}  // namespace

static size_t km_lab_stride(int N) { return (size_t)round_up(N, 16); }
static int km_pad_k(int k) { return k <= 8 ? 8 : (k <= 16 ? 16 : 32); }

size_t kmeans_workspace_bytes(int B, int D, int N, int k)
{
    const int K = km_pad_k(k);
    return sizeof(float) * (size_t)B * k * D + sizeof(float) * (size_t)B * (D * K + K) +
           sizeof(long long) * (size_t)B * k * D + sizeof(int) * (size_t)B * (k + 1) + (size_t)B * km_lab_stride(N) + 512;
}

// d_feat: [B] images of D planes, `plane_stride` floats apart, images `img_stride` floats apart.
// d_ws: workspace of kmeans_workspace_bytes(B, D, N, k), 16-byte aligned.
int kmeans_launch(const float *d_feat, size_t img_stride, int plane_stride, int B, int D, int N, int k, int iters,
                  int fix_shift, const int32_t *d_init_idx, int32_t *d_labels, float *d_centroids, void *d_ws,
                  cudaStream_t st, const float *d_affine)
{
    if (k < 1 || k > 32) return set_error(GCIS_E_INVALID, "kmeans: k=%d outside 1..32", k);
    if (iters < 1) return set_error(GCIS_E_INVALID, "kmeans: iters=%d < 1", iters);
    if (fix_shift < 0 || fix_shift > 30) return set_error(GCIS_E_INVALID, "kmeans: fix_shift=%d outside 0..30", fix_shift);
    if (B > 65535) return set_error(GCIS_E_INVALID, "kmeans: B=%d > 65535 per call", B);
    const int K = km_pad_k(k);
    auto align16 = [](char *p) { return reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15); };
    char *w = static_cast<char *>(d_ws);
    long long *sums = reinterpret_cast<long long *>(w); w += sizeof(long long) * (size_t)B * k * D;
    w = align16(w);
    float *prep = reinterpret_cast<float *>(w); w += sizeof(float) * (size_t)B * (D * K + K);
    float *cent = reinterpret_cast<float *>(w); w += sizeof(float) * (size_t)B * k * D;
    int *counts = reinterpret_cast<int *>(w); w += sizeof(int) * (size_t)B * k;
    int *done = reinterpret_cast<int *>(w); w += sizeof(int) * (size_t)B;
    unsigned char *lab8 = reinterpret_cast<unsigned char *>(align16(w));

    km_init_kernel<<<B, 256, 0, st>>>(d_feat, img_stride, plane_stride, d_init_idx, cent, prep, sums, counts, done, D, N, k, K, d_affine);
    GCIS_LAUNCH_CHECK();
    // 128-bit loads need 16-byte aligned, padded planes; otherwise one pixel per thread
    const bool vec4 = plane_stride % 4 == 0 && plane_stride >= round_up(N, 4) && img_stride % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(d_feat) & 15) == 0;
    KmParams P;
    P.feat = d_feat; P.img_stride = img_stride; P.plane_stride = plane_stride;
    P.cent = cent; P.prep = prep; P.sums = sums; P.counts = counts; P.done = done;
    P.lab8 = lab8; P.lab_stride = (int)km_lab_stride(N);
    P.affine = d_affine;
    // tile-resident pass when a tile fits twice per SM (GCIS_KM_RING=1 forces the streaming-ring pass)
    const int tp = vec4 ? kt_tile_pixels(K, D) : 0;
    P.D = D; P.N = N; P.k = k; P.chunks = tp ? ceil_div(N, tp) : ceil_div(N, KM_THREADS * (vec4 ? 4 : 1));
    P.fix_scale = (float)(1u << fix_shift);
    alignas(64) CUtensorMap tmap;
    if (tp) {
        const int rc = kt_make_tensor_map(&tmap, d_feat, img_stride, plane_stride, B, D, tp);
        if (rc) return rc;
    }
    for (int t = 0; t < iters; ++t) {
        P.labels_out = (t == iters - 1) ? d_labels : nullptr;
        P.first = t == 0;
        P.pass = t;
        int rc;
        if (tp) rc = K == 8 ? launch_tile_tp<8>(P, tmap, B, tp, st) : (K == 16 ? launch_tile_tp<16>(P, tmap, B, tp, st) : launch_tile_tp<32>(P, tmap, B, tp, st));
        else rc = vec4 ? launch_pass_k<4>(P, B, st) : launch_pass_k<1>(P, B, st);
        if (rc) return rc;
    }
    if (d_centroids)
        GCIS_CUDA_TRY(cudaMemcpyAsync(d_centroids, cent, sizeof(float) * (size_t)B * k * D, cudaMemcpyDeviceToDevice, st));
    return GCIS_OK;
}

}  // namespace gcis

#ifdef KM_TRACE
extern "C" __attribute__((visibility("default"))) int gcis_km_trace_read(long long *out, size_t bytes)
{
    return (int)cudaMemcpyFromSymbol(out, gcis::km_trace_buf, bytes);
}
#endif
