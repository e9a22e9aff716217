// jpeg.cu — image decode on the path: baseline JPEG -> interleaved RGB in HBM.
//
// Reference: BSD_metrics/script.py:25, `img = imread(img_path + name)` (scikit-image -> PIL -> libjpeg).  The
// result must be the pixels that call returns, bit for bit: the BSDS500 files are baseline sequential, 8-bit,
// Huffman coded, YCbCr 4:2:0, and libjpeg's defaults for them are the integer "islow" inverse DCT, "fancy"
// (triangle-filter) chroma upsampling and the 16-bit fixed-point YCbCr -> RGB tables.  Those three stages are
// data parallel and run here as CUDA kernels; the entropy (Huffman) decoding, a serial bit stream per image,
// runs on host threads (one image per thread) and hands the quantised coefficients over in pinned memory:
//   host    markers, Huffman tables, scan  ->  int16 coefficients [component][block][64] (natural order)
//   kernel  jpeg_idct_kernel      dequantise + 8x8 inverse DCT (jidctint.c arithmetic: 13-bit constants,
//                                 two passes, descale 11 / 18, range limit)  ->  u8 component planes
//   kernel  jpeg_colour_kernel    h2v2 / h2v1 fancy upsampling (jdsample.c) + YCbCr -> RGB (jdcolor.c) ->
//                                 [B][H][W][3] u8, the layout the segmenter takes
// Supported: SOF0 / SOF1 (baseline / extended sequential, 8 bit), 1 or 3 components, luma sampling 1x1, 2x1,
// 2x2 with chroma 1x1, restart intervals.  Anything else (progressive, arithmetic, CMYK, 12 bit) is rejected
// with GCIS_E_INVALID and the caller decodes on the host as before.
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace gcis {

namespace {

// ------------------------------------------------------------------------------------------------
// host: parser + Huffman decoder
// ------------------------------------------------------------------------------------------------

const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    bool present = false;
    uint8_t bits[17] = {0}, vals[256] = {0};
    // canonical decoding tables (ITU T.81 F.2.2.3): mincode / maxcode / valptr per code length, 9-bit lookahead
    int mincode[17], maxcode[18], valptr[17];
    uint8_t look_nbits[512], look_sym[512];
    void build()
    {
        int code = 0, k = 0;
        memset(look_nbits, 0, sizeof(look_nbits));
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
                if (l <= 9) {
                    const int first = code << (9 - l);
                    for (int j = 0; j < (1 << (9 - l)); ++j) { look_nbits[first + j] = (uint8_t)l; look_sym[first + j] = vals[k]; }
                }
            }
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int blocks_w = 0, blocks_h = 0;     // allocated blocks (whole MCUs)
    int width = 0, height = 0;          // real samples (libjpeg downsampled_width / height)
    size_t coef_off = 0;                // first int16 of this component in the image's coefficient slab
};

struct JpegHeader {
    int W = 0, H = 0, ncomp = 0, hmax = 1, vmax = 1, mcus_x = 0, mcus_y = 0, restart = 0;
    Component comp[3];
    uint16_t qt[4][64];                 // natural order
    bool qt_present[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    size_t scan_off = 0;
    size_t coef_count = 0;              // int16 per image
};

struct BitReader {
    const uint8_t *p, *end;
    uint32_t acc = 0;
    int n = 0;
    bool hit_marker = false;
    inline void fill()
    {
        while (n <= 24) {
            uint32_t b = 0;
            if (p < end && !hit_marker) {
                b = *p;
                if (b == 0xFF) {
                    if (p + 1 < end && p[1] == 0) p += 2;      // stuffed zero
                    else { hit_marker = true; b = 0; }          // a marker: feed zeros from here on
                } else {
                    ++p;
                }
            }
            acc |= b << (24 - n);
            n += 8;
        }
    }
    inline int peek(int k) { if (n < k) fill(); return (int)(acc >> (32 - k)); }
    inline void skip(int k) { acc <<= k; n -= k; }
    inline int get(int k) { if (k == 0) return 0; const int v = peek(k); skip(k); return v; }
    void restart_align() { acc = 0; n = 0; hit_marker = false; }
};

inline int huff_decode(BitReader &br, const HuffTable &t)
{
    const int look = br.peek(9);
    int l = t.look_nbits[look];
    if (l) { br.skip(l); return t.look_sym[look]; }
    int code = br.peek(16);
    for (l = 10; l <= 16; ++l) {
        const int c = code >> (16 - l);
        if (c <= t.maxcode[l] && t.maxcode[l] >= 0 && c >= t.mincode[l]) {
            br.skip(l);
            return t.vals[t.valptr[l] + c - t.mincode[l]];
        }
    }
    return -1;
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

int parse_header(const uint8_t *d, size_t n, JpegHeader &hd)
{
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return set_error(GCIS_E_INVALID, "jpeg: no SOI marker");
    size_t p = 2;
    bool sof = false;
    while (p + 4 <= n) {
        if (d[p] != 0xFF) return set_error(GCIS_E_INVALID, "jpeg: marker expected at byte %zu", p);
        while (p < n && d[p] == 0xFF) ++p;                       // fill bytes
        const int m = d[p++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) break;
        if (p + 2 > n) break;
        const size_t len = ((size_t)d[p] << 8) | d[p + 1];
        if (len < 2 || p + len > n) return set_error(GCIS_E_INVALID, "jpeg: truncated segment");
        const uint8_t *s = d + p + 2;
        const size_t sl = len - 2;
        if (m == 0xDB) {                                         // DQT
            size_t q = 0;
            while (q < sl) {
                const int pq = s[q] >> 4, tq = s[q] & 15;
                ++q;
                if (tq > 3 || q + (pq ? 128 : 64) > sl) return set_error(GCIS_E_INVALID, "jpeg: bad DQT");
                for (int i = 0; i < 64; ++i) {
                    hd.qt[tq][ZIGZAG[i]] = pq ? (uint16_t)((s[q] << 8) | s[q + 1]) : s[q];
                    q += pq ? 2 : 1;
                }
                hd.qt_present[tq] = true;
            }
        } else if (m == 0xC4) {                                  // DHT
            size_t q = 0;
            while (q + 17 <= sl) {
                const int tc = s[q] >> 4, th = s[q] & 15;
                if (tc > 1 || th > 3) return set_error(GCIS_E_INVALID, "jpeg: bad DHT");
                HuffTable &t = tc ? hd.ac[th] : hd.dc[th];
                int total = 0;
                t.bits[0] = 0;
                for (int i = 1; i <= 16; ++i) { t.bits[i] = s[q + i]; total += s[q + i]; }
                q += 17;
                if (total > 256 || q + total > sl) return set_error(GCIS_E_INVALID, "jpeg: bad DHT");
                memcpy(t.vals, s + q, total);
                q += total;
                t.present = true;
                t.build();
            }
        } else if (m == 0xC0 || m == 0xC1) {                     // SOF0 / SOF1
            if (sl < 6 || s[0] != 8) return set_error(GCIS_E_INVALID, "jpeg: only 8-bit samples are supported");
            hd.H = (s[1] << 8) | s[2]; hd.W = (s[3] << 8) | s[4]; hd.ncomp = s[5];
            if ((hd.ncomp != 1 && hd.ncomp != 3) || sl < 6 + 3 * (size_t)hd.ncomp || hd.H < 1 || hd.W < 1)
                return set_error(GCIS_E_INVALID, "jpeg: %d components are not supported", hd.ncomp);
            for (int c = 0; c < hd.ncomp; ++c) {
                Component &k = hd.comp[c];
                k.id = s[6 + 3 * c]; k.h = s[7 + 3 * c] >> 4; k.v = s[7 + 3 * c] & 15; k.tq = s[8 + 3 * c];
                if (k.tq > 3) return set_error(GCIS_E_INVALID, "jpeg: bad quantisation table index");
            }
            sof = true;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC8 && m != 0xCC) {
            return set_error(GCIS_E_INVALID, "jpeg: SOF%d (progressive / lossless / arithmetic) is not supported", m - 0xC0);
        } else if (m == 0xDD) {                                  // DRI
            if (sl >= 2) hd.restart = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {                                  // SOS
            if (!sof) return set_error(GCIS_E_INVALID, "jpeg: SOS before SOF");
            if (sl < 1 || s[0] != hd.ncomp || sl < 1 + 2 * (size_t)hd.ncomp + 3)
                return set_error(GCIS_E_INVALID, "jpeg: multi-scan files are not supported");
            for (int i = 0; i < hd.ncomp; ++i) {
                const int id = s[1 + 2 * i];
                int c = -1;
                for (int j = 0; j < hd.ncomp; ++j) if (hd.comp[j].id == id) c = j;
                if (c != i) return set_error(GCIS_E_INVALID, "jpeg: unexpected component order in the scan");
                hd.comp[c].td = s[2 + 2 * i] >> 4; hd.comp[c].ta = s[2 + 2 * i] & 15;
                if (hd.comp[c].td > 3 || hd.comp[c].ta > 3 || !hd.dc[hd.comp[c].td].present || !hd.ac[hd.comp[c].ta].present ||
                    !hd.qt_present[hd.comp[c].tq])
                    return set_error(GCIS_E_INVALID, "jpeg: scan refers to a missing table");
            }
            hd.scan_off = p + len;
            break;
        }
        p += len;
    }
    if (!sof || !hd.scan_off) return set_error(GCIS_E_INVALID, "jpeg: no frame / scan found");
    if (hd.ncomp == 1) { hd.comp[0].h = hd.comp[0].v = 1; }
    hd.hmax = hd.vmax = 1;
    for (int c = 0; c < hd.ncomp; ++c) { hd.hmax = std::max(hd.hmax, hd.comp[c].h); hd.vmax = std::max(hd.vmax, hd.comp[c].v); }
    if (hd.ncomp == 3) {
        const Component &y = hd.comp[0];
        const bool ok = hd.comp[1].h == 1 && hd.comp[1].v == 1 && hd.comp[2].h == 1 && hd.comp[2].v == 1 &&
                        ((y.h == 1 && y.v == 1) || (y.h == 2 && y.v == 1) || (y.h == 2 && y.v == 2));
        if (!ok) return set_error(GCIS_E_INVALID, "jpeg: sampling %dx%d,%dx%d,%dx%d is not supported", y.h, y.v, hd.comp[1].h,
                                  hd.comp[1].v, hd.comp[2].h, hd.comp[2].v);
    }
    hd.mcus_x = ceil_div(hd.W, 8 * hd.hmax); hd.mcus_y = ceil_div(hd.H, 8 * hd.vmax);
    size_t off = 0;
    for (int c = 0; c < hd.ncomp; ++c) {
        Component &k = hd.comp[c];
        k.blocks_w = hd.mcus_x * k.h; k.blocks_h = hd.mcus_y * k.v;
        k.width = ceil_div(hd.W * k.h, hd.hmax); k.height = ceil_div(hd.H * k.v, hd.vmax);
        k.coef_off = off;
        off += (size_t)k.blocks_w * k.blocks_h * 64;
    }
    hd.coef_count = off;
    return GCIS_OK;
}

// entropy-decode the single interleaved scan into natural-order coefficients (still quantised)
int decode_scan(const uint8_t *d, size_t n, const JpegHeader &hd, int16_t *coef)
{
    memset(coef, 0, hd.coef_count * sizeof(int16_t));
    BitReader br{d + hd.scan_off, d + n};
    int pred[3] = {0, 0, 0};
    int until_restart = hd.restart;
    for (int my = 0; my < hd.mcus_y; ++my)
        for (int mx = 0; mx < hd.mcus_x; ++mx) {
            if (hd.restart && until_restart == 0) {
                // byte-align, expect RSTn
                const uint8_t *q = br.p;
                while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
                if (q + 1 >= br.end) return set_error(GCIS_E_INVALID, "jpeg: restart marker missing");
                br.p = q + 2;
                br.restart_align();
                pred[0] = pred[1] = pred[2] = 0;
                until_restart = hd.restart;
            }
            for (int c = 0; c < hd.ncomp; ++c) {
                const Component &k = hd.comp[c];
                const HuffTable &dct = hd.dc[k.td], &act = hd.ac[k.ta];
                for (int by = 0; by < k.v; ++by)
                    for (int bx = 0; bx < k.h; ++bx) {
                        int16_t *blk = coef + k.coef_off + ((size_t)(my * k.v + by) * k.blocks_w + mx * k.h + bx) * 64;
                        int s = huff_decode(br, dct);
                        if (s < 0 || s > 11) return set_error(GCIS_E_INVALID, "jpeg: corrupt DC code");
                        if (s) pred[c] += extend(br.get(s), s);
                        blk[0] = (int16_t)pred[c];
                        for (int kk = 1; kk < 64;) {
                            const int rs = huff_decode(br, act);
                            if (rs < 0) return set_error(GCIS_E_INVALID, "jpeg: corrupt AC code");
                            const int r = rs >> 4, sz = rs & 15;
                            if (sz == 0) {
                                if (r == 15) { kk += 16; continue; }
                                break;                                   // end of block
                            }
                            kk += r;
                            if (kk > 63) return set_error(GCIS_E_INVALID, "jpeg: corrupt AC run");
                            blk[ZIGZAG[kk]] = (int16_t)extend(br.get(sz), sz);
                            ++kk;
                        }
                    }
            }
            if (hd.restart) --until_restart;
        }
    return GCIS_OK;
}

// ------------------------------------------------------------------------------------------------
// device: dequantise + inverse DCT (jidctint.c, jpeg_idct_islow), upsample + colour
// ------------------------------------------------------------------------------------------------

struct JpegDevComp {
    int blocks_w, blocks_h, width, height, h, v;
    size_t coef_off;      // int16 offset inside one image's slab
    size_t plane_off;     // u8 offset inside one image's plane slab; plane stride = blocks_w * 8
};
struct JpegParams {
    const int16_t *coef;  // [B][coef_count]
    uint8_t *planes;      // [B][plane_bytes]
    uint8_t *rgb;         // [B][H][W][3]
    size_t coef_count, plane_bytes;
    int B, H, W, ncomp, hmax, vmax;
    JpegDevComp comp[3];
    uint16_t qt[3][64];
};

constexpr int CONST_BITS = 13, PASS1_BITS = 2;
constexpr int FIX_0_298631336 = 2446, FIX_0_390180644 = 3196, FIX_0_541196100 = 4433, FIX_0_765366865 = 6270,
              FIX_0_899976223 = 7373, FIX_1_175875602 = 9633, FIX_1_501321110 = 12299, FIX_1_847759065 = 15137,
              FIX_1_961570560 = 16069, FIX_2_053119869 = 16819, FIX_2_562915447 = 20995, FIX_3_072711026 = 25172;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 1-D pass of the islow IDCT on 8 values (even part / odd part exactly as jidctint.c)
__device__ __forceinline__ void idct8(const int (&in)[8], int (&out)[8], int shift, bool pass1)
{
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * FIX_0_541196100;
    const int tmp2 = z1 + z3 * (-FIX_1_847759065);
    const int tmp3 = z1 + z2 * FIX_0_765366865;
    z2 = in[0]; z3 = in[4];
    const int tmp0 = (z2 + z3) << CONST_BITS;
    const int tmp1 = (z2 - z3) << CONST_BITS;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    int t0 = in[7], t1 = in[5], t2 = in[3], t3 = in[1];
    z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2;
    int z4 = t1 + t3;
    const int z5 = (z3 + z4) * FIX_1_175875602;
    t0 *= FIX_0_298631336; t1 *= FIX_2_053119869; t2 *= FIX_3_072711026; t3 *= FIX_1_501321110;
    z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
    z3 += z5; z4 += z5;
    t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
    out[0] = descale(tmp10 + t3, shift); out[7] = descale(tmp10 - t3, shift);
    out[1] = descale(tmp11 + t2, shift); out[6] = descale(tmp11 - t2, shift);
    out[2] = descale(tmp12 + t1, shift); out[5] = descale(tmp12 - t1, shift);
    out[3] = descale(tmp13 + t0, shift); out[4] = descale(tmp13 - t0, shift);
    (void)pass1;
}

// thread = one column (pass 1) then one row (pass 2) of a block; 8 threads per block, 32 blocks per CTA
__global__ void __launch_bounds__(256) jpeg_idct_kernel(const __grid_constant__ JpegParams P)
{
    __shared__ int ws[32][8][9];
    const int c = blockIdx.y, b = blockIdx.z;
    const JpegDevComp &k = P.comp[c];
    const int nblk = k.blocks_w * k.blocks_h;
    const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
    const int blk = blockIdx.x * 32 + lb;
    const bool ok = blk < nblk;
    if (ok) {
        const int16_t *src = P.coef + (size_t)b * P.coef_count + k.coef_off + (size_t)blk * 64;
        int in[8], out[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) in[r] = (int)src[r * 8 + t] * (int)P.qt[c][r * 8 + t];   // column t, dequantised
        // (jidctint.c short-cuts an all-zero AC column to the DC term: same value as the full formula)
        idct8(in, out, CONST_BITS - PASS1_BITS, true);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[lb][r][t] = out[r];
    }
    __syncthreads();
    if (ok) {
        int in[8], out[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) in[x] = ws[lb][t][x];                                   // row t
        idct8(in, out, CONST_BITS + PASS1_BITS + 3, false);
        const int by = blk / k.blocks_w, bx = blk - by * k.blocks_w;
        uint8_t *dst = P.planes + (size_t)b * P.plane_bytes + k.plane_off + (size_t)(by * 8 + t) * (k.blocks_w * 8) + bx * 8;
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const uint32_t v = (uint32_t)min(max(out[x] + 128, 0), 255);                    // range_limit, CENTERJSAMPLE
            if (x < 4) lo |= v << (8 * x); else hi |= v << (8 * (x - 4));
        }
        *reinterpret_cast<uint2 *>(dst) = make_uint2(lo, hi);
    }
}

// chroma sample of output pixel (y, x) after libjpeg's fancy upsampling (jdsample.c); the neighbouring rows and
// columns are clamped to the component's real extent, which is what libjpeg's edge replication amounts to
__device__ __forceinline__ int fancy_sample(const uint8_t *pl, int stride, int cw, int ch, int y, int x, int h, int v)
{
    if (h == 1 && v == 1) return pl[(size_t)y * stride + x];
    if (h == 2 && v == 1) {                                     // h2v1_fancy_upsample
        const int i = x >> 1;
        const int cur = pl[(size_t)y * stride + i];
        if (x & 1) {
            if (i == cw - 1) return cur;
            return (3 * cur + pl[(size_t)y * stride + i + 1] + 2) >> 2;
        }
        if (i == 0) return cur;
        return (3 * cur + pl[(size_t)y * stride + i - 1] + 1) >> 2;
    }
    // h2v2_fancy_upsample: vertical 3:1 blend of the nearest two rows, then horizontal 3:1 blend of column sums
    const int i = x >> 1, j = y >> 1;
    const int jn = (y & 1) ? min(j + 1, ch - 1) : max(j - 1, 0);
    const uint8_t *r0 = pl + (size_t)j * stride, *r1 = pl + (size_t)jn * stride;
    const int cur = 3 * r0[i] + r1[i];
    if (x & 1) {
        if (i == cw - 1) return (4 * cur + 7) >> 4;
        return (3 * cur + 3 * r0[i + 1] + r1[i + 1] + 7) >> 4;
    }
    if (i == 0) return (4 * cur + 8) >> 4;
    return (3 * cur + 3 * r0[i - 1] + r1[i - 1] + 8) >> 4;
}

__global__ void __launch_bounds__(256) jpeg_colour_kernel(const __grid_constant__ JpegParams P)
{
    const int b = blockIdx.y;
    const size_t npx = (size_t)P.H * P.W;
    const uint8_t *base = P.planes + (size_t)b * P.plane_bytes;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / P.W), x = (int)(i - (size_t)y * P.W);
        const JpegDevComp &k0 = P.comp[0];
        const int Y = base[k0.plane_off + (size_t)y * (k0.blocks_w * 8) + x];
        uint8_t *dst = P.rgb + ((size_t)b * npx + i) * 3;
        if (P.ncomp == 1) { dst[0] = dst[1] = dst[2] = (uint8_t)Y; continue; }
        const int hs = P.hmax, vs = P.vmax;
        const JpegDevComp &k1 = P.comp[1], &k2 = P.comp[2];
        const int cb = fancy_sample(base + k1.plane_off, k1.blocks_w * 8, k1.width, k1.height, y, x, hs, vs) - 128;
        const int cr = fancy_sample(base + k2.plane_off, k2.blocks_w * 8, k2.width, k2.height, y, x, hs, vs) - 128;
        // jdcolor.c build_ycc_rgb_table: SCALEBITS = 16, ONE_HALF = 1 << 15, FIX(x) = (int)(x * 65536 + 0.5)
        const int r = Y + ((91881 * cr + 32768) >> 16);
        const int g = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
        const int bl = Y + ((116130 * cb + 32768) >> 16);
        dst[0] = (uint8_t)min(max(r, 0), 255);
        dst[1] = (uint8_t)min(max(g, 0), 255);
        dst[2] = (uint8_t)min(max(bl, 0), 255);
    }
}

}  // namespace

}  // namespace gcis

using namespace gcis;

extern "C" {

// Frame size of a JPEG file (host only).  Returns 0 and fills h, w, components, or a negative error.
int32_t gcis_jpeg_info(const uint8_t *data, int64_t size, int32_t *h, int32_t *w, int32_t *components)
{
    if (!data || size < 4) return set_error(GCIS_E_INVALID, "jpeg: null / empty input");
    JpegHeader hd;
    const int rc = parse_header(data, (size_t)size, hd);
    if (rc) return rc;
    if (h) *h = hd.H;
    if (w) *w = hd.W;
    if (components) *components = hd.ncomp;
    return GCIS_OK;
}

// Quantised coefficients of one file (host only; test / bring-up entry point): coef must hold `cap` int16.
// Returns the number of int16 written, or a negative error.  Layout: per component [blocks_h][blocks_w][64].
int64_t gcis_jpeg_coefficients(const uint8_t *data, int64_t size, int16_t *coef, int64_t cap)
{
    if (!data || !coef) return set_error(GCIS_E_INVALID, "jpeg: null argument");
    JpegHeader hd;
    int rc = parse_header(data, (size_t)size, hd);
    if (rc) return rc;
    if ((int64_t)hd.coef_count > cap) return set_error(GCIS_E_INVALID, "jpeg: %zu coefficients, capacity %lld", hd.coef_count, (long long)cap);
    rc = decode_scan(data, (size_t)size, hd, coef);
    return rc ? rc : (int64_t)hd.coef_count;
}

// Decode B JPEG files of one common frame size straight into the device layout the segmenter takes.
//   files[i] / sizes[i]   the file bytes in host memory
//   d_rgb                 [B][H][W][3] uint8 device buffer (H, W = the files' common frame size)
// Host threads entropy-decode (n_threads <= 0: one per hardware thread, at most B); the coefficients travel
// through a pinned staging buffer; dequantisation, inverse DCT, chroma upsampling and colour conversion run on
// `stream`.  The call returns after the stream has finished (the staging buffers are freed).
int32_t gcis_jpeg_decode_batch(const uint8_t *const *files, const int64_t *sizes, int32_t B, int32_t H, int32_t W,
                               uint8_t *d_rgb, int32_t n_threads, void *stream)
{
    if (!files || !sizes || !d_rgb || B < 1) return set_error(GCIS_E_INVALID, "jpeg: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<JpegHeader> hd(B);
    for (int i = 0; i < B; ++i) {
        const int rc = parse_header(files[i], (size_t)sizes[i], hd[i]);
        if (rc) return rc;
        if (hd[i].H != H || hd[i].W != W) return set_error(GCIS_E_INVALID, "jpeg: file %d is %dx%d, expected %dx%d", i, hd[i].H, hd[i].W, H, W);
        if (hd[i].ncomp != hd[0].ncomp || hd[i].hmax != hd[0].hmax || hd[i].vmax != hd[0].vmax)
            return set_error(GCIS_E_INVALID, "jpeg: file %d has another component layout than file 0", i);
    }
    const JpegHeader &h0 = hd[0];
    JpegParams P;
    memset(&P, 0, sizeof(P));
    P.B = 1; P.H = H; P.W = W; P.ncomp = h0.ncomp; P.hmax = h0.hmax; P.vmax = h0.vmax; P.coef_count = h0.coef_count;
    size_t poff = 0;
    for (int c = 0; c < h0.ncomp; ++c) {
        const Component &k = h0.comp[c];
        P.comp[c] = JpegDevComp{k.blocks_w, k.blocks_h, k.width, k.height, k.h, k.v, k.coef_off, poff};
        poff += (size_t)k.blocks_w * 8 * k.blocks_h * 8;
    }
    P.plane_bytes = (poff + 15) & ~(size_t)15;
    int16_t *h_coef = nullptr, *d_coef = nullptr;
    uint8_t *d_planes = nullptr;
    int rc = GCIS_OK;
    auto cleanup = [&]() { if (h_coef) cudaFreeHost(h_coef); cudaFree(d_coef); cudaFree(d_planes); };
    const size_t coef_bytes = (size_t)B * h0.coef_count * sizeof(int16_t);
    if (cudaMallocHost(reinterpret_cast<void **>(&h_coef), coef_bytes) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&d_coef), coef_bytes) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&d_planes), (size_t)B * P.plane_bytes) != cudaSuccess) {
        cudaGetLastError();
        cleanup();
        return set_error(GCIS_E_NOMEM, "jpeg: staging allocation failed");
    }
    // ---- host: entropy decoding, one image per thread at a time ----
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, (int)B));
    std::atomic<int> next{0}, failed{-1};
    std::vector<std::string> errs(nt);
    auto work = [&](int tid) {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= B) break;
            if (decode_scan(files[i], (size_t)sizes[i], hd[i], h_coef + (size_t)i * h0.coef_count)) {
                errs[tid] = g_last_error;       // the error text is thread local: carry it to the caller's thread
                int expect = -1;
                failed.compare_exchange_strong(expect, tid);
            }
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t) pool.emplace_back(work, t);
        for (auto &t : pool) t.join();
    }
    if (failed.load() >= 0) { cleanup(); return set_error(GCIS_E_INVALID, "%s", errs[failed.load()].c_str()); }
    // ---- device ----
    if (cudaMemcpyAsync(d_coef, h_coef, coef_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = set_error(GCIS_E_CUDA, "jpeg: H2D copy failed");
    for (int i = 0; i < B && !rc; ++i) {
        // the quantisation tables are per file: one launch pair per image (a few microseconds each)
        JpegParams Q = P;
        Q.coef = d_coef + (size_t)i * h0.coef_count;
        Q.planes = d_planes + (size_t)i * P.plane_bytes;
        Q.rgb = d_rgb + (size_t)i * H * W * 3;
        for (int c = 0; c < h0.ncomp; ++c) memcpy(Q.qt[c], hd[i].qt[hd[i].comp[c].tq], sizeof(Q.qt[c]));
        int maxblk = 0;
        for (int c = 0; c < h0.ncomp; ++c) maxblk = std::max(maxblk, Q.comp[c].blocks_w * Q.comp[c].blocks_h);
        jpeg_idct_kernel<<<dim3(ceil_div(maxblk, 32), h0.ncomp, 1), 256, 0, st>>>(Q);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        jpeg_colour_kernel<<<dim3(std::min(ceil_div(H * W, 256), 2048), 1), 256, 0, st>>>(Q);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaGetLastError() != cudaSuccess) rc = set_error(GCIS_E_CUDA, "jpeg: kernel launch failed");
    }
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = set_error(GCIS_E_CUDA, "jpeg: %s", cudaGetErrorString(cudaGetLastError()));
    cleanup();
    return rc;
}

}  // extern "C"
