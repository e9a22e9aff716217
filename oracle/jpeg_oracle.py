"""CPU ORACLE for the image-decode step - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of what libjpeg (the decoder behind the reference's ``imread``, BSD_metrics/script.py:25) does
after entropy decoding, with its default settings: dequantisation, the integer "islow" inverse DCT
(jidctint.c), "fancy" chroma upsampling (jdsample.c h2v2 / h2v1) and YCbCr -> RGB (jdcolor.c).  The pin is PIL
itself: tests compare this restatement, and the CUDA path, with ``PIL.Image.open`` pixel for pixel (PIL is in the
image on the build container and on the GPU box; it links libjpeg-turbo, whose SIMD paths are bit-exact with the C
code restated here).  Entropy decoding is not restated: the quantised coefficients come from the product's host
decoder (``decode.jpeg_coefficients``), which the PIL comparison validates end to end."""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                   21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53,
                   60, 61, 54, 47, 55, 62, 63])


def parse_frame(data: bytes):
    """(H, W, [(h, v, tq)], {tq: 8x8 table}) from the DQT / SOF segments."""
    p, qts, frame = 2, {}, None
    while p + 4 <= len(data):
        assert data[p] == 0xFF
        m = data[p + 1]
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            p += 2
            continue
        ln = (data[p + 2] << 8) | data[p + 3]
        s = data[p + 4:p + 2 + ln]
        if m == 0xDB:
            q = 0
            while q < len(s):
                pq, tq = s[q] >> 4, s[q] & 15
                q += 1
                if pq:
                    vals = [(s[q + 2 * i] << 8) | s[q + 2 * i + 1] for i in range(64)]
                    q += 128
                else:
                    vals = list(s[q:q + 64])
                    q += 64
                t = np.zeros(64, np.int64)
                t[ZIGZAG] = vals
                qts[tq] = t.reshape(8, 8)
        elif m in (0xC0, 0xC1):
            H, W, n = (s[1] << 8) | s[2], (s[3] << 8) | s[4], s[5]
            frame = (H, W, [(s[7 + 3 * c] >> 4, s[7 + 3 * c] & 15, s[8 + 3 * c]) for c in range(n)])
        elif m == 0xDA:
            break
        p += 2 + ln
    return frame[0], frame[1], frame[2], qts


C = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137,
         f1961=16069, f2053=16819, f2562=20995, f3072=25172)


def _idct_pass(x, shift):
    """jidctint.c 1-D pass along axis -2 of x [..., 8, n] (int64), descaled by `shift`."""
    i = [x[..., k, :] for k in range(8)]
    z1 = (i[2] + i[6]) * C["f0541"]
    tmp2 = z1 + i[6] * (-C["f1847"])
    tmp3 = z1 + i[2] * C["f0765"]
    tmp0 = (i[0] + i[4]) << 13
    tmp1 = (i[0] - i[4]) << 13
    t10, t13, t11, t12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = i[7], i[5], i[3], i[1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * C["f1175"]
    t0, t1, t2, t3 = t0 * C["f0298"], t1 * C["f2053"], t2 * C["f3072"], t3 * C["f1501"]
    z1, z2, z3, z4 = z1 * -C["f0899"], z2 * -C["f2562"], z3 * -C["f1961"] + z5, z4 * -C["f0390"] + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    out = [t10 + t3, t11 + t2, t12 + t1, t13 + t0, t13 - t0, t12 - t1, t11 - t2, t10 - t3]
    half = 1 << (shift - 1)
    return np.stack([(o + half) >> shift for o in out], axis=-2)


def idct_blocks(coef, qt):
    """coef [nb, 8, 8] quantised -> samples [nb, 8, 8] uint8 (jpeg_idct_islow)."""
    x = coef.astype(np.int64) * qt
    ws = _idct_pass(x, 13 - 2)                                   # pass 1: columns
    out = _idct_pass(ws.swapaxes(-1, -2), 13 + 2 + 3)            # pass 2: rows
    return np.clip(out.swapaxes(-1, -2) + 128, 0, 255).astype(np.uint8)


def fancy_upsample(pl, h, v):
    """jdsample.c fancy upsampling of a component plane [ch, cw] (real extent) by (h, v) in {(1,1),(2,1),(2,2)}."""
    a = pl.astype(np.int64)
    if (h, v) == (1, 1):
        return a
    if v == 2:
        up, dn = np.vstack([a[:1], a[:-1]]), np.vstack([a[1:], a[-1:]])
        s = np.empty((2 * a.shape[0], a.shape[1]), np.int64)
        s[0::2], s[1::2] = 3 * a + up, 3 * a + dn               # column sums of the nearest / next-nearest row
        left, right = np.hstack([s[:, :1], s[:, :-1]]), np.hstack([s[:, 1:], s[:, -1:]])
        out = np.empty((s.shape[0], 2 * s.shape[1]), np.int64)
        out[:, 0::2], out[:, 1::2] = (3 * s + left + 8) >> 4, (3 * s + right + 7) >> 4
        out[:, 0], out[:, -1] = (4 * s[:, 0] + 8) >> 4, (4 * s[:, -1] + 7) >> 4
        return out
    left, right = np.hstack([a[:, :1], a[:, :-1]]), np.hstack([a[:, 1:], a[:, -1:]])
    out = np.empty((a.shape[0], 2 * a.shape[1]), np.int64)
    out[:, 0::2], out[:, 1::2] = (3 * a + left + 1) >> 2, (3 * a + right + 2) >> 2
    out[:, 0], out[:, -1] = a[:, 0], a[:, -1]
    return out


def decode_from_coefficients(data: bytes, coef: np.ndarray) -> np.ndarray:
    """Everything after entropy decoding: [H, W, 3] uint8."""
    H, W, comps, qts = parse_frame(data)
    hmax, vmax = max(c[0] for c in comps), max(c[1] for c in comps)
    mx, my = -(-W // (8 * hmax)), -(-H // (8 * vmax))
    planes, off = [], 0
    for (h, v, tq) in comps:
        bw, bh = mx * h, my * v
        n = bw * bh
        blk = idct_blocks(coef[off:off + n * 64].reshape(n, 8, 8), qts[tq])
        off += n * 64
        full = blk.reshape(bh, bw, 8, 8).transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)
        cw, ch = -(-W * h // hmax), -(-H * v // vmax)
        planes.append(full[:ch, :cw])
    if len(comps) == 1:
        return np.repeat(planes[0][:H, :W, None], 3, 2)
    y = planes[0][:H, :W].astype(np.int64)
    cb = fancy_upsample(planes[1], hmax, vmax)[:H, :W] - 128
    cr = fancy_upsample(planes[2], hmax, vmax)[:H, :W] - 128
    r = y + ((91881 * cr + 32768) >> 16)
    g = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16)
    b = y + ((116130 * cb + 32768) >> 16)
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)
