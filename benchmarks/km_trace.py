"""Per-CTA phase timeline of the k-means pass kernel (needs a library built with -DKM_TRACE).
    GCIS_LIB=build/libgcis_trace.so python benchmarks/km_trace.py --images 2
Prints, per pass, the median cycles a CTA spends in: prologue, stream (phase A), label/compaction,
centroid update (phase B), ticket; plus the pass's span (first CTA start -> last CTA end, globaltimer)."""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=2)
    ap.add_argument("--ctas", type=int, default=0, help="CTAs to summarise (0 = one per 1024-pixel tile)")
    a = ap.parse_args()
    import torch
    from gabor_color_image_segmentation_b200 import Plan, _lib
    from gabor_color_image_segmentation_b200.pipeline import init_indices_for
    from gabor_color_image_segmentation_b200.synth import synth_batch
    H, W, G, k = 321, 481, 5, 8
    u = min(8, a.images)
    imgs, gts = synth_batch(u, H, W, G)
    reps = (a.images + u - 1) // u
    imgs = np.concatenate([imgs] * reps)[:a.images]
    idx = init_indices_for(range(a.images), H * W, k)
    plan = Plan(H, W, max_batch=a.images, k=k, iters=20, max_gt=G, group=a.images)
    d_img = torch.from_numpy(imgs).cuda(); d_idx = torch.from_numpy(idx).cuda()
    for _ in range(3):
        plan.segment(d_img, d_idx)
    torch.cuda.synchronize()
    lib = _lib.load()
    P, C, S = 20, 1024, 8
    buf = np.zeros((P, C, S), np.int64)
    lib.gcis_km_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    rc = lib.gcis_km_trace_read(buf.ctypes.data, buf.nbytes)
    assert rc == 0, rc
    n_cta = min(C, a.ctas or 151 * a.images)
    names = ["prologue", "stream", "labels", "update", "ticket"]
    print("pass  " + "  ".join("%9s" % n for n in names) + "   cta_total   starts_spread_us")
    for p in range(P):
        t = buf[p, :n_cta]
        d = np.diff(t[:, :6], axis=1)
        med = np.median(d, axis=0)
        tot = np.median(t[:, 5] - t[:, 0])
        g = t[:, 7]
        print("%4d  " % p + "  ".join("%9.0f" % v for v in med) + "   %9.0f   %8.2f" % (tot, (g.max() - g.min()) / 1e3)
              + "   mean update %6.0f, tiles with update > 1000 cycles: %4.1f %%" % (d[:, 3].mean(), 100 * (d[:, 3] > 1000).mean()))
    if a.ctas:   # persistent tile kernel: slot 4 = end, 5 = cycles thread 0 waited for data, 6 = tiles done
        for p in (0, 1, 5, 10, 19):
            t = buf[p, :n_cta]
            tot = t[:, 4] - t[:, 0]
            print("pass %2d: kernel cycles/CTA %.0f, waiting for data %.0f (%.0f%%), tiles/CTA %.1f, cycles/tile %.0f" % (
                p, np.median(tot), np.median(t[:, 5]), 100 * np.median(t[:, 5] / tot), np.mean(t[:, 6]), np.median(tot / t[:, 6])))
        return
    fin = buf[:, :n_cta, 6] - buf[:, :n_cta, 5]
    print("finalize cycles (last CTA per image, pass 10):", fin[10][buf[10, :n_cta, 6] > 0])


if __name__ == "__main__":
    main()
