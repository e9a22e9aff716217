"""Stage times (CUDA events inside the library) of one configuration of the hot path.
    python benchmarks/config_times.py --height 1024 --width 1024 --k 32 --images 8
    python benchmarks/config_times.py --dense --colour opponent --images 16   # 8 half-octave scales x 12 orientations"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=321)
    ap.add_argument("--width", type=int, default=481)
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--scales", type=int, default=4)
    ap.add_argument("--orient", type=int, default=6)
    ap.add_argument("--colour", default="rgb")
    ap.add_argument("--dense", action="store_true", help="BASELINE config 3: the dense 8 x 12 bank (D = 288)")
    a = ap.parse_args()
    import torch
    from gabor_color_image_segmentation_b200 import GaborBank, Plan
    from gabor_color_image_segmentation_b200.pipeline import init_indices_for
    os.environ["GCIS_LANES"] = "1"
    H, W, B = a.height, a.width, a.images
    rng = np.random.default_rng(1)
    imgs = rng.integers(0, 256, (B, H, W, 3)).astype(np.uint8)
    gts = rng.integers(1, 20, (B, 1, H, W)).astype(np.uint16)
    bank = GaborBank.dense() if a.dense else (GaborBank.default(a.scales, a.orient) if (a.scales, a.orient) != (4, 6) else None)
    kw = dict(bank=bank) if bank is not None else {}
    plan = Plan(H, W, max_batch=B, k=a.k, iters=a.iters, max_gt=1, colour_space=a.colour, group=min(B, 64), **kw)
    idx = init_indices_for(range(B), H * W, a.k)
    d_img = torch.from_numpy(imgs).cuda(); d_gt = torch.from_numpy(gts.view(np.int16)).cuda(); d_idx = torch.from_numpy(idx).cuda()
    plan.pipeline_device(d_img, d_gt, d_idx); plan.fetch()
    plan.set_profiling(True)
    plan.pipeline_device(d_img, d_gt, d_idx); plan.fetch()
    st = plan.last_stage_ms()
    N, D = H * W, plan.D
    print("%dx%d D=%d k=%d images=%d:" % (H, W, D, a.k, B), {k: round(v, 3) for k, v in st.items()},
          "| k-means %.0f GB/s, Gabor %.1f us/image" % (B * a.iters * N * D * 4 / (st["kmeans"] * 1e-3) / 1e9, st["gabor"] * 1e3 / B))


if __name__ == "__main__":
    main()
