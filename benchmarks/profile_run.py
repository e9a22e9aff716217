"""One pass of the hot path over a small synthetic batch — the command profiled under ncu.
    python benchmarks/profile_run.py --images 16 --group 16 [--reps 1]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--group", type=int, default=0)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--unique", type=int, default=8)
    a = ap.parse_args()
    import torch
    from gabor_color_image_segmentation_b200 import Plan
    from gabor_color_image_segmentation_b200.pipeline import init_indices_for
    from gabor_color_image_segmentation_b200.synth import synth_batch
    H, W, G, k = 321, 481, 5, 8
    u = min(a.unique, a.images)
    imgs, gts = synth_batch(u, H, W, G)
    reps = (a.images + u - 1) // u
    imgs = np.concatenate([imgs] * reps)[:a.images]; gts = np.concatenate([gts] * reps)[:a.images]
    idx = init_indices_for(range(a.images), H * W, k)
    plan = Plan(H, W, max_batch=a.images, k=k, iters=20, max_gt=G, group=a.group)
    d_img = torch.from_numpy(imgs).cuda(); d_gt = torch.from_numpy(gts.view(np.int16)).cuda(); d_idx = torch.from_numpy(idx).cuda()
    for _ in range(a.reps):
        plan.pipeline_device(d_img, d_gt, d_idx)
        c = plan.fetch()
    torch.cuda.synchronize()
    print("ok", int(c.bd_count.sum()))


if __name__ == "__main__":
    main()
