"""Time the Gabor kernel per scale of the default bank (single-scale plans, CUDA events)."""
import os, sys, json, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gabor_color_image_segmentation_b200 import GaborBank, Plan
from gabor_color_image_segmentation_b200.synth import synth_batch
B, H, W = 32, int(os.environ.get('GB_H', 321)), 481
imgs, _ = synth_batch(4, H, W, 1)
d_img = torch.from_numpy(np.concatenate([imgs] * 8)).cuda()
full = GaborBank.default()
for name, bank in [("all", full)] + [("s%d" % s, GaborBank((f,), full.thetas)) for s, f in enumerate(full.frequencies)]:
    plan = Plan(H, W, max_batch=B, bank=bank, k=2, iters=1, max_gt=0, group=B)
    plan.gabor_features(d_img)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        plan.gabor_features(d_img)
    e1.record(); torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / 5 / B * 1e3, 1), "us/image", round(e0.elapsed_time(e1) / 5 / B * 1e6 / (H * W), 3), "ns/pixel")
    plan.close()
