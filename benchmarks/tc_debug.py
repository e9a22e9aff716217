"""Bring-up helper for the tensor-core filter-bank kernel: one gabor_features call, compared with the FP32-pipe
kernel (GCIS_GABOR_TC=0).  Usage: GCIS_LIB=build/libgcis_dbgN.so python benchmarks/tc_debug.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from gabor_color_image_segmentation_b200 import GaborBank, Plan
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 56)
    bank = GaborBank.default(2, 3) if H < 100 else GaborBank.default()
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (2, H, W, 3)).astype(np.uint8)
    d = torch.from_numpy(img).cuda()
    os.environ["GCIS_GABOR_TC"] = "0"
    ref = Plan(H, W, max_batch=2, bank=bank, k=4, iters=2, max_gt=0).gabor_features(d).cpu().numpy()
    os.environ["GCIS_GABOR_TC"] = "1"
    plan = Plan(H, W, max_batch=2, bank=bank, k=4, iters=2, max_gt=0)
    got = plan.gabor_features(d)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    err = np.abs(got - ref)
    print("ok: max |tc - fp32| = %.3e, max |ref| = %.3e, rel = %.3e" % (err.max(), np.abs(ref).max(), err.max() / np.abs(ref).max()))
    per = err.reshape(2, -1, H, W).max(axis=(0, 2, 3))
    print("per plane max err:", np.array2string(per, precision=2, max_line_width=200))


if __name__ == "__main__":
    main()
