"""Accuracy of the two filter-bank kernels against the fp64 CPU checker on full-size images (BENCH / profile helper:
it executes oracle/ as the checker, like the tests).  Prints one JSON line per kernel:
max |gpu - ref64| / max|ref64| and the largest fraction of the 1e-5 tolerance used, per scale."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from gabor_color_image_segmentation_b200 import Plan
    from gabor_color_image_segmentation_b200.synth import synth_image
    from oracle import oracle as orc
    imgs = np.stack([synth_image(i) for i in range(2)])
    want = np.stack([orc.gabor_features(im) for im in imgs])
    d = torch.from_numpy(imgs).cuda()
    for tc in ("1", "0"):
        os.environ["GCIS_GABOR_TC"] = tc
        plan = Plan(321, 481, max_batch=2, max_gt=0)
        got = plan.gabor_features(d).cpu().numpy().astype(np.float64)
        err = np.abs(got - want)
        tol = 1e-5 * np.abs(want).max() + 1e-5 * np.abs(want)
        per_scale = err.reshape(2, 3, 4, 6, 321, 481).max(axis=(0, 1, 3, 4, 5)) / np.abs(want).max()
        print(json.dumps({"kernel": "gabor_tc_kernel (row pass tcgen05 bf16x3)" if plan.uses_tensor_cores else "gabor_bank_kernel (FP32 pipe)",
                          "max_abs_err_over_max_ref": float(err.max() / np.abs(want).max()),
                          "max_fraction_of_1e-5_tolerance": float((err / tol).max()),
                          "per_scale_max_err_over_max_ref": [float(v) for v in per_scale]}))
        plan.close()


if __name__ == "__main__":
    main()
