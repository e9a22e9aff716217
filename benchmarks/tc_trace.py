"""Per-CTA phase cycles of the tensor-core filter-bank kernel's column warps (library built with -DTC_TRACE).
    GCIS_LIB=build/libgcis_tctrace.so python benchmarks/tc_trace.py"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from gabor_color_image_segmentation_b200 import Plan, _lib
    from gabor_color_image_segmentation_b200.synth import synth_batch
    B = 16
    imgs, _ = synth_batch(8, 321, 481, 1)
    imgs = np.concatenate([imgs, imgs])
    plan = Plan(321, 481, max_batch=B, max_gt=0, group=B)
    d_img = torch.from_numpy(imgs).cuda()
    for _ in range(2):
        plan.gabor_features(d_img)
    torch.cuda.synchronize()
    lib = _lib.load()
    buf = np.zeros((8192, 8), np.int64)
    lib.gcis_tc_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    assert lib.gcis_tc_trace_read(buf.ctypes.data, buf.nbytes) == 0
    names = ["setup", "wait acc", "transfer", "col pass", "taps", "tail bar", "total"]
    print("scale  CTAs  " + "  ".join("%9s" % n for n in names))
    for s in range(4):
        m = buf[(buf[:, 6] > 0) & (buf[:, 7] == s)]
        if len(m):
            print("%5d %5d  " % (s, len(m)) + "  ".join("%9.0f" % v for v in np.median(m[:, :7], axis=0)))


if __name__ == "__main__":
    main()
