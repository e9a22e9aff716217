"""Spatial (this repo's separable FFMA2 kernel) vs FFT (cuFFT through torch.fft, comparison arm
only) for the Gabor bank, per kernel size — BASELINE.json config 3 (dense 8x12 bank).

    python benchmarks/spatial_vs_fft.py [--images 16] [--orient 12]

For every scale of the dense bank a single-scale plan is timed with CUDA events; the FFT arm
reflect-pads the three colour planes by the scale's half-width, multiplies their rfft2 with the
precomputed spectra of the scale's kernels, inverts and takes the magnitude (fp32 cuFFT; its
accuracy against the fp64 oracle is printed next to the spatial kernel's)."""
import argparse
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--orient", type=int, default=12)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from gabor_color_image_segmentation_b200 import GaborBank, Plan
    from gabor_color_image_segmentation_b200.synth import synth_batch
    from oracle import oracle as orc
    H, W, B, O = 321, 481, a.images, a.orient
    imgs, _ = synth_batch(min(B, 4), H, W, 1)
    imgs = np.concatenate([imgs] * ((B + 3) // 4))[:B]
    d_img = torch.from_numpy(imgs).cuda()
    dense = GaborBank.dense()
    thetas = tuple(o * math.pi / O for o in range(O))
    rows = []
    for s, f in enumerate(dense.frequencies):
        bank = GaborBank((f,), thetas)
        h = max(bank.half_width(0, o) for o in range(O))
        plan = Plan(H, W, max_batch=B, bank=bank, colour_space="rgb", k=2, iters=1, max_gt=0, group=B)
        feat = plan.gabor_features(d_img)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize(); ev[0].record()
        for _ in range(a.reps):
            feat = plan.gabor_features(d_img)
        ev[1].record(); torch.cuda.synchronize()
        t_sp = ev[0].elapsed_time(ev[1]) / a.reps / B * 1e3          # us per image
        # ---- FFT arm ----
        x = d_img.permute(0, 3, 1, 2).float() / 255.0                 # [B,3,H,W]
        def refl(n):   # scipy 'reflect' (d c b a | a b c d | d c b a) index map for [-h, n + h)
            i = np.arange(-h, n + h) % (2 * n)
            return torch.from_numpy(np.where(i < n, i, 2 * n - 1 - i)).cuda()
        xp = x[:, :, refl(H)][:, :, :, refl(W)].contiguous()
        PH, PW = H + 2 * h, W + 2 * h
        ker = torch.zeros((O, PH, PW), dtype=torch.complex64, device="cuda")
        for o in range(O):
            g = orc.gabor_kernel(f, thetas[o])
            kh = g.shape[0] // 2
            gk = torch.from_numpy(g.astype(np.complex64)).cuda()
            # place the kernel so that circular convolution == linear convolution on the valid region
            idx = (torch.arange(-kh, kh + 1, device="cuda") % PH)[:, None], (torch.arange(-kh, kh + 1, device="cuda") % PW)[None, :]
            ker[o][idx] = gk
        Kf = torch.fft.fft2(ker)                                       # [O,PH,PW]

        def fft_arm():
            Xf = torch.fft.fft2(xp.to(torch.complex64))               # [B,3,PH,PW]
            Y = torch.fft.ifft2(Xf[:, :, None] * Kf[None, None])      # [B,3,O,PH,PW]
            return Y[..., h:h + H, h:h + W].abs()
        y = fft_arm()
        torch.cuda.synchronize(); ev[0].record()
        for _ in range(a.reps):
            y = fft_arm()
        ev[1].record(); torch.cuda.synchronize()
        t_fft = ev[0].elapsed_time(ev[1]) / a.reps / B * 1e3
        # accuracy of both arms on image 0, channel 0 against the fp64 oracle
        want = orc.gabor_features(imgs[0], orc.Bank((f,), thetas))
        scale = np.abs(want).max()
        e_sp = float(np.abs(feat[0].cpu().numpy() - want).max() / scale)
        got_fft = y[0].reshape(3 * O, H, W).cpu().numpy()
        e_fft = float(np.abs(got_fft - want).max() / scale)
        rows.append({"scale": s, "frequency": f, "kernel_side": 2 * h + 1, "spatial_us_per_image": round(t_sp, 1),
                     "fft_us_per_image": round(t_fft, 1), "spatial_max_err_rel": e_sp, "fft_max_err_rel": e_fft})
        print(json.dumps(rows[-1]))
        plan.close()
        del ker, Kf, y, xp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
