"""Synthetic BSDS-shaped data (SURVEY.md §8 d): seeded Voronoi-texture RGB images with
random Voronoi ground-truth partitions.  The dataset itself is not needed for benchmarks."""
import numpy as np

SEED0 = 1234


def voronoi_labels(rng, H, W, R, base=0):
    """H x W partition into the Voronoi cells of R random sites, labels base..base+R'-1
    (contiguous, relabelled in order of first appearance)."""
    ys = rng.integers(0, H, R).astype(np.int32)
    xs = rng.integers(0, W, R).astype(np.int32)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.int32)
    best = np.full((H, W), np.iinfo(np.int32).max, np.int32)
    lab = np.zeros((H, W), np.int32)
    for i in range(R):
        d = (yy - ys[i]) ** 2 + (xx - xs[i]) ** 2
        m = d < best
        best[m] = d[m]
        lab[m] = i
    _, inv = np.unique(lab, return_inverse=True)
    return inv.reshape(H, W).astype(np.int32) + base


def synth_image(index, H=321, W=481, seed0=SEED0):
    """uint8 H x W x 3: Voronoi regions (4..24 sites), each with a base colour, an oriented
    sinusoidal texture (0.03..0.25 cyc/px, amplitude 32) and N(0, 8^2) noise."""
    rng = np.random.default_rng(seed0 + index)
    R = int(rng.integers(4, 25))
    lab = voronoi_labels(rng, H, W, R)
    n = int(lab.max()) + 1
    base = rng.integers(0, 256, (n, 3)).astype(np.float64)
    freq = rng.uniform(0.03, 0.25, n)
    theta = rng.uniform(0, np.pi, n)
    yy, xx = np.mgrid[0:H, 0:W]
    phase = 2 * np.pi * freq[lab] * (xx * np.cos(theta[lab]) + yy * np.sin(theta[lab]))
    img = base[lab] + 32.0 * np.sin(phase)[..., None] + rng.normal(0, 8.0, (H, W, 3))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_ground_truths(index, H=321, W=481, G=5, seed0=SEED0, r_lo=2, r_hi=50):
    """[G,H,W] uint16: independent Voronoi partitions, labels 1..R_g (groundtruth.py:26 layout)."""
    rng = np.random.default_rng(seed0 + 1_000_003 * (index + 1))
    out = np.zeros((G, H, W), np.uint16)
    for g in range(G):
        while True:
            lab = voronoi_labels(rng, H, W, int(rng.integers(r_lo, r_hi + 1)), base=1)
            if lab.max() >= 2:      # a single-region ground truth raises in the reference (A.8)
                break
        out[g] = lab
    return out


def synth_batch(B, H=321, W=481, G=5, seed0=SEED0, start=0):
    imgs = np.stack([synth_image(start + i, H, W, seed0) for i in range(B)])
    gts = np.stack([synth_ground_truths(start + i, H, W, G, seed0) for i in range(B)])
    return imgs, gts
