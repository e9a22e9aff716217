"""Oracle self-checks for the builder-defined stages (Gabor bank, k-means).  Upstream holds no
code for them (parity unpinned); these tests pin the oracle's primitives to independent
implementations: scipy.ndimage.convolve for the convolution, a numpy fp64 Lloyd for k-means."""
import math

import numpy as np
from scipy import ndimage as ndi

from oracle import oracle as orc


def test_gabor_kernel_shapes_match_spec():
    bank = orc.Bank.default()
    sides = [[orc.gabor_kernel(f, th).shape[0] for th in bank.thetas] for f in bank.frequencies]
    assert sides[0] == [15, 13, 13, 15, 13, 13]
    assert [s[0] for s in sides] == [15, 29, 55, 109]
    assert [s[1] for s in sides] == [13, 25, 49, 95]
    dense = orc.Bank.dense()
    assert [orc.gabor_kernel(f, 0.0).shape[0] for f in dense.frequencies] == [15, 21, 29, 41, 55, 79, 109, 155]


def test_separable_factors_reproduce_kernel():
    for f in (0.25, 0.0625):
        for th in (0.0, math.pi / 6, math.pi / 2, 5 * math.pi / 6):
            g = orc.gabor_kernel(f, th)
            gx, gy = orc.gabor_separable(f, th)
            np.testing.assert_allclose(np.outer(gy, gx), g, rtol=0, atol=1e-15)


def test_conv2d_reflect_matches_scipy():
    rng = np.random.default_rng(0)
    for (H, W, kh, kw) in [(17, 23, 5, 7), (9, 40, 13, 13), (6, 5, 15, 15)]:   # last: kernel wider than image
        img = rng.random((H, W))
        ker = rng.standard_normal((kh, kw))
        want = ndi.convolve(img, ker, mode="reflect")
        np.testing.assert_allclose(orc.conv2d_reflect(img, ker), want, rtol=1e-12, atol=1e-12)


def test_separable_equals_direct_and_scipy():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (40, 56)).astype(np.float64) / 255.0
    for f, th in [(0.25, 0.0), (0.125, math.pi / 6), (0.0625, 2 * math.pi / 3), (0.125, math.pi / 2)]:
        g = orc.gabor_kernel(f, th)
        gx, gy = orc.gabor_separable(f, th)
        re, im = orc.conv_sep_complex(img, gx, gy)
        np.testing.assert_allclose(re, ndi.convolve(img, g.real, mode="reflect"), rtol=0, atol=1e-13)
        np.testing.assert_allclose(im, ndi.convolve(img, g.imag, mode="reflect"), rtol=0, atol=1e-13)
        np.testing.assert_allclose(re, orc.conv2d_reflect(img, g.real), rtol=0, atol=1e-13)


def test_feature_layout_and_direct_path():
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (24, 32, 3)).astype(np.uint8)
    bank = orc.Bank.default(2, 3)
    a = orc.gabor_features(img, bank)
    b = orc.gabor_features(img, bank, direct2d=True)
    assert a.shape == (3 * 2 * 3, 24, 32)
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-13)
    # d = (c*S + s)*O + o
    g = orc.gabor_kernel(bank.frequencies[1], bank.thetas[2])
    ch = img[..., 2].astype(np.float64) / 255.0
    want = np.hypot(ndi.convolve(ch, g.real, mode="reflect"), ndi.convolve(ch, g.imag, mode="reflect"))
    np.testing.assert_allclose(a[(2 * 2 + 1) * 3 + 2], want, rtol=0, atol=1e-13)


def _lloyd_numpy(f, k, T, idx):
    c = f[:, idx].T.astype(np.float64)
    lab = None
    for _ in range(T):
        d2 = ((f.T[:, None, :].astype(np.float64) - c[None]) ** 2).sum(-1)
        lab = d2.argmin(1)
        for j in range(k):
            if (lab == j).any():
                c[j] = f[:, lab == j].astype(np.float64).mean(1)
    return lab, c


def test_kmeans_oracle_against_numpy_lloyd():
    rng = np.random.default_rng(3)
    k, D, N = 4, 6, 3000
    centres = rng.random((k, D)) * 2
    f = (centres[rng.integers(0, k, N)] + 0.05 * rng.standard_normal((N, D))).T.astype(np.float32)
    idx = orc.kmeans_init_indices(N, k, 11)
    lab, cent, counts = orc.kmeans(f, k, 8, idx)
    lab_np, cent_np = _lloyd_numpy(f, k, 8, idx)
    assert (lab == lab_np).mean() > 0.999       # fp32 score form vs fp64 distances: only near-ties may differ
    np.testing.assert_allclose(cent, cent_np, rtol=0, atol=1e-5)
    assert counts.sum() == N and (np.bincount(lab, minlength=k) == counts).all()


def test_kmeans_tie_goes_to_lowest_index_and_empty_cluster_keeps_centroid():
    f = np.array([[0.0, 0.0, 1.0, 1.0, 0.5]], np.float32)     # D=1, pixel 4 equidistant
    idx = np.array([0, 2, 1], np.int32)                       # centroid 2 duplicates centroid 0 -> stays empty
    lab, cent, counts = orc.kmeans(f, 3, 1, idx)
    assert lab.tolist() == [0, 0, 1, 1, 0]
    assert counts.tolist() == [3, 2, 0]
    assert cent[2, 0] == 0.0
    np.testing.assert_allclose(cent[:2, 0], [0.5 / 3, 1.0], atol=1e-7)
