"""Oracle self-checks for the optional feature-assembly steps (north_star: "optional smoothing and
normalisation"; builder-defined, parity unpinned upstream): smoothing against scipy's 1-D
convolution, the integer-moment z-score map against numpy's mean/std, and the k-means with the folded map
against a textbook fp64 Lloyd on explicitly normalised features."""
import numpy as np
from scipy import ndimage as ndi

from oracle import oracle as orc


def _features(seed=0, H=48, W=64):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    return img, orc.Bank.default(2, 4)


def test_smoothing_matches_scipy_convolve1d():
    img, bank = _features()
    raw = orc.gabor_features(img, bank)
    sm = orc.gabor_features(img, bank, smooth=0.75)
    S, O = 2, 4
    for d in (0, 5, 13, 23):
        s = (d // O) % S
        g = orc.smoothing_taps(0.75 * orc.gabor_sigma(bank.frequencies[s]))
        assert abs(g.sum() - 1) < 1e-15 and len(g) % 2 == 1 and g[0] == g[-1]
        want = ndi.convolve1d(ndi.convolve1d(raw[d], g, axis=1, mode="reflect"), g, axis=0, mode="reflect")
        np.testing.assert_allclose(sm[d], want, rtol=0, atol=1e-14)


def test_feature_affine_against_numpy_moments():
    img, bank = _features(1, 64, 80)
    f = orc.gabor_features(img, bank).reshape(24, -1).astype(np.float32)
    f[7] = 0.125                                       # a constant plane: a = b = 0 by definition
    ab = orc.feature_affine(f)
    x = f.astype(np.float64)
    mean, sd = x.mean(1), x.std(1)
    ok = sd > 0
    assert ab[7, 0] == 0 and ab[7, 1] == 0
    np.testing.assert_allclose(ab[ok, 0], 1 / sd[ok], rtol=2e-4)       # second moment of the 16-bit fixed-point values: cross term ~ 1/sqrt(N) on this 5k-pixel image
    np.testing.assert_allclose(ab[ok, 1], -mean[ok] / sd[ok], rtol=2e-4, atol=1e-4)
    z = ab[:, :1] * f + ab[:, 1:]
    np.testing.assert_allclose(z[ok].mean(1), 0, atol=5e-4)
    np.testing.assert_allclose(z[ok].std(1), 1, atol=5e-4)


def _lloyd(X, k, T, idx):
    c = X[idx].copy()
    for _ in range(T):
        lab = np.stack([((X - c[j]) ** 2).sum(1) for j in range(k)], 1).argmin(1)
        for j in range(k):
            if (lab == j).any():
                c[j] = X[lab == j].mean(0)
    return lab, c


def test_kmeans_with_folded_normalisation_equals_lloyd_on_zscores():
    from gabor_color_image_segmentation_b200.synth import synth_image
    f = orc.gabor_features(synth_image(4, 96, 128), orc.Bank.default(3, 6)).reshape(54, -1).astype(np.float32)
    idx = orc.kmeans_init_indices(f.shape[1], 6, 5)
    ab = orc.feature_affine(f)
    lab, cent, counts = orc.kmeans(f, 6, 10, idx, affine=ab)
    z = (ab[:, :1].astype(np.float64) * f + ab[:, 1:]).T
    lab_np, cent_np = _lloyd(z, 6, 10, idx)
    assert (lab != lab_np).mean() < 1e-3
    np.testing.assert_allclose(cent, cent_np, rtol=0, atol=2e-4)       # centroids live in z space
    assert counts.sum() == f.shape[1]
    # ... and it is NOT the clustering of the raw features (the option does something)
    lab_raw, _, _ = orc.kmeans(f, 6, 10, idx)
    assert (lab_raw != lab).mean() > 0.01
    # without a map the new entry point is the old one, bit for bit
    l0, c0, _ = orc.kmeans(f, 6, 4, idx, affine=None)
    ident = np.stack([np.ones(54, np.float32), np.zeros(54, np.float32)], 1)
    l1, c1, _ = orc.kmeans(f, 6, 4, idx, affine=ident)
    np.testing.assert_array_equal(l0, l1)
    np.testing.assert_array_equal(c0, c1)
