"""CPU-side pins for the builder-defined segmenter stages and the real-data fixture.

* the oracle's fp32 score chain against an independent fp64 assignment at FULL size (321x481, D=72,
  k=8): every disagreement must be a near-tie, (second - best) / best <= 1e-5 (north_star:
  "assignments agree except where distances tie within that tolerance");
* the oracle's 20-iteration Lloyd against an independent numpy fp64 Lloyd at full size;
* the real BSDS500 fixture: decoded pixels, loader, oracle metrics == the REFERENCE's own metrics.py
  outputs (tests/golden/bsds500_golden.npz, made by oracle/make_real_fixture.py)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "bsds500")
NEAR_TIE = 1e-5


def near_tie_report(feat32, cent, labels):
    """(n_disagree, worst relative gap): fp64 assignment with `cent` vs `labels`; the gap of a
    disagreeing pixel is (d2[labels] - d2[best]) / d2[best] in fp64."""
    l64, b1, _ = orc.kmeans_assign_f64(feat32, cent)
    dis = np.flatnonzero(l64 != labels)
    if not len(dis):
        return 0, 0.0
    X = feat32[:, dis].astype(np.float64).T
    d = ((X - cent[labels[dis]].astype(np.float64)) ** 2).sum(1)
    return len(dis), float(((d - b1[dis]) / b1[dis]).max())


@pytest.fixture(scope="module")
def full_size_features():
    from gabor_color_image_segmentation_b200.synth import synth_image
    out = []
    for i in (0, 6):      # image 6 holds a pixel whose two best clusters tie within 6.6e-6 in the first pass
        f = orc.gabor_features(synth_image(i)).reshape(72, -1).astype(np.float32)
        out.append((i, f, orc.kmeans_init_indices(f.shape[1], 8, i)))
    return out


def test_fp32_score_chain_vs_fp64_assignment_full_size(full_size_features):
    for i, f, idx in full_size_features:
        for T in (1, 2, 20):
            labels, _, _ = orc.kmeans(f, 8, T, idx)
            prev = f[:, idx].T.copy() if T == 1 else orc.kmeans(f, 8, T - 1, idx)[1]   # centroids of the last assignment
            n, gap = near_tie_report(f, prev, labels)
            assert gap <= NEAR_TIE, (i, T, n, gap)
            assert n <= 4, (i, T, n)


def _lloyd_numpy_full(f32, k, T, idx):
    X = f32.T.astype(np.float64)                    # [N, D]
    c = X[idx].copy()
    lab = None
    for _ in range(T):
        d2 = np.stack([((X - c[j]) ** 2).sum(1) for j in range(k)], 1)
        lab = d2.argmin(1)                          # lowest index wins ties, like the spec
        for j in range(k):
            m = lab == j
            if m.any():
                c[j] = X[m].mean(0)
    return lab, c


def test_kmeans_full_size_against_independent_numpy_lloyd(full_size_features):
    """321x481, D=72, k=8, T=20 (config 1): the exact-integer fp32 contract vs a textbook fp64 Lloyd."""
    i, f, idx = full_size_features[0]
    lab, cent, counts = orc.kmeans(f, 8, 20, idx)
    lab_np, cent_np = _lloyd_numpy_full(f, 8, 20, idx)
    dis = np.flatnonzero(lab != lab_np)
    assert len(dis) <= 1e-4 * lab.size, len(dis)
    # rint(x * 2^24) quantisation + fp32 centroids: centroids agree to ~1e-7 absolute
    np.testing.assert_allclose(cent, cent_np, rtol=0, atol=2e-6)
    if len(dis):   # whatever differs is a near-tie under the independent centroids
        X = f[:, dis].astype(np.float64).T
        prev = orc.kmeans(f, 8, 19, idx)[1].astype(np.float64)
        d = np.stack([((X - prev[j]) ** 2).sum(1) for j in range(8)], 1)
        gap = (d[np.arange(len(dis)), lab[dis]] - d.min(1)) / d.min(1)
        assert gap.max() <= 1e-3, float(gap.max())
    assert counts.sum() == lab.size


# ---- real-data fixture -------------------------------------------------------------------------

@pytest.fixture(scope="module")
def real_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "bsds500_golden.npz"))


def _decode(fid):
    from PIL import Image
    return np.asarray(Image.open(os.path.join(FIX, "images", fid + ".jpg")))


def test_real_fixture_pixels_and_loader(real_golden):
    from gabor_color_image_segmentation_b200 import get_segmentation
    for fid in real_golden["ids"]:
        fid = str(fid)
        img = _decode(fid)
        assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == str(real_golden[fid + "/pixels_sha256"]), fid
        gts = get_segmentation(os.path.join(FIX, "truth") + "/", fid)
        assert len(gts) == int(real_golden[fid + "/n_gt"])
        assert all(g.dtype == np.uint16 and g.shape == img.shape[:2] and g.min() >= 1 for g in gts)


def test_real_fixture_oracle_metrics_equal_reference(real_golden):
    """(oracle labels of the real image, real ground truths) -> oracle metrics == the reference's own
    metrics.py outputs, integers and floats bit for bit."""
    from gabor_color_image_segmentation_b200 import get_segmentation
    keys = ["recall", "precision", "underseg", "undersegNP", "compactness", "density"]
    for fid in real_golden["ids"]:
        fid = str(fid)
        gts = get_segmentation(os.path.join(FIX, "truth") + "/", fid)
        labels = real_golden[fid + "/labels"]
        o = orc.label_counts(labels, gts)
        assert o.bd_count == int(real_golden[fid + "/bd_count"])
        np.testing.assert_array_equal(o.den_r, real_golden[fid + "/den_r"])
        np.testing.assert_array_equal(o.tp_r, real_golden[fid + "/tp_r"])
        np.testing.assert_array_equal(o.tp_p, real_golden[fid + "/tp_p"])
        np.testing.assert_array_equal(o.perim.astype(np.float64), real_golden[fid + "/perimeters"])
        m = orc.finish_metrics(o)
        assert int(m["regions"]) == int(real_golden[fid + "/regions"])
        for j, key in enumerate(keys):
            assert float(m[key]) == float(real_golden[fid + "/floats"][j]), (fid, key)


def test_real_fixture_oracle_segmentation_is_reproducible(real_golden):
    fid = str(real_golden["ids"][2])    # 3096
    img = _decode(fid)
    idx = orc.kmeans_init_indices(img.shape[0] * img.shape[1], int(real_golden["k"]), 2)
    labels, _, _ = orc.segment_image(img, int(real_golden["k"]), int(real_golden["iters"]), init_idx=idx)
    np.testing.assert_array_equal(labels, real_golden[fid + "/labels"])
