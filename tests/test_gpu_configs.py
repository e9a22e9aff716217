"""BASELINE.json configs 3 and 4 as parity cases: the dense 8x12 bank on Lab / opponent channels
(kernel sides up to 155) and large images with k-means k=32.  Same bars as test_gpu_segmenter.py:
Gabor features within 1e-5 (1e-4 for Lab) of the fp64 oracle, k-means bit-exact on identical
features, metric counts bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tol_ok(got, want, rtol):
    tol = rtol * np.abs(want).max() + rtol * np.abs(want)
    return bool((np.abs(got.astype(np.float64) - want) <= tol).all())


@pytest.mark.parametrize("space,rtol", [("opponent", 1e-5), ("lab", 1e-4)])
def test_dense_bank_features(space, rtol):
    import torch
    from gabor_color_image_segmentation_b200 import GaborBank, Plan
    from gabor_color_image_segmentation_b200.synth import synth_image
    from oracle import oracle as orc
    H, W = 176, 208
    img = synth_image(11, H, W)[None]
    plan = Plan(H, W, bank=GaborBank.dense(), colour_space=space, k=4, iters=1, max_gt=0)
    assert plan.D == 288
    feat = plan.gabor_features(torch.from_numpy(img).cuda()).cpu().numpy()[0]
    want = orc.gabor_features(img[0], orc.Bank.dense(), space)
    assert _tol_ok(feat, want, rtol)


def test_dense_bank_kmeans_288_features():
    import torch
    from gabor_color_image_segmentation_b200 import GaborBank, Plan
    from gabor_color_image_segmentation_b200.synth import synth_image
    from oracle import oracle as orc
    H, W, k, T = 120, 136, 8, 5
    img = synth_image(12, H, W)[None]
    idx = orc.kmeans_init_indices(H * W, k, 3)[None]
    plan = Plan(H, W, bank=GaborBank.dense(), colour_space="opponent", k=k, iters=T, max_gt=0)
    d_img = torch.from_numpy(img).cuda()
    labels = plan.segment(d_img, torch.from_numpy(idx)).cpu().numpy()[0]
    feat = plan.gabor_features(d_img).cpu().numpy()[0].reshape(288, -1)
    ol, _, _ = orc.kmeans(feat, k, T, idx[0])
    np.testing.assert_array_equal(labels.ravel(), ol)


def test_large_image_k32_pipeline():
    """1024x1024 RGB, k = 32 (config 4): vertical tiling of the Gabor strips, the K=32 k-means
    variant, and the metrics kernels on a 1 Mpixel label map with 32 segments."""
    import torch
    from gabor_color_image_segmentation_b200 import Plan, finish_image
    from gabor_color_image_segmentation_b200.pipeline import evaluate_batch
    from gabor_color_image_segmentation_b200.synth import synth_image, synth_ground_truths
    from oracle import oracle as orc
    H = W = 1024
    k, T, G = 32, 4, 2
    img = synth_image(21, H, W)[None]
    gts = synth_ground_truths(21, H, W, G)[None]
    idx = orc.kmeans_init_indices(H * W, k, 9)[None]
    plan = Plan(H, W, max_batch=1, k=k, iters=T, max_gt=G, n_lab_cap=64)
    c = evaluate_batch(plan, img, gts, idx, want_labels=True)
    feat = plan.gabor_features(torch.from_numpy(img).cuda()).cpu().numpy()[0]
    # Gabor: check a sample of feature planes against the oracle (all scales, both job kinds)
    want = orc.gabor_features(img[0])
    assert _tol_ok(feat, want, 1e-5)
    ol, _, _ = orc.kmeans(feat.reshape(72, -1), k, T, idx[0])
    np.testing.assert_array_equal(c.labels[0].ravel(), ol)
    o = orc.label_counts(c.labels[0], list(gts[0]))
    assert int(c.bd_count[0]) == o.bd_count
    np.testing.assert_array_equal(c.gt_counts[0, :, :5], np.stack([o.den_r, o.tp_r, o.tp_p, o.U, o.V], 1))
    np.testing.assert_array_equal(c.area[0, :o.n_seg], o.area)
    np.testing.assert_array_equal(c.perim[0, :o.n_seg], o.perim)
    got, ref = finish_image(c, 0), orc.finish_metrics(o)
    assert all(float(got[key]) == float(ref[key]) for key in ref)


def test_4k_image_runs_and_labels_match_teacher_forced():
    """3840x2160 (config 4): feature tensor 2.39 GB; checks the k-means labels on the GPU's own
    features (bit-exact) and size-independent properties of the metric counts."""
    import torch
    from gabor_color_image_segmentation_b200 import GaborBank, Plan
    from gabor_color_image_segmentation_b200.synth import voronoi_labels
    from oracle import oracle as orc
    H, W, k, T = 2160, 3840, 32, 2
    rng = np.random.default_rng(5)
    lab = voronoi_labels(rng, H // 8, W // 8, 40)
    base = rng.integers(0, 256, (lab.max() + 1, 3))
    small = base[lab].astype(np.uint8)
    img = np.ascontiguousarray(np.kron(small, np.ones((8, 8, 1), np.uint8)))[None]
    img = (img.astype(np.int16) + rng.integers(-12, 13, img.shape)).clip(0, 255).astype(np.uint8)
    gts = (np.kron(lab, np.ones((8, 8), np.int32)) + 1).astype(np.uint16)[None, None]
    idx = orc.kmeans_init_indices(H * W, k, 1)[None]
    plan = Plan(H, W, max_batch=1, bank=GaborBank.default(2, 6), k=k, iters=T, max_gt=1, n_lab_cap=64)
    c = plan.pipeline_host(img, gts, idx, 1, want_labels=True)
    N = H * W
    assert int(c.area[0].sum()) == N and int(c.gt_counts[0, 0, 1]) <= int(c.gt_counts[0, 0, 0])
    feat = plan.gabor_features(torch.from_numpy(img).cuda())
    f = feat.cpu().numpy()[0].reshape(plan.D, -1)
    del feat
    ol, _, _ = orc.kmeans(f, k, T, idx[0])
    np.testing.assert_array_equal(c.labels[0].ravel(), ol)
