"""Optional feature-assembly steps on the GPU (north_star "optional smoothing and normalisation")
against the oracle: smoothed features within 1e-5 of the fp64 oracle, the normalisation map bit-exact (it is
built from exact integer moments), and the k-means with the folded map bit-exact teacher-forced.  Both filter-bank
kernels are covered: rgb planes run the tensor-core row pass, opponent planes the FP32-pipe kernel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tol_ok(got, want, rtol=1e-5):
    tol = rtol * np.abs(want).max() + rtol * np.abs(want)
    return bool((np.abs(got.astype(np.float64) - want) <= tol).all())


@pytest.mark.parametrize("space,shape,smooth", [("rgb", (96, 128), 0.5), ("opponent", (70, 90), 1.0), ("rgb", (321, 481), 0.25)])
def test_smoothed_features_against_oracle(space, shape, smooth):
    import torch
    from gabor_color_image_segmentation_b200 import GaborBank, Plan
    from gabor_color_image_segmentation_b200.synth import synth_image
    from oracle import oracle as orc
    H, W = shape
    args = (3, 6) if H < 300 else (4, 6)
    img = synth_image(7, H, W)[None]
    plan = Plan(H, W, bank=GaborBank.default(*args), colour_space=space, k=4, iters=1, max_gt=0, smooth=smooth)
    assert plan.uses_tensor_cores == (space == "rgb")
    feat = plan.gabor_features(torch.from_numpy(img).cuda()).cpu().numpy()[0]
    want = orc.gabor_features(img[0], orc.Bank.default(*args), space, smooth=smooth)
    assert _tol_ok(feat, want)
    raw = orc.gabor_features(img[0], orc.Bank.default(*args), space)
    assert not _tol_ok(feat, raw)              # the option does something


@pytest.mark.parametrize("space,smooth", [("rgb", 0.0), ("opponent", 0.0), ("rgb", 0.5)])
def test_normalised_clustering_bit_exact(space, smooth):
    import torch
    from gabor_color_image_segmentation_b200 import GaborBank, Plan
    from gabor_color_image_segmentation_b200.synth import synth_image
    from oracle import oracle as orc
    H, W, k, T, B = 120, 160, 6, 8, 3
    imgs = np.stack([synth_image(30 + i, H, W) for i in range(B)])
    idx = np.stack([orc.kmeans_init_indices(H * W, k, i) for i in range(B)])
    plan = Plan(H, W, max_batch=B, bank=GaborBank.default(3, 6), colour_space=space, k=k, iters=T, max_gt=0,
                normalise=True, smooth=smooth)
    d_img = torch.from_numpy(imgs).cuda()
    feat = plan.gabor_features(d_img)
    f = feat.cpu().numpy().reshape(B, plan.D, -1)
    # the map from the standalone moments kernel: bit-exact against the oracle's integer-moment definition
    ab = plan.feature_affine(feat).cpu().numpy()
    for b in range(B):
        np.testing.assert_array_equal(ab[b].view(np.uint32), orc.feature_affine(f[b]).view(np.uint32))
    # whole segmenter (moments taken in the filter bank's / the smoothing's epilogue): labels bit-exact teacher-forced
    labels = plan.segment(d_img, torch.from_numpy(idx)).cpu().numpy().reshape(B, -1)
    lab2, cent2 = plan.kmeans(feat, torch.from_numpy(idx))           # caller-supplied features: same result
    np.testing.assert_array_equal(lab2.cpu().numpy().reshape(B, -1), labels)
    for b in range(B):
        ol, oc, _ = orc.kmeans(f[b], k, T, idx[b], affine=orc.feature_affine(f[b]))
        np.testing.assert_array_equal(labels[b], ol)
        np.testing.assert_array_equal(cent2.cpu().numpy()[b].view(np.uint32), oc.view(np.uint32))
        plain, _, _ = orc.kmeans(f[b], k, T, idx[b])
        assert (plain != ol).mean() > 1e-3                            # normalisation changes the clustering


def test_segmenter_slot_options_and_pipeline():
    """The drop-in callable and the batch pipeline accept the options; metrics of the labels stay oracle-exact."""
    from gabor_color_image_segmentation_b200 import Plan, gabor_kmeans_segment, finish_image
    from gabor_color_image_segmentation_b200.pipeline import evaluate_batch, init_indices_for
    from gabor_color_image_segmentation_b200.synth import synth_batch
    from oracle import oracle as orc
    B, H, W, G, k = 2, 96, 128, 3, 5
    imgs, gts = synth_batch(B, H, W, G)
    lab = gabor_kmeans_segment(imgs[0], n_clusters=k, n_iter=6, normalise=True, smooth=0.5)
    assert lab.shape == (H, W) and 0 <= lab.min() and lab.max() < k
    plan = Plan(H, W, max_batch=B, k=k, iters=6, max_gt=G, normalise=True, smooth=0.5)
    c = evaluate_batch(plan, imgs, gts, init_indices_for(range(B), H * W, k), want_labels=True)
    np.testing.assert_array_equal(c.labels[0], lab)                   # same seed convention (image index 0)
    for b in range(B):
        o = orc.label_counts(c.labels[b], list(gts[b]))
        got, want = finish_image(c, b), orc.finish_metrics(o)
        assert all(float(got[key]) == float(want[key]) for key in want)
