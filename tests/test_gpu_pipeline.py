"""Whole path through the host entry point of the C ABI (gcis_pipeline_host)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pipeline_host_matches_stagewise_and_oracle():
    import torch
    from gabor_color_image_segmentation_b200 import Plan, finish_image
    from gabor_color_image_segmentation_b200.pipeline import evaluate_batch, init_indices_for
    from gabor_color_image_segmentation_b200.synth import synth_batch
    from oracle import oracle as orc
    B, H, W, G, k, T = 5, 321, 481, 5, 8, 20
    imgs, gts = synth_batch(B, H, W, G)
    idx = init_indices_for(range(B), H * W, k)
    plan = Plan(H, W, max_batch=4, k=k, iters=T, max_gt=G, n_lab_cap=64)   # B > max_batch: chunked
    c = evaluate_batch(plan, imgs, gts, idx, want_labels=True)
    assert c.labels.shape == (B, H, W)
    # stage-wise device path gives the same labels
    lab2 = plan.segment(torch.from_numpy(imgs[:4]).cuda(), torch.from_numpy(idx[:4])).cpu().numpy()
    np.testing.assert_array_equal(lab2, c.labels[:4])
    for b in range(B):
        # metrics of the GPU's labels: integer-exact and float-bit-equal against the oracle
        o = orc.label_counts(c.labels[b], list(gts[b]))
        assert int(c.bd_count[b]) == o.bd_count
        np.testing.assert_array_equal(c.gt_counts[b, :, :5], np.stack([o.den_r, o.tp_r, o.tp_p, o.U, o.V], 1))
        np.testing.assert_array_equal(c.area[b, :o.n_seg], o.area)
        np.testing.assert_array_equal(c.perim[b, :o.n_seg], o.perim)
        want = orc.finish_metrics(o)
        got = finish_image(c, b)
        for key in want:
            assert float(got[key]) == float(want[key]), (b, key)
    # k-means teacher-forced on the GPU's own features at full size: bit-exact labels
    feat = plan.gabor_features(torch.from_numpy(imgs[:1]).cuda()).cpu().numpy()[0].reshape(72, -1)
    ol, _, _ = orc.kmeans(feat, k, T, idx[0])
    np.testing.assert_array_equal(c.labels[0].ravel(), ol)


def test_pipeline_independent_of_batching_and_grouping():
    from gabor_color_image_segmentation_b200 import Plan
    from gabor_color_image_segmentation_b200.pipeline import evaluate_batch, init_indices_for, records_to_array
    from gabor_color_image_segmentation_b200.synth import synth_batch
    B, H, W, G, k = 6, 120, 160, 3, 5
    imgs, gts = synth_batch(B, H, W, G)
    idx = init_indices_for(range(B), H * W, k)
    base = None
    for (mb, group) in [(6, 1), (6, 2), (6, 6), (4, 3), (1, 1)]:
        plan = Plan(H, W, max_batch=mb, k=k, iters=8, max_gt=G, group=group)
        rec = records_to_array(evaluate_batch(plan, imgs, gts, idx))
        if base is None:
            base = rec
        np.testing.assert_array_equal(rec, base)
        plan.close()


def test_pipeline_with_pinned_tensors_and_ragged_gt():
    import torch
    from gabor_color_image_segmentation_b200 import Plan
    from gabor_color_image_segmentation_b200.pipeline import init_indices_for
    from gabor_color_image_segmentation_b200.synth import synth_batch
    B, H, W, G, k = 3, 64, 96, 4, 4
    imgs, gts = synth_batch(B, H, W, G)
    idx = init_indices_for(range(B), H * W, k)
    plan = Plan(H, W, max_batch=B, k=k, iters=4, max_gt=G)
    n_gt = np.array([4, 1, 3], np.int32)
    a = plan.pipeline_host(imgs, gts, idx, B, n_gt)
    pin = lambda x: torch.from_numpy(x).pin_memory()
    b = plan.pipeline_host(pin(imgs), pin(gts.view(np.int16)), pin(idx), B, n_gt)
    np.testing.assert_array_equal(a.gt_counts, b.gt_counts)
    assert (a.gt_counts[1, 1:] == 0).all() and (a.gt_counts[2, 3:] == 0).all()


def test_mixed_shapes_and_region_scores():
    """Landscape + portrait images with ragged annotator counts through evaluate_mixed, and the
    region scores from the pipeline's contingency tables."""
    from gabor_color_image_segmentation_b200 import Plan, metrics, gabor_kmeans_segment
    from gabor_color_image_segmentation_b200.pipeline import evaluate_mixed, evaluate_batch, init_indices_for
    from gabor_color_image_segmentation_b200.region_scores import region_scores
    from gabor_color_image_segmentation_b200.synth import synth_image, synth_ground_truths
    shapes = [(96, 128), (128, 96), (96, 128)]
    imgs = [synth_image(i, h, w) for i, (h, w) in enumerate(shapes)]
    gts = [list(synth_ground_truths(i, h, w, 2 + i)) for i, (h, w) in enumerate(shapes)]
    res = evaluate_mixed(imgs, gts, k=5, iters=6)
    for i, (img, gt) in enumerate(zip(imgs, gts)):
        labels = gabor_kmeans_segment(img, n_clusters=5, n_iter=6, seed=i)
        m = metrics(img, labels, gt)
        m.set_metrics()
        want = m.get_metrics()
        for key in want:
            assert float(res[i][key]) == float(want[key]), (i, key)
    # region scores from the plan's device-side tables == from label_counts_host on the same labels
    from gabor_color_image_segmentation_b200 import label_counts_host
    H, W, G, k = 96, 128, 2, 5
    plan = Plan(H, W, max_batch=1, k=k, iters=6, max_gt=G, n_lab_cap=64)
    g2 = np.stack(gts[0])[None].astype(np.uint16)
    c = evaluate_batch(plan, imgs[0][None], g2, init_indices_for([0], H * W, k), want_labels=True)
    a = region_scores(plan.fetch_hist(1))
    b = region_scores(label_counts_host(c.labels, g2, n_seg_cap=k, n_lab_cap=64, want_hist=True).hist)
    for key in a:
        assert a[key][0] == b[key][0]
