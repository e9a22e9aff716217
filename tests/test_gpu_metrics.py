"""GPU label-comparison kernels (through the C ABI) against the reference's golden vectors and
the CPU oracle.  Integers bit-exact; floats bit-equal (finished on the host in reference order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLOAT_KEYS = ["recall", "precision", "underseg", "undersegNP", "compactness", "density"]


def _voronoi(rng, H, W, R, base=0):
    from gabor_color_image_segmentation_b200.synth import voronoi_labels
    return voronoi_labels(rng, H, W, R, base)


def test_metrics_class_matches_reference_golden(golden):
    from gabor_color_image_segmentation_b200 import metrics
    for name in [str(n) for n in golden["names"]]:
        lb = golden[name + "/lb"]
        gts = list(golden[name + "/gt"])
        size = int(golden[name + "/size"])
        m = metrics(None, lb, gts)
        m.set_boundary_recall(size)
        m.set_boundary_precision(size)
        m.set_density()
        m.set_undersegmentation()
        m.set_compactness()
        d = m.get_metrics()
        want = golden[name + "/floats"]
        assert int(d["regions"]) == int(golden[name + "/regions"]), name
        for i, k in enumerate(FLOAT_KEYS):
            assert float(d[k]) == float(want[i]), (name, k, d[k], want[i])
        np.testing.assert_array_equal(m.perimeters, golden[name + "/perimeters"], err_msg=name)
        # attribute / dict types the reference exposes
        assert isinstance(d["recall"], float) and isinstance(d["precision"], float)
        assert isinstance(d["underseg"], np.float64) and isinstance(d["density"], np.float64)
        assert (m.nx, m.ny) == lb.shape


def test_img_truth_attribute(golden):
    from gabor_color_image_segmentation_b200 import metrics
    from oracle import oracle as orc
    lb = golden["small_2/lb"]; gts = list(golden["small_2/gt"])
    m = metrics(None, lb, gts)
    for got, t in zip(m.img_truth, gts):
        np.testing.assert_array_equal(got, orc.find_boundaries(t))


def test_counts_match_oracle_random_batches():
    from gabor_color_image_segmentation_b200 import label_counts_host
    from oracle import oracle as orc
    rng = np.random.default_rng(11)
    # (B, H, W, G, regions, gt regions, size): ragged tiles, tiles crossing 32x64 borders, k-means-like
    # and SLIC-like region counts (shared-memory and global histogram paths)
    cases = [(3, 33, 65, 2, 8, 6, 5), (2, 70, 130, 3, 40, 30, 5), (1, 321, 481, 5, 8, 50, 5),
             (1, 321, 481, 2, 300, 60, 5), (2, 64, 64, 1, 5, 4, 3), (2, 31, 63, 2, 5, 4, 7),
             (1, 1, 50, 1, 3, 3, 5), (1, 50, 1, 1, 3, 3, 5), (2, 40, 40, 2, 6, 5, 1), (1, 45, 77, 2, 9, 7, 9)]
    for (B, H, W, G, R, Rg, size) in cases:
        lbs = np.stack([_voronoi(rng, H, W, R) for _ in range(B)])
        gts = np.stack([np.stack([_voronoi(rng, H, W, Rg, base=1) for _ in range(G)]) for _ in range(B)]).astype(np.uint16)
        c = label_counts_host(lbs, gts, dil_recall=size, want_hist=True)
        for b in range(B):
            o = orc.label_counts(lbs[b], list(gts[b]), size)
            tag = (B, H, W, G, R, size, b)
            assert int(c.n_seg[b]) == o.n_seg, tag
            assert int(c.bd_count[b]) == o.bd_count, tag
            np.testing.assert_array_equal(c.gt_counts[b, :, 0], o.den_r, err_msg=str(tag))
            np.testing.assert_array_equal(c.gt_counts[b, :, 1], o.tp_r, err_msg=str(tag))
            np.testing.assert_array_equal(c.gt_counts[b, :, 2], o.tp_p, err_msg=str(tag))
            np.testing.assert_array_equal(c.gt_counts[b, :, 3], o.U, err_msg=str(tag))
            np.testing.assert_array_equal(c.gt_counts[b, :, 4], o.V, err_msg=str(tag))
            np.testing.assert_array_equal(c.area[b, :o.n_seg], o.area[:o.n_seg], err_msg=str(tag))
            np.testing.assert_array_equal(c.perim[b, :o.n_seg], o.perim[:o.n_seg], err_msg=str(tag))
            np.testing.assert_array_equal(c.n_lab[b], o.n_lab, err_msg=str(tag))
            for g in range(G):
                h = c.hist[b, g, :o.n_seg, :int(o.n_lab[g])]
                np.testing.assert_array_equal(h, o.hist[g], err_msg=str(tag))
                assert int(c.gt_counts[b, g, 5]) == int((o.hist[g].astype(np.int64) ** 2).sum())
                assert int(c.gt_counts[b, g, 6]) == int((o.hist[g].sum(0).astype(np.int64) ** 2).sum())


def test_ragged_ground_truth_counts():
    from gabor_color_image_segmentation_b200 import label_counts_host
    from oracle import oracle as orc
    rng = np.random.default_rng(12)
    B, H, W, G = 3, 48, 80, 4
    lbs = np.stack([_voronoi(rng, H, W, 7) for _ in range(B)])
    gts = np.stack([np.stack([_voronoi(rng, H, W, 6, base=1) for _ in range(G)]) for _ in range(B)]).astype(np.uint16)
    n_gt = np.array([4, 2, 1], np.int32)
    c = label_counts_host(lbs, gts, n_gt=n_gt)
    for b in range(B):
        o = orc.label_counts(lbs[b], list(gts[b, :n_gt[b]]))
        np.testing.assert_array_equal(c.gt_counts[b, :n_gt[b], :5],
                                      np.stack([o.den_r, o.tp_r, o.tp_p, o.U, o.V], 1))
        assert (c.gt_counts[b, n_gt[b]:] == 0).all()


def test_reference_error_behaviour():
    from gabor_color_image_segmentation_b200 import metrics
    lb1 = np.zeros((8, 8), np.int64)
    gt2 = np.tile(np.arange(8) // 4 + 1, (8, 1))
    m = metrics(None, lb1, [gt2])
    with pytest.raises(ZeroDivisionError):
        m.set_boundary_precision()                      # metrics.py:94
    lb2 = np.tile(np.arange(8) // 4, (8, 1))
    with pytest.raises(ZeroDivisionError):
        metrics(None, lb2, [np.ones((8, 8), np.int64)]).set_boundary_recall()   # metrics.py:72
    with pytest.raises(ZeroDivisionError):
        metrics(None, lb2, []).set_boundary_recall()    # metrics.py:74
    with pytest.raises(ZeroDivisionError):
        metrics(None, lb2, []).set_undersegmentation()  # metrics.py:145
    with pytest.raises(ValueError):
        metrics(None, lb2, [np.ones((8, 9), np.int64)]).set_boundary_recall()   # shape mismatch
    with pytest.raises(ValueError):
        metrics(None, lb2 - 1, [gt2]).set_metrics()     # negative labels rejected explicitly
    # float labels are truncated like astype('int') (metrics.py:43)
    a = metrics(None, lb2 + 0.7, [gt2]); a.set_metrics()
    b = metrics(None, lb2, [gt2]); b.set_metrics()
    assert a.get_metrics() == b.get_metrics()
    # precision ignores `size` (metrics.py:93), recall honours it (metrics.py:69)
    rng = np.random.default_rng(3)
    lb = _voronoi(rng, 40, 60, 6); gt = _voronoi(rng, 40, 60, 5, base=1)
    m1 = metrics(None, lb, [gt]); m1.set_boundary_precision(1); m1.set_boundary_recall(1)
    m5 = metrics(None, lb, [gt]); m5.set_boundary_precision(5); m5.set_boundary_recall(5)
    assert m1.precision == m5.precision and m1.recall < m5.recall


def test_label_capacity_is_enforced():
    from gabor_color_image_segmentation_b200 import label_counts_host
    lbs = np.zeros((1, 16, 16), np.int32); lbs[0, 3, 3] = 9
    gts = np.ones((1, 1, 16, 16), np.uint16)
    with pytest.raises(IndexError):
        label_counts_host(lbs, gts, n_seg_cap=4)
    gts[0, 0, 5, 5] = 70
    with pytest.raises(IndexError):
        label_counts_host(np.zeros((1, 16, 16), np.int32), gts, n_lab_cap=8)


def test_full_size_properties():
    """Size-independent properties at BASELINE's full image size (481x321, G=5)."""
    from gabor_color_image_segmentation_b200 import label_counts_host
    from gabor_color_image_segmentation_b200.synth import synth_ground_truths
    rng = np.random.default_rng(5)
    B, H, W, G = 4, 321, 481, 5
    lbs = np.stack([_voronoi(rng, H, W, 8) for _ in range(B)])
    gts = np.stack([synth_ground_truths(i, H, W, G) for i in range(B)])
    c = label_counts_host(lbs, gts, want_hist=True)
    N = H * W
    assert (c.area.sum(1) == N).all()
    assert (c.hist.sum((2, 3)) == N).all()                          # every pixel counted once per GT
    assert (c.hist.sum(3) == c.area[:, None, :]).all()              # row sums are the areas
    assert (c.gt_counts[..., 1] <= c.gt_counts[..., 0]).all()       # tp_r <= |bd(gt)|
    assert (c.gt_counts[..., 2] <= c.bd_count[:, None]).all()       # tp_p <= |bd(lb)|
    assert (c.gt_counts[..., 3] <= c.gt_counts[..., 4]).all()       # U <= V
    # a segmentation scored against itself: perfect recall/precision, zero undersegmentation
    same = label_counts_host(lbs, (lbs + 1).astype(np.uint16)[:, None], n_lab_cap=16)
    assert (same.gt_counts[:, 0, 0] == same.bd_count).all()
    assert (same.gt_counts[:, 0, 1] == same.bd_count).all() and (same.gt_counts[:, 0, 2] == same.bd_count).all()
    assert (same.gt_counts[:, 0, 3] == 0).all() and (same.gt_counts[:, 0, 4] == 0).all()
    # batch invariance
    one = label_counts_host(lbs[2:3], gts[2:3])
    np.testing.assert_array_equal(one.gt_counts[0], c.gt_counts[2])


def test_region_scores_from_gpu_tables_against_sklearn():
    """PRI / VoI / covering (beyond the reference, builder-defined) from the GPU contingency tables,
    checked against scikit-learn's pair-counting and information scores on the label maps."""
    from sklearn.metrics import mutual_info_score, rand_score
    from gabor_color_image_segmentation_b200 import label_counts_host
    from gabor_color_image_segmentation_b200.region_scores import region_scores
    rng = np.random.default_rng(17)
    B, H, W, G = 2, 90, 120, 3
    lbs = np.stack([_voronoi(rng, H, W, 7) for _ in range(B)])
    gts = np.stack([np.stack([_voronoi(rng, H, W, 6, base=1) for _ in range(G)]) for _ in range(B)]).astype(np.uint16)
    c = label_counts_host(lbs, gts, want_hist=True)
    s = region_scores(c.hist)
    for b in range(B):
        ri, voi, cov = [], [], []
        for g in range(G):
            x, y = lbs[b].ravel(), gts[b, g].ravel().astype(np.int64)
            ri.append(rand_score(y, x))
            hx = mutual_info_score(x, x); hy = mutual_info_score(y, y); mi = mutual_info_score(x, y)
            voi.append((hx + hy - 2 * mi) / np.log(2))
            acc = 0.0
            for j in np.unique(y):
                m = y == j
                best = max(((x == i) & m).sum() / ((x == i) | m).sum() for i in np.unique(x[m]))
                acc += m.sum() * best
            cov.append(acc / x.size)
        assert abs(s["pri"][b] - np.mean(ri)) < 1e-12
        assert abs(s["voi"][b] - np.mean(voi)) < 1e-9
        assert abs(s["covering"][b] - np.mean(cov)) < 1e-12
