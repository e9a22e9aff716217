"""Generate tests/golden/metrics_golden.npz by running the REFERENCE's own
``BSD_metrics/metrics.py`` (unmodified, imported from /root/reference) in the build
container.  TEST INFRASTRUCTURE.  The reference cannot travel to the GPU box, so the
vectors it produces are committed; re-run with

    python oracle/make_golden.py

scikit-image / matplotlib are absent from this image; the reference is imported over the
scipy-based stand-in in oracle/ref_shim — state that caveat with
every parity claim: "the reference's code over a scipy restatement of scikit-image".
"""
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/BSD_metrics"
OUT = os.path.join(HERE, "..", "tests", "golden", "metrics_golden.npz")


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, "ref_shim"))
    sys.path.insert(0, REF)
    import metrics as ref_metrics        # noqa: E402  (reference module)
    import groundtruth as ref_gt         # noqa: E402
    return ref_metrics, ref_gt


def voronoi(rng, H, W, R, base=0):
    ys = rng.integers(0, H, R); xs = rng.integers(0, W, R)
    yy, xx = np.mgrid[0:H, 0:W]
    d = (yy[..., None] - ys) ** 2 + (xx[..., None] - xs) ** 2
    lab = np.argmin(d, axis=-1)
    # relabel to contiguous 0..R'-1 in order of first appearance
    _, inv = np.unique(lab, return_inverse=True)
    return inv.reshape(H, W) + base


def run_case(ref_metrics, lb, gts, size=5):
    m = ref_metrics.metrics(None, lb, gts)
    m.set_boundary_recall(size)
    m.set_boundary_precision(size)
    m.set_density()
    m.set_undersegmentation()
    m.set_compactness()
    d = m.get_metrics()
    bd = ref_metrics.find_boundaries(m.lb)
    return {
        "regions": np.int64(d["regions"]),
        "floats": np.array([d["recall"], d["precision"], d["underseg"], d["undersegNP"],
                            d["compactness"], d["density"]], np.float64),
        "perimeters": np.asarray(m.perimeters, np.float64),
        "bd_count": np.int64(np.sum(bd)),
        "den_r": np.array([np.sum(t) for t in m.img_truth], np.int64),
        "tp_r": np.array([np.sum(ref_metrics.dilation(bd, ref_metrics.rectangle(size, size)) * t)
                          for t in m.img_truth], np.int64),
        "tp_p": np.array([np.sum(bd * ref_metrics.dilation(t, ref_metrics.rectangle(5, 5)))
                          for t in m.img_truth], np.int64),
    }


def main():
    ref_metrics, ref_gt = _import_reference()
    cwd = os.getcwd()
    os.chdir(REF)  # groundtruth.py:39 reads ./data/truth/
    store = {}
    names = []

    def add(name, lb, gts, size=5):
        r = run_case(ref_metrics, lb, gts, size)
        names.append(name)
        store[name + "/lb"] = np.asarray(lb)
        store[name + "/gt"] = np.stack(gts).astype(np.uint16)
        store[name + "/size"] = np.int64(size)
        for k, v in r.items():
            store[name + "/" + k] = v
        print(name, r["regions"], r["floats"])

    # (1) real BSDS500 ground truths with the formula label map
    for fid in ["2092", "100007", "3096", "33039"]:
        gts = ref_gt.get_segment_from_filename(fid)
        H, W = gts[0].shape
        yy, xx = np.mgrid[0:H, 0:W]
        add("bsds_%s_grid" % fid, (yy // 64) * 8 + (xx // 64), gts)
    # one real fixture with a k-means-like (8 regions) and a SLIC-like (300 regions) partition
    gts = ref_gt.get_segment_from_filename("3096")
    rng = np.random.default_rng(7)
    add("bsds_3096_vor8", voronoi(rng, 321, 481, 8), gts)
    add("bsds_3096_vor300", voronoi(rng, 321, 481, 300), gts)

    # (2) small seeded cases: ragged shapes, G = 1..4, dilation sizes 1/3/5/7
    rng = np.random.default_rng(2024)
    for i, (H, W, G, size) in enumerate([(20, 30, 1, 5), (20, 30, 1, 1), (37, 53, 3, 5), (64, 48, 4, 3),
                                         (33, 65, 2, 7), (16, 16, 2, 5), (50, 7, 2, 5), (9, 40, 3, 5)]):
        lb = voronoi(rng, H, W, int(rng.integers(2, 12)))
        gts = [voronoi(rng, H, W, int(rng.integers(2, 9)), base=1) for _ in range(G)]
        add("small_%d" % i, lb, gts, size)
    # tiny formula case of SURVEY.md Appendix B
    yy, xx = np.mgrid[0:20, 0:30]
    add("tiny_formula", yy // 10, [xx // 10 + 1])
    # 1-based and non-contiguous label maps
    lb = voronoi(rng, 40, 60, 6)
    gts = [voronoi(rng, 40, 60, 5, base=1) for _ in range(2)]
    add("one_based", lb + 1, gts)
    add("non_contiguous", lb * 3 + 2, gts)
    # pure noise labels (worst case for boundary counts)
    add("noise", rng.integers(0, 5, (24, 31)), [rng.integers(1, 4, (24, 31)) for _ in range(2)])

    os.chdir(cwd)
    store["names"] = np.array(names)
    buf = io.BytesIO()
    np.savez_compressed(buf, **store)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "wb") as f:
        f.write(buf.getvalue())
    print("wrote", OUT, len(buf.getvalue()), "bytes")


if __name__ == "__main__":
    main()
