"""Generate the real-data fixture: tests/golden/bsds500/ + tests/golden/bsds500_golden.npz.
TEST INFRASTRUCTURE; run in the build container (it reads /root/reference):

    python oracle/make_real_fixture.py

What it commits
  tests/golden/bsds500/images/<id>.jpg, truth/<id>.mat
      eight BSDS500 files copied byte for byte from the reference checkout's DATA directory
      (BSD_metrics/data/Berkeley/*/, BSD_metrics/data/truth/*/).  They are data, not reference
      source.  Attribution: Berkeley Segmentation Data Set and Benchmarks 500 (BSDS500),
      P. Arbelaez, M. Maire, C. Fowlkes and J. Malik, "Contour Detection and Hierarchical Image
      Segmentation", IEEE TPAMI 33(5), 2011; images and human annotations distributed by the
      Berkeley Computer Vision Group for non-commercial research and educational use.
  tests/golden/bsds500_golden.npz, per image id:
      <id>/pixels_sha256   sha256 of the H x W x 3 uint8 array PIL (libjpeg-turbo) decodes: what script.py:25's
                           imread returns; the fixture test re-decodes and compares, and a future in-repo
                           decoder must reproduce it
      <id>/labels          the ORACLE's segmentation of the decoded image (default bank 4x6, rgb, k=8, T=20,
                           init = default_rng(index).choice), uint8 [H][W]
      <id>/floats, regions, perimeters, bd_count, den_r, tp_r, tp_p
                           outputs of the REFERENCE's own BSD_metrics/metrics.py (imported unmodified over
                           oracle/ref_shim, see make_golden.py) for (labels, real ground truths): the
                           whole script.py:30-38 loop body with the oracle in the segmenter slot.
Caveat carried with every parity claim: "the reference's code over a scipy restatement of scikit-image".
"""
import hashlib
import io
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference/BSD_metrics"
DST = os.path.join(ROOT, "tests", "golden", "bsds500")
OUT = os.path.join(ROOT, "tests", "golden", "bsds500_golden.npz")

# (id, split): landscape and one portrait image, 5..7 annotators, the four SURVEY Appendix-B ids first
IDS = [("2092", "train"), ("100007", "test"), ("3096", "val"), ("33039", "val"),
       ("8068", "test"), ("12003", "train"), ("35028", "test"), ("41006", "test")]
K, T = 8, 20


def main():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    _import_reference, run_case = mg._import_reference, mg.run_case
    from oracle import oracle as orc
    from PIL import Image
    ref_metrics, ref_gt = _import_reference()
    os.makedirs(os.path.join(DST, "images"), exist_ok=True)
    os.makedirs(os.path.join(DST, "truth"), exist_ok=True)
    store = {"ids": np.array([i for i, _ in IDS]), "k": np.int64(K), "iters": np.int64(T)}
    for index, (fid, split) in enumerate(IDS):
        jpg = os.path.join(REF, "data", "Berkeley", split, fid + ".jpg")
        mat = os.path.join(REF, "data", "truth", split, fid + ".mat")
        shutil.copyfile(jpg, os.path.join(DST, "images", fid + ".jpg"))
        shutil.copyfile(mat, os.path.join(DST, "truth", fid + ".mat"))
        img = np.asarray(Image.open(jpg))                                  # script.py:25
        gts = ref_gt.get_segmentation(os.path.join(REF, "data", "truth", split) + "/", fid)   # groundtruth.py:16
        H, W = img.shape[:2]
        idx = orc.kmeans_init_indices(H * W, K, index)
        labels, _, _ = orc.segment_image(img, K, T, init_idx=idx)         # the segmenter slot, script.py:30
        r = run_case(ref_metrics, labels, gts)                            # script.py:36-37
        store[fid + "/pixels_sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest())
        store[fid + "/labels"] = labels.astype(np.uint8)
        store[fid + "/n_gt"] = np.int64(len(gts))
        for k, v in r.items():
            store[fid + "/" + k] = v
        print(fid, img.shape, len(gts), r["regions"], r["floats"])
    buf = io.BytesIO()
    np.savez_compressed(buf, **store)
    with open(OUT, "wb") as f:
        f.write(buf.getvalue())
    print("wrote", OUT, len(buf.getvalue()), "bytes")


if __name__ == "__main__":
    main()
