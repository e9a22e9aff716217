"""CPU-side tests: the C-ABI library loads and exports what include/gcis.h declares, the host
mirror of the reference interface finishes floats exactly like the reference, loaders and the
synthetic generator behave, and the product never reaches for the oracle."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gabor_color_image_segmentation_b200")


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gcis.h")).read()
    return sorted(set(re.findall(r"GCIS_API[^;(]*?\b(gcis_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gabor_color_image_segmentation_b200 import _lib
    _lib.build()
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.gcis_version() == 101


def test_sass_is_sm100a():
    import subprocess
    from gabor_color_image_segmentation_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback_without_device():
    from gabor_color_image_segmentation_b200 import _lib, GaborBank
    lib = _lib.load()
    if lib.gcis_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from gabor_color_image_segmentation_b200 import Plan, label_counts_host, metrics
    with pytest.raises(_lib.GcisError):
        Plan(32, 32)
    with pytest.raises(_lib.GcisError):
        label_counts_host(np.zeros((1, 8, 8), np.int32), np.ones((1, 1, 8, 8), np.uint16))
    with pytest.raises(_lib.GcisError):
        metrics(None, np.zeros((8, 8), int), [np.ones((8, 8), int)]).set_metrics()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/gcis_oracle.c:orc_kmeans", ""), os.path.join(dirpath, f)


def test_bank_description_matches_oracle():
    from gabor_color_image_segmentation_b200 import GaborBank
    from oracle import oracle as orc
    for bank, obank in [(GaborBank.default(), orc.Bank.default()), (GaborBank.dense(), orc.Bank.dense())]:
        assert bank.frequencies == obank.frequencies and bank.thetas == obank.thetas
        for s in range(len(bank.frequencies)):
            for o in range(len(bank.thetas)):
                gx, gy = bank.separable(s, o)
                ox, oy = orc.gabor_separable(obank.frequencies[s], obank.thetas[o])
                assert len(gx) == len(ox) == 2 * bank.half_width(s, o) + 1
                np.testing.assert_allclose(gx, ox, rtol=0, atol=1e-15)
                np.testing.assert_allclose(gy, oy, rtol=0, atol=1e-15)


def test_host_float_finishing_is_bit_equal_to_reference(golden):
    """Product host code (metrics.finish_image) on the oracle's integer counts vs the golden floats."""
    from gabor_color_image_segmentation_b200.engine import BatchCounts
    from gabor_color_image_segmentation_b200.metrics import finish_image
    from oracle import oracle as orc
    keys = ["recall", "precision", "underseg", "undersegNP", "compactness", "density"]
    for name in [str(n) for n in golden["names"]]:
        o = orc.label_counts(golden[name + "/lb"], list(golden[name + "/gt"]), int(golden[name + "/size"]))
        G = len(o.den_r)
        gc = np.zeros((1, G, 8), np.int64)
        gc[0, :, 0], gc[0, :, 1], gc[0, :, 2], gc[0, :, 3], gc[0, :, 4] = o.den_r, o.tp_r, o.tp_p, o.U, o.V
        c = BatchCounts(o.H, o.W, np.array([o.bd_count], np.int64), gc, o.area[None].astype(np.int32),
                        o.perim[None].astype(np.int32), np.array([o.n_seg], np.int32), o.n_lab[None],
                        np.zeros(1, np.int32), np.array([G], np.int32))
        got = finish_image(c, 0)
        for i, k in enumerate(keys):
            assert float(got[k]) == float(golden[name + "/floats"][i]), (name, k)
        assert int(got["regions"]) == int(golden[name + "/regions"])


def test_groundtruth_loader_and_packing(tmp_path):
    from scipy.io import savemat
    from gabor_color_image_segmentation_b200 import get_segmentation, get_segment_from_filename, pack_ground_truths
    rng = np.random.default_rng(0)
    segs = [rng.integers(1, 9, (12, 17)).astype(np.uint16) for _ in range(3)]
    gt = np.empty((1, 3), object)
    for i, s in enumerate(segs):
        inner = np.empty((1, 1), dtype=[("Segmentation", object), ("Boundaries", object)])
        inner[0, 0] = (s, (s > 4).astype(np.uint8))
        gt[0, i] = inner
    for split in ("train", "val"):
        os.makedirs(tmp_path / "truth" / split)
    savemat(str(tmp_path / "truth" / "val" / "4242.mat"), {"groundTruth": gt})
    got = get_segmentation(str(tmp_path / "truth" / "val") + "/", "4242.mat")
    assert len(got) == 3 and got[0].dtype == np.uint16
    for a, b in zip(got, segs):
        np.testing.assert_array_equal(a, b)
    got2 = get_segment_from_filename("4242", path=str(tmp_path / "truth") + "/")
    assert len(got2) == 3
    assert get_segment_from_filename("nope", path=str(tmp_path / "truth") + "/") == []
    packed, n_gt = pack_ground_truths([got, got[:1]], max_gt=4)
    assert packed.shape == (2, 4, 12, 17) and packed.dtype == np.uint16 and n_gt.tolist() == [3, 1]
    np.testing.assert_array_equal(packed[0, 2], segs[2])
    assert (packed[1, 1:] == 0).all()
    with pytest.raises(ValueError):
        pack_ground_truths([got], max_gt=2)


def _write_seg(path, lab, zero_based=True):
    """Run-length encode a label map the way the BSDS300 .seg files do."""
    H, W = lab.shape
    rows = []
    for r in range(H):
        c = 0
        while c < W:
            e = c
            while e + 1 < W and lab[r, e + 1] == lab[r, c]:
                e += 1
            rows.append("%d %d %d %d" % (lab[r, c], r, c, e))
            c = e + 1
    hdr = ["format ascii cr", "date Thu Jul 26 14:34:31 2001", "image 1", "user 1", "width %d" % W, "height %d" % H,
           "segments %d" % (lab.max() + 1), "gray 0", "invert 0", "flipflop 0", "data"]
    path.write_text("\n".join(hdr + rows) + "\n")


def test_seg_reader_round_trip_and_errors(tmp_path):
    from gabor_color_image_segmentation_b200.groundtruth import read_seg
    from gabor_color_image_segmentation_b200.synth import synth_ground_truths
    lab = synth_ground_truths(5, 37, 53, 1)[0].astype(np.int64) - 1      # 0-based like the .seg files
    f = tmp_path / "a.seg"
    _write_seg(f, lab)
    got = read_seg(str(f))
    assert got.dtype == np.uint16 and got.shape == lab.shape
    np.testing.assert_array_equal(got, lab + 1)                          # .mat convention: labels 1..R
    np.testing.assert_array_equal(read_seg(str(f), one_based=False), lab)
    # a hole in the coverage and a run outside the image are rejected
    txt = f.read_text().splitlines()
    (tmp_path / "hole.seg").write_text("\n".join(txt[:-1]) + "\n")
    with pytest.raises(ValueError):
        read_seg(str(tmp_path / "hole.seg"))
    (tmp_path / "oob.seg").write_text("\n".join(txt + ["0 99 0 3"]) + "\n")
    with pytest.raises(ValueError):
        read_seg(str(tmp_path / "oob.seg"))


def test_synthetic_generator_is_deterministic_and_bsds_shaped():
    from gabor_color_image_segmentation_b200 import synth
    a, g = synth.synth_image(7), synth.synth_ground_truths(7)
    assert a.shape == (321, 481, 3) and a.dtype == np.uint8
    assert g.shape == (5, 321, 481) and g.dtype == np.uint16
    np.testing.assert_array_equal(a, synth.synth_image(7))
    np.testing.assert_array_equal(g, synth.synth_ground_truths(7))
    assert not np.array_equal(a, synth.synth_image(8))
    for t in g:
        labs = np.unique(t)
        assert labs[0] == 1 and labs[-1] == len(labs) and len(labs) >= 2      # 1..R contiguous


def test_kmeans_init_indices_match_spec():
    from gabor_color_image_segmentation_b200 import kmeans_init_indices
    from oracle import oracle as orc
    a = kmeans_init_indices(154401, 8, 5)
    np.testing.assert_array_equal(a, np.random.default_rng(5).choice(154401, 8, replace=False))
    np.testing.assert_array_equal(a, orc.kmeans_init_indices(154401, 8, 5))
    assert len(set(a.tolist())) == 8
