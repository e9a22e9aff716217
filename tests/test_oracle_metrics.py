"""The CPU oracle's metrics stage against the golden vectors produced by the reference's own
BSD_metrics/metrics.py (oracle/make_golden.py).  Integers equal, floats bit-equal."""
import numpy as np
import pytest

from oracle import oracle as orc

FLOAT_KEYS = ["recall", "precision", "underseg", "undersegNP", "compactness", "density"]


def _cases(golden):
    return [str(n) for n in golden["names"]]


def test_golden_has_appendix_b_values(golden):
    # SURVEY.md Appendix B, copied by hand from the survey: pins the golden file itself.
    f = golden["bsds_2092_grid/floats"]
    assert repr(float(f[0])) == "0.15976118173849757"
    assert repr(float(f[1])) == "0.12170293695828396"
    assert repr(float(f[2])) == "0.1862895040465134"
    assert repr(float(f[3])) == "0.34630049583320616"
    assert repr(float(f[4])) == "0.8035023542808041"
    assert repr(float(f[5])) == "0.05935194720241449"
    assert repr(float(golden["bsds_33039_grid/floats"][4])) == "0.8035023542808043"
    assert list(golden["bsds_2092_grid/den_r"]) == [6944, 5178, 6631, 5351, 7628, 5033, 8001]
    assert list(golden["bsds_2092_grid/tp_r"]) == [1065, 946, 917, 1008, 1075, 849, 1168]
    assert list(golden["bsds_2092_grid/tp_p"]) == [1181, 1087, 1001, 1195, 1121, 991, 1231]
    assert int(golden["bsds_2092_grid/bd_count"]) == 9164


def test_oracle_matches_reference_on_every_golden_case(golden):
    for name in _cases(golden):
        lb = golden[name + "/lb"]
        gts = list(golden[name + "/gt"])
        size = int(golden[name + "/size"])
        c = orc.label_counts(lb, gts, size)
        assert c.n_seg == int(golden[name + "/regions"]), name
        assert c.bd_count == int(golden[name + "/bd_count"]), name
        np.testing.assert_array_equal(c.den_r, golden[name + "/den_r"], err_msg=name)
        np.testing.assert_array_equal(c.tp_r, golden[name + "/tp_r"], err_msg=name)
        np.testing.assert_array_equal(c.tp_p, golden[name + "/tp_p"], err_msg=name)
        np.testing.assert_array_equal(c.perim.astype(np.float64), golden[name + "/perimeters"], err_msg=name)
        got = orc.finish_metrics(c)
        want = golden[name + "/floats"]
        for i, k in enumerate(FLOAT_KEYS):
            assert float(got[k]) == float(want[i]), (name, k, got[k], want[i])


def test_appendix_b_integer_counts(golden):
    c = orc.label_counts(golden["bsds_100007_grid/lb"], list(golden["bsds_100007_grid/gt"]))
    assert list(c.U) == [16065, 18648, 37085, 24240, 35229]
    assert list(c.V) == [32130, 37296, 74170, 48480, 66336]
    c = orc.label_counts(golden["bsds_3096_grid/lb"], list(golden["bsds_3096_grid/gt"]))
    assert list(c.U) == [8149, 11874, 26759, 8081, 7940]
    assert list(c.V) == [16298, 23748, 53364, 16162, 15880]
    area = c.area.reshape(6, 8)
    assert (area[:5, :7] == 4096).all() and (area[:5, 7] == 2112).all()
    assert list(area[5]) == [64] * 7 + [33]
    perim = c.perim.reshape(6, 8)
    assert (perim[:5, :7] == 252).all() and (perim[:5, 7] == 190).all()
    assert list(perim[5]) == [64] * 7 + [33]


def test_boundaries_and_dilation_against_scipy():
    """Closed-form A.1/A.2 vs the scipy calls scikit-image makes."""
    from scipy import ndimage as ndi
    rng = np.random.default_rng(5)
    cross = ndi.generate_binary_structure(2, 1)
    for _ in range(40):
        H, W = int(rng.integers(1, 24)), int(rng.integers(1, 24))
        x = rng.integers(0, 4, (H, W)).astype(np.int64)
        want = ndi.grey_dilation(x, footprint=cross) != ndi.grey_erosion(x, footprint=cross)
        np.testing.assert_array_equal(orc.find_boundaries(x), want)
        b = rng.random((H, W)) < 0.15
        for size in (1, 2, 3, 4, 5, 6, 7):
            fp = np.ones((size, size), np.uint8)
            want = ndi.grey_dilation(b.astype(np.uint8), footprint=fp[::-1, ::-1]).astype(bool)
            np.testing.assert_array_equal(orc.dilate_square(b, size), want, err_msg=str(size))


def test_reference_error_behaviour():
    lb = np.zeros((8, 8), np.int64)                     # single region
    gt = np.tile(np.arange(8) // 4 + 1, (8, 1))
    with pytest.raises(ZeroDivisionError):
        orc.finish_metrics(orc.label_counts(lb, [gt]))  # metrics.py:94
    lb2 = np.tile(np.arange(8) // 4, (8, 1))
    with pytest.raises(ZeroDivisionError):
        orc.finish_metrics(orc.label_counts(lb2, [np.ones((8, 8), np.int64)]))  # metrics.py:72
    with pytest.raises(ZeroDivisionError):
        orc.finish_metrics(orc.label_counts(lb2, []))   # metrics.py:74
