"""SLIC (the segmenter the reference calls, BSD_metrics/script.py:11,30) on the GPU against the oracle's restatement of
scikit-image's published algorithm: labels bit-exact (float64 arithmetic in the same order), plus the properties the
algorithm guarantees.  scikit-image itself is absent from the image: parity with a given release is unpinned."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape,n_segments,compactness,connect", [((321, 481), 300, 10.0, True), ((96, 128), 60, 10.0, True),
                                                                  ((120, 90), 100, 20.0, False), ((64, 64), 16, 1.0, True)])
def test_slic_labels_equal_the_oracle(shape, n_segments, compactness, connect):
    from gabor_color_image_segmentation_b200 import slic
    from gabor_color_image_segmentation_b200.synth import synth_image
    from oracle import oracle as orc
    for seed in (0, 1):
        img = synth_image(70 + seed, *shape)
        got = slic(img, n_segments=n_segments, compactness=compactness, enforce_connectivity=connect)
        want = orc.slic(img, n_segments, compactness, 10, connect, 1)
        np.testing.assert_array_equal(got, want)


def test_slic_on_a_real_image_through_the_reference_loop_body():
    """script.py:25-38 with the reference's own segmenter call: imread -> slic(img, n_segments=300, compactness=10.0)
    -> ground truths -> metrics; labels equal the oracle's, segments are 4-connected, metrics equal the oracle's."""
    from PIL import Image
    from scipy import ndimage as ndi
    from gabor_color_image_segmentation_b200 import get_segmentation, metrics, slic
    from oracle import oracle as orc
    fix = os.path.join(ROOT, "tests", "golden", "bsds500")
    img = np.asarray(Image.open(os.path.join(fix, "images", "3096.jpg")))
    labels = slic(img, n_segments=300, compactness=10.0)
    np.testing.assert_array_equal(labels, orc.slic(img, 300, 10.0))
    ids = np.unique(labels)
    assert 200 <= len(ids) <= 300 and ids.min() >= 0
    assert all(ndi.label(labels == v)[1] == 1 for v in ids)          # every superpixel is one connected component
    gts = get_segmentation(os.path.join(fix, "truth") + "/", "3096")
    m = metrics(img, labels, gts)
    m.set_metrics()
    want = orc.finish_metrics(orc.label_counts(labels, gts))
    got = m.get_metrics()
    assert all(float(got[k]) == float(want[k]) for k in want)
    assert 0.5 < got["recall"] <= 1.0                                  # superpixels hug the human boundaries
