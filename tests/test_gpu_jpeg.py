"""Image decode on the GPU: decode_jpeg_batch == PIL (the decoder behind the reference's
imread, BSD_metrics/script.py:25) pixel for pixel, on the real BSDS500 fixture and on every accepted sampling layout;
and the dataset driver gives identical results with the GPU decoder."""
import io
import os

import numpy as np
import pytest
from PIL import Image

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "bsds500")


def _encode(arr, **kw):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG", **kw)
    return buf.getvalue()


def test_fixture_batches_equal_pil():
    from gabor_color_image_segmentation_b200 import decode_jpeg_batch, imread_gpu
    files = sorted(os.listdir(os.path.join(FIX, "images")))
    blobs = {f: open(os.path.join(FIX, "images", f), "rb").read() for f in files}
    by_shape = {}
    for f in files:
        by_shape.setdefault(Image.open(io.BytesIO(blobs[f])).size, []).append(f)
    assert len(by_shape) == 2                                  # landscape and portrait
    for names in by_shape.values():
        got = decode_jpeg_batch([blobs[f] for f in names], threads=3).cpu().numpy()
        for n, f in enumerate(names):
            np.testing.assert_array_equal(got[n], np.asarray(Image.open(io.BytesIO(blobs[f]))), err_msg=f)
    one = imread_gpu(os.path.join(FIX, "images", files[0])).cpu().numpy()
    np.testing.assert_array_equal(one, np.asarray(Image.open(os.path.join(FIX, "images", files[0]))))


@pytest.mark.parametrize("shape,kw", [((64, 80), dict(quality=90, subsampling=2)), ((37, 53), dict(quality=75, subsampling=2)),
                                      ((41, 30), dict(quality=60, subsampling=1)), ((33, 47), dict(quality=95, subsampling=0)),
                                      ((17, 9), dict(quality=30, subsampling=2)), ((321, 481), dict(quality=100, subsampling=2)),
                                      ((45, 61), dict(quality=80, grey=True))])
def test_sampling_layouts_equal_pil(shape, kw):
    from gabor_color_image_segmentation_b200 import decode_jpeg_batch
    from gabor_color_image_segmentation_b200.synth import synth_image
    kw = dict(kw)
    grey = kw.pop("grey", False)
    blobs = []
    for i in range(3):
        img = synth_image(50 + i, *shape)
        blobs.append(_encode(img[..., 1] if grey else img, **kw))
    got = decode_jpeg_batch(blobs).cpu().numpy()
    for i, b in enumerate(blobs):
        np.testing.assert_array_equal(got[i], np.asarray(Image.open(io.BytesIO(b)).convert("RGB")))


def test_dataset_driver_with_gpu_decode_is_identical():
    from gabor_color_image_segmentation_b200 import evaluate_dataset
    a = evaluate_dataset(os.path.join(FIX, "images"), os.path.join(FIX, "truth"), k=8, iters=6, want_labels=True)
    b = evaluate_dataset(os.path.join(FIX, "images"), os.path.join(FIX, "truth"), k=8, iters=6, want_labels=True,
                         gpu_decode=True, chunk=3)
    assert [r[0] for r in a] == [r[0] for r in b]
    for ra, rb in zip(a, b):
        np.testing.assert_array_equal(ra[2], rb[2])
        assert all(float(ra[1][k]) == float(rb[1][k]) for k in ra[1])
