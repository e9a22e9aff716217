"""Image decode, host side: the library's entropy decoder (no GPU needed) + the numpy restatement of libjpeg's
post-entropy stages (oracle/jpeg_oracle.py) against PIL, the decoder behind the reference's imread
(BSD_metrics/script.py:25): pixel-exact on the real BSDS500 fixture and on every sampling layout the decoder accepts."""
import hashlib
import io
import os

import numpy as np
import pytest
from PIL import Image

from gabor_color_image_segmentation_b200 import _lib, decode
from oracle import jpeg_oracle as jo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "bsds500", "images")


def _encode(arr, **kw):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG", **kw)
    return buf.getvalue()


def _synthetic(h, w, seed=0):
    from gabor_color_image_segmentation_b200.synth import synth_image
    return synth_image(seed, h, w)


def test_fixture_files_decode_to_pil_pixels():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "bsds500_golden.npz"))
    for fid in gold["ids"]:
        data = open(os.path.join(FIX, str(fid) + ".jpg"), "rb").read()
        want = np.asarray(Image.open(io.BytesIO(data)))
        assert decode.jpeg_info(data) == (want.shape[0], want.shape[1], 3)
        got = jo.decode_from_coefficients(data, decode.jpeg_coefficients(data))
        np.testing.assert_array_equal(got, want)
        assert hashlib.sha256(got.tobytes()).hexdigest() == str(gold[str(fid) + "/pixels_sha256"])


@pytest.mark.parametrize("shape,kw", [
    ((64, 80), dict(quality=90, subsampling=2)),      # 4:2:0, whole MCUs
    ((37, 53), dict(quality=75, subsampling=2)),      # 4:2:0, ragged edges in both directions
    ((41, 30), dict(quality=60, subsampling=1)),      # 4:2:2
    ((33, 47), dict(quality=95, subsampling=0)),      # 4:4:4
    ((17, 9), dict(quality=30, subsampling=2)),       # smaller than two MCUs, coarse quantisation
    ((50, 70), dict(quality=100, subsampling=2)),     # 16-bit-free but quantiser = 1
    ((40, 56), dict(quality=85, subsampling=2, optimize=True)),   # optimised Huffman tables
])
def test_sampling_layouts_against_pil(shape, kw):
    data = _encode(_synthetic(*shape, seed=shape[0]), **kw)
    want = np.asarray(Image.open(io.BytesIO(data)))
    got = jo.decode_from_coefficients(data, decode.jpeg_coefficients(data))
    np.testing.assert_array_equal(got, want)


def test_greyscale_and_restart_intervals():
    g = _synthetic(45, 61)[..., 1]
    data = _encode(g, quality=80)
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    np.testing.assert_array_equal(jo.decode_from_coefficients(data, decode.jpeg_coefficients(data)), want)
    try:
        data = _encode(_synthetic(70, 90), quality=80, subsampling=2, restart_marker_blocks=3)
    except TypeError:
        pytest.skip("this Pillow cannot write restart markers")
    if b"\xff\xdd" not in data:
        pytest.skip("this Pillow ignored restart_marker_blocks")
    want = np.asarray(Image.open(io.BytesIO(data)))
    np.testing.assert_array_equal(jo.decode_from_coefficients(data, decode.jpeg_coefficients(data)), want)


def test_unsupported_flavours_are_rejected_not_misdecoded():
    prog = _encode(_synthetic(40, 40), quality=80, progressive=True)
    with pytest.raises(_lib.GcisError):
        decode.jpeg_info(prog)
    with pytest.raises(_lib.GcisError):
        decode.jpeg_info(b"not a jpeg at all")
    cmyk = io.BytesIO()
    Image.fromarray(_synthetic(24, 24)).convert("CMYK").save(cmyk, format="JPEG")
    with pytest.raises(_lib.GcisError):
        decode.jpeg_info(cmyk.getvalue())
    data = _encode(_synthetic(40, 40), quality=80)
    with pytest.raises(_lib.GcisError):                 # truncated scan: the Huffman decoder must notice, not crash
        decode.jpeg_coefficients(data[:len(data) // 3])
