"""Image decode for the path (BSD_metrics/script.py:25: ``img = imread(img_path + name)``): baseline JPEG files ->
uint8 RGB tensors in HBM, bit-identical to what PIL / libjpeg return.  Huffman decoding
runs on host threads inside libgcis.so, inverse DCT + chroma upsampling + colour conversion on the GPU."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib


def jpeg_info(data: bytes):
    """(H, W, components) of a JPEG file held in memory; raises GcisError for flavours the decoder rejects."""
    h, w, n = C.c_int32(), C.c_int32(), C.c_int32()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    _lib.check(_lib.load().gcis_jpeg_info(buf, len(data), C.byref(h), C.byref(w), C.byref(n)), "gcis_jpeg_info")
    return h.value, w.value, n.value


def jpeg_coefficients(data: bytes) -> np.ndarray:
    """Quantised DCT coefficients (int16, per component [blocks_h][blocks_w][64], natural order) - host only."""
    H, W, n = jpeg_info(data)
    cap = 4 * ((H + 15) // 16 * 16) * ((W + 15) // 16 * 16)
    out = np.zeros(cap, np.int16)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    cnt = _lib.check(_lib.load().gcis_jpeg_coefficients(buf, len(data), out.ctypes.data, cap), "gcis_jpeg_coefficients")
    return out[:cnt]


def decode_jpeg_batch(blobs: Sequence[bytes], threads: int = 0):
    """JPEG files (bytes) of ONE frame size -> CUDA uint8 tensor [B, H, W, 3] (the segmenter's input layout)."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.GcisError("no CUDA device: gabor_color_image_segmentation_b200 has no CPU path")
    B = len(blobs)
    if B < 1:
        raise ValueError("no files")
    H, W, _ = jpeg_info(blobs[0])
    keep = [(C.c_uint8 * len(b)).from_buffer_copy(b) for b in blobs]
    ptrs = (C.c_void_p * B)(*[C.addressof(k) for k in keep])
    sizes = (C.c_int64 * B)(*[len(b) for b in blobs])
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().gcis_jpeg_decode_batch(ptrs, sizes, B, H, W, out.data_ptr(), int(threads),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "gcis_jpeg_decode_batch")
    return out


def imread_gpu(path: str):
    """One JPEG file -> CUDA uint8 tensor [H, W, 3]."""
    with open(path, "rb") as f:
        return decode_jpeg_batch([f.read()])[0]
