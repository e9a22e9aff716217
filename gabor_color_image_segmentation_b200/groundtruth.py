"""Drop-in for the reference's ``BSD_metrics/groundtruth.py``: BSDS500 ``.mat`` ground truth
-> list of uint16 label maps (host I/O; the GPU path starts after this)."""
import os

import numpy as np
from scipy.io import loadmat

__all__ = ["get_segmentation", "get_segment_from_filename", "pack_ground_truths"]


def get_segmentation(path, filename):
    """Load groundtruth on BSD500 for the specified image (groundtruth.py:16-29):
    ``f['groundTruth'][0][i][0][0][0]`` is the i-th annotator's ``Segmentation``."""
    f = loadmat(path + filename)
    data = f['groundTruth'][0]
    groundtruth = []
    for img in data:
        groundtruth.append(img[0][0][0])
    return groundtruth


def get_segment_from_filename(filename, path="./data/truth/"):
    """All ground truths of image ``filename`` from every split folder under ``path``
    (groundtruth.py:33-50; the reference hard-codes ``./data/truth/``, kept as the default)."""
    list_dir = os.listdir(path)
    filename = filename + '.mat'
    segments = []
    for folder in list_dir:
        list_img = os.listdir(path + folder + "/")
        if filename in list_img:
            segments.extend(get_segmentation(path + folder + "/", filename))
    return segments


def pack_ground_truths(per_image_segments, max_gt=None):
    """Ragged list (images x annotators) of H x W maps -> ([B,G,H,W] uint16, n_gt [B] int32),
    the layout the batched GPU entry points take.  Unused slots are zero-filled."""
    B = len(per_image_segments)
    G = max_gt or max(len(s) for s in per_image_segments)
    H, W = np.asarray(per_image_segments[0][0]).shape
    out = np.zeros((B, G, H, W), np.uint16)
    n_gt = np.zeros(B, np.int32)
    for b, segs in enumerate(per_image_segments):
        if len(segs) > G:
            raise ValueError("image %d has %d ground truths, capacity %d" % (b, len(segs), G))
        n_gt[b] = len(segs)
        for g, s in enumerate(segs):
            s = np.asarray(s)
            if s.shape != (H, W):
                raise ValueError("image %d ground truth %d has shape %s, expected %s" % (b, g, s.shape, (H, W)))
            out[b, g] = s
    return out, n_gt
