"""Drop-in for the reference's ``BSD_metrics/groundtruth.py``: BSDS500 ``.mat`` ground truth
-> list of uint16 label maps (host I/O; the GPU path starts after this)."""
import os

import numpy as np
from scipy.io import loadmat

__all__ = ["get_segmentation", "get_segment_from_filename", "pack_ground_truths", "read_seg"]


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
            if s.size and (s.min() < 0 or s.max() > 65535):
                raise ValueError("image %d ground truth %d: labels outside 0..65535 would wrap in uint16" % (b, g))
            out[b, g] = s
    return out, n_gt


def read_seg(filename, one_based=True):
    """BSDS300 human segmentation (``data/Humans/{color,gray}/<user>/<image>.seg``, SURVEY.md section 2
    #7: ASCII header ending in a ``data`` line, then run-length rows ``label row col_start col_end``,
    columns inclusive) -> H x W uint16 label map.  No reference code reads these files; the decoder
    lets the same evaluation run against BSDS300 annotations.  ``one_based`` shifts the labels to
    1..segments, the convention of the ``.mat`` ground truths (column 0 of the contingency table empty)."""
    width = height = None
    with open(filename, "r") as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "width":
                width = int(tok[1])
            elif tok[0] == "height":
                height = int(tok[1])
            elif tok[0] == "format" and len(tok) > 1 and tok[1] != "ascii":
                raise ValueError("%s: unsupported .seg format %r" % (filename, " ".join(tok[1:])))
            elif tok[0] == "data":
                break
        if width is None or height is None:
            raise ValueError("%s: .seg header lacks width/height" % filename)
        runs = np.loadtxt(f, dtype=np.int64, ndmin=2)
    if runs.size == 0 or runs.shape[1] != 4:
        raise ValueError("%s: expected rows of 'label row col_start col_end'" % filename)
    lab, row, c0, c1 = runs.T
    if row.min() < 0 or row.max() >= height or c0.min() < 0 or c1.max() >= width or (c1 < c0).any() or lab.min() < 0:
        raise ValueError("%s: run outside the %dx%d image" % (filename, height, width))
    # expand the runs with one difference array per row: +label at col_start, -label after col_end
    out = np.zeros((height, width + 1), np.int64)
    covered = np.zeros((height, width + 1), np.int64)
    v = lab + (1 if one_based else 0)
    np.add.at(out, (row, c0), v)
    np.add.at(out, (row, c1 + 1), -v)
    np.add.at(covered, (row, c0), 1)
    np.add.at(covered, (row, c1 + 1), -1)
    covered = np.cumsum(covered, axis=1)[:, :width]
    if (covered != 1).any():
        raise ValueError("%s: runs overlap or leave %d pixels uncovered" % (filename, int((covered != 1).sum())))
    return np.cumsum(out, axis=1)[:, :width].astype(np.uint16)
