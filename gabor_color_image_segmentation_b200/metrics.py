"""Drop-in for the reference's ``BSD_metrics/metrics.py``: same class name, constructor,
methods, attributes, dict keys and error behaviour; the pixel work runs on the GPU.

The GPU returns the integer counts of SURVEY.md A.3–A.7; the floats are finished here on
the host in the reference's own evaluation order, which makes them bit-identical to the
reference's (metrics.py:70-74, 89-96, 129-146, 157, 192-201).
"""
from __future__ import annotations

from math import pi

import numpy as np

from . import _lib
from .engine import BatchCounts, label_counts_host

__all__ = ["metrics", "finish_image", "METRIC_KEYS"]

METRIC_KEYS = ("regions", "recall", "precision", "underseg", "undersegNP", "compactness", "density")


def _recall(c: BatchCounts, b: int):
    G = int(c.n_gt[b])
    recall = 0
    for g in range(G):                                        # metrics.py:71-72
        recall += float(int(c.gt_counts[b, g, 1])) / float(int(c.gt_counts[b, g, 0]))
    recall /= G                                               # metrics.py:74
    return recall


def _precision(c: BatchCounts, b: int):
    G = int(c.n_gt[b])
    precision = 0
    global_score = float(int(c.bd_count[b]))                  # metrics.py:90
    for g in range(G):                                        # metrics.py:92-94
        precision += float(int(c.gt_counts[b, g, 2])) / global_score
    precision /= G                                            # metrics.py:96
    return precision


def _underseg(c: BatchCounts, b: int):
    G = int(c.n_gt[b])
    N = c.H * c.W
    und = 0.
    und_np = 0.
    for g in range(G):                                        # metrics.py:129-143
        u = np.float64(int(c.gt_counts[b, g, 3])) + 0.
        u /= N
        und += u
        unp = np.float64(int(c.gt_counts[b, g, 4])) + 0.
        unp /= N
        und_np += unp
    und /= G                                                  # metrics.py:145-146
    und_np /= G
    return und, und_np


def _density(c: BatchCounts, b: int):
    return np.int64(c.bd_count[b]) / float(c.H * c.W)         # metrics.py:157


def _compactness(c: BatchCounts, b: int, n_segments: int):
    compactness = 0
    max_area = float(c.H * c.W)
    for i in range(n_segments):                               # metrics.py:195-201
        area = np.int64(c.area[b, i])
        perimeter = np.float64(c.perim[b, i])
        ratio = area / max_area
        if perimeter > 0:
            compactness += 4 * pi * ratio * area / pow(perimeter, 2)
    return compactness


def finish_image(c: BatchCounts, b: int) -> dict:
    """All seven reference outputs of image ``b`` from its integer record (get_metrics keys)."""
    n_segments = np.int64(c.n_seg[b])
    und, und_np = _underseg(c, b)
    return {"regions": n_segments, "recall": _recall(c, b), "precision": _precision(c, b),
            "underseg": und, "undersegNP": und_np, "compactness": _compactness(c, b, int(n_segments)),
            "density": _density(c, b)}


class metrics:

    """
    Compute the metrics associated to a segmentation for the
    Berkeley Segmentation Dataset (GPU-backed; reference: metrics.py:18-255)
    """

    def __init__(self, img, lb, segments_truth):
        self.img = img
        self.lb = np.asarray(lb).astype('int')                # metrics.py:43 (truncates floats)
        self.nx, self.ny = self.lb.shape
        self.segments_truth = segments_truth
        self.n_segments = np.max(self.lb) + 1                 # metrics.py:51
        self._cache = {}
        self._img_truth = None

    # GT boundary maps (metrics.py:47-49) are only materialised if a caller asks for them.
    @property
    def img_truth(self):
        if self._img_truth is None:
            self._img_truth = [find_boundaries(np.asarray(t)) for t in self.segments_truth]
        return self._img_truth

    def _counts(self, size=5) -> BatchCounts:
        if size not in self._cache:
            if self.lb.min() < 0:
                raise ValueError("negative labels are not supported (the reference silently wraps them)")
            for t in self.segments_truth:                     # numpy ValueError at metrics.py:72,93
                if np.asarray(t).shape != self.lb.shape:
                    raise ValueError("operands could not be broadcast together with shapes %s %s"
                                     % (self.lb.shape, np.asarray(t).shape))
            G = len(self.segments_truth)
            for t in self.segments_truth:
                t = np.asarray(t)
                if t.size and (t.min() < 0 or t.max() > 65535):
                    raise ValueError("ground-truth labels must lie in 0..65535 (uint16, the dtype groundtruth.py:26 yields)")
            gts = (np.stack([np.asarray(t) for t in self.segments_truth]).astype(np.uint16)[None]
                   if G else np.zeros((1, 0, self.nx, self.ny), np.uint16))
            self._cache[size] = label_counts_host(self.lb[None].astype(np.int32), gts, dil_recall=int(size),
                                                  n_seg_cap=int(self.n_segments))
        return self._cache[size]

    def set_boundary_recall(self, size=5):
        self.recall = _recall(self._counts(size), 0)

    def set_boundary_precision(self, size=5):
        # the reference ignores `size` here: rectangle(5, 5) is hard-coded (metrics.py:93)
        self.precision = _precision(self._counts(5), 0)

    def set_undersegmentation(self):
        self.undersegmentation, self.undersegmentationNP = _underseg(self._counts(5), 0)

    def set_density(self):
        self.density = _density(self._counts(5), 0)

    def perimeter(self):
        self.perimeters = self._counts(5).perim[0, :int(self.n_segments)].astype(np.float64)

    def set_compactness(self):
        self.perimeter()
        self.compactness = _compactness(self._counts(5), 0, int(self.n_segments))

    def set_metrics(self):
        self.set_boundary_recall()
        self.set_boundary_precision()
        self.set_density()
        self.set_undersegmentation()
        self.set_compactness()

    def display_metrics(self):
        print("Regions: " + str(self.n_segments) +
              " Recall: " + str(self.recall) +
              " Precision: " + str(self.precision) +
              " Undersegmentation: " + str(self.undersegmentation) +
              " Undersegmentation (NP) " + str(self.undersegmentationNP) +
              " Compactness " + str(self.compactness) +
              " Density " + str(self.density))

    def get_metrics(self):
        return {"regions": self.n_segments,
                "recall": self.recall,
                "precision": self.precision,
                "underseg": self.undersegmentation,
                "undersegNP": self.undersegmentationNP,
                "compactness": self.compactness,
                "density": self.density}


def find_boundaries(x: np.ndarray) -> np.ndarray:
    """skimage ``find_boundaries`` defaults on the GPU, for ``img_truth``."""
    x32 = np.ascontiguousarray(x, np.int32)
    out = np.empty(x32.shape, np.uint8)
    _lib.check(_lib.load().gcis_find_boundaries_host(x32.ctypes.data, 1, x32.shape[0], x32.shape[1],
                                                     out.ctypes.data), "gcis_find_boundaries_host")
    return out.astype(bool)


def finish_batch(c: BatchCounts) -> dict:
    """Vectorised ``finish_image`` over a whole batch: arrays [B] per key.  Every image sees the
    same IEEE operations in the same order as the scalar replay (loops over g and i stay
    sequential; only the batch axis is vectorised), so results are bit-equal to it.  Requires a
    uniform ground-truth count; images that would raise ZeroDivisionError yield inf/nan here."""
    B = len(c.bd_count)
    G = int(c.n_gt[0]) if B else 0
    if B and not (c.n_gt == G).all():
        raise ValueError("finish_batch needs the same number of ground truths for every image")
    N = c.H * c.W
    gc = c.gt_counts
    with np.errstate(divide="ignore", invalid="ignore"):
        recall = np.zeros(B); precision = np.zeros(B); und = np.zeros(B); und_np = np.zeros(B)
        gs = c.bd_count.astype(np.float64)
        for g in range(G):
            recall = recall + gc[:, g, 1].astype(np.float64) / gc[:, g, 0].astype(np.float64)
            precision = precision + gc[:, g, 2].astype(np.float64) / gs
            und = und + gc[:, g, 3].astype(np.float64) / N
            und_np = und_np + gc[:, g, 4].astype(np.float64) / N
        recall = recall / G; precision = precision / G; und = und / G; und_np = und_np / G
        density = c.bd_count / float(N)
        compact = np.zeros(B)
        max_area = float(N)
        for i in range(c.area.shape[1]):
            area = c.area[:, i].astype(np.int64)
            per = c.perim[:, i].astype(np.float64)
            term = 4 * pi * (area / max_area) * area / (per * per)
            compact = compact + np.where((per > 0) & (i < c.n_seg), term, 0.0)
    return {"regions": c.n_seg.astype(np.int64), "recall": recall, "precision": precision, "underseg": und,
            "undersegNP": und_np, "compactness": compact, "density": density}
