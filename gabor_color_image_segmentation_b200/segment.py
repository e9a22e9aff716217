"""The segmenter for the reference's ``slic`` slot (BSD_metrics/script.py:30):
``labels = gabor_kmeans_segment(img)`` -> H x W integer label map with labels 0..k-1.

Gabor bank -> per-pixel magnitude features -> k-means, all on the GPU."""
from __future__ import annotations

import numpy as np

from . import _lib
from .engine import GaborBank, Plan, kmeans_init_indices

_PLANS = {}


def _plan_for(H, W, bank, colour_space, feature, k, iters, max_batch=1, normalise=False, smooth=0.0):
    key = (H, W, bank, colour_space, feature, k, iters, max_batch, bool(normalise), float(smooth))
    if key not in _PLANS:
        _PLANS[key] = Plan(H, W, max_batch=max_batch, bank=bank, colour_space=colour_space, feature=feature,
                           k=k, iters=iters, max_gt=0, normalise=normalise, smooth=smooth)
    return _PLANS[key]


def gabor_kmeans_segment(img, n_clusters=8, n_iter=20, seed=0, bank: GaborBank = None, colour_space="rgb",
                         feature="magnitude", init_idx=None, normalise=False, smooth=0.0):
    """img: H x W x 3 uint8 (what ``skimage.io.imread`` returns) -> H x W int32 labels.
    ``smooth`` > 0 smooths every magnitude plane with a Gaussian of sigma = smooth * sigma_s; ``normalise`` clusters
    on per-feature z-scores."""
    import torch
    img = np.ascontiguousarray(img)
    if img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
        raise ValueError("img must be H x W x 3 uint8")
    H, W = img.shape[:2]
    bank = bank or GaborBank.default()
    plan = _plan_for(H, W, bank, colour_space, feature, int(n_clusters), int(n_iter), normalise=normalise, smooth=smooth)
    if init_idx is None:
        init_idx = kmeans_init_indices(H * W, int(n_clusters), seed)
    d_img = torch.from_numpy(img).cuda()[None]
    labels = plan.segment(d_img, torch.from_numpy(np.asarray(init_idx, np.int32))[None])
    return labels[0].cpu().numpy()


def slic(image, n_segments=100, compactness=10.0, max_num_iter=10, enforce_connectivity=True, start_label=1):
    """Stand-in for ``skimage.segmentation.slic`` at BSD_metrics/script.py:11,30 (same name, same leading keywords and
    defaults): H x W x 3 uint8 -> H x W integer labels from ``start_label``.  scikit-image's published algorithm
    (float64 CIELAB, regular-grid seeds, ``max_num_iter`` assignment / update rounds, connectivity enforcement) on the
    GPU; parity with a given scikit-image release is unpinned.  Options this build does not have
    (sigma, spacing, masks, slic_zero, non-RGB input) are not accepted."""
    img = np.ascontiguousarray(image)
    if img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
        raise ValueError("image must be H x W x 3 uint8")
    H, W = img.shape[:2]
    labels = np.empty((H, W), np.int32)
    _lib.check(_lib.load().gcis_slic_host(img.ctypes.data, H, W, int(n_segments), float(compactness), int(max_num_iter),
                                          int(bool(enforce_connectivity)), int(start_label), labels.ctypes.data), "gcis_slic_host")
    return labels
