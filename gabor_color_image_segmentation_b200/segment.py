"""The segmenter for the reference's ``slic`` slot (BSD_metrics/script.py:30):
``labels = gabor_kmeans_segment(img)`` -> H x W integer label map with labels 0..k-1.

Gabor bank -> per-pixel magnitude features -> k-means, all on the GPU."""
from __future__ import annotations

import numpy as np

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
