"""Batch driver: the reference's loop (BSD_metrics/script.py:22-38) over many images,
sharded by image over the GPUs of one box (SURVEY.md §8 e).

Each rank owns the images ``i`` with ``i % world == rank`` and runs the whole per-image path
locally; the only exchange is the reduction of the per-image metric sums (one NCCL
all-reduce of 7 float64) or, for the bit-exact check, an all-gather of the integer records."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import BatchCounts, Plan, kmeans_init_indices
from .metrics import METRIC_KEYS, finish_image

SUM_KEYS = METRIC_KEYS[1:]  # recall, precision, underseg, undersegNP, compactness, density


def shard_indices(n_images: int, rank: int, world: int) -> np.ndarray:
    """Image indices owned by ``rank`` (round-robin)."""
    return np.arange(rank, n_images, world, dtype=np.int64)


def init_indices_for(indices, n_pixels: int, k: int, seed: int = 0) -> np.ndarray:
    """[len(indices), k] initial-centroid pixels; seeded per GLOBAL image index so the result
    does not depend on the sharding."""
    return np.stack([kmeans_init_indices(n_pixels, k, seed + int(i)) for i in indices]) \
        if len(indices) else np.zeros((0, k), np.int32)


def records_to_array(c: BatchCounts) -> np.ndarray:
    """Integer record per image as one int64 row: [n_seg, n_gt, bd, area[k], perim[k], gt_counts[G*8]]."""
    B = len(c.bd_count)
    gc = np.asarray(c.gt_counts, np.int64)
    width = int(np.prod(gc.shape[1:]))      # explicit: reshape(B, -1) cannot infer a width for an empty shard
    return np.concatenate([c.n_seg.astype(np.int64).reshape(B, 1), c.n_gt.astype(np.int64).reshape(B, 1),
                           np.asarray(c.bd_count, np.int64).reshape(B, 1), c.area.astype(np.int64).reshape(B, -1) if B
                           else np.zeros((0, c.area.shape[1]), np.int64),
                           c.perim.astype(np.int64).reshape(B, -1) if B else np.zeros((0, c.perim.shape[1]), np.int64),
                           gc.reshape(B, width)], axis=1)


def array_to_records(a: np.ndarray, H: int, W: int, k: int, G: int) -> BatchCounts:
    B = a.shape[0]
    return BatchCounts(H, W, a[:, 2].copy(), a[:, 3 + 2 * k:].reshape(B, max(G, 1), -1).copy(),
                       a[:, 3:3 + k].astype(np.int32), a[:, 3 + k:3 + 2 * k].astype(np.int32),
                       a[:, 0].astype(np.int32), np.zeros((B, max(G, 1)), np.int32), np.zeros(B, np.int32),
                       a[:, 1].astype(np.int32))


def metric_sums(c: BatchCounts) -> np.ndarray:
    """[7] float64: sum over the batch of the six reference scores, and the image count."""
    acc = np.zeros(len(SUM_KEYS) + 1, np.float64)
    for b in range(len(c.bd_count)):
        m = finish_image(c, b)
        for i, key in enumerate(SUM_KEYS):
            acc[i] += float(m[key])
        acc[-1] += 1.0
    return acc


def reduce_sums(local: np.ndarray, device=None) -> np.ndarray:
    """Sum the 7-vector over ranks (NCCL on GPUs, gloo on CPU); identity without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    t = torch.from_numpy(local.copy())
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def gather_records(local: np.ndarray, indices: np.ndarray, n_images: int, device=None) -> Optional[np.ndarray]:
    """All-gather the integer records and put them back in global image order, so the floats
    finished from them are identical for 1/2/4/8 ranks.  Every rank returns the full table."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = np.zeros((n_images, local.shape[1]), np.int64)
        out[indices] = local
        return out
    world = dist.get_world_size()
    width = local.shape[1]
    per = (n_images + world - 1) // world
    pad = np.zeros((per, width + 1), np.int64)
    pad[:len(indices), 0] = indices + 1          # 0 marks padding
    pad[:len(indices), 1:] = local
    t = torch.from_numpy(pad)
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = np.zeros((n_images, width), np.int64)
    for p in parts:
        p = p.cpu().numpy()
        ok = p[:, 0] > 0
        out[p[ok, 0] - 1] = p[ok, 1:]
    return out


def evaluate_batch(plan: Plan, imgs: np.ndarray, gts: np.ndarray, init_idx: np.ndarray,
                   n_gt: Optional[np.ndarray] = None, want_labels: bool = False) -> BatchCounts:
    """Host arrays in, integer records out, through the C ABI's host entry point."""
    imgs = np.asarray(imgs)
    gts = np.asarray(gts)
    if imgs.dtype != np.uint8:
        raise ValueError("imgs must be uint8 [B,H,W,3] (as imread returns them)")
    if gts.size and (gts.min() < 0 or gts.max() > 65535):
        raise ValueError("ground-truth labels must lie in 0..65535 (uint16, the dtype groundtruth.py:26 yields)")
    if gts.size and int(gts.max()) + 1 > plan.n_lab_cap:
        raise ValueError(f"a ground truth has {int(gts.max()) + 1} labels (max+1) but the plan was built with "
                         f"n_lab_cap={plan.n_lab_cap}; build the plan with n_lab_cap >= {int(gts.max()) + 1}")
    imgs = np.ascontiguousarray(imgs)
    gts = np.ascontiguousarray(gts, np.uint16)
    init_idx = np.ascontiguousarray(init_idx, np.int32)
    return plan.pipeline_host(imgs, gts, init_idx, imgs.shape[0], n_gt, want_labels)


def empty_counts(plan: Plan) -> BatchCounts:
    """Record set of an empty shard (a rank that owns no image): every collective still sees the same widths."""
    G = max(plan.max_gt, 1)
    z = lambda *s, dt=np.int32: np.zeros(s, dt)
    return BatchCounts(plan.H, plan.W, z(0, dt=np.int64), z(0, G, 8, dt=np.int64), z(0, plan.k), z(0, plan.k), z(0),
                       z(0, G), z(0), z(0))


def dataset_scores(sums: np.ndarray) -> dict:
    """Mean of each per-image score over the dataset (builder-defined aggregation; the
    reference aggregates nothing across images, script.py:22-38)."""
    n = sums[-1]
    return {k: float(sums[i] / n) for i, k in enumerate(SUM_KEYS)} | {"images": int(n)}


def evaluate_mixed(images, ground_truths, k: int = 8, iters: int = 20, seed: int = 0, plans: Optional[dict] = None,
                   **plan_kwargs):
    """The reference's driver loop (script.py:22-38) over a list of images of MIXED shapes — BSDS500
    holds both 321x481 and 481x321 images and 4..9 annotators per image.  Images are grouped by
    shape, each group runs through one plan, and the per-image metric dicts come back in input
    order.  `ground_truths[i]` is the list `get_segment_from_filename` returns for image i."""
    from .groundtruth import pack_ground_truths
    from .metrics import finish_image
    plans = {} if plans is None else plans
    by_shape = {}
    for i, img in enumerate(images):
        by_shape.setdefault(tuple(np.asarray(img).shape[:2]), []).append(i)
    out = [None] * len(images)
    for (H, W), idxs in by_shape.items():
        G = max(len(ground_truths[i]) for i in idxs)
        n_lab = max(int(np.max(g)) for i in idxs for g in ground_truths[i]) + 1
        key = (H, W, G, k, iters)
        if key not in plans or plans[key].n_lab_cap < n_lab or plans[key].max_batch < len(idxs):
            plans[key] = Plan(H, W, max_batch=len(idxs), k=k, iters=iters, max_gt=G,
                              n_lab_cap=max(64, n_lab), **plan_kwargs)
        plan = plans[key]
        imgs = np.stack([np.asarray(images[i], np.uint8) for i in idxs])
        gts, n_gt = pack_ground_truths([ground_truths[i] for i in idxs], max_gt=G)
        init = init_indices_for(idxs, H * W, k, seed)
        c = evaluate_batch(plan, imgs, gts, init, n_gt)
        for n, i in enumerate(idxs):
            out[i] = finish_image(c, n)
    return out
