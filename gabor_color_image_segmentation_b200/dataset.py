"""The reference's driver (BSD_metrics/script.py:19-38) over a whole directory: for every image, decode it,
segment it, load its human ground truths and score the segmentation; here as ONE batched GPU pass per image
shape instead of a per-image loop.

    results = evaluate_dataset("data/Berkeley/train/", "data/truth/train/")
    print_like_script(results)          # the two lines script.py:27,38 print per image

Host side: JPEG decode (script.py:25, ``imread``) and ``.mat`` parsing
(groundtruth.py:16-32) run in a thread pool straight into PINNED staging buffers, one [n,H,W,3] uint8 and one
[n,G,H,W] uint16 tensor per image shape (BSDS500 mixes 321x481 and 481x321 and 4..9 annotators), which
``gcis_pipeline_host`` then streams to the device while the kernels of the previous chunk run."""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence

import numpy as np

from .engine import Plan
from .groundtruth import get_segmentation
from .metrics import finish_image
from .pipeline import init_indices_for

IMAGE_EXT = (".jpg", ".jpeg", ".png", ".bmp")


def _find_truth(truth_dir: str, name: str) -> str:
    """Directory (with trailing slash) that holds ``<name>.mat``: truth_dir itself or one of its split
    sub-directories (groundtruth.py:39-48 scans ./data/truth/*/ the same way)."""
    if os.path.exists(os.path.join(truth_dir, name + ".mat")):
        return os.path.join(truth_dir, "")
    for sub in sorted(os.listdir(truth_dir)):
        if os.path.exists(os.path.join(truth_dir, sub, name + ".mat")):
            return os.path.join(truth_dir, sub, "")
    raise FileNotFoundError("no ground truth %s.mat under %s" % (name, truth_dir))


def evaluate_dataset(image_dir: str, truth_dir: str, names: Optional[Sequence[str]] = None, k: int = 8, iters: int = 20,
                     seed: int = 0, workers: Optional[int] = None, chunk: int = 200, want_labels: bool = False,
                     gpu_decode: bool = False, **plan_kwargs):
    """-> list of (name, metrics dict[, labels]) in sorted file-name order.  ``names`` (without extension)
    selects a subset.  Image i of the sorted list is clustered from initial centroids seeded with ``seed + i``.
    ``plan_kwargs`` go to :class:`Plan` (bank, colour_space, feature, normalise, smooth ...).
    ``gpu_decode=True`` decodes the JPEG files on the GPU (decode.decode_jpeg_batch: Huffman on host threads, inverse
    DCT / upsampling / colour conversion as kernels; same pixels as PIL) and runs the device-resident pipeline."""
    import torch
    from PIL import Image
    files = sorted(f for f in os.listdir(image_dir) if f.lower().endswith(IMAGE_EXT))
    if names is not None:
        want = set(names)
        files = [f for f in files if os.path.splitext(f)[0] in want]
    if not files:
        return []
    stems = [os.path.splitext(f)[0] for f in files]
    workers = workers or min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(workers) as ex:
        # pass 1: shapes (no decode) and ground truths (they fix G and the label capacity of each bucket)
        sizes = list(ex.map(lambda f: Image.open(os.path.join(image_dir, f)).size, files))          # (W, H)
        segs = list(ex.map(lambda s: get_segmentation(_find_truth(truth_dir, s), s), stems))
        buckets = {}
        for i, (w, h) in enumerate(sizes):
            buckets.setdefault((h, w), []).append(i)
        out = [None] * len(files)
        for (H, W), idxs in buckets.items():
            n = len(idxs)
            G = max(len(segs[i]) for i in idxs)
            n_lab = max(int(np.max(g)) for i in idxs for g in segs[i]) + 1
            imgs = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
            gts = torch.zeros((n, G, H, W), dtype=torch.int16).pin_memory()
            imgs_np, gts_np = imgs.numpy(), gts.numpy().view(np.uint16)
            n_gt = np.zeros(n, np.int32)

            blobs = [None] * n

            def load(slot_i):
                slot, i = slot_i
                if gpu_decode:
                    with open(os.path.join(image_dir, files[i]), "rb") as fh:
                        blobs[slot] = fh.read()
                else:
                    a = np.asarray(Image.open(os.path.join(image_dir, files[i])).convert("RGB"))      # script.py:25
                    if a.shape != (H, W, 3):
                        raise ValueError("%s: decoded shape %s differs from its header" % (files[i], a.shape))
                    imgs_np[slot] = a
                for g, sgm in enumerate(segs[i]):
                    sgm = np.asarray(sgm)
                    if sgm.shape != (H, W):
                        raise ValueError("%s: ground truth %d has shape %s, image is %s" % (stems[i], g, sgm.shape, (H, W)))
                    gts_np[slot, g] = sgm
                n_gt[slot] = len(segs[i])
            list(ex.map(load, enumerate(idxs)))
            plan = Plan(H, W, max_batch=min(n, chunk), k=k, iters=iters, max_gt=G, n_lab_cap=max(64, n_lab), **plan_kwargs)
            init = init_indices_for(idxs, H * W, k, seed)
            if gpu_decode:
                from .decode import decode_jpeg_batch
                from .engine import BatchCounts
                parts = []
                for a in range(0, n, plan.max_batch):          # device-resident chunks of the plan's capacity
                    b = min(n, a + plan.max_batch)
                    d_img = decode_jpeg_batch(blobs[a:b])
                    plan.pipeline_device(d_img, gts[a:b].cuda(non_blocking=True), torch.from_numpy(init[a:b]),
                                         torch.from_numpy(n_gt[a:b]))
                    parts.append(plan.fetch(want_labels=want_labels))
                c = parts[0] if len(parts) == 1 else BatchCounts(
                    H, W, *[np.concatenate([getattr(p_, f) for p_ in parts]) for f in
                            ("bd_count", "gt_counts", "area", "perim", "n_seg", "n_lab", "status", "n_gt")],
                    None, np.concatenate([p_.labels for p_ in parts]) if want_labels else None)
            else:
                c = plan.pipeline_host(imgs, gts, torch.from_numpy(init), n, n_gt, want_labels=want_labels)
            for slot, i in enumerate(idxs):
                m = finish_image(c, slot)
                out[i] = (stems[i], m, c.labels[slot].copy()) if want_labels else (stems[i], m)
            plan.close()
    return out


def format_like_display_metrics(m: dict) -> str:
    """The line ``metrics.display_metrics()`` prints (metrics.py:231-237), from a ``get_metrics()`` dict."""
    return ("Regions: " + str(m["regions"]) + " Recall: " + str(m["recall"]) + " Precision: " + str(m["precision"]) +
            " Undersegmentation: " + str(m["underseg"]) + " Undersegmentation (NP) " + str(m["undersegNP"]) +
            " Compactness " + str(m["compactness"]) + " Density " + str(m["density"]))


def print_like_script(results) -> None:
    """Per image the two lines the reference driver prints (script.py:27 and :38)."""
    for r in results:
        print("Processing image " + r[0])
        print(format_like_display_metrics(r[1]))
