"""Region scores beyond the reference, from the contingency tables the GPU already builds
(metrics.py:115-126 computes the same table for the undersegmentation errors): Probabilistic
Rand Index, Variation of Information and segmentation covering.

Builder-defined (the reference has no such code, parity unpinned):
  PRI      = mean over ground truths of RI = 1 - [ (sum_i a_i^2 + sum_j b_j^2)/2 - sum_ij n_ij^2 ] / C(N,2)
  VoI      = mean over ground truths of H(S) + H(G) - 2 I(S;G), entropies in bits
  covering = mean over ground truths of (1/N) sum_j b_j max_i n_ij / (a_i + b_j - n_ij)
with n_ij the table, a_i its row sums (segments), b_j its column sums (ground-truth regions).
The tables are integers computed on the GPU; these few hundred float operations per image
finish on the host in float64."""
import numpy as np


def rand_index(h: np.ndarray) -> float:
    h = h.astype(np.int64)
    N = int(h.sum())
    a = h.sum(1); b = h.sum(0)
    pairs = N * (N - 1) // 2
    disagree = (int((a * a).sum()) + int((b * b).sum())) // 2 - int((h * h).sum())
    return 1.0 - disagree / pairs


def variation_of_information(h: np.ndarray) -> float:
    h = h.astype(np.float64)
    N = h.sum()
    p = h / N
    pa = p.sum(1); pb = p.sum(0)
    ent = lambda q: float(-(q[q > 0] * np.log2(q[q > 0])).sum())
    nz = p > 0
    mi = float((p[nz] * np.log2(p[nz] / (pa[:, None] * pb[None, :])[nz])).sum())
    return ent(pa) + ent(pb) - 2.0 * mi


def covering(h: np.ndarray) -> float:
    h = h.astype(np.float64)
    N = h.sum()
    a = h.sum(1); b = h.sum(0)
    union = a[:, None] + b[None, :] - h
    iou = np.where(union > 0, h / np.where(union > 0, union, 1.0), 0.0)
    return float((b * iou.max(0)).sum() / N)


def region_scores(hist: np.ndarray, n_gt=None) -> dict:
    """hist [B,G,nS,nL] -> {'pri','voi','covering'}: arrays [B] (mean over each image's ground truths)."""
    B, G = hist.shape[:2]
    n_gt = np.full(B, G, np.int32) if n_gt is None else np.asarray(n_gt)
    out = {k: np.zeros(B) for k in ("pri", "voi", "covering")}
    for b in range(B):
        g_n = int(n_gt[b])
        for g in range(g_n):
            out["pri"][b] += rand_index(hist[b, g])
            out["voi"][b] += variation_of_information(hist[b, g])
            out["covering"][b] += covering(hist[b, g])
        for k in out:
            out[k][b] /= g_n
    return out
