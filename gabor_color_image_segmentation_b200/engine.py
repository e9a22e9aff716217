"""Plan wrapper: PyTorch tensors in and out, CUDA through the C ABI (include/gcis.h).

PyTorch is used for device memory and streams only; all arithmetic runs in libgcis.so.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib


@dataclass(frozen=True)
class GaborBank:
    """Filter bank.  Default: 4 octave-spaced scales x 6 orientations."""
    frequencies: tuple
    thetas: tuple
    bandwidth: float = 1.0
    n_stds: float = 3.0

    @staticmethod
    def default(n_scales: int = 4, n_orient: int = 6, f0: float = 0.25, octave_step: float = 1.0) -> "GaborBank":
        return GaborBank(tuple(f0 * 2.0 ** (-octave_step * s) for s in range(n_scales)),
                         tuple(o * math.pi / n_orient for o in range(n_orient)))

    @staticmethod
    def dense() -> "GaborBank":
        return GaborBank.default(8, 12, 0.25, 0.5)

    def half_width(self, s: int, o: int) -> int:
        return _lib.check(_lib.load().gcis_gabor_half_width(self.frequencies[s], self.thetas[o],
                                                            self.bandwidth, self.n_stds))

    def separable(self, s: int, o: int):
        """Complex 1-D factors (gx, gy) exactly as the CUDA path builds them."""
        h = self.half_width(s, o)
        n = 2 * h + 1
        arrs = [np.zeros(n, np.float64) for _ in range(4)]
        _lib.check(_lib.load().gcis_gabor_separable(self.frequencies[s], self.thetas[o], self.bandwidth, self.n_stds,
                                                    *[a.ctypes.data for a in arrs], n))
        return arrs[0] + 1j * arrs[1], arrs[2] + 1j * arrs[3]


@dataclass
class BatchCounts:
    """Integer records of a batch (SURVEY.md A.3–A.7): the bit-exact contract."""
    H: int
    W: int
    bd_count: np.ndarray     # [B] int64
    gt_counts: np.ndarray    # [B][G][8] int64: den_r, tp_r, tp_p, U, V, sum n_ij^2, sum b_j^2, -
    area: np.ndarray         # [B][n_seg_cap] int32
    perim: np.ndarray        # [B][n_seg_cap] int32
    n_seg: np.ndarray        # [B] int32
    n_lab: np.ndarray        # [B][G] int32
    status: np.ndarray       # [B] int32
    n_gt: np.ndarray         # [B] int32
    hist: Optional[np.ndarray] = None   # [B][G][n_seg_cap][n_lab_cap] int32
    labels: Optional[np.ndarray] = None  # [B][H][W] int32


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.GcisError("no CUDA device: gabor_color_image_segmentation_b200 has no CPU path")
    return torch


class Plan:
    """Device workspaces + bank tables for one image shape and configuration."""

    def __init__(self, height: int, width: int, max_batch: int = 1, bank: Optional[GaborBank] = None,
                 colour_space: str = "rgb", feature: str = "magnitude", k: int = 8, iters: int = 20,
                 fix_shift: int = 24, max_gt: int = 5, n_lab_cap: int = 64, dil_recall: int = 5, group: int = 0,
                 normalise: bool = False, smooth: float = 0.0):
        self.lib = _lib.load()
        self.bank = bank or GaborBank.default()
        self.H, self.W, self.N = int(height), int(width), int(height) * int(width)
        self.max_batch, self.k, self.iters = int(max_batch), int(k), int(iters)
        self.max_gt, self.n_lab_cap, self.fix_shift = int(max_gt), int(n_lab_cap), int(fix_shift)
        self._f = (C.c_double * len(self.bank.frequencies))(*self.bank.frequencies)
        self._t = (C.c_double * len(self.bank.thetas))(*self.bank.thetas)
        cfg = _lib.GcisConfig(self.H, self.W, self.max_batch, _lib.COLOUR[colour_space],
                              len(self.bank.frequencies), len(self.bank.thetas), self._f, self._t,
                              self.bank.bandwidth, self.bank.n_stds, _lib.FEATURE[feature], self.k, self.iters,
                              self.fix_shift, self.max_gt, self.n_lab_cap, int(dil_recall), int(group),
                              int(bool(normalise)), float(smooth))
        self.normalise, self.smooth = bool(normalise), float(smooth)
        h = C.c_void_p()
        _lib.check(self.lib.gcis_plan_create(C.byref(cfg), C.byref(h)), "gcis_plan_create")
        self._h = h
        self.D = _lib.check(self.lib.gcis_plan_feature_dim(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.gcis_plan_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.gcis_plan_workspace_bytes(self._h))

    def launch_group(self, B: int) -> int:
        """Images per kernel launch for a batch of B images."""
        return _lib.check(self.lib.gcis_plan_launch_group(self._h, int(B)))

    @property
    def uses_tensor_cores(self) -> bool:
        return bool(_lib.check(self.lib.gcis_plan_uses_tensor_cores(self._h)))

    @staticmethod
    def _stream():
        return C.c_void_p(_torch().cuda.current_stream().cuda_stream)

    def _chk_img(self, img):
        torch = _torch()
        if img.dtype != torch.uint8 or img.dim() != 4 or tuple(img.shape[1:]) != (self.H, self.W, 3) or not img.is_cuda:
            raise ValueError(f"img must be a CUDA uint8 tensor [B,{self.H},{self.W},3]")
        return img.contiguous()

    # ---- stages (device tensors) ----
    def gabor_features(self, img):
        torch = _torch()
        img = self._chk_img(img)
        B = img.shape[0]
        feat = torch.empty((B, self.D, self.H, self.W), dtype=torch.float32, device=img.device)
        _lib.check(self.lib.gcis_gabor_features(self._h, img.data_ptr(), B, feat.data_ptr(), self._stream()),
                   "gcis_gabor_features")
        return feat

    def _chk_init(self, init_idx, B):
        if tuple(init_idx.shape) != (B, self.k):
            raise ValueError(f"init_idx must be [{B},{self.k}] (one pixel index per cluster), got {tuple(init_idx.shape)}")

    def kmeans(self, feat, init_idx):
        torch = _torch()
        if (not isinstance(feat, torch.Tensor) or not feat.is_cuda or feat.dtype != torch.float32 or feat.dim() != 4
                or tuple(feat.shape[1:]) != (self.D, self.H, self.W)):
            raise ValueError(f"feat must be a CUDA float32 tensor [B,{self.D},{self.H},{self.W}]")
        if not 1 <= feat.shape[0] <= self.max_batch:
            raise ValueError(f"feat holds {feat.shape[0]} images, plan capacity is {self.max_batch}")
        feat = feat.contiguous()
        B = feat.shape[0]
        self._chk_init(init_idx, B)
        idx = init_idx.to(device=feat.device, dtype=torch.int32).contiguous()
        labels = torch.empty((B, self.H, self.W), dtype=torch.int32, device=feat.device)
        cent = torch.empty((B, self.k, self.D), dtype=torch.float32, device=feat.device)
        _lib.check(self.lib.gcis_kmeans(self._h, feat.data_ptr(), B, idx.data_ptr(), labels.data_ptr(),
                                        cent.data_ptr(), self._stream()), "gcis_kmeans")
        return labels, cent

    def feature_affine(self, feat):
        """[B,D,2] float32 (a, b) of the per-feature z-score z = a*x + b of a feature tensor (normalise=True plans)."""
        torch = _torch()
        if (not feat.is_cuda or feat.dtype != torch.float32 or feat.dim() != 4
                or tuple(feat.shape[1:]) != (self.D, self.H, self.W) or not 1 <= feat.shape[0] <= self.max_batch):
            raise ValueError(f"feat must be a CUDA float32 tensor [B<={self.max_batch},{self.D},{self.H},{self.W}]")
        feat = feat.contiguous()
        out = torch.empty((feat.shape[0], self.D, 2), dtype=torch.float32, device=feat.device)
        _lib.check(self.lib.gcis_feature_affine(self._h, feat.data_ptr(), feat.shape[0], out.data_ptr(), self._stream()),
                   "gcis_feature_affine")
        return out

    def segment(self, img, init_idx):
        torch = _torch()
        img = self._chk_img(img)
        B = img.shape[0]
        self._chk_init(init_idx, B)
        idx = init_idx.to(device=img.device, dtype=torch.int32).contiguous()
        labels = torch.empty((B, self.H, self.W), dtype=torch.int32, device=img.device)
        _lib.check(self.lib.gcis_segment_device(self._h, img.data_ptr(), B, idx.data_ptr(), labels.data_ptr(),
                                                self._stream()), "gcis_segment_device")
        return labels

    # ---- whole path ----
    def pipeline_device(self, img, gt, init_idx, n_gt=None):
        """Inputs resident in HBM; results stay on the device until fetch()."""
        torch = _torch()
        img = self._chk_img(img)
        B = img.shape[0]
        if gt.dtype != torch.uint16 and gt.dtype != torch.int16:
            raise ValueError("gt must be uint16 [B,G,H,W]")
        if tuple(gt.shape) != (B, self.max_gt, self.H, self.W):
            raise ValueError(f"gt must be [B,{self.max_gt},{self.H},{self.W}]")
        gt = gt.contiguous()
        self._chk_init(init_idx, B)
        if n_gt is not None and tuple(n_gt.shape) != (B,):
            raise ValueError(f"n_gt must be [{B}]")
        idx = init_idx.to(device=img.device, dtype=torch.int32).contiguous()
        ng = n_gt.to(device=img.device, dtype=torch.int32).contiguous() if n_gt is not None else None
        _lib.check(self.lib.gcis_pipeline_device(self._h, img.data_ptr(), gt.data_ptr(),
                                                 ng.data_ptr() if ng is not None else None, idx.data_ptr(), B,
                                                 self._stream()), "gcis_pipeline_device")
        self._last = (B, ng)
        return B

    def _alloc_out(self, B, want_labels):
        G = max(self.max_gt, 1)
        out = dict(bd=np.zeros(B, np.int64), gc=np.zeros((B, G, _lib.GT_SLOTS), np.int64),
                   area=np.zeros((B, self.k), np.int32), perim=np.zeros((B, self.k), np.int32),
                   n_lab=np.zeros((B, G), np.int32), status=np.zeros(B, np.int32),
                   labels=np.zeros((B, self.H, self.W), np.int32) if want_labels else None)
        return out

    def _counts(self, o, B, n_gt):
        st = o["status"]
        if st.any():
            bad = int(np.flatnonzero(st)[0])
            if int(st[bad]) & _lib.ST_LAB_OVER:
                raise _lib.GcisError(f"image {bad}: a ground truth has max(gt)+1 = {int(o['n_lab'][bad].max())} regions, "
                                     f"above the plan's n_lab_cap = {self.n_lab_cap}; build the plan with a larger "
                                     "n_lab_cap (pipeline.evaluate_batch / evaluate_mixed size it from the data)")
            raise _lib.GcisError(f"image {bad}: label out of range (status {int(st[bad])})")
        ng = np.full(B, self.max_gt, np.int32) if n_gt is None else np.asarray(n_gt, np.int32)
        n_seg = np.zeros(B, np.int32)
        for b in range(B):
            nz = np.flatnonzero(o["area"][b])
            n_seg[b] = int(nz[-1]) + 1 if nz.size else 0
        return BatchCounts(self.H, self.W, o["bd"], o["gc"], o["area"], o["perim"], n_seg, o["n_lab"], st, ng,
                           None, o["labels"])

    def fetch(self, want_labels: bool = False) -> BatchCounts:
        B, ng = self._last
        o = self._alloc_out(B, want_labels)
        _lib.check(self.lib.gcis_pipeline_fetch(self._h, B, o["bd"].ctypes.data, o["gc"].ctypes.data,
                                                o["area"].ctypes.data, o["perim"].ctypes.data,
                                                o["n_lab"].ctypes.data, o["status"].ctypes.data,
                                                o["labels"].ctypes.data if want_labels else None, self._stream()),
                   "gcis_pipeline_fetch")
        return self._counts(o, B, ng.cpu().numpy() if ng is not None else None)

    def fetch_hist(self, B: int) -> np.ndarray:
        """Contingency tables [B,G,k,n_lab_cap] int32 of the last pipeline call (up to max_batch images)."""
        h = np.zeros((B, max(self.max_gt, 1), self.k, self.n_lab_cap), np.int32)
        _lib.check(self.lib.gcis_pipeline_fetch_hist(self._h, B, h.ctypes.data, self._stream()),
                   "gcis_pipeline_fetch_hist")
        return h

    def pipeline_host(self, img_ptr, gt_ptr, init_ptr, B: int, n_gt: Optional[np.ndarray] = None,
                      want_labels: bool = False) -> BatchCounts:
        """Host buffers in, host records out (copies inside).  *_ptr: objects with a host address —
        numpy arrays or (pinned) CPU torch tensors of the documented shapes."""
        def addr(x):
            return x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr()

        def chk(x, name, shape, itemsize):
            # the C ABI reads shape-many elements from a raw address: a wrong shape or dtype would be an
            # out-of-bounds read inside libgcis.so, so it is rejected here
            if isinstance(x, np.ndarray):
                ok = x.flags["C_CONTIGUOUS"] and x.dtype.itemsize == itemsize and x.dtype.kind in "iu"
            else:
                ok = (not x.is_cuda) and x.is_contiguous() and x.element_size() == itemsize and not x.dtype.is_floating_point
            if not ok or tuple(x.shape) != shape:
                raise ValueError(f"{name} must be a contiguous host array of shape {shape} with {itemsize}-byte integers, "
                                 f"got shape {tuple(x.shape)} dtype {x.dtype}")
        B = int(B)
        if B < 1:
            raise ValueError("B must be >= 1")
        chk(img_ptr, "imgs", (B, self.H, self.W, 3), 1)
        if self.max_gt > 0:
            chk(gt_ptr, "gts", (B, self.max_gt, self.H, self.W), 2)
        chk(init_ptr, "init_idx", (B, self.k), 4)
        o = self._alloc_out(B, want_labels)
        ng = None if n_gt is None else np.ascontiguousarray(n_gt, np.int32)
        if ng is not None and (ng.shape != (B,) or ng.min() < 0 or ng.max() > self.max_gt):
            raise ValueError(f"n_gt must be [{B}] with values in 0..{self.max_gt}")
        _lib.check(self.lib.gcis_pipeline_host(self._h, addr(img_ptr), addr(gt_ptr) if self.max_gt > 0 else None,
                                               ng.ctypes.data if ng is not None else None, addr(init_ptr), B,
                                               o["bd"].ctypes.data, o["gc"].ctypes.data, o["area"].ctypes.data,
                                               o["perim"].ctypes.data, o["n_lab"].ctypes.data,
                                               o["status"].ctypes.data,
                                               o["labels"].ctypes.data if want_labels else None),
                   "gcis_pipeline_host")
        return self._counts(o, B, ng)

    def set_profiling(self, on: bool):
        _lib.check(self.lib.gcis_plan_set_profiling(self._h, int(on)))

    def last_stage_ms(self):
        ms = (C.c_float * 4)()
        _lib.check(self.lib.gcis_plan_last_stage_ms(self._h, ms))
        return dict(zip(("colour", "gabor", "kmeans", "metrics"), [float(v) for v in ms]))


def kmeans_init_indices(n_pixels: int, k: int, seed: int) -> np.ndarray:
    """Initial centroid pixels: default_rng(seed).choice(N, k, replace=False)."""
    return np.random.default_rng(seed).choice(n_pixels, k, replace=False).astype(np.int32)


def label_counts_host(lbs: np.ndarray, gts: np.ndarray, n_gt: Optional[Sequence[int]] = None,
                      n_seg_cap: Optional[int] = None, n_lab_cap: Optional[int] = None, dil_recall: int = 5,
                      want_hist: bool = False) -> BatchCounts:
    """BSD_metrics integer counts for a batch of label maps (metrics.py:25-201) on the GPU.
    lbs [B,H,W] int32, gts [B,G,H,W] uint16."""
    lib = _lib.load()
    lbs = np.ascontiguousarray(lbs, np.int32)
    gts = np.ascontiguousarray(gts, np.uint16)
    B, H, W = lbs.shape
    G = gts.shape[1] if gts.ndim == 4 else 0
    if G and gts.shape != (B, G, H, W):
        raise ValueError("ground-truth shape mismatch")
    if n_seg_cap is None:
        n_seg_cap = max(int(lbs.max()) + 1, 1)
    if n_lab_cap is None:
        n_lab_cap = max(int(gts.max()) + 1, 1) if G else 1
    Gs = max(G, 1)
    bd = np.zeros(B, np.int64); gc = np.zeros((B, Gs, _lib.GT_SLOTS), np.int64)
    area = np.zeros((B, n_seg_cap), np.int32); perim = np.zeros((B, n_seg_cap), np.int32)
    hist = np.zeros((B, Gs, n_seg_cap, n_lab_cap), np.int32) if want_hist else None
    n_seg = np.zeros(B, np.int32); n_lab = np.zeros((B, Gs), np.int32); status = np.zeros(B, np.int32)
    ng = None if n_gt is None else np.ascontiguousarray(n_gt, np.int32)
    rc = lib.gcis_label_metrics_host(lbs.ctypes.data, gts.ctypes.data if G else None,
                                     ng.ctypes.data if ng is not None else None, B, H, W, G, n_seg_cap, n_lab_cap,
                                     dil_recall, bd.ctypes.data, gc.ctypes.data, area.ctypes.data,
                                     perim.ctypes.data, hist.ctypes.data if want_hist else None, n_seg.ctypes.data,
                                     n_lab.ctypes.data, status.ctypes.data)
    if rc == -3:
        bad = int(np.flatnonzero(status)[0])
        if status[bad] & _lib.ST_NEG_LABEL:
            raise ValueError(f"image {bad}: negative labels are not supported")
        raise IndexError(f"image {bad}: label exceeds capacity (status {int(status[bad])})")
    _lib.check(rc, "gcis_label_metrics_host")
    return BatchCounts(H, W, bd, gc, area, perim, n_seg, n_lab, status,
                       np.full(B, G, np.int32) if ng is None else ng, hist, None)
