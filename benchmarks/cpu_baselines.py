"""CPU baselines timed beside the GPU run (BASELINE.md "CPU-baseline plan", steps 3-4).  BENCH INFRASTRUCTURE:
this module executes oracle/ (the C port) and oracle/_ref (the reference's own metrics.py, unmodified) as the
things MEASURED for the cpu_baseline legs only; nothing here is on the product path.

  reference_metrics   the REFERENCE's own ``metrics(img, lb, gts).set_metrics()`` (BSD_metrics/metrics.py:208-217,
                      pure-Python pixel loops) from oracle/_ref over the scipy stand-in for scikit-image:
                      single thread (the reference has no parallelism) and a multiprocessing.Pool over images.
  strong              the "strong CPU baseline": scipy.signal.fftconvolve Gabor bank + sklearn KMeans (Lloyd,
                      n_init=1, 20 iterations) + the reference metrics, single thread and Pool.
  port                the repo's C oracle of the whole path, one image per host thread (bench.py).
"""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W, G, K, ITERS = 321, 481, 5, 8, 20


def _import_reference_metrics():
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "metrics.py")):
        return None
    for p in (os.path.join(ROOT, "oracle", "ref_shim"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    import metrics as ref_metrics   # the reference module (oracle/_ref/metrics.py)
    return ref_metrics


def _inputs(i):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from gabor_color_image_segmentation_b200.synth import synth_image, synth_ground_truths, voronoi_labels
    img = synth_image(20_000 + i, H, W)
    gts = list(synth_ground_truths(20_000 + i, H, W, G))
    lb = voronoi_labels(np.random.default_rng(i), H, W, K)
    return img, lb, gts


def _ref_metrics_one(i):
    """One image through the reference's own class; returns seconds spent inside set_metrics()."""
    ref_metrics = _import_reference_metrics()
    img, lb, gts = _inputs(i)
    t0 = time.perf_counter()
    m = ref_metrics.metrics(img, lb, gts)
    m.set_metrics()
    return time.perf_counter() - t0


def _gabor_fft(img):
    """72 magnitude planes by scipy.signal.fftconvolve (reflect padding), the bank of DESIGN.md section 3."""
    from scipy.signal import fftconvolve
    planes = [img[..., c].astype(np.float64) / 255.0 for c in range(3)]
    feats = []
    for c in range(3):
        for s in range(4):
            f = 0.25 * 2.0 ** (-s)
            sig = math.sqrt(math.log(2) / 2) / math.pi * 3.0 / f
            for o in range(6):
                th = o * math.pi / 6
                h = math.ceil(max(abs(3 * sig * math.cos(th)), abs(3 * sig * math.sin(th)), 1))
                y, x = np.mgrid[-h:h + 1, -h:h + 1]
                rx = x * math.cos(th) + y * math.sin(th)
                g = np.exp(-0.5 * (x * x + y * y) / sig ** 2) / (2 * math.pi * sig * sig) * np.exp(2j * math.pi * f * rx)
                pad = np.pad(planes[c], h, mode="symmetric")
                feats.append(np.abs(fftconvolve(pad, g, mode="valid")))
    return np.stack(feats)


def _strong_one(i):
    """FFT Gabor + sklearn k-means + reference metrics for one image; seconds."""
    from sklearn.cluster import KMeans
    try:
        from threadpoolctl import threadpool_limits
    except Exception:   # noqa: BLE001
        threadpool_limits = None
    ref_metrics = _import_reference_metrics()
    img, _, gts = _inputs(i)
    t0 = time.perf_counter()
    feat = _gabor_fft(img).reshape(72, -1).T.astype(np.float32)
    init = feat[np.random.default_rng(i).choice(feat.shape[0], K, replace=False)]
    km = KMeans(n_clusters=K, init=init, n_init=1, max_iter=ITERS, tol=0.0, algorithm="lloyd")
    if threadpool_limits is not None:
        with threadpool_limits(limits=1):
            labels = km.fit_predict(feat)
    else:
        labels = km.fit_predict(feat)
    m = ref_metrics.metrics(img, labels.reshape(H, W), gts)
    m.set_metrics()
    return time.perf_counter() - t0


def _pool_rate(fn, n, procs):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")          # the parent may hold a CUDA context: never fork it
    with ctx.Pool(procs) as pool:
        pool.map(_warm, range(procs))      # imports done before the clock starts
        t0 = time.perf_counter()
        pool.map(fn, range(n), chunksize=1)
        dt = time.perf_counter() - t0
    return n / dt, dt


def _warm(_):
    _import_reference_metrics()
    import scipy.signal   # noqa: F401
    import sklearn.cluster   # noqa: F401
    return 0


def measure(procs=None, n_single=2):
    """dict with the reference-metrics and strong-baseline rates, or {'unavailable': why}."""
    procs = procs or os.cpu_count() or 1
    if _import_reference_metrics() is None:
        return {"unavailable": "oracle/_ref/metrics.py is absent (run __graft_entry__.build() where /root/reference exists)"}
    out = {}
    t = [_ref_metrics_one(i) for i in range(n_single)]
    pool_rate, pool_s = _pool_rate(_ref_metrics_one, 2 * procs, procs)
    out["reference_metrics"] = {
        "what": "the reference's own metrics.set_metrics() (BSD_metrics/metrics.py:208-217) from oracle/_ref, unmodified, "
                "over the scipy stand-in for scikit-image; 321x481 labels with 8 regions, 5 ground truths",
        "single_thread_s_per_image": float(np.mean(t)), "single_thread_images_per_s": float(1.0 / np.mean(t)),
        "pool_images_per_s": float(pool_rate), "pool_processes": procs,
        "sample": "%d images single thread, %d images over the pool (%.1f s)" % (n_single, 2 * procs, pool_s)}
    t = [_strong_one(0)]
    pool_rate, pool_s = _pool_rate(_strong_one, procs, procs)
    out["strong"] = {
        "what": "scipy.signal.fftconvolve Gabor bank (72 complex convolutions) + sklearn KMeans(lloyd, n_init=1, 20 "
                "iterations) + the reference's metrics.set_metrics(); not reference code for the first two stages "
                "(the reference has none)",
        "single_thread_s_per_image": float(np.mean(t)), "single_thread_images_per_s": float(1.0 / np.mean(t)),
        "pool_images_per_s": float(pool_rate), "pool_processes": procs,
        "sample": "1 image single thread, %d images over the pool (%.1f s)" % (procs, pool_s)}
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(measure(), indent=1))
