#!/usr/bin/env python
"""bench.py — images/sec of the hot path (Gabor bank + k-means + BSD metrics) on synthetic
BSDS-shaped data (481x321 RGB, 5 Voronoi ground truths per image).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host cores)

A step = one pass of the whole path over one batch of `--images` images per GPU (BASELINE.json
configs[1]: a 200-image BSDS-test-shaped batch).  Weak scaling: every rank processes its own batch.
  value  : images/s with inputs already resident in HBM (device timing, CUDA events, max over ranks)
  e2e    : images/s through the C ABI's host entry point (pinned host buffers in, records out)
The oracle (oracle/) and oracle/_ref (the reference's own metrics.py, copied there by build()) are executed only
for the cpu_baseline / --impl reference legs (benchmarks/cpu_baselines.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, G, K_CLUSTERS, ITERS = 321, 481, 5, 8, 20
METRIC = "images/sec (481x321 RGB, Gabor+cluster+BSD eval)"


def make_data(n, start=0, workers=None):
    """n synthetic images + ground truths (SURVEY.md §8 d), generated in parallel on the host."""
    from concurrent.futures import ProcessPoolExecutor
    from gabor_color_image_segmentation_b200.synth import synth_image, synth_ground_truths
    workers = workers or min(os.cpu_count() or 1, 16)
    idx = list(range(start, start + n))
    if n <= 4 or workers <= 1:
        imgs = [synth_image(i, H, W) for i in idx]
        gts = [synth_ground_truths(i, H, W, G) for i in idx]
    else:
        with ProcessPoolExecutor(workers) as ex:
            imgs = list(ex.map(synth_image, idx, chunksize=4))
            gts = list(ex.map(synth_ground_truths, idx, chunksize=4))
    return np.stack(imgs), np.stack(gts)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu_index, [], threading.Event()
        self.ready = threading.Event()   # set once the first sample exists: start_and_wait() returns only then

    def mark(self):
        """Start of the timed region: drop the idle samples (the last one stays as a fallback for a very short region)."""
        self.idle_last = self.samples[-1] if self.samples else None
        self.samples = []

    def start_and_wait(self, timeout=5.0):
        """NVML initialisation can take longer than a short timed region: wait for the first sample before timing."""
        self.start()
        self.ready.wait(timeout)
        return self

    def run(self):
        if self._run_nvml():
            return
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
                    self.ready.set()
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def _run_nvml(self):
        """Same fields through NVML (nvidia_ml_py): one sample every 20 ms instead of one nvidia-smi process
        every few hundred ms, so even a 300 ms timed region is covered by a dozen samples."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            reasons_fn(h)
        except Exception:
            return False
        bits = [0x8, 0x40, 0x20, 0x4]   # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap

        def sample():
            r = reasons_fn(h)
            try:
                power = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                power = 0.0
            return ([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(max_sm), str(power)] +
                    ["Active" if r & b else "Not Active" for b in bits])

        try:
            sample()            # a box whose NVML cannot answer falls back to nvidia-smi polling
        except Exception:
            return False
        while not self.stop_flag.is_set():
            try:
                self.samples.append(sample())
                self.ready.set()
            except Exception:
                pass
            self.stop_flag.wait(0.02)
        return True

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples and getattr(self, "idle_last", None):
            self.samples = [self.idle_last]
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples), "power_w_max": max(float(s[2]) for s in self.samples)}


def cpu_pipeline_rate(n_images, threads):
    """Oracle port of the whole path on `threads` host threads; returns (images/s, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    from gabor_color_image_segmentation_b200.pipeline import init_indices_for
    orc.lib()
    imgs, gts = make_data(n_images, start=10_000, workers=1)   # no fork once CUDA is up
    idx = init_indices_for(range(n_images), H * W, K_CLUSTERS)

    def one(i):
        _oracle_image(orc, imgs[i], gts[i], idx[i])
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(one, range(n_images)))
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def _oracle_image(orc, img, gts, idx):
    labels, _, _ = orc.segment_image(img, K_CLUSTERS, ITERS, init_idx=idx)
    c = orc.label_counts(labels, list(gts))
    return orc.finish_metrics(c)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = max(threads, 4)
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_pipeline_rate(min(threads, 2), threads)
    rates, secs = [], []
    for _ in range(args.steps):
        r, dt = cpu_pipeline_rate(sample, threads)
        rates.append(r); secs.append(dt)
    value = sample * len(secs) / sum(secs)
    extra = {}
    if not args.no_ref_metrics:
        from benchmarks import cpu_baselines
        extra = cpu_baselines.measure(threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: synthetic BSDS-shaped 321x481 RGB, bank 4x6, k=8, T=20, G=5",
                       "images_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": "%d images per step, one image per host thread, C oracle "
                                       "(fp64 separable Gabor + fp32-exact k-means + metrics)" % sample, **extra},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from gabor_color_image_segmentation_b200 import Plan, _lib
    from gabor_color_image_segmentation_b200.metrics import finish_batch
    from gabor_color_image_segmentation_b200.pipeline import SUM_KEYS, init_indices_for, reduce_sums

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU path; use --impl reference for the CPU arm)")
    B = args.images
    # every rank gets its own images (weak scaling); generate `unique` of them and cycle.
    # Host-side generation forks workers, so it runs before CUDA/NCCL are initialised.
    unique = min(B, args.unique)
    pinned = pin_to_local_cores(local, world)
    imgs_u, gts_u = make_data(unique, start=rank * 100_000, workers=max(1, min(16, (os.cpu_count() or 1) // max(world, 1))))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    reps = (B + unique - 1) // unique
    imgs = np.concatenate([imgs_u] * reps)[:B]
    gts = np.concatenate([gts_u] * reps)[:B]
    idx = init_indices_for(range(rank * B, rank * B + B), H * W, K_CLUSTERS)

    plan = Plan(H, W, max_batch=B, k=K_CLUSTERS, iters=ITERS, max_gt=G, n_lab_cap=64, group=args.group)
    launch_group = plan.launch_group(B)          # images per kernel launch (200 -> 4 launches of 50)
    uses_tc = plan.uses_tensor_cores
    d_img = torch.from_numpy(imgs).to(dev)
    d_gt = torch.from_numpy(gts.view(np.int16)).to(dev)
    d_idx = torch.from_numpy(idx).to(dev)
    h_img = torch.from_numpy(imgs).pin_memory()
    h_gt = torch.from_numpy(gts.view(np.int16)).pin_memory()
    h_idx = torch.from_numpy(idx).pin_memory()

    def finish(c):
        m = finish_batch(c)
        local_sums = np.array([m[k].sum() for k in SUM_KEYS] + [float(len(c.bd_count))])
        return reduce_sums(local_sums, dev)

    def step_device():
        plan.pipeline_device(d_img, d_gt, d_idx)
        return finish(plan.fetch())

    def step_host():
        return finish(plan.pipeline_host(h_img, h_gt, h_idx, B))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        sums = step_device()
    # ---- value: device-resident inputs, CUDA events, max over ranks ----
    sampler = ClockSampler(local).start_and_wait()
    barrier()
    sampler.mark()            # keep only the samples taken under load
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sums = step_device()
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    # ---- e2e: host buffers through the C ABI ----
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sums_h = step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    else:
        e2e_ms = e2e_s * 1e3
    if not os.environ.get("GCIS_BENCH_NOCHECK"):   # (set only for wrong-on-purpose timing experiments)
        assert np.array_equal(sums, sums_h), "device-resident and host-buffer passes disagree"

    # ---- per-stage device times (CUDA events inside the library) for the roofline ----
    # The timed runs above overlap the Gabor kernel of one group with the k-means passes of the
    # previous one on two streams; a kernel's own duration needs it alone on the device, so the
    # stage times come from a single-stream plan over the same images.
    plan.close()
    os.environ["GCIS_LANES"] = "1"
    plan1 = Plan(H, W, max_batch=B, k=K_CLUSTERS, iters=ITERS, max_gt=G, n_lab_cap=64, group=args.group)
    plan1.pipeline_device(d_img, d_gt, d_idx)
    plan1.fetch()
    plan1.set_profiling(True)
    plan1.pipeline_device(d_img, d_gt, d_idx)
    plan1.fetch()
    stage = plan1.last_stage_ms()
    plan1.close()
    os.environ.pop("GCIS_LANES", None)
    latency = one_image_latency(imgs[0], gts[0]) if rank == 0 else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        N, D = H * W, 72
        km_bytes = B * (ITERS * N * D * 4 + N * 4)                  # T*N*D*4 + N*4 per image
        gabor_bytes = B * N * (3 + D * 4)
        gabor_flop = B * 6.98e9                                     # SURVEY.md §8(d), complex-separable count
        fma_peak = measure_fma_peak()
        km_gbs = km_bytes / (stage["kmeans"] * 1e-3) / 1e9
        gb_tfs = gabor_flop / (stage["gabor"] * 1e-3) / 1e12
        total_imgs = B * world * args.steps
        dominant = "kmeans" if stage["kmeans"] >= stage["gabor"] else "gabor"
        # DRAM bytes of ONE km_tile_kernel launch from the committed ncu capture taken at this launch size
        # (profiles/r02_km_traffic.json); null when the capture was made at another launch size
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_km_traffic.json")))
            if int(tr["images_per_launch"]) == launch_group:
                traffic, traffic_src = float(tr["dram_bytes_per_launch"]), tr["source"]
        except Exception:
            pass
        roof_km = {"kernel": "km_tile_kernel<8,256,2> (%d launches per %d-image launch group, timed with their km_finalize_kernel)"
                             % (ITERS, launch_group), "bound": "hbm",
                   "achieved": km_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": km_gbs / hbm_peak,
                   "traffic": traffic, "algorithmic_bytes_per_launch": (N * D * 4) * launch_group,
                   "images_per_launch": launch_group, "traffic_source": traffic_src,
                   "peak_source": peak_src, "ms_per_step": stage["kmeans"]}
        fp32_peak = fma_peak.get("ffma_rrr_tflops")
        # FLOPs that still run on the FP32 pipe: the column pass, 12 of the 19 real taps per pixel, channel and scale of
        # the executed 3.66 GFLOP/image; the row pass runs as bf16 MMAs on the tensor cores
        # (ncu, profiles/r02_top_kernels_ncu.md: 16.7 M warp-level FFMA2 per image = 2.14 GFLOP after the triangular sweep
        # blocks and the thin last strips; the formula's 2.31 G is the useful count and is what is reported)
        col_flop = B * 3.66e9 * (12.0 / 19.0 if uses_tc else 1.0)
        r1_tfs = B * 3.66e9 / (stage["gabor"] * 1e-3) / 1e12
        roof_gb = {"kernel": "gabor_tc_kernel (row pass: tcgen05 bf16x3 MMA, column pass: FFMA2)" if uses_tc else "gabor_bank_kernel",
                   "bound": "fp32", "achieved": gb_tfs, "peak": fp32_peak, "unit": "TFLOP/s",
                   "frac": (gb_tfs / fp32_peak) if fp32_peak else None,
                   "note": "achieved = SURVEY 8(d) count (6.98 GFLOP/image, un-paired complex-separable) / event time; with the row "
                           "pass on the tensor cores this can exceed the FP32-pipe peak" if uses_tc else
                           "achieved = SURVEY 8(d) count (6.98 GFLOP/image) / event time",
                   "fp32_pipe_executed": {"tflops": col_flop / (stage["gabor"] * 1e-3) / 1e12,
                                          "frac": (col_flop / (stage["gabor"] * 1e-3) / 1e12 / fp32_peak) if fp32_peak else None,
                                          "what": "FLOPs executed on the FP32 pipe per image: %.2f G" % (col_flop / B / 1e9)},
                   "by_round1_executed_count": {"tflops": r1_tfs, "frac": (r1_tfs / fp32_peak) if fp32_peak else None,
                                                "what": "3.66 GFLOP/image, the FLOPs round 1's all-FP32 kernel executed (VERDICT r01 item 2's yardstick)"},
                   "traffic": None, "peak_source": "measured in this run (benchmarks/fma_peak)",
                   "hbm_gbs": gabor_bytes / (stage["gabor"] * 1e-3) / 1e9, "ms_per_step": stage["gabor"]}
        line = {
            "metric": METRIC, "value": total_imgs / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: %d-image synthetic BSDS-shaped batch per GPU, 321x481 RGB, "
                                   "bank 4 scales x 6 orientations, k-means k=8 T=20, BSD metrics vs 5 ground truths" % B,
                       "images_per_step_per_gpu": B, "unique_images": unique, "images_per_launch": launch_group,
                       "l2": "inputs + feature tensor (%.1f GB per step) far exceed the 126 MB L2" % (B * N * D * 4 / 1e9)},
            "e2e": {"value": total_imgs / (e2e_ms * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": int(h_img.numel() + h_gt.numel() * 2 + h_idx.numel() * 4),
                    "d2h_bytes_per_step": int(B * (8 + G * 8 * 8 + 2 * K_CLUSTERS * 4 + G * 4 + 4))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "host_pinning": pinned,
            "roofline": roof_km if dominant == "kmeans" else roof_gb,
            "roofline_other": roof_gb if dominant == "kmeans" else roof_km,
            "stage_ms_per_step": stage,
            "latency_ms_1img": latency,
            "fma_peak": fma_peak,
            "dataset_scores": {k: float(sums[i] / sums[-1]) for i, k in enumerate(SUM_KEYS)},
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            n = max(threads, 4)
            rate, dt = cpu_pipeline_rate(n, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": "%d images, one per host thread, %.1f s: C oracle (fp64 separable Gabor + "
                                              "fp32-exact k-means + BSD metrics)" % (n, dt)}
            if not args.no_ref_metrics:
                from benchmarks import cpu_baselines
                line["cpu_baseline"].update(cpu_baselines.measure(threads))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def pin_to_local_cores(local, world_local):
    """Keep this rank's threads (and the pinned staging buffers it first-touches) on the host cores nearest to its GPU:
    the GPU's NUMA node when the box has several, otherwise an even slice of the cores, so that N ranks do not
    migrate over one another's caches.  Returns a description for the bench line."""
    try:
        import torch
        cpus = sorted(os.sched_getaffinity(0))
        pci = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        node = -1
        if pci is not None:
            import glob
            for path in glob.glob("/sys/bus/pci/devices/*:%02x:*/numa_node" % pci):
                node = int(open(path).read().strip())
                break
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()] \
            if os.path.isdir("/sys/devices/system/node") else []
        if node >= 0 and len(nodes) > 1:
            txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
            want = set()
            for part in txt.split(","):
                a, _, b = part.partition("-")
                want.update(range(int(a), int(b or a) + 1))
            mine = [c for c in cpus if c in want] or cpus
            how = "numa node %d" % node
        else:
            per = max(1, len(cpus) // max(world_local, 1))
            mine = cpus[local * per:(local + 1) * per] or cpus
            how = "core slice (single NUMA node)"
        os.sched_setaffinity(0, mine)
        return {"cores": len(mine), "how": how}
    except Exception as e:  # noqa: BLE001
        return {"cores": None, "how": "not pinned: %s" % e}


def run_gpu_strong(args):
    """BASELINE config 5: ONE fixed batch of --total-images synthetic images sharded over the ranks
    (image i -> rank i mod N, pipeline.shard_indices), NCCL sum of the metric sums every step and one all-gather
    of the integer records; the gathered table (and therefore every score finished from it in global image
    order) must be identical for 1/2/4/8 GPUs: the line carries its sha256.
    Image i of the batch is synthetic image (i mod --unique) clustered from its own initial centroids (seeded by
    the global index i), so all --total-images results differ while only --unique images are generated."""
    import hashlib
    import torch
    import torch.distributed as dist
    from gabor_color_image_segmentation_b200 import Plan, _lib
    from gabor_color_image_segmentation_b200.metrics import finish_batch
    from gabor_color_image_segmentation_b200 import pipeline as pl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    total, U, B = args.total_images, min(args.unique, args.total_images), args.images
    if U % world:
        raise SystemExit("--unique must be a multiple of the number of GPUs in strong-scaling mode")
    mine = pl.shard_indices(total, rank, world)                       # global indices of this rank
    uniq = np.arange(rank, U, world)                                  # synthetic images this rank ever sees
    pinned = pin_to_local_cores(local, world)
    imgs_u, gts_u = make_data_indices(uniq, workers=max(1, min(16, (os.cpu_count() or 1) // max(world, 1))))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pos = {int(u): n for n, u in enumerate(uniq)}
    sel = np.array([pos[int(i) % U] for i in mine], np.int64)         # this rank's images as rows of imgs_u
    chunks = [np.arange(a, min(a + B, len(mine))) for a in range(0, len(mine), B)]
    plan = Plan(H, W, max_batch=B, k=K_CLUSTERS, iters=ITERS, max_gt=G, n_lab_cap=64, group=args.group)
    idx_all = pl.init_indices_for(mine, H * W, K_CLUSTERS)
    # inputs of every step, host (pinned) and device.  Steps whose images coincide (the batch cycles through the
    # rank's --unique / N synthetic images) share one payload; only the initial-centroid indices differ per step.
    payload_h, payload_d, h_chunks, d_in = {}, {}, [], []
    for ch in chunks:
        key = sel[ch].tobytes()
        if key not in payload_h:
            payload_h[key] = (torch.from_numpy(imgs_u[sel[ch]]).pin_memory(), torch.from_numpy(gts_u[sel[ch]].view(np.int16)).pin_memory())
            payload_d[key] = tuple(t.to(dev) for t in payload_h[key])
        ix = torch.from_numpy(idx_all[ch]).pin_memory()
        h_chunks.append(payload_h[key] + (ix,))
        d_in.append(payload_d[key] + (ix.to(dev),))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(host):
        recs, sums = [], np.zeros(7)
        for n, ch in enumerate(chunks):
            if host:
                im, gt, ix = h_chunks[n]
                c = plan.pipeline_host(im, gt, ix, len(ch))
            else:
                plan.pipeline_device(*d_in[n])
                c = plan.fetch()
            m = finish_batch(c)
            sums = sums + pl.reduce_sums(np.array([m[k].sum() for k in pl.SUM_KEYS] + [float(len(ch))]), dev)
            recs.append(pl.records_to_array(c))
        table = pl.gather_records(np.concatenate(recs) if recs else np.zeros((0, 3 + 2 * K_CLUSTERS + G * 8), np.int64),
                                  mine, total, dev)
        return table, sums

    plan.pipeline_device(*d_in[0]); plan.fetch()                      # warm-up (3 passes)
    plan.pipeline_device(*d_in[0]); plan.fetch()
    plan.pipeline_host(*h_chunks[0], len(chunks[0]))
    sampler = ClockSampler(local).start_and_wait()
    barrier()
    sampler.mark()            # keep only the samples taken under load
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    table, sums = run(host=False)
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    barrier()
    t0 = time.perf_counter()
    table_h, sums_h = run(host=True)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    assert np.array_equal(table, table_h), "device-resident and host-buffer passes disagree"
    if rank == 0:
        back = pl.array_to_records(table, H, W, K_CLUSTERS, G)
        m = finish_batch(back)                                        # floats finished in GLOBAL image order
        scores = {k: float(np.sum(m[k]) / total) for k in pl.SUM_KEYS}
        h2d = sum(int(a.numel() + b.numel() * 2 + c.numel() * 4) for a, b, c in h_chunks)
        line = {"metric": METRIC, "value": total / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": len(chunks),
                "warmup": 3, "ms_per_step": ms / max(len(chunks), 1), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[4]: one fixed %d-image synthetic batch (321x481 RGB, bank 4x6, k=8 T=20, 5 ground "
                                       "truths) sharded over the GPUs, image i -> rank i mod N; NCCL all-reduce of the metric sums "
                                       "per %d-image step and one all-gather of the integer records" % (total, B),
                           "total_images": total, "unique_images": U, "images_per_step_per_gpu": B,
                           "images_per_launch": plan.launch_group(B),
                           "l2": "inputs + feature tensor far exceed the 126 MB L2"},
                "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d // max(len(chunks), 1),
                        "d2h_bytes_per_step": int(B * (8 + G * 8 * 8 + 2 * K_CLUSTERS * 4 + G * 4 + 4))},
                "gpu_launches": int(launches), "clocks": clocks, "host_pinning": pinned,
                "records_table_sha256": hashlib.sha256(np.ascontiguousarray(table).tobytes()).hexdigest(),
                "records_table_shape": list(table.shape),
                "dataset_scores": scores,
                "dataset_scores_reduced": {k: float(sums[i] / sums[-1]) for i, k in enumerate(pl.SUM_KEYS)}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def make_data_indices(indices, workers=None):
    """Synthetic images + ground truths for explicit global indices."""
    from concurrent.futures import ProcessPoolExecutor
    from gabor_color_image_segmentation_b200.synth import synth_image, synth_ground_truths
    idx = [int(i) for i in indices]
    workers = workers or min(os.cpu_count() or 1, 16)
    if len(idx) <= 4 or workers <= 1:
        imgs = [synth_image(i, H, W) for i in idx]
        gts = [synth_ground_truths(i, H, W, G) for i in idx]
    else:
        with ProcessPoolExecutor(workers) as ex:
            imgs = list(ex.map(synth_image, idx, chunksize=4))
            gts = list(ex.map(synth_ground_truths, idx, chunksize=4))
    return np.stack(imgs), np.stack(gts)


CONFIGS = {
    # name: (H, W, k, images, colour space, dense bank, BASELINE.json config it belongs to)
    "k16": (321, 481, 16, 64, "rgb", False, "configs[1] shape with k=16"),
    "k32": (1024, 1024, 32, 8, "rgb", False, "configs[3]: 1024x1024 RGB, k-means k=32"),
    "k32_4k": (2160, 3840, 32, 2, "rgb", False, "configs[3]: 3840x2160 RGB, k-means k=32"),
    "dense": (321, 481, 8, 32, "opponent", True, "configs[2]: dense bank 8 scales x 12 orientations on opponent channels (D = 288)"),
    "dense_lab": (321, 481, 8, 32, "lab", True, "configs[2]: dense bank 8 scales x 12 orientations on Lab channels (D = 288)"),
    "normalised": (321, 481, 8, 200, "rgb", False, "configs[1] with per-feature normalisation (DESIGN.md 3.6)"),
    "smoothed": (321, 481, 8, 100, "rgb", False, "configs[1] with feature smoothing 0.5 sigma_s and normalisation"),
    "portrait": (481, 321, 8, 200, "rgb", False, "configs[1] on portrait images (481 rows x 321 columns, as a third of BSDS500)"),
}


def run_config(args):
    """Secondary BASELINE configurations as driver-visible records: one JSON line with the stage times of a
    device-resident pass (CUDA events inside the library, single-stream plan) and each stage's roofline fraction."""
    import torch
    from gabor_color_image_segmentation_b200 import GaborBank, Plan
    from gabor_color_image_segmentation_b200.pipeline import init_indices_for
    from gabor_color_image_segmentation_b200.synth import synth_image, synth_ground_truths
    Hc, Wc, k, B, space, dense, what = CONFIGS[args.config]
    B = args.images if args.images != 200 or args.config in ("normalised", "portrait") else B
    os.environ["GCIS_LANES"] = "1"
    torch.cuda.set_device(0)
    small = max(1, max(Hc, Wc) // 1024)                      # large images: upscaled synthetic content + noise
    rng = np.random.default_rng(5)
    imgs, gts = [], []
    for i in range(min(B, 8)):
        im = synth_image(300 + i, Hc // small, Wc // small)
        gt = synth_ground_truths(300 + i, Hc // small, Wc // small, 1)
        if small > 1:
            im = np.kron(im, np.ones((small, small, 1), np.uint8))[:Hc, :Wc]
            im = (im.astype(np.int16) + rng.integers(-10, 11, im.shape)).clip(0, 255).astype(np.uint8)
            gt = np.kron(gt, np.ones((1, small, small), np.uint16))[:, :Hc, :Wc]
        imgs.append(np.ascontiguousarray(im)); gts.append(np.ascontiguousarray(gt))
    imgs = np.stack([imgs[i % len(imgs)] for i in range(B)]); gts = np.stack([gts[i % len(gts)] for i in range(B)])
    kw = {}
    if dense:
        kw["bank"] = GaborBank.dense()
    if args.config == "normalised":
        kw["normalise"] = True
    if args.config == "smoothed":
        kw.update(normalise=True, smooth=0.5)
    plan = Plan(Hc, Wc, max_batch=B, k=k, iters=ITERS, max_gt=1, colour_space=space, n_lab_cap=64, **kw)
    idx = init_indices_for(range(B), Hc * Wc, k)
    d_img = torch.from_numpy(imgs).cuda(); d_gt = torch.from_numpy(gts.view(np.int16)).cuda(); d_idx = torch.from_numpy(idx).cuda()
    for _ in range(max(args.warmup, 3)):
        plan.pipeline_device(d_img, d_gt, d_idx); plan.fetch()
    sampler = ClockSampler(0).start_and_wait()
    torch.cuda.synchronize()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        plan.pipeline_device(d_img, d_gt, d_idx); plan.fetch()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.summary()
    plan.set_profiling(True)
    plan.pipeline_device(d_img, d_gt, d_idx); plan.fetch()
    st = plan.last_stage_ms()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    fma = measure_fma_peak()
    fp32 = fma.get("ffma_rrr_tflops") or 72.5
    N, D = Hc * Wc, plan.D
    K = 8 if k <= 8 else (16 if k <= 16 else 32)
    km_bytes = B * ITERS * N * D * 4.0
    km_flop = B * ITERS * N * 2.0 * K * D                      # score FMAs actually executed (clusters padded to K)
    t_hbm, t_fp = km_bytes / (hbm * 1e9), km_flop / (fp32 * 1e12)
    km_s = st["kmeans"] * 1e-3
    line = {"metric": METRIC, "value": B / (ms * 1e-3), "unit": "images/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %d images of %dx%d, D=%d, k=%d, T=%d, colour=%s" % (what, B, Hc, Wc, D, k, ITERS, space),
                       "name": args.config, "images_per_launch": plan.launch_group(B), "tensor_core_row_pass": plan.uses_tensor_cores},
            "stage_ms_per_step": st, "clocks": clocks,
            "roofline": {"kernel": "k-means passes", "bound": "hbm" if t_hbm >= t_fp else "fp32",
                         "achieved": km_bytes / km_s / 1e9 if t_hbm >= t_fp else km_flop / km_s / 1e12,
                         "peak": hbm if t_hbm >= t_fp else fp32, "unit": "GB/s" if t_hbm >= t_fp else "TFLOP/s",
                         "frac": max(t_hbm, t_fp) / km_s, "hbm_gbs": km_bytes / km_s / 1e9, "fp32_tflops": km_flop / km_s / 1e12,
                         "traffic": None},
            "gabor_us_per_image": st["gabor"] * 1e3 / B, "fma_peak": fma}
    print(json.dumps(line))
    return 0


def one_image_latency(img, gts, reps=20):
    """BASELINE config 1 through the drop-in surface: one script.py iteration = labels = segmenter(img);
    metrics(img, labels, gts).set_metrics().  Host arrays in, floats out; median wall-clock ms."""
    import torch
    from gabor_color_image_segmentation_b200 import gabor_kmeans_segment, metrics
    gl = list(gts)

    def once():
        t0 = time.perf_counter()
        labels = gabor_kmeans_segment(img, n_clusters=K_CLUSTERS, n_iter=ITERS)
        t1 = time.perf_counter()
        m = metrics(img, labels, gl)
        m.set_metrics()
        t2 = time.perf_counter()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    t = np.array([once() for _ in range(reps)])
    return {"segment": float(np.median(t[:, 0])), "metrics": float(np.median(t[:, 1])),
            "total": float(np.median(t.sum(1))), "reps": reps,
            "what": "gabor_kmeans_segment(img) + metrics(img, labels, 5 ground truths).set_metrics(), one 321x481 image, "
                    "host arrays in and out, median wall clock"}


def measure_fma_peak():
    exe = os.path.join(ROOT, "benchmarks", "fma_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=200, help="images per step per GPU")
    ap.add_argument("--unique", type=int, default=200, help="distinct synthetic images generated per rank (cycled if fewer than --images)")
    ap.add_argument("--no-ref-metrics", action="store_true", help="skip timing the reference's own metrics.py and the strong CPU baseline")
    ap.add_argument("--group", type=int, default=0, help="images per launch group (0 = library default, 64)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --images per GPU per step (default, BASELINE configs[1]); strong: one fixed --total-images batch "
                         "sharded over the GPUs (BASELINE configs[4])")
    ap.add_argument("--total-images", type=int, default=10000, help="size of the fixed batch in strong-scaling mode")
    ap.add_argument("--config", default="headline", choices=["headline"] + sorted(CONFIGS),
                    help="headline = BASELINE configs[1] (the metric's configuration); the others are the secondary configurations")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.scaling == "strong":
        return run_gpu_strong(args)
    if args.config != "headline":
        return run_config(args)
    return run_gpu(args)


def _keep_stdout_for_the_json_line():
    """Native libraries write to file descriptor 1 (NCCL prints its version there): point it at stderr and keep the
    real stdout for Python's prints, i.e. the one JSON line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _keep_stdout_for_the_json_line()
    rc = main()
    sys.stdout.flush()
    sys.exit(rc)
