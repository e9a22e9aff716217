"""Shard invariance on real GPUs (BASELINE config 5): the gathered integer records of a
batch sharded over 2 GPUs (one process per GPU, NCCL) equal the 1-GPU table bit for bit, and so do the dataset
scores finished from it in global image order.  Needs two devices; skipped otherwise (the CPU twin with gloo is
tests/test_multigpu_host.py)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_IMAGES, H, W, G, K, T = 13, 120, 160, 3, 5, 8     # 13 images on 2 ranks: uneven shards


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _table(rank, world, dev):
    import torch
    from gabor_color_image_segmentation_b200 import Plan
    from gabor_color_image_segmentation_b200 import pipeline as pl
    from gabor_color_image_segmentation_b200.synth import synth_image, synth_ground_truths
    idx = pl.shard_indices(N_IMAGES, rank, world)
    imgs = np.stack([synth_image(int(i), H, W) for i in idx])
    gts = np.stack([synth_ground_truths(int(i), H, W, G) for i in idx])
    plan = Plan(H, W, max_batch=4, k=K, iters=T, max_gt=G)          # several chunks per rank
    c = pl.evaluate_batch(plan, imgs, gts, pl.init_indices_for(idx, H * W, K))
    sums = pl.reduce_sums(pl.metric_sums(c), dev)
    return pl.gather_records(pl.records_to_array(c), idx, N_IMAGES, dev), sums


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    table, sums = _table(rank, world, dev)
    if rank == 0:
        np.save(out, {"table": table, "sums": sums}, allow_pickle=True)
    dist.destroy_process_group()


def test_two_gpu_records_equal_one_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from gabor_color_image_segmentation_b200 import pipeline as pl
    out = str(tmp_path / "r0.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out, allow_pickle=True).item()
    torch.cuda.set_device(0)
    one, sums1 = _table(0, 1, torch.device("cuda", 0))
    np.testing.assert_array_equal(got["table"], one)
    a = pl.metric_sums(pl.array_to_records(got["table"], H, W, K, G))
    b = pl.metric_sums(pl.array_to_records(one, H, W, K, G))
    np.testing.assert_array_equal(a, b)                               # floats finished in global order: bit-equal
    np.testing.assert_allclose(got["sums"], sums1, rtol=1e-13)        # NCCL-summed per-rank sums: order differs
    assert got["sums"][-1] == N_IMAGES
