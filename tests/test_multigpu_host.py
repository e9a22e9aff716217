"""N>1 host logic on CPU (gloo, world_size 2): sharding, reduction of metric sums and the
all-gather of integer records.  The per-image engine is stubbed with the CPU oracle here
(tests may use it as the checker); on the GPU box the same code runs over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _records_via_oracle(indices, H, W, G, k):
    """Integer records for the given global image indices, computed by the oracle."""
    sys.path.insert(0, ROOT)
    from gabor_color_image_segmentation_b200.engine import BatchCounts
    from gabor_color_image_segmentation_b200.synth import voronoi_labels, synth_ground_truths
    from oracle import oracle as orc
    B = len(indices)
    gc = np.zeros((B, G, 8), np.int64); bd = np.zeros(B, np.int64)
    area = np.zeros((B, k), np.int32); perim = np.zeros((B, k), np.int32); n_seg = np.zeros(B, np.int32)
    for n, i in enumerate(indices):
        lb = voronoi_labels(np.random.default_rng(int(i)), H, W, k)
        o = orc.label_counts(lb, list(synth_ground_truths(int(i), H, W, G)))
        bd[n] = o.bd_count; n_seg[n] = o.n_seg
        gc[n, :, 0], gc[n, :, 1], gc[n, :, 2], gc[n, :, 3], gc[n, :, 4] = o.den_r, o.tp_r, o.tp_p, o.U, o.V
        area[n, :o.n_seg] = o.area; perim[n, :o.n_seg] = o.perim
    return BatchCounts(H, W, bd, gc, area, perim, n_seg, np.zeros((B, G), np.int32), np.zeros(B, np.int32),
                       np.full(B, G, np.int32))


def _worker(rank, world, port, n_images, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gabor_color_image_segmentation_b200 import pipeline as pl
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W, G, k = 40, 56, 2, 5
    idx = pl.shard_indices(n_images, rank, world)
    c = _records_via_oracle(idx, H, W, G, k)
    sums = pl.reduce_sums(pl.metric_sums(c))
    table = pl.gather_records(pl.records_to_array(c), idx, n_images)
    if rank == 0:
        np.save(out, {"sums": sums, "table": table}, allow_pickle=True)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [7, 1])   # 1 image on 2 ranks: rank 1 owns an EMPTY shard and must still reach the collectives
def test_two_rank_reduction_equals_single_process(tmp_path, n_images):
    sys.path.insert(0, ROOT)
    from gabor_color_image_segmentation_b200 import pipeline as pl
    H, W, G, k = 40, 56, 2, 5
    out = str(tmp_path / "r0.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_images, out), nprocs=2, join=True)
    got = np.load(out, allow_pickle=True).item()
    c = _records_via_oracle(np.arange(n_images), H, W, G, k)
    table = pl.records_to_array(c)
    np.testing.assert_array_equal(got["table"], table)               # integer records: identical
    one = pl.metric_sums(c)
    np.testing.assert_allclose(got["sums"], one, rtol=1e-14)         # float sums: order differs by rank
    assert got["sums"][-1] == n_images
    # finishing floats from the gathered table in image order is shard-independent, bit for bit
    back = pl.array_to_records(got["table"], H, W, k, G)
    np.testing.assert_array_equal(pl.metric_sums(back), one)
    assert set(pl.dataset_scores(one)) == {"recall", "precision", "underseg", "undersegNP", "compactness",
                                           "density", "images"}


def test_shard_indices_partition():
    sys.path.insert(0, ROOT)
    from gabor_color_image_segmentation_b200.pipeline import shard_indices, init_indices_for
    for n in (0, 1, 7, 200):
        for world in (1, 2, 4, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            allv = np.sort(np.concatenate(parts)) if n else np.zeros(0, np.int64)
            np.testing.assert_array_equal(allv, np.arange(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    a = init_indices_for([3, 5], 1000, 4, seed=9)
    b = init_indices_for([5], 1000, 4, seed=9)
    np.testing.assert_array_equal(a[1], b[0])           # seeded by global index, not by position
