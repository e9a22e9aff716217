"""gabor_color_image_segmentation_b200 — B200-native (sm_100a) Gabor-bank segmentation and
BSD_metrics evaluation behind the reference's call surface.

    from gabor_color_image_segmentation_b200 import *   # metrics, get_segmentation,
                                                         # get_segment_from_filename (script.py:13-14)
    labels = gabor_kmeans_segment(img)                   # the slic slot (script.py:30)
    m = metrics(img, labels, segments); m.set_metrics(); m.display_metrics()

All pixel work runs in ``libgcis.so`` (hand-written CUDA, C ABI in include/gcis.h); there is
no CPU fallback."""
from .groundtruth import get_segmentation, get_segment_from_filename, pack_ground_truths
from .metrics import metrics, finish_image, find_boundaries
from .engine import GaborBank, Plan, BatchCounts, kmeans_init_indices, label_counts_host
from .segment import gabor_kmeans_segment, slic
from .region_scores import region_scores
from .dataset import evaluate_dataset, print_like_script
from .decode import decode_jpeg_batch, imread_gpu, jpeg_info

__all__ = ["metrics", "get_segmentation", "get_segment_from_filename", "gabor_kmeans_segment", "slic",
           "GaborBank", "Plan", "BatchCounts", "kmeans_init_indices", "label_counts_host",
           "pack_ground_truths", "finish_image", "find_boundaries", "region_scores",
           "evaluate_dataset", "print_like_script", "decode_jpeg_batch", "imread_gpu", "jpeg_info"]
