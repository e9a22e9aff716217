"""ctypes binding of libgcis.so (include/gcis.h).  No CPU fallback: if the CUDA library
is missing or no device is present, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCIS_LIB") or os.path.join(_PKG, "libgcis.so")
CSRC = os.path.join(_PKG, "csrc")
SOURCES = ["plan.cu", "gabor.cu", "gabor_tc.cu", "features.cu", "kmeans.cu", "label_metrics.cu", "jpeg.cu", "slic.cu"]

GT_SLOTS = 8
COLOUR = {"rgb": 0, "opponent": 1, "lab": 2}
FEATURE = {"magnitude": 0, "energy": 1}
ST_NEG_LABEL, ST_SEG_OVER, ST_LAB_OVER = 1, 2, 4


class GcisError(RuntimeError):
    pass


class GcisConfig(C.Structure):
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32), ("max_batch", C.c_int32), ("colour_space", C.c_int32),
        ("n_scales", C.c_int32), ("n_orient", C.c_int32),
        ("frequencies", C.POINTER(C.c_double)), ("thetas", C.POINTER(C.c_double)),
        ("bandwidth", C.c_double), ("n_stds", C.c_double),
        ("feature", C.c_int32), ("k", C.c_int32), ("iters", C.c_int32), ("fix_shift", C.c_int32),
        ("max_gt", C.c_int32), ("n_lab_cap", C.c_int32), ("dil_recall", C.c_int32), ("group", C.c_int32),
        ("normalise", C.c_int32), ("smooth", C.c_double),
    ]


def nvcc_command(out=LIB_PATH):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
            "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_PKG, "..", "include", "gcis.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or needs_build():
        subprocess.check_call(nvcc_command())
    return LIB_PATH


_lib = None
_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64

# name -> (restype, argtypes): every symbol include/gcis.h declares
SIGNATURES = {
    "gcis_version": (_i32, []),
    "gcis_last_error": (C.c_char_p, []),
    "gcis_launch_count": (_i64, []),
    "gcis_device_count": (_i32, []),
    "gcis_gabor_half_width": (_i32, [C.c_double] * 4),
    "gcis_gabor_separable": (_i32, [C.c_double] * 4 + [_vp] * 4 + [_i32]),
    "gcis_plan_create": (_i32, [C.POINTER(GcisConfig), C.POINTER(_vp)]),
    "gcis_plan_destroy": (None, [_vp]),
    "gcis_plan_feature_dim": (_i32, [_vp]),
    "gcis_plan_workspace_bytes": (_i64, [_vp]),
    "gcis_plan_launch_group": (_i32, [_vp, _i32]),
    "gcis_plan_uses_tensor_cores": (_i32, [_vp]),
    "gcis_gabor_features": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "gcis_kmeans": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "gcis_feature_affine": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "gcis_segment_device": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "gcis_label_metrics_device": (_i32, [_vp, _vp, _vp] + [_i32] * 7 + [_vp] * 8 + [_vp]),
    "gcis_label_metrics_host": (_i32, [_vp, _vp, _vp] + [_i32] * 7 + [_vp] * 8),
    "gcis_find_boundaries_host": (_i32, [_vp, _i32, _i32, _i32, _vp]),
    "gcis_pipeline_device": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "gcis_pipeline_fetch": (_i32, [_vp, _i32] + [_vp] * 7 + [_vp]),
    "gcis_pipeline_fetch_hist": (_i32, [_vp, _i32, _vp, _vp]),
    "gcis_pipeline_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32] + [_vp] * 7),
    "gcis_jpeg_info": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "gcis_jpeg_coefficients": (_i64, [_vp, _i64, _vp, _i64]),
    "gcis_jpeg_decode_batch": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp]),
    "gcis_slic_host": (_i32, [_vp, _i32, _i32, _i32, C.c_double, _i32, _i32, _i32, _vp]),
    "gcis_plan_set_profiling": (_i32, [_vp, _i32]),
    "gcis_plan_last_stage_ms": (_i32, [_vp, _vp]),
}


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GcisError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(this package has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "gcis") -> int:
    if rc < 0:
        msg = load().gcis_last_error().decode("utf-8", "replace")
        raise GcisError(f"{what} failed ({rc}): {msg}")
    return rc


def launch_count() -> int:
    return int(load().gcis_launch_count())
