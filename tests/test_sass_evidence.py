"""The built library must contain the Blackwell instructions DESIGN.md claims (no GPU needed: cuobjdump reads the
in-tree libgcis.so).  Guards against a change that silently loses them, e.g. the compiler moving the filter bank's
column taps from uniform registers back into ordinary ones (profiles/r02_gabor_tc.md)."""
import re
import shutil
import subprocess

import pytest


@pytest.fixture(scope="module")
def sass_by_kernel():
    from gabor_color_image_segmentation_b200 import _lib
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-sass", _lib.build()], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError) as e:   # no CUDA toolkit on this machine
        pytest.skip("cuobjdump not usable: %s" % e)
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
        elif cur is not None and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
            cur.append(line)
    return kernels


def _kernel(kernels, *needles):
    hits = [v for k, v in kernels.items() if all(n in k for n in needles)]
    assert hits, "no kernel matching %r" % (needles,)
    return hits[0]


def _count(lines, pattern):
    return sum(1 for l in lines if re.search(pattern, l))


def test_filter_bank_kernel_uses_tcgen05_tma_and_uniform_taps(sass_by_kernel):
    k = _kernel(sass_by_kernel, "gabor_tc_kernel", "ILb0E")
    assert _count(k, r"\bUTC[A-Z]*MMA") >= 1          # tcgen05.mma
    assert _count(k, r"\bLDTM") >= 1                  # tcgen05.ld
    assert _count(k, r"\bUTMALDG") >= 1               # TMA tensor loads
    ffma2, uniform = _count(k, r"\bFFMA2\b"), _count(k, r"\bFFMA2\b.*\bUR\d+")
    assert ffma2 >= 1000
    assert uniform >= 0.7 * ffma2, "column taps are no longer uniform-register operands (%d of %d FFMA2)" % (uniform, ffma2)


def test_kmeans_tile_kernel_uses_tma_and_integer_tensor_cores(sass_by_kernel):
    k = _kernel(sass_by_kernel, "km_tile_kernel", "ILi8ELi256ELi2")
    assert _count(k, r"\bUTMALDG") >= 1               # TMA tensor boxes
    assert _count(k, r"\bSYNCS") >= 1                 # mbarriers
    assert _count(k, r"\bFFMA2\b") >= 32              # packed fp32 score chain
    assert _count(k, r"\bIMMA") >= 4                  # exact integer update GEMM
