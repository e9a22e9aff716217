"""Recipe for oracle/_ref/: a byte-for-byte copy of the reference's two modules on the metrics path,
so that the GPU box (which has no /root/reference) can TIME the reference's own code as the CPU baseline
(bench.py: cpu_baseline.reference_metrics, cpu_baseline.strong).  TEST / BENCH INFRASTRUCTURE.

    python oracle/fetch_ref.py          # also called by __graft_entry__.build() when /root/reference exists

oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so it travels
with the snapshot like a built .so.  The files are used unmodified, imported over oracle/ref_shim (the scipy
stand-in for scikit-image / matplotlib)."""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/BSD_metrics"
DST = os.path.join(HERE, "_ref")
FILES = ["metrics.py", "groundtruth.py"]


def fetch() -> bool:
    """Copy the reference modules if the reference checkout is present; True when oracle/_ref is populated."""
    if os.path.isdir(SRC):
        os.makedirs(DST, exist_ok=True)
        for f in FILES:
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


if __name__ == "__main__":
    print("oracle/_ref populated:", fetch())
