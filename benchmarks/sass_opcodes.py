"""Per-kernel counts of the SASS opcodes that prove (or disprove) a Blackwell-native kernel, from the built
library (runs in the build container, no GPU needed):

    python benchmarks/sass_opcodes.py > profiles/r02_sass_opcodes.md

UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, UBLKCP = bulk copy, SYNCS = mbarrier,
FFMA2 = packed fp32 FMA, IMMA/HMMA = legacy warp-level tensor-core MMA (mma.sync)."""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PATTERNS = [("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM", r"\bSTTM"),
            ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"), ("UTMALDG (TMA tensor load)", r"\bUTMALDG"), ("UBLKCP (bulk copy)", r"\bUBLKCP"),
            ("SYNCS (mbarrier)", r"\bSYNCS"), ("FFMA2", r"\bFFMA2"), ("FFMA2 with a uniform-register pair operand", r"\bFFMA2\b.*\bUR\d+"),
            ("LDCU (uniform constant load)", r"\bLDCU"), ("FFMA", r"\bFFMA\b"), ("IMMA (mma.sync s8)", r"\bIMMA"),
            ("HMMA (mma.sync f16)", r"\bHMMA"), ("ACQBULK/PDL", r"\bACQBULK|\bPREEXIT"), ("total instructions", r"^\s+/\*[0-9a-f]{4,}\*/")]


def main():
    from gabor_color_image_segmentation_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|gcis::", "", name).split("(")[0]
            cur = counts.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        for label, pat in PATTERNS:
            if re.search(pat, line):
                cur[label] += 1
    labels = [l for l, _ in PATTERNS]
    print("# SASS opcode counts per kernel of `libgcis.so` (`cuobjdump -sass`, sm_100a)\n")
    print("| kernel | " + " | ".join(labels) + " |")
    print("|---|" + "---|" * len(labels))
    for name, c in counts.items():
        print("| `%s` | " % name + " | ".join(str(c.get(l, 0)) for l in labels) + " |")


if __name__ == "__main__":
    main()
