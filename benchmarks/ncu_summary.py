"""Summarise an ncu report (run in the build container: `ncu -i` needs no GPU).
    python benchmarks/ncu_summary.py gpurun_out/km_prof.ncu-rep > profiles/r01_km_pass.md"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (shared memory)"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu summary of `%s` (`ncu --set full --clock-control none`; per launch, cold cache, serialised)\n" % rep.split("/")[-1])
    print("| launch | kernel | " + " | ".join(n for _, n in METRICS if _ in idx) + " |")
    print("|---|---|" + "---|" * sum(1 for m, _ in METRICS if m in idx))
    for n, r in enumerate(data):
        cells = []
        for m, _ in METRICS:
            if m in idx:
                v = r[idx[m]]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                cells.append(v + " " + units[idx[m]])
        print("| %d | %s | %s |" % (n, r[idx["Kernel Name"]].split("(")[0][-40:], " | ".join(cells)))


if __name__ == "__main__":
    main()
