"""Condenses an `ncu --page raw --csv` dump into one markdown row per kernel launch.
usage: python profiles/ncu_summary.py raw.csv [more.csv ...] > summary.md"""
import csv
import sys

COLS = [
    ("time", "gpu__time_duration.sum"),
    ("DRAM rd", "dram__bytes_read.sum"),
    ("DRAM wr", "dram__bytes_write.sum"),
    ("DRAM %", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("fp64 %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("occ %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("warp insts", "smsp__inst_executed.sum"),
    ("elig/cyc", "smsp__warps_eligible.avg.per_cycle_active"),
]


def fmt(v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if u == "byte":
        return "%.2f MB" % (x / 1e6)
    if u == "Kbyte":
        return "%.2f MB" % (x / 1e3)
    if u == "Mbyte":
        return "%.2f MB" % x
    if u == "Gbyte":
        return "%.2f MB" % (x * 1e3)
    if u == "ns":
        return "%.2f us" % (x / 1e3)
    if u in ("us", "usecond"):
        return "%.2f us" % x
    if u in ("ms", "msecond"):
        return "%.2f us" % (x * 1e3)
    if x >= 1e6:
        return "%.3g" % x
    return "%.4g" % x


def main():
    print("| kernel | grid | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = {n: i for i, n in enumerate(hdr)}
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            name = r[idx["Kernel Name"]].replace("<unnamed>::", "").replace("void ", "")
            name = name.split("(")[0][:60]
            cells = []
            for _, m in COLS:
                i = idx.get(m)
                cells.append(fmt(r[i], units[i]) if i is not None else "-")
            print("| %s | %s | %s |" % (name, r[idx["Grid Size"]].strip("()").split(",")[0], " | ".join(cells)))


if __name__ == "__main__":
    main()
