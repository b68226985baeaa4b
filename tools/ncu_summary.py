"""Summarise ncu output for profiles/: `--set full` report -> per-kernel table; launch-list csv -> time shares.

  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep > profiles/rN_full.md
  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rN_launches.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
           ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
           ("launch__registers_per_thread", "regs"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
           ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("smsp__inst_executed.sum", "warp inst"),
           ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("launch__grid_size", "grid")]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), n) for m, n in METRICS if m in hdr]
    print("| kernel | " + " | ".join(n for _, n in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        cells = []
        for i, _ in cols:
            v = r[i]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {units[i]}".strip() if len(units[i]) < 8 else v)
        print(f"| {name} | " + " | ".join(cells) + " |")


def launches(path):
    text = open(path).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    tot = collections.Counter()
    cnt = collections.Counter()
    unit = ""
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        tot[name] += float(r["Metric Value"].replace(",", ""))
        cnt[name] += 1
        unit = r["Metric Unit"]
    s = sum(tot.values())
    print(f"| kernel | launches | total {unit} | share |")
    print("|---|---|---|---|")
    for k, v in tot.most_common():
        print(f"| {k} | {cnt[k]} | {v:.0f} | {100 * v / s:.1f} % |")


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2])
