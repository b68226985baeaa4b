"""Per-kernel numbers bench.py quotes from the committed ncu captures -> profiles/r02_ncu_metrics.json.

  python tools/ncu_metrics.py C3=gpurun_out/r02_prof_C3.ncu-rep:64 C2=...:64 C5=...:8      (name=report:frames_per_launch)

For every kernel of a `ncu --set full` report (last launch of each name): duration, DRAM read / write bytes, L2 and L1 hit
rates, warp instructions, issue utilisation, occupancy; per frame where that makes sense."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = {"gpu__time_duration.sum": "ns", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
     "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "smsp__inst_executed.sum": "warp_inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
     "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct", "launch__registers_per_thread": "regs",
     "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst"}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}


def read(path, frames):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
        k = {}
        for m, short in M.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    k[short] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
                except ValueError:
                    pass
        k["ms"] = k.pop("ns", 0.0) / 1e6
        k["frames_per_launch"] = frames
        k["dram_bytes_per_frame"] = (k.get("dram_rd", 0) + k.get("dram_wr", 0)) / frames
        k["warp_inst_per_frame"] = k.get("warp_inst", 0) / frames
        res[name] = k
    return res


if __name__ == "__main__":
    out = {}
    for arg in sys.argv[1:]:
        name, _, rest = arg.partition("=")
        path, _, frames = rest.partition(":")
        out[name] = read(path, int(frames or 64))
    dst = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")
    json.dump(out, open(dst, "w"), indent=1)
    for cfg, ks in out.items():
        for k, v in ks.items():
            print(cfg, k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
