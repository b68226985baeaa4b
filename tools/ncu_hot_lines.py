"""Per-source-line hot spots of one kernel from an `ncu --set full --import-source on` report.

  python tools/ncu_hot_lines.py gpurun_out/prof.ncu-rep k_shade [top=40] [cubin=k_render]

ncu's CSV source page carries metrics only in the SASS view, so the SASS rows (address, stall samples,
executed instructions) are joined by instruction offset with `nvdisasm -g` line info of the cubin
extracted from stillleben_b200/libslb.so (must be the build that was profiled).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_map(cubin_name, kern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "stillleben_b200", "libslb.so")], cwd=tmp, capture_output=True)
    path = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.startswith(cubin_name)][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", path], capture_output=True, text=True).stdout
    m, cur, inside = {}, ("?", 0), False
    for ln in out.splitlines():
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        f = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if f:
            cur = (os.path.basename(f.group(1)), int(f.group(2)))
            continue
        a = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if a:
            m[int(a.group(1), 16)] = cur
    return m


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cubin = sys.argv[4] if len(sys.argv) > 4 else "k_render"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    body = [r for r in rows if r and r[0].startswith("0x") and len(r) >= len(hdr)]
    base = int(body[0][0], 16)
    lm = line_map(cubin, kern)
    ci_s, ci_i = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    agg = collections.defaultdict(lambda: [0, 0])
    tot_s = tot_i = 0
    for r in body:
        off = int(r[0], 16) - base
        key = lm.get(off, ("?", 0))
        s, i = int(r[ci_s] or 0), int(r[ci_i] or 0)
        agg[key][0] += s
        agg[key][1] += i
        tot_s += s
        tot_i += i
    src_cache = {}

    def src(f, line):
        if f not in src_cache:
            p = os.path.join(ROOT, "stillleben_b200", "csrc", f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        L = src_cache[f]
        return L[line - 1].strip()[:100] if 0 < line <= len(L) else ""

    print(f"kernel {kern}: {tot_s} stall samples, {tot_i} executed warp instructions, {len(body)} SASS instructions")
    print()
    print("| samples % | inst % | file:line | source |")
    print("|---|---|---|---|")
    for (f, line), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {100.0 * s / max(1, tot_s):.1f} | {100.0 * i / max(1, tot_i):.1f} | {f}:{line} | `{src(f, line)}` |")


if __name__ == "__main__":
    main()
