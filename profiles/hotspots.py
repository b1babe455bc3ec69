"""Per-source-line summary of an ncu report captured with --import-source on: share of executed warp instructions, share
of warp-stall samples, threads active per instruction and shared-memory wavefronts for the lines that matter.
Usage: python profiles/hotspots.py gpurun_out/r1_tables.ncu-rep k_stats k_eightmer k_sketch32 > profiles/r1/source_hotspots.txt"""
import collections
import csv
import subprocess
import sys


def hotspots(rep, kernel, out):
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel, "--print-source", "cuda,sass"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdrs = [r for r in rows if r and r[0] == "Line No"]
    if not hdrs:
        out.write(f"{kernel}: not in {rep}\n\n")
        return
    h = hdrs[0]
    ie, sm, te, wf = (h.index(x) for x in ("Instructions Executed", "# Samples", "Thread Instructions Executed", "L1 Wavefronts Shared"))
    cur = None
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    src = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        key = (cur, ln)
        if r[1].strip():
            src[key] = r[1].strip()[:100]
        try:
            a = agg[key]
            a[0] += int(r[ie] or 0); a[1] += int(r[sm] or 0); a[2] += int(r[te] or 0); a[3] += int(r[wf] or 0)
        except (ValueError, IndexError):
            pass
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    out.write(f"== {kernel}: {tot / 1e6:.1f} M warp instructions (source-mapped), {ts} stall samples; lines with >= 1.5 % of either\n")
    out.write("file                 line  inst%  stall%  thr/inst  smem-wavefronts  source\n")
    for k in sorted(agg):
        v = agg[k]
        if v[0] >= tot * 0.015 or v[1] >= ts * 0.015:
            out.write(f"{(k[0] or '?')[:20]:20s} {k[1]:4d}  {v[0] / tot * 100:5.1f}  {v[1] / ts * 100:6.1f}  {v[2] / max(v[0], 1):8.1f}  {v[3] / 1e6:13.1f} M  {src.get(k, '')}\n")
    out.write("\n")


if __name__ == "__main__":
    for kern in sys.argv[2:]:
        hotspots(sys.argv[1], kern, sys.stdout)
