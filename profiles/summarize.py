"""Turn the ncu reports brought back in gpurun_out/ into the text summaries committed under profiles/<round>/.
Usage: python profiles/summarize.py r2   (reads gpurun_out/r2_*.ncu-rep and gpurun_out/r2_launches.csv; writes
profiles/r2/ncu_full_summary.txt, traffic.json (read by bench.py for roofline.traffic), launches.csv, launch_summary.txt)"""
import collections
import csv
import re
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
] + ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k for k in (
    "long_scoreboard", "short_scoreboard", "mio_throttle", "lg_throttle", "barrier", "not_selected", "wait",
    "math_pipe_throttle", "branch_resolving", "no_instruction")]


def kernel_name(s):  # "void k_stats<1>(EngineView, ...)" -> "k_stats"
    return re.sub(r"<.*>", "", re.sub(r"^void ", "", s.split("(")[0])).strip()


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    tag = sys.argv[1]
    gout = os.path.join(ROOT, "gpurun_out")
    dst = os.path.join(ROOT, "profiles", tag)
    os.makedirs(dst, exist_ok=True)
    seen = set()
    lines = ["ncu --set full --clock-control none --import-source on -k regex:... ; bench.py --steps 1 --warmup 1 --no-cpu-baseline --bgzf-records 0",
             "(cfg 2: one 957 MB batch of 3.30 M records per launch).  Per-launch values; times under ncu are cold-cache and serialised",
             "(compare shares, not absolutes).  First captured launch of every kernel.", ""]
    traffic = {}
    for name in ("tables2", "tables", "frame", "inflate"):
        rep = os.path.join(gout, f"{tag}_{name}.ncu-rep")
        if not os.path.exists(rep):
            continue
        hdr, units, rows = raw(rep)
        kn = hdr.index("Kernel Name")
        for r in rows:
            k = kernel_name(r[kn])
            if k in seen:
                continue
            seen.add(k)
            lines.append("kernel".ljust(90) + k)
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    lines.append(m.ljust(90) + f"{r[i]} {units[i]}")
            lines.append("")
            try:
                rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                traffic[k] = float(r[rd]) * scale.get(units[rd], 1.0) + float(r[wr]) * scale.get(units[wr], 1.0)
            except Exception:
                pass
    open(os.path.join(dst, "ncu_full_summary.txt"), "w").write("\n".join(lines))
    # bench.py reads the per-launch DRAM traffic of the dominant kernel family from here
    fam = {"k_stats": traffic.get("k_stats"), "k_eightmer": traffic.get("k_eightmer"), "k_sketch": traffic.get("k_sketch32v2", traffic.get("k_sketch32")),
           "k_cov": sum(v for k, v in traffic.items() if k.startswith("k_cov_")) or None,
           "per_kernel": traffic,
           "_source": f"dram__bytes_read.sum + dram__bytes_write.sum per launch from gpurun_out/{tag}_tables.ncu-rep (ncu --set full --clock-control none, "
                      f"this round's final build, one 957 MB batch of 3.30 M cfg 2 records per launch); summarised by profiles/summarize.py {tag}"}
    json.dump(fam, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
    # launch list
    lcsv = os.path.join(gout, f"{tag}_launches.csv")
    if os.path.exists(lcsv):
        rows = list(csv.reader(open(lcsv)))
        hdr = None
        agg = collections.OrderedDict()
        keep = []
        for r in rows:
            if "Kernel Name" in r:
                hdr = r
                keep.append(r)
                continue
            if hdr and len(r) == len(hdr):
                keep.append(r)
                d = dict(zip(hdr, r))
                try:
                    v = float(d["Metric Value"].replace(",", ""))
                except ValueError:
                    continue
                v *= {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9}.get(d.get("Metric Unit", "ns"), 1.0)   # -> ns
                a = agg.setdefault(kernel_name(d["Kernel Name"]), [0, 0.0])
                a[0] += 1
                a[1] += v
        with open(os.path.join(dst, "launches.csv"), "w", newline="") as f:
            csv.writer(f).writerows(keep)
        tot = sum(a[1] for a in agg.values())
        out = ["ncu --metrics gpu__time_duration.sum --clock-control none; bench.py --steps 2 --warmup 1 --no-cpu-baseline --bgzf-records 3300000 (cfg 2, 10 M records:",
               "kernel-only passes over 3 resident batches of 3.3 M records, streaming passes over 12 slices of 256 MB, BGZF passes).  Serialised, cold-cache times.", "",
               "%-22s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share")]
        for k, (n, t) in agg.items():
            out.append("%-22s %6d %12.1f %10.1f %7.3f" % (k, n, t / 1e3, t / 1e3 / n, t / tot))
        open(os.path.join(dst, "launch_summary.txt"), "w").write("\n".join(out) + "\n")
        print("\n".join(out))


if __name__ == "__main__":
    main()
