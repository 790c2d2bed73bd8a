#!/usr/bin/env python
"""Summarise an `ncu --set full` report that holds several kernels (scripts/gpu_all_kernels.sh) into
profiles/<tag>_all_kernels.txt: one line per kernel (its last captured launch = steady state).
usage: ncu_kernels_summary.py <tag> [raw-page .csv | report.ncu-rep]"""
import csv, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "all_kernels_raw.csv")
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}

def val(r, name):
    try:
        v = float(r[col[name]].replace(",", ""))
    except Exception:
        return float("nan")
    u = units[col[name]]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6, "s": 1e9,
                "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1.0)

last, count = {}, {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    last[name] = r
    count[name] = count.get(name, 0) + 1

out = os.path.join(ROOT, "profiles", f"{tag}_all_kernels.txt")
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none, scripts/all_kernels.py: the last captured launch of every kernel\n"
            "# (grid kernels on 4096^2, droplet kernels with 4 Mi droplets on 8192^2; times are single cold-ish launches under\n"
            "#  the profiler - use them for the bound each kernel sits on, not as bench values)\n"
            "# dram% = gpu__dram_throughput pct of peak; issue% = smsp__issue_active pct of peak; occ% = achieved warps active\n")
    f.write(f"{'kernel':40s} {'n':>3s} {'time_us':>9s} {'grid':>8s} {'blk':>4s} {'regs':>4s} {'occ%':>5s} {'dramMB':>8s} {'GB/s':>7s} "
            f"{'dram%':>6s} {'issue%':>6s} {'L2hit%':>6s} {'L1hit%':>6s}\n")
    for name, r in sorted(last.items(), key=lambda kv: -val(kv[1], "gpu__time_duration.sum")):
        t = val(r, "gpu__time_duration.sum")      # ns
        db = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        f.write(f"{name[:40]:40s} {count[name]:3d} {t / 1e3:9.1f} {int(val(r, 'launch__grid_size')):8d} {int(val(r, 'launch__block_size')):4d} "
                f"{int(val(r, 'launch__registers_per_thread')):4d} {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} "
                f"{db / 1e6:8.1f} {db / t:7.0f} {val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
                f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} {val(r, 'lts__t_sector_hit_rate.pct'):6.1f} "
                f"{val(r, 'l1tex__t_sector_hit_rate.pct'):6.1f}\n")
print(open(out).read())
