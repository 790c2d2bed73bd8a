#!/usr/bin/env python
"""Turn the ncu outputs a GPU visit brought back (gpurun_out/) into the small tracked summaries
under profiles/.  usage: ncu_summary.py <tag> [launches.csv] [full.ncu-rep] [kernel substring]
Writes profiles/<tag>_launches.txt, profiles/<tag>_<kernel>.txt and .json (bench.py reads
dram_bytes_per_launch from the .json for roofline.traffic)."""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches.csv")
rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "fused_full.ncu-rep")
kern = sys.argv[4] if len(sys.argv) > 4 else "k_fused"      # k_fused_step (one warp group) or k_fused_ws (two)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
    H = rows[hdr]; ik, iv, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[hdr + 1:]:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
        v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: every launch of the bench command\n"
                f"# (cold-cache, serialised: use the SHARES, not the absolute times)\n")
        f.write(f"{'kernel':44s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, v in tot.most_common():
            f.write(f"{k:44s} {cnt[k]:8d} {v:12.1f} {v / cnt[k]:10.1f} {100 * v / T:6.1f}%\n")
    print(open(os.path.join(ROOT, "profiles", f"{tag}_launches.txt")).read())

if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    row = next(r for r in rows[2:] if kern in r[hdr.index("Kernel Name")])
    m = {h: (v, u) for h, u, v in zip(hdr, units, row)}
    def g(name, default=None):
        try:
            return float(m[name][0].replace(",", ""))
        except Exception:
            return default
    def unit_scale(name):
        u = m[name][1]
        return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    dr = g("dram__bytes_read.sum") * unit_scale("dram__bytes_read.sum")
    dw = g("dram__bytes_write.sum") * unit_scale("dram__bytes_write.sum")
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "inst_executed", "thread_inst_executed", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
            "sass__inst_executed_global_loads", "sass__inst_executed_global_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    name = "fused_step" if kern == "k_fused" else re.sub(r"\W+", "_", kern)
    with open(os.path.join(ROOT, "profiles", f"{tag}_{name}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, one launch of {row[hdr.index('Kernel Name')]}\n")
        for k in keys:
            if k in m:
                f.write(f"{k:95s} {m[k][0]:>16s} {m[k][1]}\n")
        f.write(f"{'dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch)':95s} {dr + dw:16.0f} byte\n")
    js = {"kernel": row[hdr.index("Kernel Name")], "duration_us": g("gpu__time_duration.sum"), "dram_bytes_per_launch": dr + dw,
          "dram_read_bytes": dr, "dram_write_bytes": dw, "inst_executed": g("inst_executed"),
          "registers": g("launch__registers_per_thread"), "ipc_active": g("sm__inst_executed.avg.per_cycle_active"),
          # map size of the captured launch (bench.py reports roofline.traffic only for a run of the same size)
          "cells_per_launch": int(os.environ.get("NCU_CELLS", 4096 * 4096))}
    json.dump(js, open(os.path.join(ROOT, "profiles", f"{tag}_{name}.json"), "w"), indent=1)
    print(open(os.path.join(ROOT, "profiles", f"{tag}_{name}.txt")).read())
