#!/usr/bin/env python3
"""Condense gpurun_out/ ncu outputs into small tracked files under profiles/.

  python tools/summarise_profile.py TAG
reads gpurun_out/launches_TAG.csv (ncu --metrics gpu__time_duration.sum launch list) and
gpurun_out/prof_TAG.ncu-rep (ncu --set full capture of the render kernel) and writes
profiles/TAG_launches.csv, profiles/TAG_launch_shares.txt and profiles/TAG_k_render_metrics.txt.
"""
import csv
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "sm__maximum_warps_per_active_cycle_pct"]


def main():
    tag = sys.argv[1]
    go = os.path.join(ROOT, "gpurun_out")
    pd = os.path.join(ROOT, "profiles")
    os.makedirs(pd, exist_ok=True)
    lpath = os.path.join(go, "launches_%s.csv" % tag)
    if os.path.exists(lpath):
        rows = []
        with open(lpath, newline="") as f:
            lines = [ln for ln in f if ln.startswith('"')]
        rd = csv.reader(lines)
        hdr = next(rd)
        ik, iv, ib, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Block Size"), hdr.index("Grid Size")
        for r in rd:
            rows.append((r[ik].split("(")[0].replace("void ", ""), r[ig], r[ib], float(r[iv].replace(",", ""))))
        with open(os.path.join(pd, "%s_launches.csv" % tag), "w") as f:
            f.write("id,kernel,grid,block,gpu_time_ns\n")
            for i, (k, g, b, ns) in enumerate(rows):
                f.write('%d,"%s","%s","%s",%.0f\n' % (i, k, g, b, ns))
        agg = OrderedDict()
        for k, g, b, ns in rows:
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += ns
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(pd, "%s_launch_shares.txt" % tag), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: %d launches, %.3f ms total\n"
                    "# (cold-cache, serialised: compare shares, not absolutes)\n" % (len(rows), tot / 1e6))
            for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("%-40s launches %4d  total %10.3f us  avg %10.3f us  share %5.1f %%\n"
                        % (k, n, ns / 1e3, ns / 1e3 / n, 100 * ns / tot))
    rep = os.path.join(go, "prof_%s.ncu-rep" % tag)
    if os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader([ln for ln in out.splitlines() if ln.startswith('"')]))
        hdr, units = rows[0], rows[1]
        with open(os.path.join(pd, "%s_k_render_metrics.txt" % tag), "w") as f:
            f.write("# ncu --set full --clock-control none --import-source on -k regex:k_render (one launch each)\n")
            for vals in rows[2:]:
                name = vals[hdr.index("Kernel Name")]
                f.write("kernel: %s grid %s block %s\n" % (name[:120], vals[hdr.index("Grid Size")], vals[hdr.index("Block Size")]))
                for i, h in enumerate(hdr):
                    if h in KEYS:
                        f.write("  %-62s %-14s %s\n" % (h, units[i], vals[i]))
        import json
        kr = [v for v in rows[2:] if "k_render" in v[hdr.index("Kernel Name")]]
        if kr:
            v = kr[-1]
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(v[hdr.index("dram__bytes_read.sum")]) * unit[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(v[hdr.index("dram__bytes_write.sum")]) * unit[units[hdr.index("dram__bytes_write.sum")]]
            json.dump({"traffic_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                       "kernel": v[hdr.index("Kernel Name")].split("(")[0],
                       "source": "profiles/%s_k_render_metrics.txt (ncu --set full, one launch)" % tag},
                      open(os.path.join(pd, "%s_traffic.json" % tag), "w"))
    print("profiles written for", tag)


if __name__ == "__main__":
    main()
