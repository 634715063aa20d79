#!/usr/bin/env python3
"""Per-config ncu captures -> profiles/r02_kernels.json (read by bench.py: roofline.traffic and roofline.secondary)
and profiles/r02_<config>_metrics.txt.

  python tools/summarise_kernels.py TAG config=gpurun_out/prof_X.ncu-rep[:kernel-regex] ...

Each report holds `ncu --set full --clock-control none --import-source on` launches of one config's render kernel;
the LAST launch whose name matches is summarised (per launch: DRAM bytes, instructions, L2 hit rate, sectors per
request, pipe utilisation).
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__shared_mem_per_block_static"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    tag = sys.argv[1]
    pd = os.environ.get("EU_PROFILE_DIR") or os.path.join(ROOT, "profiles")
    os.makedirs(pd, exist_ok=True)
    kpath = os.path.join(pd, "%s_kernels.json" % tag)
    allk = json.load(open(kpath)) if os.path.exists(kpath) else {}
    for spec in sys.argv[2:]:
        config, rest = spec.split("=", 1)
        rep, _, rx = rest.partition(":")
        if not os.path.exists(rep):
            print("missing", rep)
            continue
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader([ln for ln in out.splitlines() if ln.startswith('"')]))
        if len(rows) < 3:
            print("empty", rep)
            continue
        hdr, units = rows[0], rows[1]
        ik = hdr.index("Kernel Name")
        cand = [v for v in rows[2:] if re.search(rx or "k_render", v[ik])]
        if not cand:
            print("no kernel matching", rx, "in", rep)
            continue
        v = cand[-1]

        def get(name, scale=False):
            if name not in hdr:
                return None
            i = hdr.index(name)
            x = num(v[i])
            if x is not None and scale:
                x *= UNIT.get(units[i], 1.0)
            return x
        rd, wr = get("dram__bytes_read.sum", True), get("dram__bytes_write.sum", True)
        req, sec = get("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        rel = os.path.join("profiles", "%s_%s_metrics.txt" % (tag, config.replace("/", "_")))
        allk[config] = {"kernel": v[ik].split("(")[0].replace("void ", ""), "dram_bytes": (rd or 0) + (wr or 0),
                        "dram_read_bytes": rd, "dram_write_bytes": wr, "inst_executed": get("smsp__inst_executed.sum"),
                        "time_us_under_ncu": (get("gpu__time_duration.sum") or 0.0) *
                        {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[hdr.index("gpu__time_duration.sum")], 1.0), "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
                        "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct"),
                        "global_ld_sectors_per_request": (sec / req) if req and sec else None,
                        "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "pipe_fma_pct": get("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active"),
                        "pipe_alu_pct": get("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
                        "pipe_xu_pct": get("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active"),
                        "pipe_lsu_pct": get("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active"),
                        "l1tex_throughput_pct": get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                        "dram_throughput_pct": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        "registers": get("launch__registers_per_thread"),
                        "source": rel + " (ncu --set full --clock-control none, one launch)"}
        with open(os.path.join(pd, os.path.basename(rel)), "w") as f:
            f.write("# ncu --set full --clock-control none --import-source on (one launch of the config's render kernel)\n")
            f.write("kernel: %s grid %s block %s\n" % (v[ik][:140], v[hdr.index("Grid Size")], v[hdr.index("Block Size")]))
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write("  %-62s %-14s %s\n" % (h, units[i], v[i]))
        print(config, allk[config]["kernel"], "time", allk[config]["time_us_under_ncu"], "us")
    json.dump(allk, open(kpath, "w"), indent=1)


if __name__ == "__main__":
    main()
