#!/usr/bin/env python
"""Text summary of ncu --set full captures for profiles/: the metrics DESIGN.md argues with, per report, plus the
hot source lines (scripts/ncu_lines.py) when a cubin with line info is given.
Usage: scripts/ncu_summary.py <title> <report.ncu-rep>[:cubin:kernel-substring] ... > profiles/ncu_summary_<tag>.txt
       scripts/ncu_summary.py --traffic <name>=<report.ncu-rep> ...            > profiles/ncu_traffic_<tag>.json
"""
import csv, io, json, os, subprocess, sys

WANT = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALLS = ["no_instruction", "wait", "barrier", "math_pipe_throttle", "short_scoreboard", "long_scoreboard", "branch_resolving", "not_selected",
          "mio_throttle", "dispatch_stall", "lg_throttle", "membar", "sleeping"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    if sys.argv[1] == "--traffic":
        res = {"kernels": {}}
        for a in sys.argv[2:]:
            name, rep = a.split("=")
            m = raw(rep)

            def by(k):
                v, u = m[k]
                return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            res["kernels"][name] = {"dram_bytes_per_launch": int(by("dram__bytes_read.sum") + by("dram__bytes_write.sum")),
                                    "kernel": m["Kernel Name"][0], "gpu_time_ms_under_ncu": float(m["gpu__time_duration.sum"][0]),
                                    "source": os.path.basename(rep) + ": dram__bytes_read.sum + dram__bytes_write.sum of one launch"}
        print(json.dumps(res, indent=1))
        return
    print("# " + sys.argv[1])
    here = os.path.dirname(os.path.abspath(__file__))
    for a in sys.argv[2:]:
        parts = a.split(":")
        rep = parts[0]
        m = raw(rep)
        print("\n## " + os.path.basename(rep).replace(".ncu-rep", ""))
        for k in WANT:
            if k in m:
                print(f"{k:90s} {m[k][1]:>14s} {m[k][0]}")
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in m:
                print(f"{k:90s} {m[k][1]:>14s} {m[k][0]}")
        if len(parts) == 3:
            print(f"\n### hot source lines, {parts[2]}")
            print(subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), rep, parts[1], parts[2], "24"], capture_output=True, text=True).stdout)


if __name__ == "__main__":
    main()
