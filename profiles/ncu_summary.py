"""Summarise .ncu-rep captures (ncu --set full --import-source on) as JSON: the handful of metrics
DESIGN.md / bench.py quote, and the top stall reasons of every captured kernel.

    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [...] > profiles/x_summary.json
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def ncu_csv(path, page, extra=()):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summarize(path):
    r = ncu_csv(path, "raw")
    hdr, units = r[0], r[1]
    kernels = []
    for row in r[2:]:
        name = row[hdr.index("Kernel Name")]
        k = {"kernel": name.split("(")[0], "id": row[hdr.index("ID")]}
        for key in KEYS:
            if key in hdr:
                k[key] = {"value": float(row[hdr.index(key)].replace(",", "")), "unit": units[hdr.index(key)]}
        stalls = sorted(((float(row[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                         for i, h in enumerate(hdr)
                         if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)[:6]
        k["stalls_per_issue"] = {n: round(v, 2) for v, n in stalls}
        kernels.append(k)
    return kernels


if __name__ == "__main__":
    print(json.dumps({p: summarize(p) for p in sys.argv[1:]}, indent=1))
