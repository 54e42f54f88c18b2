#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into a small JSON: duration, DRAM bytes, stall ratios, pipe use per kernel.
usage: ncu_summary.py report.ncu-rep [out.json]"""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "sm__ops_path_tensor_src_fp64.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    res = []
    for row in rows[2:]:
        d = dict(zip(h, row))
        k = {"kernel": d["Kernel Name"][:70], "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for key in KEYS:
            if key in d and d[key] != "":
                k[key] = d[key] + " " + rows[1][h.index(key)]
        for key in h:
            if key.startswith("smsp__average_warps_issue_stalled_") and key.endswith("_per_issue_active.ratio"):
                v = float(d[key] or 0)
                if v >= 0.05:
                    k.setdefault("stalls_per_issue", {})[key[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
        res.append(k)
    txt = json.dumps({"report": rep, "kernels": res}, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
