#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + hottest source lines.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("kernel:", data[0][hdr.index("Kernel Name")])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-75s %-12s %s" % (k, units[i], " ".join(r[i] for r in data)))
    # stall reasons
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            v = float(data[0][i])
            if v > 0.15:
                print("stall %-60s %.2f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                             capture_output=True, text=True).stdout
        fname, lines, hdr = "", [], None
        for r in csv.reader(io.StringIO(out)):
            if len(r) == 2 and r[0] == "File Path":
                fname = r[1].split("/")[-1]
            elif r and r[0] == "Line No":
                hdr = r
            elif hdr and len(r) == len(hdr) and r[0] != "":
                lines.append((fname, r))
        cs, cx = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")

        def num(x):
            try:
                return float(x)
            except Exception:
                return 0.0
        tot = sum(num(r[cs]) for _, r in lines) or 1
        totx = sum(num(r[cx]) for _, r in lines) or 1
        print("total samples %d, instructions executed %d" % (tot, totx))
        key = cx if "--by-inst" in sys.argv else cs
        lines.sort(key=lambda fr: -num(fr[1][key]))
        for f, r in lines[:n]:
            print("%5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100 * num(r[cs]) / tot, 100 * num(r[cx]) / totx, f, r[0],
                                                          r[1].strip()[:110]))


if __name__ == "__main__":
    main()
