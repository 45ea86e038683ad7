#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) into the handful of numbers DESIGN.md / bench.py quote:
duration, DRAM bytes, issue-slot / pipe utilisation, occupancy, stall mix.  Usage: ncu_summary.py report.ncu-rep [out.md]"""
import csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy (elapsed)"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (active)"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor pipe instructions"), ("sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (UMMA) pipe active"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"), ("launch__registers_per_thread", "registers/thread"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp instructions"), ("sm__icc_request_hit_rate.pct", "instruction-cache hit rate"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (smem)"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global)"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle")]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out.append("### %s  grid %s block %s\n" % (d["Kernel Name"], d.get("Grid Size"), d.get("Block Size")))
        out.append("| metric | value | unit |\n|---|---|---|")
        for k, label in KEYS:
            if k in d and d[k] != "":
                out.append("| %s (`%s`) | %s | %s |" % (label, k, d[k], u[k]))
        out.append("")
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "a").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
