#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into a small, committable text file under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_k_mc_thread.txt

Prints the metrics the roofline discussion in DESIGN.md / bench.py refers to: duration, issue-slot utilisation, pipe
utilisations, occupancy and its limiters, stall reasons, shared-memory bank conflicts, DRAM traffic, and an opcode
histogram of executed warp instructions.
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    report, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    lines = []
    for record in rows[2:]:
        index = {h: i for i, h in enumerate(header)}
        lines.append("kernel: %s  grid %s block %s" % (record[index["Kernel Name"]], record[index.get("Grid Size", 0)], record[index.get("Block Size", 0)]))
        for name in METRICS:
            if name in index:
                lines.append("  %-78s %-16s %s" % (name, units[index[name]], record[index[name]]))
        lines.append("  stall reasons (warps per issue-active cycle):")
        for name in sorted(index):
            if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
                value = float(record[index[name]] or 0)
                if value >= 0.05:
                    lines.append("    %-40s %.3f" % (name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], value))
    src = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        header = rows[1]
        index = {h: i for i, h in enumerate(header)}
        ops = collections.Counter()
        total = 0
        for record in rows[2:]:
            if len(record) <= index["Instructions Executed"]:
                continue
            tokens = record[index["Source"]].split()
            if not tokens:
                continue
            op = tokens[1] if tokens[0].startswith("@") and len(tokens) > 1 else tokens[0]
            try:
                count = int(record[index["Instructions Executed"]] or 0)
            except ValueError:       # a report with several kernels repeats the header rows
                continue
            ops[op.split(".")[0]] += count
            total += count
        lines.append("  executed warp instructions by opcode (total %d, %d SASS lines):" % (total, len(rows) - 2))
        for op, count in ops.most_common(16):
            lines.append("    %-10s %14d  %5.1f%%" % (op, count, 100.0 * count / max(total, 1)))
    with open(out, "w") as handle:
        handle.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
