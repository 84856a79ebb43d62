#!/usr/bin/env python
"""Turns an Nsight Compute report into the per-kernel summary CSV kept under profiles/.

usage: python tools/ncu_summary.py gpurun_out/<capture>.ncu-rep [more.ncu-rep ...] > profiles/<name>_summary.csv

Values are exported in BASE units (`ncu --print-units base`: bytes, nanoseconds, plain counts, percent), one row per
captured launch, so that a column means the same thing in every row (the round-1 summaries took the unit of the first
kernel for all of them).  Columns: the launch, its duration, issue / instruction-cache / occupancy figures, DRAM traffic
and throughput, L1 / L2 hit rates and the main stall reasons per issued instruction."""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration_ns"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_limit_regs_blocks"),
    ("launch__occupancy_limit_shared_mem", "occ_limit_smem_blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("sm__icc_request_hit_rate.pct", "icc_hit_pct"),
    ("dram__bytes_read.sum", "dram_read_bytes"),
    ("dram__bytes_write.sum", "dram_write_bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_throughput_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
]


def main():
    out = csv.writer(sys.stdout)
    out.writerow(["report", "launch_id", "kernel"] + [short for _, short in METRICS] + ["dram_bytes_total"])
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], check=True, stdout=subprocess.PIPE,
                             text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr = rows[0]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            if len(r) != len(hdr):
                continue
            vals = []
            for name, _ in METRICS:
                v = r[idx[name]] if name in idx else ""
                vals.append(v.replace(",", ""))
            try:
                total = float(vals[12]) + float(vals[13])
            except ValueError:
                total = ""
            out.writerow([rep.split("/")[-1], r[idx["ID"]], r[idx["Kernel Name"]].split("(")[0]] + vals + [total])


if __name__ == "__main__":
    main()
