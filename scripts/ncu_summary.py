#!/usr/bin/env python
"""Summarise an .ncu-rep (exported with `ncu -i X --page raw --csv`) into the handful of metrics the
roofline discussion needs.  Usage: python scripts/ncu_summary.py raw.csv [substring ...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe", "sm__inst_executed_pipe_tensor", "pipe_tensor_cycles_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct", "lts__t_bytes.sum ",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum ", "sm__cycles_active.avg", "lts__t_sectors_op_atom.sum",
        "lts__t_sectors_op_red.sum", "smsp__average_warp", "issue_stalled", "l1tex__m_xbar2l1tex",
        "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "smsp__issue_active.avg.pct",
        "launch__grid_size", "launch__cluster", "sm__ctas_launched", "smsp__cycles_elapsed.avg.per_second"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    keys = sys.argv[2:] or KEYS
    for r in rows[2:]:
        print("==", r[4][:100], "grid", r[8], "block", r[7])
        for h, u, v in zip(hdr, units, r):
            if any(k.strip() in h for k in keys):
                print(f"  {h[:95]:95s} {v:>18s} {u}")


if __name__ == "__main__":
    main()
