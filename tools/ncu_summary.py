#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (raw page as CSV) into one block per launch:
  ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv [title]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
        ("gpu__time_duration.sum", "duration"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (legacy HMMA counter)"),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "xbar->L1 bytes"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct", "issue active %"),
        ("launch__registers_per_thread", "regs/thread"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall long scoreboard %"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ]
tens = [h for h in hdr if "tensor" in h and "pct" in h]
print(f"== {title}: {len(rows) - 2} launches (ncu --set full --clock-control none; cold-cache, serialised replays)")
for r in rows[2:]:
    print("-" * 100)
    for key, label in want:
        if key in hdr:
            i = hdr.index(key)
            print(f"  {label:36s} {r[i][:90]} {units[i]}")
    for key in tens[:6]:
        i = hdr.index(key)
        print(f"  {key[:70]:70s} {r[i]} {units[i]}")
