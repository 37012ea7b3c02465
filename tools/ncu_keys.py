#!/usr/bin/env python
"""Print the key metrics of an .ncu-rep (run here, no GPU): python tools/ncu_keys.py file.ncu-rep [more-substrings...]"""
import csv, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
for v in rows[2:]:
    for i, n in enumerate(h):
        if n in want or any(e in n for e in extra):
            print("%-90s %s %s" % (n, v[i], u[i]))
    print("-" * 40)
