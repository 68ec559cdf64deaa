#!/usr/bin/env python
"""Key counters of an .ncu-rep (one block per profiled launch): python tools/ncu_summary.py rep [regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
default = r"^(Kernel Name|gpu__time_duration.sum|launch__registers_per_thread|launch__grid_size|launch__block_size|launch__occupancy_limit.*|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|dram__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|l1tex__data_pipe_lsu_wavefronts.sum|sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|sm__cycles_elapsed.max|lts__t_sectors_op_red.sum|lts__t_sectors_op_atom.sum|lts__t_sector_hit_rate.pct|lts__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed|smsp__cycles_active.avg|sm__inst_executed_pipe_lsu.sum|smsp__inst_executed_pipe_uniform.sum|lts__t_sectors_srcunit_tex_op_read.sum|lts__t_requests_srcunit_tex_op_red.sum)$"
pat = pat or re.compile(default)
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if pat.search(h):
            print(f"{h:75s} {r[i]:>22s} {units[i]}")
    print("---")
