#!/usr/bin/env python
"""Print the metrics we track from an .ncu-rep (raw page) - used to write profiles/*.md."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','launch__occupancy_limit_registers',
 'launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active',
 'dram__bytes_write.sum','dram__bytes_read.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.sum','smsp__inst_executed_pipe_fp64.sum',
 'sm__cycles_elapsed.avg','smsp__cycles_active.avg','sm__cycles_active.avg',
 'smsp__average_warp_latency_issue_stalled_wait.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_pipe_lsu.sum','lts__t_bytes.sum',
 'sm__cycles_elapsed.avg.per_second']
def main(path, extra):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('kernel:', r[hdr.index('Kernel Name')][:110])
        for i,h in enumerate(hdr):
            if h in WANT or any(e in h for e in extra):
                print(f'  {h:80s} {r[i]} {units[i]}')
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
