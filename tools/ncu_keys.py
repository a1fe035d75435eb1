"""Print the handful of ncu raw-page metrics we track, one column per profiled launch."""
import csv, subprocess, sys
KEYS = ['Kernel Name','Grid Size','gpu__time_duration.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__inst_executed_pipe_fp64','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active',
 'sm__pipe_tensor_op_dmma','smsp__inst_executed_pipe_tensor','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__cycles_active.avg','smsp__cycles_active.avg',
 'stalled_barrier_per_issue','stalled_long_scoreboard_per_issue','stalled_short_scoreboard_per_issue','stalled_math_pipe_throttle_per_issue',
 'stalled_wait_per_issue','stalled_mio_throttle_per_issue','stalled_not_selected_per_issue','stalled_dispatch_stall_per_issue','stalled_lg_throttle_per_issue','stalled_no_instruction_per_issue','stalled_tensor']
out = subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for ci,(h,u) in enumerate(zip(hdr,units)):
    if any(k in h for k in KEYS) and 'pcsamp' not in h:
        print(f'{h[:95]:95s} {u[:12]:12s} ' + ' '.join(f'{r[ci][:18]:>18s}' for r in rows[2:]))
