#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/run_clip.py --frames 2 --clips 1 --mode tc3"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_cin3_px4_kernel|conv_cout" -c 6 -f -o gpurun_out/r2_ncu_full_stem $CMD > gpurun_out/r2_ncu_full_stem.log 2>&1
ncu -i gpurun_out/r2_ncu_full_stem.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_stem.csv 2>/dev/null
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r2_ncu_full_stem.csv')))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio']
idx=[hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print('----')
    for i in idx: print(hdr[i][:70], '=', r[i][:90], units[i])
P
