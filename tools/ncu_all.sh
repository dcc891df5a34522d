#!/bin/bash
# ONE ncu pass over every kernel launch of a warmed 2-frame clip (split-precision engine, eager launches) with the metric list the
# roofline discussion uses, then a per-kernel summary (tools/ncu_by_kernel.py -> profiles/r2_ncu_per_kernel.{json,md}); plus a
# `--set full --import-source on` capture of one launch of the three heaviest kernels.  Run on the GPU box:
#   gpurun --timeout 1500 -- 'bash tools/ncu_all.sh'
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,lts__t_bytes.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"
CMD="python tools/run_clip.py --frames 2 --clips 2 --mode tc3"
timeout 1200 ncu --metrics "$M" --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2_ncu_all.csv $CMD > gpurun_out/r2_ncu_all.log 2>&1
tail -2 gpurun_out/r2_ncu_all.log
python tools/ncu_by_kernel.py gpurun_out/r2_ncu_all.csv gpurun_out/r2_ncu_per_kernel > gpurun_out/r2_ncu_by_kernel.log 2>&1; tail -3 gpurun_out/r2_ncu_by_kernel.log
head -50 gpurun_out/r2_ncu_per_kernel.md
for spec in "conv_tc_3x3|conv_tc_kernel<3, (0|false), 3, (0|false)>|260" "conv_tc_1x1|conv_tc_kernel<3, (0|false), 1, (0|false)>|300" "attn_tc|attn_tc_kernel|14"; do
  label="${spec%%|*}"; rest="${spec#*|}"; skip="${rest##*|}"; regex="${rest%|*}"
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s "$skip" -c 1 \
      -f -o "gpurun_out/r2_ncu_full_$label" $CMD > "gpurun_out/r2_ncu_full_$label.log" 2>&1
  if [ -f "gpurun_out/r2_ncu_full_$label.ncu-rep" ]; then
    ncu -i "gpurun_out/r2_ncu_full_$label.ncu-rep" --page raw --csv > "gpurun_out/r2_ncu_full_$label.csv" 2>/dev/null
    python tools/ncu_summary.py "gpurun_out/r2_ncu_full_$label.csv" "$CMD" "$label, one launch, --set full" > "gpurun_out/r2_ncu_full_$label.json" 2>/dev/null
    echo "full $label: $(head -c 600 gpurun_out/r2_ncu_full_$label.json | tr '\n' ' ')"
  else echo "full $label: no report"; tail -3 "gpurun_out/r2_ncu_full_$label.log"; fi
done
