#!/bin/bash
# refresh of profiles/r2_ncu_per_kernel.* and profiles/r2_sanitizer.md on the final build (one GPU, ~20 min)
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size"
CMD="python tools/run_clip.py --frames 2 --clips 1 --mode tc3"
( time timeout 1100 ncu --metrics "$M" --clock-control none --kernel-name-base demangled --launch-skip 650 --launch-count 1200 --csv --log-file gpurun_out/r2_ncu_all.csv $CMD ) > gpurun_out/r2_ncu_all.log 2>&1
tail -4 gpurun_out/r2_ncu_all.log
python tools/ncu_by_kernel.py gpurun_out/r2_ncu_all.csv gpurun_out/r2_ncu_per_kernel > gpurun_out/r2_ncu_by_kernel.log 2>&1; tail -3 gpurun_out/r2_ncu_by_kernel.log
head -50 gpurun_out/r2_ncu_per_kernel.md | cut -c1-200
for tool in memcheck racecheck; do
  ( time timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -q -m gpu -x \
      -k "3x3_64_64 or linear_splitk or (fused_attention_matches_torch and kvpack and 3-256) or (fused_window and kvpack and 16) or split_16sq_512 or split_32sq_256_n2 or epi_ragged or out_proj_256x512 or lockstep4_1024" ) > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool: $(grep -E 'ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|real' gpurun_out/r2_sanitizer_$tool.log | tr '\n' ' ')"
done
