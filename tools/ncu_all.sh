#!/bin/bash
# ONE ncu pass over the kernel launches of a 2-frame clip (split-precision engine, eager launches; the ~650 one-off weight
# transposes / panel repacks of engine creation are skipped) with the few metrics the roofline discussion uses, then a
# per-kernel summary (tools/ncu_by_kernel.py -> gpurun_out/r2_ncu_per_kernel.{json,md}); plus `--set full --import-source on`
# captures of a few launches of the two heaviest kernels.  ~1 s of ncu overhead per launch: bounded by --launch-count.
#   gpurun --timeout 1800 -- 'bash tools/ncu_all.sh'
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size"
CMD="python tools/run_clip.py --frames 2 --clips 1 --mode tc3"
timeout 1000 ncu --metrics "$M" --clock-control none --kernel-name-base demangled --launch-skip 650 --launch-count 1400 --csv --log-file gpurun_out/r2_ncu_all.csv $CMD > gpurun_out/r2_ncu_all.log 2>&1
tail -2 gpurun_out/r2_ncu_all.log
python tools/ncu_by_kernel.py gpurun_out/r2_ncu_all.csv gpurun_out/r2_ncu_per_kernel > gpurun_out/r2_ncu_by_kernel.log 2>&1; tail -3 gpurun_out/r2_ncu_by_kernel.log
head -45 gpurun_out/r2_ncu_per_kernel.md
for spec in "conv_tc|conv_tc_kernel|60|6" "attn_tc|attn_tc_kernel|2|1"; do
  label="${spec%%|*}"; rest="${spec#*|}"; regex="${rest%%|*}"; rest="${rest#*|}"; skip="${rest%%|*}"; cnt="${rest#*|}"
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s "$skip" -c "$cnt" \
      -f -o "gpurun_out/r2_ncu_full_$label" $CMD > "gpurun_out/r2_ncu_full_$label.log" 2>&1
  if [ -f "gpurun_out/r2_ncu_full_$label.ncu-rep" ]; then
    ncu -i "gpurun_out/r2_ncu_full_$label.ncu-rep" --page raw --csv > "gpurun_out/r2_ncu_full_$label.csv" 2>/dev/null
    python tools/ncu_summary.py "gpurun_out/r2_ncu_full_$label.csv" "$CMD" "$label, --set full" > "gpurun_out/r2_ncu_full_$label.json" 2>/dev/null
    echo "full $label: ok ($(grep -c 'Kernel Name' gpurun_out/r2_ncu_full_$label.json) launches)"
  else echo "full $label: no report"; tail -3 "gpurun_out/r2_ncu_full_$label.log"; fi
done
# sanitizer passes over op-level tests of the two tcgen05 kernels (bounded; the tools slow kernels down 10-100x)
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -q -m gpu -x \
      -k "3x3_64_64 or linear_splitk or fused_attention_matches_torch and 128-64 or split_16sq_512 or epi_ragged" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool: rc=$? $(grep -E 'ERROR SUMMARY|passed|failed|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_$tool.log | tr '\n' ' ')"
done
