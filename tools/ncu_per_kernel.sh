#!/bin/bash
# One `ncu --set full` capture per kernel family of the hot path (north_star: "each kernel ships with an ncu capture reporting
# achieved HBM GB/s and tensor-pipe utilisation").  Run on the GPU box; ~40 replays per captured launch, so each family is
# limited to 2 launches of a warmed T=2 clip.  Writes gpurun_out/ncu_<family>.ncu-rep and a JSON summary per family
# (tools/ncu_summary.py) -- copy the JSONs to profiles/.
#   gpurun --timeout 1500 -- 'bash tools/ncu_per_kernel.sh'
mkdir -p gpurun_out
CMD="python tools/run_clip.py --frames 2 --clips 2 --mode tc3"
# family label | kernel-name regex | launches to skip (land in the second, warmed clip)
FAMILIES=(
  "conv_tc_3x3|conv_tc_kernel<3, (0|false), 3|400"
  "conv_tc_1x1|conv_tc_kernel<3, (0|false), 1|600"
  "conv_tc_s2|conv_tc_kernel<3, (0|false), 2|20"
  "splitk_reduce|splitk_reduce_kernel|700"
  "gn_partial|gn_partial_kernel|180"
  "gn_small|gn_small_kernel|160"
  "layernorm|layernorm_v4_kernel|80"
  "bgemm32|bgemm32_kernel|70"
  "softmax_rows|softmax_rows_v4_kernel|40"
  "softmax_expect2|softmax_expect2_kernel|2"
  "conv_stem|conv_cin3_px4_kernel|4"
  "conv_head|conv_cout4_kernel|2"
  "flow_warp|flow_warp_kernel|1"
  "argmax_gather|argmax_gather_kernel|2"
  "window_partition|window_partition_kernel|96"
)
for spec in "${FAMILIES[@]}"; do
  label="${spec%%|*}"; rest="${spec#*|}"; skip="${rest##*|}"; regex="${rest%|*}"     # the regex itself may contain '|' 
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s "$skip" -c 2 \
      -f -o "gpurun_out/ncu_$label" $CMD > "gpurun_out/ncu_$label.log" 2>&1
  if [ -f "gpurun_out/ncu_$label.ncu-rep" ]; then
    ncu -i "gpurun_out/ncu_$label.ncu-rep" --page raw --csv > "gpurun_out/ncu_$label.csv" 2>/dev/null
    python tools/ncu_summary.py "gpurun_out/ncu_$label.csv" > "gpurun_out/ncu_$label.json" 2>/dev/null
    echo "ncu $label: $(python - "gpurun_out/ncu_$label.json" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1])); l = d["launches"][0] if isinstance(d, dict) else d[0]
    print(l.get("gpu__time_duration.sum"), "| dram rd", l.get("dram__bytes_read.sum"), "wr", l.get("dram__bytes_write.sum"),
          "| tensor pipe", l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"))
except Exception as e:
    print("summary failed:", e)
P
)"
  else
    echo "ncu $label: no report (see gpurun_out/ncu_$label.log)"
  fi
done
