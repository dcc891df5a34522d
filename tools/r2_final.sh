#!/bin/bash
# End-of-round evidence run (one GPU): full GPU test tier, headline bench line with its baselines, the reference arm on the same
# clip, configs 3 and 5, timelines, and the ncu capture of the dominant conv shape that bench.py's roofline.traffic quotes.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r2_final_gpu_tests.log 2>&1
tail -70 gpurun_out/r2_final_gpu_tests.log | cut -c1-400
timeout 900 python bench.py --steps 10 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
tail -c 2500 gpurun_out/r2_final_bench_n1.json; tail -3 gpurun_out/r2_final_bench_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
cut -c1-300 gpurun_out/r2_final_bench_reference.json
timeout 400 python bench.py --config 3 --no-cpu-baseline --no-eager-baseline --steps 3 > gpurun_out/r2_final_bench_config3.json 2> gpurun_out/r2_final_bench_config3.err
cut -c1-250 gpurun_out/r2_final_bench_config3.json
timeout 500 python bench.py --config 5 --batch-clips 4 --no-cpu-baseline --no-eager-baseline --steps 2 > gpurun_out/r2_final_bench_config5_lock4.json 2> gpurun_out/r2_final_bench_config5.err
cut -c1-250 gpurun_out/r2_final_bench_config5_lock4.json
timeout 300 python bench.py --clips-per-step 4 --batch-clips 4 --no-cpu-baseline --no-eager-baseline --steps 4 > gpurun_out/r2_final_bench_lockstep4.json 2> gpurun_out/r2_final_bench_lockstep4.err
cut -c1-250 gpurun_out/r2_final_bench_lockstep4.json
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --out gpurun_out/r2_final_tl_noflow > gpurun_out/r2_final_timeline_noflow_T4.txt 2>&1
KEEP_NO_SIDE=1 timeout 300 python tools/timeline.py --frames 5 --out gpurun_out/r2_final_tl_inline > gpurun_out/r2_final_timeline_gmflow_inline_T5.txt 2>&1
grep -A12 "== last frame" gpurun_out/r2_final_timeline_noflow_T4.txt | head -16
CMD="python tools/run_clip.py --frames 2 --clips 1 --mode tc3"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel<\(int\)3, \(bool\)0, \(int\)3, \(bool\)0>" -s 11 -c 16 \
    -f -o gpurun_out/r2_ncu_full_conv3x3 $CMD > gpurun_out/r2_ncu_full_conv3x3.log 2>&1
if [ -f gpurun_out/r2_ncu_full_conv3x3.ncu-rep ]; then
  ncu -i gpurun_out/r2_ncu_full_conv3x3.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_conv3x3.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/r2_ncu_full_conv3x3.csv "$CMD" "conv_tc_kernel<3,false,3,false>, 16 launches after the first 11, --set full" > gpurun_out/r2_ncu_full_conv3x3.json
  python - <<'P'
import json
d=json.load(open('gpurun_out/r2_ncu_full_conv3x3.json'))
for l in d['launches']: print(l['Grid Size'], l['gpu__time_duration.sum'], 'rd', l['dram__bytes_read.sum'], 'wr', l['dram__bytes_write.sum'], 'tensor', l['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'][:5])
P
else tail -5 gpurun_out/r2_ncu_full_conv3x3.log; fi
