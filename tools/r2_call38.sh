#!/bin/bash
mkdir -p gpurun_out
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --out gpurun_out/r2_final_tl_noflow > gpurun_out/r2_final_timeline_noflow_T4.txt 2>&1
KEEP_NO_SIDE=1 timeout 300 python tools/timeline.py --frames 5 --out gpurun_out/r2_final_tl_inline > gpurun_out/r2_final_timeline_gmflow_inline_T5.txt 2>&1
grep -A6 "== last frame" gpurun_out/r2_final_timeline_noflow_T4.txt | head -8
timeout 900 python bench.py --steps 10 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
tail -c 600 gpurun_out/r2_final_bench_n1.json; tail -2 gpurun_out/r2_final_bench_n1.err
