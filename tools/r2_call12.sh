#!/bin/bash
mkdir -p gpurun_out
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --clips 4 --batch-clips 4 --out gpurun_out/r2_tl_lock4 > gpurun_out/r2_timeline_lockstep4.txt 2>&1
grep -A30 "== last frame" gpurun_out/r2_timeline_lockstep4.txt | head -40
bash tools/ab.sh "one_clip|" "lock4_b4|AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 4" "lock8_b8|AB_BENCH_ARGS=--clips-per-step 8 --batch-clips 8" "lock4_b8_rep2|AB_BENCH_ARGS=--clips-per-step 8 --batch-clips 4" "lock4_skipflow|KEEP_DEBUG_SKIP_FLOW=1;AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 4" | tee gpurun_out/r2_ab12.txt
