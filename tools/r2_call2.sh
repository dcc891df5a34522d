#!/bin/bash
# Round-2 second GPU call: GPU test tier (no -x), then GroupNorm-statistics fusion A/B + timeline, layer times.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r2_gpu_tests2.log 2>&1
tail -70 gpurun_out/r2_gpu_tests2.log
bash tools/ab.sh "gn_fused|" "gn_standalone|KEEP_GN_EPILOGUE=0" "cluster8|KEEP_TC_CLUSTER=8" "cluster4|KEEP_TC_CLUSTER=4" | tee gpurun_out/r2_ab2.txt
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --out gpurun_out/r2_tl_gnfused > gpurun_out/r2_timeline_gnfused.txt 2>&1
grep -A32 "== last frame" gpurun_out/r2_timeline_gnfused.txt | head -60
timeout 300 python tools/layer_times.py --mode tc3 --frames 5 > gpurun_out/r2_layer_times.txt 2>&1; head -80 gpurun_out/r2_layer_times.txt
