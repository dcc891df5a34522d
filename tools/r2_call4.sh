#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r2_gpu_tests4.log 2>&1
tail -45 gpurun_out/r2_gpu_tests4.log
bash tools/ab.sh "pdl_light_on|" | tee gpurun_out/r2_ab4.txt
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --out gpurun_out/r2_tl_pdl > gpurun_out/r2_timeline_pdl.txt 2>&1
grep -A28 "== last frame" gpurun_out/r2_timeline_pdl.txt | head -45
cp comfyui-keep_b200/libkeep_b200.so /tmp/lib_on.so
KEEP_NVCC_EXTRA="-DKEEP_PDL_LIGHT_TRIGGER=0" python comfyui-keep_b200/build.py --force > gpurun_out/r2_rebuild.log 2>&1
bash tools/ab.sh "pdl_light_off|" | tee -a gpurun_out/r2_ab4.txt
cp /tmp/lib_on.so comfyui-keep_b200/libkeep_b200.so
bash tools/ab.sh "pdl_light_on_again|" "lockstep4_b4|AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 4" | tee -a gpurun_out/r2_ab4.txt
