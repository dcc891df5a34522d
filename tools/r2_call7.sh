#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh "default|" "skip_flow|KEEP_DEBUG_SKIP_FLOW=1" "flow_inline|KEEP_NO_SIDE=1" "side_sms_64|KEEP_SIDE_SMS=64" "side_sms_148|KEEP_SIDE_SMS=148" "side_short4|KEEP_SIDE_SMS=-4" | tee gpurun_out/r2_ab7.txt
cp comfyui-keep_b200/libkeep_b200.so /tmp/lib_default.so
KEEP_NVCC_EXTRA="-DKEEP_PDL_TINY_TRIGGER=1" python comfyui-keep_b200/build.py --force > gpurun_out/r2_rebuild7.log 2>&1
bash tools/ab.sh "tiny_on|" "tiny_on_skip_flow|KEEP_DEBUG_SKIP_FLOW=1" | tee -a gpurun_out/r2_ab7.txt
cp /tmp/lib_default.so comfyui-keep_b200/libkeep_b200.so
