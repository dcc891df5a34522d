#!/bin/bash
# A/Bs: tiny-kernel early trigger, conv late trigger, tensor-core CFA-32 attention; parity subset for the new attention path
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzzz_T20.py -q -m gpu ) > gpurun_out/r2_gpu_tests5.log 2>&1
tail -12 gpurun_out/r2_gpu_tests5.log
bash tools/ab.sh "tiny_on|" "tiny_on_mha_simt|KEEP_MHA_TC_MIN_L=100000" | tee gpurun_out/r2_ab5.txt
cp comfyui-keep_b200/libkeep_b200.so /tmp/lib_default.so
KEEP_NVCC_EXTRA="-DKEEP_PDL_TINY_TRIGGER=0" python comfyui-keep_b200/build.py --force > gpurun_out/r2_rebuild5a.log 2>&1
bash tools/ab.sh "tiny_off|" | tee -a gpurun_out/r2_ab5.txt
KEEP_NVCC_EXTRA="-DKEEP_PDL_CONV_TRIGGER=1" python comfyui-keep_b200/build.py --force > gpurun_out/r2_rebuild5b.log 2>&1
bash tools/ab.sh "tiny_on_convtrig|" | tee -a gpurun_out/r2_ab5.txt
KEEP_NVCC_EXTRA="-DKEEP_PDL_CONV_TRIGGER=1 -DKEEP_PDL_TINY_TRIGGER=0" python comfyui-keep_b200/build.py --force > gpurun_out/r2_rebuild5c.log 2>&1
bash tools/ab.sh "tiny_off_convtrig|" | tee -a gpurun_out/r2_ab5.txt
cp /tmp/lib_default.so comfyui-keep_b200/libkeep_b200.so
bash tools/ab.sh "tiny_on_again|" | tee -a gpurun_out/r2_ab5.txt
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --out gpurun_out/r2_tl_5 > gpurun_out/r2_timeline_5.txt 2>&1
grep -A28 "== last frame" gpurun_out/r2_timeline_5.txt | head -45
