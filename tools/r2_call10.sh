#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -q -m gpu ) > gpurun_out/r2_gpu_tests10a.log 2>&1
tail -12 gpurun_out/r2_gpu_tests10a.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzzz_T20.py tests/test_gpu_zz_asian.py -q -m gpu ) > gpurun_out/r2_gpu_tests10b.log 2>&1
tail -6 gpurun_out/r2_gpu_tests10b.log
bash tools/ab.sh "stacked64_on|" | tee gpurun_out/r2_ab10.txt
cp comfyui-keep_b200/libkeep_b200.so /tmp/lib_default.so
KEEP_NVCC_EXTRA="-DKEEP_TC_STACKED=0" python comfyui-keep_b200/build.py --force > gpurun_out/r2_rebuild10.log 2>&1
bash tools/ab.sh "stacked64_off|" | tee -a gpurun_out/r2_ab10.txt
cp /tmp/lib_default.so comfyui-keep_b200/libkeep_b200.so
bash tools/ab.sh "stacked64_on_again|" "gm_fuse_qkv|KEEP_GM_FUSE_QKV=1" | tee -a gpurun_out/r2_ab10.txt
KEEP_GM_FUSE_QKV=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "free_running_T3 or free_T2" > gpurun_out/r2_parity_gm_fuse.log 2>&1; tail -3 gpurun_out/r2_parity_gm_fuse.log
timeout 200 python tools/layer_times.py --mode tc3 --frames 3 > gpurun_out/r2_layer_times_stacked.txt 2>&1; head -12 gpurun_out/r2_layer_times_stacked.txt
bash tools/ab.sh "split_target_112|KEEP_TC_SPLIT_TARGET=112" "split_target_148|KEEP_TC_SPLIT_TARGET=148" "split_target_64|KEEP_TC_SPLIT_TARGET=64" "side_48|KEEP_SIDE_SMS=48" "side_80|KEEP_SIDE_SMS=80" "flow_chunk_8|KEEP_FLOW_CHUNK=8" "flow_chunk_2|KEEP_FLOW_CHUNK=2" "lq_chunk_20|KEEP_LQ_CHUNK=20" "lq_chunk_5|KEEP_LQ_CHUNK=5" | tee -a gpurun_out/r2_ab10.txt
