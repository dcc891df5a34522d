#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh "all640|" "all640_b|"
KEEP_NVCC_EXTRA="-DKEEP_TC_THREADS_1X1=704" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
bash tools/ab.sh "lin704|" "lin704_b|"
KEEP_NVCC_EXTRA="-DKEEP_TC_THREADS_1X1=576" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
bash tools/ab.sh "lin576|" "lin576_b|"
KEEP_NVCC_EXTRA="-DKEEP_TC_THREADS=576 -DKEEP_TC_THREADS_1X1=640" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
bash tools/ab.sh "conv576|" "conv576_b|"
