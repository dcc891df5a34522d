#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
echo "== 640 thr, 2 groups"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish | grep -E "prod:A_FULL|mma:issued|epi:done|cta"
bash tools/ab.sh "g2_640|" "g2_640_b|"
KEEP_NVCC_EXTRA="-DKEEP_TC_THREADS=576 -DKEEP_TC_THREADS_1X1=640" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
echo "== 576 thr, 2 groups"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish | grep -E "prod:A_FULL|mma:issued|epi:done|cta"
bash tools/ab.sh "g2_576|" "g2_576_b|"
KEEP_NVCC_EXTRA="-DKEEP_TC_GROUPS_3X3=1" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
bash tools/ab.sh "g1_640|" "g1_640_b|"
