#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh "t704|" "t704_b|"
KEEP_NVCC_EXTRA="-DKEEP_TC_THREADS=640" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
echo "== 640 threads"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish | grep -E "prod:loads|mma:issued|epi|cta"
bash tools/ab.sh "t640|" "t640_b|"
