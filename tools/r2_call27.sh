#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
echo "== split-K 16^2 512->512, partial32"; timeout 120 python tools/tc_trace.py 1 512 16 16 512 3 3 swish | grep -E "mma:issued|epi|cta"
bash tools/ab.sh "p32|" "p32_b|"
KEEP_NVCC_EXTRA="-DKEEP_TC_PARTIAL32=0" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
echo "== split-K 16^2 512->512, 16 columns per round trip"; timeout 120 python tools/tc_trace.py 1 512 16 16 512 3 3 swish | grep -E "mma:issued|epi|cta"
bash tools/ab.sh "p16|" "p16_b|"
