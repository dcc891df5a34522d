#!/bin/bash
mkdir -p gpurun_out
echo "== LD2=1"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish | grep -E "epi|mma:issued|cta"
bash tools/ab.sh "ld2|" "ld2_b|"
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
KEEP_NVCC_EXTRA="-DKEEP_TC_EPI_LD2=0" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
echo "== LD2=0"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish | grep -E "epi|mma:issued|cta"
bash tools/ab.sh "ld1|" "ld1_b|"
