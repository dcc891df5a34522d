#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh "restored|"
KEEP_NVCC_EXTRA="-DKEEP_TC_TRACE_FINE=1" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
echo "== 64->64 3x3 @512^2 tc3 + GN, fine trace"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish
echo "== 64->64 plain, fine trace"; timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish
