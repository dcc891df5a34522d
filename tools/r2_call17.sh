#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh "nopf|KEEP_TC_PREFETCH=0" "pf|" "nopf2|KEEP_TC_PREFETCH=0" "pf2|"
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -3
KEEP_NVCC_EXTRA="-DKEEP_TC_TRACE_FINE=1" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
echo "== 64->64 3x3 @512^2 tc3 + GN, prefetch"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish
