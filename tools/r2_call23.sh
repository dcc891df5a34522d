#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
echo "== LD2 at 96 regs"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish | grep -E "mma:issued|epi|cta"
bash tools/ab.sh "ld2|" "ld2_b|"
KEEP_NVCC_EXTRA="-DKEEP_TC_EPI_LD2=0" python comfyui-keep_b200/build.py --force > /dev/null 2>&1
bash tools/ab.sh "ld1|" "ld1_b|"
