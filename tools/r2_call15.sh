#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -4
echo "== 64->64 3x3 @512^2 tc3 + GN, pair"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish
echo "== same, no pair"; KEEP_TC_PAIR=0 TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish
bash tools/ab.sh "nopair|KEEP_TC_PAIR=0" "pair|" "nopair2|KEEP_TC_PAIR=0" "pair2|"
