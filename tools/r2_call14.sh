#!/bin/bash
mkdir -p gpurun_out
echo "== 64->64 3x3 @512^2 tc3, plain epilogue"; timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish
echo "== same, GroupNorm statistics in the epilogue"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 64 512 512 64 3 3 swish
echo "== 128->128 3x3 @256^2 tc3 plain"; timeout 120 python tools/tc_trace.py 1 128 256 256 128 3 3 swish
echo "== same + GN"; TC_TRACE_GN=1 timeout 120 python tools/tc_trace.py 1 128 256 256 128 3 3 swish
bash tools/ab.sh "noflow_base|KEEP_DEBUG_SKIP_FLOW=1;KEEP_GN_REDUCE_FINAL=0;KEEP_LN_REDUCE=0" "noflow_both|KEEP_DEBUG_SKIP_FLOW=1" "lock4_base|KEEP_GN_REDUCE_FINAL=0;KEEP_LN_REDUCE=0;AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 4" "lock4_both|AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 4"
