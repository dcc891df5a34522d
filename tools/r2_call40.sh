#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "groupnorm" 2>&1 | tail -2
bash tools/ab.sh "tail4|" "sepfin|KEEP_GN_REDUCE_FINAL=0" "tail4_b|" "sepfin_b|KEEP_GN_REDUCE_FINAL=0"
