#!/bin/bash
mkdir -p gpurun_out
( KEEP_ATTNBLOCK_TC=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzzz_T20.py tests/test_gpu_zz_asian.py tests/test_gpu_zzz_batch.py -q -x ) 2>&1 | tail -6
bash tools/ab.sh "simt_attnblock|" "tc_attnblock|KEEP_ATTNBLOCK_TC=1" "simt_b|" "tc_b|KEEP_ATTNBLOCK_TC=1"
