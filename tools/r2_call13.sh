#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -k "groupnorm or layernorm" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzz_batch.py tests/test_gpu_zzzz_T20.py -x -q 2>&1 | tail -5
bash tools/ab.sh "base|KEEP_GN_REDUCE_FINAL=0;KEEP_LN_REDUCE=0" "gnfin|KEEP_LN_REDUCE=0" "lnred|KEEP_GN_REDUCE_FINAL=0" "both|" "base2|KEEP_GN_REDUCE_FINAL=0;KEEP_LN_REDUCE=0" "both2|" | tee gpurun_out/r2_ab13.txt
