#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "fused" ) > gpurun_out/r2_gpu_tests9a.log 2>&1
tail -5 gpurun_out/r2_gpu_tests9a.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzzz_T20.py -q -m gpu ) > gpurun_out/r2_gpu_tests9b.log 2>&1
tail -5 gpurun_out/r2_gpu_tests9b.log
bash tools/ab.sh "fused_attn_v3|" "unfused_attn|KEEP_FUSED_ATTN=0" | tee gpurun_out/r2_ab9.txt
KEEP_NO_SIDE=1 timeout 300 python tools/timeline.py --frames 5 --out gpurun_out/r2_tl_9 > gpurun_out/r2_timeline_gmflow_inline_T5_v3.txt 2>&1
head -12 gpurun_out/r2_timeline_gmflow_inline_T5_v3.txt
