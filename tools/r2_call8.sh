#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "fused" ) > gpurun_out/r2_gpu_tests8a.log 2>&1
tail -8 gpurun_out/r2_gpu_tests8a.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzzz_T20.py tests/test_gpu_lifecycle.py -q -m gpu ) > gpurun_out/r2_gpu_tests8b.log 2>&1
tail -8 gpurun_out/r2_gpu_tests8b.log
bash tools/ab.sh "fused_attn_v2|" "unfused_attn|KEEP_FUSED_ATTN=0" "mha_simt|KEEP_FUSED_MHA=0" "skip_flow|KEEP_DEBUG_SKIP_FLOW=1" | tee gpurun_out/r2_ab8.txt
KEEP_NO_SIDE=1 timeout 300 python tools/timeline.py --frames 5 --out gpurun_out/r2_tl_8 > gpurun_out/r2_timeline_gmflow_inline_T5_v2.txt 2>&1
head -16 gpurun_out/r2_timeline_gmflow_inline_T5_v2.txt
