#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lifecycle.py tests/test_gpu_parity.py tests/test_gpu_zzz_batch.py -x -q 2>&1 | tail -4
bash tools/ab.sh "serial|KEEP_LQ_OVERLAP=0" "overlap|" "serial2|KEEP_LQ_OVERLAP=0" "overlap2|" "overlap_noflow|KEEP_DEBUG_SKIP_FLOW=1" "serial_noflow|KEEP_DEBUG_SKIP_FLOW=1;KEEP_LQ_OVERLAP=0"
