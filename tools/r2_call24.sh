#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh "side64|" "side48|KEEP_SIDE_SMS=48" "side80|KEEP_SIDE_SMS=80" "side100|KEEP_SIDE_SMS=100" "chunk3|KEEP_FLOW_CHUNK=3" "chunk4|KEEP_FLOW_CHUNK=4" "chunk1|KEEP_FLOW_CHUNK=1" "lock4|AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 4"
