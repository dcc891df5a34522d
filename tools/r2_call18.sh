#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -k "attention" 2>&1 | tail -4
echo "== attention trace, kv pack"; timeout 120 python tools/attn_trace.py 16
echo "== attention trace, convert per tile"; KEEP_ATTN_PACK=0 timeout 120 python tools/attn_trace.py 16
bash tools/ab.sh "nopack|KEEP_ATTN_PACK=0" "pack|" "nopack2|KEEP_ATTN_PACK=0" "pack2|"
