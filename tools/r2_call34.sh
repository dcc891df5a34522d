#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -q -m gpu -x ) 2>&1 | tail -8
bash tools/ab.sh "mha256_default|" "mha1024|KEEP_MHA_TC_MIN_L=1024"
