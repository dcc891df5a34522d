#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh "base|" "tgt56|KEEP_TC_SPLIT_TARGET=56" "tgt112|KEEP_TC_SPLIT_TARGET=112" "tgt148|KEEP_TC_SPLIT_TARGET=148" "min64|KEEP_TC_SPLIT_MIN=64" "mha256|KEEP_MHA_TC_MIN_L=256" "base2|"
