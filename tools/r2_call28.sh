#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh "base|" "lq20|KEEP_LQ_CHUNK=20" "lq5|KEEP_LQ_CHUNK=5" "gn3|KEEP_GN_EPILOGUE=3" "gn2|KEEP_GN_EPILOGUE=2" "unstacked|KEEP_TC_RESIDENT=1" "base2|"
