#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -m gpu ) > gpurun_out/r2_gpu_tests3.log 2>&1
tail -40 gpurun_out/r2_gpu_tests3.log
bash tools/ab.sh "gn0|KEEP_GN_EPILOGUE=0" "gn1|KEEP_GN_EPILOGUE=1" "gn2|KEEP_GN_EPILOGUE=2" "gn3|KEEP_GN_EPILOGUE=3" | tee gpurun_out/r2_ab3.txt
# role timelines of CTA 0: n cin h w cout mode(3 = split precision) ksz act
for spec in "8 128 64 64 128 3 1 none" "8 256 64 64 1024 3 1 none" "8 1024 64 64 128 3 1 none" "1 512 16 16 512 3 3 swish" "1 512 16 16 512 3 1 none" "1 256 64 64 256 3 3 swish" "1 64 512 512 64 3 3 swish" "1 128 256 256 128 3 3 swish"; do
  echo "== trace $spec"; timeout 120 python tools/tc_trace.py $spec 2>&1 | tail -12
done > gpurun_out/r2_traces.txt 2>&1
cat gpurun_out/r2_traces.txt
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --out gpurun_out/r2_tl_gn3 > gpurun_out/r2_timeline_gn3.txt 2>&1
grep -A28 "== last frame" gpurun_out/r2_timeline_gn3.txt | head -50
