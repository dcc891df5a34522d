#!/bin/bash
# precision table: generator blocks >= j with plain fp16 operands (1 MMA pass) in the split-precision engine; error under the
# teacher-forced bar (T = 2 and T = 20, codes forced) and speed
mkdir -p gpurun_out
for j in 99 20 17 14; do
  echo "== KEEP_GEN_FAST_FROM=$j"
  KEEP_GEN_FAST_FROM=$j timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzzz_T20.py -q -k "stagewise_teacher_forced_T2 or codes_forced_every_frame" 2>&1 | grep -E "forced_T2\[tc3\]|T20_codes_forced|passed|failed"
done
bash tools/ab.sh "pass3|" "fast20|KEEP_GEN_FAST_FROM=20" "fast17|KEEP_GEN_FAST_FROM=17" "fast14|KEEP_GEN_FAST_FROM=14"
