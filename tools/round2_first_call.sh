#!/bin/bash
# First gpurun call of a round: the whole GPU test tier (timed), the headline bench line, and the A/Bs left unmeasured
# at the end of round 1 (lockstep clip batching, bf16 "wide" operands on the general config).  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2_gpu_tests.log 2>&1
tail -3 gpurun_out/r2_gpu_tests.log
timeout 400 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 600 gpurun_out/r2_bench_n1.json
bash tools/ab.sh \
  "one_clip|" \
  "replicas2_b4|AB_BENCH_ARGS=--clips-per-step 4" \
  "lockstep2_b4_seq|KEEP_BENCH_REPLICAS=1;AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 2" \
  "lockstep2_b4_rep2|AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 2" \
  "lockstep4_b4|AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 4" \
  "lockstep2_b2|AB_BENCH_ARGS=--clips-per-step 2 --batch-clips 2" \
  "wide_general|KEEP_FORCE_FLAGS=16" \
  "gm_fuse_qkv|KEEP_GM_FUSE_QKV=1" | tee gpurun_out/r2_ab.txt
# parity of the wide flag on the general config (teacher-forced + free-running, tc3)
KEEP_FORCE_FLAGS=16 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k tc3 > gpurun_out/r2_parity_wide_general.log 2>&1
tail -3 gpurun_out/r2_parity_wide_general.log
# GMFlow fused q|k|v projections: flows / free-running parity with the knob on
KEEP_GM_FUSE_QKV=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "free_running_T3" > gpurun_out/r2_parity_gm_fuse.log 2>&1
tail -3 gpurun_out/r2_parity_gm_fuse.log
# where a lockstep frame's time goes (single stream: no GMFlow), next to the per-clip timeline
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --out gpurun_out/r2_tl_single > gpurun_out/r2_timeline_single.txt 2>&1
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --clips 2 --batch-clips 2 --out gpurun_out/r2_tl_lock2 > gpurun_out/r2_timeline_lockstep2.txt 2>&1
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --clips 4 --batch-clips 4 --out gpurun_out/r2_tl_lock4 > gpurun_out/r2_timeline_lockstep4.txt 2>&1
tail -n +1 gpurun_out/r2_timeline_lockstep2.txt | head -40
