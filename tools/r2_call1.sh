#!/bin/bash
# Round-2 first GPU call: full GPU test tier, headline bench (with the eager-PyTorch-on-B200 and CPU baselines), the same-config
# reference arm (2 steps), configs 3 and 5 on one GPU, lockstep / wide-operand A/Bs, per-frame timeline.  All output -> gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/r2_call1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_box.txt; nproc >> gpurun_out/r2_box.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2_box.txt
( time timeout 1100 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2_gpu_tests.log 2>&1
tail -60 gpurun_out/r2_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 1500 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_T20.json 2> gpurun_out/r2_bench_reference.err
cat gpurun_out/r2_bench_reference_T20.json | cut -c1-400
timeout 400 python bench.py --config 3 --no-cpu-baseline --no-eager-baseline --steps 3 > gpurun_out/r2_bench_config3_n1.json 2> gpurun_out/r2_bench_config3.err
cut -c1-300 gpurun_out/r2_bench_config3_n1.json; tail -3 gpurun_out/r2_bench_config3.err
for bc in 1 2 4; do
  timeout 500 python bench.py --config 5 --batch-clips $bc --no-cpu-baseline --no-eager-baseline --steps 2 --warmup 3 > gpurun_out/r2_bench_config5_n1_lock$bc.json 2> gpurun_out/r2_bench_config5_lock$bc.err
  echo "config5 lock$bc: $(cut -c1-200 gpurun_out/r2_bench_config5_n1_lock$bc.json)"; tail -2 gpurun_out/r2_bench_config5_lock$bc.err
done
export AB_EXTRA="--no-eager-baseline"
bash tools/ab.sh \
  "one_clip|" \
  "replicas2_b4|AB_BENCH_ARGS=--clips-per-step 4" \
  "lockstep2_b4_rep2|AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 2" \
  "lockstep4_b4|AB_BENCH_ARGS=--clips-per-step 4 --batch-clips 4" \
  "wide_general|KEEP_FORCE_FLAGS=16" | tee gpurun_out/r2_ab.txt
KEEP_FORCE_FLAGS=16 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k tc3 > gpurun_out/r2_parity_wide_general.log 2>&1
tail -25 gpurun_out/r2_parity_wide_general.log
KEEP_DEBUG_SKIP_FLOW=1 timeout 300 python tools/timeline.py --frames 4 --out gpurun_out/r2_tl_single > gpurun_out/r2_timeline_single.txt 2>&1
tail -n +1 gpurun_out/r2_timeline_single.txt | head -75
