#!/bin/bash
# multi-GPU check of bench.py exactly as the driver launches it (torchrun, one rank per GPU): config 2 (one clip per GPU + gather)
# and config 5 (512-frame stream, clips round-robin).  usage: gpurun --gpus N -- 'bash tools/r2_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 1200 gpurun_out/r2_bench_n$N.json; tail -3 gpurun_out/r2_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --config 5 --batch-clips 4 --steps 2 --warmup 3 > gpurun_out/r2_bench_config5_n$N.json 2> gpurun_out/r2_bench_config5_n$N.err
cut -c1-400 gpurun_out/r2_bench_config5_n$N.json; tail -3 gpurun_out/r2_bench_config5_n$N.err
[ -n "$SKIP_REF" ] || python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/r2_bench_reference_n$N.json 2> gpurun_out/r2_bench_reference_n$N.err
[ -n "$SKIP_REF" ] || cut -c1-300 gpurun_out/r2_bench_reference_n$N.json
