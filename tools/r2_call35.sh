#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
tail -c 1800 gpurun_out/r2_final_bench_n1.json; tail -3 gpurun_out/r2_final_bench_n1.err
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
