#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_lifecycle.py -x -q -k default_engine 2>&1 | grep -E "lifecycle|passed|failed"; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
