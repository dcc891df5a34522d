#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "small" 2>&1 | tail -2
CMD="python tools/run_clip.py --frames 2 --clips 1 --mode tc3"
timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum --clock-control none --kernel-name-base demangled -k "regex:conv_cin3_px4_kernel|conv_cout" -c 6 --csv --log-file gpurun_out/r2_stem_head.csv $CMD > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_stem_head.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d['Kernel Name'][:60], d['Grid Size'], d['Metric Name'], d['Metric Value'])
P
bash tools/ab.sh "small_v2|" "small_v2_b|"
