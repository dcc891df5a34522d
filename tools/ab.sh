#!/bin/bash
# A/B harness for the GPU box: each line = label + env assignments; prints frames/s of a short bench per setting.
# usage: tools/ab.sh "label1|ENV=..;ENV2=.." "label2|..."     (AB_BENCH_ARGS=<extra bench.py arguments> in a spec applies to that line)
mkdir -p gpurun_out
for spec in "$@"; do
  label="${spec%%|*}"; envs="${spec#*|}"
  ( IFS=';'; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done; unset IFS
    timeout 240 python bench.py --no-cpu-baseline --no-eager-baseline --steps ${AB_STEPS:-4} --warmup 3 $AB_BENCH_ARGS > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
    python - "$label" <<'P'
import json,sys
lab=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/ab_%s.json'%lab).read().strip().splitlines()[-1])
    print("AB %-24s value %.2f e2e %.2f ms/step %.2f conv-family TF/s %.1f"%(lab,d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['achieved']))
except Exception as e:
    print("AB %-24s FAILED %r"%(lab,e)); print(open('gpurun_out/ab_%s.err'%lab).read()[-1500:])
P
  )
done
