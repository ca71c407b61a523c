#!/bin/bash
# Dev experiment (under gpurun): cfg5 bench lines of library variants built into .variants/lib_<name>.so
# ("base" = the in-tree library).  usage: tools/run_var.sh <name>[:ENV=VAL] ...
set -u
O=gpurun_out
cp vectorizedbayesiannetwork_b200/libvbn_cuda.so /tmp/base.so
for spec in "$@"; do
  v=${spec%%:*}; envs=""; [ "$spec" != "$v" ] && envs=${spec#*:}
  if [ "$v" = base ]; then cp /tmp/base.so vectorizedbayesiannetwork_b200/libvbn_cuda.so; else cp .variants/lib_$v.so vectorizedbayesiannetwork_b200/libvbn_cuda.so; fi
  env $envs timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-others --queries-per-gpu 1250 > $O/v_$v.json 2> $O/v_$v.err
  echo "$v rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$O/v_$v.json').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],2), 'kernel_ms', d['roofline']['kernel_ms_avg'])
except Exception as e:
    print('no line', e)
PY
)"
done
cp /tmp/base.so vectorizedbayesiannetwork_b200/libvbn_cuda.so
