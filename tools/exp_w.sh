#!/bin/bash
# Dev experiment (under gpurun): bench lines of the given workloads.  usage: tools/exp_w.sh <tag> <workload>...
set -u
O=gpurun_out; TAG=$1; shift
for w in "$@"; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-others > $O/${TAG}_$w.json 2> $O/${TAG}_$w.err
  echo "$w rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$O/${TAG}_$w.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('ms/step', round(d['ms_per_step'],3), 'kernel_ms', r.get('kernel_ms_avg'), 'frac', r.get('frac'), 'value', '%.3e'%d['value'])
except Exception as e:
    print('no line', e)
PY
)"
done
