#!/bin/bash
# Dev experiment (run under gpurun, one GPU): cfg5 bench lines of tcgen05 kernel variants + ncu captures.
set -u
O=gpurun_out
run_bench() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > $O/x_$name.json 2> $O/x_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$O/x_$name.json').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],2), 'kernel_ms', d['roofline']['kernel_ms_avg'], 'value', '%.3e'%d['value'])
except Exception as e:
    print('no line', e)
PY
)"
}
prof() { # name, env...
  local name=$1; shift
  CMD="python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --queries-per-gpu 296"
  env "$@" timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:schedule_tc_kernel -s 3 -c 1 -f -o $O/x_prof_$name $CMD > $O/x_ncu_$name.log 2>&1
  tail -1 $O/x_ncu_$name.log
}
if [ "${TESTS:-1}" = "1" ]; then
  timeout 600 python -m pytest tests/test_tc_path.py tests/test_parity_golden.py -x -q -m gpu > $O/x_tests.log 2>&1
  echo "tests rc=$? $(tail -1 $O/x_tests.log)"
fi
for v in ${VARIANTS:-"4x1:0 2x2:0"}; do
  shape=${v%%:*}; dbg=${v##*:}
  run_bench ${shape}_d${dbg} VBN_TC_SHAPE=$shape VBN_TC_TUNE=$dbg
done
for v in ${PROFS:-"4x1"}; do prof $v VBN_TC_SHAPE=$v; done
