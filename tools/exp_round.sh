#!/bin/bash
# Dev experiment (run under gpurun, one GPU): parity tests of the tcgen05 path, then cfg5 bench lines.
set -u
O=gpurun_out
TAG=${TAG:-e}
if [ "${TESTS:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_tc_path.py tests/test_parity_golden.py tests/test_stream_replay.py -x -q -m gpu > $O/${TAG}_tests.log 2>&1
  echo "tests rc=$? $(tail -1 $O/${TAG}_tests.log)"
fi
for q in ${QUERIES:-1250}; do
  timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-others --queries-per-gpu $q > $O/${TAG}_q$q.json 2> $O/${TAG}_q$q.err
  echo "q=$q rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$O/${TAG}_q$q.json').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],2), 'kernel_ms', d['roofline']['kernel_ms_avg'], 'value', '%.3e'%d['value'])
except Exception as e:
    print('no line', e)
PY
)"
done
for t in ${TUNES:-}; do
  VBN_TC_TUNE=$t timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-others --queries-per-gpu 1250 > $O/${TAG}_t$t.json 2> $O/${TAG}_t$t.err
  echo "tune=$t rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$O/${TAG}_t$t.json').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],2), 'kernel_ms', d['roofline']['kernel_ms_avg'], 'value', '%.3e'%d['value'])
except Exception as e:
    print('no line', e)
PY
)"
done
for sh in ${SHAPES:-}; do
  VBN_TC_SHAPE=$sh timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-others --queries-per-gpu 1250 > $O/${TAG}_s$sh.json 2> $O/${TAG}_s$sh.err
  echo "shape=$sh rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$O/${TAG}_s$sh.json').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],2), 'kernel_ms', d['roofline']['kernel_ms_avg'], 'value', '%.3e'%d['value'])
except Exception as e:
    print('no line', e)
PY
)"
done
