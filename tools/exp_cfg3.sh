#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_reference_semantics.py tests/test_stream_replay.py tests/test_parity_golden.py -x -q -m gpu > $O/x_tests3.log 2>&1; echo "tests rc=$? $(tail -1 $O/x_tests3.log)"
for sh in default 3 4 6 9; do
  if [ $sh = default ]; then E=""; else E="VBN_SHAPE=$sh"; fi
  env $E timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-others > $O/x_c3_$sh.json 2> $O/x_c3_$sh.err
  echo "shape $sh rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$O/x_c3_$sh.json').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],2), 'kernel_ms', d['roofline']['kernel_ms_avg'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('no line', e)
PY
)"
done
