#!/bin/bash
# usage (under gpurun): tools/prof_one.sh <tag> <workload> <kernel-regex> [bench args...]: one --set full capture
set -u
TAG=$1; W=$2; K=$3; shift 3
CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --no-others $*"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log
