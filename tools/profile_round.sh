#!/bin/bash
# Round profile capture (run under gpurun, one GPU): per-launch device times of one bench command per
# workload, then one `--set full` capture of each dominant kernel.  Outputs land in gpurun_out/.
set -u
R=${1:-r02}
declare -A ARGS=( [cfg5]="--queries-per-gpu 296" [cfg2]="--samples 250000" [cfg3]="--queries-per-gpu 1024" [cfg4]="--queries-per-gpu 303104" )
# mangled-name regexes: cfg3 also launches small schedule_kernel<2,128,true,..> instances while it builds its lookup
# tables, which a plain "schedule_kernel" filter would capture instead of the dominant <4,256,false,..> kernel
declare -A KREG=( [cfg5]="schedule_tc_kernel" [cfg2]="schedule_kernelILi4ELi256ELb0" [cfg3]="schedule_kernelILi4ELi256ELb0" [cfg4]="kde_log_prob" )
for w in cfg5 cfg2 cfg3 cfg4; do
  CMD="python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-others ${ARGS[$w]}"
  $CMD > gpurun_out/${R}_plain_$w.json 2> gpurun_out/${R}_plain_$w.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_$w.csv $CMD > gpurun_out/${R}_ncu_launches_$w.log 2>&1
  $CMD > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:${KREG[$w]} -s 3 -c 1 -f -o gpurun_out/${R}_prof_$w $CMD > gpurun_out/${R}_ncu_full_$w.log 2>&1
  tail -1 gpurun_out/${R}_ncu_full_$w.log
  # gpurun brings back at most 64 MiB: keep the raw-metric export, drop the report (cfg5's is kept)
  ncu -i gpurun_out/${R}_prof_$w.ncu-rep --page raw --csv > gpurun_out/${R}_prof_$w.raw.csv 2>/dev/null
  if [ "$w" != "cfg5" ]; then rm -f gpurun_out/${R}_prof_$w.ncu-rep; fi
done
# the tensor-core KDE kernel (Dp = 7, Dx = 1, 50 000 stored points): timing, then one full capture
python tools/bench_kde_tc.py 50000 262144 7 1 > gpurun_out/${R}_kde_tc.txt 2>&1
python tools/bench_kde_tc.py 200000 262144 14 2 >> gpurun_out/${R}_kde_tc.txt 2>&1
ncu --set full --clock-control none --kernel-name-base mangled -k regex:kde_tc_kernel -s 1 -c 1 -f -o gpurun_out/${R}_prof_kdetc python tools/bench_kde_tc.py 50000 65536 7 1 > gpurun_out/${R}_ncu_full_kdetc.log 2>&1
ncu -i gpurun_out/${R}_prof_kdetc.ncu-rep --page raw --csv > gpurun_out/${R}_prof_kdetc.raw.csv 2>/dev/null
rm -f gpurun_out/${R}_prof_kdetc.ncu-rep
python tools/latency_cfg1.py > gpurun_out/${R}_cfg1_latency.txt 2>&1
tail -1 gpurun_out/${R}_cfg1_latency.txt
