#!/bin/bash
# usage: tools/prof_summary.sh <report.ncu-rep> <object-with-cubin.o> <kernel-substring> <groups.py>
set -e
REP=$1; OBJ=$2; KN=$3; GROUPS_FILE=$4
T=$(mktemp -d)
( cd $T && cuobjdump -xelf all $OBJ >/dev/null )
ncu -i $REP --page source --csv > $T/src.csv 2>/dev/null
ncu -i $REP --page raw --csv > $T/raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('$T/raw.csv')))
hdr=rows[0]; vals=rows[2]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__sass_inst_executed_op_local_ld.sum','smsp__sass_inst_executed_op_local_st.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum']
for h,v in zip(hdr,vals):
    if h in want: print(h,v)
    elif h.startswith('smsp__average_warps_issue_stalled') and h.endswith('ratio') and float(v)>0.15: print('  ',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),round(float(v),2))
PY
python /root/repo/tools/linemap_groups.py $T/src.csv $T/*.cubin $KN $GROUPS_FILE
echo $T
