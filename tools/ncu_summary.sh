#!/bin/bash
# tools/ncu_summary.sh <report.ncu-rep> <profiles/prefix>: the three text summaries committed under profiles/
# (details page, selected raw metrics, per-source-line instructions + stall samples).
set -e
rep=$1; out=$2
ncu -i $rep --page details > ${out}_details.txt
ncu -i $rep --page raw --csv | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr,units,vals=rows[0],rows[1],rows[2]
import re
pat=re.compile(r'dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum|sm__cycles_active\.avg$|sm__inst_executed_pipe_(adu|alu|cbu|fma|fp64|lsu|uniform|xu)\.avg\.pct|sm__pipe_(alu|fma)_cycles_active\.avg\.pct|smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct|smsp__thread_inst_executed_per_inst_executed\.ratio|sm__warps_active\.avg\.pct|launch__registers_per_thread|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate\.pct|smsp__cycles_active\.avg$|sm__throughput\.avg\.pct|launch__grid_size|launch__occupancy_limit')
for h,u,v in zip(hdr,units,vals):
    if pat.search(h): print(h,u,v)
" > ${out}_raw_selected.txt
ncu -i $rep --page source --csv --print-source cuda,sass > /tmp/ncu_both.csv 2>/dev/null
python3 $(dirname $0)/ncu_lines.py /tmp/ncu_both.csv 45 > ${out}_lines.txt
