#!/bin/bash
# Instruction / divergence / duration counters of every libvlidar kernel of one scan (run under gpurun).
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"k_" --csv --log-file gpurun_out/ncu_inst.csv python tools/profile_scan.py ${1:-710} 1 > gpurun_out/ncu_inst.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/ncu_inst.csv')) if len(r)>10]
h=rows[0]; k=h.index('Kernel Name'); m=h.index('Metric Name'); v=h.index('Metric Value'); i=h.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[i], r[k].split('(')[0][-28:]), {})[r[m]]=float(r[v].replace(',',''))
tot=0
for (id_,name),mm in d.items():
    tot+=mm.get('smsp__inst_executed.sum',0)
    print("%-30s %8.1f us  %10.0f inst  lanes %5.1f  warps_active %5.1f%%"%(name, mm.get('gpu__time_duration.sum',0)/1e3, mm.get('smsp__inst_executed.sum',0), mm.get('smsp__thread_inst_executed_per_inst_executed.ratio',0), mm.get('sm__warps_active.avg.pct_of_peak_sustained_active',0)))
print("total warp-instructions %.1f M"%(tot/1e6))
PY
