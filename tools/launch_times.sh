#!/bin/bash
# usage: tools/launch_times.sh <out.csv> [lib.so] [streams] [waves]  -- per-kernel average duration (us) under ncu, cold-cache serialised launches
out=$1; lib=${2:-}; n=${3:-32768}; t=${4:-3}
if [ -n "$lib" ]; then export SOLO_B200_LIB=$lib; fi
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out python tools/ncu_target.py $n $t dec > /dev/null 2>&1
python - "$out" <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki].split("(")[0].replace("<unnamed>::","")
    agg.setdefault(k,[]).append(float(r[vi])/1000.0)
for k,v in agg.items():
    if k.startswith("sb_") and "init" not in k: print("%-32s n=%d avg %.1f us (last %.1f)"%(k,len(v),sum(v)/len(v),v[-1]))
PY
