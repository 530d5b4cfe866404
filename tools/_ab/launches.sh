#!/bin/bash
# usage: launches.sh <name> <skip> <count> <command...>: per-launch device times (ncu, cold cache: compare shares)
name=$1; skip=$2; count=$3; shift 3
"$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$name.log; exit 1; }
tail -2 gpurun_out/plain_$name.log
ncu --metrics gpu__time_duration.sum --clock-control none -s $skip -c $count --csv --log-file gpurun_out/launches_$name.csv "$@" > gpurun_out/ncu_$name.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_$name.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
seq=[(r[ik],float(r[iv].replace(",",""))) for r in rows[1:]]
tot=sum(v for _,v in seq)
print("launches",len(seq),"total us",tot/1e3)
for k,v in seq: print("%9.1f us  %s"%(v/1e3,k[:110]))
PY
