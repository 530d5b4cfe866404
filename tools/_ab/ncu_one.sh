#!/bin/bash
# usage: ncu_one.sh <name> <kernel regex> <command...>   (one ncu capture per gpurun call)
name=$1; regex=$2; shift 2
OUT=gpurun_out/ncu_$name; mkdir -p $OUT
"$@" > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:$regex" -c 1 -f -o /tmp/prof_$name "$@" > $OUT/ncu.log 2>&1
ncu -i /tmp/prof_$name.ncu-rep --page details > $OUT/details.txt 2>/dev/null
ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i /tmp/prof_$name.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_hot.py 40 > $OUT/hot.txt
tail -3 $OUT/plain.log
