#!/bin/bash
# Round profile capture (run under gpurun): one ncu --set full capture per hot
# kernel at the size bench.py times it, exported as text / csv (the .ncu-rep files
# are large), a digest of all of them (summary.json), and the launch list of a
# short bench run.  Every capture runs only after the same command exited 0
# without ncu.
# usage: tools/capture_profiles.sh <outdir>
set -u
OUT=${1:-gpurun_out/profiles}
mkdir -p "$OUT"
cap () {  # name, kernel regex, command...
	local name=$1 regex=$2; shift 2
	"$@" > "$OUT/plain_$name.log" 2>&1 || { echo "$name: plain run failed"; tail -3 "$OUT/plain_$name.log"; return; }
	ncu --set full --clock-control none --import-source on -k "regex:$regex" -c 1 -f \
		-o /tmp/prof_$name "$@" > "$OUT/ncu_$name.log" 2>&1
	ncu -i /tmp/prof_$name.ncu-rep --page details > "$OUT/${name}_details.txt" 2>/dev/null
	ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > "$OUT/${name}_raw.csv" 2>/dev/null
	echo "$name: done ($(grep -m1 -E 'Duration' "$OUT/${name}_details.txt" | tr -s ' '))"
}
hot () {  # name
	ncu -i /tmp/prof_$1.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_hot.py 30 > "$OUT/$1_hot.txt"
}
cap chain vm_pair4 python tools/prof_cases.py --case chain --iters 2
hot chain
cap chain4 vm_pair4 python tools/prof_cases.py --case chain4 --iters 2
cap exp vm_flat python tools/prof_cases.py --case exp --iters 2
cap add vm_flat python tools/prof_cases.py --case add --iters 2
cap mulext vm_flat python tools/prof_cases.py --case mulext --iters 2
cap permute vm_tile4 python tools/prof_cases.py --case permute --iters 2
cap rsum_cols fold python tools/prof_cases.py --case rsum_cols --iters 2
cap rsum_rows fold python tools/prof_cases.py --case rsum_rows --iters 2
cap rsum_all fold python tools/prof_cases.py --case rsum_all --iters 2
cap convert convert python tools/prof_cases.py --case convert --iters 2
cap rand rand python tools/prof_cases.py --case rand --iters 2
cap general "gather|general|push|pull" python tools/prof_cases.py --case general --iters 2
cap sgemm_nn sgemm_tma python tools/gemm_bench.py --size 8192 --layouts nn --iters 1
hot sgemm_nn
cap sgemm_persist sgemm_persist python tools/gemm_bench.py --mnk 8192,1024,784 --layouts nn --iters 1
cap dgemm_nn dgemm_tma python tools/gemm_bench.py --size 8192 --dtype f64 --layouts nn --iters 1
cap tf32x3_nn tf32x3_kernel python tools/gemm_bench.py --size 8192 --precision 1 --layouts nn --iters 1
cap skinny_rows skinny_rows python tools/gemm_bench.py --mnk 65536,10,1024 --layouts nn --iters 1
cap skinny_cols skinny_cols python tools/gemm_bench.py --mnk 1024,10,65536 --layouts tn --iters 1
python tools/summarize_profiles.py "$OUT" > "$OUT/summary.json"
python bench.py --steps 3 --warmup 3 --skip-cpu --no-traffic-probe > "$OUT/plain_bench.log" 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
	--log-file "$OUT/launches_bench.csv" python bench.py --steps 3 --warmup 3 --skip-cpu --no-traffic-probe --skip-verify > "$OUT/ncu_bench.log" 2>&1
echo "launch list: done"
