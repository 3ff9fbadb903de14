#!/bin/bash
# full-step ncu capture of the map pipeline (exported to CSV on the box; the .ncu-rep is too large to travel), bench
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none -k regex:'map_kernel|prim_kernel' -s 32 -c 16 -o /tmp/${TAG}_full -f \
	python bench.py --steps 1 --warmup 3 --streams 1 --cpu-sample 0 > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture exit $?"
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
ls -la /tmp/${TAG}_full.ncu-rep $OUT/${TAG}_full_raw.csv
shift
for extra in "$@"; do
	echo "== $extra"
	( time timeout 900 bash -c "$extra" ) 2>&1 | tail -c 4000
done
