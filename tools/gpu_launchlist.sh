#!/bin/bash
TAG=${1:-ll}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
	python bench.py --steps 2 --warmup 3 --streams 1 --cpu-sample 0 > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list exit $?"; wc -l $OUT/${TAG}_launches.csv
