#!/bin/bash
# parity tests, bench, full-step ncu capture of the map pipeline, compute-sanitizer on the small script
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
( time timeout 600 python bench.py --phases ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 2800 $OUT/${TAG}_bench.json; tail -16 $OUT/${TAG}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'map_kernel|prim_kernel' -s 32 -c 16 -o $OUT/${TAG}_full -f \
	python bench.py --steps 1 --warmup 3 --streams 1 --cpu-sample 0 > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:transpose -c 6 --csv --log-file $OUT/${TAG}_transpose.csv \
	python bench.py --steps 1 --warmup 3 --streams 1 --cpu-sample 0 > $OUT/${TAG}_ncu_t.log 2>&1
grep transpose $OUT/${TAG}_transpose.csv | tail -3 | cut -c1-60,200-
( timeout 400 compute-sanitizer --tool memcheck python tests/_san_small.py ) > $OUT/${TAG}_memcheck.log 2>&1
tail -3 $OUT/${TAG}_memcheck.log
( timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis python tests/_san_small.py ) > $OUT/${TAG}_racecheck.log 2>&1
tail -3 $OUT/${TAG}_racecheck.log
