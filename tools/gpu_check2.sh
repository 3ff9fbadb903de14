#!/bin/bash
# shorter visit: parity tests + bench (+ optional extra bench variants given as further args)
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
( time timeout 600 python bench.py ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 2500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
( time timeout 600 python bench.py --steps 15 --cpu-sample 0 ) > $OUT/${TAG}_bench_k15.json 2> $OUT/${TAG}_bench_k15.err
tail -c 2500 $OUT/${TAG}_bench_k15.json
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
tail -c 1500 $OUT/${TAG}_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:transpose -c 8 --csv --log-file $OUT/${TAG}_transpose.csv \
	python bench.py --steps 1 --warmup 3 --streams 1 --cpu-sample 0 > $OUT/${TAG}_ncu_t.log 2>&1
grep transpose $OUT/${TAG}_transpose.csv | tail -3
