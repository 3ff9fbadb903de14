#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
( time timeout 600 python bench.py ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 1500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 120 tools/ubench/ub > $OUT/${TAG}_ubench.txt 2>&1
cat $OUT/${TAG}_ubench.txt
