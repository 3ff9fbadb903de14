#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one full ncu capture of the top kernels.
# Everything lands in gpurun_out/<tag>_*.
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $OUT/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
( time timeout 600 python bench.py --phases ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.json
tail -20 $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
	python bench.py --steps 1 --warmup 3 --streams 1 --cpu-sample 0 > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list exit $?"
if [ "$2" = "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'map_|prim' -s 30 -c 12 -o $OUT/${TAG}_full -f \
	python bench.py --steps 1 --warmup 3 --streams 1 --cpu-sample 0 > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture exit $?"
fi
