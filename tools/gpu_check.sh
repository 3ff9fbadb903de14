#!/bin/bash
# One GPU-box visit.  Everything lands in gpurun_out/<tag>_*.
#   tools/gpu_check.sh TAG            parity tests, smoke, bench line, ncu launch list of one bench command, launch list of
#                                     the streaming (clust_filt=False) step, single-clip latency
#   tools/gpu_check.sh TAG full       + `ncu --set full` of ONE step's map-pipeline launches, exported to CSV on the box
#                                       (the .ncu-rep is too large to travel) and summarised into
#                                       gpurun_out/<tag>_map_kernel_traffic.json (copy to profiles/map_kernel_traffic.json)
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $OUT/${TAG}_smi.txt 2>&1
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -3
( time timeout 900 python bench.py ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 2500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
	python bench.py --steps 2 --warmup 3 --streams 1 --cpu-sample 0 --c5-clips 0 > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list exit $?"; wc -l $OUT/${TAG}_launches.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'map_stream|fill_centres|spline|interp_eval|lowpass|smooth|boxes_kernel|clip_scores|index_tables|border' -c 120 --csv \
	--log-file $OUT/${TAG}_stream_launches.csv python tools/stream_step.py > $OUT/${TAG}_stream.log 2>&1
python tools/lat_single.py > $OUT/${TAG}_lat.log 2>&1; head -4 $OUT/${TAG}_lat.log
if [ "$2" = "full" ]; then
	timeout 900 ncu --set full --clock-control none -k regex:'map_kernel|prim_kernel' -s 32 -c 16 -o /tmp/${TAG}_full -f \
		python bench.py --steps 1 --warmup 3 --streams 1 --cpu-sample 0 --c5-clips 0 > $OUT/${TAG}_ncu_full.log 2>&1
	echo "full capture exit $?"
	ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
	python tools/ncu_summarise.py $OUT/${TAG}_full_raw.csv $OUT/${TAG}_map_kernel_traffic.json \
		"gpurun_out/${TAG}_full_raw.csv (ncu --set full --clock-control none, 200 clips, ONE step = the 16 launches of the map pipeline: front, prim x5, back x5, mono x5)"
fi
