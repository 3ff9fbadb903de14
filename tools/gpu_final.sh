#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, both bench arms
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
tail -c 700 $OUT/${TAG}_bench_ref.json
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --phases ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value %.3f M (%.2f ms)  e2e %.3f M (%.2f ms)  launches %d' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['gpu_launches']))
print('roofline', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'stream', d['roofline_streaming_stages']['frac'], 'iou', d['iou_stage']['frac'], 'crop', d['crop_stage']['frac'])
print('single', d['single_clip']['ms_per_clip'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
PY
tail -16 $OUT/${TAG}_bench.err
