#!/bin/bash
# A/B of the Prim launch shapes (RVB_PRIM_VARIANT, one digit per size class)
OUT=gpurun_out
mkdir -p $OUT
for v in 0000 1000 0100 0010 0001 1111 0000; do
	RVB_PRIM_VARIANT=$v timeout 300 python bench.py --steps 10 --cpu-sample 0 > $OUT/var_$v.json 2> $OUT/var_$v.err
	python - <<PY
import json
d=json.load(open('$OUT/var_$v.json'))
print('$v', 'value %.3f M  %.2f ms/step   e2e %.3f M  map pipeline %.2f ms' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['kernel_ms_per_step']))
PY
done
