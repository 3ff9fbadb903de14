"""Single-clip latency of the drop-in entry point (BASELINE configs[0]) and where it goes: the C-ABI call alone (prebuilt
batch, host buffers), engine.run (packing + call + unpacking) and smart_vid_crop (dict plumbing on top)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from retargetvid_b200 import smartVidCrop as svc, synth  # noqa: E402


def med(fn, n=23):
	lat = []
	for _ in range(n):
		t0 = time.perf_counter()
		fn()
		lat.append((time.perf_counter() - t0) * 1e3)
	lat = sorted(lat[3:])
	return lat[len(lat) // 2], lat[0]


vd1 = synth.make_clip(**synth.config_clips(1)[0])
CP = svc.sc_init_crop_params()
CP['out_ratio'] = '1:3'
eng = svc._engine(0)
print('smart_vid_crop            median %.3f ms  min %.3f' % med(lambda: svc.smart_vid_crop('c1.mp4', CP, save_vid=False, vid_data=dict(vd1))))
print('engine.run detail+filtered median %.3f ms  min %.3f' % med(lambda: eng.run([vd1], CP, ['1:3'], detail=True, want_filtered=True)))
print('engine.run boxes only      median %.3f ms  min %.3f' % med(lambda: eng.run([vd1], CP, ['1:3'], detail=False)))
try:
	print('stages (front, prim+back, -, pipeline) ms', eng.ctx.last_stage_ms(), 'map kernels', eng.ctx.last_map_kernel_ms())
except Exception as e:
	print('stage times unavailable:', e)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
	svc.smart_vid_crop('c1.mp4', CP, save_vid=False, vid_data=dict(vd1))
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
