"""Single-clip latency of the drop-in entry point (BASELINE configs[0]) with a per-stage split."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from retargetvid_b200 import smartVidCrop as svc, synth
vd1 = synth.make_clip(**synth.config_clips(1)[0])
CP = svc.sc_init_crop_params(); CP['out_ratio'] = '1:3'
lat = []
for i in range(23):
	t0 = time.perf_counter()
	VD, res = svc.smart_vid_crop('c1.mp4', CP, save_vid=False, vid_data=dict(vd1))
	lat.append((time.perf_counter() - t0) * 1e3)
lat = sorted(lat[3:])
eng = svc._engine(0)
st = eng.ctx.last_stage_ms() if hasattr(eng.ctx, 'last_stage_ms') else None
print(os.environ.get('RVB_CHAIN_LEVELS'), os.environ.get('RVB_NO_SPLIT'), 'median %.2f ms  min %.2f  stages(front,prim,back,pipeline) %s' % (lat[len(lat)//2], lat[0], st))
