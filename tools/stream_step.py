"""The streaming stages alone (clust_filt=False) on the configs[2] workload, device-resident: a few calls for an ncu
launch list (`ncu --metrics gpu__time_duration.sum ... python tools/stream_step.py`) and the CUDA-event time per step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from retargetvid_b200 import _cabi  # noqa: E402
from retargetvid_b200 import smartVidCrop as svc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
vds = bench.make_workload(n, 0)
ctx = _cabi.Context(0)
st = torch.cuda.Stream()
ctx.set_stream(st.cuda_stream)
torch.cuda.set_stream(st)
wl = bench.Workload(vds, ['1:3', '3:1'], 1, torch, _cabi)
CP = svc.sc_init_crop_params()
CP['clust_filt'] = False
p = _cabi.params_from_crop_params(CP)
for _ in range(3):
	ctx.crop_track_batch(p, wl.b_dev[0])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10):
	ctx.crop_track_batch(p, wl.b_dev[0])
e1.record(st)
torch.cuda.synchronize()
print('streaming step %.3f ms (%d maps, %d frames)' % (e0.elapsed_time(e1) / 10, wl.NM, wl.NF))
import time  # noqa: E402
t0 = time.perf_counter()
for _ in range(10):
	ctx.crop_track_batch(p, wl.b_dev[0])
t1 = time.perf_counter()
torch.cuda.synchronize()
print('host time per call %.3f ms; sections (us): tables, pack+upload, map input, map launches, track launches, results = %s'
	% ((t1 - t0) * 100, ['%.0f' % v for v in ctx.last_host_us()]))
CP['clust_filt'] = True
p2 = _cabi.params_from_crop_params(CP)
for _ in range(3):
	ctx.crop_track_batch(p2, wl.b_dev[0])
torch.cuda.synchronize()
print('default pipeline, host sections (us): %s' % ['%.0f' % v for v in ctx.last_host_us()])
