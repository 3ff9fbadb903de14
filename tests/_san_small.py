"""Small run of every kernel family for compute-sanitizer (tools: memcheck, racecheck); not collected by pytest.
usage: compute-sanitizer --tool memcheck python tests/_san_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from retargetvid_b200 import _cabi, smartVidCrop as svc, synth  # noqa: E402
from retargetvid_b200.engine import CropEngine  # noqa: E402

e = CropEngine(0)
vds = [synth.make_clip(100 + i, fc=40 + 30 * i, shot_starts=[18] if i else []) for i in range(2)]
for best in (False, True):
	CP = svc.sc_init_crop_params(use_best_settings=best)
	r = e.run(vds, CP, ['1:3', '3:1'], detail=True, want_filtered=True)
	print('ok', best, r[0].boxes[0][0], r[1].status)
CP = svc.sc_init_crop_params()
r = e.run(vds, CP, ['1:3'], detail=True, want_filtered='hwn')
print('hwn', r[1].filtered_hwn.shape)
for rtype, factor in ((2, 4), (1, 2), (3, 3)):
	CP = svc.sc_init_crop_params()
	CP.update(dict(t_threshold=90, hdbscan_min=5, hdbscan_min_samples=3, resize_factor=factor, resize_type=rtype))
	print('resize', rtype, factor, e.run(vds, CP, ['4:5'])[0].boxes[0][0])
CP = svc.sc_init_crop_params()
CP['clust_filt'] = False
print(e.run(vds, CP, ['1:3'])[0].boxes[0][0])
# IoU: two videos of odd lengths, 3 annotators, per-frame output
rng = np.random.default_rng(0)
lens = [70, 33]
foff = np.array([0, 70, 103], dtype=np.int64)


def boxes(n):
	b = rng.integers(0, 600, (n, 4)).astype(np.int32)
	b[:, 2:] = np.maximum(b[:, 2:], b[:, :2])
	return b


method, annot = boxes(103), np.stack([boxes(103) for _ in range(3)])
neval = np.array(lens, dtype=np.int32)
fiou = np.empty((3, 103))
acc = np.zeros((2, 3, 2), dtype=np.uint64)
ib = _cabi.rvb_iou_batch()
ib.n_videos, ib.n_users, ib.mem_space = 2, 3, _cabi.RVB_MEM_HOST
ib.frame_offset, ib.n_eval = foff.ctypes.data, neval.ctypes.data
ib.method_boxes, ib.annot_boxes, ib.frame_iou, ib.acc = method.ctypes.data, annot.ctypes.data, fiou.ctypes.data, acc.ctypes.data
e.ctx.iou_batch(ib)
print('iou', float(fiou.mean()))
frames = rng.integers(0, 256, (5, 36, 64, 3)).astype(np.uint8)
bb = np.array([[3 + i, 0, 3 + i + 21, 36] for i in range(5)], dtype=np.int32)
print('crop', e.ctx.crop_frames(frames, bb).shape)
