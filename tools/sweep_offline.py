"""Second half of a sweep whose oracle side was computed elsewhere (no GPU needed for that part):
    python tools/sweep_offline.py make first:last out.pkl      # oracle on the host cores -> compact expectations
    python tools/sweep_offline.py check out.pkl                # CUDA against them (GPU box)
Filtered maps are compared through per-map SHA-1 digests."""
import hashlib
import multiprocessing as mp
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_at_size as t  # noqa: E402


def _digest(filt_hwn):
	return [hashlib.sha1(np.ascontiguousarray(filt_hwn[:, :, m]).tobytes()).hexdigest() for m in range(filt_hwn.shape[2])]


def _one(i):
	vd, over, ratios, ex = t._random_case(i)
	w = t._oracle_or_error((vd, over, ratios, ex['cvrg_window'], ex['np_int']))
	if isinstance(w, str):
		return i, w
	return i, dict(filt=_digest(w[0]['filt']), dxs=w[0]['dxs'], bbs=[x['bbs'] for x in w], cvrg=[x['cvrg'] for x in w])


if sys.argv[1] == 'make':
	a, b = sys.argv[2].split(':')
	with mp.get_context('fork').Pool(os.cpu_count() or 1) as pool:
		res = dict(pool.map(_one, range(int(a), int(b)), chunksize=2))
	pickle.dump(res, open(sys.argv[3], 'wb'))
	print('cases', len(res), 'reference raises on', sum(isinstance(v, str) for v in res.values()))
else:
	from retargetvid_b200 import _cabi, smartVidCrop as svc
	from retargetvid_b200.engine import CropEngine
	res = pickle.load(open(sys.argv[2], 'rb'))
	e = CropEngine(0)
	bad = 0
	for i in sorted(res):
		want = res[i]
		vd, over, ratios, ex = t._random_case(i)
		CP = svc.sc_init_crop_params()
		CP.update(over)
		r = e.run([vd], CP, ratios, detail=True, want_filtered='hwn', raise_on_clip_error=False, cvrg_window=ex['cvrg_window'], np_int=ex['np_int'])[0]
		if isinstance(want, str) or r.status != 0:
			ok = (isinstance(want, str) and r.status == _cabi.RVB_ERR_NO_CENTRES) or r.status == _cabi.RVB_ERR_CAPACITY
			if not ok:
				bad += 1
			print(i, 'status', r.status, 'oracle', want if isinstance(want, str) else 'ok', '' if ok else 'MISMATCH')
			continue
		dm = sum(a != b for a, b in zip(_digest(r.filtered_hwn), want['filt']))
		dxs = float(np.max(np.abs(r.series[4] - want['dxs'])))
		db = [int((r.boxes[k] != want['bbs'][k]).any(axis=1).sum()) for k in range(len(ratios))]
		dc = [int(float(r.cvrg_scores[k]) != want['cvrg'][k]) for k in range(len(ratios))] if CP['exit_on_low_cvrg'] else []
		if dm or dxs > 1e-6 or any(db) or any(dc):
			bad += 1
			print(i, 'maps', dm, 'dxs %.2e' % dxs, 'boxes', db, 'cvrg', dc, 'size %dx%d' % (vd['h_process'], vd['w_process']), over, ex)
	print('cases', len(res), 'bad', bad)
