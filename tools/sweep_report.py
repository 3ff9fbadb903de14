"""Runs the random-parameter sweep of tests/test_gpu_at_size.py case by case and prints one line per failing case
(stage that differs first, parameters): python tools/sweep_report.py [n_cases] [case,case,...|first:last|-] [batches]
(`batches`: also 10 calls of 8 random clips each -- mixed frame sizes, lengths and cuts -- under one random parameter set)"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_at_size as t  # noqa: E402
from retargetvid_b200 import _cabi, smartVidCrop as svc  # noqa: E402
from retargetvid_b200.engine import CropEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
only = None
if len(sys.argv) > 2 and sys.argv[2] not in ('', '-'):
	if ':' in sys.argv[2]:
		a, b = sys.argv[2].split(':')
		only = list(range(int(a), int(b)))
	else:
		only = [int(v) for v in sys.argv[2].split(',')]
idx = only if only else list(range(n))
cases = [t._random_case(i) for i in idx]
with mp.get_context('fork').Pool(max(1, min(len(cases), os.cpu_count() or 1))) as pool:
	wants = pool.map(t._oracle_or_error, [(vd, over, ratios, ex['cvrg_window'], ex['np_int']) for vd, over, ratios, ex in cases], chunksize=1)
e = CropEngine(0)
bad = 0
for i, ((vd, over, ratios, ex), want) in zip(idx, zip(cases, wants)):
	CP = svc.sc_init_crop_params()
	CP.update(over)
	res = e.run([vd], CP, ratios, detail=True, want_filtered=True, raise_on_clip_error=False, cvrg_window=ex['cvrg_window'], np_int=ex['np_int'])[0]
	if res.status != 0 or isinstance(want, str):
		print(i, 'status', res.status, 'oracle', want if isinstance(want, str) else 'ok')
		continue
	filt = np.transpose(res.filtered, (1, 2, 0))
	dm = (filt != want[0]['filt']).any(axis=(0, 1))
	ddx = float(np.nanmax(np.abs(res.dx - want[0]['dx']))) if len(res.dx) else 0.0
	dxs = float(np.max(np.abs(res.series[4] - want[0]['dxs'])))
	db = [int((res.boxes[k] != want[k]['bbs']).any(axis=1).sum()) for k in range(len(ratios))]
	if CP['exit_on_low_cvrg']:
		db += [int(float(res.cvrg_scores[k]) != want[k]['cvrg']) for k in range(len(ratios))]
	if dm.any() or ddx > 1e-9 or dxs > 1e-6 or any(db):
		bad += 1
		npts = [int((want[0]['filt'][:, :, m] > 0).sum()) for m in np.nonzero(dm)[0][:4]]
		if only and len(only) <= 16:
			print('  ratios', ratios, 'cvrg got', [float(v) for v in res.cvrg_scores], 'want', [w['cvrg'] for w in want], 'dims', [list(d) for d in res.dims])
			for k in range(len(ratios)):
				df = np.nonzero((res.boxes[k] != want[k]['bbs']).any(axis=1))[0]
				if len(df):
					print('  ratio', ratios[k], 'frames', df[:8].tolist(), 'got', res.boxes[k][df[0]].tolist(), 'want', want[k]['bbs'][df[0]].tolist(), 'dxs', float(want[0]['dxs'][df[0]]), float(res.series[4][df[0]]), 'dys', float(want[0]['dys'][df[0]]), float(res.series[5][df[0]]))
		print(i, 'maps differing %d of %d (first %s, kept points there %s)' % (int(dm.sum()), len(dm), np.nonzero(dm)[0][:6].tolist(), npts),
			'dx %.2e dxs %.2e boxes %s' % (ddx, dxs, db), 'size %dx%d' % (vd['h_process'], vd['w_process']), over, ex)
print('cases', n, 'bad', bad)

# ---- batches: 8 random clips (mixed frame sizes, lengths, cuts) x one random parameter set per call ----
if len(sys.argv) > 3 and sys.argv[3] == 'batches':
	nb = 10
	jobs, meta = [], []
	for b in range(nb):
		_, over, ratios, ex = t._random_case(5000 + b)
		clips = [t._random_case(6000 + 8 * b + j)[0] for j in range(8)]
		meta.append((over, ratios, ex, clips))
		jobs += [(vd, over, ratios, ex['cvrg_window'], ex['np_int']) for vd in clips]
	with mp.get_context('fork').Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
		wants = pool.map(t._oracle_or_error, jobs, chunksize=1)
	bad = 0
	for b, (over, ratios, ex, clips) in enumerate(meta):
		CP = svc.sc_init_crop_params()
		CP.update(over)
		rs = e.run(clips, CP, ratios, detail=True, want_filtered=True, raise_on_clip_error=False, cvrg_window=ex['cvrg_window'], np_int=ex['np_int'])
		for j, (vd, res) in enumerate(zip(clips, rs)):
			want = wants[8 * b + j]
			if isinstance(want, str) or res.status != 0:
				ok = isinstance(want, str) and res.status == _cabi.RVB_ERR_NO_CENTRES or res.status == _cabi.RVB_ERR_CAPACITY
				print('batch', b, 'clip', j, 'status', res.status, 'oracle', want if isinstance(want, str) else 'ok', '' if ok else 'MISMATCH')
				bad += 0 if ok else 1
				continue
			filt = np.transpose(res.filtered, (1, 2, 0))
			dm = int((filt != want[0]['filt']).any(axis=(0, 1)).sum())
			dxs = float(np.max(np.abs(res.series[4] - want[0]['dxs'])))
			db = [int((res.boxes[k] != want[k]['bbs']).any(axis=1).sum()) for k in range(len(ratios))]
			if dm or dxs > 1e-6 or any(db):
				bad += 1
				print('batch', b, 'clip', j, 'maps', dm, 'dxs %.2e' % dxs, 'boxes', db, 'size %dx%d' % (vd['h_process'], vd['w_process']), over, ex)
	print('batches', nb, 'clips', 8 * nb, 'bad', bad)
