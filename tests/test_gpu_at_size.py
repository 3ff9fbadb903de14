"""Parity of the CUDA path against the oracle on the workloads bench.py times (BASELINE.json configs[2], [3], [4]),
not on look-alikes: the clips whose cut blend (smartVidCrop.py:2369-2373) inflates a map into the largest point-count
classes, the 10 000-frame 1080p multi-shot clip with LOESS and the padding fallback, and corpus clips x 4 ratios.
Needs a GPU: pytest -m gpu.  The oracle runs one clip per process on the host cores."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from helpers import loess_tolerance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
	from retargetvid_b200.engine import CropEngine
	e = CropEngine(0)
	yield e
	e.close()


def _oracle_one(args):
	from oracle import sc_oracle
	vd, over, ratios, cvrg_window = args[:4]
	np_int = args[4] if len(args) > 4 else False
	outs = []
	for r in ratios:
		CP = sc_oracle.sc_init_crop_params()
		CP.update(over)
		CP['out_ratio'] = r
		o = sc_oracle.smart_vid_crop_oracle(vd, CP, cvrg_window=cvrg_window, np_int=np_int)
		outs.append(dict(bbs=np.array(o['bbs'], dtype=np.int32), filt=o['smaps_filtered'], dxs=np.array(o['dxs'], dtype=np.float64),
						dys=np.array(o['dys'], dtype=np.float64), cvrg=o.get('mean_cvrg_score'),
						dx=np.array([np.nan if v is None else v for v in o['dx']], dtype=np.float64)))
	return outs


def _oracle_many(vds, over, ratios, cvrg_window='reference'):
	procs = min(len(vds), os.cpu_count() or 1)
	with mp.get_context('fork').Pool(procs) as pool:
		return pool.map(_oracle_one, [(vd, over, ratios, cvrg_window) for vd in vds], chunksize=1)


def _loess_tol(vd, CP):
	"""the largest loess_tolerance over the shots of a clip, with the window smartVidCrop.py:1669-1671 gives each"""
	tol = 1e-9
	for s0, s1 in np.asarray(vd['segmentation']):
		cl = int(s1 - s0 + 1)
		if cl < 10:
			continue
		win = min(int(float(vd['fr']) * float(CP['loess_w_secs'])), cl - 2)
		win -= (win % 2 == 0)
		tol = max(tol, loess_tolerance(cl, win))
	return tol


def _check_clip(res, want, ratios, vd, tag, tol_floor=1e-8):
	filt = np.transpose(res.filtered, (1, 2, 0))
	assert np.array_equal(filt, want[0]['filt']), '%s: filtered maps differ in %d maps' % (
		tag, int((filt != want[0]['filt']).any(axis=(0, 1)).sum()))
	assert np.max(np.abs(res.dx - want[0]['dx'])) <= 1e-9, tag
	max_cl = int(max(s[1] - s[0] + 1 for s in vd['segmentation']))
	tol = max(loess_tolerance(max_cl), tol_floor)
	assert np.max(np.abs(res.series[4] - want[0]['dxs'])) <= tol, (tag, np.max(np.abs(res.series[4] - want[0]['dxs'])), tol)
	assert np.max(np.abs(res.series[5] - want[0]['dys'])) <= tol, tag
	scale = min(vd['w_process'] / vd['w_orig'], vd['h_process'] / vd['h_orig'])
	n_flips = 0
	for k, r in enumerate(ratios):
		got, ref = res.boxes[k], want[k]['bbs']
		diff = np.nonzero((got != ref).any(axis=1))[0]
		for f in diff:
			# a box may differ by one pixel only where the reference centre sits within the LOESS tolerance of an
			# integer before int() (smartVidCrop.py:998-999)
			vx = want[0]['dxs'][f] / (vd['w_process'] / vd['w_orig'])
			vy = want[0]['dys'][f] / (vd['h_process'] / vd['h_orig'])
			near = min(abs(vx - round(vx)), abs(vy - round(vy)))
			assert near <= tol / scale, (tag, r, int(f), got[f], ref[f])
			assert np.max(np.abs(got[f] - ref[f])) <= 1
		n_flips += len(diff)
	return n_flips


def test_config3_bench_clips_vs_oracle(engine):
	"""configs[2] exactly as bench.py builds it: the 8 clips with the largest post-blend point counts (the maps of
	the 3072 / 4096 / 8192 classes with the default min_cluster_size 26) and 4 random ones: filtered maps and boxes
	of both ratios bit-exact against the oracle."""
	import bench
	from retargetvid_b200 import smartVidCrop as svc
	vds = bench.make_workload(200, 0)
	CP = svc.sc_init_crop_params()
	ratios = ['1:3', '3:1']
	res = engine.run(vds, CP, ratios, detail=True)
	peak = np.array([int(r.map_info[:, 0].max()) for r in res])
	assert all(r.status == 0 for r in res)
	order = np.argsort(-peak, kind='stable')
	pick = [int(i) for i in order[:8]]
	rng = np.random.default_rng(1)
	pick += [int(i) for i in rng.choice([int(i) for i in order[8:]], 4, replace=False)]
	assert peak[pick[0]] > 3072, 'the workload no longer reaches the large capacity classes (%d)' % peak[pick[0]]
	sub = [vds[i] for i in pick]
	got = engine.run(sub, CP, ratios, detail=True, want_filtered=True)
	want = _oracle_many(sub, {}, ratios)
	flips = 0
	for i, g, w in zip(pick, got, want):
		# the batch of 200 and the batch of 12 must agree with each other too
		assert np.array_equal(g.boxes, res[i].boxes)
		flips += _check_clip(g, w, ratios, vds[i], 'c3 clip %d (peak %d points)' % (i, peak[i]))
	assert flips == 0, flips


def test_config4_long_multishot_clip_vs_oracle(engine):
	"""configs[3]: one 10 000-frame 1920x1080 clip, shots of 60-600 frames, 9:16 target, LOESS on, coverage score
	with the crop-sized window and the padding decision."""
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vd = synth.make_clip(**synth.config_clips(4)[0])
	assert vd['fc'] == 10000 and len(vd['segmentation']) > 20
	over = dict(exit_on_low_cvrg=True, loess_filt=1)
	CP = svc.sc_init_crop_params()
	CP.update(over)
	ratios = ['9:16']
	res = engine.run([vd], CP, ratios, detail=True, want_filtered=True, cvrg_window='crop')[0]
	assert res.status == 0
	want = _oracle_many([vd], over, ratios, cvrg_window='crop')[0]
	flips = _check_clip(res, want, ratios, vd, 'c4')
	assert flips <= 2, flips
	assert float(res.cvrg_scores[0]) == want[0]['cvrg']
	# the drop-in entry point takes the same padding decision as the oracle's score implies
	CP['out_ratio'] = '9:16'
	VD, info = svc.smart_vid_crop('c4.mp4', CP, save_vid=False, vid_data=dict(vd), cvrg_window='crop')
	assert info['coverage_score'] == want[0]['cvrg']
	assert info['result'] == ('padded' if want[0]['cvrg'] < CP['t_cvrg'] else 'smart cropped')
	if info['result'] == 'smart cropped':
		assert np.array_equal(np.array(VD['bbs'], dtype=np.int32), res.boxes[0])


def test_config5_corpus_clips_vs_oracle(engine):
	"""configs[4]: 12 clips of the 2 000-clip corpus x {1:3, 3:1, 9:16, 4:5}, all four ratios from one pass."""
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	specs = synth.config_clips(5)
	rng = np.random.default_rng(5)
	pick = sorted(int(i) for i in rng.choice(len(specs), 12, replace=False))
	vds = [synth.make_clip(**specs[i]) for i in pick]
	CP = svc.sc_init_crop_params()
	ratios = ['1:3', '3:1', '9:16', '4:5']
	got = engine.run(vds, CP, ratios, detail=True, want_filtered=True)
	want = _oracle_many(vds, {}, ratios)
	flips = 0
	for i, g, w, vd in zip(pick, got, want, vds):
		assert g.status == 0
		flips += _check_clip(g, w, ratios, vd, 'c5 clip %d' % i)
	assert flips == 0, flips


N_RANDOM_CASES = int(os.environ.get('RVB_TEST_RANDOM_CASES', '32'))


def _oracle_or_error(args):
	"""the oracle's outputs, or 'TypeError' where the reference raises on a clip without any salient pixel: TypeError from
	float(None) in interp_handler (smartVidCrop.py:1533) when a shot has fewer than 3 maps, else ValueError from int(NaN) in
	sc_compute_bb (:998) after interp1d turned the Nones into NaNs -- the library reports both as RVB_ERR_NO_CENTRES"""
	try:
		return _oracle_one(args)
	except TypeError:
		return 'TypeError'
	except ValueError as e:
		if 'NaN' in str(e):
			return 'TypeError'
		raise


def _random_case(i):
	"""Seeded random crop parameters + clip geometry: values a user of the reference can set (smartVidCrop.py:135-183)."""
	from retargetvid_b200 import synth
	rng = np.random.default_rng(77000 + i)
	over = dict(
		t_threshold=int(rng.choice([60, 90, 120, 150, 200])),
		hdbscan_min=int(rng.choice([5, 12, 26, 40])),
		hdbscan_min_samples=[None, 3, 8][int(rng.integers(0, 3))],
		select_sum=int(rng.choice([1, 2])),
		op_close=bool(rng.integers(0, 2)),
		com_km=bool(rng.integers(0, 4) > 0),
		value_bias=float(rng.choice([1.0, 0.5])),
		lp_filt=int(rng.integers(0, 4) > 0),
		lp_order=int(rng.choice([2, 3, 5, 7])),
		lp_cutoff=float(rng.choice([1.0, 2.0, 3.5])),
		loess_filt=int(rng.integers(0, 3) > 0),
		loess_degree=int(rng.choice([1, 2])),
		loess_w_secs=float(rng.choice([1, 2, 3])),
		shift_time=int(rng.choice([0, 0, 4])),
		t_border=int(rng.choice([-1, -1, 10])),
	)
	if over['hdbscan_min_samples'] is None:
		del over['hdbscan_min_samples']
	if rng.integers(0, 4) == 0:
		over.update(resize_factor=int(rng.choice([2, 3, 4])), resize_type=int(rng.choice([1, 2, 3])))
	if rng.integers(0, 5) == 0:
		over['clust_filt'] = False
	extra = dict(np_int=False, cvrg_window='reference')
	if rng.integers(0, 3) == 0:
		# focus stability (smartVidCrop.py:1337-1455, 2424-2473), with the numpy the reference pins (np.int exists) or a newer one
		over.update(focus_stability=True, foces_stab_t=float(rng.choice([30, 60, 100])), foces_stab_s=float(rng.choice([0.5, 1.0, 2.0])),
					min_d_jump=int(rng.choice([1, 5, 10])))
		extra['np_int'] = bool(rng.integers(0, 2))
	if rng.integers(0, 4) == 0:
		over.update(exit_on_low_cvrg=True, t_cvrg=float(rng.choice([0.3, 0.6])))
		extra['cvrg_window'] = ['reference', 'crop'][int(rng.integers(0, 2))]
	fc = int(rng.integers(40, 150))
	n_cuts = int(rng.integers(0, 4))
	cuts = sorted(set(int(v) for v in rng.integers(3, fc - 3, n_cuts)))
	size = [(640, 360), (1920, 1080), (480, 360), (360, 640)][int(rng.integers(0, 4))]
	vd = synth.make_clip(88000 + i, fc=fc, fr=float(rng.choice([12.0, 24.0, 25.0, 30.0, 60.0])), w_orig=size[0], h_orig=size[1],
						shot_starts=cuts, skip=int(rng.choice([1, 3, 6, 6, 8])),
						kind=['blobs', 'blobs', 'blobs', 'noise', 'few_points', 'single_pixel'][int(rng.integers(0, 6))])
	if rng.integers(0, 3) == 0:
		# maps without a salient pixel at the start, in the middle, at the end (sc_handle_empty_centers, smartVidCrop.py:1221-1300)
		n = vd['fc_sel']
		for m in set([0, int(rng.integers(0, n)), int(rng.integers(0, n)), n - 1][:int(rng.integers(1, 5))]):
			vd['smaps'][:, :, m] = 0
	if rng.integers(0, 6) == 0:
		over['t_threshold'] = float(over['t_threshold']) + 0.5      # smaps < t with a fractional t
	ratios = [['1:3', '3:1'], ['9:16'], ['4:5', '1:1'], ['16:9', '1:3']][int(rng.integers(0, 4))]
	return vd, over, ratios, extra


def test_random_parameter_sweep_vs_oracle(engine):
	"""N_RANDOM_CASES seeded random combinations of crop parameters, frame rates, sampling steps, frame sizes (landscape, portrait,
	4:3) and shot layouts: the CUDA path against the oracle, every case on its own (integer stages bit-exact, float
	stages within the stated tolerances)."""
	from retargetvid_b200 import _cabi
	from retargetvid_b200 import smartVidCrop as svc
	cases = [_random_case(i) for i in range(N_RANDOM_CASES)]
	procs = min(len(cases), os.cpu_count() or 1)
	with mp.get_context('fork').Pool(procs) as pool:
		wants = pool.map(_oracle_or_error, [(vd, over, ratios, ex['cvrg_window'], ex['np_int']) for vd, over, ratios, ex in cases], chunksize=1)
	flips = 0
	failures = []
	over_capacity = []
	for i, ((vd, over, ratios, ex), want) in enumerate(zip(cases, wants)):
		CP = svc.sc_init_crop_params()
		CP.update(over)
		res = engine.run([vd], CP, ratios, detail=True, want_filtered=True, raise_on_clip_error=False,
						cvrg_window=ex['cvrg_window'], np_int=ex['np_int'])[0]
		if isinstance(want, str):
			if res.status != _cabi.RVB_ERR_NO_CENTRES:
				failures.append((i, 'the reference raises TypeError, status %d' % res.status, over))
			continue
		if res.status == _cabi.RVB_ERR_CAPACITY:
			# the documented limit: a map with more than 8 192 salient pixels fails that clip
			over_capacity.append(i)
			continue
		if res.status != 0:
			failures.append((i, 'status %d' % res.status, over))
			continue
		# A Butterworth filter of high order and low cut-off is ill conditioned in transfer-function form: scipy's own output
		# then depends on LAPACK's rounding inside lfilter_zi and on the order of the filter's operations at the 1e-7 .. 1e-5
		# level, so that is how far any restatement can be pinned (the library says which filters these are)
		tol_floor = 1e-8
		if CP['lp_filt']:
			well = _cabi.debug_butter(int(CP['lp_order']), float(CP['lp_cutoff']) / (0.5 * float(vd['fr'])))[3]
			tol_floor = 1e-8 if well else 2e-5
		try:
			flips += _check_clip(res, want, ratios, vd, 'case %d' % i, max(tol_floor, _loess_tol(vd, CP)))
			if CP['exit_on_low_cvrg']:
				for k in range(len(ratios)):
					assert float(res.cvrg_scores[k]) == want[k]['cvrg'], ('coverage score', ratios[k])
		except AssertionError as e:
			failures.append((i, str(e)[:300], over))
	assert not failures, failures
	# (every flip was checked above: one pixel, at a frame whose reference centre sits within the tolerance of an integer
	# before int() -- common when com_km=False makes the centres integers and focus stability freezes them)
	assert flips <= sum(vd['fc'] for vd, _, _, _ in cases) // 50, flips
	assert len(over_capacity) <= max(1, len(cases) // 16), over_capacity


def test_random_batches_vs_oracle(engine):
	"""4 calls of 6 random clips each -- mixed frame sizes (several process sizes in one call), lengths, cuts and map
	kinds -- under one random parameter set per call: every clip equals the oracle run on that clip alone."""
	from retargetvid_b200 import _cabi
	from retargetvid_b200 import smartVidCrop as svc
	meta, jobs = [], []
	for b in range(4):
		_, over, ratios, ex = _random_case(5000 + b)
		clips = [_random_case(6000 + 8 * b + j)[0] for j in range(6)]
		meta.append((over, ratios, ex, clips))
		jobs += [(vd, over, ratios, ex['cvrg_window'], ex['np_int']) for vd in clips]
	with mp.get_context('fork').Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
		wants = pool.map(_oracle_or_error, jobs, chunksize=1)
	failures = []
	k = 0
	for b, (over, ratios, ex, clips) in enumerate(meta):
		CP = svc.sc_init_crop_params()
		CP.update(over)
		rs = engine.run(clips, CP, ratios, detail=True, want_filtered=True, raise_on_clip_error=False,
						cvrg_window=ex['cvrg_window'], np_int=ex['np_int'])
		for j, (vd, res) in enumerate(zip(clips, rs)):
			want = wants[k]
			k += 1
			if isinstance(want, str):
				if res.status != _cabi.RVB_ERR_NO_CENTRES:
					failures.append((b, j, 'the reference raises TypeError, status %d' % res.status))
				continue
			if res.status == _cabi.RVB_ERR_CAPACITY:
				continue
			tol_floor = 1e-8
			if CP['lp_filt'] and not _cabi.debug_butter(int(CP['lp_order']), float(CP['lp_cutoff']) / (0.5 * float(vd['fr'])))[3]:
				tol_floor = 2e-5
			try:
				assert res.status == 0
				_check_clip(res, want, ratios, vd, 'batch %d clip %d' % (b, j), max(tol_floor, _loess_tol(vd, CP)))
			except AssertionError as e:
				failures.append((b, j, str(e)[:300], over))
	assert not failures, failures
