"""Parity of the CUDA path (through the C ABI) against the golden fixtures produced by the
unmodified reference and against the oracle.  Needs a GPU: pytest -m gpu."""
import os
import statistics

import numpy as np
import pytest

from helpers import CLIP_NAMES, GOLDEN, fixture_np_int, load_clip_fixture, loess_tolerance

pytestmark = pytest.mark.gpu

GOLD_1_3 = [0.5063503965471644, 0.5085459028456193, 0.48639442473709943, 0.4985176465467487,
			0.49074824192623134, 0.5055689465632044]
GOLD_3_1 = [0.7071752389525412, 0.7247261735288933, 0.7027747813757874, 0.7011597429628194,
			0.7137894895151101, 0.7360647297774207]


@pytest.fixture(scope='module')
def engine():
	from retargetvid_b200.engine import CropEngine
	e = CropEngine(0)
	yield e
	e.close()


def _eval_fixture():
	z = np.load(os.path.join(GOLDEN, 'eval_fixture.npz'))
	vids = [int(v) for v in z['vid_inds']]
	annots = []
	for u in range(1, 7):
		per = {}
		for ar in ('1-3', '3-1'):
			boxes = z['annot_%d_%s' % (u, ar)].astype(np.int32)
			lens = z['annot_len_%d_%s' % (u, ar)]
			offs = np.concatenate([[0], np.cumsum(lens)])
			per[ar] = {v: boxes[offs[i]:offs[i + 1]] for i, v in enumerate(vids)}
		annots.append(per)
	method = {}
	for ar in ('1-3', '3-1'):
		boxes = z['method_%s' % ar].astype(np.int32)
		lens = z['method_len_%s' % ar]
		offs = np.concatenate([[0], np.cumsum(lens)])
		method[ar] = {v: boxes[offs[i]:offs[i + 1]] for i, v in enumerate(vids)}
	frame_counts = {v: len(annots[0]['1-3'][v]) for v in vids}
	return method, annots, frame_counts, str(z['eval_csv'])


def test_iou_golden_vector(engine):
	"""Config 2: shipped results vs 6 annotators, every number of BASELINE.md section 2 bit for bit."""
	from retargetvid_b200 import retargetvid_eval as rev
	method, annots, frame_counts, csv = _eval_fixture()
	ev = rev.evaluate_run(engine.ctx, method, annots, frame_counts)
	assert ev['1-3']['per_user'] == GOLD_1_3
	assert ev['3-1']['per_user'] == GOLD_3_1
	assert ev['1-3']['mean'] == 49.93542598610112
	assert ev['3-1']['mean'] == 71.42816926854287
	line = csv.splitlines()[-1].split(',')
	assert ['%05.3f' % ev['1-3'][k] for k in ('worst', 'best', 'mean')] == line[1:4]
	assert ['%05.3f' % ev['3-1'][k] for k in ('worst', 'best', 'mean')] == line[12:15]


def test_iou_per_frame_and_exact_mean_vs_oracle(engine):
	from oracle import eval_oracle
	from retargetvid_b200 import retargetvid_eval as rev
	method, annots, frame_counts, _ = _eval_fixture()
	vids = sorted(method['1-3'].keys())[:12]
	sub = {v: method['1-3'][v] for v in vids}
	ann = [{v: a['1-3'][v] for v in vids} for a in annots]
	vid_iou, fiou, order = rev.evaluate_arrays(engine.ctx, sub, ann, frame_counts, want_frame_iou=True)
	off = 0
	for i, v in enumerate(order):
		n = frame_counts[v]
		for u in range(6):
			ref = [eval_oracle.bb_intersection_over_union(eval_oracle.clamp0(list(ann[u][v][f])),
														eval_oracle.clamp0(list(sub[v][f]))) for f in range(n)]
			assert list(fiou[u, off:off + n]) == ref
			assert vid_iou[i][u] == statistics.mean(ref)
		off += n


def test_iou_edge_cases(engine):
	"""negative coordinates are clamped, disjoint boxes give 0, short method files stop the mean early."""
	from oracle import eval_oracle
	from retargetvid_b200 import retargetvid_eval as rev
	rng = np.random.default_rng(5)
	method = {1: rng.integers(-50, 700, (40, 4)).astype(np.int32), 2: np.array([[0, 0, 10, 10]] * 7, dtype=np.int32)}
	for v in method:
		method[v][:, 2:] = np.maximum(method[v][:, 2:], method[v][:, :2])
	annots = [{1: rng.integers(0, 640, (50, 4)).astype(np.int32), 2: np.array([[100, 100, 120, 130]] * 9, dtype=np.int32)}]
	annots[0][1][:, 2:] = np.maximum(annots[0][1][:, 2:], annots[0][1][:, :2])
	fc = {1: 50, 2: 9}
	vid_iou, _, order = rev.evaluate_arrays(engine.ctx, method, annots, fc)
	for i, v in enumerate(order):
		assert vid_iou[i][0] == eval_oracle.video_iou([list(b) for b in method[v]], [list(b) for b in annots[0][v]], fc[v])


def test_cluster_labels_match_library_fixture(engine):
	"""sc_clustering_filt's fit_predict (smartVidCrop.py:1099): labels equal to the stand-in library's."""
	from retargetvid_b200 import _cabi
	z = np.load(os.path.join(GOLDEN, 'hdbscan_fixture.npz'))
	bad = []
	for i in range(int(z['count'])):
		P = z['P_%d' % i].astype(np.int64)
		want = z['L_%d' % i].astype(np.int64)
		mcs, ms = [int(v) for v in z['cfg_%d' % i]]
		m = np.zeros((140, 250), dtype=np.uint8)
		m[P[:, 0], P[:, 1]] = 200
		p = _cabi.rvb_params()
		engine.ctx.lib.rvb_params_default(p, 0)
		p.hdbscan_min = mcs
		p.hdbscan_min_samples = 0 if ms < 0 else ms
		got = engine.ctx.debug_cluster_labels(p, m)
		if not np.array_equal(got, want):
			bad.append((i, len(P), int((got != want).sum())))
	assert not bad, bad


def _nan_close(a, b, tol):
	a = np.asarray(a, dtype=np.float64)
	b = np.asarray(b, dtype=np.float64)
	assert a.shape == b.shape
	return float(np.max(np.abs(a - b))) if a.size else 0.0


@pytest.mark.parametrize('name', CLIP_NAMES)
def test_clip_fixture_end_to_end(engine, name):
	"""Every stage output of the reference's smart_vid_crop on a synthetic clip."""
	from retargetvid_b200 import smartVidCrop as svc
	vd, over, ratios, fx = load_clip_fixture(name)
	CP = svc.sc_init_crop_params()
	CP.update(over)
	res = engine.run([vd], CP, ratios, detail=True, want_filtered=True, np_int=fixture_np_int(fx))[0]
	assert res.status == 0
	if CP['focus_stability']:
		# line-sampling means (smartVidCrop.py:1395-1455) and the centres before the freeze
		assert _nan_close(res.jumps, fx['jumps'], 0) <= 1e-9
		assert _nan_close(res.dxnf, fx['dxnf'], 0) <= 1e-9
	# integer stages: bit-exact
	filt = np.transpose(res.filtered, (1, 2, 0))
	assert np.array_equal(filt, fx['smaps_filtered']), 'filtered maps differ in %d maps' % int(
		(filt != fx['smaps_filtered']).any(axis=(0, 1)).sum())
	for k, r in enumerate(ratios):
		tag = r.replace(':', '-')
		assert list(res.dims[k]) == list(fx['dims_' + tag])
	# float stages: stated tolerances (process pixels)
	assert _nan_close(res.dx, fx['dx'], 0) <= 1e-9
	assert _nan_close(res.dy, fx['dy'], 0) <= 1e-9
	assert _nan_close(res.series[0], fx['dxi'], 0) <= 1e-9
	assert _nan_close(res.series[1], fx['dyi'], 0) <= 1e-9
	assert _nan_close(res.series[2], fx['dxl'], 0) <= 1e-8
	assert _nan_close(res.series[3], fx['dyl'], 0) <= 1e-8
	max_cl = int(max(s[1] - s[0] + 1 for s in vd['segmentation']))
	tol = max(loess_tolerance(max_cl), 1e-8)
	assert _nan_close(res.series[4], fx['dxs_pre'], 0) <= tol
	assert _nan_close(res.series[5], fx['dys_pre'], 0) <= tol
	# boxes: bit-exact, except frames whose reference centre sits within tol of an integer
	# boundary before the int() truncation (smartVidCrop.py:998-999) -- enumerated, must be explained
	scale_w = vd['w_process'] / vd['w_orig']
	scale_h = vd['h_process'] / vd['h_orig']
	for k, r in enumerate(ratios):
		want = fx['bbs_' + r.replace(':', '-')]
		got = res.boxes[k]
		diff = np.nonzero((got != want).any(axis=1))[0]
		for f in diff:
			vx = fx['dxs_pre'][f] / scale_w
			vy = fx['dys_pre'][f] / scale_h
			near = min(abs(vx - round(vx)), abs(vy - round(vy)))
			assert near <= tol / min(scale_w, scale_h), (name, r, int(f), got[f], want[f])
			assert np.max(np.abs(got[f] - want[f])) <= 1
		assert len(diff) <= max(1, len(want) // 100), (name, r, len(diff))


def test_smooth_series_vs_pyloess_fixture(engine):
	"""loess_handler -> pyloess.Loess.estimate (pyloess.py:61-95), degree 2, as the hot path calls it."""
	from retargetvid_b200 import _cabi
	z = np.load(os.path.join(GOLDEN, 'loess_fixture.npz'))
	for cl in (12, 60, 300, 1283):
		y = z['y_%d' % cl]
		want = z['est_%d' % cl]
		w = int(z['w_%d' % cl])
		p = _cabi.rvb_params()
		engine.ctx.lib.rvb_params_default(p, 0)
		p.lp_filt = 0
		# window = min(int(fr * w_secs), cl - 2), made odd: pick fr so that it equals the fixture's
		p.loess_w_secs = 1.0
		fr = float(w) if w % 2 == 1 else float(w + 1)
		lp, sm = engine.ctx.debug_smooth_series(p, y, fr)
		assert np.array_equal(lp, y)
		assert np.max(np.abs(sm - want)) <= loess_tolerance(cl), (cl, np.max(np.abs(sm - want)))
	# narrow / wide windows and degree 1 (other frame rates, loess_w_secs, loess_degree)
	for cl, w, deg in z['extra']:
		key = '%d_%d_%d' % (cl, w, deg)
		p = _cabi.rvb_params()
		engine.ctx.lib.rvb_params_default(p, 0)
		p.lp_filt = 0
		p.loess_w_secs = 1.0
		p.loess_degree = int(deg)
		lp, sm = engine.ctx.debug_smooth_series(p, z['y_' + key], float(w))
		assert np.max(np.abs(sm - z['est_' + key])) <= loess_tolerance(int(cl), int(w)), (key, np.max(np.abs(sm - z['est_' + key])))


def test_batch_equals_single(engine):
	"""Batching many clips and ratios in one launch gives exactly the per-clip results."""
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vds = [synth.make_clip(900 + i, fc=60 + 17 * i, shot_starts=[30] if i % 2 else []) for i in range(5)]
	CP = svc.sc_init_crop_params()
	ratios = ['1:3', '3:1', '9:16', '4:5']
	batch = engine.run(vds, CP, ratios)
	for i, vd in enumerate(vds):
		one = engine.run([vd], CP, ratios)[0]
		assert np.array_equal(one.boxes, batch[i].boxes)
		assert np.array_equal(one.dx, batch[i].dx)
		assert np.array_equal(one.series, batch[i].series)


def test_drop_in_smart_vid_crop(engine, tmp_path):
	"""The reference-facing call: vid_data pickle in temp_path -> (VD, smart_crop_results) and the text files."""
	import pickle
	from retargetvid_b200 import smartVidCrop as svc
	vd, over, ratios, fx = load_clip_fixture('c1_default')
	with open(os.path.join(tmp_path, 'clipA.pkl'), 'wb') as fp:
		pickle.dump(vd, fp)
	CP = svc.sc_init_crop_params()
	CP['out_ratio'] = '1:3'
	VD, res = svc.smart_vid_crop(os.path.join('somewhere', 'clipA.mp4'), CP, temp_path=str(tmp_path), save_vid=False)
	assert res['result'] == 'smart cropped'
	assert VD['bbs'] == [[int(v) for v in bb] for bb in fx['bbs_1-3']]
	assert res['info'] == ' (360x640)->(140x250)->(360x120)->(360x120)\n'
	assert isinstance(VD['bbs'][0][0], int) and 't__clustering' in res and 't_total' in res
	svc.write_result_files(str(tmp_path), 'clipA_1-3', VD, res)
	with open(os.path.join(tmp_path, 'clipA_1-3.txt')) as fp:
		lines = fp.read().splitlines()
	assert lines[0] == '%d,%d,%d,%d' % tuple(fx['bbs_1-3'][0]) and len(lines) == vd['fc']


def test_error_behaviour(engine):
	from retargetvid_b200 import _cabi
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	CP = svc.sc_init_crop_params()
	with pytest.raises(TypeError):  # the reference raises TypeError (float(None)) when every map is empty
		svc.smart_vid_crop('x.mp4', CP, save_vid=False, vid_data=synth.make_clip(1, fc=60, kind='empty'))
	CP2 = dict(CP)
	CP2['t_threshold'] = 0      # every pixel salient: far beyond RVB_MAX_POINTS -> loud capacity error
	with pytest.raises(_cabi.RvbError) as ei:
		svc.smart_vid_crop('x.mp4', CP2, save_vid=False, vid_data=synth.make_clip(2, fc=30))
	assert ei.value.code == _cabi.RVB_ERR_CAPACITY
	CP3 = svc.sc_init_crop_params(use_best_settings=True)
	CP3['resize_type'] = 5     # not one of the reference's three interpolation modes (smartVidCrop.py:1078-1084): must be loud
	with pytest.raises(NotImplementedError):
		svc.smart_vid_crop('x.mp4', CP3, save_vid=False, vid_data=synth.make_clip(3, fc=30))


def _run_raw(engine, vds, CP, ratios, maps_kind, maps_arrays):
	"""Direct C-ABI call with an explicit map layout (host memory)."""
	import ctypes as C
	from retargetvid_b200 import _cabi
	nc = len(vds)
	clips = (_cabi.rvb_clip * nc)()
	shots, tinds = [], []
	mo = fo = so = 0
	ptrs = (C.c_void_p * nc)()
	for i, vd in enumerate(vds):
		c = clips[i]
		c.n_maps, c.n_frames, c.n_shots = vd['fc_sel'], vd['fc'], len(vd['segmentation'])
		c.h_orig, c.w_orig, c.fr = vd['h_orig'], vd['w_orig'], vd['fr']
		c.map_offset, c.frame_offset, c.shot_offset = mo, fo, so
		shots.append(np.concatenate([vd['segmentation'], vd['segmentation_sel']], axis=1))
		tinds.append(np.asarray(vd['true_inds'], dtype=np.int32))
		ptrs[i] = maps_arrays[i].ctypes.data
		mo += c.n_maps
		fo += c.n_frames
		so += c.n_shots
	shots = np.ascontiguousarray(np.concatenate(shots), dtype=np.int32)
	tinds = np.ascontiguousarray(np.concatenate(tinds), dtype=np.int32)
	b = _cabi.rvb_batch()
	b.n_clips, b.h_process, b.w_process, b.n_ratios = nc, vds[0]['h_process'], vds[0]['w_process'], len(ratios)
	b.maps_kind, b.mem_space = maps_kind, _cabi.RVB_MEM_HOST
	b.row_stride = maps_arrays[0].shape[-1] if maps_kind == _cabi.RVB_MAPS_U8_NHW else 0
	for r, s in enumerate(ratios):
		a, bb = s.split(':')
		b.ratio_w[r], b.ratio_h[r] = float(a), float(bb)
	b.clips = clips
	b.shots = shots.ctypes.data
	b.true_inds = tinds.ctypes.data
	b.clip_maps = ptrs
	boxes = np.empty((len(ratios), fo, 4), dtype=np.int32)
	filt = np.empty((mo, b.h_process, b.w_process), dtype=np.uint8)
	b.boxes = boxes.ctypes.data
	b.filtered_maps = filt.ctypes.data
	b.row_stride_out = b.w_process
	engine.ctx.crop_track_batch(_cabi.params_from_crop_params(CP), b)
	return boxes, filt


def test_map_layouts_agree(engine):
	"""reference layout [H,W,N], device-native [N,H,256] and float32 log-saliency entries."""
	from retargetvid_b200 import _cabi
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vds = [synth.make_clip(700 + i, fc=90, shot_starts=[45] if i else [], keep_logp=True) for i in range(3)]
	CP = svc.sc_init_crop_params()
	ratios = ['1:3', '9:16']
	hwn = [np.ascontiguousarray(vd['smaps']) for vd in vds]
	b0, f0 = _run_raw(engine, vds, CP, ratios, _cabi.RVB_MAPS_U8_HWN, hwn)
	nhw = []
	for vd in vds:
		a = np.zeros((vd['fc_sel'], 140, 256), dtype=np.uint8)
		a[:, :, :250] = np.transpose(vd['smaps'], (2, 0, 1))
		a[:, :, 250:] = 77   # padding bytes are the caller's garbage and must be ignored
		nhw.append(a)
	b1, f1 = _run_raw(engine, vds, CP, ratios, _cabi.RVB_MAPS_U8_NHW, nhw)
	assert np.array_equal(b0, b1) and np.array_equal(f0, f1)


def test_float32_entry(engine):
	"""a1 (unisal/train.py:1270-1274) fused in front of the path: the uint8 quantisation may differ from
	numpy's by one grey level only where v*255 is within float32 rounding of an integer (CUDA rounds exp
	correctly, numpy/torch do not); such pixels are enumerated and must be rare."""
	from oracle import sc_oracle
	from retargetvid_b200 import _cabi
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vd = synth.make_clip(811, fc=60, keep_logp=True)
	logp = np.ascontiguousarray(vd['_logp'], dtype=np.float32)
	CP = svc.sc_init_crop_params()
	CP['clust_filt'] = False
	CP['t_threshold'] = 0       # filtered map == quantised map
	_, q = _run_raw(engine, [vd], CP, ['1:3'], _cabi.RVB_MAPS_F32_NHW, [logp])
	want = np.stack([sc_oracle.normalise_u8(logp[i]) for i in range(logp.shape[0])])
	diff = q.astype(int) - want.astype(int)
	assert np.max(np.abs(diff)) <= 1
	flips = np.argwhere(diff != 0)
	assert len(flips) <= q.size // 20000, len(flips)
	for (i, y, x) in flips:
		e = np.exp(logp[i].astype(np.float64))
		v = e[y, x] / e.max() * 255.0
		assert abs(v - round(v)) < 1e-4, (i, y, x, v)
	# and the full path from float32 maps equals the path from the quantised maps it produced
	CP = svc.sc_init_crop_params()
	b_f32, _ = _run_raw(engine, [vd], CP, ['1:3'], _cabi.RVB_MAPS_F32_NHW, [logp])
	vd_q = dict(vd)
	vd_q['smaps'] = np.ascontiguousarray(np.transpose(q, (1, 2, 0)))
	b_u8, _ = _run_raw(engine, [vd_q], CP, ['1:3'], _cabi.RVB_MAPS_U8_HWN, [vd_q['smaps']])
	assert np.array_equal(b_f32, b_u8)


def test_full_size_properties(engine):
	"""BASELINE configs[2] size (200 clips x 2 ratios): size-independent properties of the output."""
	import bench
	from retargetvid_b200 import smartVidCrop as svc
	vds = bench.make_workload(200, 0)
	CP = svc.sc_init_crop_params()
	res = engine.run(vds, CP, ['1:3', '3:1'], detail=True)
	rng = np.random.default_rng(0)
	for i, (vd, r) in enumerate(zip(vds, res)):
		assert r.status == 0
		b13, b31 = r.boxes[0], r.boxes[1]
		# crop size and containment (sc_compute_bb's clamps): 120x360 and 640x213 inside 640x360
		assert np.all(b13[:, 2] - b13[:, 0] == 120) and np.all(b13[:, 1] == 0) and np.all(b13[:, 3] == 360)
		assert np.all(b31[:, 3] - b31[:, 1] == 213) and np.all(b31[:, 0] == 0) and np.all(b31[:, 2] == 640)
		assert b13[:, 0].min() >= 0 and b13[:, 2].max() <= 640 and b31[:, 1].min() >= 0 and b31[:, 3].max() <= 360
		# both ratios come from the same centre track
		cx = (r.series[4] / (250.0 / 640.0)).astype(np.int64)
		assert np.array_equal(np.clip(cx - 60, 0, 520), b13[:, 0])
		# centres lie inside the map, kept pixels never exceed the thresholded pixels by more than closing can add
		assert np.all((r.dx >= 0) & (r.dx <= 249) & (r.dy >= 0) & (r.dy <= 139))
	# a random subset recomputed alone gives bit-identical results (batch independence, no cross-clip state)
	for i in rng.choice(len(vds), 6, replace=False):
		one = engine.run([vds[i]], CP, ['1:3', '3:1'], detail=True)[0]
		assert np.array_equal(one.boxes, res[i].boxes) and np.array_equal(one.series, res[i].series)
	# idempotence of the filter: feeding the filtered maps back leaves them unchanged when no blend applies
	vd = dict(vds[3])
	first = engine.run([vd], CP, ['1:3'], detail=True, want_filtered=True)[0]
	vd2 = dict(vd)
	vd2['smaps'] = np.ascontiguousarray(np.transpose(first.filtered, (1, 2, 0)))
	again = engine.run([vd2], CP, ['1:3'], detail=True, want_filtered=True)[0]
	seg_cuts = set(int(s[0]) for s in vd['segmentation_sel'])
	free = [k for k in range(3, vd['fc_sel']) if not ({k - 2, k - 1, k} & seg_cuts)]
	kept = (first.filtered[free] > 0).sum()
	assert (again.filtered[free] > 0).sum() <= kept * 1.05 + 5


def test_iou_properties_full_size(engine):
	"""config 2 size: IoU(a, a) == 1 for every frame, symmetry, and a checksum of the per-video sums."""
	from retargetvid_b200 import retargetvid_eval as rev
	method, annots, frame_counts, _ = _eval_fixture()
	ann = [a['3-1'] for a in annots]
	self_iou, _, _ = rev.evaluate_arrays(engine.ctx, annots[0]['3-1'], [annots[0]['3-1']], frame_counts)
	assert all(v[0] == 1.0 for v in self_iou)
	ab, _, _ = rev.evaluate_arrays(engine.ctx, annots[1]['3-1'], [annots[2]['3-1']], frame_counts)
	ba, _, _ = rev.evaluate_arrays(engine.ctx, annots[2]['3-1'], [annots[1]['3-1']], frame_counts)
	assert ab == ba


def test_coverage_score_crop_window_and_padding_fallback(engine):
	"""a7 with the Appendix B-1 window (config 4: 1080p, multi-shot, 9:16, padding fallback enabled):
	scores equal the oracle's restatement of smartVidCrop.py:1310-1331 bit for bit; the reference window
	reproduces its 0.0; a low score makes smart_vid_crop return result='padded' without boxes."""
	from oracle import sc_oracle
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vd = synth.make_clip(4001, fc=150, w_orig=1920, h_orig=1080, shot_starts=[70, 76])
	CP = svc.sc_init_crop_params()
	CP['exit_on_low_cvrg'] = True
	ratios = ['9:16', '3:1', '16:9']
	res = engine.run([vd], CP, ratios, detail=True, want_filtered=True, cvrg_window='crop')[0]
	filt = np.ascontiguousarray(np.transpose(res.filtered, (1, 2, 0)))
	for k, r in enumerate(ratios):
		mode, wf, hf = sc_oracle.sc_calc_dest_size(vd['w_orig'], vd['h_orig'], r)
		want, per_map = sc_oracle.sc_compute_cvrg_score(filt, mode, vd['w_process'], vd['h_process'], 'crop',
														wf, hf, vd['w_orig'], vd['h_orig'])
		assert float(res.cvrg_scores[k]) == want, (r, float(res.cvrg_scores[k]), want)
	ref0 = engine.run([vd], CP, ratios, detail=True, cvrg_window='reference')[0]
	assert list(ref0.cvrg_scores) == [0.0, 0.0, 0.0]
	# the whole oracle pipeline on this clip (shots of 6 and 74 frames, LOESS on)
	CP2 = dict(CP)
	CP2['out_ratio'] = '9:16'
	want = sc_oracle.smart_vid_crop_oracle(vd, CP2, cvrg_window='crop')
	assert np.array_equal(filt, want['smaps_filtered'])
	assert np.array_equal(res.boxes[0], np.array(want['bbs'], dtype=np.int32))
	# padding fallback: a coverage score can never reach t_cvrg=1.01
	CP3 = dict(CP2)
	CP3['t_cvrg'] = 1.01
	VD, info = svc.smart_vid_crop('c4.mp4', CP3, save_vid=False, vid_data=dict(vd), cvrg_window='crop')
	assert info['result'] == 'padded' and 'bbs' not in VD and info['coverage_score'] == want['mean_cvrg_score']


def test_focus_stability_sampling_paths(engine):
	"""focus stability with integer centres (com_km=False) so that vertical / horizontal moves occur, and with
	the pinned-numpy behaviour (np_int) for diagonal moves: jumps, frozen centres and boxes equal the oracle."""
	from oracle import sc_oracle
	from retargetvid_b200 import _cabi
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vd = synth.make_clip(5150, fc=200, shot_starts=[90])
	for np_int in (False, True):
		CP = svc.sc_init_crop_params()
		CP.update(dict(focus_stability=True, com_km=False, min_d_jump=1, foces_stab_t=200, foces_stab_s=3.0, out_ratio='1:3'))
		want = sc_oracle.smart_vid_crop_oracle(vd, CP, np_int=np_int)
		assert _cabi.params_from_crop_params(CP, np_int=np_int).np_int_compat == int(np_int)
		res = engine.run([vd], CP, ['1:3'], detail=True, np_int=np_int)[0]
		assert np.allclose(res.jumps, np.array(want['jumps'], dtype=np.float64), rtol=0, atol=1e-9), np_int
		assert np.array_equal(res.dx, np.array(want['dx'], dtype=np.float64))
		assert np.array_equal(res.dxnf, np.array(want['dxnf'], dtype=np.float64))
		assert np.array_equal(res.boxes[0], np.array(want['bbs'], dtype=np.int32))
		assert any(j != 255 for j in want['jumps'])


@pytest.mark.parametrize('w_orig,h_orig', [(480, 360), (360, 640), (1000, 250)])
def test_other_process_sizes(engine, w_orig, h_orig):
	"""4:3 (187x250 maps), portrait (250x140) and very wide (62x250) inputs against the oracle."""
	from oracle import sc_oracle
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vd = synth.make_clip(9000 + w_orig, fc=80, w_orig=w_orig, h_orig=h_orig, shot_starts=[37])
	CP = svc.sc_init_crop_params()
	CP['out_ratio'] = '1:1'
	res = engine.run([vd], CP, ['1:1', '9:16'], detail=True, want_filtered=True)[0]
	want = sc_oracle.smart_vid_crop_oracle(vd, CP)
	assert np.array_equal(np.transpose(res.filtered, (1, 2, 0)), want['smaps_filtered'])
	assert np.array_equal(res.boxes[0], np.array(want['bbs'], dtype=np.int32))
	assert np.max(np.abs(res.series[4] - np.array(want['dxs']))) <= 1e-8


def test_ragged_batch_and_tiny_clips(engine):
	"""clips of 1, 2, 3 frames, a clip whose every shot is shorter than the filtfilt pad, mixed process sizes in one call."""
	from oracle import sc_oracle
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vds = [synth.make_clip(9100, fc=1), synth.make_clip(9101, fc=2), synth.make_clip(9102, fc=3),
		synth.make_clip(9103, fc=40, shot_starts=[9, 17, 26, 31]), synth.make_clip(9104, fc=30, w_orig=480, h_orig=360),
		synth.make_clip(9105, fc=25, kind='single_pixel')]
	CP = svc.sc_init_crop_params()
	CP['out_ratio'] = '3:1'
	res = engine.run(vds, CP, ['3:1'], detail=True)
	for vd, r in zip(vds, res):
		want = sc_oracle.smart_vid_crop_oracle(vd, CP)
		assert r.status == 0
		assert np.array_equal(r.boxes[0], np.array(want['bbs'], dtype=np.int32))


def test_streaming_kernel_clust_filt_off(engine):
	"""clust_filt=False (smartVidCrop.py:2354): the one-warp-per-map streaming kernel equals the fused kernel
	(taken when filtered maps are requested) and the oracle, for the centroid and the argmax variant."""
	from oracle import sc_oracle
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vds = [synth.make_clip(9300 + i, fc=70 + 9 * i, shot_starts=[33] if i % 2 else []) for i in range(4)]
	vds.append(synth.make_clip(9310, fc=40, kind='few_points'))
	for com_km in (True, False):
		CP = svc.sc_init_crop_params()
		CP.update(dict(clust_filt=False, com_km=com_km, out_ratio='4:5'))
		fast = engine.run(vds, CP, ['4:5'], detail=True)
		slow = engine.run(vds, CP, ['4:5'], detail=True, want_filtered=True)
		for vd, a, b in zip(vds, fast, slow):
			assert np.array_equal(a.boxes, b.boxes) and np.array_equal(a.dx, b.dx) and np.array_equal(a.map_scores, b.map_scores)
			assert np.array_equal(a.map_info[:, 0], b.map_info[:, 0])
			want = sc_oracle.smart_vid_crop_oracle(vd, CP)
			assert np.array_equal(a.boxes[0], np.array(want['bbs'], dtype=np.int32))
			assert np.allclose(a.map_scores, want['mean_sal_scores'], rtol=0, atol=1e-12)
