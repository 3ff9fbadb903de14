"""Generates the committed golden fixtures by running the UNMODIFIED reference
in the build container (needs /root/reference; never runs on the GPU box).

    python tests/golden/make_golden.py [--only eval|clips|loess|hdbscan]

Outputs (all under tests/golden/):
  eval_fixture.npz     the 6 annotator zips + results/smartvidcrop as int16 arrays,
                       and the numbers the unmodified retargetvid_eval.py prints
  clip_*.npz           synthetic vid_data inputs + every stage output of the
                       reference smart_vid_crop (hdbscan -> sklearn stand-in)
  loess_fixture.npz    pyloess.Loess.estimate outputs (degree 1 and 2)
  hdbscan_fixture.npz  sklearn.cluster.HDBSCAN labels on point sets from maps

numpy's argsort is forced onto its portable (non-SIMD) introsort with
NPY_DISABLE_CPU_FEATURES, because the clustering result depends on how the
unstable sort permutes tied edge weights (DESIGN.md "tie rule"); that is the
only sort numpy had in the versions the reference pins (README.md:85-92).
"""
import contextlib
import io
import os
import shutil
import subprocess
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SCALAR_SORT = ("AVX512F AVX512CD AVX512_KNL AVX512_KNM AVX512_SKX AVX512_CLX "
			"AVX512_CNL AVX512_ICL AVX512_SPR AVX2")

if os.environ.get('NPY_DISABLE_CPU_FEATURES') != SCALAR_SORT:
	env = dict(os.environ)
	env['NPY_DISABLE_CPU_FEATURES'] = SCALAR_SORT
	sys.exit(subprocess.call([sys.executable] + sys.argv, env=env))

import numpy as np  # noqa: E402

sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402
from retargetvid_b200 import synth  # noqa: E402

VID_INDS = list(range(1, 101)) + list(range(601, 701))
ARS = ['1-3', '3-1']


def _read_boxes(text):
	return np.array([[int(v) for v in l.split(',')] for l in text.splitlines()], dtype=np.int16).reshape(-1, 4)


def make_eval_fixture():
	ref = ref_harness.REF_ROOT
	out = {}
	offsets = {}
	for u in range(1, 7):
		z = zipfile.ZipFile(os.path.join(ref, 'annotations', 'annotator_%d.zip' % u))
		for ar in ARS:
			chunks = []
			lens = []
			for v in VID_INDS:
				b = _read_boxes(z.read('annotator_%d/%03d_%s.txt' % (u, v, ar)).decode())
				chunks.append(b)
				lens.append(len(b))
			out['annot_%d_%s' % (u, ar)] = np.concatenate(chunks)
			out['annot_len_%d_%s' % (u, ar)] = np.array(lens, dtype=np.int32)
	for ar in ARS:
		chunks = []
		lens = []
		for v in VID_INDS:
			with open(os.path.join(ref, 'results', 'smartvidcrop', '%03d_%s.txt' % (v, ar))) as fp:
				b = _read_boxes(fp.read())
			chunks.append(b)
			lens.append(len(b))
		out['method_%s' % ar] = np.concatenate(chunks)
		out['method_len_%s' % ar] = np.array(lens, dtype=np.int32)
	out['vid_inds'] = np.array(VID_INDS, dtype=np.int32)

	# run the unmodified evaluator on a writable copy (it unzips next to itself)
	with tempfile.TemporaryDirectory() as td:
		shutil.copy(os.path.join(ref, 'retargetvid_eval.py'), td)
		shutil.copytree(os.path.join(ref, 'annotations'), os.path.join(td, 'annotations'))
		shutil.copytree(os.path.join(ref, 'results'), os.path.join(td, 'results'))
		r = subprocess.run([sys.executable, 'retargetvid_eval.py'], cwd=td, capture_output=True, text=True)
		with open(os.path.join(td, 'eval_current.txt')) as fp:
			csv = fp.read()
		# full precision per-annotator means by importing nothing: re-run its loop
		# semantics is the oracle's job; here we keep what the script printed.
	out['eval_csv'] = np.array(csv)
	out['eval_stdout_tail'] = np.array('\n'.join(r.stdout.splitlines()[-4:]))
	np.savez_compressed(os.path.join(HERE, 'eval_fixture.npz'), **out)
	print('eval fixture written;', csv.splitlines()[-1])


def _capture(vd, CP):
	"""Reference smart_vid_crop + a copy of dxs/dys taken just before
	sc_compute_bb truncates them in place (smartVidCrop.py:995-999)."""
	ref = ref_harness.load_reference()
	pre = {}
	orig_bb = ref.sc_compute_bb

	def spy(vid_data, crop_params, verbose=False):
		pre['dxs'] = [float(v) for v in vid_data['dxs']]
		pre['dys'] = [float(v) for v in vid_data['dys']]
		return orig_bb(vid_data, crop_params, verbose=verbose)
	ref.sc_compute_bb = spy
	try:
		VD, res = ref_harness.run_reference(vd, CP)
	finally:
		ref.sc_compute_bb = orig_bb
	return VD, res, pre


def _f(a):
	return np.array([np.nan if v is None else float(v) for v in a], dtype=np.float64)


CLIPS = {
	# name: (make_clip kwargs, crop-param overrides, out ratios)
	'c1_default': (dict(seed=1000, fc=300), {}, ['1:3', '3:1']),
	'multishot': (dict(seed=2001, fc=260, shot_starts=[90, 97, 110, 200]), {}, ['1:3', '9:16']),
	'fr25': (dict(seed=2002, fc=180, fr=25.0, shot_starts=[61]), {}, ['4:5']),
	'hd1080': (dict(seed=2003, fc=200, w_orig=1920, h_orig=1080, shot_starts=[120]), {}, ['9:16', '3:1']),
	'constant': (dict(seed=2004, fc=120, kind='constant'), {}, ['1:3']),
	'noise': (dict(seed=2005, fc=100, kind='noise', shot_starts=[50]), {}, ['1:3']),
	'few_points': (dict(seed=2006, fc=90, kind='few_points'), {}, ['3:1']),
	'sumsel_min5': (dict(seed=2007, fc=150, shot_starts=[70]),
					dict(select_sum=1, hdbscan_min=5, hdbscan_min_samples=3, t_threshold=90), ['1:3']),
	'noclose_nolp': (dict(seed=2008, fc=150), dict(op_close=False, lp_filt=0), ['1:3']),
	'savgol_argmax': (dict(seed=2009, fc=150, shot_starts=[80]),
					dict(loess_filt=0, com_km=False, lp_cutoff=1, lp_order=2), ['1:3']),
	'border': (dict(seed=2010, fc=120), dict(t_border=10), ['1:3', '3:1']),
	'best_settings': (dict(seed=2012, fc=240, shot_starts=[100]), 'BEST', ['1:3', '3:1']),
	'best_hd_fr25': (dict(seed=2013, fc=150, w_orig=1920, h_orig=1080, fr=25.0), 'BEST', ['9:16']),
	# parameters no other fixture moves: time shift, degree-1 LOESS with a 1 s window, a 3rd-order low-pass at 3 Hz
	'shift_deg1': (dict(seed=2014, fc=200, shot_starts=[90]),
					dict(shift_time=5, loess_degree=1, loess_w_secs=1, lp_order=3, lp_cutoff=3), ['1:3', '4:5']),
	# clustering on a 4x down-scaled copy taken with INTER_NEAREST (resize_type 3)
	'resize_nearest': (dict(seed=2015, fc=120, shot_starts=[50]),
					dict(t_threshold=90, hdbscan_min=5, hdbscan_min_samples=3, resize_factor=4, resize_type=3, select_sum=1), ['3:1']),
	# the ISM-2021 preset as it behaves on the numpy the reference pins (README.md:85-92, numpy < 1.24): np.int still
	# exists there, so get_points_on_line (smartVidCrop.py:1377,1384) samples diagonal moves instead of raising into
	# its own except.  'NPINT' runs the unmodified reference with the alias restored (np.int = int).
	'best_npint': (dict(seed=2017, fc=220, shot_starts=[120]), 'BEST_NPINT', ['1:3', '3:1']),
	# oracle-only for now (tests/helpers.py ORACLE_ONLY_NAMES): border detection on a 1080p multi-shot clip, and a 3 s
	# LOESS window with value_bias / t_threshold / lp_cutoff moved
	'border_hd_multishot': (dict(seed=2018, fc=260, w_orig=1920, h_orig=1080, shot_starts=[70, 150]), dict(t_border=12), ['3:1', '9:16']),
	'loess_w3_bias': (dict(seed=2019, fc=320, shot_starts=[200]),
					dict(loess_w_secs=3, value_bias=0.5, t_threshold=140, lp_cutoff=1.5, hdbscan_min=15), ['4:5']),
	# exactly 2x: cv2.resize turns INTER_LINEAR into INTER_AREA (resize.cpp, is_area_fast); 4:3 input -> 187 x 250 maps, odd height
	'resize_area2': (dict(seed=2020, fc=110, w_orig=480, h_orig=360, shot_starts=[60]),
					dict(t_threshold=90, hdbscan_min=5, hdbscan_min_samples=3, resize_factor=2, resize_type=1, select_sum=1), ['9:16']),
	# clustering on a 4x down-scaled copy taken with INTER_CUBIC (resize_type 2), dominant cluster by maximum
	'resize_cubic': (dict(seed=2021, fc=130, shot_starts=[40, 90]),
					dict(t_threshold=90, hdbscan_min=5, hdbscan_min_samples=3, resize_factor=4, resize_type=2), ['1:3', '4:5']),
	# combinations the random sweep (tests/test_gpu_at_size.py) singled out: clustering off with a resize factor set (the
	# reference then never enters sc_clustering_filt, so no down-/up-scaling round trip; only the centroid samples by the
	# factor), and integer (argmax) centres frozen by focus stability under Savitzky-Golay smoothing
	'noclust_resize': (dict(seed=2022, fc=140, shot_starts=[45, 100]),
					dict(clust_filt=False, resize_factor=4, resize_type=1, t_threshold=90), ['1:3', '4:5']),
	'argmax_focus': (dict(seed=2023, fc=160, shot_starts=[70]),
					dict(com_km=False, focus_stability=True, foces_stab_t=60, foces_stab_s=1.0, min_d_jump=5, loess_filt=0, lp_order=3, lp_cutoff=2.0), ['1:3', '3:1']),
	# a different sampling table: every 3rd frame gets a map, 24 fps
	'skip3_fr24': (dict(seed=2016, fc=150, fr=24.0, skip=3, shot_starts=[75]), {}, ['9:16']),
}


def _with_empties(vd, rng):
	"""Zero a few maps (start, middle run, end) so sc_handle_empty_centers runs."""
	n = vd['fc_sel']
	for i in [0, 1, n // 2, n // 2 + 1, n // 2 + 2, n - 1]:
		vd['smaps'][:, :, i] = 0
	return vd


def make_clip_fixtures():
	ref = ref_harness.load_reference()
	specs = dict(CLIPS)
	if specs.pop('__skip_empties__', 0) != None:
		specs['empties'] = (dict(seed=2011, fc=240, shot_starts=[100]), {}, ['1:3'])
	for name, (kw, over, ratios) in specs.items():
		vd = synth.make_clip(**kw)
		if name == 'border_hd_multishot':
			vd['smaps'][:9, :, :] = 0
			vd['smaps'][-14:, :, :] = 0
			vd['smaps'][:, :17, :] = 0
		if name == 'border':
			# blank borders: zero 12 rows top, 20 cols right in every map
			vd['smaps'][:12, :, :] = 0
			vd['smaps'][:, -20:, :] = 0
		if name == 'empties':
			vd = _with_empties(vd, None)
		out = {}
		out['kw'] = np.array(repr(kw))
		out['over'] = np.array(repr('BEST' if over == 'BEST_NPINT' else over))
		out['ratios'] = np.array(ratios)
		for k in ('smaps', 'segmentation', 'segmentation_sel'):
			out['in_' + k] = np.asarray(vd[k])
		out['in_true_inds'] = np.array(vd['true_inds'], dtype=np.int32)
		out['in_inds_to_orig'] = np.array(vd['inds_to_orig'], dtype=np.int32)
		out['in_scalars'] = np.array([vd['fr'], vd['fc'], vd['fc_sel'], vd['h_orig'], vd['w_orig'],
									vd['h_process'], vd['w_process']], dtype=np.float64)
		for r in ratios:
			best = isinstance(over, str) and over.startswith('BEST')
			npint = isinstance(over, str) and over.endswith('NPINT')
			CP = ref.sc_init_crop_params(use_best_settings=best)
			if not best:
				CP.update(over)
			CP['out_ratio'] = r
			if npint:
				np.int = int      # the alias numpy < 1.24 had; removed again below
			try:
				VD, res, pre = _capture(vd, CP)
			finally:
				if npint:
					del np.int
			out['np_int'] = np.array(1 if npint else 0)
			tag = r.replace(':', '-')
			out['bbs_' + tag] = np.array(VD['bbs'], dtype=np.int32)
			out['dims_' + tag] = np.array([VD['conversion_mode'], VD['w_final'], VD['h_final'],
										VD['fbb_w'], VD['fbb_h'], VD['border_t'], VD['border_b'],
										VD['border_l'], VD['border_r']], dtype=np.int32)
			if r == ratios[0]:
				out['smaps_filtered'] = np.asarray(VD['smaps'])
				out['dxnf'] = _f(VD['dxnf'])
				out['jumps'] = _f(VD['jumps'])
				out['dx'] = _f(VD['dx'])
				out['dy'] = _f(VD['dy'])
				out['dxi'] = _f(VD['dxi'])
				out['dyi'] = _f(VD['dyi'])
				out['dxl'] = _f(VD['dxl'])
				out['dyl'] = _f(VD['dyl'])
				out['dxs_pre'] = _f(pre['dxs'])
				out['dys_pre'] = _f(pre['dys'])
		np.savez_compressed(os.path.join(HERE, 'clip_%s.npz' % name), **out)
		print('clip fixture', name, 'N=%d F=%d' % (vd['fc_sel'], vd['fc']), 'first box', out['bbs_' + ratios[0].replace(':', '-')][0])


def make_loess_fixture():
	pyloess = ref_harness.load_pyloess()
	xx = np.array([0.5578196, 2.0217271, 2.5773252, 3.4140288, 4.3014084, 4.7448394, 5.1073781,
				6.5411662, 6.7216176, 7.2600583, 8.1335874, 9.1224379, 11.9296663, 12.3797674,
				13.2728619, 4.2767453, 15.3731026, 15.6476637, 18.5605355, 18.5866354, 18.7572812])
	yy = np.array([18.63654, 103.49646, 150.35391, 190.51031, 208.70115, 213.71135, 228.49353,
				233.55387, 234.55054, 223.89225, 227.68339, 223.91982, 168.01999, 164.95750,
				152.61107, 160.78742, 168.55567, 152.42658, 221.70702, 222.69040, 243.18828])
	lo = pyloess.Loess(xx, yy)
	out = dict(xx=xx, yy=yy)
	out['main_deg1_w7'] = np.array([lo.estimate(x, window=7, use_matrix=False, degree=1) for x in xx])
	out['main_deg2_w7'] = np.array([lo.estimate(x, window=7, use_matrix=False, degree=2) for x in xx])
	# the way the hot path calls it (smartVidCrop.py:1637-1638): integer frame
	# axis, degree 2, odd window
	rng = np.random.default_rng(7)
	for cl, w in ((12, 9), (60, 58 - 1), (300, 59), (1283, 59)):
		t = np.array(list(range(cl)))
		y = 120 + 40 * np.sin(t / 37.0) + rng.normal(0, 2.0, cl)
		lo = pyloess.Loess(t, y)
		out['y_%d' % cl] = y
		out['w_%d' % cl] = np.array(w)
		out['est_%d' % cl] = np.array([lo.estimate(j, window=w, use_matrix=False, degree=2) for j in range(cl)])
	# narrow and wide windows, both degrees (what other frame rates / loess_w_secs / loess_degree give): key suffix cl_w_deg
	extra = ((67, 11, 2), (130, 11, 2), (150, 23, 1), (400, 35, 2), (600, 179, 2), (200, 119, 1), (90, 87, 2))
	for cl, w, deg in extra:
		t = np.array(list(range(cl)))
		y = 120 + 40 * np.sin(t / 37.0) + 15 * np.sin(t / 5.0) + rng.normal(0, 2.0, cl)
		lo = pyloess.Loess(t, y)
		out['y_%d_%d_%d' % (cl, w, deg)] = y
		out['est_%d_%d_%d' % (cl, w, deg)] = np.array([lo.estimate(j, window=w, use_matrix=False, degree=deg) for j in range(cl)])
	out['extra'] = np.array(extra)
	np.savez_compressed(os.path.join(HERE, 'loess_fixture.npz'), **out)
	print('loess fixture written')


def make_hdbscan_fixture():
	"""Point sets from thresholded synthetic maps + the labels the stand-in
	library returns (what smartVidCrop.py:1099 would receive)."""
	import warnings
	warnings.filterwarnings('ignore')
	out = {}
	idx = 0
	# (the first six rows are the round-1 fixture; the rest covers the (min_cluster_size, min_samples) combinations the
	# random parameter sweep draws, on every kind of map, 2 maps each)
	configs = [(11, 'blobs', 26, None, 120, 2), (12, 'blobs', 26, None, 120, 2), (13, 'blobs', 5, 3, 90, 2),
			(14, 'noise', 26, None, 120, 2), (15, 'noise', 5, 3, 120, 2), (16, 'blobs', 10, None, 150, 2)]
	seed = 100
	for mcs in (5, 12, 26, 40, 80):
		for ms in (None, 3, 8, 15):
			for kind, thr in (('blobs', 90), ('noise', 120), ('few_points', 60), ('blobs', 150)):
				configs.append((seed, kind, mcs, ms, thr, 5))
				seed += 1
	for seed, kind, mcs, ms, thr, step in configs:
		vd = synth.make_clip(seed, fc=40, shot_starts=[20], kind=kind)
		maps = np.transpose(vd['smaps'], (2, 0, 1))
		for i in range(0, maps.shape[0], step):
			m = maps[i].copy()
			m[m < thr] = 0
			ys, xs = np.nonzero(m)
			if len(ys) <= mcs + 1:
				continue
			P = np.stack([ys, xs], 1)
			if (ms or mcs) + 1 > len(P):
				continue      # (the library raises: min_samples must be at most the number of points)
			lab = ref_harness._SklearnHDBSCANStandIn(min_cluster_size=mcs, min_samples=ms, metric='sqeuclidean',
													cluster_selection_method='eom', allow_single_cluster=True).fit_predict(P)
			out['P_%d' % idx] = P.astype(np.int16)
			out['L_%d' % idx] = lab.astype(np.int16)
			out['cfg_%d' % idx] = np.array([mcs, -1 if ms is None else ms], dtype=np.int32)
			idx += 1
	out['count'] = np.array(idx)
	np.savez_compressed(os.path.join(HERE, 'hdbscan_fixture.npz'), **out)
	print('hdbscan fixture written:', idx, 'point sets')


if __name__ == '__main__':
	only = sys.argv[sys.argv.index('--only') + 1] if '--only' in sys.argv else None
	if '--clip' in sys.argv:
		keep = sys.argv[sys.argv.index('--clip') + 1].split(',')
		for k in list(CLIPS.keys()):
			if k not in keep:
				del CLIPS[k]
		CLIPS['__skip_empties__'] = None
	if only in (None, 'eval'):
		make_eval_fixture()
	if only in (None, 'loess'):
		make_loess_fixture()
	if only in (None, 'hdbscan'):
		make_hdbscan_fixture()
	if only in (None, 'clips'):
		make_clip_fixtures()
