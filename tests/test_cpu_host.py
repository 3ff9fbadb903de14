"""CPU-only tests: the C ABI library loads and exports every declared symbol, host logic
(sharding, parameter mirror, synthetic data), oracle pieces against golden vectors."""
import os
import re
import statistics
import subprocess
import sys

import numpy as np
import pytest

from helpers import GOLDEN, loess_tolerance

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
	import __graft_entry__ as ge
	ge.build()
	from retargetvid_b200 import _cabi
	lib = _cabi.load_library()
	hdr = open(os.path.join(ROOT, 'include', 'retargetvid_b200.h')).read()
	declared = set(re.findall(r'\b(rvb_[a-z_0-9]+)\s*\(', hdr))
	assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
	for name in declared:
		assert getattr(lib, name) is not None
	assert b'retargetvid_b200' in lib.rvb_version()


def test_struct_layouts_match_header_sizes():
	"""ctypes mirrors of the ABI structs: sizes as the C compiler lays them out."""
	from retargetvid_b200 import _cabi
	src = '#include "%s"\n#include <stdio.h>\nint main(){printf("%%zu %%zu %%zu %%zu", sizeof(rvb_params), sizeof(rvb_clip), sizeof(rvb_batch), sizeof(rvb_iou_batch));return 0;}' % os.path.join(ROOT, 'include', 'retargetvid_b200.h')
	import tempfile
	with tempfile.TemporaryDirectory() as td:
		with open(os.path.join(td, 's.c'), 'w') as fp:
			fp.write(src)
		subprocess.check_call(['gcc', '-o', os.path.join(td, 's'), os.path.join(td, 's.c')])
		out = subprocess.check_output([os.path.join(td, 's')]).decode().split()
	import ctypes
	got = [ctypes.sizeof(_cabi.rvb_params), ctypes.sizeof(_cabi.rvb_clip), ctypes.sizeof(_cabi.rvb_batch), ctypes.sizeof(_cabi.rvb_iou_batch)]
	assert [int(v) for v in out] == got


def test_product_fails_loudly_without_gpu():
	"""No CPU fallback: without a CUDA device the context creation raises."""
	import torch
	if torch.cuda.is_available():
		pytest.skip('a GPU is present')
	from retargetvid_b200 import _cabi
	with pytest.raises(_cabi.RvbError):
		_cabi.Context(0)


def test_product_never_imports_the_oracle():
	for dirpath, _, files in os.walk(os.path.join(ROOT, 'retargetvid_b200')):
		for f in files:
			if f.endswith('.py'):
				src = open(os.path.join(dirpath, f)).read()
				assert 'oracle' not in re.sub(r'#.*', '', src).replace('"""', ''), f


def test_crop_params_mirror_the_reference_dict():
	from oracle import sc_oracle
	from retargetvid_b200 import smartVidCrop as svc
	for best in (False, True):
		a = svc.sc_init_crop_params(use_best_settings=best)
		b = sc_oracle.sc_init_crop_params(use_best_settings=best)
		assert a == b and list(a.keys()) == list(b.keys()) and len(a) == 31
	assert svc.smart_crop_version() == '1.4.0'


def test_lpt_sharding_balances_and_covers():
	from retargetvid_b200 import sharding
	rng = np.random.default_rng(0)
	costs = [int(v) for v in rng.integers(40, 220, 200)]
	for world in (1, 2, 4, 8):
		shards = sharding.lpt_shards(costs, world)
		assert sorted(sum(shards, [])) == list(range(200))
		loads = [sum(costs[i] for i in s) for s in shards]
		assert max(loads) - min(loads) <= max(costs)


def test_gloo_two_ranks_shard_and_gather(tmp_path):
	"""world_size 2 on CPU (gloo): each rank takes its LPT shard, results gathered on the host equal
	the single-process result (the data path has no collective; the gather is host plumbing)."""
	script = os.path.join(tmp_path, 'w.py')
	with open(script, 'w') as fp:
		fp.write('''
import os, sys
sys.path.insert(0, %r)
import numpy as np
import torch.distributed as dist
from retargetvid_b200 import sharding
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
costs = [50 + (7 * i) %% 90 for i in range(23)]
mine = sharding.my_shard(costs, rank, world)
local = [np.full((costs[i], 4), i, dtype=np.int32) for i in mine]     # stands for the boxes of video i
out = sharding.gather_results(local, mine, world, len(costs), dist)
assert all(o is not None and o.shape == (costs[i], 4) and int(o[0, 0]) == i for i, o in enumerate(out))
if rank == 0:
	print('GATHER_OK', len(out))
dist.destroy_process_group()
''' % ROOT)
	r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
						'--master-addr', '127.0.0.1', '--master-port', '29641', script], capture_output=True, text=True, timeout=300)
	assert 'GATHER_OK 23' in r.stdout, r.stdout + r.stderr


def test_synthetic_sampling_table_follows_the_reference_rule():
	from retargetvid_b200 import synth
	ti, i2o = synth.sampling_table(40, [13, 14], skip=6)
	assert ti == [0, 6, 12, 13, 14, 20, 26, 32, 38, 39]
	assert i2o[12] == 2 and i2o[13] == 3 and i2o[19] == 4 and i2o[39] == 9
	vd = synth.make_clip(7, fc=40, shot_starts=[13, 14])
	assert vd['segmentation'].tolist() == [[0, 12], [13, 13], [14, 39]]
	assert vd['segmentation_sel'].tolist() == [[0, 2], [3, 3], [4, 9]]
	assert vd['smaps'].shape == (140, 250, 10) and vd['smaps'].dtype == np.uint8


def test_oracle_hdbscan_matches_library_fixture():
	"""hdbscan_port against the labels the stand-in library produced: 157 point sets over 21 (min_cluster_size, min_samples)
	combinations and four kinds of maps."""
	from oracle import hdbscan_port
	z = np.load(os.path.join(GOLDEN, 'hdbscan_fixture.npz'))
	for i in range(int(z["count"])):
		P = z['P_%d' % i].astype(np.int64)
		mcs, ms = [int(v) for v in z['cfg_%d' % i]]
		got = hdbscan_port.fit_predict(P, mcs, None if ms < 0 else ms)
		assert np.array_equal(got, z['L_%d' % i].astype(np.int64)), i


def test_numpy_argsort_emulation():
	"""numpy_aquicksort == np.argsort on this box when numpy's SIMD sort is disabled (portable introsort)."""
	code = '''
import sys
sys.path.insert(0, %r)
import numpy as np
from oracle import hdbscan_port as hp
rng = np.random.default_rng(3)
for trial in range(120):
	n = int(rng.integers(1, 2500))
	v = rng.integers(1, 6, n) if trial %% 2 else np.where(rng.uniform(size=n) < 0.7, 9, rng.integers(9, 60, n))
	assert np.array_equal(np.argsort(v.astype(np.float64)), hp.numpy_aquicksort(v)), (trial, n)
print('SORT_OK')
''' % ROOT
	env = dict(os.environ)
	env['NPY_DISABLE_CPU_FEATURES'] = ('AVX512F AVX512CD AVX512_KNL AVX512_KNM AVX512_SKX AVX512_CLX AVX512_CNL '
									'AVX512_ICL AVX512_SPR AVX2')
	r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=300)
	assert 'SORT_OK' in r.stdout, r.stdout + r.stderr


def test_oracle_pyloess_and_eval_golden():
	from oracle import eval_oracle, sc_oracle
	z = np.load(os.path.join(GOLDEN, 'loess_fixture.npz'))
	xx, yy = z['xx'], z['yy']
	n_xx = (xx - xx.min()) / (xx.max() - xx.min())
	n_yy = (yy - yy.min()) / (yy.max() - yy.min())
	for deg, key in ((1, 'main_deg1_w7'), (2, 'main_deg2_w7')):
		got = [sc_oracle.loess_estimate(n_xx, n_yy, yy.min(), yy.max(), xx.min(), xx.max(), x, 7, deg) for x in xx]
		assert np.allclose(got, z[key], rtol=0, atol=1e-9)
	for cl in (12, 60, 300):
		got = sc_oracle.loess_handler(np.arange(cl), z['y_%d' % cl], 1, int(z['w_%d' % cl]), 2)
		# the reference's pinv of the uncentred normal equations is itself only reproducible to this level
		# across BLAS/SIMD code paths (the fixture was generated with numpy's SIMD dispatch off)
		assert np.max(np.abs(np.array(got) - z['est_%d' % cl])) <= loess_tolerance(cl)
	# narrow / wide windows and degree 1: the tolerance follows (shot length / window)^4
	for cl, w, deg in z['extra']:
		key = '%d_%d_%d' % (cl, w, deg)
		got = sc_oracle.loess_handler(np.arange(cl), z['y_' + key], 1, int(w), int(deg))
		assert np.max(np.abs(np.array(got) - z['est_' + key])) <= loess_tolerance(int(cl), int(w)), key
	# evaluator golden on the first 25 videos, 1-3, annotator 1
	e = np.load(os.path.join(GOLDEN, 'eval_fixture.npz'))
	lens = e['annot_len_1_1-3']
	offs = np.concatenate([[0], np.cumsum(lens)])
	ml = e['method_len_1-3']
	moffs = np.concatenate([[0], np.cumsum(ml)])
	vals = []
	for i in range(25):
		a = e['annot_1_1-3'][offs[i]:offs[i + 1]].astype(int).tolist()
		m = e['method_1-3'][moffs[i]:moffs[i + 1]].astype(int).tolist()
		vals.append(eval_oracle.video_iou(m, a, len(a)))
	assert all(0.0 <= v <= 1.0 for v in vals) and abs(statistics.mean(vals) - 0.5) < 0.25


def test_cv_resize_restatement_matches_opencv():
	"""oracle/cv_resize.py against cv2 itself (the library the reference calls, smartVidCrop.py:1080,1158,1184)."""
	cv2 = pytest.importorskip('cv2')
	from oracle import cv_resize
	rng = np.random.default_rng(0)
	for t in range(40):
		H, W = (140, 250) if t % 2 == 0 else (int(rng.integers(20, 200)), int(rng.integers(20, 256)))
		m = rng.integers(0, 256, (H, W)).astype(np.uint8)
		if t % 3 == 0:
			m[m < 150] = 0
		for f in (4.0, 3.0, 1.5):
			a = cv2.resize(m, None, fx=1.0 / f, fy=1.0 / f, interpolation=cv2.INTER_LINEAR)
			assert np.array_equal(a, cv_resize.resize_linear_u8(m, fx=1.0 / f, fy=1.0 / f))
			assert np.array_equal(cv2.resize(a, (W, H), interpolation=cv2.INTER_LINEAR), cv_resize.resize_linear_u8(a, dsize_wh=(W, H)))
			assert np.array_equal(cv2.resize(m, None, fx=1.0 / f, fy=1.0 / f, interpolation=cv2.INTER_NEAREST),
								cv_resize.resize_nearest_u8(m, 1.0 / f, 1.0 / f))
		# INTER_CUBIC (resize_type 2, smartVidCrop.py:1081-1082) at the factors whose sampling cv2 (IPP build) shares with
		# OpenCV's own code; the process size and odd sizes
		for f in (4.0, 2.0, 3.0, 5.0, 1.5, 2.5):
			if t % 2 == 0 or f in (4.0, 2.0):
				assert np.array_equal(cv2.resize(m, None, fx=1.0 / f, fy=1.0 / f, interpolation=cv2.INTER_CUBIC),
									cv_resize.resize_cubic_u8(m, 1.0 / f, 1.0 / f)), (H, W, f)
		# exactly 1/2: OpenCV takes the INTER_AREA fast path for INTER_LINEAR (odd sizes have partial last cells)
		half = cv2.resize(m, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_LINEAR)
		assert np.array_equal(half, cv_resize.resize_area2_u8(m))
		assert np.array_equal(cv2.resize(half, (W, H), interpolation=cv2.INTER_LINEAR), cv_resize.resize_linear_u8(half, dsize_wh=(W, H)))


def test_bench_c5_shards_partition_the_corpus():
	"""bench.py --workload c5 (BASELINE configs[4]): the ranks' shards are disjoint, cover the corpus and are
	balanced by map count (per-video sharding, no collective)."""
	sys.path.insert(0, ROOT)
	import bench
	from retargetvid_b200 import sharding, synth
	specs = synth.config_clips(5, n_clips=64)
	costs = [len(synth.sampling_table(sp['fc'], sp.get('shot_starts', ()), 6)[0]) for sp in specs]
	world = 4
	shards = sharding.lpt_shards(costs, world)
	assert sorted(i for sh in shards for i in sh) == list(range(64))
	loads = [sum(costs[i] for i in sh) for sh in shards]
	assert max(loads) - min(loads) <= max(costs)
	vds = bench.make_workload(16, 1, 2, 'c5')
	want = [synth.config_clips(5, n_clips=16)[i] for i in sharding.my_shard(
		[len(synth.sampling_table(sp['fc'], (), 6)[0]) for sp in synth.config_clips(5, n_clips=16)], 1, 2)]
	assert [v['fc'] for v in vds] == [sp['fc'] for sp in want]
	assert all(v['fc_sel'] == len(synth.sampling_table(v['fc'], (), 6)[0]) for v in vds)


def test_tie_flip_report_is_committed_and_bounded():
	"""tests/golden/tie_flips.json (tests/golden/make_tie_flips.py): the unmodified reference under this box's stock
	numpy (SIMD argsort) against the committed fixtures (portable introsort).  Every difference is a +-1 px box flip,
	and the oracle with that one line changed (np.argsort of the edge weights) reproduces the stock run bit for bit,
	i.e. every flip traces to how the unstable sort permutes tied edge weights."""
	import json
	rep = json.load(open(os.path.join(GOLDEN, 'tie_flips.json')))
	clips = rep['clips']
	assert len(clips) >= 5
	assert sum(c['maps_differing'] for c in clips.values()) > 0      # the effect is real on this box ...
	for name, c in clips.items():
		assert c['max_abs_box_delta'] <= 1, name                      # ... and never more than one pixel
		assert c['oracle_with_stock_sort_equals_stock_reference'], name
		assert c['boxes_differing'] <= c['boxes'] // 8, name
	# fixed-point (oracle, kernels) vs float64 (libraries) stability sums: the closest excess-of-mass decision that is
	# not an exact tie is many orders of magnitude away from float64 rounding (1e-16 per term); exact ties are the
	# trivial 0 == 0 of two empty sums, or equal in float64 too
	assert rep['eom_decisions'] > 1000
	assert rep['eom_nonzero_ties'] == 0
	assert rep['eom_min_relative_gap'] > 1e-9, rep['eom_min_relative_gap']


def test_stock_sort_flips_live_are_tie_permutations_only():
	"""On whatever CPU this runs: the oracle with np.argsort (stock) and with the portable introsort emulation differ
	only through the order of tied edge weights -- same multiset of weights per rank, and boxes within one pixel."""
	from oracle import hdbscan_port as hp
	from oracle import sc_oracle
	from helpers import load_clip_fixture
	vd, over, ratios, fx = load_clip_fixture('fr25')
	CP = sc_oracle.sc_init_crop_params()
	CP.update(over)
	CP['out_ratio'] = ratios[0]
	a = sc_oracle.smart_vid_crop_oracle(vd, CP, cluster_fn=lambda X: hp.fit_predict(X, CP['hdbscan_min'], CP['hdbscan_min_samples'], True, sort='stock'))
	b = fx['bbs_' + ratios[0].replace(':', '-')]
	assert np.max(np.abs(np.array(a['bbs'], dtype=np.int64) - b)) <= 1
	rng = np.random.default_rng(0)
	w = np.where(rng.uniform(size=900) < 0.7, 9, rng.integers(9, 30, 900))
	s1 = np.argsort(w.astype(np.float64))
	s2 = hp.numpy_aquicksort(w)
	assert np.array_equal(w[s1], w[s2])          # both sort; they may only permute equal weights differently


def test_multi_gpu_engine_shards_and_gathers_in_input_order(monkeypatch):
	"""MultiGpuCropEngine (the product's per-video sharding, smartVidCrop.py:2722-2726 in parallel): LPT shards by map
	count, one engine per device, results back in input order.  CropEngine is replaced by a recorder (no GPU here)."""
	from retargetvid_b200 import engine as eng_mod
	calls = {}

	class FakeEngine(object):
		def __init__(self, device=0):
			self.device = device

		def run(self, vds, CP, ratios, **kw):
			calls.setdefault(self.device, []).extend(vd['id'] for vd in vds)
			return [('res', vd['id'], self.device, tuple(ratios)) for vd in vds]

		def close(self):
			pass
	monkeypatch.setattr(eng_mod, 'CropEngine', FakeEngine)
	rng = np.random.default_rng(4)
	vds = [dict(id=i, fc_sel=int(rng.integers(40, 220))) for i in range(37)]
	m = eng_mod.MultiGpuCropEngine([0, 1, 2, 3])
	out = m.run(vds, {}, ['1:3', '3:1'], detail=False)
	assert [o[1] for o in out] == list(range(37))
	assert sorted(sum(calls.values(), [])) == list(range(37)) and set(calls) == {0, 1, 2, 3}
	loads = [sum(vds[i]['fc_sel'] for i in calls[d]) for d in range(4)]
	assert max(loads) - min(loads) <= max(v['fc_sel'] for v in vds)
	with pytest.raises(ValueError):
		eng_mod.MultiGpuCropEngine([0, 0])


def test_result_text_format_and_parser_match_python():
	"""a16 (smartVidCrop.py:2783-2785) and its reader (retargetvid_eval.py:152-159): the library's formatter writes the
	bytes '%d,%d,%d,%d\\n' % ... does, its parser returns what int(c[k]) of line.split(',') returns and rejects what
	Python rejects (host code of the C ABI: runs without a GPU)."""
	from retargetvid_b200 import _cabi
	rng = np.random.default_rng(3)
	b = rng.integers(-3000, 5000, (5000, 4)).astype(np.int32)
	b[0] = [0, -1, 2147483647, -2147483648]
	want = ''.join('%d,%d,%d,%d\n' % tuple(r) for r in b.tolist())
	got = _cabi.format_boxes_txt(b)
	assert got == want.encode()
	assert np.array_equal(_cabi.parse_boxes_txt(got), b)
	assert _cabi.format_boxes_txt(np.zeros((0, 4), dtype=np.int32)) == b''
	assert _cabi.parse_boxes_txt('').shape == (0, 4)
	# what int() accepts: blanks around a field, a plus sign, further fields, \\r\\n line ends, no newline at the end
	text = ' 1, -2 ,+3,4,99\r\n5,6,7,8'
	rows = [ln.split(',') for ln in text.splitlines()]
	py = [[int(c[0]), int(c[1]), int(c[2]), int(c[3])] for c in rows]
	assert _cabi.parse_boxes_txt(text).tolist() == py
	# what Python rejects (IndexError / ValueError) is an error here as well, with the line number
	for bad in ('1,2,3\n', '1,2,3,4\n\n5,6,7,8\n', '1,2,x,4\n', '1,2,3.0,4\n', '1,2,3,\n', '1 2,3,4,5\n'):
		with pytest.raises((ValueError, IndexError)):
			[[int(c[0]), int(c[1]), int(c[2]), int(c[3])] for c in [ln.split(',') for ln in bad.splitlines()]]
		with pytest.raises(_cabi.RvbError):
			_cabi.parse_boxes_txt(bad)


def test_butterworth_design_matches_scipy_and_conditioning_flag():
	"""scipy.signal.butter + lfilter_zi as the library designs them on the host (smartVidCrop.py:1601-1605), and the probe
	that decides whether filtfilt may run in parallel chunks: yes for every well-conditioned filter (the defaults of both
	presets among them), no for high order at low cut-off, where only scipy's sequential order reproduces scipy."""
	from scipy import signal
	from retargetvid_b200 import _cabi
	for order in (1, 2, 3, 4, 5, 6, 7, 8):
		for cut, fr in ((2.0, 30.0), (1.0, 30.0), (2.0, 24.0), (3.5, 25.0), (5.0, 30.0)):
			wn = cut / (0.5 * fr)
			b, a, zi, ok = _cabi.debug_butter(order, wn)
			rb, ra = signal.butter(order, wn, btype='low')
			rzi = signal.lfilter_zi(rb, ra)
			assert np.max(np.abs(b - rb) / np.abs(rb)) < 1e-13 and np.max(np.abs(a - ra) / np.abs(ra)) < 1e-13, (order, cut, fr)
			if ok:
				assert np.max(np.abs(zi - rzi) / np.abs(rzi)) < 1e-11, (order, cut, fr)
	assert _cabi.debug_butter(5, 2.0 / 15.0)[3]            # ICIP-2021 defaults at 30 fps
	assert _cabi.debug_butter(5, 2.0 / 12.0)[3]            # ... at 24 fps
	assert _cabi.debug_butter(2, 1.0 / 15.0)[3]            # ISM-2021 preset (lp_order 2, lp_cutoff 1)
	assert not _cabi.debug_butter(5, 1.0 / 15.0)[3]
	assert not _cabi.debug_butter(7, 2.0 / 15.0)[3]
	with pytest.raises(_cabi.RvbError):
		_cabi.debug_butter(5, 1.5)                          # Wn >= 1: scipy raises, the reference falls back
