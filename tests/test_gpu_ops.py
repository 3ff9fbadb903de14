"""GPU tests of the entry points around the kernels: the torch custom op fed with device-resident tensors on the
caller's stream, per-video sharding over several GPUs (8e), per-clip error isolation in a batch, and
the lattice-local Prim against the all-pairs Prim.  Needs a GPU: pytest -m gpu."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
	from retargetvid_b200.engine import CropEngine
	e = CropEngine(0)
	yield e
	e.close()


def _clips(n, seed0=7700):
	from retargetvid_b200 import synth
	return [synth.make_clip(seed0 + i, fc=70 + 23 * i, shot_starts=[31] if i % 2 else [], keep_logp=True) for i in range(n)]


def test_torch_op_device_tensors_on_caller_stream(engine):
	"""torch.ops.retargetvid_b200.crop_track: uint8 [N,H,256] and float32 log-saliency [N,H,W] tensors that live on the
	device (where UNISAL leaves them, unisal/train.py:852-854 before the .cpu()), launched on a non-default stream:
	same boxes and centres as the host-buffer path, and nothing but the status leaves the device."""
	import torch
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import torch_op
	vds = _clips(4)
	CP = svc.sc_init_crop_params()
	ratios = ['1:3', '9:16']
	want = engine.run(vds, CP, ratios, detail=True)
	want_boxes = np.concatenate([w.boxes for w in want], axis=1)
	want_dx = np.concatenate([w.dx for w in want])
	st = torch.cuda.Stream()
	with torch.cuda.stream(st):
		u8 = torch.zeros((sum(v['fc_sel'] for v in vds), 140, 256), dtype=torch.uint8, device='cuda')
		mo = 0
		for vd in vds:
			n = vd['fc_sel']
			u8[mo:mo + n, :, :250] = torch.from_numpy(np.ascontiguousarray(np.transpose(vd['smaps'], (2, 0, 1)))).cuda()
			u8[mo:mo + n, :, 250:] = 99          # row padding is the producer's garbage
			mo += n
		boxes, centres, status = torch_op.crop_track(u8, vds, CP, ratios)
		assert boxes.is_cuda and centres.is_cuda and boxes.dtype == torch.int32
		f32 = torch.from_numpy(np.ascontiguousarray(np.concatenate([vd['_logp'] for vd in vds]), dtype=np.float32)).cuda()
		boxes_f, centres_f, status_f = torch_op.crop_track(f32, vds, CP, ratios)
	st.synchronize()
	assert status.cpu().tolist() == [0] * len(vds) and status_f.cpu().tolist() == [0] * len(vds)
	assert np.array_equal(boxes.cpu().numpy(), want_boxes)
	assert np.array_equal(centres.cpu().numpy()[0], want_dx)
	# float32 entry: the fused uint8 post-process may differ from numpy's by one grey level on a few pixels
	# (tests/test_gpu_parity.py::test_float32_entry); on these clips the boxes are the same
	assert np.array_equal(boxes_f.cpu().numpy(), want_boxes)


def test_multi_gpu_sharded_equals_single_gpu(engine):
	"""smart_vid_crop_batch(devices=[0, 1]): per-video LPT shards, one context and host thread per GPU, results gathered
	on the host in input order == the single-GPU result."""
	import torch
	if torch.cuda.device_count() < 2:
		pytest.skip('needs 2 GPUs')
	from retargetvid_b200 import smartVidCrop as svc
	vds = _clips(9, seed0=7800)
	CP = svc.sc_init_crop_params()
	ratios = ['1:3', '3:1', '4:5']
	one = svc.smart_vid_crop_batch(vds, CP, ratios, device=0)
	two = svc.smart_vid_crop_batch(vds, CP, ratios, devices=[0, 1])
	for a, b in zip(one, two):
		assert a.status == 0 and b.status == 0
		assert np.array_equal(a.boxes, b.boxes) and np.array_equal(a.dx, b.dx) and np.array_equal(a.series, b.series)


def test_one_failing_clip_does_not_take_the_batch_down(engine):
	"""A clip with a map of more salient pixels than RVB_MAX_POINTS reports RVB_ERR_CAPACITY for itself; the other
	clips of the batch keep results identical to running them alone (the reference's loop would also only lose that
	one video)."""
	from retargetvid_b200 import _cabi
	from retargetvid_b200 import smartVidCrop as svc
	vds = _clips(3, seed0=7900)
	bad = dict(vds[1])
	sm = bad['smaps'].copy()
	sm[:, :, 2] = 200          # 35 000 salient pixels in one map
	bad['smaps'] = sm
	batch = [vds[0], bad, vds[2]]
	CP = svc.sc_init_crop_params()
	res = svc.smart_vid_crop_batch(batch, CP, ['1:3'])
	assert res[1].status == _cabi.RVB_ERR_CAPACITY
	for i in (0, 2):
		alone = svc.smart_vid_crop_batch([batch[i]], CP, ['1:3'])[0]
		assert res[i].status == 0 and np.array_equal(res[i].boxes, alone.boxes)
	with pytest.raises(_cabi.RvbError):
		svc.smart_vid_crop_batch(batch, CP, ['1:3'], raise_on_clip_error=True)


def test_lattice_local_prim_equals_all_pairs_prim():
	"""fprim_kernel (lattice offsets + bucket queue + lazy far keys) against the round-1 all-pairs prim_kernel
	(the default; RVB_FRONTIER_PRIM=1 selects the lattice-local one) on blobs, salt noise (every step a stall), sparse maps, cuts (blended chain maps) and the
	min_cluster_size 5 / min_samples 3 setting (core distances below 9): identical filtered maps, per-map records,
	centres and boxes."""
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	from retargetvid_b200.engine import CropEngine
	vds = [synth.make_clip(8000 + i, fc=90 + 11 * i, shot_starts=[40, 47] if i % 2 else []) for i in range(6)]
	vds += [synth.make_clip(8100, fc=60, kind='noise', shot_starts=[25]), synth.make_clip(8101, fc=50, kind='few_points'),
			synth.make_clip(8102, fc=40, kind='single_pixel'), synth.make_clip(8103, fc=45, kind='constant')]
	outs = {}
	for over in ({}, dict(hdbscan_min=5, hdbscan_min_samples=3, t_threshold=90, select_sum=1)):
		CP = svc.sc_init_crop_params()
		CP.update(over)
		for dense in (True, False):
			if dense:
				os.environ.pop('RVB_FRONTIER_PRIM', None)
			else:
				os.environ['RVB_FRONTIER_PRIM'] = '1'
			e = CropEngine(0)
			try:
				outs[dense] = e.run(vds, CP, ['1:3', '3:1'], detail=True, want_filtered=True, raise_on_clip_error=False)
			finally:
				e.close()
				os.environ.pop('RVB_FRONTIER_PRIM', None)
		for i, (a, b) in enumerate(zip(outs[True], outs[False])):
			assert a.status == b.status, i
			assert np.array_equal(a.filtered, b.filtered), (i, over)
			assert np.array_equal(a.map_info, b.map_info), (i, over)
			assert np.array_equal(a.boxes, b.boxes) and np.array_equal(a.dx, b.dx, equal_nan=True), (i, over)


def _write_boxes(path, boxes):
	with open(path, 'w') as fp:
		fp.write(''.join('%d,%d,%d,%d\n' % tuple(int(v) for v in b) for b in boxes))


def test_evaluator_cli_reproduces_the_reference_csv(tmp_path, capsys, monkeypatch):
	"""python -m retargetvid_b200.retargetvid_eval on the shipped results + the 6 annotators (written out from the
	committed fixture): the CSV line the unmodified retargetvid_eval.py prints, and its validity
	report (retargetvid_eval.py:102-121)."""
	from helpers import GOLDEN
	from retargetvid_b200 import retargetvid_eval as rev
	z = np.load(os.path.join(GOLDEN, 'eval_fixture.npz'))
	vids = [int(v) for v in z['vid_inds']]
	ann_dir = tmp_path / 'annotations'
	res_dir = tmp_path / 'results' / 'smartvidcrop'
	res_dir.mkdir(parents=True)
	for u in range(1, 7):
		d = ann_dir / ('annotator_%d' % u)
		d.mkdir(parents=True)
		for ar in ('1-3', '3-1'):
			lens = z['annot_len_%d_%s' % (u, ar)]
			offs = np.concatenate([[0], np.cumsum(lens)])
			boxes = z['annot_%d_%s' % (u, ar)]
			for i, v in enumerate(vids):
				_write_boxes(d / ('%03d_%s.txt' % (v, ar)), boxes[offs[i]:offs[i + 1]])
	for ar in ('1-3', '3-1'):
		lens = z['method_len_%s' % ar]
		offs = np.concatenate([[0], np.cumsum(lens)])
		boxes = z['method_%s' % ar]
		for i, v in enumerate(vids):
			_write_boxes(res_dir / ('%03d_%s.txt' % (v, ar)), boxes[offs[i]:offs[i + 1]])
	monkeypatch.chdir(tmp_path)
	lines = rev.main([str(tmp_path / 'results'), '--annotations', str(ann_dir)])
	want = str(z['eval_csv']).splitlines()[-1].split(',')
	got = lines[-1].split(',')
	assert got[1:4] == want[1:4] and got[12:15] == want[12:15] and got[-1] == '0'
	out = capsys.readouterr().out
	assert '(file errors:0 + frame count errors:0)' in out and 'valid runs::' in out
	assert os.path.isfile(tmp_path / 'eval_current.txt')


def test_iou_per_annotator_length_and_malformed_boxes(engine):
	"""The reference breaks out of the frame loop per annotator (retargetvid_eval.py:163-179): an annotator with a
	shorter list is averaged over its own frames only.  A box with x2 < x1 is not an IoU in [0, 1]: loud error."""
	from oracle import eval_oracle
	from retargetvid_b200 import _cabi
	from retargetvid_b200 import retargetvid_eval as rev
	rng = np.random.default_rng(11)

	def boxes(n):
		b = rng.integers(0, 600, (n, 4)).astype(np.int32)
		b[:, 2:] = np.maximum(b[:, 2:], b[:, :2])
		return b
	method = {5: boxes(60), 9: boxes(33)}
	annots = [{5: boxes(60), 9: boxes(40)}, {5: boxes(41), 9: boxes(40)}]       # annotator 2 stops early on video 5
	fc = {5: 60, 9: 40}
	vid_iou, _, order = rev.evaluate_arrays(engine.ctx, method, annots, fc)
	for i, v in enumerate(order):
		for u in range(2):
			want = eval_oracle.video_iou([list(b) for b in method[v]], [list(b) for b in annots[u][v]], fc[v])
			assert vid_iou[i][u] == want, (v, u)
	# a box with x2 < x1 has a negative area; its intersection with anything is 0, so the IoU is 0 / (aA + aB): 0.0 like the
	# reference unless the union is exactly 0, where the reference raises ZeroDivisionError -> loud error here too
	bad = {5: method[5].copy(), 9: method[9]}
	bad[5][3] = [5, 0, 0, 0]                 # area (0 - 5 + 1) * 1 = -4
	ann_bad = [{5: annots[0][5].copy(), 9: annots[0][9]}, annots[1]]
	ann_bad[0][5][3] = [0, 0, 1, 1]          # area 4: union 0
	with pytest.raises(_cabi.RvbError):
		rev.evaluate_arrays(engine.ctx, bad, ann_bad, fc)
	ok = {5: method[5].copy(), 9: method[9]}
	ok[5][3] = [400, 10, 100, 50]            # x2 < x1 with a non-zero union: IoU 0.0, as the reference computes it
	vid_iou2, _, _ = rev.evaluate_arrays(engine.ctx, ok, annots, fc)
	assert vid_iou2[0][0] == eval_oracle.video_iou([list(b) for b in ok[5]], [list(b) for b in annots[0][5]], fc[5])


def test_renderer_crop_equals_numpy_slicing(engine):
	"""rvb_crop_frames / smartVidCrop.crop_frames: out[f] == frame[f][y1:y2, x1:x2, :] (sc_renderer,
	smartVidCrop.py:1906-1912) for row lengths that are multiples of 8 and of 4 bytes and for odd ones (202 x 3 bytes),
	boxes from the crop track itself."""
	from retargetvid_b200 import _cabi
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vd = synth.make_clip(8300, fc=40)
	CP = svc.sc_init_crop_params()
	res = svc.smart_vid_crop_batch([vd], CP, ['1:3', '9:16', '4:5', '3:1'])[0]
	rng = np.random.default_rng(3)
	frames = rng.integers(0, 256, (vd['fc'], vd['h_orig'], vd['w_orig'], 3)).astype(np.uint8)
	for r in range(4):
		bbs = res.boxes[r]
		got = svc.crop_frames(frames, bbs)
		for f in (0, 7, vd['fc'] - 1):
			x1, y1, x2, y2 = (int(v) for v in bbs[f])
			assert np.array_equal(got[f], frames[f][y1:y2, x1:x2, :]), (r, f)
	# single channel, and a box that does not have the common size is refused
	gray = np.ascontiguousarray(frames[:, :, :, 0:1])
	got = svc.crop_frames(gray, res.boxes[0])
	x1, y1, x2, y2 = (int(v) for v in res.boxes[0][5])
	assert np.array_equal(got[5], gray[5][y1:y2, x1:x2, :])
	bad = np.array(res.boxes[0])
	bad[3, 2] += 1
	with pytest.raises(_cabi.RvbError):
		svc.crop_frames(frames, bad)


def test_iou_division_and_exact_mean_on_random_boxes_at_size(engine):
	"""300k frames x 3 annotators of random well-formed boxes (extents 1 .. 8191 and a few beyond, so both integer paths
	run; disjoint pairs; videos of 1, 31, 32, 33, 255 ... frames so that every padding case of the warp slots occurs):
	every per-frame IoU equals numpy's IEEE division of the two integers bit for bit (retargetvid_eval.py:10-27), every
	per-(video, annotator) mean equals the exactly rounded mean of those doubles (statistics.mean, :193)."""
	import statistics
	from fractions import Fraction
	from retargetvid_b200 import _cabi
	rng = np.random.default_rng(2024)
	lens = [1, 31, 32, 33, 255, 256, 257, 613, 7, 1000] * 80
	lens += [int(v) for v in rng.integers(1, 2000, 120)]
	V, U = len(lens), 3
	foff = np.zeros(V + 1, dtype=np.int64)
	foff[1:] = np.cumsum(lens)
	NF = int(foff[-1])

	def boxes(n, big_every):
		w = rng.integers(1, 8192, n)
		h = rng.integers(1, 8192, n)
		big = (np.arange(n) % big_every) == 0
		w[big] += 9000
		h[big] = h[big] % 4000 + 1             # unions stay below 2^28 (the fixed point holds IoUs >= 2^-28 exactly)
		x1 = rng.integers(0, 12000, n)
		y1 = rng.integers(0, 12000, n)
		# a third of the boxes small and close together so that intersections are common
		near = rng.random(n) < 0.6
		x1[near] %= 300
		y1[near] %= 300
		return np.stack([x1, y1, x1 + w - 1, y1 + h - 1], axis=1).astype(np.int32)
	method = boxes(NF, 997)
	annot = np.stack([boxes(NF, 1013 + u) for u in range(U)])
	neval = np.tile(np.array(lens, dtype=np.int32)[:, None], (1, U))
	neval[5, 1] = max(1, lens[5] // 2)          # one annotator stops early
	fiou = np.empty((U, NF), dtype=np.float64)
	acc = np.zeros((V, U, 2), dtype=np.uint64)
	ib = _cabi.rvb_iou_batch()
	ib.n_videos, ib.n_users, ib.mem_space = V, U, _cabi.RVB_MEM_HOST
	ib.frame_offset = foff.ctypes.data
	ib.n_eval_user = neval.ctypes.data
	ib.method_boxes, ib.annot_boxes, ib.frame_iou, ib.acc = method.ctypes.data, annot.ctypes.data, fiou.ctypes.data, acc.ctypes.data
	engine.ctx.iou_batch(ib)
	m = method.astype(np.int64)
	for u in range(U):
		g = annot[u].astype(np.int64)
		iw = np.maximum(0, np.minimum(g[:, 2], m[:, 2]) - np.maximum(g[:, 0], m[:, 0]) + 1)
		ih = np.maximum(0, np.minimum(g[:, 3], m[:, 3]) - np.maximum(g[:, 1], m[:, 1]) + 1)
		inter = iw * ih
		uni = (g[:, 2] - g[:, 0] + 1) * (g[:, 3] - g[:, 1] + 1) + (m[:, 2] - m[:, 0] + 1) * (m[:, 3] - m[:, 1] + 1) - inter
		want = inter.astype(np.float64) / uni.astype(np.float64)
		assert (inter > 0).sum() > NF // 10
		assert np.array_equal(fiou[u].view(np.uint64), want.view(np.uint64)), int((fiou[u] != want).sum())
		for v in list(range(0, V, 37)) + [5]:
			n = int(neval[v, u])
			vals = want[foff[v]:foff[v] + n]
			exact = sum((Fraction(float(x)) for x in vals), Fraction(0)) / n
			got = engine.ctx.iou_mean_from_acc(acc[v, u, 0], acc[v, u, 1], n)
			assert got == float(exact) == statistics.mean([float(x) for x in vals]), (v, u)


def test_filtered_maps_in_the_reference_layout(engine):
	"""RVB_FILTERED_HWN: the filtered maps come back per clip as [H][W][n_maps] (what smart_vid_crop leaves in
	vid_data['smaps'], smartVidCrop.py:2366-2373), transposed on the device -- equal to the [N][H][W] output, for clips
	with 1 .. >60 maps (several map tiles, every alignment of the per-pixel runs)."""
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	CP = svc.sc_init_crop_params()
	vds = [synth.make_clip(8800 + i, fc=fc, shot_starts=ss) for i, (fc, ss) in enumerate(
		[(6, []), (45, []), (300, [150]), (413, [100, 101, 300]), (1000, [500]), (59, []), (367, [])])]
	a = engine.run(vds, CP, ['1:3'], detail=False, want_filtered=True)
	b = engine.run(vds, CP, ['1:3'], detail=False, want_filtered='hwn')
	assert max(int(vd['fc_sel']) for vd in vds) > 120
	for ra, rb, vd in zip(a, b, vds):
		assert rb.filtered_hwn.shape == (vd['h_process'], vd['w_process'], int(vd['fc_sel']))
		assert np.array_equal(np.transpose(ra.filtered, (1, 2, 0)), rb.filtered_hwn)
		assert np.array_equal(ra.boxes, rb.boxes)


@pytest.mark.parametrize('rtype,factor', [(2, 4), (2, 2), (2, 3), (2, 5), (1, 3), (1, 5), (3, 2)])
def test_downscaled_clustering_vs_oracle(engine, rtype, factor):
	"""resize_factor != 1 (smartVidCrop.py:1078-1084,1158,1184): clustering on a down-scaled copy -- INTER_LINEAR (1),
	INTER_CUBIC (2), INTER_NEAREST (3) -- scaled back with INTER_LINEAR; filtered maps and boxes equal the oracle's
	(oracle/cv_resize.py is checked against cv2 in the CPU suite, the factor-4 paths against reference fixtures)."""
	from oracle import sc_oracle
	from retargetvid_b200 import smartVidCrop as svc
	from retargetvid_b200 import synth
	vd = synth.make_clip(9100 + 10 * rtype + factor, fc=80, shot_starts=[33])
	CP = svc.sc_init_crop_params()
	CP.update(dict(t_threshold=90, hdbscan_min=5, hdbscan_min_samples=3, resize_factor=factor, resize_type=rtype))
	CP['out_ratio'] = '1:3'
	res = engine.run([vd], CP, ['1:3'], detail=True, want_filtered='hwn')[0]
	want = sc_oracle.smart_vid_crop_oracle(vd, dict(CP))
	bad = (res.filtered_hwn != want['smaps_filtered']).any(axis=(0, 1))
	assert not bad.any(), 'filtered maps differ in maps %s' % np.nonzero(bad)[0].tolist()
	assert np.array_equal(res.boxes[0], np.array(want['bbs'], dtype=np.int32))
