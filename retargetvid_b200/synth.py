"""Synthetic ``vid_data`` generator (DHF1K-shaped clips) for tests and benchmarks.

The UNISAL saliency CNN and TransNet shot detector are out of scope (SURVEY.md
section 2), so saliency maps and shot boundaries are synthesised here.  The
dictionary produced has exactly the input keys that the reference's ingest
stage leaves behind (``smartVidCrop.py:480-489,549-554``) so the same object
can be pickled and fed to the reference ``smart_vid_crop`` through its
``temp_path`` cache (``smartVidCrop.py:2244-2256``) as well as to this package.

Recipe: SURVEY.md section 8(d) "Synthetic inputs".
"""
from __future__ import annotations

import numpy as np

# Line counts of annotations/annotator_1/*_1-3.txt in the reference (the 200
# RetargetVid / DHF1K clip lengths, videos 001-100 and 601-700).  Kept as data
# so the benchmark does not need /root/reference at run time.
DHF1K_LENGTHS = None  # filled lazily from tests/golden/dhf1k_lengths.npy or fallback


def _dhf1k_lengths():
	global DHF1K_LENGTHS
	if DHF1K_LENGTHS is None:
		import os
		p = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dhf1k_lengths.npy')
		DHF1K_LENGTHS = np.load(p)
	return DHF1K_LENGTHS


def process_dims(w_orig, h_orig, max_input_d=250):
	"""(SAL_H, SAL_W) as the reference computes them (smartVidCrop.py:252-254)."""
	dsr = float(max(w_orig, h_orig)) / max_input_d
	return int(h_orig / dsr), int(w_orig / dsr)


def sampling_table(fc, shot_starts, skip=6):
	"""Which frames get a saliency map (smartVidCrop.py:386-399).

	A frame is sampled if it is the first frame, ``skip`` frames after the last
	sampled one, the first frame of a shot, or the last frame of the video.
	Returns (true_inds [N], inds_to_orig [F]).
	"""
	starts = set(int(s) for s in shot_starts)
	true_inds = []
	inds_to_orig = []
	for f in range(fc):
		if (len(true_inds) == 0) or (f == true_inds[-1] + skip) or (f in starts) or (f == fc - 1):
			true_inds.append(f)
		inds_to_orig.append(len(true_inds) - 1)
	return true_inds, inds_to_orig


def make_segmentation(fc, shot_starts, inds_to_orig):
	"""Inclusive [start, end] shot table in frame and in map indices
	(smartVidCrop.py:457-474)."""
	starts = sorted(set([0] + [int(s) for s in shot_starts if 0 < int(s) < fc]))
	seg = []
	for k, s in enumerate(starts):
		e = (starts[k + 1] - 1) if k + 1 < len(starts) else fc - 1
		seg.append([s, e])
	seg = np.array(seg, dtype=np.int32)
	seg_sel = np.copy(seg)
	for i in range(seg.shape[0]):
		for j in range(2):
			seg_sel[i][j] = inds_to_orig[seg[i][j]]
	return seg, seg_sel


def blob_log_saliency(rng, n_maps, H, W, cut_maps=(), n_blobs=None):
	"""fp32 log-probability maps [N, H, W]: 1-3 anisotropic Gaussian blobs doing a
	reflected random walk over a low uniform background, normalised to sum 1
	per map (the UNISAL output form, unisal/model.py:497)."""
	yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
	out = np.empty((n_maps, H, W), dtype=np.float32)
	cut_maps = set(int(c) for c in cut_maps)

	def new_scene():
		nb = int(rng.integers(1, 4)) if n_blobs is None else n_blobs
		cx = rng.uniform(0.15 * W, 0.85 * W, nb)
		cy = rng.uniform(0.15 * H, 0.85 * H, nb)
		sx = rng.uniform(6.0, 20.0, nb)
		sy = rng.uniform(6.0, 20.0, nb)
		pk = rng.uniform(0.5, 1.0, nb)
		pk[int(rng.integers(0, nb))] = 1.0
		bg = rng.uniform(0.0, 0.05)
		return nb, cx, cy, sx, sy, pk, bg

	nb, cx, cy, sx, sy, pk, bg = new_scene()
	for i in range(n_maps):
		if i in cut_maps and i > 0:
			nb, cx, cy, sx, sy, pk, bg = new_scene()
		p = np.full((H, W), 0.0)
		for b in range(nb):
			p += pk[b] * np.exp(-0.5 * (((xx - cx[b]) / sx[b]) ** 2 + ((yy - cy[b]) / sy[b]) ** 2))
		p += bg * rng.uniform(0.0, 1.0, (H, W))
		p = p / p.sum()
		out[i] = np.log(np.maximum(p, 1e-30)).astype(np.float32)
		# reflected random walk of the blob centres, 3 px per map
		cx = cx + rng.normal(0.0, 3.0, nb)
		cy = cy + rng.normal(0.0, 3.0, nb)
		cx = np.where(cx < 0, -cx, cx)
		cx = np.where(cx > W - 1, 2 * (W - 1) - cx, cx)
		cy = np.where(cy < 0, -cy, cy)
		cy = np.where(cy > H - 1, 2 * (H - 1) - cy, cy)
	return out


def quantise_u8(logp):
	"""The UNISAL post-process that defines the hot path's uint8 input
	(unisal/train.py:1270-1274): exp, divide by the map max, *255, truncate."""
	out = np.empty(logp.shape, dtype=np.uint8)
	for i in range(logp.shape[0]):
		s = np.exp(logp[i])
		s = (s / np.amax(s)) * 255.0
		out[i] = s.astype('uint8')
	return out


def make_clip(seed, fc=300, w_orig=640, h_orig=360, fr=30.0, shot_starts=(), skip=6,
			max_input_d=250, kind='blobs', keep_logp=False):
	"""One synthetic clip as a reference-shaped ``vid_data`` dict.

	kind: 'blobs' (default recipe), 'empty' (all-zero maps), 'single_pixel',
	'few_points' (fewer non-zero pixels than hdbscan_min+1), 'constant'
	(identical maps, so the centre series is constant -> LOESS NaN path),
	'noise' (uniform random maps, adversarial for clustering).
	"""
	rng = np.random.default_rng(seed)
	H, W = process_dims(w_orig, h_orig, max_input_d)
	true_inds, inds_to_orig = sampling_table(fc, shot_starts, skip)
	seg, seg_sel = make_segmentation(fc, shot_starts, inds_to_orig)
	N = len(true_inds)
	cut_maps = [int(s) for s in seg_sel[:, 0]]
	logp = None
	if kind == 'blobs':
		logp = blob_log_saliency(rng, N, H, W, cut_maps)
		maps = quantise_u8(logp)
	elif kind == 'constant':
		logp = blob_log_saliency(rng, 1, H, W, n_blobs=1)
		logp = np.repeat(logp, N, axis=0)
		maps = quantise_u8(logp)
	elif kind == 'empty':
		maps = np.zeros((N, H, W), dtype=np.uint8)
	elif kind == 'single_pixel':
		maps = np.zeros((N, H, W), dtype=np.uint8)
		for i in range(N):
			maps[i, int(rng.integers(0, H)), int(rng.integers(0, W))] = 255
	elif kind == 'few_points':
		maps = np.zeros((N, H, W), dtype=np.uint8)
		for i in range(N):
			k = int(rng.integers(2, 20))
			ys = rng.integers(0, H, k)
			xs = rng.integers(0, W, k)
			maps[i, ys, xs] = rng.integers(120, 256, k).astype(np.uint8)
	elif kind == 'noise':
		maps = rng.integers(0, 256, (N, H, W)).astype(np.uint8)
		# keep the number of surviving pixels moderate: sparse salt noise + one blob
		keep = rng.uniform(0, 1, (N, H, W)) < 0.01
		blob = quantise_u8(blob_log_saliency(rng, N, H, W, cut_maps, n_blobs=1))
		maps = np.where(keep, maps, blob).astype(np.uint8)
	else:
		raise ValueError('unknown clip kind %r' % kind)

	vd = {}
	# reference layout: [H, W, N], frame index fastest (smartVidCrop.py:286)
	vd['smaps'] = np.ascontiguousarray(np.transpose(maps, (1, 2, 0)))
	vd['segmentation'] = seg
	vd['segmentation_sel'] = seg_sel
	vd['true_inds'] = list(true_inds)
	vd['inds_to_orig'] = list(inds_to_orig)
	vd['fr'] = float(fr)
	vd['fc'] = int(fc)
	vd['fc_sel'] = int(N)
	vd['h_orig'] = int(h_orig)
	vd['w_orig'] = int(w_orig)
	vd['h_process'] = int(H)
	vd['w_process'] = int(W)
	vd['times'] = {'read_init': 0.0, '_read': 0.0, '_read_shot_det': 0.0,
				'_read_sal_det': 0.0, 'read_tidy': 0.0}
	if keep_logp and logp is not None:
		vd['_logp'] = logp
	return vd


def random_shot_starts(rng, fc, lo=60, hi=600):
	"""Shot boundaries drawn i.i.d. uniform lo..hi frames (config C4)."""
	starts = []
	f = int(rng.integers(lo, hi + 1))
	while f < fc - 1:
		starts.append(f)
		f += int(rng.integers(lo, hi + 1))
	return starts


def config_clips(config, n_clips=None, rank=0, world=1):
	"""Clip descriptors (seed, kwargs) for the BASELINE.json configs.

	config 1: one 640x360 300-frame single-shot clip.
	config 3: 200 clips with the DHF1K length distribution.
	config 4: one 1920x1080 10 000-frame multi-shot clip.
	config 5: 2 000 clips drawn with replacement from the config-3 lengths.
	Seeds follow SURVEY.md 8(d): default_rng(1000*config + clip_index).
	"""
	if config == 1:
		return [dict(seed=1000, fc=300)]
	if config == 3:
		lens = _dhf1k_lengths()
		n = len(lens) if n_clips is None else n_clips
		out = []
		for i in range(n):
			rng = np.random.default_rng(3000000 + i)
			fc = int(lens[i % len(lens)])
			# about a third of the clips get one or two cuts
			starts = []
			if rng.uniform() < 0.35:
				starts = sorted(set(int(x) for x in rng.integers(40, fc - 40, int(rng.integers(1, 3)))))
			out.append(dict(seed=1000 * 3 + i + 1000003 * rank, fc=fc, shot_starts=starts))
		return out
	if config == 4:
		rng = np.random.default_rng(4000)
		return [dict(seed=4000, fc=10000, w_orig=1920, h_orig=1080,
					shot_starts=random_shot_starts(rng, 10000))]
	if config == 5:
		lens = _dhf1k_lengths()
		n = 2000 if n_clips is None else n_clips
		rng = np.random.default_rng(5000)
		draw = rng.integers(0, len(lens), n)
		return [dict(seed=1000 * 5 + i, fc=int(lens[draw[i]])) for i in range(n)][rank::world]
	raise ValueError('unknown config %r' % config)
