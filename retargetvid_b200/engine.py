"""Host-side batching over the C ABI: packs reference-shaped ``vid_data`` dicts
into one ``rvb_crop_track_batch`` call per process size and unpacks the results.

Host logic only (dict plumbing); all arithmetic runs in the CUDA library.
"""
import ctypes as C
import threading

import numpy as np

from . import _cabi, sharding


def parse_ratio(out_ratio):
	"""'a:b' -> (a, b) as floats (smartVidCrop.py:950-952)."""
	c = str(out_ratio).split(':')
	return float(c[0]), float(c[1])


class ClipResult(object):
	"""Per-clip outputs of one batched call (numpy views / python scalars)."""
	__slots__ = ('status', 'boxes', 'dx', 'dy', 'dxnf', 'dynf', 'jumps', 'empty', 'series', 'map_scores', 'mean_sal_score',
				'cvrg_scores', 'dims', 'map_info', 'filtered', 'filtered_hwn')


class CropEngine(object):
	def __init__(self, device=0):
		self.ctx = _cabi.Context(device)
		self.device = device

	def close(self):
		self.ctx.close()

	def run(self, vds, CP, ratios, detail=True, want_filtered=False, cvrg_window='reference', raise_on_clip_error=True,
			np_int=False):
		"""vds: list of vid_data dicts; ratios: list of 'a:b' strings.
		np_int: focus stability samples diagonal moves as the reference does on the numpy it pins (np.int exists,
		smartVidCrop.py:1377,1384) instead of returning 255 for them as it does on numpy >= 1.24.
		raise_on_clip_error=False: a clip that fails on its own (RVB_ERR_CAPACITY: a map with more salient pixels than
		RVB_MAX_POINTS; RVB_ERR_NO_CENTRES: no salient pixel at all) is reported through ClipResult.status and the
		other clips of the batch keep their results.
		Returns a list of ClipResult in input order."""
		params = _cabi.params_from_crop_params(CP, cvrg_window, np_int)
		results = [None] * len(vds)
		groups = {}
		for i, vd in enumerate(vds):
			groups.setdefault((int(vd['h_process']), int(vd['w_process'])), []).append(i)
		for (H, W), idxs in groups.items():
			self._run_group([vds[i] for i in idxs], idxs, results, params, ratios, H, W, detail, want_filtered,
							raise_on_clip_error)
		return results

	def _run_group(self, vds, idxs, results, params, ratios, H, W, detail, want_filtered, raise_on_clip_error):
		nc = len(vds)
		R = len(ratios)
		clips = (_cabi.rvb_clip * nc)()
		shots = []
		tinds = []
		keep = []
		ptrs = (C.c_void_p * nc)()
		mo = fo = so = 0
		for i, vd in enumerate(vds):
			sm = vd['smaps']
			if not (isinstance(sm, np.ndarray) and sm.dtype == np.uint8 and sm.ndim == 3 and sm.shape[0] == H and sm.shape[1] == W):
				raise ValueError('vid_data[%d]["smaps"] must be uint8 [h_process, w_process, fc_sel]' % i)
			sm = np.ascontiguousarray(sm)
			keep.append(sm)
			N = int(sm.shape[2])
			seg = np.asarray(vd['segmentation'], dtype=np.int64).reshape(-1, 2)
			ssel = np.asarray(vd['segmentation_sel'], dtype=np.int64).reshape(-1, 2)
			if N != int(vd['fc_sel']) or N != len(vd['true_inds']) or seg.shape != ssel.shape:
				raise ValueError('vid_data[%d]: inconsistent fc_sel / true_inds / segmentation' % i)
			c = clips[i]
			c.n_maps = N
			c.n_frames = int(vd['fc'])
			c.n_shots = int(seg.shape[0])
			c.h_orig = int(vd['h_orig'])
			c.w_orig = int(vd['w_orig'])
			c.fr = float(vd['fr'])
			c.map_offset = mo
			c.frame_offset = fo
			c.shot_offset = so
			shots.append(np.concatenate([seg, ssel], axis=1))
			tinds.append(np.asarray(vd['true_inds'], dtype=np.int32))
			ptrs[i] = sm.ctypes.data
			mo += N
			fo += c.n_frames
			so += c.n_shots
		shots = np.ascontiguousarray(np.concatenate(shots, axis=0), dtype=np.int32)
		tinds = np.ascontiguousarray(np.concatenate(tinds), dtype=np.int32)
		NM, NF = mo, fo
		b = _cabi.rvb_batch()
		b.n_clips = nc
		b.h_process = H
		b.w_process = W
		b.row_stride = 0
		b.maps_kind = _cabi.RVB_MAPS_U8_HWN
		b.mem_space = _cabi.RVB_MEM_HOST
		b.n_ratios = R
		for r, s in enumerate(ratios):
			b.ratio_w[r], b.ratio_h[r] = parse_ratio(s)
		b.clips = clips
		b.shots = shots.ctypes.data
		b.true_inds = tinds.ctypes.data
		b.maps = None
		b.clip_maps = ptrs
		boxes = np.empty((R, NF, 4), dtype=np.int32)
		b.boxes = boxes.ctypes.data
		status = np.zeros(nc, dtype=np.int32)
		b.clip_status = status.ctypes.data
		dims = np.zeros((nc, R, 9), dtype=np.int32)
		b.clip_dims = dims.ctypes.data
		cscores = np.zeros((nc, 1 + R), dtype=np.float64)
		b.clip_scores = cscores.ctypes.data
		centres = series = empty = mscores = minfo = filt = cnf = None
		if detail:
			centres = np.empty((2, NM), dtype=np.float64)
			cnf = np.full((3, NM), 255.0, dtype=np.float64)
			b.centres_nf = cnf.ctypes.data
			series = np.empty((6, NF), dtype=np.float64)
			empty = np.empty(NM, dtype=np.uint8)
			mscores = np.empty(NM, dtype=np.float64)
			minfo = np.empty((NM, 4), dtype=np.int32)
			b.centres = centres.ctypes.data
			b.series = series.ctypes.data
			b.empty = empty.ctypes.data
			b.map_scores = mscores.ctypes.data
			b.map_info = minfo.ctypes.data
		if want_filtered == 'hwn':
			# the reference's layout: per clip [H][W][n_maps], packed back to back
			filt = np.empty(NM * H * W, dtype=np.uint8)
			b.filtered_maps = filt.ctypes.data
			b.filtered_layout = _cabi.RVB_FILTERED_HWN
		elif want_filtered:
			filt = np.empty((NM, H, W), dtype=np.uint8)
			b.filtered_maps = filt.ctypes.data
			b.row_stride_out = W
		err = None
		try:
			self.ctx.crop_track_batch(params, b)
		except _cabi.RvbError as e:
			if e.code not in (_cabi.RVB_ERR_CAPACITY, _cabi.RVB_ERR_NO_CENTRES) or raise_on_clip_error:
				raise
			err = e
		for i in range(nc):
			c = clips[i]
			res = ClipResult()
			res.status = int(status[i])
			f0, f1 = c.frame_offset, c.frame_offset + c.n_frames
			m0, m1 = c.map_offset, c.map_offset + c.n_maps
			res.boxes = boxes[:, f0:f1, :]
			res.dims = dims[i]
			res.mean_sal_score = float(cscores[i, 0])
			res.cvrg_scores = cscores[i, 1:]
			res.dx = res.dy = res.empty = res.series = res.map_scores = res.map_info = res.filtered = None
			res.dxnf = res.dynf = res.jumps = None
			if detail:
				res.dxnf = cnf[0, m0:m1]
				res.dynf = cnf[1, m0:m1]
				res.jumps = cnf[2, m0:m1]
				res.dx = centres[0, m0:m1]
				res.dy = centres[1, m0:m1]
				res.empty = empty[m0:m1]
				res.series = series[:, f0:f1]
				res.map_scores = mscores[m0:m1]
				res.map_info = minfo[m0:m1]
			res.filtered_hwn = None
			if want_filtered == 'hwn':
				res.filtered_hwn = filt[m0 * H * W:m1 * H * W].reshape(H, W, m1 - m0)
			elif want_filtered:
				res.filtered = filt[m0:m1]
			results[idxs[i]] = res
		del keep
		return err


class MultiGpuCropEngine(object):
	"""Per-video sharding over the GPUs of one box: replaces the serial loop over videos of the
	reference's driver (smartVidCrop.py:2722-2726).  One CropEngine (context, stream, workspace) and one host thread per
	device, videos assigned longest-processing-time-first by their number of saliency maps, all target ratios of a video
	on the same device, results gathered on the host in input order.  No collective: videos never interact."""

	def __init__(self, devices):
		devices = [int(d) for d in devices]
		if not devices or len(set(devices)) != len(devices):
			raise ValueError('devices must be a non-empty list of distinct CUDA device indices')
		self.devices = devices
		self.engines = [CropEngine(d) for d in devices]

	def close(self):
		for e in self.engines:
			e.close()

	def shards(self, vds):
		return sharding.lpt_shards([int(vd['fc_sel']) for vd in vds], len(self.engines))

	def run(self, vds, CP, ratios, **kw):
		"""Same arguments and result as CropEngine.run."""
		shards = self.shards(vds)
		parts = [None] * len(self.engines)
		errs = [None] * len(self.engines)

		def work(k):
			try:
				if shards[k]:
					parts[k] = self.engines[k].run([vds[i] for i in shards[k]], CP, ratios, **kw)
				else:
					parts[k] = []
			except BaseException as e:      # re-raised on the calling thread
				errs[k] = e
		ths = [threading.Thread(target=work, args=(k,)) for k in range(len(self.engines))]
		for t in ths:
			t.start()
		for t in ths:
			t.join()
		for e in errs:
			if e is not None:
				raise e
		out = [None] * len(vds)
		for k, idxs in enumerate(shards):
			for i, r in zip(idxs, parts[k]):
				out[i] = r
		return out
