"""ctypes binding of the C ABI declared in include/retargetvid_b200.h.

The shared library is built in-tree (retargetvid_b200/lib/libretargetvid_b200.so)
by ``__graft_entry__.build()`` / ``python -m retargetvid_b200.build``.  There is no
CPU fallback: if the library is missing or CUDA is unavailable every entry
point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libretargetvid_b200.so')

RVB_OK = 0
RVB_ERR_INVALID = -1
RVB_ERR_CUDA = -2
RVB_ERR_UNSUPPORTED = -3
RVB_ERR_CAPACITY = -4
RVB_ERR_NO_CENTRES = -5
RVB_MAX_POINTS = 8192
RVB_MAX_RATIOS = 8
RVB_MEM_HOST = 0
RVB_MEM_DEVICE = 1
RVB_FILTERED_NHW, RVB_FILTERED_HWN = 0, 1
RVB_MAPS_U8_NHW = 0
RVB_MAPS_U8_HWN = 1
RVB_MAPS_F32_NHW = 2

EXPORTS = ['rvb_version', 'rvb_last_error', 'rvb_ctx_create', 'rvb_ctx_destroy', 'rvb_ctx_set_stream',
		'rvb_ctx_synchronize', 'rvb_ctx_launch_count', 'rvb_ctx_last_map_kernel_ms', 'rvb_ctx_last_stage_ms', 'rvb_params_default',
		'rvb_crop_track_batch', 'rvb_iou_batch_run', 'rvb_iou_mean_from_acc', 'rvb_debug_cluster_labels',
		'rvb_debug_smooth_series', 'rvb_ctx_phase_cycles', 'rvb_crop_frames', 'rvb_ctx_last_iou_kernel_ms', 'rvb_ctx_last_host_us', 'rvb_format_boxes_txt', 'rvb_parse_boxes_txt', 'rvb_debug_butter']


class RvbError(RuntimeError):
	def __init__(self, code, msg):
		RuntimeError.__init__(self, 'retargetvid_b200 error %d: %s' % (code, msg))
		self.code = code


class rvb_params(C.Structure):
	_fields_ = [('t_threshold', C.c_int32), ('clust_filt', C.c_int32), ('hdbscan_min', C.c_int32),
				('hdbscan_min_samples', C.c_int32), ('select_sum', C.c_int32), ('op_close', C.c_int32),
				('com_km', C.c_int32), ('t_border', C.c_int32), ('loess_filt', C.c_int32),
				('loess_degree', C.c_int32), ('lp_filt', C.c_int32), ('lp_order', C.c_int32),
				('shift_time', C.c_int32), ('exit_on_low_cvrg', C.c_int32), ('cvrg_window', C.c_int32),
				('resize_type', C.c_int32), ('focus_stability', C.c_int32), ('min_d_jump', C.c_int32), ('skip', C.c_int32),
				('np_int_compat', C.c_int32), ('foces_stab_t', C.c_double), ('foces_stab_s', C.c_double),
				('loess_w_secs', C.c_double), ('lp_cutoff', C.c_double),
				('resize_factor', C.c_double), ('t_cvrg', C.c_double)]


class rvb_clip(C.Structure):
	_fields_ = [('n_maps', C.c_int32), ('n_frames', C.c_int32), ('n_shots', C.c_int32), ('h_orig', C.c_int32),
				('w_orig', C.c_int32), ('reserved0', C.c_int32), ('fr', C.c_double), ('map_offset', C.c_int64),
				('frame_offset', C.c_int64), ('shot_offset', C.c_int64)]


class rvb_batch(C.Structure):
	_fields_ = [('n_clips', C.c_int32), ('h_process', C.c_int32), ('w_process', C.c_int32),
				('row_stride', C.c_int32), ('maps_kind', C.c_int32), ('mem_space', C.c_int32),
				('n_ratios', C.c_int32), ('reserved0', C.c_int32),
				('ratio_w', C.c_double * RVB_MAX_RATIOS), ('ratio_h', C.c_double * RVB_MAX_RATIOS),
				('clips', C.POINTER(rvb_clip)), ('shots', C.c_void_p), ('true_inds', C.c_void_p),
				('maps', C.c_void_p), ('boxes', C.c_void_p), ('centres', C.c_void_p), ('centres_nf', C.c_void_p),
				('empty', C.c_void_p),
				('series', C.c_void_p), ('map_scores', C.c_void_p), ('clip_scores', C.c_void_p),
				('clip_dims', C.c_void_p), ('filtered_maps', C.c_void_p), ('row_stride_out', C.c_int32),
				('filtered_layout', C.c_int32), ('map_info', C.c_void_p), ('clip_status', C.c_void_p),
				('clip_maps', C.POINTER(C.c_void_p))]


class rvb_iou_batch(C.Structure):
	_fields_ = [('n_videos', C.c_int32), ('n_users', C.c_int32), ('mem_space', C.c_int32), ('reserved0', C.c_int32),
				('frame_offset', C.c_void_p), ('n_eval', C.c_void_p), ('method_boxes', C.c_void_p),
				('annot_boxes', C.c_void_p), ('frame_iou', C.c_void_p), ('acc', C.c_void_p), ('n_eval_user', C.c_void_p), ('n_bad', C.c_void_p)]


_lib = None


def load_library():
	"""Loads the CUDA library; raises if it has not been built (no fallback)."""
	global _lib
	if _lib is not None:
		return _lib
	if not os.path.isfile(LIB_PATH):
		raise RvbError(RVB_ERR_CUDA, 'CUDA library not built: %s is missing (run `python -m retargetvid_b200.build`)' % LIB_PATH)
	lib = C.CDLL(LIB_PATH)
	lib.rvb_version.restype = C.c_char_p
	lib.rvb_last_error.restype = C.c_char_p
	lib.rvb_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
	lib.rvb_ctx_destroy.argtypes = [C.c_void_p]
	lib.rvb_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
	lib.rvb_ctx_synchronize.argtypes = [C.c_void_p]
	lib.rvb_ctx_launch_count.argtypes = [C.c_void_p]
	lib.rvb_ctx_launch_count.restype = C.c_int64
	lib.rvb_ctx_last_map_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
	lib.rvb_ctx_last_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
	lib.rvb_params_default.argtypes = [C.POINTER(rvb_params), C.c_int]
	lib.rvb_ctx_phase_cycles.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
	lib.rvb_crop_track_batch.argtypes = [C.c_void_p, C.POINTER(rvb_params), C.POINTER(rvb_batch)]
	lib.rvb_iou_batch_run.argtypes = [C.c_void_p, C.POINTER(rvb_iou_batch)]
	lib.rvb_iou_mean_from_acc.argtypes = [C.POINTER(C.c_uint64), C.c_int64]
	lib.rvb_iou_mean_from_acc.restype = C.c_double
	lib.rvb_debug_cluster_labels.argtypes = [C.c_void_p, C.POINTER(rvb_params), C.c_void_p, C.c_int32, C.c_int32,
											C.c_void_p, C.POINTER(C.c_int32)]
	lib.rvb_debug_smooth_series.argtypes = [C.c_void_p, C.POINTER(rvb_params), C.c_void_p, C.c_int32, C.c_double,
											C.c_void_p, C.c_void_p]
	lib.rvb_ctx_last_iou_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
	lib.rvb_crop_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
									C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
	lib.rvb_ctx_last_host_us.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
	lib.rvb_format_boxes_txt.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
	lib.rvb_debug_butter.argtypes = [C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
	lib.rvb_parse_boxes_txt.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
	_lib = lib
	return lib


def debug_butter(order, wn):
	"""(b, a, zi, chunked_ok) of the low-pass filter the library designs for scipy.signal.butter(order, wn)."""
	lib = load_library()
	b, a, zi = np.zeros(order + 1), np.zeros(order + 1), np.zeros(order)
	ok = C.c_int32()
	check(lib.rvb_debug_butter(order, float(wn), b.ctypes.data, a.ctypes.data, zi.ctypes.data, C.byref(ok)))
	return b, a, zi, bool(ok.value)


def format_boxes_txt(boxes):
	"""[F, 4] boxes -> the bytes of a result file, '%d,%d,%d,%d\\n' per frame (smartVidCrop.py:2783-2785)."""
	lib = load_library()
	bb = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
	buf = C.create_string_buffer(max(1, 48 * len(bb)))
	n = C.c_int64()
	check(lib.rvb_format_boxes_txt(bb.ctypes.data, len(bb), buf, len(buf), C.byref(n)))
	return buf.raw[:n.value]


def parse_boxes_txt(text):
	"""The text of a result / annotation file -> int32 [F, 4], as retargetvid_eval.py:152-159 reads it; a line that
	int() or the indexing would reject raises RvbError naming the line."""
	lib = load_library()
	data = text.encode() if isinstance(text, str) else bytes(text)
	cap = data.count(b'\n') + data.count(b'\r') + 1
	out = np.empty((cap, 4), dtype=np.int32)
	n = C.c_int64()
	check(lib.rvb_parse_boxes_txt(data, len(data), out.ctypes.data, cap, C.byref(n)))
	return out[:n.value].copy()


def check(code):
	if code != RVB_OK:
		raise RvbError(code, load_library().rvb_last_error().decode('utf-8', 'replace'))


def params_from_crop_params(CP, cvrg_window='reference', np_int=False):
	"""crop_params dict (sc_init_crop_params keys) -> rvb_params."""
	p = rvb_params()
	# `smaps < t` on uint8 values with a fractional t keeps exactly the values >= ceil(t) (smartVidCrop.py:1057)
	p.t_threshold = int(-(-float(CP['t_threshold']) // 1))
	p.clust_filt = 1 if CP['clust_filt'] else 0
	p.hdbscan_min = int(CP['hdbscan_min'])
	p.hdbscan_min_samples = 0 if CP['hdbscan_min_samples'] is None else int(CP['hdbscan_min_samples'])
	p.select_sum = int(CP['select_sum'])
	p.op_close = 1 if CP['op_close'] else 0
	p.com_km = 1 if CP['com_km'] else 0
	p.t_border = int(CP['t_border'])
	p.loess_filt = 1 if CP['loess_filt'] else 0
	p.loess_degree = int(CP['loess_degree'])
	p.lp_filt = 1 if CP['lp_filt'] else 0
	p.lp_order = int(CP['lp_order'])
	p.shift_time = int(CP['shift_time'])
	p.exit_on_low_cvrg = 1 if CP['exit_on_low_cvrg'] else 0
	p.cvrg_window = 1 if cvrg_window == 'crop' else 0
	p.resize_type = int(CP['resize_type'])
	p.focus_stability = 1 if CP['focus_stability'] else 0
	p.min_d_jump = int(CP['min_d_jump'])
	p.skip = int(CP['skip'])
	p.np_int_compat = 1 if np_int else 0
	p.foces_stab_t = float(CP['foces_stab_t'])
	p.foces_stab_s = float(CP['foces_stab_s'])
	p.loess_w_secs = float(CP['loess_w_secs'])
	p.lp_cutoff = float(CP['lp_cutoff'])
	p.resize_factor = float(CP['resize_factor'])
	p.t_cvrg = float(CP['t_cvrg'])
	return p


class Context(object):
	"""One per device; owns the library's workspace and stream."""

	def __init__(self, device=0):
		self.lib = load_library()
		h = C.c_void_p()
		check(self.lib.rvb_ctx_create(int(device), C.byref(h)))
		self.handle = h
		self.device = int(device)

	def close(self):
		if self.handle is not None and self.handle.value:
			self.lib.rvb_ctx_destroy(self.handle)
			self.handle = None

	def __del__(self):
		try:
			self.close()
		except Exception:
			pass

	def set_stream(self, cuda_stream_ptr):
		check(self.lib.rvb_ctx_set_stream(self.handle, C.c_void_p(cuda_stream_ptr)))

	def synchronize(self):
		check(self.lib.rvb_ctx_synchronize(self.handle))

	def launch_count(self):
		return int(self.lib.rvb_ctx_launch_count(self.handle))

	def last_stage_ms(self):
		"""(front, Prim, back, whole map pipeline) CUDA-event ms of the last crop_track call (split pipeline)"""
		out = (C.c_float * 4)()
		check(self.lib.rvb_ctx_last_stage_ms(self.handle, out))
		return [float(x) for x in out]

	def last_map_kernel_ms(self):
		ms = C.c_float()
		n = C.c_int32()
		check(self.lib.rvb_ctx_last_map_kernel_ms(self.handle, C.byref(ms), C.byref(n)))
		return float(ms.value), int(n.value)

	def last_iou_kernel_ms(self):
		ms = C.c_float()
		check(self.lib.rvb_ctx_last_iou_kernel_ms(self.handle, C.byref(ms)))
		return float(ms.value)

	def phase_cycles(self, enable=True):
		out = (C.c_uint64 * 16)()
		check(self.lib.rvb_ctx_phase_cycles(self.handle, 1 if enable else 0, out))
		return [int(v) for v in out]

	def crop_track_batch(self, params, batch):
		check(self.lib.rvb_crop_track_batch(self.handle, C.byref(params), C.byref(batch)))

	def last_host_us(self):
		out = (C.c_double * 6)()
		check(self.lib.rvb_ctx_last_host_us(self.handle, out))
		return [float(v) for v in out]

	def iou_batch(self, batch):
		check(self.lib.rvb_iou_batch_run(self.handle, C.byref(batch)))

	def iou_mean_from_acc(self, lo, hi, n):
		a = (C.c_uint64 * 2)(int(lo), int(hi))
		return float(self.lib.rvb_iou_mean_from_acc(a, int(n)))

	def crop_frames(self, frames, boxes):
		"""frames: uint8 [F, H, W, C] numpy array; boxes: [F, 4] x1,y1,x2,y2 of one size.  Returns uint8 [F, h, w, C]
		(the per-frame crop of the reference's renderer, smartVidCrop.py:1906-1912)."""
		fr = np.ascontiguousarray(frames, dtype=np.uint8)
		bb = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
		F, H, W, Cn = fr.shape
		ow, oh = int(bb[0, 2] - bb[0, 0]), int(bb[0, 3] - bb[0, 1])
		out = np.empty((F, oh, ow, Cn), dtype=np.uint8)
		check(self.lib.rvb_crop_frames(self.handle, fr.ctypes.data, F, H, W, Cn, bb.ctypes.data, oh, ow, out.ctypes.data, RVB_MEM_HOST))
		return out

	def debug_cluster_labels(self, params, map_hw):
		m = np.ascontiguousarray(map_hw, dtype=np.uint8)
		labels = np.empty(RVB_MAX_POINTS, dtype=np.int32)
		n = C.c_int32()
		check(self.lib.rvb_debug_cluster_labels(self.handle, C.byref(params), m.ctypes.data, m.shape[0], m.shape[1],
												labels.ctypes.data, C.byref(n)))
		return labels[:n.value].copy()

	def debug_smooth_series(self, params, series, fr):
		x = np.ascontiguousarray(series, dtype=np.float64)
		lp = np.empty_like(x)
		sm = np.empty_like(x)
		check(self.lib.rvb_debug_smooth_series(self.handle, C.byref(params), x.ctypes.data, len(x), float(fr),
											lp.ctypes.data, sm.ctypes.data))
		return lp, sm
