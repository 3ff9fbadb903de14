"""torch custom op over the C ABI: ``torch.ops.retargetvid_b200.crop_track``.

SURVEY.md 8f-3: today the UNISAL output goes ``.cpu()`` -> numpy -> uint8 -> ``[H,W,N]`` host array
(unisal/train.py:852-854,1270-1274, smartVidCrop.py:420) before the crop selection runs on the CPU.  This op consumes
the saliency tensor where the CNN leaves it -- on the device, on the caller's CUDA stream -- either as the float32
log-saliency ``[N,H,W]`` that model.py:494-498 returns (the uint8 post-process of train.py:1270-1274 is then fused in
front of the path) or as uint8 ``[N,H,W]`` / ``[N,H,256]``, and returns the boxes as a device tensor.  Nothing is copied
to the host except the few bytes of per-clip status.

PyTorch is plumbing here (tensor memory, the current stream, the dispatcher); all arithmetic is in the CUDA library.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi

_CTX = {}


def _ctx(device_index, stream_ptr):
	"""One library context (workspace) per (device, CUDA stream): calls on different streams may be in flight together."""
	key = (device_index, int(stream_ptr))
	if key not in _CTX:
		_CTX[key] = _cabi.Context(device_index)
		_CTX[key].set_stream(stream_ptr)
	return _CTX[key]


_IPARAMS = ['t_threshold', 'clust_filt', 'hdbscan_min', 'hdbscan_min_samples', 'select_sum', 'op_close', 'com_km', 't_border',
			'loess_filt', 'loess_degree', 'lp_filt', 'lp_order', 'shift_time', 'exit_on_low_cvrg', 'cvrg_window', 'resize_type',
			'focus_stability', 'min_d_jump', 'skip', 'np_int_compat']
_FPARAMS = ['foces_stab_t', 'foces_stab_s', 'loess_w_secs', 'lp_cutoff', 'resize_factor', 't_cvrg']


def pack_params(CP, cvrg_window='reference', np_int=False):
	"""crop_params dict -> (int list, float list) in the field order of rvb_params."""
	p = _cabi.params_from_crop_params(CP, cvrg_window, np_int)
	return [int(getattr(p, k)) for k in _IPARAMS], [float(getattr(p, k)) for k in _FPARAMS]


def pack_clips(vds):
	"""vid_data dicts -> the host-side metadata tensors of the op: clips int64 [nc, 5] (n_maps, n_frames, n_shots,
	h_orig, w_orig), fr float64 [nc], shots int32 [sum S, 4], true_inds int32 [sum N]."""
	clips = torch.tensor([[int(vd['fc_sel']), int(vd['fc']), len(vd['segmentation']), int(vd['h_orig']), int(vd['w_orig'])]
						for vd in vds], dtype=torch.int64)
	fr = torch.tensor([float(vd['fr']) for vd in vds], dtype=torch.float64)
	shots = torch.from_numpy(np.ascontiguousarray(np.concatenate(
		[np.concatenate([np.asarray(vd['segmentation']).reshape(-1, 2), np.asarray(vd['segmentation_sel']).reshape(-1, 2)], axis=1)
		for vd in vds]), dtype=np.int32))
	tinds = torch.from_numpy(np.ascontiguousarray(np.concatenate([np.asarray(vd['true_inds']) for vd in vds]), dtype=np.int32))
	return clips, fr, shots, tinds


torch.library.define(
	'retargetvid_b200::crop_track',
	'(Tensor maps, Tensor clips, Tensor fr, Tensor shots, Tensor true_inds, float[] ratio_w, float[] ratio_h, '
	'int[] iparams, float[] fparams) -> (Tensor, Tensor, Tensor)')


@torch.library.impl('retargetvid_b200::crop_track', 'CUDA')
def _crop_track_cuda(maps, clips, fr, shots, true_inds, ratio_w, ratio_h, iparams, fparams):
	"""maps: device tensor, float32 [N,H,W] (log-saliency, W = process width) or uint8 [N,H,W'] where W' is the row
	stride, a multiple of 4: 256 means the device-native layout of a 250-pixel-wide map (6 ignored padding bytes per row),
	anything else is taken as the process width itself.
	Returns (boxes int32 [R, sum F, 4] = x1,y1,x2,y2, centres float64 [2, sum N], status int32 [n_clips]), all on the device."""
	if not maps.is_cuda or maps.dim() != 3 or not maps.is_contiguous():
		raise ValueError('maps must be a contiguous CUDA tensor [N, H, W]')
	dev = maps.device.index if maps.device.index is not None else torch.cuda.current_device()
	stream_ptr = torch.cuda.current_stream(maps.device).cuda_stream
	ctx = _ctx(dev, stream_ptr)
	nc = int(clips.shape[0])
	R = len(ratio_w)
	carr = (_cabi.rvb_clip * nc)()
	mo = fo = so = 0
	cl = clips.cpu().tolist()
	frl = fr.cpu().tolist()
	for i in range(nc):
		c = carr[i]
		c.n_maps, c.n_frames, c.n_shots, c.h_orig, c.w_orig = cl[i]
		c.fr = frl[i]
		c.map_offset, c.frame_offset, c.shot_offset = mo, fo, so
		mo += c.n_maps
		fo += c.n_frames
		so += c.n_shots
	if mo != int(maps.shape[0]):
		raise ValueError('clips describe %d maps, the tensor holds %d' % (mo, int(maps.shape[0])))
	shots_h = shots.cpu().contiguous().to(torch.int32)
	tinds_h = true_inds.cpu().contiguous().to(torch.int32)
	p = _cabi.rvb_params()
	for k, v in zip(_IPARAMS, iparams):
		setattr(p, k, int(v))
	for k, v in zip(_FPARAMS, fparams):
		setattr(p, k, float(v))
	b = _cabi.rvb_batch()
	b.n_clips, b.h_process, b.n_ratios = nc, int(maps.shape[1]), R
	b.mem_space = _cabi.RVB_MEM_DEVICE
	if maps.dtype == torch.float32:
		b.maps_kind, b.w_process, b.row_stride = _cabi.RVB_MAPS_F32_NHW, int(maps.shape[2]), 0
	elif maps.dtype == torch.uint8:
		# the process width of a 16:9 input is 250 (smartVidCrop.py:252-254); a row stride of 256 carries 6 padding bytes
		ws = int(maps.shape[2])
		b.maps_kind, b.row_stride = _cabi.RVB_MAPS_U8_NHW, ws
		b.w_process = 250 if ws == 256 else ws
		if ws % 4:
			raise ValueError('uint8 maps need a row stride that is a multiple of 4 (pad the rows, e.g. to 256)')
	else:
		raise ValueError('maps must be float32 (log-saliency) or uint8')
	for r in range(R):
		b.ratio_w[r], b.ratio_h[r] = float(ratio_w[r]), float(ratio_h[r])
	b.clips = carr
	b.shots = shots_h.data_ptr()
	b.true_inds = tinds_h.data_ptr()
	b.maps = maps.data_ptr()
	boxes = torch.empty((R, fo, 4), dtype=torch.int32, device=maps.device)
	centres = torch.empty((2, mo), dtype=torch.float64, device=maps.device)
	status = torch.empty((nc,), dtype=torch.int32, device=maps.device)
	b.boxes = boxes.data_ptr()
	b.centres = centres.data_ptr()
	b.clip_status = status.data_ptr()
	ctx.crop_track_batch(p, b)
	return boxes, centres, status


@torch.library.register_fake('retargetvid_b200::crop_track')
def _crop_track_fake(maps, clips, fr, shots, true_inds, ratio_w, ratio_h, iparams, fparams):
	nf = int(clips[:, 1].sum())
	return (maps.new_empty((len(ratio_w), nf, 4), dtype=torch.int32), maps.new_empty((2, maps.shape[0]), dtype=torch.float64),
			maps.new_empty((clips.shape[0],), dtype=torch.int32))


def crop_track(maps, vds, CP, out_ratios, cvrg_window='reference'):
	"""Convenience wrapper: device saliency tensor + reference-shaped vid_data dicts (only their metadata is read)
	-> (boxes int32 [R, sum F, 4], centres float64 [2, sum N], status int32 [n_clips]), all on the device, computed on the
	current CUDA stream."""
	clips, fr, shots, tinds = pack_clips(vds)
	ip, fp = pack_params(CP, cvrg_window)
	rw, rh = [], []
	for s in out_ratios:
		a, bb = str(s).split(':')
		rw.append(float(a))
		rh.append(float(bb))
	return torch.ops.retargetvid_b200.crop_track(maps, clips, fr, shots, tinds, rw, rh, ip, fp)
