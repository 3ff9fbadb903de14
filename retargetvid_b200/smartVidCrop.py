"""Drop-in host API of SmartVidCrop's crop-selection path, running on the B200 library.

Mirrors the public surface of the reference's smartVidCrop.py for this path
(SURVEY.md 8b): ``sc_init_crop_params`` (smartVidCrop.py:132-209),
``smart_vid_crop`` (:2218-2614), ``smart_crop_version`` (:2617-2618), the
``vid_data`` / ``smart_crop_results`` dict keys and the two text file formats
(:2778-2785).  Ingest (video decode, TransNet, UNISAL; :234-836) and rendering
(:1801-2213) are out of scope: ``vid_data`` comes from the reference's own ingest
cache (``temp_path/<video>.pkl``, :2244-2256) or is passed directly.

All arithmetic is done by the CUDA library through ``retargetvid_b200._cabi``;
there is no CPU fallback.
"""
import os
import pickle
import time

import numpy as np

from . import _cabi
from .engine import CropEngine, MultiGpuCropEngine

_ENGINES = {}


def _engine(device=0):
	if device not in _ENGINES:
		_ENGINES[device] = CropEngine(device)
	return _ENGINES[device]


def _multi_engine(devices):
	key = tuple(int(d) for d in devices)
	if key not in _ENGINES:
		_ENGINES[key] = MultiGpuCropEngine(key)
	return _ENGINES[key]


def _pick_engine(device, devices):
	if devices is not None and len(devices) > 1:
		return _multi_engine(devices)
	return _engine(device if not devices else int(devices[0]))


# Initiates the SmartVidCrop method's parameters to the default settings
# (same 31 keys, defaults and preset as smartVidCrop.py:132-209)
def sc_init_crop_params(print_dict=False, use_best_settings=False):
	crop_params = {}
	crop_params['out_ratio'] = "4:5"
	crop_params['max_input_d'] = 250
	crop_params['skip'] = 6
	crop_params['read_batch'] = 2000
	crop_params['resize_factor'] = 1.0
	crop_params['resize_type'] = 1
	crop_params['op_close'] = True
	crop_params['value_bias'] = 1.0
	crop_params['exit_on_spread_sal'] = False
	crop_params['exit_on_low_cvrg'] = False
	crop_params['com_km'] = True
	crop_params['clust_filt'] = True
	crop_params['select_sum'] = 2
	crop_params['min_d_jump'] = 10
	crop_params['focus_stability'] = False
	crop_params['foces_stab_t'] = 60
	crop_params['foces_stab_s'] = 1.5
	crop_params['hdbscan_min'] = 26
	crop_params['hdbscan_min_samples'] = None
	crop_params['shift_time'] = 0
	crop_params['loess_filt'] = 1
	crop_params['loess_w_secs'] = 2
	crop_params['loess_degree'] = 2
	crop_params['lp_filt'] = 1
	crop_params['lp_cutoff'] = 2
	crop_params['lp_order'] = 5
	crop_params['t_sal'] = 40
	crop_params['t_cvrg'] = 0.60
	crop_params['t_threshold'] = 120
	crop_params['t_border'] = -1
	crop_params['t_cut'] = 120
	if use_best_settings:
		crop_params['t_threshold'] = 90
		crop_params['hdbscan_min'] = 5
		crop_params['hdbscan_min_samples'] = 3
		crop_params['min_d_jump'] = 1
		crop_params['resize_factor'] = 4
		crop_params['op_close'] = True
		crop_params['value_bias'] = 1.0
		crop_params['select_sum'] = 1
		crop_params['focus_stability'] = True
		crop_params['foces_stab_t'] = 60
		crop_params['foces_stab_s'] = 1.5
		crop_params['t_border'] = -1
		crop_params['lp_filt'] = 1
		crop_params['lp_cutoff'] = 1
		crop_params['lp_order'] = 2
		crop_params['loess_filt'] = 0
	if print_dict:
		for x in crop_params.keys():
			print(x, ':', crop_params[x])
	return crop_params


def smart_crop_version():
	return '1.4.0'


def _check_supported(CP):
	if float(CP['resize_factor']) != 1.0:
		if not float(CP['resize_factor']) > 1.0:
			raise NotImplementedError('resize_factor=%r: up-scaling before the clustering is not built' % CP['resize_factor'])
		if CP['resize_type'] not in (1, 2, 3):
			# (the reference then takes no branch at smartVidCrop.py:1078-1084 and clusters the full-size map)
			raise NotImplementedError('resize_type=%r: bilinear (1), cubic (2) and nearest (3) are built' % (CP['resize_type'],))


def _times_dict(vid_dur, t_map, t_total, ingest_times):
	"""Same keys and '%7.3fs, %6.3f%%' format as sc_all_times (smartVidCrop.py:113-124);
	keys starting with '_' are summed into 'total'."""
	keys = ['_read', '_read_shot_det', '_read_sal_det', '_calc_dest_size', '_border_det', '_check_mean_sal',
			'_thresh', '_clustering', '_check_cvrg', '_center_of_mass', '_center_empty_handle', '_focus_stability',
			'_interpolation', '_smooth', '_bb', '_shift']
	vals = {k: 0.0 for k in keys}
	for k in ('_read', '_read_shot_det', '_read_sal_det'):
		vals[k] = float(ingest_times.get(k, 0.0))
	vals['_clustering'] = t_map                      # fused map kernel: threshold .. centre of mass
	vals['_smooth'] = max(t_total - t_map, 0.0)      # everything after it, incl. copies
	out = {}
	sum_t = 0.0
	sum_p = 0.0
	for k in keys:
		sum_t += vals[k]
		sum_p += (vals[k] / vid_dur) * 100.0
		out[k] = '%7.3fs, %6.3f%%' % (vals[k], (vals[k] / vid_dur) * 100.0)
	for k in ('read_init', 'read_tidy', 'clust_init', 'render', 'copy_sound'):
		v = float(ingest_times.get(k, 0.0)) if k.startswith('read') else 0.0
		out[k] = '%7.3fs, %6.3f%%' % (v, (v / vid_dur) * 100.0)
	out['total'] = '%7.3fs, %6.3f%%' % (sum_t, sum_p)
	return out


def _fill_vd(VD, CP, res, ratio_index, want_smaps):
	"""Writes the stage outputs into the vid_data dict with the reference's keys."""
	d = res.dims[ratio_index]
	VD['conversion_mode'] = int(d[0])
	VD['w_final'] = int(d[1])
	VD['h_final'] = int(d[2])
	VD['fbb_w'] = int(d[3])
	VD['fbb_h'] = int(d[4])
	VD['border_t'], VD['border_b'], VD['border_l'], VD['border_r'] = int(d[5]), int(d[6]), int(d[7]), int(d[8])
	VD['mean_sal_scores'] = np.array(res.map_scores)
	VD['mean_sal_score'] = res.mean_sal_score if CP['exit_on_spread_sal'] else None
	VD['mean_cvrg_score'] = float(res.cvrg_scores[ratio_index]) if CP['exit_on_low_cvrg'] else None
	dx = np.asarray(res.dx, dtype=np.float64).tolist()
	dy = np.asarray(res.dy, dtype=np.float64).tolist()
	VD['dx'] = dx
	VD['dy'] = dy
	VD['dxnf'] = np.asarray(res.dxnf, dtype=np.float64).tolist()
	VD['dynf'] = np.asarray(res.dynf, dtype=np.float64).tolist()
	jumps = np.asarray(res.jumps, dtype=np.float64).tolist()
	VD['jumps'] = [255 if v == 255.0 else v for v in jumps]
	VD['jumps_inds'] = [i for i in range(1, len(jumps)) if CP['focus_stability'] and jumps[i] < CP['foces_stab_t']]
	s = res.series
	VD['dxi'] = s[0].tolist()
	VD['dyi'] = s[1].tolist()
	VD['dxl'] = s[2].tolist()
	VD['dyl'] = s[3].tolist()
	# sc_compute_bb truncates dxs/dys in place to original-size ints (smartVidCrop.py:995-999)
	scale_h = float(VD['h_process']) / float(VD['h_orig'])
	scale_w = float(VD['w_process']) / float(VD['w_orig'])
	VD['dxs'] = [int(v / scale_w) for v in s[4]]
	VD['dys'] = [int(v / scale_h) for v in s[5]]
	ts = []
	for seg in np.asarray(VD['segmentation']):
		ts += list(range(int(seg[1]) - int(seg[0]) + 1))
	VD['ts'] = ts
	VD['bbs'] = np.asarray(res.boxes[ratio_index]).tolist()      # python ints, [x1, y1, x2, y2] per frame
	if want_smaps and res.filtered_hwn is not None:
		VD['smaps'] = res.filtered_hwn         # [H, W, N] as the reference leaves it (transposed on the GPU)
	elif want_smaps and res.filtered is not None:
		VD['smaps'] = np.ascontiguousarray(np.transpose(res.filtered, (1, 2, 0)))
	return VD


def _pad_decision(CP, res, ratio_index):
	"""The two exits of smart_vid_crop that replace smart-cropping by padding (smartVidCrop.py:2310-2320, :2380-2395),
	with the decisions of SURVEY.md Appendix B-2 / B-3."""
	do_pad = False
	if CP['exit_on_spread_sal'] and res.mean_sal_score > CP['t_sal']:
		do_pad = True   # Appendix B-3: compare mean_sal_score (the reference reads an unset key)
	if CP['exit_on_low_cvrg'] and float(res.cvrg_scores[ratio_index]) < CP['t_cvrg']:
		do_pad = True
	return do_pad


def _results_dict(VD, CP, do_pad, t_dict):
	"""smart_crop_results with the reference's keys in the reference's order (smartVidCrop.py:2374,2544,2552,2581-2610)."""
	r = {}
	r['cuts_clust'] = 0
	r['result'] = 'padded' if do_pad else 'smart cropped'
	r['info'] = ' (%dx%d)->(%dx%d)->(%dx%d)->(%dx%d)\n' % \
		(VD['h_orig'], VD['w_orig'], VD['h_process'], VD['w_process'],
		VD['h_final'], VD['w_final'], VD['fbb_h'], VD['fbb_w'])
	params_string = ''
	for cpk in CP.keys():
		params_string += ' %-18s : %s\n' % (cpk, str(CP[cpk]))
	r['params'] = params_string
	r['mean_sal_score'] = VD['mean_sal_score']
	r['mean_sal_score_t'] = CP['t_sal']
	r['coverage_score'] = VD['mean_cvrg_score']
	r['coverage_score_t'] = CP['t_cvrg']
	for k in t_dict.keys():
		if k.startswith('_'):
			r['t_' + k] = t_dict[k]
	for k in t_dict.keys():
		if not k.startswith('_'):
			r['t_' + k] = t_dict[k]
	return r


_PAD_DROPS = ('bbs', 'dx', 'dy', 'dxnf', 'dynf', 'dxi', 'dyi', 'dxl', 'dyl', 'dxs', 'dys', 'ts')


def smart_vid_crop_batch(vid_datas, CP=None, out_ratios=None, device=0, detail=True, want_filtered=False,
						cvrg_window='reference', devices=None, raise_on_clip_error=False):
	"""Batched entry point: many videos x many target ratios in one pass.

	vid_datas: list of vid_data dicts (ingest output, smartVidCrop.py:480-489).
	out_ratios: list of 'a:b' strings (default [CP['out_ratio']]).
	devices: list of CUDA device indices: the videos are sharded per video over these GPUs (one context and host thread
	per device, longest first by map count, results gathered on the host in input order) -- the serial loop over videos
	of the reference's driver (smartVidCrop.py:2722-2726) run in parallel.  Default: the single `device`.
	Returns a list (per video) of engine.ClipResult; boxes[r] is the [fc, 4] int32 array of x1,y1,x2,y2 for
	out_ratios[r].  A video that fails on its own (ClipResult.status != 0: more salient pixels in a map than
	RVB_MAX_POINTS, or no salient pixel at all) does not take the others down unless raise_on_clip_error is set.
	"""
	if CP is None:
		CP = sc_init_crop_params()
	_check_supported(CP)
	if out_ratios is None:
		out_ratios = [CP['out_ratio']]
	return _pick_engine(device, devices).run(vid_datas, CP, list(out_ratios), detail=detail, want_filtered=want_filtered,
											cvrg_window=cvrg_window, raise_on_clip_error=raise_on_clip_error)


def smart_vid_crop(video_path, CP=None,
				demo_fn='', final_vid_fn='', plots_fn='',
				frames_dir='', temp_path=None,
				verbose=False, save_vid=True,
				callback_progress=None, callback_session=None, callback_status=None,
				copy_sound=False, vid_data=None, device=0, cvrg_window='reference'):
	"""Same signature and return value as the reference (smartVidCrop.py:2218-2223,2614):
	returns (VD, smart_crop_results).  Extra keyword arguments: ``vid_data`` (skip the
	pickle cache), ``device``, ``cvrg_window``."""
	smart_crop_results = {}
	if CP is None:
		CP = sc_init_crop_params()
	_check_supported(CP)
	if save_vid and (final_vid_fn or demo_fn or plots_fn or frames_dir):
		# rendering (sc_renderer / sc_render_padded, smartVidCrop.py:1801-2213) is out of scope: asking for an output
		# video is an error; the reference's default save_vid=True with no output file name set has nothing to write
		# and runs the crop selection only
		raise NotImplementedError('rendering (sc_renderer / sc_render_padded) is out of scope; call with save_vid=False')

	VD = vid_data
	ingest_times = {}
	if VD is None:
		# the reference's ingest cache: <temp_path>/<video file name without extension>.pkl
		vid_fn = os.path.basename(video_path).split('.')[0]
		cands = []
		if temp_path is not None:
			cands.append(os.path.join(temp_path, vid_fn + '.pkl'))
		if str(video_path).endswith('.pkl'):
			cands.append(video_path)
		for fn in cands:
			if os.path.isfile(fn):
				with open(fn, 'rb') as fp:
					VD = pickle.load(fp)
				break
		if VD is None:
			raise NotImplementedError('ingest (decode + TransNet + UNISAL) is out of scope: provide the reference\'s '
									'vid_data pickle in temp_path or pass vid_data=')
	if 'smaps' not in VD:
		raise ValueError('vid_data pickle without saliency maps')
	ingest_times = dict(VD.get('times', {}))

	if (callback_status is not None) and (callback_session is not None):
		callback_status(callback_session, 'sc', 'SC PROCESSING', 'smart-cropping main process')

	VD['segm_backup'] = np.asarray(VD['segmentation']).copy()
	t0 = time.perf_counter()
	eng = _engine(device)
	res = eng.run([VD], CP, [CP['out_ratio']], detail=True, want_filtered='hwn', cvrg_window=cvrg_window,
				raise_on_clip_error=False)[0]
	t_total = time.perf_counter() - t0
	t_map = eng.ctx.last_map_kernel_ms()[0] / 1000.0
	if res.status == _cabi.RVB_ERR_NO_CENTRES:
		# the reference reaches float(None) in interp_handler (smartVidCrop.py:1533)
		raise TypeError("float() argument must be a string or a real number, not 'NoneType' (no non-empty saliency map)")
	if res.status != _cabi.RVB_OK:
		raise _cabi.RvbError(res.status, 'a saliency map has more than %d salient pixels' % _cabi.RVB_MAX_POINTS)

	do_pad = _pad_decision(CP, res, 0)
	VD = _fill_vd(VD, CP, res, 0, want_smaps=True)
	if do_pad:
		# Appendix B-2: the reference's pad path raises KeyError('dx'); return the state its
		# caller anticipates ("Bounding boxes are not available", smartVidCrop.py:2788-2790)
		for k in _PAD_DROPS:
			VD.pop(k, None)
	t_dict = _times_dict(VD['fc'] / VD['fr'], t_map, t_total, ingest_times)
	smart_crop_results.update(_results_dict(VD, CP, do_pad, t_dict))
	if verbose:
		print(' Times::')
		for k, v in t_dict.items():
			print('   %-21s : %s' % (k, v))
	# (the reference ends with gc.collect() to drop its frame buffers, smartVidCrop.py:2612; nothing of that size lives here
	# and a full collection costs more than the whole GPU pass of a clip)
	return VD, smart_crop_results


def crop_frames(frames, bbs, device=0):
	"""The per-frame crop of the reference's renderer (sc_renderer, smartVidCrop.py:1906-1912) on the GPU:
	frames uint8 [F, h_orig, w_orig, C], bbs = VD['bbs'] -> uint8 [F, fbb_h, fbb_w, C] with
	out[f] == frames[f][y1:y2, x1:x2, :].  Decoding and encoding the video stay with the caller (out of scope)."""
	return _engine(device).ctx.crop_frames(frames, bbs)


def write_result_files(results_out, suffix, vid_data, info_dict):
	"""The two text outputs of the reference's driver (smartVidCrop.py:2778-2785):
	<suffix>_info.txt (key:value lines) and <suffix>.txt (x1,y1,x2,y2 per frame)."""
	os.makedirs(results_out, exist_ok=True)
	with open(os.path.join(results_out, suffix + '_info.txt'), 'w') as stfp:
		for k in info_dict.keys():
			stfp.write(k + ':' + str(info_dict[k]) + '\n')
	if 'bbs' not in vid_data:
		return      # padded: "Bounding boxes are not available" (smartVidCrop.py:2788-2790)
	with open(os.path.join(results_out, suffix + '.txt'), 'wb') as bbfp:
		bbfp.write(_cabi.format_boxes_txt(vid_data['bbs']))      # '%d,%d,%d,%d\n' per frame


def process_pickles(pickle_paths, results_out_top, aspect_ratios_to_test=('1:3', '3:1'), crop_params=None,
					test_name='default_config', device=0, devices=None, cvrg_window='reference'):
	"""The reference's batch driver (smartVidCrop.py:2722-2785) over ingest pickles: for every
	(aspect ratio, video) writes results/<test_name>/<vid>_<a>-<b>.txt and _info.txt with the keys and the padding
	decisions of smart_vid_crop.  All ratios of a video are evaluated from one GPU pass; devices=[...] shards the
	videos over several GPUs.  Returns the per-video ClipResults (status != 0: that video failed, nothing written)."""
	CP = sc_init_crop_params() if crop_params is None else dict(crop_params)
	_check_supported(CP)
	vds = []
	names = []
	for pth in pickle_paths:
		with open(pth, 'rb') as fp:
			vds.append(pickle.load(fp))
		names.append(os.path.basename(pth).split('.')[0])
	t0 = time.perf_counter()
	eng = _pick_engine(device, devices)
	results = eng.run(vds, CP, list(aspect_ratios_to_test), detail=True, cvrg_window=cvrg_window, raise_on_clip_error=False)
	t_total = time.perf_counter() - t0
	t_map = eng.ctx.last_map_kernel_ms()[0] / 1000.0 if hasattr(eng, 'ctx') else 0.0
	results_out = os.path.join(results_out_top, str(test_name))
	for vd, name, res in zip(vds, names, results):
		if res.status != _cabi.RVB_OK:
			# this video failed on its own (the reference would have raised inside its loop); the others are written
			continue
		share = float(vd['fc_sel']) / float(sum(v['fc_sel'] for v in vds))
		for r, orp in enumerate(aspect_ratios_to_test):
			cp = dict(CP)
			cp['out_ratio'] = orp
			do_pad = _pad_decision(cp, res, r)
			VD = _fill_vd(dict(vd), cp, res, r, want_smaps=False)
			t_dict = _times_dict(VD['fc'] / VD['fr'], t_map * share, t_total * share, dict(vd.get('times', {})))
			info = _results_dict(VD, cp, do_pad, t_dict)
			if do_pad:
				for k in _PAD_DROPS:
					VD.pop(k, None)
			write_result_files(results_out, name + '_' + str(orp.replace(':', '-')), VD, info)
	return results
