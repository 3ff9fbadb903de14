"""Shared helpers for the tests: fixture loading and vid_data reconstruction."""
import ast
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

CLIP_NAMES = ['c1_default', 'multishot', 'fr25', 'hd1080', 'constant', 'noise', 'few_points',
			'sumsel_min5', 'noclose_nolp', 'savgol_argmax', 'border', 'empties', 'best_settings', 'best_hd_fr25',
			'shift_deg1', 'resize_nearest', 'skip3_fr24', 'best_npint', 'border_hd_multishot', 'loess_w3_bias', 'resize_area2', 'resize_cubic']

ORACLE_ONLY_NAMES = []   # every reference fixture is also run through the CUDA path


def load_clip_fixture(name):
	z = np.load(os.path.join(GOLDEN, 'clip_%s.npz' % name))
	fx = {k: z[k] for k in z.files}
	sc = fx['in_scalars']
	vd = dict(smaps=fx['in_smaps'], segmentation=fx['in_segmentation'],
			segmentation_sel=fx['in_segmentation_sel'],
			true_inds=[int(v) for v in fx['in_true_inds']],
			inds_to_orig=[int(v) for v in fx['in_inds_to_orig']],
			fr=float(sc[0]), fc=int(sc[1]), fc_sel=int(sc[2]), h_orig=int(sc[3]), w_orig=int(sc[4]),
			h_process=int(sc[5]), w_process=int(sc[6]))
	over = ast.literal_eval(str(fx['over']))
	if over == 'BEST':
		from oracle import sc_oracle
		base = sc_oracle.sc_init_crop_params()
		best = sc_oracle.sc_init_crop_params(use_best_settings=True)
		over = {k: v for k, v in best.items() if base[k] != v}
	ratios = [str(r) for r in fx['ratios']]
	return vd, over, ratios, fx


def fixture_np_int(fx):
	"""True for fixtures generated with numpy's old ``np.int`` alias restored (the behaviour of the numpy the reference
	pins): the oracle and the library are then run with their np_int switch."""
	return bool(int(fx['np_int'])) if 'np_int' in fx else False


def loess_tolerance(cl):
	"""Stated tolerance (process pixels) between an accurate LOESS solve and the
	reference's pinv of the uncentred normal equations, whose own error grows
	with shot length (5e-12 @60, 2e-8 @300, 5e-5 @2000, 3.5e-2 @10000)."""
	return max(1e-9, 4e-8 * (cl / 300.0) ** 4)
