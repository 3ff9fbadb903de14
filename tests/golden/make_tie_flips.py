"""Enumerates what changes when the UNMODIFIED reference runs under this machine's stock numpy (numpy >= 2.0 sorts
with an AVX-512 / AVX2 SIMD routine) instead of numpy's portable introsort, the only sort in the numpy the reference
pins and the one the committed fixtures, the oracle and the CUDA kernels reproduce (DESIGN.md "tie rule").

    python tests/golden/make_tie_flips.py            # needs /root/reference; writes tests/golden/tie_flips.json

For every listed fixture the reference is run as a user of this box would run it (no NPY_DISABLE_CPU_FEATURES) and
compared with the committed fixture: filtered maps that differ, boxes that differ, max |delta| of a box coordinate.
To show that every flip traces to a clustering tie and to nothing else, the oracle is run in the same process with
``sort='stock'`` (np.argsort of the edge weights as this numpy does it -- the one line that differs): it must
reproduce the stock-numpy reference bit for bit.  Also records the smallest relative gap between the competing
stabilities of every excess-of-mass decision on these clips (the oracle and the kernels sum stabilities in exact
2^-46 fixed point, the libraries in float64; the decisions can only differ when that gap is ~1e-11).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if os.environ.get('NPY_DISABLE_CPU_FEATURES'):
	sys.exit('run without NPY_DISABLE_CPU_FEATURES: this script measures the stock numpy sort')

import numpy as np  # noqa: E402

sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ref_harness  # noqa: E402
from helpers import load_clip_fixture  # noqa: E402
from oracle import hdbscan_port, sc_oracle  # noqa: E402

NAMES = ['c1_default', 'multishot', 'fr25', 'hd1080', 'noise', 'sumsel_min5', 'border', 'loess_w3_bias']


def main():
	ref = ref_harness.load_reference()
	report = {'numpy': np.__version__, 'clips': {}, 'note': 'stock = this numpy\'s default argsort (SIMD dispatch depends on the CPU); '
			'portable = NPY_DISABLE_CPU_FEATURES as in make_golden.py'}
	gaps = []
	for name in NAMES:
		vd, over, ratios, fx = load_clip_fixture(name)
		r = ratios[0]
		tag = r.replace(':', '-')
		CP = ref.sc_init_crop_params()
		CP.update(over)
		CP['out_ratio'] = r
		vd_ref = dict(vd)
		vd_ref['times'] = {'read_init': 0.0, '_read': 0.0, '_read_shot_det': 0.0, '_read_sal_det': 0.0, 'read_tidy': 0.0}
		VD, _ = ref_harness.run_reference(vd_ref, CP)
		maps_stock = np.asarray(VD['smaps'])
		bbs_stock = np.array(VD['bbs'], dtype=np.int32)
		maps_port = fx['smaps_filtered']
		bbs_port = fx['bbs_' + tag]
		dmap = (maps_stock != maps_port).reshape(-1, maps_port.shape[2]).any(axis=0)
		dbox = (bbs_stock != bbs_port).any(axis=1)
		# the oracle with the one line changed
		OCP = sc_oracle.sc_init_crop_params()
		OCP.update(over)
		OCP['out_ratio'] = r

		def cluster_stock(X, CPo=OCP):
			return hdbscan_port.fit_predict(X, CPo['hdbscan_min'], CPo['hdbscan_min_samples'], True, sort='stock')

		def cluster_port(X, CPo=OCP):
			return hdbscan_port.fit_predict(X, CPo['hdbscan_min'], CPo['hdbscan_min_samples'], True, gaps=gaps)
		out = sc_oracle.smart_vid_crop_oracle(vd, OCP, cluster_fn=cluster_stock)
		sc_oracle.smart_vid_crop_oracle(vd, OCP, cluster_fn=cluster_port)
		ok_maps = bool(np.array_equal(out['smaps_filtered'], maps_stock))
		ok_boxes = bool(np.array_equal(np.array(out['bbs'], dtype=np.int32), bbs_stock))
		# how different a flipped map is: pixels whose kept / dropped state changed
		px = [int(((maps_stock[:, :, i] > 0) != (maps_port[:, :, i] > 0)).sum()) for i in np.nonzero(dmap)[0]]
		report['clips'][name] = {
			'ratio': r, 'maps': int(maps_port.shape[2]), 'maps_differing': int(dmap.sum()),
			'boxes': int(len(bbs_port)), 'boxes_differing': int(dbox.sum()),
			'max_abs_box_delta': int(np.abs(bbs_stock.astype(np.int64) - bbs_port).max()),
			'max_pixels_changed_in_a_map': int(max(px) if px else 0),
			'oracle_with_stock_sort_equals_stock_reference': ok_maps and ok_boxes,
		}
		print(name, report['clips'][name])
	g = np.array([x[0] for x in gaps], dtype=np.float64)
	big = np.array([x[1] for x in gaps], dtype=np.float64)
	report['eom_decisions'] = int(len(g))
	report['eom_trivial_zero_ties'] = int(((g == 0) & (big == 0)).sum())      # 0 == 0: exact in float64 as well
	report['eom_nonzero_ties'] = int(((g == 0) & (big > 0)).sum())
	nz = g[g > 0]
	report['eom_min_relative_gap'] = float(nz.min()) if len(nz) else None
	with open(os.path.join(HERE, 'tie_flips.json'), 'w') as fp:
		json.dump(report, fp, indent=1, sort_keys=True)
	print('eom decisions %d, min relative gap of the non-ties %s, trivial 0 == 0 ties %d, non-zero ties %d' % (
		len(g), report['eom_min_relative_gap'], report['eom_trivial_zero_ties'], report['eom_nonzero_ties']))


if __name__ == '__main__':
	main()
