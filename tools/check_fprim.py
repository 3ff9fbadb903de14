"""GPU check: the lattice-local Prim (fprim_kernel) against the all-pairs Prim (the default; RVB_FRONTIER_PRIM=1 selects the lattice-local one) on bench clips:
boxes, centres, per-map records and filtered maps must be identical.  Prints the stage times and work counters.

    python tools/check_fprim.py [n_clips]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retargetvid_b200 import smartVidCrop as svc  # noqa: E402
from retargetvid_b200 import synth  # noqa: E402
from retargetvid_b200.engine import CropEngine  # noqa: E402


def run(vds, dense, want_filtered):
	if dense:
		os.environ.pop('RVB_FRONTIER_PRIM', None)
	else:
		os.environ['RVB_FRONTIER_PRIM'] = '1'
	eng = CropEngine(0)
	CP = svc.sc_init_crop_params()
	eng.run(vds, CP, ['1:3', '3:1'], detail=True, want_filtered=want_filtered)
	eng.ctx.phase_cycles(True)
	t0 = time.perf_counter()
	res = eng.run(vds, CP, ['1:3', '3:1'], detail=True, want_filtered=want_filtered)
	dt = time.perf_counter() - t0
	cyc = eng.ctx.phase_cycles(False)
	try:
		st = eng.ctx.last_stage_ms()
	except Exception:
		st = None
	eng.close()
	return res, dt, cyc, st


def main():
	n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
	vds = [synth.make_clip(**sp) for sp in synth.config_clips(3, n_clips=n)]
	out = {}
	if '--frontier-only' in sys.argv:      # (for ncu captures)
		res, dt, cyc, st = run(vds, False, False)
		print('frontier call %.1f ms, stage ms %s, prim cycles %d, work %s' % (dt * 1e3, st, cyc[3], cyc[11:16]))
		return 0
	for dense in (True, False):
		res, dt, cyc, st = run(vds, dense, True)
		out[dense] = res
		print('dense' if dense else 'frontier', 'call %.1f ms, stage ms (front, prim, back, all) %s' % (dt * 1e3, st))
		print('  prim cycles %d, work [steps, stalls, far pair-vectors, near updates, batches] %s' % (cyc[3], cyc[11:16]))
	bad = 0
	for i, (a, b) in enumerate(zip(out[True], out[False])):
		ok = (np.array_equal(a.boxes, b.boxes) and np.array_equal(a.dx, b.dx) and np.array_equal(a.dy, b.dy)
			and np.array_equal(a.map_info, b.map_info) and np.array_equal(a.filtered, b.filtered))
		if not ok:
			bad += 1
			print('clip %d differs: boxes %d, maps %d' % (i, int((a.boxes != b.boxes).sum()),
				int((a.filtered != b.filtered).reshape(len(a.filtered), -1).any(axis=1).sum())))
	print('clips %d, differing %d' % (len(vds), bad))
	return 1 if bad else 0


if __name__ == '__main__':
	sys.exit(main())
