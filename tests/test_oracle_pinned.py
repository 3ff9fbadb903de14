"""Pins the oracle (oracle/) against fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from helpers import CLIP_NAMES, ORACLE_ONLY_NAMES, fixture_np_int, load_clip_fixture
from oracle import sc_oracle


def _nan_eq(a, b, tol):
	a = np.asarray(a, dtype=np.float64)
	b = np.asarray(b, dtype=np.float64)
	assert a.shape == b.shape
	assert np.array_equal(np.isnan(a), np.isnan(b))
	m = ~np.isnan(a)
	if m.any():
		assert np.max(np.abs(a[m] - b[m])) <= tol, np.max(np.abs(a[m] - b[m]))


@pytest.mark.parametrize('name', CLIP_NAMES + ORACLE_ONLY_NAMES)
def test_oracle_matches_reference_fixture(name):
	vd, over, ratios, fx = load_clip_fixture(name)
	for k, r in enumerate(ratios):
		CP = sc_oracle.sc_init_crop_params()
		CP.update(over)
		CP['out_ratio'] = r
		out = sc_oracle.smart_vid_crop_oracle(vd, CP, np_int=fixture_np_int(fx))
		tag = r.replace(':', '-')
		dims = fx['dims_' + tag]
		assert [out['conversion_mode'], out['w_final'], out['h_final'], out['fbb_w'], out['fbb_h']] == list(dims[:5])
		assert list(out['borders']) == list(dims[5:9])
		if k == 0:
			# integer stages: bit-exact
			assert np.array_equal(out['smaps_filtered'], fx['smaps_filtered'])
			# centroid: sklearn KMeans(1) == unweighted mean to ~1e-13 px
			if CP['focus_stability']:
				# the line-sampling means (smartVidCrop.py:1395-1455) and the centres before the freeze
				_nan_eq(out['jumps'], fx['jumps'], 1e-9)
				_nan_eq([np.nan if v is None else v for v in out['dxnf']], fx['dxnf'], 1e-9)
			_nan_eq([np.nan if v is None else v for v in out['dx']], fx['dx'], 1e-9)
			_nan_eq([np.nan if v is None else v for v in out['dy']], fx['dy'], 1e-9)
			_nan_eq(out['dxi'], fx['dxi'], 1e-9)
			_nan_eq(out['dyi'], fx['dyi'], 1e-9)
			_nan_eq(out['dxl'], fx['dxl'], 1e-9)
			_nan_eq(out['dyl'], fx['dyl'], 1e-9)
			_nan_eq(out['dxs'], fx['dxs_pre'], 1e-9)
			_nan_eq(out['dys'], fx['dys_pre'], 1e-9)
		assert np.array_equal(np.array(out['bbs'], dtype=np.int32), fx['bbs_' + tag])
