"""Import the UNMODIFIED reference ``smartVidCrop.py`` in this container.

Only usable where /root/reference exists (the build container), never on the
GPU box: it is used by ``make_golden.py`` to generate the committed fixtures
and by the oracle-pinning tests (skipped when the reference is absent).

The reference imports model code at import time (smartVidCrop.py:22-83), so
seven modules are stubbed (recipe: SURVEY.md Appendix C).  ``hdbscan`` (the
un-vendored hdbscan==0.8.26, README.md:87) is replaced by
``sklearn.cluster.HDBSCAN(algorithm='brute')`` with ``min_samples + 1`` -- see
DESIGN.md "oracle" for why the +1.
"""
import contextlib
import io
import os
import pickle
import sys
import tempfile
import types

import numpy as np

REF_ROOT = os.environ.get('RVB_REFERENCE_ROOT', '/root/reference')


def reference_available():
	return os.path.isfile(os.path.join(REF_ROOT, 'smartVidCrop.py'))


class _SklearnHDBSCANStandIn(object):
	"""Maps the hdbscan==0.8.26 constructor used at smartVidCrop.py:2340-2348
	onto sklearn.cluster.HDBSCAN."""

	def __init__(self, min_cluster_size=5, min_samples=None, metric='euclidean',
				approx_min_span_tree=True, gen_min_span_tree=False,
				cluster_selection_method='eom', core_dist_n_jobs=4,
				allow_single_cluster=False):
		from sklearn.cluster import HDBSCAN
		ms = min_cluster_size if min_samples is None else min_samples
		self._impl = HDBSCAN(min_cluster_size=min_cluster_size, min_samples=ms + 1,
							metric=metric, algorithm='brute',
							cluster_selection_method=cluster_selection_method,
							allow_single_cluster=allow_single_cluster, copy=True)

	def fit_predict(self, X):
		return self._impl.fit_predict(np.asarray(X, dtype=np.float64))


def _install_stubs():
	def mod(name, **attrs):
		m = types.ModuleType(name)
		for k, v in attrs.items():
			setattr(m, k, v)
		sys.modules[name] = m
		return m

	mod('ffmpeg')
	iv = mod('imutils.video', FileVideoStream=object)
	mod('imutils', video=iv)
	plt = mod('matplotlib.pyplot')
	mod('matplotlib', pyplot=plt)

	class _TF(types.ModuleType):
		def __getattr__(self, name):
			if name.startswith('__'):
				raise AttributeError(name)
			return lambda *a, **k: None
	sys.modules['tensorflow'] = _TF('tensorflow')

	class ShotTransNetParams(object):
		pass

	class ShotTransNet(object):
		def __init__(self, params, session=None):
			pass
	mod('transnetv1_handler', ShotTransNetParams=ShotTransNetParams, ShotTransNet=ShotTransNet)
	mod('unisal_handler', init_unisal_for_images=lambda *a, **k: None)
	mod('hdbscan', HDBSCAN=_SklearnHDBSCANStandIn)


_REF = None


def load_reference():
	"""Returns the imported reference module (cached)."""
	global _REF
	if _REF is not None:
		return _REF
	if not reference_available():
		raise RuntimeError('reference not present at %s' % REF_ROOT)
	import torch  # before the stubs: torch's import walks sys.modules
	_install_stubs()
	real_device = torch.device
	paths = [REF_ROOT, os.path.join(REF_ROOT, '3rd_party_libs', 'loess')]
	for p in paths:
		if p not in sys.path:
			sys.path.insert(0, p)
	try:
		# smartVidCrop.py:72 asks for cuda:0; there is no GPU in the build container
		torch.device = lambda *a, **k: real_device('cpu')
		with contextlib.redirect_stdout(io.StringIO()):
			import importlib
			_REF = importlib.import_module('smartVidCrop')
	finally:
		torch.device = real_device
	return _REF


def run_reference(vd, CP, quiet=True):
	"""Runs the reference smart_vid_crop (smartVidCrop.py:2218) on a synthetic
	vid_data dict through its temp_path pickle cache.  Returns (VD, results)."""
	ref = load_reference()
	vd = {k: (np.copy(v) if isinstance(v, np.ndarray) else (list(v) if isinstance(v, list) else v))
		for k, v in vd.items() if not k.startswith('_')}
	with tempfile.TemporaryDirectory() as td:
		ref.vid_fn = 'synthetic'  # global read at smartVidCrop.py:2245
		with open(os.path.join(td, 'synthetic.pkl'), 'wb') as fp:
			pickle.dump(vd, fp)
		out = io.StringIO()
		with contextlib.redirect_stdout(out if quiet else sys.stdout):
			VD, res = ref.smart_vid_crop('synthetic.mp4', dict(CP), temp_path=td, save_vid=False)
	return VD, res


def load_pyloess():
	p = os.path.join(REF_ROOT, '3rd_party_libs', 'loess')
	if p not in sys.path:
		sys.path.insert(0, p)
	import pyloess
	return pyloess
