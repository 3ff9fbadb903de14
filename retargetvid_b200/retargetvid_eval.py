"""RetargetVid evaluator on the B200 library: same inputs, CSV and numbers as the
reference's retargetvid_eval.py (annotations/annotator_*.zip or directories,
results/<run>/<vid>_<ar>.txt -> eval_current.txt).

    python -m retargetvid_b200.retargetvid_eval [results_dir] [--annotations DIR]

The per-frame IoU and the exact per-video sums are computed by the CUDA kernel
(rvb_iou_batch_run); the host only parses text, and takes the means of 200
per-video values per annotator with statistics.mean as the reference does
(retargetvid_eval.py:243-246).
"""
import os
import statistics
import sys
import time
import zipfile

import numpy as np

from . import _cabi

VID_INDS = list(range(1, 101)) + list(range(601, 701))
ARS = ['1-3', '3-1']


def _parse_boxes(text):
	"""One box per line, int(c[0]) .. int(c[3]) of line.split(',') (retargetvid_eval.py:152-159): the library's parser."""
	return _cabi.parse_boxes_txt(text)


def load_annotations(annotations_dir, n_users=6):
	"""annots[user][ar][vid] -> int32 [frames, 4]; reads annotator_<u>/ or annotator_<u>.zip
	(retargetvid_eval.py:45-82) without extracting next to the script."""
	annots = []
	for u in range(1, n_users + 1):
		d = os.path.join(annotations_dir, 'annotator_%d' % u)
		z = None
		if not os.path.isdir(d):
			zp = os.path.join(annotations_dir, 'annotator_%d.zip' % u)
			if not os.path.isfile(zp):
				raise FileNotFoundError('"annotator_%d" directory and zip not found' % u)
			z = zipfile.ZipFile(zp)
		per_ar = {}
		for ar in ARS:
			per_ar[ar] = {}
			for v in VID_INDS:
				fn = '%03d_%s.txt' % (v, ar)
				if z is not None:
					text = z.read('annotator_%d/%s' % (u, fn)).decode()
				else:
					with open(os.path.join(d, fn)) as fp:
						text = fp.read()
				per_ar[ar][v] = _parse_boxes(text)
		annots.append(per_ar)
	return annots


def evaluate_arrays(ctx, method, annots, frame_counts, want_frame_iou=False):
	"""method[vid] -> [F,4] boxes; annots[user][vid] -> [F,4]; frame_counts[vid] -> int.
	Returns (vid_iou [V][U] exactly rounded means, frame_iou or None, vids)."""
	vids = sorted(method.keys())
	U = len(annots)
	V = len(vids)
	offs = np.zeros(V + 1, dtype=np.int64)
	neval = np.zeros((V, U), dtype=np.int32)
	for i, v in enumerate(vids):
		# the reference stops, per annotator, at the first frame either list lacks (retargetvid_eval.py:163-179)
		for u in range(U):
			n = min(int(frame_counts[v]), len(method[v]), len(annots[u][v]))
			if n < 1:
				raise statistics.StatisticsError('mean requires at least one data point')
			neval[i, u] = n
		offs[i + 1] = offs[i] + int(neval[i].max())
	NF = int(offs[V])
	mb = np.zeros((NF, 4), dtype=np.int32)
	ab = np.zeros((U, NF, 4), dtype=np.int32)
	for i, v in enumerate(vids):
		mb[offs[i]:offs[i] + int(neval[i].max())] = method[v][:int(neval[i].max())]
		for u in range(U):
			ab[u, offs[i]:offs[i] + neval[i, u]] = annots[u][v][:neval[i, u]]
	acc = np.zeros((V, U, 2), dtype=np.uint64)
	fiou = np.empty((U, NF), dtype=np.float64) if want_frame_iou else None
	b = _cabi.rvb_iou_batch()
	b.n_videos = V
	b.n_users = U
	b.mem_space = _cabi.RVB_MEM_HOST
	b.frame_offset = offs.ctypes.data
	b.n_eval = None
	b.n_eval_user = neval.ctypes.data
	b.method_boxes = mb.ctypes.data
	b.annot_boxes = ab.ctypes.data
	b.frame_iou = fiou.ctypes.data if want_frame_iou else None
	b.acc = acc.ctypes.data
	# malformed boxes (x2 < x1, empty union) make the call fail loudly: RvbError(RVB_ERR_INVALID)
	ctx.iou_batch(b)
	vid_iou = [[ctx.iou_mean_from_acc(acc[i, u, 0], acc[i, u, 1], int(neval[i, u])) for u in range(U)] for i in range(V)]
	return vid_iou, fiou, vids


def evaluate_run(ctx, method_by_ar, annots, frame_counts):
	"""One run: {ar: {vid: boxes}} -> {ar: dict(per_user, worst, best, mean)} (x100 like the CSV)."""
	out = {}
	for ar in method_by_ar.keys():
		ann = [a[ar] for a in annots]
		vid_iou, _, vids = evaluate_arrays(ctx, method_by_ar[ar], ann, frame_counts)
		users = [statistics.mean([vid_iou[i][u] for i in range(len(vids))]) for u in range(len(annots))]
		out[ar] = dict(per_user=users, worst=min(users) * 100, best=max(users) * 100,
					mean=statistics.mean(users) * 100)
	return out


def _scrape_info(fn, stats):
	"""retargetvid_eval.py:197-222"""
	if not os.path.isfile(fn):
		return
	with open(fn) as fp:
		info_raw = fp.read().splitlines()
	for k in info_raw:
		if '%' in k:
			key = k.split(':')[0].strip().lower()
			val = float(k.split(',')[1].replace('%', '').strip())
			stats.setdefault(key, []).append(val)
		else:
			for tag in ('cuts_clust', 'cuts_extra', 'no_extra_cuts'):
				if tag + ':' in k:
					stats.setdefault(tag, []).append(int(k.split(':')[1].strip()))


def main(argv=None):
	argv = list(sys.argv[1:] if argv is None else argv)
	root_path = os.getcwd()
	annotations_dir = os.path.join(root_path, 'annotations')
	if '--annotations' in argv:
		i = argv.index('--annotations')
		annotations_dir = argv[i + 1]
		del argv[i:i + 2]
	runs_folder = argv[0] if argv else 'results'
	print(' Read results from "%s" directory' % runs_folder)
	print(' loading annotations...')
	annots = load_annotations(annotations_dir)
	print(' ...found annotations from %d users' % len(annots))
	frame_counts = {v: len(annots[0]['1-3'][v]) for v in VID_INDS}
	runs = sorted(os.path.split(f)[-1] for f in os.scandir(runs_folder) if f.is_dir())
	# validity report of the reference (retargetvid_eval.py:102-121): counted, never rejected
	print(' Checking runs validity...')
	for run in runs:
		file_errors_count = 0
		frame_count_errors_count = 0
		for v in VID_INDS:
			for ar in ARS:
				fn = os.path.join(runs_folder, run, '%03d_%s.txt' % (v, ar))
				if not os.path.isfile(fn):
					file_errors_count += 1
				else:
					with open(fn) as fp:
						n_lines = len(fp.read().splitlines())
					if abs(frame_counts[v] - n_lines) > 1:
						frame_count_errors_count += 1
		print(' - %-30s (file errors:%d + frame count errors:%d)' % (run, file_errors_count, frame_count_errors_count))
	print(' valid runs::')
	for run in runs:
		print(' - %s' % run)
	ctx = _cabi.Context(0)
	lines = []
	header = ('%-36s' + ',%-6s' * 23) % ('Method', 'Worst', 'Best', 'Mean', 'ttm', 'tta', 'tcm', 'tca', 'ccm', 'cca',
										'ecm', 'eca', 'Worst', 'Best', 'Mean', 'ttm', 'tta', 'tcm', 'tca', 'ccm', 'cca',
										'ecm', 'eca', 'mf')
	print(' Processing runs...')
	for i_run, run in enumerate(runs):
		t0 = time.time()
		missing = 0
		method = {}
		stats = {}
		for ar in ARS:
			method[ar] = {}
			stats[ar] = {}
			for v in VID_INDS:
				fn = os.path.join(runs_folder, run, '%03d_%s.txt' % (v, ar))
				if not os.path.isfile(fn):
					missing += 1
					continue
				with open(fn) as fp:
					method[ar][v] = _parse_boxes(fp.read())
				_scrape_info(os.path.join(runs_folder, run, '%03d_%s_info.txt' % (v, ar)), stats[ar])
		ev = evaluate_run(ctx, method, annots, frame_counts)
		run_name = run.replace('_', ',')
		if 'mt=1.0_rf=' not in run_name:
			run_name = run_name.replace('_mt=1.0', '_mt=1.0_rf=1')
		s = '%-36s,' % (run_name.replace('_', ','))
		for ar in ARS:
			st = stats[ar]

			def mx(k):
				return max(st[k]) if k in st else -1

			def av(k):
				return statistics.mean(st[k]) if k in st else -1
			s += '%05.3f,%05.3f,%05.3f,%05.3f,%05.3f,%05.3f,%05.3f,%05.3f,%05.3f,%05.3f,%05.3f,' % (
				ev[ar]['worst'], ev[ar]['best'], ev[ar]['mean'], mx('t_total'), av('t_total'),
				mx('t__clustering'), av('t__clustering'), mx('cuts_clust'), av('cuts_clust'),
				mx('cuts_extra'), av('cuts_extra'))
		s += '%d' % missing
		lines.append(s)
		print(' %3d/%3d: %s ---> %.3fs' % (i_run + 1, len(runs), run, time.time() - t0))
	with open('eval_current.txt', 'w') as fp:
		print('\n Evaluation:')
		print(header)
		fp.write(header + '\n')
		for s in lines:
			print(s)
			fp.write(s + '\n')
	ctx.close()
	return lines


if __name__ == '__main__':
	main()
