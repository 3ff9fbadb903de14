#!/usr/bin/env python
"""Benchmark of the saliency-map -> crop-track hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--clips C]

A step is one pass of the hot path over one batch of synthetic DHF1K-shaped clips
(BASELINE.json configs[2]: 200 clips x {1:3, 3:1}); with N GPUs every rank gets its own
batch of the same shape (weak scaling, videos are independent, no collective on the data
path).  `value` is device-resident throughput, `e2e` goes through the C ABI with pinned
HOST buffers in the reference's own [H,W,N] layout (H2D + D2H inside the timed region).
Throughput counts SOURCE frames: every frame gets one box per target ratio from one pass,
the ratios do not multiply the count (the CPU arm runs one ratio per pass, as the reference
does).  The same run also measures BASELINE.json configs[4] (`strong_c5`: ONE 2 000-clip
corpus sharded per video over the N ranks, 4 ratios, boxes gathered on rank 0's host).
`--impl reference` times the CPU restatement of the reference's path (the oracle, with
the clustering delegated to scikit-learn's HDBSCAN as in the survey) on the host cores.
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATIOS = ['1:3', '3:1']
METRIC = 'saliency-map->crop frames/sec (source frames; each gets one box per target ratio from one pass)'
UNIT = 'frames/s'
H, W, WPS = 140, 250, 256
ALGO_BYTES_PER_MAP = H * W + 17          # uint8 entry, + the (cx, cy, empty) record
ALGO_BYTES_PER_BOX = 16


def _make(spec):
	from retargetvid_b200 import synth
	return synth.make_clip(**spec)


def make_workload(n_clips, rank, world=1, workload='c3'):
	"""c3 (default): BASELINE.json configs[2], every rank its own n_clips clips (weak scaling).
	c5: configs[4], ONE corpus of n_clips clips sharded per video over the ranks, longest first by map count
	(retargetvid_b200/sharding.py); the map count follows from the frame count alone."""
	from retargetvid_b200 import sharding, synth
	if workload == 'c5':
		specs = synth.config_clips(5, n_clips=n_clips)
		costs = [len(synth.sampling_table(sp['fc'], sp.get('shot_starts', ()), 6)[0]) for sp in specs]
		specs = [specs[i] for i in sharding.my_shard(costs, rank, world)]
	else:
		specs = synth.config_clips(3, n_clips=n_clips, rank=rank)
	procs = min(len(specs), max(1, (os.cpu_count() or 1)))
	if procs > 1:
		with mp.get_context('fork').Pool(procs) as pool:
			return pool.map(_make, specs, chunksize=4)
	return [_make(s) for s in specs]


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (a restatement of the reference's path), all host cores
# ---------------------------------------------------------------------------------------------
def _cpu_one(args):
	import warnings
	warnings.filterwarnings('ignore')
	from oracle import sc_oracle
	from sklearn.cluster import HDBSCAN
	vd, ratio = args
	CP = sc_oracle.sc_init_crop_params()
	CP['out_ratio'] = ratio
	k = CP['hdbscan_min'] if CP['hdbscan_min_samples'] is None else CP['hdbscan_min_samples']
	clusterer = HDBSCAN(min_cluster_size=CP['hdbscan_min'], min_samples=k + 1, metric='sqeuclidean',
						algorithm='brute', cluster_selection_method='eom', allow_single_cluster=True, copy=True)

	def cluster_fn(X):
		return clusterer.fit_predict(np.asarray(X, dtype=np.float64))
	out = sc_oracle.smart_vid_crop_oracle(vd, CP, cluster_fn=cluster_fn)
	return len(out['bbs'])


def cpu_baseline(vds, n_sample, steps=1, warmup=0, budget_s=150.0):
	"""Times the CPU path on a bounded sample: the first n_sample clips x one ratio, one clip
	per process, all cores.  The sample shrinks if steps x warmup passes would exceed the
	time budget.  Returns (frames/s, cores, description, seconds per step)."""
	cores = os.cpu_count() or 1
	sample = [(vd, RATIOS[0]) for vd in vds[:n_sample]]
	procs = min(cores, len(sample))
	with mp.get_context('fork').Pool(procs) as pool:
		t0 = time.perf_counter()
		pool.map(_cpu_one, sample, chunksize=1)          # untimed: imports, first-call costs
		t1 = time.perf_counter() - t0
		passes = steps + max(0, warmup - 1)
		if passes * t1 > budget_s and len(sample) > procs:
			keep = max(procs, int(len(sample) * budget_s / (passes * t1)))
			sample = sample[:keep]
		for _ in range(max(0, warmup - 1)):
			pool.map(_cpu_one, sample, chunksize=1)
		t0 = time.perf_counter()
		for _ in range(steps):
			pool.map(_cpu_one, sample, chunksize=1)
		dt = (time.perf_counter() - t0) / steps
	frames = sum(vd['fc'] for vd, _ in sample)
	maps = sum(vd['fc_sel'] for vd, _ in sample)
	desc = ('first %d clips of the workload x ratio %s (%d frames, %d maps) per step, one clip per process on '
			'%d host cores; oracle/sc_oracle.py with sklearn.cluster.HDBSCAN(brute) as the clustering library'
			% (len(sample), RATIOS[0], frames, maps, procs))
	return frames / dt, procs, desc, dt


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
	def __init__(self, index):
		self.index = index
		self.proc = None
		self.path = None

	def start(self):
		try:
			fd, self.path = tempfile.mkstemp(suffix='.csv')
			os.close(fd)
			q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
				'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
				'clocks_event_reasons.sw_power_cap')
			self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
										'--format=csv,noheader,nounits', '-lms', '20'],
										stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
		except Exception:
			self.proc = None

	def stop(self):
		out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
		if self.proc is None:
			return out
		self.proc.terminate()
		try:
			self.proc.wait(timeout=5)
		except Exception:
			self.proc.kill()
		try:
			rows = [l.strip().split(', ') for l in open(self.path) if l.strip()]
			sm = sorted(float(r[0]) for r in rows if len(r) >= 7)
			if sm:
				out['sm_mhz'] = sm[len(sm) // 2]
				out['sm_max_mhz'] = float(rows[0][1])
				names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
				for k, nme in enumerate(names):
					if any(r[3 + k].strip().lower().startswith('active') for r in rows if len(r) >= 7):
						out['reasons'].append(nme)
				out['samples'] = len(sm)
			os.unlink(self.path)
		except Exception:
			pass
		return out


def bind_near_gpu(index):
	"""Pins this process (and the threads and pinned buffers it creates from now on) to the CPUs NVML reports as local
	to the GPU, so that the e2e uploads do not cross the socket interconnect.  Returns the number of CPUs kept."""
	try:
		import pynvml
		import torch
		pynvml.nvmlInit()
		pr = torch.cuda.get_device_properties(index)
		try:
			bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
			h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
		except Exception:
			h = pynvml.nvmlDeviceGetHandleByIndex(index)
		pynvml.nvmlDeviceSetCpuAffinity(h)
		return len(os.sched_getaffinity(0))
	except Exception:
		return None


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class Workload(object):
	"""One rank's batch: host metadata (as the ABI requires), the maps in pinned host memory in the reference's
	[H,W,N] layout (e2e entry) and in device memory in the native [N][H][256] layout (device-resident entry), and one
	set of output buffers per context in flight."""

	def __init__(self, vds, ratios, nctx, torch, _cabi):
		self.vds, self.ratios, self._cabi = vds, list(ratios), _cabi
		nc = self.nc = len(vds)
		self.R = len(ratios)
		NM = self.NM = sum(v['fc_sel'] for v in vds)
		NF = self.NF = sum(v['fc'] for v in vds)
		NS = sum(len(v['segmentation']) for v in vds)
		self.clips = (_cabi.rvb_clip * nc)()
		self.shots = np.zeros((NS, 4), dtype=np.int32)
		self.tinds = np.zeros(NM, dtype=np.int32)
		mo = fo = so = 0
		for i, vd in enumerate(vds):
			c = self.clips[i]
			c.n_maps, c.n_frames, c.n_shots = vd['fc_sel'], vd['fc'], len(vd['segmentation'])
			c.h_orig, c.w_orig, c.fr = vd['h_orig'], vd['w_orig'], vd['fr']
			c.map_offset, c.frame_offset, c.shot_offset = mo, fo, so
			self.shots[so:so + c.n_shots, 0:2] = vd['segmentation']
			self.shots[so:so + c.n_shots, 2:4] = vd['segmentation_sel']
			self.tinds[mo:mo + c.n_maps] = vd['true_inds']
			mo += c.n_maps
			fo += c.n_frames
			so += c.n_shots
		# e2e input: pinned host memory, reference layout [H,W,N] per clip, packed back to back
		self.host_maps = torch.empty(NM * H * W, dtype=torch.uint8).pin_memory()
		hm = self.host_maps.numpy()
		self.ptrs = (C.c_void_p * nc)()
		off = 0
		for i, vd in enumerate(vds):
			n = vd['fc_sel'] * H * W
			hm[off:off + n] = vd['smaps'].reshape(-1)
			self.ptrs[i] = self.host_maps.data_ptr() + off
			off += n
		self.host_boxes = [torch.empty((self.R, NF, 4), dtype=torch.int32).pin_memory() for _ in range(nctx)]
		# device-resident input: device-native layout uint8 [N][H][256]
		self.dev_maps = torch.zeros((NM, H, WPS), dtype=torch.uint8, device='cuda')
		mo = 0
		for vd in vds:
			n = vd['fc_sel']
			self.dev_maps[mo:mo + n, :, :W] = torch.from_numpy(np.ascontiguousarray(np.transpose(vd['smaps'], (2, 0, 1)))).cuda()
			mo += n
		self.dev_boxes = [torch.empty((self.R, NF, 4), dtype=torch.int32, device='cuda') for _ in range(nctx)]
		torch.cuda.synchronize()
		self.b_dev = [self.batch(True, k) for k in range(nctx)]
		self.b_host = [self.batch(False, k) for k in range(nctx)]

	def batch(self, device_resident, k=0):
		_cabi = self._cabi
		b = _cabi.rvb_batch()
		b.n_clips, b.h_process, b.w_process, b.n_ratios = self.nc, H, W, self.R
		for r, s in enumerate(self.ratios):
			a, bb = s.split(':')
			b.ratio_w[r], b.ratio_h[r] = float(a), float(bb)
		b.clips = self.clips
		b.shots = self.shots.ctypes.data
		b.true_inds = self.tinds.ctypes.data
		if device_resident:
			b.maps_kind, b.mem_space, b.row_stride = _cabi.RVB_MAPS_U8_NHW, _cabi.RVB_MEM_DEVICE, WPS
			b.maps = self.dev_maps.data_ptr()
			b.boxes = self.dev_boxes[k].data_ptr()
		else:
			b.maps_kind, b.mem_space = _cabi.RVB_MAPS_U8_HWN, _cabi.RVB_MEM_HOST
			b.maps = None
			b.clip_maps = self.ptrs
			b.boxes = self.host_boxes[k].data_ptr()
		return b

	def release(self):
		self.host_maps = self.dev_maps = self.host_boxes = self.dev_boxes = None


def _kernel_source_hash():
	"""Hash of the device code of the map pipeline (the kernels profiles/map_kernel_traffic.json was captured on)."""
	import hashlib
	h = hashlib.sha256()
	d = os.path.join(ROOT, 'retargetvid_b200', 'csrc')
	for f in ('map_kernel.cuh', 'prim_kernel.cuh', 'prim_retire.inc'):      # (fprim_kernel.cuh is opt-in and not in the capture)
		h.update(open(os.path.join(d, f), 'rb').read())
	return h.hexdigest()[:16]


def gpu_arm(args, rank, world, local_rank):
	import torch
	from retargetvid_b200 import _cabi
	from retargetvid_b200 import smartVidCrop as svc
	torch.cuda.set_device(local_rank)
	dist = None
	gloo = None
	if world > 1:
		import torch.distributed as dist
		dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
		gloo = dist.new_group(backend='gloo')     # host-side gather of the boxes (no collective on the data path)

	c5_main = args.workload == 'c5'
	ratios = ['1:3', '3:1', '9:16', '4:5'] if c5_main else list(RATIOS)
	vds = make_workload(args.clips, rank, world, args.workload)
	all_cpus = os.sched_getaffinity(0)
	near_cpus = None if args.no_bind else bind_near_gpu(local_rank)

	# NCTX contexts, each with its own stream and workspace: consecutive batches are independent, so
	# they are pipelined (the latency tail of one batch's few very large maps overlaps the next batch)
	NCTX = max(1, args.streams)
	ctxs = [_cabi.Context(local_rank) for _ in range(NCTX)]
	streams = [torch.cuda.Stream() for _ in range(NCTX)]
	for cx, st in zip(ctxs, streams):
		cx.set_stream(st.cuda_stream)
	ctx = ctxs[0]
	torch.cuda.set_stream(streams[0])
	CP = svc.sc_init_crop_params()
	params = _cabi.params_from_crop_params(CP)
	wl = Workload(vds, ratios, NCTX, torch, _cabi)
	nc, R, NM, NF = wl.nc, wl.R, wl.NM, wl.NF

	def barrier():
		if dist is not None:
			dist.barrier()
		torch.cuda.synchronize()

	def reduce_max(ms):
		if dist is not None:
			t = torch.tensor([ms], device='cuda', dtype=torch.float64)
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
			ms = float(t.item())
		return ms

	def launches():
		return sum(cx.launch_count() for cx in ctxs)

	def timed_device(w, steps):
		"""device-resident, asynchronous calls round-robin over the contexts; CUDA events on stream 0
		bracket all streams (they wait for the start event, stream 0 waits for their end events)."""
		barrier()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		l0 = launches()
		e0.record(streams[0])
		for st in streams[1:]:
			st.wait_event(e0)
		for k in range(steps):
			ctxs[k % NCTX].crop_track_batch(params, w.b_dev[k % NCTX])
		for st in streams[1:]:
			ev = torch.cuda.Event()
			ev.record(st)
			streams[0].wait_event(ev)
		e1.record(streams[0])
		barrier()
		return reduce_max(e0.elapsed_time(e1)), launches() - l0

	def timed_host(w, steps, after=None):
		"""end to end: synchronous host-buffer calls, one host thread per context.  `after` (optional) runs on every
		rank after its last call, inside the timed region (the host gather of the sharded workload)."""
		barrier()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		errs = []

		def worker(k):
			try:
				for i in range(k, steps, NCTX):
					ctxs[k].crop_track_batch(params, w.b_host[k])
			except Exception as e:  # pragma: no cover
				errs.append(e)
		t0 = time.perf_counter()
		e0.record(streams[0])
		ths = [threading.Thread(target=worker, args=(k,)) for k in range(NCTX)]
		for t in ths:
			t.start()
		for t in ths:
			t.join()
		e1.record(streams[0])
		if after is not None:
			after()
		torch.cuda.synchronize()
		wall_ms = (time.perf_counter() - t0) * 1e3
		barrier()
		if errs:
			raise errs[0]
		return reduce_max(e0.elapsed_time(e1)), reduce_max(wall_ms)

	def warm(w):
		for _ in range(args.warmup):
			for k in range(NCTX):
				ctxs[k].crop_track_batch(params, w.b_dev[k])
		for _ in range(max(1, args.warmup // 2)):
			for k in range(NCTX):
				ctxs[k].crop_track_batch(params, w.b_host[k])
		torch.cuda.synchronize()
		first = w.dev_boxes[0].cpu().numpy().copy()
		for k in range(NCTX):
			assert np.array_equal(first, w.host_boxes[k].numpy()), 'device-resident and host-buffer paths disagree'
			assert np.array_equal(first, w.dev_boxes[k].cpu().numpy())
		return first

	first_boxes = warm(wl)
	sampler = ClockSampler(local_rank)
	if rank == 0:
		sampler.start()
	ms_dev, launches_n = timed_device(wl, args.steps)
	ms_e2e, _ = timed_host(wl, args.steps)
	clocks = sampler.stop() if rank == 0 else None

	# roofline of the dominant kernels (the map pipeline): separate un-pipelined pass, the library's own CUDA events
	# around the map launches on the launching stream
	barrier()
	map_ms, map_launches = 0.0, 0
	stage_ms = [0.0, 0.0, 0.0, 0.0]
	for _ in range(args.steps):
		ctx.crop_track_batch(params, wl.b_dev[0])
		ms, nl = ctx.last_map_kernel_ms()
		map_ms += ms
		map_launches += nl
		for i, v in enumerate(ctx.last_stage_ms()):
			stage_ms[i] += v / args.steps
	barrier()

	# what the Prim launches executed against what the all-pairs formulation needs (work counters of the kernel)
	prim = None
	if rank == 0:
		info = torch.empty((NM, 4), dtype=torch.int32, device='cuda')
		bb = wl.batch(True, 0)
		bb.map_info = info.data_ptr()
		ctx.phase_cycles(True)
		ctx.crop_track_batch(params, bb)
		cyc = ctx.phase_cycles(False)
		torch.cuda.synchronize()
		npts = info[:, 0].cpu().numpy().astype(np.int64)
		clustered = npts > CP['hdbscan_min'] + 1
		dense_pairs = int((npts[clustered] * (npts[clustered] - 1) // 2).sum())
		steps_w, stalls_w, far_w, near_w, batches_w = (int(v) for v in cyc[11:16])
		prim = {'what': 'Prim launches of one step (the maps of the split pipeline; cut-adjacent chains run in the monolithic kernel): '
						'rvb::prim_kernel = all pairs, registers, dp4a (default) or rvb::fprim_kernel = lattice-local (RVB_FRONTIER_PRIM=1)',
				'kernel': 'fprim_kernel (lattice-local)' if batches_w else 'prim_kernel (all pairs)',
				'clustered_maps': int(clustered.sum()),
				'all_pairs_formulation_pair_updates': dense_pairs,
				'front_ms_per_step': stage_ms[0], 'prim_and_back_ms_per_step': stage_ms[1],
				'map_pipeline_ms_per_step': stage_ms[3],
				'phase_split_sm_cycles': {k: int(v) for k, v in zip(
					['load', 'threshold+compact', 'core distances', 'prim', 'argsort emulation', 'cartesian tree', 'condensed bfs',
					'fall-out', 'eom+labels', 'rebuild+closing', 'results'], cyc[:11])}}
		if batches_w:
			prim.update({'prim_steps': steps_w, 'batches': batches_w, 'steps_per_batch': steps_w / max(1, batches_w), 'stalls': stalls_w,
						'near_key_updates': near_w, 'far_pair_updates': far_w * 32,
						'executed_over_all_pairs': (far_w * 32 + near_w) / max(1, dense_pairs)})
		del info

	# (i) of SURVEY.md H2: the streaming stages alone (threshold, mean saliency, centroid, track, boxes), i.e. the
	# same call with the clustering filter switched off -- this is the part of the path that is HBM-bound
	CP_s = dict(CP)
	CP_s['clust_filt'] = False
	params_s = _cabi.params_from_crop_params(CP_s)
	for _ in range(3):
		ctx.crop_track_batch(params_s, wl.b_dev[0])
	barrier()
	stream_map_ms = 0.0
	es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	es0.record(streams[0])
	for _ in range(args.steps):
		ctx.crop_track_batch(params_s, wl.b_dev[0])
		stream_map_ms += ctx.last_map_kernel_ms()[0]
	es1.record(streams[0])
	barrier()
	stream_ms = es0.elapsed_time(es1)

	# what the link alone allows for the e2e entry: one plain pinned-host -> device copy of a step's maps
	scratch_dev = torch.empty(NM * H * W, dtype=torch.uint8, device='cuda')
	h2d_ms = None
	for _ in range(3):
		barrier()
		eh0, eh1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		eh0.record(streams[0])
		scratch_dev.copy_(wl.host_maps, non_blocking=True)
		eh1.record(streams[0])
		torch.cuda.synchronize()
		t = eh0.elapsed_time(eh1)
		h2d_ms = t if h2d_ms is None else min(h2d_ms, t)
	del scratch_dev

	# the maps produced on the device (where UNISAL leaves them) through the torch custom op on the
	# caller's stream -- the e2e of the real pipeline, without the host link
	dev_prod = None
	if rank == 0:
		from retargetvid_b200 import torch_op
		clips_t, fr_t, shots_t, tinds_t = torch_op.pack_clips(vds)
		ip, fp = torch_op.pack_params(CP)
		rw = [float(s.split(':')[0]) for s in ratios]
		rh = [float(s.split(':')[1]) for s in ratios]
		for _ in range(2):
			out = torch.ops.retargetvid_b200.crop_track(wl.dev_maps, clips_t, fr_t, shots_t, tinds_t, rw, rh, ip, fp)
		torch.cuda.synchronize()
		assert np.array_equal(out[0].cpu().numpy(), first_boxes), 'torch op and C-ABI call disagree'
		# one call in flight: the latency of a step through the op
		t0 = time.perf_counter()
		for _ in range(args.steps):
			out = torch.ops.retargetvid_b200.crop_track(wl.dev_maps, clips_t, fr_t, shots_t, tinds_t, rw, rh, ip, fp)
			first_box = out[0][0, 0].cpu()      # the caller reads a result: synchronises the step
		dt1 = (time.perf_counter() - t0) / args.steps
		# NCTX calls in flight on as many streams (the op keeps one context per stream), every result read back
		outs = [None] * NCTX
		for k in range(NCTX):
			with torch.cuda.stream(streams[k]):
				outs[k] = torch.ops.retargetvid_b200.crop_track(wl.dev_maps, clips_t, fr_t, shots_t, tinds_t, rw, rh, ip, fp)
		torch.cuda.synchronize()
		t0 = time.perf_counter()
		for i in range(args.steps):
			k = i % NCTX
			with torch.cuda.stream(streams[k]):
				if i >= NCTX:
					first_box = outs[k][0][0, 0].cpu()      # the result of the call issued NCTX steps ago
				outs[k] = torch.ops.retargetvid_b200.crop_track(wl.dev_maps, clips_t, fr_t, shots_t, tinds_t, rw, rh, ip, fp)
		for k in range(NCTX):
			with torch.cuda.stream(streams[k]):
				first_box = outs[k][0][0, 0].cpu()
		torch.cuda.synchronize()
		dt = (time.perf_counter() - t0) / args.steps
		dev_prod = {'what': 'torch.ops.retargetvid_b200.crop_track on device-resident uint8 [N,140,256] maps on the caller\'s streams, every '
							'result read back by the host (wall clock of the python calls); value: %d calls in flight, latency_ms: one' % NCTX,
					'value': NF / dt, 'unit': UNIT, 'ms_per_step': dt * 1e3, 'latency_ms_one_call_in_flight': dt1 * 1e3}
		del outs
		del out, first_box

	# stage 6 (IoU evaluation, retargetvid_eval.py:133-194) as its own streaming measurement: the step's frames tiled
	# 16x (synthetic annotator boxes, 6 annotators), device-resident, exact 128-bit accumulation per (video, annotator)
	iou = None
	if rank == 0:
		TILE, U = 16, 6
		nv = nc * TILE
		foff = np.zeros(nv + 1, dtype=np.int64)
		foff[1:] = np.cumsum(np.tile(np.array([v['fc'] for v in vds], dtype=np.int64), TILE))
		nev = np.tile(np.array([v['fc'] for v in vds], dtype=np.int32), TILE)
		nfi = int(foff[-1])
		method = wl.dev_boxes[0][0].repeat(TILE, 1).contiguous()
		g = torch.Generator(device='cuda').manual_seed(7)
		x1 = torch.randint(0, 500, (U, nfi, 1), device='cuda', generator=g, dtype=torch.int32)
		y1 = torch.zeros((U, nfi, 1), device='cuda', dtype=torch.int32)
		annot = torch.cat([x1, y1, x1 + 120, y1 + 360], dim=2).contiguous()
		acc = torch.zeros((nv, U, 2), dtype=torch.int64, device='cuda')
		ib = _cabi.rvb_iou_batch()
		ib.n_videos, ib.n_users, ib.mem_space = nv, U, _cabi.RVB_MEM_DEVICE
		ib.frame_offset = foff.ctypes.data
		ib.n_eval = nev.ctypes.data
		ib.method_boxes, ib.annot_boxes, ib.frame_iou, ib.acc = method.data_ptr(), annot.data_ptr(), None, acc.data_ptr()
		for _ in range(3):
			ctx.iou_batch(ib)
		torch.cuda.synchronize()
		ei0, ei1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		ei0.record(streams[0])
		iou_k_ms = 0.0
		for _ in range(args.steps):
			ctx.iou_batch(ib)
			iou_k_ms += ctx.last_iou_kernel_ms() / args.steps       # (synchronises: the calls do not overlap the host work)
		ei1.record(streams[0])
		torch.cuda.synchronize()
		iou_ms = ei0.elapsed_time(ei1) / args.steps
		iou_bytes = nfi * (16 + 16 * U)
		iou = {'what': 'rvb_iou_batch_run (iou_kernel): %d frames x %d annotators per call, device-resident boxes, one call per step '
						'(includes the per-call frame->video table upload)' % (nfi, U),
				'ious_per_sec': nfi * U / (iou_k_ms / 1e3), 'ms_per_call': iou_ms, 'kernel_ms_per_call': iou_k_ms,
				'algorithmic_bytes_per_call': iou_bytes, 'achieved_gbs': iou_bytes / (iou_k_ms / 1e3) / 1e9,
				'note': 'achieved_gbs / frac are for the kernel (CUDA events around its launch inside the library); ms_per_call adds the host-side '
						'packing and upload of the per-video tables, which bounds back-to-back calls'}
		del method, annot, acc

	# the renderer's per-frame crop (smartVidCrop.py:1906-1912) on device-resident frames: pure data
	# movement, rows of 360 bytes cut out of rows of 1920 bytes at an arbitrary byte offset
	crop = None
	if rank == 0:
		nfr = 1000
		frames_t = torch.randint(0, 256, (nfr, 360, 640, 3), dtype=torch.uint8, device='cuda')
		bx = wl.dev_boxes[0][0, :nfr].cpu().numpy().copy() if not c5_main else None
		if bx is not None and bx.shape[0] == nfr:
			oh, ow = int(bx[0, 3] - bx[0, 1]), int(bx[0, 2] - bx[0, 0])
			out_t = torch.empty((nfr, oh, ow, 3), dtype=torch.uint8, device='cuda')
			lib = _cabi.load_library()

			def crop_call():
				_cabi.check(lib.rvb_crop_frames(ctx.handle, frames_t.data_ptr(), nfr, 360, 640, 3, bx.ctypes.data, oh, ow, out_t.data_ptr(), _cabi.RVB_MEM_DEVICE))
			for _ in range(3):
				crop_call()
			torch.cuda.synchronize()
			ec0, ec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			ec0.record(streams[0])
			for _ in range(args.steps):
				crop_call()
			ec1.record(streams[0])
			torch.cuda.synchronize()
			crop_ms = ec0.elapsed_time(ec1) / args.steps
			f0 = 17
			x1, y1, x2, y2 = (int(v) for v in bx[f0])
			assert torch.equal(out_t[f0], frames_t[f0, y1:y2, x1:x2, :])
			cbytes = 2 * nfr * oh * ow * 3
			crop = {'what': 'rvb_crop_frames (crop_frames_kernel): %d device-resident 640x360x3 frames -> %dx%d crops along the 1:3 track of the step '
							'(read + write of the cropped pixels; includes the per-call upload of the boxes)' % (nfr, ow, oh),
					'ms_per_call': crop_ms, 'algorithmic_bytes_per_call': cbytes, 'achieved_gbs': cbytes / (crop_ms / 1e3) / 1e9}
			del out_t
		del frames_t

	# BASELINE.json configs[0]: ONE 300-frame clip through the drop-in python entry point (host numpy maps in, boxes out)
	c1 = None
	if rank == 0:
		from retargetvid_b200 import synth
		vd1 = synth.make_clip(**synth.config_clips(1)[0])
		CP1 = svc.sc_init_crop_params()
		CP1['out_ratio'] = '1:3'
		lat = []
		for i in range(13):
			t0 = time.perf_counter()
			VD1, _res = svc.smart_vid_crop('c1.mp4', CP1, save_vid=False, vid_data=dict(vd1), device=local_rank)
			lat.append((time.perf_counter() - t0) * 1e3)
		lat = sorted(lat[3:])
		c1 = {'what': 'configs[0]: smart_vid_crop() on one 640x360 clip of %d frames (%d maps), ratio 1:3, wall clock of the python call, '
					'median of 10 after 3 warm-up calls' % (vd1['fc'], vd1['fc_sel']),
				'ms_per_clip': lat[len(lat) // 2], 'frames_per_sec': vd1['fc'] / (lat[len(lat) // 2] / 1e3)}

	tot_frames, tot_maps, tot_clips = world * NF, world * NM, world * nc
	if dist is not None and c5_main:      # ranks hold different shards of one corpus
		t = torch.tensor([NF, NM, nc], device='cuda', dtype=torch.int64)
		dist.all_reduce(t)
		tot_frames, tot_maps, tot_clips = (int(x) for x in t.tolist())
	value = tot_frames * args.steps / (ms_dev / 1e3)
	e2e = tot_frames * args.steps / (ms_e2e / 1e3)

	# ---- BASELINE.json configs[4]: ONE corpus sharded per video over the ranks, 4 ratios, boxes gathered on rank 0 ----
	strong = None
	if not c5_main and args.c5_clips > 0:
		wl.release()
		del wl
		torch.cuda.empty_cache()
		ratios5 = ['1:3', '3:1', '9:16', '4:5']
		vds5 = make_workload(args.c5_clips, rank, world, 'c5')
		w5 = Workload(vds5, ratios5, NCTX, torch, _cabi)
		sizes = torch.tensor([w5.NF, w5.NM, w5.nc], device='cuda', dtype=torch.int64)
		all_sizes = [sizes.clone() for _ in range(world)]
		if dist is not None:
			dist.all_gather(all_sizes, sizes)
		else:
			all_sizes = [sizes]
		nf_all = [int(t[0]) for t in all_sizes]
		corpus_frames, corpus_maps, corpus_clips = (int(sum(int(t[i]) for t in all_sizes)) for i in range(3))
		warm(w5)
		steps5 = max(1, min(args.steps, 3))
		ms5_dev, _ = timed_device(w5, steps5)
		# host gather: every rank's boxes [4][NF_rank][4] end up in rank 0's host memory in one padded buffer per rank
		pad = max(nf_all)
		gather_bufs = [torch.empty((len(ratios5), pad, 4), dtype=torch.int32) for _ in range(world)] if (dist is not None and rank == 0) else None
		send = torch.zeros((len(ratios5), pad, 4), dtype=torch.int32) if dist is not None else None

		def gather_boxes():
			if dist is None:
				return
			send[:, :w5.NF, :] = w5.host_boxes[0]
			dist.gather(send, gather_bufs, dst=0, group=gloo)
		_, wall5 = timed_host(w5, steps5, after=gather_boxes)
		if dist is not None and rank == 0:
			assert np.array_equal(gather_bufs[0][:, :w5.NF, :].numpy(), w5.host_boxes[0].numpy())
		strong = {'workload': 'BASELINE.json configs[4]: one corpus of %d synthetic DHF1K-shaped 640x360 clips (%d frames, %d maps) sharded per '
							'video over %d GPU(s), longest first by map count, x ratios %s; no collective on the data path' % (
								corpus_clips, corpus_frames, corpus_maps, world, ','.join(ratios5)),
				'scaling': 'strong', 'steps': steps5,
				'value': corpus_frames * steps5 / (ms5_dev / 1e3), 'unit': UNIT, 'ms_per_step': ms5_dev / steps5,
				'e2e': {'value': corpus_frames * steps5 / (wall5 / 1e3), 'unit': UNIT, 'ms_per_step': wall5 / steps5,
						'h2d_bytes_per_step': int(corpus_maps * H * W), 'd2h_bytes_per_step': int(len(ratios5) * corpus_frames * 16),
						'gather': 'boxes of every rank gathered into rank 0 host memory (torch.distributed gloo group) inside the timed region; '
								'wall clock between barriers, max over ranks' if dist is not None else 'single rank: the boxes are already in its host memory'},
				'frames_per_rank': nf_all}
		w5.release()
		del w5

	line = None
	if rank == 0:
		peaks = {}
		try:
			peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
		except Exception:
			pass
		peak = float(peaks.get('hbm_gbs', 6650.0))
		algo = NM * ALGO_BYTES_PER_MAP
		achieved = algo * args.steps / (map_ms / 1e3) / 1e9
		traffic, traffic_note = None, 'no ncu capture committed for this build'
		try:
			prof = json.load(open(os.path.join(ROOT, 'profiles', 'map_kernel_traffic.json')))
			if prof.get('kernel_source_hash') == _kernel_source_hash():
				traffic, traffic_note = prof.get('dram_bytes_per_step'), prof.get('source')
			else:
				traffic_note = ('profiles/map_kernel_traffic.json was captured on other kernel sources (hash %s, now %s): not reported'
								% (prof.get('kernel_source_hash'), _kernel_source_hash()))
				sys.stderr.write('bench.py: ' + traffic_note + '\n')
		except Exception:
			pass
		os.sched_setaffinity(0, all_cpus)      # the CPU arm gets every host core back
		cpu_v, cpu_cores, cpu_desc, _ = cpu_baseline(vds, args.cpu_sample) if (world == 1 and args.cpu_sample > 0) else (None, None, None, None)
		stream_achieved = NM * ALGO_BYTES_PER_MAP * args.steps / (stream_ms / 1e3) / 1e9
		line = {
			'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
			'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'strong' if c5_main else 'weak', 'vs_baseline': None,
			'dtype': 'u8 maps / int32 lattice arithmetic / f64 track', 'data': 'synthetic',
			'config': {'workload': ('BASELINE.json configs[4]: one corpus of %d synthetic DHF1K-shaped 640x360 clips sharded per video over %d GPU(s) '
								'(longest first by map count) x ratios %s, default (ICIP-2021) crop params' % (tot_clips, world, ','.join(ratios))) if c5_main else
								('BASELINE.json configs[2]: %d synthetic DHF1K-shaped 640x360 clips per GPU x ratios %s, '
								'default (ICIP-2021) crop params' % (nc, ','.join(ratios))),
					'total_clips': tot_clips, 'total_frames': tot_frames, 'total_maps': tot_maps,
					'clips_per_gpu': nc, 'frames_per_gpu': NF, 'maps_per_gpu': NM, 'ratios': ratios,
					'boxes_per_frame': R, 'boxes_per_sec': value * R,
					'maps_per_sec': tot_maps * args.steps / (ms_dev / 1e3),
					'input_bytes_per_gpu': NM * H * WPS, 'l2': 'inputs larger than L2 (no flush needed)',
					'value_entry': 'device-resident uint8 [N][140][256]', 'e2e_entry': 'pinned host uint8 [H][W][N] per clip',
					'batches_in_flight': NCTX,
					'counting': 'source frames per second; rounds before r02 counted frames x ratios (2x these figures for 2 ratios)',
					'same_config_as_cpu_arm': False,
					'cpu_arm_difference': 'the CPU arm (cpu_baseline / --impl reference) evaluates ONE target ratio per pass over a bounded sample of the same clips, as the reference does; this arm evaluates every ratio of the workload from one pass and counts each source frame once, so the frames/s ratio understates the work done per frame and equals the maps/s ratio (maps per frame are fixed by the sampling rule)'},
			'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(NM * H * W), 'd2h_bytes_per_step': int(R * NF * 16),
					'ms_per_step': ms_e2e / args.steps,
					'h2d_copy_alone_ms': h2d_ms, 'h2d_copy_alone_gbs': NM * H * W / (h2d_ms / 1e3) / 1e9,
					'cpus_near_gpu': near_cpus,
					'note': 'h2d_copy_alone_* = one plain cudaMemcpyAsync of the same pinned buffer: the floor the host link sets for an e2e step'},
			'e2e_device_producer': dev_prod,
			'gpu_launches': int(launches_n),
			'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
						'traffic': traffic, 'traffic_source': traffic_note,
						'kernel': 'map pipeline of a step: rvb::map_kernel<256,16,Front> -> per size class rvb::prim_kernel<NW,KMAX> -> rvb::map_kernel<NT,TPT,Back> '
								'(the size classes one after the other on the main stream), cut-adjacent chains in rvb::map_kernel<NT,TPT,Mono> x5 on a side stream '
								'(joined before the end event)',
						'algorithmic_bytes_per_step': algo, 'kernel_ms_per_step': map_ms / args.steps,
						'kernel_launches_per_step': map_launches / args.steps,
						'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s'},
			'roofline_streaming_stages': {'what': 'same call with clust_filt=False: threshold, mean saliency, centroid, empty fill, interpolation, low-pass, LOESS, boxes; '
										'frac is of the WHOLE step (frac_map_kernel_alone: the streaming map kernel only)',
										'bound': 'hbm', 'achieved': stream_achieved, 'peak': peak, 'unit': 'GB/s', 'frac': stream_achieved / peak,
										'frac_map_kernel_alone': NM * ALGO_BYTES_PER_MAP * args.steps / (stream_map_ms / 1e3) / 1e9 / peak,
										'kernel_ms_per_step': stream_map_ms / args.steps, 'ms_per_step': stream_ms / args.steps,
										'frames_per_sec': NF * args.steps / (stream_ms / 1e3)},
			'strong_c5': strong,
			'single_clip': c1,
			'prim_stage': prim,
			'iou_stage': iou,
			'crop_stage': crop,
			'clocks': clocks,
		}
		if iou is not None:
			iou['peak_gbs'] = peak
			iou['frac'] = iou['achieved_gbs'] / peak
		if crop is not None:
			crop['peak_gbs'] = peak
			crop['frac'] = crop['achieved_gbs'] / peak
		if cpu_v is not None:
			line['cpu_baseline'] = {'value': cpu_v, 'unit': UNIT, 'cores': cpu_cores, 'kind': 'port', 'sample': cpu_desc}
	if dist is not None:
		dist.barrier()
		dist.destroy_process_group()
	for cx in ctxs:
		cx.close()
	return line


def reference_arm(args, rank, world):
	if rank != 0:
		return None
	vds = make_workload(max(args.cpu_sample, 1), 0)
	v, cores, desc, dt = cpu_baseline(vds, args.cpu_sample, steps=args.steps, warmup=args.warmup)
	NM = sum(x['fc_sel'] for x in vds[:args.cpu_sample])
	NF = sum(x['fc'] for x in vds[:args.cpu_sample])
	return {
		'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
		'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
		'dtype': 'u8 maps / f64 (numpy, scipy, scikit-learn on the CPU)', 'data': 'synthetic',
		'config': {'workload': 'BASELINE.json configs[2] (bounded sample of the same clips): %d clips x ratio %s per step'
							% (args.cpu_sample, RATIOS[0]), 'frames_per_step': NF, 'maps_per_step': NM},
		'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
		'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
		'gpu_launches': 0,
	}


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type=int, default=1)
	ap.add_argument('--steps', type=int, default=20)
	ap.add_argument('--warmup', type=int, default=3)
	ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
	ap.add_argument('--workload', default='c3', choices=['c3', 'c5'],
					help='c3: BASELINE configs[2], --clips clips per GPU, 2 ratios, weak scaling (default, the metric is quoted on it); '
						'c5: configs[4], one corpus of --clips clips (default 2000) sharded over the GPUs, 4 ratios, strong scaling')
	ap.add_argument('--clips', type=int, default=None, help='clips per GPU (c3, default 200) or in the corpus (c5, default 2000)')
	ap.add_argument('--cpu-sample', type=int, default=16, help='clips in the bounded CPU sample')
	ap.add_argument('--c5-clips', type=int, default=2000, help='clips of the configs[4] corpus measured after the headline workload (strong_c5 block); 0 skips it')
	ap.add_argument('--streams', type=int, default=4, help='contexts/streams used to pipeline consecutive batches')
	ap.add_argument('--no-bind', action='store_true', help='do not pin the process to the CPUs local to its GPU')
	ap.add_argument('--phases', action='store_true', help='also print the per-phase SM-cycle split of the map kernel (stderr)')
	args = ap.parse_args()
	rank = int(os.environ.get('RANK', '0'))
	world = int(os.environ.get('WORLD_SIZE', '1'))
	local_rank = int(os.environ.get('LOCAL_RANK', '0'))
	if args.clips is None:
		args.clips = 2000 if args.workload == 'c5' else 200
	# the contract is ONE JSON line on stdout: anything a library prints there meanwhile (NCCL announces its version on
	# stdout) goes to stderr instead
	sys.stdout.flush()
	saved_stdout = os.dup(1)
	os.dup2(2, 1)
	if args.impl == 'reference':
		line = reference_arm(args, rank, world)
	else:
		if args.warmup < 3:
			args.warmup = 3
		line = gpu_arm(args, rank, world, local_rank)
	sys.stdout.flush()
	os.dup2(saved_stdout, 1)
	os.close(saved_stdout)
	if rank == 0 and line is not None:
		print(json.dumps(line), flush=True)


if __name__ == '__main__':
	main()
