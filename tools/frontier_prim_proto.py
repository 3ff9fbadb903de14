"""Prototype (numpy) of the lattice-local "frontier" Prim of prim_kernel.cuh: checks that order and weights equal
oracle.hdbscan_port.prim_order and counts the work of each part.

    python tools/frontier_prim_proto.py [n_clips] [R0]

near keys: updates from tree nodes within d^2 <= R0 (lattice ring offsets), applied when a node joins.
far keys : updates from ALL tree nodes, applied lazily -- only when the smallest near key exceeds R0 (a "stall":
           every edge that is left is longer than R0), and then only for the nodes added since the last stall.
A step whose smallest near key is <= R0 is exact without the far keys: a missing update has d^2 > R0, hence a weight
> R0, so it can neither lower the minimum nor join the set of points that tie for it.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hdbscan_port as hp  # noqa: E402
from retargetvid_b200 import synth  # noqa: E402

INF = 1 << 30


def ring_offsets(R0):
	r = int(np.sqrt(R0))
	offs = [(dy, dx, dy * dy + dx * dx) for dy in range(-r, r + 1) for dx in range(-r, r + 1)
			if 0 < dy * dy + dx * dx <= R0]
	return np.array(offs, dtype=np.int64)


def frontier_prim(P, core, R0, stats):
	P = np.asarray(P, dtype=np.int64)
	core = np.asarray(core, dtype=np.int64)
	n = len(P)
	H, W = 256, 256
	grid = -np.ones((H + 64, W + 64), dtype=np.int64)
	grid[P[:, 0] + 32, P[:, 1] + 32] = np.arange(n)
	offs = ring_offsets(R0)
	near = np.full(n, INF, dtype=np.int64)
	far = np.full(n, INF, dtype=np.int64)
	in_tree = np.zeros(n, dtype=bool)
	order = [0]
	weight = []
	in_tree[0] = True
	synced = 0
	cur = 0
	stalls = 0
	sync_pairs = 0
	local_updates = 0
	stall_steps = []
	for step in range(n - 1):
		nb = grid[P[cur, 0] + 32 + offs[:, 0], P[cur, 1] + 32 + offs[:, 1]]
		ok = nb >= 0
		nbi = nb[ok]
		d2 = offs[ok, 2]
		live = ~in_tree[nbi]
		nbi, d2 = nbi[live], d2[live]
		mr = np.maximum(np.maximum(d2, core[cur]), core[nbi])
		imp = mr < near[nbi]
		local_updates += int(imp.sum())
		near[nbi[imp]] = mr[imp]
		nk = np.where(in_tree, INF, near)
		L = int(nk.min())
		if L <= R0:
			new = int(np.argmin(nk))
			w = L
		else:
			stalls += 1
			stall_steps.append(step)
			N = np.nonzero(~in_tree)[0]
			T = np.array(order[synced:], dtype=np.int64)
			sync_pairs += len(N) * len(T)
			for s in range(0, len(T), 256):
				t = T[s:s + 256]
				d = ((P[N][:, None, :] - P[t][None, :, :]) ** 2).sum(axis=2)
				m = np.maximum(np.maximum(d, core[N][:, None]), core[t][None, :]).min(axis=1)
				far[N] = np.minimum(far[N], m)
			synced = len(order)
			k = np.minimum(near, far)
			k = np.where(in_tree, INF, k)
			new = int(np.argmin(k))
			w = int(k[new])
		order.append(new)
		weight.append(w)
		in_tree[new] = True
		cur = new
	stats.update(n=n, stalls=stalls, sync_pairs=sync_pairs, local_updates=local_updates, n_offs=len(offs),
				dense_pairs=n * (n - 1) // 2, stall_steps=stall_steps)
	return np.array(order), np.array(weight)


def maps_of(n_clips, stride=7, with_blend=True):
	specs = synth.config_clips(3, n_clips=n_clips)
	for sp in specs:
		vd = synth.make_clip(**sp)
		sm = vd['smaps']
		for m in range(0, sm.shape[2], stride):
			a = sm[:, :, m].copy()
			a[a < 120] = 0
			yield a
			if with_blend and m + 1 < sm.shape[2] and (m // stride) % 4 == 0:
				# a blend-like map: (thresholded next + this) / 2 with uint8 wrap (smartVidCrop.py:2369-2373)
				b = sm[:, :, m + 1].copy()
				b[b < 120] = 0
				yield ((b + a).astype(float) / 2).astype(np.uint8)


def main():
	n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 3
	R0 = int(sys.argv[2]) if len(sys.argv) > 2 else 25
	tot = dict(maps=0, bad=0, stalls=0, sync_pairs=0, dense_pairs=0, local_updates=0, steps=0)
	for a in maps_of(n_clips):
		P = np.argwhere(a > 0)
		if len(P) < 30:
			continue
		core = hp.core_distances(P, 26)
		o1, w1 = hp.prim_order(P, core)
		st = {}
		o2, w2 = frontier_prim(P, core, R0, st)
		ok = np.array_equal(o1, o2) and np.array_equal(w1, w2)
		tot['maps'] += 1
		tot['bad'] += 0 if ok else 1
		tot['steps'] += st['n'] - 1
		for k in ('stalls', 'sync_pairs', 'dense_pairs', 'local_updates'):
			tot[k] += st[k]
		print('n=%5d stalls=%3d sync=%8d dense=%8d upd/step=%.2f maxw=%d %s %s' % (
			st['n'], st['stalls'], st['sync_pairs'], st['dense_pairs'], st['local_updates'] / max(1, st['n'] - 1),
			int(w1.max()), 'ok' if ok else 'MISMATCH', st['stall_steps'][:6]))
	print(tot)
	print('offsets per step %d; sync pairs / dense pairs = %.3f; stalls per map %.2f; updates per step %.2f' % (
		len(ring_offsets(R0)), tot['sync_pairs'] / max(1, tot['dense_pairs']), tot['stalls'] / max(1, tot['maps']),
		tot['local_updates'] / max(1, tot['steps'])))


if __name__ == '__main__':
	main()
