"""Prototype (numpy) of the cell-restricted exact Prim planned for prim_kernel: checks that the order and
weights equal oracle.hdbscan_port.prim_order and counts the pair updates saved.

Idea: points are split into cells by empty bands of >= G columns/rows (XY cut), so every cross-cell pair has
d^2 > T = G^2.  While the tree grows inside one cell only that cell's points are updated ("live"); the step is
exact as long as the chosen key is below min(frozen minimum, (T+1) << 13).  Otherwise a "jump" finds the exact
global minimum with a bichromatic search pruned by bounding boxes.

    python tools/lazy_prim_proto.py [n_clips]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hdbscan_port as hp  # noqa: E402
from retargetvid_b200 import synth  # noqa: E402

SH = 13
INF = (1 << 32) - 1


def cells_of(P, G):
	def bands(v, size):
		occ = np.zeros(size + 1, dtype=bool)
		occ[v] = True
		band = np.zeros(size + 1, dtype=np.int64)
		b = 0
		gap = 0
		seen = False
		for t in range(size + 1):
			if occ[t]:
				if seen and gap >= G:
					b += 1
				seen = True
				gap = 0
			else:
				gap += 1
			band[t] = b
		return band[v], b + 1
	by, ny = bands(P[:, 0], 256)
	bx, nx = bands(P[:, 1], 256)
	return by * nx + bx


def lazy_prim(P, core, G=8, stats=None):
	P = np.asarray(P, dtype=np.int64)
	core = np.asarray(core, dtype=np.int64)
	n = len(P)
	T = G * G
	cell = cells_of(P, G)
	key = np.full(n, INF, dtype=np.int64)
	in_tree = np.zeros(n, dtype=bool)
	idx = np.arange(n)
	order = [0]
	weight = []
	pairs = 0
	jump_pairs = 0
	jumps = 0
	dense_pairs = 0

	def mr(i, js):
		d = ((P[js] - P[i]) ** 2).sum(axis=1)
		return np.maximum(np.maximum(core[js], core[i]), d)

	cur = 0
	in_tree[0] = True
	live_cell = cell[0]
	bound = (T + 1) << SH
	tree_list = [0]
	while len(order) < n:
		nt = ~in_tree
		dense_pairs += int(nt.sum())
		live = np.nonzero(nt & (cell == live_cell))[0]
		if len(live):
			k = (mr(cur, live) << SH) | live
			key[live] = np.minimum(key[live], k)
			pairs += len(live)
			g = int(key[live].min())
		else:
			g = INF
		if g >= bound:
			# jump: exact global minimum.  Upper bound U from the stored keys or a couple of refinement scans.
			jumps += 1
			N = np.nonzero(nt)[0]
			S = np.array(tree_list)
			U = int(key[N].min()) >> SH
			if U > (1 << 18):
				i0 = cur
				for _ in range(2):
					m = mr(i0, N)
					j0 = N[int(np.argmin(m))]
					jump_pairs += len(N)
					d = ((P[S] - P[j0]) ** 2).sum(axis=1)
					i0 = S[int(np.argmin(d))]
					jump_pairs += len(S)
				U = int(max(((P[i0] - P[j0]) ** 2).sum(), core[i0], core[j0]))
				key[j0] = min(key[j0], (U << SH) | j0)
			# prune by bounding boxes

			def box_d2(Q, lo, hi):
				d = np.maximum(np.maximum(lo - Q, Q - hi), 0)
				return (d ** 2).sum(axis=1)
			lo, hi = P[N].min(axis=0), P[N].max(axis=0)
			S2 = S[box_d2(P[S], lo, hi) <= U]
			lo, hi = P[S2].min(axis=0), P[S2].max(axis=0)
			N2 = N[box_d2(P[N], lo, hi) <= U]
			jump_pairs += len(S) + len(N)
			for i in S2:
				k = (mr(i, N2) << SH) | N2
				key[N2] = np.minimum(key[N2], k)
			jump_pairs += len(S2) * len(N2)
			g = int(key[N].min())
			new = g & ((1 << SH) - 1)
			live_cell = cell[new]
			frozen = N[(cell[N] != live_cell)]
			fmin = int(key[frozen].min()) if len(frozen) else INF
			bound = min(fmin, (T + 1) << SH)
			# the jump itself is committed even if g >= bound (it is exact)
		new = g & ((1 << SH) - 1)
		order.append(new)
		weight.append(g >> SH)
		in_tree[new] = True
		key[new] = INF
		tree_list.append(new)
		cur = new
	if stats is not None:
		stats.update(pairs=pairs, jump_pairs=jump_pairs, jumps=jumps, dense_pairs=dense_pairs, n=n,
				cells=len(np.unique(cell)))
	return np.array(order), np.array(weight)


def main():
	n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 3
	G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
	specs = synth.config_clips(3, n_clips=n_clips)
	tot = dict(pairs=0, jump_pairs=0, dense_pairs=0, jumps=0, maps=0, bad=0)
	for sp in specs:
		vd = synth.make_clip(**sp)
		sm = vd['smaps']
		for m in range(0, sm.shape[2], 7):
			a = sm[:, :, m].copy()
			a[a < 120] = 0
			P = np.argwhere(a > 0)
			if len(P) < 30:
				continue
			core = hp.core_distances(P, 26)
			o1, w1 = hp.prim_order(P, core)
			st = {}
			o2, w2 = lazy_prim(P, core, G, st)
			ok = np.array_equal(o1, o2) and np.array_equal(w1, w2)
			tot['maps'] += 1
			tot['bad'] += 0 if ok else 1
			for k in ('pairs', 'jump_pairs', 'dense_pairs', 'jumps'):
				tot[k] += st[k]
			print('n=%5d cells=%d jumps=%3d dense=%8d lazy=%8d + jump %8d  ratio %.2f %s' % (
				st['n'], st['cells'], st['jumps'], st['dense_pairs'], st['pairs'], st['jump_pairs'],
				st['dense_pairs'] / max(1, st['pairs'] + st['jump_pairs']), 'ok' if ok else 'MISMATCH'))
	print(tot, 'overall ratio %.2f' % (tot['dense_pairs'] / max(1, tot['pairs'] + tot['jump_pairs'])))


if __name__ == '__main__':
	main()
