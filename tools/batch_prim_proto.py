"""Prototype (numpy) of the batched lattice-local Prim of fprim_kernel.cuh: the K lowest-index members of the lowest
non-empty bucket are expanded together (one per lane), and the longest prefix that sequential Prim would also pop, in that
order, is committed.  Checks order and weights against oracle.hdbscan_port.prim_order and prints the batch statistics.

    python tools/batch_prim_proto.py [n_clips] [K]

Cut rule (f_1 < ... < f_K the batch, all at level L, i the 1-based position of the node that makes the update):
  * an update that gives a point a level below L            -> at most i pops
  * an update that gives a point q outside the batch level L -> at most max(i, rank(q)) pops, rank(q) = #{m : f_m < q}
    (q would have to be popped before f_{rank(q)+1}, and it is a candidate once f_i is in the tree)
Updates are min operations, so the committed prefix can be applied in any order.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hdbscan_port as hp  # noqa: E402
from frontier_prim_proto import maps_of, ring_offsets  # noqa: E402

R0 = 18
LEVELS = [1, 2, 4, 5, 8, 9, 10, 13, 16, 17, 18]
INF = 99


def lev(w):
	return LEVELS.index(int(w)) if w <= R0 else 11


def batch_prim(P, core, K, stats):
	P = np.asarray(P, dtype=np.int64)
	n = len(P)
	grid = -np.ones((256 + 64, 256 + 64), dtype=np.int64)
	grid[P[:, 0] + 32, P[:, 1] + 32] = np.arange(n)
	offs = ring_offsets(R0)
	olev = np.array([lev(d) for d in offs[:, 2]])
	clev = np.array([lev(c) for c in core])
	klev = np.full(n, 11, dtype=np.int64)      # 11 = not reached within R0
	in_tree = np.zeros(n, dtype=bool)
	far = np.full(n, 1 << 40, dtype=np.int64)
	synced = 0
	order = [0]
	weight = []

	def updates_of(t):
		nb = grid[P[t, 0] + 32 + offs[:, 0], P[t, 1] + 32 + offs[:, 1]]
		ok = nb >= 0
		j = nb[ok]
		nl = np.maximum(np.maximum(olev[ok], clev[t]), clev[j])
		keep = (~in_tree[j]) & (nl < klev[j])
		return j[keep], nl[keep]

	def join(t, w):
		order.append(int(t))
		weight.append(int(w))
		in_tree[t] = True
		klev[t] = 0

	in_tree[0] = True
	klev[0] = 0
	j, nl = updates_of(0)
	np.minimum.at(klev, j, nl)
	ks = []
	stalls = 0
	while len(order) < n:
		cand_lev = np.where(in_tree, INF, klev)
		L = int(cand_lev.min())
		if L >= 11:
			# stall: lazy far keys
			stalls += 1
			N = np.nonzero(~in_tree)[0]
			T = np.array(order[synced:], dtype=np.int64)
			for s in range(0, len(T), 256):
				t = T[s:s + 256]
				d = ((P[N][:, None, :] - P[t][None, :, :]) ** 2).sum(axis=2)
				m = np.maximum(np.maximum(d, core[N][:, None]), core[t][None, :]).min(axis=1)
				far[N] = np.minimum(far[N], m)
			synced = len(order)
			k = np.where(in_tree, 1 << 41, far)
			g = int(np.argmin(k))
			join(g, k[g])
			j, nl = updates_of(g)
			np.minimum.at(klev, j, nl)
			continue
		f = np.nonzero(cand_lev == L)[0][:K]
		kb = len(f)
		lim = kb
		ups = []
		for i in range(1, kb + 1):          # pass 1, all "lanes" on the pre-batch state
			j, nl = updates_of(f[i - 1])
			ups.append((j, nl))
			for q, l in zip(j, nl):
				if l < L:
					lim = min(lim, i)
				elif l == L and q not in f:
					lim = min(lim, max(i, int(np.searchsorted(f, q))))
		p = min(lim, n - len(order))
		for i in range(p):                  # pass 2: commit
			np.minimum.at(klev, ups[i][0], ups[i][1])
		for i in range(p):
			join(f[i], LEVELS[L])
		ks.append(p)
	stats.update(ks=ks, stalls=stalls)
	return np.array(order), np.array(weight)


def main():
	n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 2
	K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
	allk = []
	bad = 0
	maps = 0
	for a in maps_of(n_clips, stride=9):
		P = np.argwhere(a > 0)
		if len(P) < 30:
			continue
		core = hp.core_distances(P, 26)
		o1, w1 = hp.prim_order(P, core)
		st = {}
		o2, w2 = batch_prim(P, core, K, st)
		ok = np.array_equal(o1, o2) and np.array_equal(w1, w2)
		bad += 0 if ok else 1
		maps += 1
		allk += st['ks']
		print('n=%5d batches %4d mean k %.1f stalls %d %s' % (len(P), len(st['ks']), np.mean(st['ks']), st['stalls'], 'ok' if ok else 'MISMATCH'))
	k = np.array(allk)
	print('maps %d bad %d; batches %d, steps in batches %d, mean k %.2f; share of steps in batches with k >= 8: %.3f, k == %d: %.3f' % (
		maps, bad, len(k), k.sum(), k.mean(), k[k >= 8].sum() / k.sum(), K, k[k == K].sum() / k.sum()))


if __name__ == '__main__':
	main()
