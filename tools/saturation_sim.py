"""How much of the all-pairs Prim's update work a "frozen" set would save: a live point whose key weight equals its own
core distance can never improve (every candidate weight is >= its core), so it needs no update, only a place in the
argmin.  Simulates prim_kernel's slot schedule (live points in K register slots per thread, K from 24 down, compaction
when the live points fit fewer slots) with and without moving saturated points out of the slots at every compaction.

    python tools/saturation_sim.py [n_clips]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hdbscan_port as hp  # noqa: E402
from frontier_prim_proto import maps_of  # noqa: E402

SLOTS = [1, 2, 3, 4, 6, 8, 12, 16, 20, 24]


def slots_for(live, nt):
	need = (live + nt - 1) // nt
	for k in SLOTS:
		if k >= need:
			return k
	return SLOTS[-1]


def nt_for(n):
	return 32 if n <= 768 else 64 if n <= 1536 else 128 if n <= 3072 else 256


def simulate(P, core):
	P = np.asarray(P, dtype=np.int64)
	core = np.asarray(core, dtype=np.int64)
	n = len(P)
	nt = nt_for(n)
	key = np.full(n, 1 << 40, dtype=np.int64)
	in_tree = np.zeros(n, dtype=bool)
	in_tree[0] = True
	cur = 0
	# baseline schedule
	base_cost = 0       # slot updates executed (K x NT per step)
	sat_steps = 0       # live point-steps spent saturated
	live_steps = 0
	# frozen schedule
	in_slots = ~in_tree.copy()
	k_base = slots_for(n - 1, nt)
	k_frz = k_base
	frz_cost = 0
	compactions = 0
	for step in range(n - 1):
		d = ((P - P[cur]) ** 2).sum(axis=1)
		mr = np.maximum(np.maximum(d, core), core[cur])
		key = np.where(in_tree, key, np.minimum(key, mr))
		live = ~in_tree
		sat = live & (key == core)
		live_steps += int(live.sum())
		sat_steps += int(sat.sum())
		base_cost += k_base * nt
		frz_cost += k_frz * nt
		kk = np.where(in_tree, 1 << 41, key * 8192 + np.arange(n))
		cur = int(np.argmin(kk))
		in_tree[cur] = True
		in_slots[cur] = False
		nlive = n - 2 - step
		if slots_for(nlive, nt) < k_base:
			k_base = slots_for(nlive, nt)
		# frozen variant: compaction when the points still in slots fit fewer slots; saturated ones leave at that time
		cnt = int(in_slots.sum())
		if slots_for(cnt, nt) < k_frz:
			live2 = ~in_tree
			in_slots = live2 & ~(key == core)
			k_frz = slots_for(int(in_slots.sum()), nt)
			compactions += 1
	return base_cost, frz_cost, sat_steps, live_steps, compactions


def main():
	n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 2
	tot = np.zeros(5)
	for a in maps_of(n_clips, stride=11):
		P = np.argwhere(a > 0)
		if len(P) < 60 or len(P) > 4096:
			continue
		core = hp.core_distances(P, 26)
		r = simulate(P, core)
		tot += np.array(r)
		print('n=%5d slot-updates %9d -> %9d (%.3f), saturated share of live point-steps %.3f, compactions %d' % (
			len(P), r[0], r[1], r[1] / r[0], r[2] / r[3], r[4]))
	print('total: slot updates x%.3f, saturated share %.3f' % (tot[1] / tot[0], tot[2] / tot[3]))


if __name__ == '__main__':
	main()
