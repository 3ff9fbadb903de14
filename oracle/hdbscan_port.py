"""ORACLE (test infrastructure, never shipped or measured as the product).

CPU restatement of the density clustering that the reference calls at
smartVidCrop.py:1099 (``hdbs_clusterer.fit_predict(X)``) with the constructor
arguments of smartVidCrop.py:2340-2348.

The arithmetic lives in the third-party package hdbscan==0.8.26
(README.md:87), which is neither vendored in the reference nor installed here.
It is restated from the published algorithm (Campello, Moulavi, Sander 2013;
McInnes, Healy, Astels 2017) following, step by step, the code path of the
same code base as ported into scikit-learn (sklearn/cluster/_hdbscan/, the
stand-in the survey used), with ``metric='sqeuclidean'`` which forces the
O(n^2) brute/generic route in both packages:

  hdbscan.py:171   _hdbscan_brute            -> pairwise sq. distances
  _reachability.pyx:42,85  mutual_reachability_graph  (core = k-th smallest of a row)
  _linkage.pyx:61  mst_from_mutual_reachability  (Prim from node 0, first-index argmin,
                                                   edge recorded as (previous node, new node))
  hdbscan.py:148   _process_mst              (argsort of edge weights)
  _linkage.pyx:226 make_single_linkage       (union-find, left = cluster of previous node)
  _tree.pyx        _condense_tree / _compute_stability / _get_clusters(eom,
                   allow_single_cluster=True) / _do_labelling

min_samples convention: hdbscan==0.8.26 takes the core distance at sorted-row
index ``min_samples`` (self at index 0); scikit-learn takes index
``min_samples - 1`` and documents the +1 offset itself (hdbscan.py "Notes").
``k`` below is the hdbscan convention, i.e. the distance to the k-th nearest
OTHER point.

Two places where the libraries are implementation defined (PARITY UNPINNED against
hdbscan==0.8.26 itself -- the reference holds no golden vector for it; PINNED against
sklearn's port of the same code on synthetic maps: tests/test_oracle_pinned.py,
tests/golden/hdbscan_fixture.npz):

 1. ties between equal edge weights: both libraries use an UNSTABLE
    ``np.argsort``; all coordinates are integer pixels so weights tie massively.
    The oracle reproduces the permutation of numpy's portable introsort
    (``numpy_aquicksort`` below, the default ``sort='numpy'`` of
    ``single_linkage``); the fixtures are generated with numpy's SIMD sorts
    disabled so that the stand-in library uses the same one
    (tests/golden/tie_flips.json counts what changes under the SIMD sort).
    ``sort='stable'`` is kept for experiments only.
 2. stability sums: the libraries accumulate float64 ``(lambda - birth) * size``
    in condensed-tree row order.  The oracle accumulates exactly in 2^-46
    fixed point (lambda_fix(w) = floor(2^46 / w), w an exact integer), which is
    order independent; excess-of-mass decisions can differ from float64 only if
    two competing stabilities agree to ~1e-11 (tests/test_oracle_pinned.py
    records the smallest relative gap seen on the fixtures).
"""
import numpy as np

LAMBDA_FRAC_BITS = 46


def lambda_fix(w):
	"""floor(2^46 / w) for a positive integer mutual-reachability weight."""
	return (1 << LAMBDA_FRAC_BITS) // int(w)


def core_distances(P, k):
	"""k-th smallest squared distance to another point (row sorted, self at 0).
	_reachability.pyx:132-136 with further_neighbor_idx = k."""
	P = np.asarray(P, dtype=np.int64)
	n = P.shape[0]
	core = np.empty(n, dtype=np.int64)
	step = max(1, (1 << 24) // max(n, 1))
	for s in range(0, n, step):
		d = ((P[s:s + step, None, :] - P[None, :, :]) ** 2).sum(axis=2)
		core[s:s + step] = np.partition(d, k, axis=1)[:, k]
	return core


def prim_order(P, core):
	"""Prim on the mutual-reachability graph from node 0 (_linkage.pyx:97-112).
	Returns (order[n], weight[n-1]): order[i+1] joined with weight[i]; the edge
	the libraries record is (order[i], order[i+1], weight[i])."""
	P = np.asarray(P, dtype=np.int64)
	n = P.shape[0]
	BIG = np.iinfo(np.int64).max
	min_reach = np.full(n, BIG, dtype=np.int64)
	in_tree = np.zeros(n, dtype=bool)
	order = np.empty(n, dtype=np.int64)
	weight = np.empty(n - 1, dtype=np.int64)
	cur = 0
	order[0] = 0
	for i in range(n - 1):
		in_tree[cur] = True
		d = ((P - P[cur]) ** 2).sum(axis=1)
		mr = np.maximum(np.maximum(core, core[cur]), d)
		np.minimum(min_reach, mr, out=min_reach)
		min_reach[in_tree] = BIG
		new = int(np.argmin(min_reach))  # first index among ties
		order[i + 1] = new
		weight[i] = min_reach[new]
		cur = new
	return order, weight


def numpy_aquicksort(v):
	"""The permutation ``np.argsort(v)`` (default kind) returns with numpy's
	portable, non-SIMD introsort -- ``aquicksort_<double>`` in
	numpy/_core/src/npysort/quicksort.cpp, unchanged since numpy 1.x and the only
	implementation in the numpy versions the reference pins (README.md:85-92).
	Median-of-3 Hoare partition, insertion sort below 17 elements, largest
	partition pushed on the stack, heapsort once the depth budget is spent."""
	v = [int(x) for x in v]
	num = len(v)
	t = list(range(num))
	if num < 2:
		return np.array(t, dtype=np.int64)
	SMALL = 15
	stack = []
	pl, pr = 0, num - 1
	cdepth = (num.bit_length() - 1) * 2  # npy_get_msb(num) * 2
	while True:
		popped_by_heapsort = False
		if cdepth < 0:
			_aheapsort(v, t, pl, pr - pl + 1)
			popped_by_heapsort = True
		if not popped_by_heapsort:
			while (pr - pl) > SMALL:
				pm = pl + ((pr - pl) >> 1)
				if v[t[pm]] < v[t[pl]]:
					t[pm], t[pl] = t[pl], t[pm]
				if v[t[pr]] < v[t[pm]]:
					t[pr], t[pm] = t[pm], t[pr]
				if v[t[pm]] < v[t[pl]]:
					t[pm], t[pl] = t[pl], t[pm]
				vp = v[t[pm]]
				pi = pl
				pj = pr - 1
				t[pm], t[pj] = t[pj], t[pm]
				while True:
					pi += 1
					while v[t[pi]] < vp:
						pi += 1
					pj -= 1
					while vp < v[t[pj]]:
						pj -= 1
					if pi >= pj:
						break
					t[pi], t[pj] = t[pj], t[pi]
				pk = pr - 1
				t[pi], t[pk] = t[pk], t[pi]
				if pi - pl < pr - pi:
					stack.append((pi + 1, pr, None))
					pr = pi - 1
				else:
					stack.append((pl, pi - 1, None))
					pl = pi + 1
				cdepth -= 1
				stack[-1] = (stack[-1][0], stack[-1][1], cdepth)
			# insertion sort
			for pi in range(pl + 1, pr + 1):
				vi = t[pi]
				vp = v[vi]
				pj = pi
				while pj > pl and vp < v[t[pj - 1]]:
					t[pj] = t[pj - 1]
					pj -= 1
				t[pj] = vi
		if not stack:
			break
		pl, pr, cdepth = stack.pop()
	return np.array(t, dtype=np.int64)


def _aheapsort(v, t, start, n):
	"""aheapsort_<double> (numpy/_core/src/npysort/heapsort.cpp) on t[start:start+n]."""
	# the C code uses a 1-based view: a = tosort - 1
	def A(i):
		return t[start + i - 1]

	def setA(i, val):
		t[start + i - 1] = val
	for l in range(n >> 1, 0, -1):
		tmp = A(l)
		i = l
		j = l << 1
		while j <= n:
			if j < n and v[A(j)] < v[A(j + 1)]:
				j += 1
			if v[tmp] < v[A(j)]:
				setA(i, A(j))
				i = j
				j += j
			else:
				break
		setA(i, tmp)
	while n > 1:
		tmp = A(n)
		setA(n, A(1))
		n -= 1
		i = 1
		j = 2
		while j <= n:
			if j < n and v[A(j)] < v[A(j + 1)]:
				j += 1
			if v[tmp] < v[A(j)]:
				setA(i, A(j))
				i = j
				j += j
			else:
				break
		setA(i, tmp)


def single_linkage(order, weight, sort='numpy'):
	"""_process_mst + make_single_linkage (_linkage.pyx:226-272).
	Rows: (left, right, weight, size); node ids >= n are internal nodes."""
	n = len(order)
	if sort == 'numpy':
		rows = numpy_aquicksort(weight)
	elif sort == 'stock':
		# whatever this machine's numpy does (numpy >= 2.0: an AVX-512 / AVX2 SIMD sort picked at run time); float64 like
		# the libraries' weight column
		rows = np.argsort(np.asarray(weight, dtype=np.float64))
	else:
		rows = np.argsort(weight, kind='stable')
	parent = list(range(2 * n - 1))
	size = [1] * n + [0] * (n - 1)

	def find(x):
		r = x
		while parent[r] != r:
			r = parent[r]
		while parent[x] != r:
			parent[x], x = r, parent[x]
		return r

	hier = []
	nxt = n
	for e in rows:
		a = find(int(order[e]))
		b = find(int(order[e + 1]))
		hier.append((a, b, int(weight[e]), size[a] + size[b]))
		size[nxt] = size[a] + size[b]
		parent[a] = nxt
		parent[b] = nxt
		nxt += 1
	return hier


def _bfs(hier, n, root):
	"""bfs_from_hierarchy (_tree.pyx): level order, left before right."""
	out = []
	q = [root]
	while q:
		out.extend(q)
		nq = []
		for x in q:
			if x >= n:
				nq.append(hier[x - n][0])
				nq.append(hier[x - n][1])
		q = nq
	return out


def condense_tree(hier, n, min_cluster_size):
	"""_condense_tree (_tree.pyx).  Rows: (parent, child, weight, child_size);
	lambda = 1/weight is kept as the integer weight."""
	root = 2 * (n - 1)
	relabel = {root: n}
	next_label = n + 1
	ignore = set()
	rows = []
	for node in _bfs(hier, n, root):
		if node in ignore or node < n:
			continue
		left, right, w, _ = hier[node - n]
		lc = hier[left - n][3] if left >= n else 1
		rc = hier[right - n][3] if right >= n else 1
		if lc >= min_cluster_size and rc >= min_cluster_size:
			relabel[left] = next_label
			next_label += 1
			rows.append((relabel[node], relabel[left], w, lc))
			relabel[right] = next_label
			next_label += 1
			rows.append((relabel[node], relabel[right], w, rc))
		elif lc < min_cluster_size and rc < min_cluster_size:
			for side in (left, right):
				for sub in _bfs(hier, n, side):
					if sub < n:
						rows.append((relabel[node], sub, w, 1))
					ignore.add(sub)
		elif lc < min_cluster_size:
			relabel[right] = relabel[node]
			for sub in _bfs(hier, n, left):
				if sub < n:
					rows.append((relabel[node], sub, w, 1))
				ignore.add(sub)
		else:
			relabel[left] = relabel[node]
			for sub in _bfs(hier, n, right):
				if sub < n:
					rows.append((relabel[node], sub, w, 1))
				ignore.add(sub)
	return rows


def compute_stability(rows, n):
	"""_compute_stability (_tree.pyx) in exact fixed point (module docstring 2)."""
	birth = {n: 0}
	for (p, c, w, s) in rows:
		birth[c] = lambda_fix(w)
	stab = {}
	for (p, c, w, s) in rows:
		stab[p] = stab.get(p, 0) + (lambda_fix(w) - birth[p]) * s
	return stab


def get_clusters(rows, stab, n, allow_single_cluster=True, gaps=None):
	"""_get_clusters (eom, epsilon 0, no max size) + _do_labelling (_tree.pyx).
	gaps: optional list that receives (relative gap |children - node| / max, max) of every excess-of-mass comparison."""
	node_list = sorted(stab.keys(), reverse=True)
	if not allow_single_cluster:
		node_list = node_list[:-1]
	cluster_rows = [r for r in rows if r[3] > 1]
	children = {}
	for (p, c, w, s) in cluster_rows:
		children.setdefault(p, []).append(c)
	is_cluster = {c: True for c in node_list}
	stab = dict(stab)
	for node in node_list:
		sub = sum(stab[c] for c in children.get(node, []))
		if gaps is not None and children.get(node):
			# (relative gap, larger of the two sums): a gap of 0 with a sum of 0 is the trivial 0 == 0 of two empty sums
			gaps.append((abs(sub - stab[node]) / float(max(sub, stab[node], 1)), max(sub, stab[node])))
		if sub > stab[node]:
			is_cluster[node] = False
			stab[node] = sub
		else:
			q = list(children.get(node, []))
			while q:
				x = q.pop()
				is_cluster[x] = False
				q.extend(children.get(x, []))
	clusters = sorted(c for c in is_cluster if is_cluster[c])
	cmap = {c: i for i, c in enumerate(clusters)}
	cset = set(clusters)

	# _do_labelling: every row whose child is not a selected cluster links the
	# child to its parent; a point's cluster is the top of its chain.
	up = {}
	for (p, c, w, s) in rows:
		if c not in cset:
			up[c] = p
	root_cluster = n
	labels = np.full(n, -1, dtype=np.int64)
	point_w = {}
	root_min_w = None
	for (p, c, w, s) in rows:
		if c < n:
			point_w[c] = w
		if p == root_cluster:
			root_min_w = w if root_min_w is None else min(root_min_w, w)
	for pt in range(n):
		c = pt
		while c in up:
			c = up[c]
		if c != root_cluster:
			labels[pt] = cmap[c]
		elif len(clusters) == 1 and allow_single_cluster:
			# lambda >= max lambda among the root's rows  <=>  weight <= min weight
			if point_w[pt] <= root_min_w:
				labels[pt] = cmap[c]
	return labels


def fit_predict(P, min_cluster_size, min_samples=None, allow_single_cluster=True,
				return_debug=False, sort='numpy', gaps=None):
	"""Labels (-1 = noise) for integer points P [n,2] in (row, col) order.
	Mirrors HDBSCAN(min_cluster_size, min_samples, metric='sqeuclidean',
	cluster_selection_method='eom', allow_single_cluster=True).fit_predict."""
	P = np.asarray(P, dtype=np.int64)
	n = P.shape[0]
	k = min_cluster_size if min_samples is None else min_samples
	k = min(n - 1, k)  # hdbscan_.py: min_samples = min(size - 1, min_samples)
	if k == 0:
		k = 1
	core = core_distances(P, k)
	order, weight = prim_order(P, core)
	hier = single_linkage(order, weight, sort=sort)
	rows = condense_tree(hier, n, min_cluster_size)
	stab = compute_stability(rows, n)
	labels = get_clusters(rows, stab, n, allow_single_cluster, gaps=gaps)
	if return_debug:
		return labels, dict(core=core, order=order, weight=weight, rows=rows, stab=stab)
	return labels
