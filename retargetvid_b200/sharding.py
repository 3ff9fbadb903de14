"""Per-video sharding across GPUs: videos never interact, so a rank takes a
subset of the videos and there is no collective on the data path; results are gathered on
the host.  Longest-processing-time-first by number of saliency maps balances the ranks."""


def lpt_shards(costs, world):
	"""costs[i] = work of video i (its number of maps).  Returns world lists of indices whose
	total costs are balanced greedily (largest first onto the least loaded rank)."""
	order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
	loads = [0] * world
	shards = [[] for _ in range(world)]
	for i in order:
		r = min(range(world), key=lambda k: (loads[k], k))
		shards[r].append(i)
		loads[r] += costs[i]
	for s in shards:
		s.sort()
	return shards


def my_shard(costs, rank, world):
	return lpt_shards(costs, world)[rank]


def gather_results(local_results, local_indices, world, total, dist=None):
	"""Host gather of per-video results (python objects, e.g. [fc,4] int32 box arrays) into input
	order.  With dist=None (single process) it only reorders."""
	out = [None] * total
	if dist is None or world == 1:
		for i, r in zip(local_indices, local_results):
			out[i] = r
		return out
	gathered = [None] * world
	dist.all_gather_object(gathered, list(zip(local_indices, local_results)))
	for part in gathered:
		for i, r in part:
			out[i] = r
	return out
