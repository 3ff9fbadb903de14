// Prim on the mutual-reachability graph, lattice-local ("frontier") formulation.  One warp owns one map.
//
// Same result as prim_segment (map_kernel.cuh) and the all-pairs prim_kernel it replaces -- _linkage.pyx:97-112: start at
// point 0, np.argmin = lowest index among equal weights -- but a node that joins the tree only touches the salient
// pixels within d^2 <= kFrR0 of it (60 lattice offsets, found through the occupancy bit mask) instead of every point
// outside the tree, and the minimum is taken from a bucket queue instead of a scan over all keys:
//
//   * near keys.  key_j = min over tree nodes t with d^2(t, j) <= kFrR0 of max(d^2, core_t, core_j), kept only when it is
//     <= kFrR0.  Such a weight is one of 11 values (sums of two squares), and "weight -> level" is monotone, so the level
//     of a mutual-reachability weight is the max of three levels; the queue is 11 bit maps over the point indices and
//     "lowest weight, then lowest index" is the first set bit of the first non-empty map.
//   * why that is exact.  An update that was skipped has d^2 > kFrR0 or a core distance > kFrR0, hence weight > kFrR0.
//     While the smallest near key L is <= kFrR0 no skipped update can lower the minimum or add a point to the set that
//     ties for it, so the node the library's argmin picks and its weight are the ones found here.
//   * open pixels.  A point whose key has reached the level of its own core distance can never improve ("saturated").
//     A bit map over the lattice holds the points outside the tree that are not saturated; a node reads the 9 rows of
//     its neighbourhood from it with funnel shifts and only the set bits (the thin crescent ahead of the flood front,
//     ~7 of 60 neighbours) are looked up, instead of testing 60 offsets.
//   * batches.  The flood of a blob pops long runs of consecutive pixels.  The 32 lowest members f_1 < ... < f_K of the
//     lowest bucket (level L) are expanded together, one per lane, against the state before the batch; the longest prefix
//     sequential Prim would pop in that order is committed: an update that gives a point a level below L ends the prefix
//     at its node i, an update that lifts a point q outside the batch to level L ends it at max(i, #{m : f_m < q}) (q is a
//     candidate once f_i is in the tree and would be popped before the next larger batch node).  Key updates are min
//     operations, so the committed lanes apply theirs in any order (atomicMin; the bucket bits are moved afterwards by
//     the lane whose level won).  81 % of the steps of the bench maps fall into full batches of 32.
//   * stalls.  When every near bucket is empty all remaining edges are longer than kFrR0 (a jump to another blob, the
//     rim of a blob, sparse pixels).  Then the exact keys are completed lazily: far_j = min over ALL tree nodes of
//     max(d^2, core_t, core_j), brought up to date only for the nodes that joined since the previous stall (all pairs,
//     registers, dp4a), so the total far work never exceeds the n^2/2 pair updates of the dense formulation and is
//     0 while a blob is being flooded.
//
// In : scr_pinfo[off + j] = {core_j, (y << 8) | x}  (front kernel), points in row-major order
// Out: scr_pkey[off + s]  = (weight of the edge that added the (s+1)-th node << 13) | node index
#pragma once
#include "map_kernel.cuh"

namespace rvb {

constexpr int kFrR0 = 18;                  // local radius (squared)
constexpr int kFrOffsets = 60;             // lattice offsets with 0 < d^2 <= kFrR0
constexpr int kFrIters = (kFrOffsets + 31) / 32;
// weights <= kFrR0 that a squared lattice distance (and hence a core distance) can take: bit w set
constexpr uint32_t kFrLevelMask = (1u << 1) | (1u << 2) | (1u << 4) | (1u << 5) | (1u << 8) | (1u << 9) | (1u << 10) | (1u << 13) |
								  (1u << 16) | (1u << 17) | (1u << 18);
constexpr int kFrLevels = 11;              // levels 0 .. 10
constexpr uint32_t kFrLevInf = 11;         // "weight > kFrR0" / "not reached": bucket 11 exists but is never scanned
constexpr int kFrSyncSlots = 12;           // far sync: non-tree points per lane held in registers
constexpr int kFrBatch = 32;               // nodes expanded together (one per lane)
constexpr int kFrUpd = 12;                 // key updates a lane can hold for the commit; more ends the batch before its node
// the offsets in row-major order with the level of their squared length (batched expansion: compile-time constants)
__host__ __device__ constexpr int fr_dy(int o) {
	constexpr int8_t t[kFrOffsets] = {-4, -4, -4, -3, -3, -3, -3, -3, -3, -3, -2, -2, -2, -2, -2, -2, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4};
	return t[o];
}
__host__ __device__ constexpr int fr_dx(int o) {
	constexpr int8_t t[kFrOffsets] = {-1, 0, 1, -3, -2, -1, 0, 1, 2, 3, -3, -2, -1, 0, 1, 2, 3, -4, -3, -2, -1, 0, 1, 2, 3, 4, -4, -3, -2, -1, 1, 2, 3, 4, -4, -3, -2, -1, 0, 1, 2, 3, 4, -3, -2, -1, 0, 1, 2, 3, -3, -2, -1, 0, 1, 2, 3, -1, 0, 1};
	return t[o];
}
__host__ __device__ constexpr int fr_ol(int o) {
	constexpr uint8_t t[kFrOffsets] = {9, 8, 9, 10, 7, 6, 5, 6, 7, 10, 7, 4, 3, 2, 3, 4, 7, 9, 6, 3, 1, 0, 1, 3, 6, 9, 8, 5, 2, 0, 0, 2, 5, 8, 9, 6, 3, 1, 0, 1, 3, 6, 9, 7, 4, 3, 2, 3, 4, 7, 10, 7, 6, 5, 6, 7, 10, 9, 8, 9};
	return t[o];
}

// per-point record: bits 0..15 (y << 8) | x, 16..19 level of the core distance, 20..23 level of the key (0 once in the
// tree: nothing is < 0, and a key of level 0 cannot be improved either)
constexpr uint32_t kFrKeyShift = 20, kFrCoreShift = 16;

struct FrOffsetTable {
	int8_t dy[kFrIters * 32];
	int8_t dx[kFrIters * 32];
	uint8_t lev[kFrIters * 32];            // level of d^2; padding entries: dy = dx = 0, level kFrLevInf
	uint32_t level_w[16];                  // weight of level l
};
__constant__ FrOffsetTable c_froffs;

struct FPrimArgs {
	const int *list;       // maps of this size class
	const int *list_len;
	int *head;
	const MapOut *out;     // n_points
	const int *scr_off;
	const uint2 *scr_pinfo;
	uint32_t *scr_pkey;
	uint32_t *scr_far;     // stall path: far keys
	uint32_t *scr_alist;   // stall path: indices of the points outside the tree (from the front) and of the tree nodes
	                       // not yet applied to the far keys (from the back)
	int cap;               // points per map this launch has shared memory for
	int H, W, RS;          // lattice; RS = 32-bit words per row of the occupancy mask
	unsigned long long *phase_cycles;
	unsigned long long *work;   // optional [5]: steps, stalls, far pair updates / 32, near updates, batches
};

// bytes of dynamic shared memory for a capacity of `cap` points on an H x W lattice
__host__ __device__ inline int fprim_smem_bytes(int cap, int H, int W) {
	const int RS = (W + 31) >> 5;
	const int nw = cap >> 5;
	int o = 0;
	o += H * RS * 4;                       // occ
	o += ((H * RS * 2) + 15) & ~15;        // wbase
	o = (o + 15) & ~15;
	o += cap * 4;                          // pk
	o += (kFrLevels + 1) * nw * 4;         // buckets (+ the dummy one)
	o += nw * 4 * 2;                       // alive, synced
	o += H * (RS + 2) * 4;                 // open (one guard word on either side of a row)
	return (o + 15) & ~15;
}

__host__ __device__ inline int fr_level_of(uint32_t w) {
	// weights above kFrR0 (anything that is not a level weight cannot occur below it) -> kFrLevInf
	if (w > (uint32_t)kFrR0) return (int)kFrLevInf;
#ifdef __CUDA_ARCH__
	return __popc(kFrLevelMask & ((1u << w) - 1u));
#else
	return __builtin_popcount(kFrLevelMask & ((1u << w) - 1u));
#endif
}

// shared-memory accesses of the hot loop through 32-bit addresses; volatile keeps their order, no memory clobber lets
// the ALU work be scheduled around them (phase boundaries are __syncwarp()s)
__device__ __forceinline__ uint32_t fr_lds32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint32_t fr_lds16(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void fr_sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ void fr_red_and(uint32_t addr, uint32_t v) { asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ void fr_red_or(uint32_t addr, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v)); }

// rows of the neighbourhood d^2 <= kFrR0: half width of row dy = -4 .. 4, and the bit offset of each row in the packed
// 61-bit list of open neighbours (rows 0..4 in word A, rows 5..8 in word B)
__host__ __device__ constexpr int fr_row_hw(int r) { return (r == 0 || r == 8) ? 1 : (r == 1 || r == 2 || r == 6 || r == 7) ? 3 : 4; }

template <int CAP>
__global__ void __launch_bounds__(32, 16) fprim_kernel(const FPrimArgs a) {
	constexpr int NWORDS = CAP / 32;
	extern __shared__ __align__(16) uint8_t fsm[];
	__shared__ uint32_t uq[32 * kFrUpd];      // per lane: the key updates of its batch node, (old level << 24) | (level << 16) | point
	__shared__ uint16_t sel[kFrBatch];        // the batch: lowest members of the lowest bucket
	const int lane = threadIdx.x;
	const int H = a.H, RS = a.RS, RSO = a.RS + 2;
	uint32_t *occ = reinterpret_cast<uint32_t *>(fsm);
	uint16_t *wbase = reinterpret_cast<uint16_t *>(fsm + H * RS * 4);
	const int o_pk = (H * RS * 4 + (((H * RS * 2) + 15) & ~15) + 15) & ~15;
	uint32_t *pk = reinterpret_cast<uint32_t *>(fsm + o_pk);
	uint32_t *bm = reinterpret_cast<uint32_t *>(fsm + o_pk + CAP * 4);
	uint32_t *alive = bm + (kFrLevels + 1) * NWORDS;
	uint32_t *synced = alive + NWORDS;
	uint32_t *open = synced + NWORDS;         // [H][RS + 2]: points outside the tree whose key can still improve

	while (true) {
		int m = -1;
		if (lane == 0) {
			const int i = atomicAdd(a.head, 1);
			m = (i < *a.list_len) ? a.list[i] : -1;
		}
		m = __shfl_sync(0xffffffffu, m, 0);
		if (m < 0) break;
		const long long t0 = clock64();
		const int n = a.out[m].n_points;
		const size_t off = (size_t)a.scr_off[m];
		const uint2 *gp = a.scr_pinfo + off;
		uint32_t *pkey_out = a.scr_pkey + off;
		uint32_t *far = a.scr_far + off;
		uint32_t *alist = a.scr_alist + off;

		// ---- setup: occupancy mask, first point index of every mask word, per-point records, empty buckets -----------
		for (int i = lane; i < H * RS; i += 32) occ[i] = 0u;
		for (int i = lane; i < (kFrLevels + 3) * NWORDS + H * RSO; i += 32) bm[i] = 0u;   // buckets, alive, synced, open
		__syncwarp();
		for (int j0 = 0; j0 < n; j0 += 32) {
			const int j = j0 + lane;
			uint2 pi = make_uint2(0u, 0xFFFFFFFFu);
			if (j < n) pi = gp[j];
			const uint32_t xy = pi.y;
			// points are in row-major order: the first point of a mask word is the one whose predecessor lies in another word
			uint32_t pv = __shfl_up_sync(0xffffffffu, xy, 1);
			if (lane == 0) pv = (j > 0 && j < n) ? gp[j - 1].y : 0xFFFFFFFFu;
			if (j < n) {
				const uint32_t clev = (uint32_t)fr_level_of(pi.x & 0x1FFFFu);
				pk[j] = (xy & 0xFFFFu) | (clev << kFrCoreShift) | (kFrLevInf << kFrKeyShift);
				const int y = (int)(xy >> 8), x = (int)(xy & 0xFFu);
				const int wi = y * RS + (x >> 5);
				atomicOr(&occ[wi], 1u << (x & 31));
				if (clev < kFrLevInf) atomicOr(&open[y * RSO + 1 + (x >> 5)], 1u << (x & 31));
				const bool first = (j == 0) || ((int)(pv >> 8) * RS + (int)((pv & 0xFFu) >> 5) != wi);
				if (first) wbase[wi] = (uint16_t)j;
			}
		}
		for (int w = lane; w < NWORDS; w += 32) {
			const int lo = w * 32;
			alive[w] = (n >= lo + 32) ? 0xFFFFFFFFu : ((n > lo) ? ((1u << (n - lo)) - 1u) : 0u);
		}
		__syncwarp();

		uint32_t lvl_any = 0u;      // bit l: bucket l may be non-empty (uniform)
		bool far_valid = false;     // far[] holds the keys of the previous stall
		unsigned long long w_stalls = 0, w_far = 0, w_near = 0, w_batches = 0;
		int step = 0;               // edges found
		// the batch to expand: kb nodes (lane < kb holds f), all at level L; forced: ONE node that is already in the Prim
		// order (the root, a node found by a stall, a node whose updates do not fit a lane's queue)
		int kb = 1, f = 0, L = (int)kFrLevInf;
		bool forced = true;
		while (true) {
			// ---- pass 1: every lane expands its node against the state before the batch (reads only) --------------------
			++w_batches;
			const int f_last = __shfl_sync(0xffffffffu, f, kb - 1);
			const int ucap = forced ? 32 * kFrUpd : kFrUpd;
			int lim = kFrBatch;        // pops the batch may commit as far as this lane can tell
			int cnt = 0;               // updates queued by this lane
			int qmin = 0x7fffffff;     // smallest new level-L candidate that lies between this node and the last batch node
			if (lane < kb) {
				const uint32_t tv = pk[f];
				const int cy = (int)((tv >> 8) & 0xFFu), cx = (int)(tv & 0xFFu);
				const uint32_t clev = (tv >> kFrCoreShift) & 0xFu;
				// the open pixels of the 9 rows, packed: rows 0..4 -> A (3 + 7 + 7 + 9 + 9 bits), rows 5..8 -> B (9 + 7 + 7 + 3)
				unsigned long long A = 0ull, B = 0ull;
#pragma unroll
				for (int r = 0; r < 9; ++r) {
					constexpr int kShift[9] = {0, 3, 10, 17, 26, 0, 9, 16, 23};
					const int hw = fr_row_hw(r);
					const int y = cy + r - 4;
					uint32_t bits = 0u;
					if ((unsigned)y < (unsigned)H) {
						const int x0 = cx - hw;
						const uint32_t *row = open + y * RSO + ((x0 + 32) >> 5);
						bits = __funnelshift_r(row[0], row[1], x0 & 31) & ((1u << (2 * hw + 1)) - 1u);
						if (r == 4) bits &= ~(1u << hw);      // the node itself
					}
					if (r < 5) A |= (unsigned long long)bits << kShift[r];
					else B |= (unsigned long long)bits << kShift[r];
				}
				while ((A | B) != 0ull) {
					// next open neighbour: bit position -> (row, column)
					int r, b;
					if (A != 0ull) {
						const int p = __ffsll((long long)A) - 1;
						A &= A - 1ull;
						r = (p >= 26) ? 4 : (p >= 17) ? 3 : (p >= 10) ? 2 : (p >= 3) ? 1 : 0;
						b = p - ((p >= 26) ? 26 : (p >= 17) ? 17 : (p >= 10) ? 10 : (p >= 3) ? 3 : 0);
					} else {
						const int p = __ffsll((long long)B) - 1;
						B &= B - 1ull;
						r = (p >= 23) ? 8 : (p >= 16) ? 7 : (p >= 9) ? 6 : 5;
						b = p - ((p >= 23) ? 23 : (p >= 16) ? 16 : (p >= 9) ? 9 : 0);
					}
					const int hw = (r == 0 || r == 8) ? 1 : (r == 1 || r == 2 || r == 6 || r == 7) ? 3 : 4;
					const int dy = r - 4, dx = b - hw;
					const int y = cy + dy, x = cx + dx;
					const uint32_t olev = (uint32_t)__popc(kFrLevelMask & ((1u << (dy * dy + dx * dx)) - 1u));
					const int wi = y * RS + (x >> 5);
					const uint32_t word = occ[wi];
					const int j = (int)wbase[wi] + __popc(word & ((1u << (x & 31)) - 1u));
					const uint32_t v = pk[j];
					const uint32_t nlev = max(max(olev, clev), (v >> kFrCoreShift) & 0xFu);
					if (nlev < (v >> kFrKeyShift)) {
						if (cnt < ucap) uq[lane * kFrUpd + cnt] = (nlev << 16) | (uint32_t)j;
						++cnt;
						if ((int)nlev < L || ((int)nlev == L && j < f)) lim = min(lim, lane + 1);
						else if ((int)nlev == L && j < f_last) qmin = min(qmin, j);
					}
				}
				if (cnt > ucap) lim = min(lim, lane);      // cannot hold its updates: the batch ends before this node
			}
			int p = kb;
			if (!forced) {
				// rank of a new candidate between the batch nodes: #{m : f_m < q}; q is popped before the next larger one
				uint32_t need = __ballot_sync(0xffffffffu, qmin != 0x7fffffff);
				while (need != 0u) {
					const int src = __ffs(need) - 1;
					need &= need - 1u;
					const int q = __shfl_sync(0xffffffffu, qmin, src);
					const int rank = __popc(__ballot_sync(0xffffffffu, lane < kb && f < q));
					if (lane == src) lim = min(lim, max(lane + 1, rank));
				}
				p = __reduce_min_sync(0xffffffffu, lim);
				p = min(min(p, kb), n - 1 - step);
			}
			const uint32_t wL = c_froffs.level_w[L];
			if (p == 0) {
				// the first node alone has more updates than a lane's queue: pop it and expand it with the whole queue
				const int f0 = __shfl_sync(0xffffffffu, f, 0);
				if (lane == 0) pkey_out[step] = (wL << kKeyShift) | (uint32_t)f0;
				++step;
				kb = 1; forced = true;
				__syncwarp();
				continue;
			}
			// ---- pass 2: the first p lanes commit.  Keys first (atomicMin), then the bucket bits by the lane that won ------
			const bool mine = lane < p;
			const int ncommit = mine ? min(cnt, ucap) : 0;
			const int maxc = __reduce_max_sync(0xffffffffu, ncommit);
			__syncwarp();
			for (int u = 0; u < maxc; ++u) {
				if (u < ncommit) {
					const uint32_t e = uq[lane * kFrUpd + u];
					const uint32_t j = e & 0xFFFFu, nlev = (e >> 16) & 0xFu;
					const uint32_t old = atomicMin(&pk[j], (pk[j] & 0x000FFFFFu) | (nlev << kFrKeyShift));
					const uint32_t olv = old >> kFrKeyShift;
					uq[lane * kFrUpd + u] = e | ((olv > nlev) ? (olv << 24) : 0u);      // old level 0 never occurs here: "not lowered"
				}
			}
			__syncwarp();
			uint32_t newlv = 0u;
			for (int u = 0; u < maxc; ++u) {
				if (u < ncommit) {
					const uint32_t e = uq[lane * kFrUpd + u];
					const uint32_t olv = e >> 24;
					if (olv != 0u) {
						const uint32_t j = e & 0xFFFFu, nlev = (e >> 16) & 0xFu;
						const uint32_t jb = 1u << (j & 31);
						atomicAnd(&bm[olv * NWORDS + (j >> 5)], ~jb);             // (bucket kFrLevInf is a dummy)
						const uint32_t vj = pk[j];
						if ((vj >> kFrKeyShift) == nlev) {
							atomicOr(&bm[nlev * NWORDS + (j >> 5)], jb);
							newlv |= 1u << nlev;
							if (nlev == ((vj >> kFrCoreShift) & 0xFu)) {
								// saturated: the key has reached the level of the point's own core distance
								const int y = (int)((vj >> 8) & 0xFFu), x = (int)(vj & 0xFFu);
								atomicAnd(&open[y * RSO + 1 + (x >> 5)], ~(1u << (x & 31)));
							}
						}
						if (a.work != nullptr) ++w_near;
					}
				}
			}
			// retire the committed nodes
			if (mine) {
				if (!forced) pkey_out[step + lane] = (wL << kKeyShift) | (uint32_t)f;
				const uint32_t tv = pk[f];
				pk[f] = tv & 0x000FFFFFu;
				atomicAnd(&alive[f >> 5], ~(1u << (f & 31)));
				atomicAnd(&bm[L * NWORDS + (f >> 5)], ~(1u << (f & 31)));
				const int y = (int)((tv >> 8) & 0xFFu), x = (int)(tv & 0xFFu);
				atomicAnd(&open[y * RSO + 1 + (x >> 5)], ~(1u << (x & 31)));
			}
			lvl_any |= __reduce_or_sync(0xffffffffu, newlv);
			if (!forced) step += p;
			__syncwarp();
			if (step >= n - 1) break;
			// ---- the lowest members of the lowest non-empty bucket: the next batch -------------------------------------------
			L = -1;
			kb = 0;
			while (lvl_any != 0u) {
				const int l = __ffs(lvl_any) - 1;
				const uint32_t *b = bm + l * NWORDS;
				int total = 0;
				for (int wb0 = 0; wb0 < NWORDS && total < kFrBatch; wb0 += 32) {
					const int w = wb0 + lane;
					uint32_t mm = (w < NWORDS) ? b[w] : 0u;
					const int c = __popc(mm);
					int inc = c;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) {
						const int t = __shfl_up_sync(0xffffffffu, inc, o);
						if (lane >= o) inc += t;
					}
					int pos = total + inc - c;
					while (mm != 0u && pos < kFrBatch) {
						sel[pos++] = (uint16_t)(w * 32 + __ffs(mm) - 1);
						mm &= mm - 1u;
					}
					total += __shfl_sync(0xffffffffu, inc, 31);
				}
				if (total > 0) { L = l; kb = min(total, kFrBatch); break; }
				lvl_any &= ~(1u << l);
			}
			__syncwarp();
			forced = false;
			if (L >= 0) {
				f = (lane < kb) ? (int)sel[lane] : 0;
				continue;
			}
			// ---- stall: every edge that is left is longer than kFrR0.  Complete the keys with the tree nodes that
			// joined since the last stall (all of them at the first one) and take the exact minimum.
			++w_stalls;
			{
				// alist[0 .. cnt): the points outside the tree; alist[n - tcnt .. n): the tree nodes not applied yet
				int acnt = 0, tcnt = 0;
				for (int wb0 = 0; wb0 < NWORDS; wb0 += 32) {
					const int w = wb0 + lane;
					uint32_t ma = 0u, mt = 0u;
					if (w < NWORDS) {
						ma = alive[w];
						mt = ~ma & ~synced[w];
						if (w * 32 + 32 > n) mt &= (n > w * 32) ? ((1u << (n - w * 32)) - 1u) : 0u;
					}
					int ia = __popc(ma), itn = __popc(mt);
					const int ca = ia, ct = itn;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) {
						const int ta = __shfl_up_sync(0xffffffffu, ia, o);
						const int tt = __shfl_up_sync(0xffffffffu, itn, o);
						if (lane >= o) { ia += ta; itn += tt; }
					}
					int pa = acnt + ia - ca;
					while (ma) {
						const int bpos = __ffs(ma) - 1;
						ma &= ma - 1u;
						alist[pa++] = (uint32_t)(w * 32 + bpos);
					}
					int pt = n - 1 - (tcnt + itn - ct);
					while (mt) {
						const int bpos = __ffs(mt) - 1;
						mt &= mt - 1u;
						alist[pt--] = (uint32_t)(w * 32 + bpos);
					}
					acnt += __shfl_sync(0xffffffffu, ia, 31);
					tcnt += __shfl_sync(0xffffffffu, itn, 31);
				}
				__syncwarp();
				uint32_t best = 0xFFFFFFFFu;
				for (int b0 = 0; b0 < acnt; b0 += 32 * kFrSyncSlots) {
					uint32_t sxy[kFrSyncSlots], sc[kFrSyncSlots], sf[kFrSyncSlots];
					const int nslots = min(kFrSyncSlots, (acnt - b0 + 31) >> 5);
#pragma unroll
					for (int s = 0; s < kFrSyncSlots; ++s) {
						const int e = b0 + s * 32 + lane;
						sxy[s] = 0u; sc[s] = 0x1FFFFu; sf[s] = 0x3FFFFu;
						if (e < acnt) {
							const int j = (int)alist[e];
							sxy[s] = pk[j] & 0xFFFFu;
							sc[s] = gp[j].x & 0x1FFFFu;
							if (far_valid) sf[s] = far[j];
						}
					}
					for (int tb = 0; tb < tcnt; tb += 32) {
						// 32 tree nodes per batch: one per lane, broadcast by shuffles
						const int te = tb + lane;
						uint32_t txy = 0u, tc = 0x1FFFFu;
						if (te < tcnt) {
							const int t = (int)alist[n - 1 - te];
							txy = pk[t] & 0xFFFFu;
							tc = gp[t].x & 0x1FFFFu;
						}
						const int nb = min(32, tcnt - tb);
						for (int q = 0; q < nb; ++q) {
							const uint32_t bxy = __shfl_sync(0xffffffffu, txy, q);
							const uint32_t bc = __shfl_sync(0xffffffffu, tc, q);
#pragma unroll
							for (int s = 0; s < kFrSyncSlots; ++s) {
								if (s < nslots) {
									const uint32_t ad = __vabsdiffu4(sxy[s], bxy);
									const uint32_t d2 = __dp4a(ad, ad, 0u);
									sf[s] = min(sf[s], max(d2, max(sc[s], bc)));
								}
							}
						}
						if (a.work != nullptr) w_far += (unsigned long long)nb * (unsigned long long)nslots;
					}
#pragma unroll
					for (int s = 0; s < kFrSyncSlots; ++s) {
						const int e = b0 + s * 32 + lane;
						if (e < acnt) {
							const uint32_t j = alist[e];
							far[j] = sf[s];
							best = min(best, (sf[s] << kKeyShift) | j);
						}
					}
				}
				for (int w = lane; w < NWORDS; w += 32) synced[w] = ~alive[w];
				far_valid = true;
				const uint32_t gk = __reduce_min_sync(0xffffffffu, best);
				if (lane == 0) pkey_out[step] = gk;
				++step;
				f = (int)(gk & kKeyIdxMask);
				kb = 1; forced = true; L = (int)kFrLevInf;
				__syncwarp();
			}
		}
		if (lane == 0) {
			if (a.phase_cycles != nullptr) atomicAdd(&a.phase_cycles[3], (unsigned long long)(clock64() - t0));
			if (a.work != nullptr) {
				atomicAdd(&a.work[0], (unsigned long long)(n > 0 ? n - 1 : 0));
				atomicAdd(&a.work[1], w_stalls);
			}
		}
		if (a.work != nullptr) {
			// w_far is uniform (counted by every lane alike, added once); w_near is per lane
			w_near = __reduce_add_sync(0xffffffffu, (unsigned)w_near);
			if (lane == 0) { atomicAdd(&a.work[2], w_far); atomicAdd(&a.work[3], w_near); atomicAdd(&a.work[4], w_batches); }
		}
		__syncwarp();
	}
}

}  // namespace rvb
