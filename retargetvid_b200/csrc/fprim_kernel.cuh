// Prim on the mutual-reachability graph, lattice-local ("frontier") formulation.  One warp owns one map.
//
// Same result as prim_segment (map_kernel.cuh) and the dense prim_kernel it replaces -- _linkage.pyx:97-112: start at
// point 0, np.argmin = lowest index among equal weights -- but a node that joins the tree only touches the salient
// pixels within d^2 <= kFrR0 of it (80 lattice offsets, found through the occupancy bit mask) instead of every point
// outside the tree, and the minimum is taken from a bucket queue instead of a scan over all keys:
//
//   * near keys.  key_j = min over tree nodes t with d^2(t, j) <= kFrR0 of max(d^2, core_t, core_j), kept only when it is
//     <= kFrR0.  Such a weight is one of 13 values (sums of two squares), so the queue is 13 bit maps over the point
//     indices; "lowest weight, then lowest index" is the first set bit of the first non-empty map.
//   * why that is exact.  An update that was skipped has d^2 > kFrR0, hence weight > kFrR0.  While the smallest near key
//     L is <= kFrR0 no skipped update can lower the minimum or add a point to the set that ties for it, so the node the
//     library's argmin picks and its weight are the ones found here.
//   * stalls.  When every near bucket is empty all remaining edges are longer than kFrR0 (a jump to another blob, sparse
//     pixels).  Then the exact keys are completed lazily: far_j = min over ALL tree nodes of max(d^2, core_t, core_j),
//     brought up to date only for the nodes that joined since the previous stall (dense, registers, dp4a), so the total
//     far work never exceeds the n^2/2 pair updates of the dense formulation and is 0 for a single blob.
//
// In : scr_pinfo[off + j] = {core_j, (y << 8) | x}  (front kernel), points in row-major order
// Out: scr_pkey[off + s]  = (weight of the edge that added the (s+1)-th node << 13) | node index
#pragma once
#include "map_kernel.cuh"

namespace rvb {

constexpr int kFrR0 = 25;                  // local radius (squared)
constexpr int kFrOffsets = 80;             // lattice offsets with 0 < d^2 <= kFrR0
constexpr int kFrIters = (kFrOffsets + 31) / 32;
// weights <= kFrR0 that a squared lattice distance (and hence a core distance) can take: bit w set
constexpr uint32_t kFrLevelMask = (1u << 1) | (1u << 2) | (1u << 4) | (1u << 5) | (1u << 8) | (1u << 9) | (1u << 10) | (1u << 13) |
								  (1u << 16) | (1u << 17) | (1u << 18) | (1u << 20) | (1u << 25);
constexpr int kFrLevels = 13;
constexpr uint32_t kFrKeyInf = 0x7FFFu;    // key field of a point no tree node has reached within kFrR0
constexpr int kFrSyncSlots = 16;           // far sync: non-tree points per lane held in registers

struct FrOffsetTable {
	int8_t dy[kFrIters * 32];
	int8_t dx[kFrIters * 32];
	uint32_t d2[kFrIters * 32];            // padding entries: dy = dx = 0, d2 = 0x7FFFFF (never <= kFrR0)
	uint32_t level_w[32];                  // weight of level l
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
	uint32_t *scr_alist;   // stall path: indices of the points outside the tree
	int cap;               // points per map this launch has shared memory for
	int H, W, RS;          // lattice; RS = 32-bit words per row of the occupancy mask
	unsigned long long *phase_cycles;
	unsigned long long *work;   // optional [4]: steps, stalls, far pair updates / 32, near updates
};

// bytes of dynamic shared memory for a capacity of `cap` points on an H x W lattice
__host__ __device__ inline int fprim_smem_bytes(int cap, int H, int W) {
	const int RS = (W + 31) >> 5;
	const int nw = cap >> 5;
	int o = 0;
	o += H * RS * 4;                 // occ
	o += ((H * RS * 2) + 15) & ~15;  // wbase
	o = (o + 15) & ~15;
	o += cap * 4;                    // pk
	o += cap * 2;                    // pxy
	o += kFrLevels * nw * 4;         // buckets
	o += nw * 4 * 2;                 // alive, synced
	return (o + 15) & ~15;
}

__device__ __forceinline__ int fr_level(uint32_t w) { return __popc(kFrLevelMask & ((1u << w) - 1u)); }

template <int CAP>
__global__ void __launch_bounds__(32) fprim_kernel(const FPrimArgs a) {
	constexpr int NWORDS = CAP / 32;
	constexpr int WPL = (NWORDS + 31) / 32;   // bitmap words per lane
	extern __shared__ __align__(16) uint8_t fsm[];
	const int lane = threadIdx.x;
	const int H = a.H, W = a.W, RS = a.RS;
	uint32_t *occ = reinterpret_cast<uint32_t *>(fsm);
	uint16_t *wbase = reinterpret_cast<uint16_t *>(fsm + H * RS * 4);
	const int o_pk = (H * RS * 4 + (((H * RS * 2) + 15) & ~15) + 15) & ~15;
	uint32_t *pk = reinterpret_cast<uint32_t *>(fsm + o_pk);
	uint16_t *pxy = reinterpret_cast<uint16_t *>(fsm + o_pk + CAP * 4);
	uint32_t *bm = reinterpret_cast<uint32_t *>(fsm + o_pk + CAP * 6);
	uint32_t *alive = bm + kFrLevels * NWORDS;
	uint32_t *synced = alive + NWORDS;

	int ody[kFrIters], odx[kFrIters];
	uint32_t od2[kFrIters];
#pragma unroll
	for (int it = 0; it < kFrIters; ++it) {
		ody[it] = c_froffs.dy[it * 32 + lane];
		odx[it] = c_froffs.dx[it * 32 + lane];
		od2[it] = c_froffs.d2[it * 32 + lane];
	}

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
		for (int i = lane; i < (kFrLevels + 2) * NWORDS; i += 32) bm[i] = 0u;   // buckets, alive, synced
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
				pk[j] = (pi.x & 0x1FFFFu) | (kFrKeyInf << 17);
				pxy[j] = (uint16_t)xy;
				const int y = (int)(xy >> 8), x = (int)(xy & 0xFFu);
				const int wi = y * RS + (x >> 5);
				atomicOr(&occ[wi], 1u << (x & 31));
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
		unsigned long long w_stalls = 0, w_far = 0, w_near = 0;
		int cur = 0;
		for (int step = 0; step < n - 1; ++step) {
			// ---- the node `cur` joins the tree: retire it, then push its weights to the pixels around it ---------------
			const uint32_t cv = pk[cur];
			const uint32_t cxy = pxy[cur];
			const uint32_t cc = cv & 0x1FFFFu;
			__syncwarp();
			if (lane == 0) {
				pk[cur] = cc;       // key field 0: in the tree, no update can pass `mr < key`
				alive[cur >> 5] &= ~(1u << (cur & 31));
			}
			__syncwarp();
			const int cy = (int)(cxy >> 8), cx = (int)(cxy & 0xFFu);
			uint32_t newlv = 0u;
#pragma unroll
			for (int it = 0; it < kFrIters; ++it) {
				const int y = cy + ody[it], x = cx + odx[it];
				const bool inb = ((unsigned)y < (unsigned)H) && ((unsigned)x < (unsigned)W);
				const int wi = y * RS + (x >> 5);
				const uint32_t word = inb ? occ[wi] : 0u;
				const uint32_t bit = 1u << (x & 31);
				if (word & bit) {
					const int j = (int)wbase[wi] + __popc(word & (bit - 1u));
					const uint32_t v = pk[j];
					const uint32_t cj = v & 0x1FFFFu, kj = v >> 17;
					const uint32_t mr = max(max(od2[it], cc), cj);
					if (mr < min(kj, (uint32_t)kFrR0 + 1u)) {
						pk[j] = cj | (mr << 17);
						const uint32_t jb = 1u << (j & 31);
						const int jw = j >> 5;
						if (kj != kFrKeyInf) atomicAnd(&bm[fr_level(kj) * NWORDS + jw], ~jb);
						const int nl = fr_level(mr);
						atomicOr(&bm[nl * NWORDS + jw], jb);
						newlv |= 1u << nl;
						if (a.work != nullptr) ++w_near;
					}
				}
			}
			lvl_any |= __reduce_or_sync(0xffffffffu, newlv);
			__syncwarp();
			// ---- the next node: first set bit of the first non-empty bucket ----------------------------------------------
			uint32_t g = 0xFFFFFFFFu;   // (weight << 13) | index
			while (lvl_any != 0u) {
				const int l = __ffs(lvl_any) - 1;
				const uint32_t *b = bm + l * NWORDS;
				uint32_t idx = 0xFFFFFFFFu;
#pragma unroll
				for (int k = WPL - 1; k >= 0; --k) {
					const int w = lane + 32 * k;
					if (w < NWORDS) {
						const uint32_t mm = b[w];
						if (mm) idx = (uint32_t)(w * 32 + __ffs(mm) - 1);
					}
				}
				idx = __reduce_min_sync(0xffffffffu, idx);
				if (idx != 0xFFFFFFFFu) {
					if (lane == 0) bm[l * NWORDS + (idx >> 5)] &= ~(1u << (idx & 31));
					g = (c_froffs.level_w[l] << kKeyShift) | idx;
					break;
				}
				lvl_any &= ~(1u << l);
			}
			if (g == 0xFFFFFFFFu) {
				// ---- stall: every edge that is left is longer than kFrR0.  Complete the keys with the tree nodes that
				// joined since the last stall (all of them at the first one) and take the exact minimum.
				++w_stalls;
				int cnt = 0;
				for (int wb = 0; wb < NWORDS; wb += 32) {
					const int w = wb + lane;
					uint32_t mm = (w < NWORDS) ? alive[w] : 0u;
					const int c = __popc(mm);
					int inc = c;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) {
						const int t = __shfl_up_sync(0xffffffffu, inc, o);
						if (lane >= o) inc += t;
					}
					int pos = cnt + inc - c;
					while (mm) {
						const int bpos = __ffs(mm) - 1;
						mm &= mm - 1u;
						alist[pos++] = (uint32_t)(w * 32 + bpos);
					}
					cnt += __shfl_sync(0xffffffffu, inc, 31);
				}
				__syncwarp();
				uint32_t best = 0xFFFFFFFFu;
				for (int b0 = 0; b0 < cnt; b0 += 32 * kFrSyncSlots) {
					uint32_t sxy[kFrSyncSlots], sc[kFrSyncSlots], sf[kFrSyncSlots];
#pragma unroll
					for (int s = 0; s < kFrSyncSlots; ++s) {
						const int e = b0 + s * 32 + lane;
						sxy[s] = 0u; sc[s] = 0x1FFFFu; sf[s] = 0x3FFFFu;
						if (e < cnt) {
							const int j = (int)alist[e];
							sxy[s] = pxy[j];
							sc[s] = pk[j] & 0x1FFFFu;
							if (far_valid) sf[s] = far[j];
						}
					}
					for (int tw = 0; tw < NWORDS; ++tw) {
						uint32_t tm = ~alive[tw] & ~synced[tw];
						if (tw * 32 + 32 > n) tm &= (n > tw * 32) ? ((1u << (n - tw * 32)) - 1u) : 0u;
						while (tm) {
							const int t = tw * 32 + __ffs(tm) - 1;
							tm &= tm - 1u;
							const uint32_t txy = pxy[t];
							const uint32_t tc = pk[t] & 0x1FFFFu;
#pragma unroll
							for (int s = 0; s < kFrSyncSlots; ++s) {
								const uint32_t ad = __vabsdiffu4(sxy[s], txy);
								const uint32_t d2 = __dp4a(ad, ad, 0u);
								sf[s] = min(sf[s], max(d2, max(sc[s], tc)));
							}
							if (a.work != nullptr) w_far += (unsigned long long)min(kFrSyncSlots, (cnt - b0 + 31) >> 5);
						}
					}
#pragma unroll
					for (int s = 0; s < kFrSyncSlots; ++s) {
						const int e = b0 + s * 32 + lane;
						if (e < cnt) {
							const uint32_t j = alist[e];
							far[j] = sf[s];
							best = min(best, (sf[s] << kKeyShift) | j);
						}
					}
				}
				for (int w = lane; w < NWORDS; w += 32) synced[w] = ~alive[w];
				far_valid = true;
				g = __reduce_min_sync(0xffffffffu, best);
				__syncwarp();
			}
			if (lane == 0) pkey_out[step] = g;
			cur = (int)(g & kKeyIdxMask);
		}
		if (lane == 0) {
			if (a.phase_cycles != nullptr) atomicAdd(&a.phase_cycles[3], (unsigned long long)(clock64() - t0));
			if (a.work != nullptr) {
				atomicAdd(&a.work[0], (unsigned long long)(n > 0 ? n - 1 : 0));
				atomicAdd(&a.work[1], w_stalls);
			}
		}
		if (a.work != nullptr) {
			// w_far is uniform (counted once per warp by lane 0 below); w_near is per lane
			w_near = __reduce_add_sync(0xffffffffu, (unsigned)w_near);
			if (lane == 0) { atomicAdd(&a.work[2], w_far); atomicAdd(&a.work[3], w_near); }
		}
		__syncwarp();
	}
}

}  // namespace rvb
