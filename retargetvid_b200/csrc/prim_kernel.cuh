// Prim on the mutual-reachability graph as its own launch (split pipeline: front -> prim_kernel -> back).
//
// Same arithmetic and tie rule as prim_segment in map_kernel.cuh (_linkage.pyx:97-112: start at point 0,
// np.argmin = lowest index among equal weights), but ONE map is owned by 1, 2 or 4 warps instead of a
// 256/512-thread CTA.  Per added node the all-pairs update costs ~6 instructions per 32 live points, while
// the argmin/broadcast around it costs ~30 instructions PER WARP: with 8-16 warps per map that overhead was
// 2-3x the update itself.  A map needs only 12 bytes of shared memory per point here (the 64-bit point
// record + the live-key list), so 12 warps' worth of maps stay resident per SM and hide each other's
// redux/shared-memory latency.
//
// In : scr_pinfo[off + j] = {core_j, (y << 8) | x}  (front kernel)
// Out: scr_pkey[off + s]  = (weight of the edge that added the (s+1)-th node << 13) | node index
#pragma once
#include "map_kernel.cuh"

namespace rvb {

struct PrimArgs {
	const int *list;       // maps of this size class
	const int *list_len;
	int *head;
	const MapOut *out;     // n_points
	const int *scr_off;
	const uint2 *scr_pinfo;
	uint32_t *scr_pkey;
	int cap;               // points per map this launch has shared memory for
	unsigned long long *phase_cycles;
};

// The owner of the node just added retires its register slot: key = "never again", core = "in the tree".
static_assert(kInTreeCore == 0x1FFFFu, "prim_retire.inc writes the literal");
#include "prim_retire.inc"

// slot counts a segment can run with: 1 2 3 4 6 8 12 16 20 ...
__host__ __device__ constexpr int prim_next_slots(int k) { return k < 4 ? k + 1 : k < 8 ? k + 2 : k + 4; }

// `nsteps` nodes are added; the live points sit in the first K of the KMAX register slots per thread.  Only this
// loop is specialised on K (fully unrolled, no per-slot predicate, all K updates independent of each other); the
// code that loads and compacts the slots is shared.
template <int NT, int KMAX, int K>
__device__ __forceinline__ void prim_steps(uint32_t (&pxy)[KMAX], uint32_t (&pc)[KMAX], uint32_t (&key)[KMAX],
											const uint32_t (&idc)[KMAX], const uint2 *pinfo, uint32_t *pk, uint32_t (*wmin)[32],
											int &cur, int step, const int step_end) {
	constexpr int NW = NT / 32;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint32_t cxy = pinfo[cur].y;
	uint32_t cc = pinfo[cur].x & 0x1FFFFu;
	const uint32_t a_info = smem_u32addr(pinfo);
	const uint32_t a_wm = smem_u32addr(&wmin[0][0]);
	const uint32_t a_wm_mine = a_wm + 4u * (uint32_t)warp, a_wm_lane = a_wm + 4u * (uint32_t)min(lane, NW - 1);
	for (; step < step_end; ++step, ++pk) {
		uint32_t best = 0xFFFFFFFFu;
#pragma unroll
		for (int i = 0; i < K; ++i) {
			const uint32_t ad = __vabsdiffu4(pxy[i], cxy);
			const uint32_t d2 = __dp4a(ad, ad, 0u);
			const uint32_t mr = max(d2, max(pc[i], cc));
			uint32_t k;
			asm("mad.lo.u32 %0, %1, 8192, %2;" : "=r"(k) : "r"(mr), "r"(idc[i]));
			key[i] = min(key[i], k);
			best = min(best, key[i]);
		}
		uint32_t g = __reduce_min_sync(0xffffffffu, best);
		if constexpr (NW > 1) {
			const uint32_t par = (uint32_t)(step & 1) * 128u;   // wmin[step & 1]
			sts_u32(a_wm_mine + par, g);
			__syncthreads();
			// every lane reads all NW warp minima with one or two broadcast vector loads: no second warp reduction on
			// the per-step critical path
			if constexpr (NW == 2) {
				uint32_t v0, v1;
				asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(a_wm + par) : "memory");
				g = min(v0, v1);
			} else if constexpr (NW == 4) {
				uint32_t v0, v1, v2, v3;
				asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a_wm + par) : "memory");
				g = min(min(v0, v1), min(v2, v3));
			} else if constexpr (NW == 8) {
				uint32_t v0, v1, v2, v3, v4, v5, v6, v7;
				asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a_wm + par) : "memory");
				asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7) : "r"(a_wm + par + 16u) : "memory");
				g = min(min(min(v0, v1), min(v2, v3)), min(min(v4, v5), min(v6, v7)));
			} else {
				g = lds_u32(a_wm_lane + par);   // lanes >= NW re-read the last entry: harmless for a minimum
				g = __reduce_min_sync(0xffffffffu, g);
			}
		}
		const uint32_t cu = g & kKeyIdxMask;
		cur = (int)cu;
		if (tid == 0) *pk = g;
		uint32_t ix, iy;
		asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ix), "=r"(iy) : "r"(a_info + 8u * cu) : "memory");
		const uint32_t so = ix >> 17;
		cxy = iy;
		cc = ix & 0x1FFFFu;
		if ((int)(so & 0x3FFu) == tid) prim_retire_slot(pc, key, (int)(so >> 10));
	}
}

template <int NT, int KMAX, int K = 1>
__device__ __forceinline__ void prim_steps_dispatch(int kslots, uint32_t (&pxy)[KMAX], uint32_t (&pc)[KMAX], uint32_t (&key)[KMAX],
													 const uint32_t (&idc)[KMAX], const uint2 *pinfo, uint32_t *pk,
													 uint32_t (*wmin)[32], int &cur, int step, const int step_end) {
	if constexpr (K >= KMAX) {
		prim_steps<NT, KMAX, KMAX>(pxy, pc, key, idc, pinfo, pk, wmin, cur, step, step_end);
	} else {
		if (kslots <= K) prim_steps<NT, KMAX, K>(pxy, pc, key, idc, pinfo, pk, wmin, cur, step, step_end);
		else prim_steps_dispatch<NT, KMAX, prim_next_slots(K)>(kslots, pxy, pc, key, idc, pinfo, pk, wmin, cur, step, step_end);
	}
}

// One segment: the live points (lkey: their keys, which carry the point index) are spread over `kslots` register
// slots per thread, `nsteps` nodes are added, and the survivors go back to lkey if another segment follows.
template <int NT, int KMAX>
__device__ __forceinline__ void prim_seg_warps(uint2 *pinfo, uint32_t *lkey, uint32_t *pkey_out, uint32_t (*wmin)[32],
												int *live_cnt, PrimState &st, int kslots, int nsteps, bool write_back) {
	constexpr int NW = NT / 32;
	const int tid = threadIdx.x;
	uint32_t pxy[KMAX], pc[KMAX], key[KMAX], idc[KMAX];
#pragma unroll
	for (int i = 0; i < KMAX; ++i) {
		const int pos = tid + i * NT;
		key[i] = 0xFFFFFFFFu;
		idc[i] = 0u;
		pxy[i] = 0u;
		pc[i] = kInTreeCore;
		if (pos < st.live) {
			const uint32_t k = lkey[pos];
			const uint32_t j = k & kKeyIdxMask;
			key[i] = k;
			idc[i] = j;
			const uint2 pi = pinfo[j];
			pxy[i] = pi.y;
			pc[i] = pi.x & 0x1FFFFu;
			pinfo[j].x = pc[i] | ((uint32_t)(tid | (i << 10)) << 17);
		}
	}
	if (tid == 0) *live_cnt = 0;
	if constexpr (NW == 1) __syncwarp(); else __syncthreads();
	int cur = st.cur;
	prim_steps_dispatch<NT, KMAX>(kslots, pxy, pc, key, idc, pinfo, pkey_out + st.step, wmin, cur, st.step, st.step + nsteps);
	if (write_back) {
#pragma unroll
		for (int i = 0; i < KMAX; ++i) {
			if (pc[i] != kInTreeCore) {
				const int p = atomicAdd(live_cnt, 1);
				lkey[p] = key[i];
			}
		}
	}
	st.step += nsteps;
	st.live -= nsteps;
	st.cur = cur;
	if constexpr (NW == 1) __syncwarp(); else __syncthreads();
}

// warps per SM the register budget allows: 24 slots -> 128 registers -> 16 warps, 16 slots -> 20, 12 slots -> 24
__host__ __device__ constexpr int prim_warps_per_sm(int kmax) { return kmax >= 24 ? 16 : kmax >= 16 ? 20 : 24; }

template <int NW, int KMAX>
__global__ void __launch_bounds__(32 * NW, prim_warps_per_sm(KMAX) / NW) prim_kernel(const PrimArgs a) {
	constexpr int NT = 32 * NW;
	extern __shared__ __align__(16) uint8_t psm[];
	__shared__ __align__(16) uint32_t wmin[2][32];
	__shared__ int s_map, s_live;
	uint2 *pinfo = reinterpret_cast<uint2 *>(psm);
	uint32_t *lkey = reinterpret_cast<uint32_t *>(psm + (size_t)8 * a.cap);
	const int tid = threadIdx.x;
	while (true) {
		if constexpr (NW == 1) __syncwarp(); else __syncthreads();
		if (tid == 0) {
			const int i = atomicAdd(a.head, 1);
			s_map = (i < *a.list_len) ? a.list[i] : -1;
		}
		if constexpr (NW == 1) __syncwarp(); else __syncthreads();
		const int m = s_map;
		if (m < 0) break;
		const long long t0 = clock64();
		const int n = a.out[m].n_points;
		const size_t off = (size_t)a.scr_off[m];
		const uint2 *gp = a.scr_pinfo + off;
		uint32_t *pkey_out = a.scr_pkey + off;
#pragma unroll 4
		for (int j = tid; j < n; j += NT) pinfo[j] = gp[j];
		// every point but the root, "not reached yet": the largest weight field, index in the low bits
		for (int j = tid + 1; j < n; j += NT) lkey[j - 1] = 0xFFFFE000u | (uint32_t)j;
		if constexpr (NW == 1) __syncwarp(); else __syncthreads();
		PrimState ps;
		ps.step = 0; ps.live = n - 1; ps.cur = 0;
		while (ps.step < n - 1) {
			// slots per thread this segment runs with, and the live count at which the next smaller size fits
			int kslots = (ps.live + NT - 1) / NT;
			int ks = 1, prev = 0;
			while (ks < kslots) { prev = ks; ks = prim_next_slots(ks); }
			int nsteps = (prev > 0) ? (ps.live - prev * NT) : ps.live;
			nsteps = min(nsteps, n - 1 - ps.step);
			const bool more = (ps.step + nsteps) < (n - 1);
			prim_seg_warps<NT, KMAX>(pinfo, lkey, pkey_out, wmin, &s_live, ps, kslots, nsteps, more);
		}
		if (a.phase_cycles != nullptr && tid == 0) atomicAdd(&a.phase_cycles[3], (unsigned long long)(clock64() - t0));
	}
}

}  // namespace rvb
