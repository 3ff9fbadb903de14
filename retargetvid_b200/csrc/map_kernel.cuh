// Fused per-map kernel: (normalise) -> threshold -> cut blend -> dominant-cluster filter
// (HDBSCAN, bit-exact with oracle/hdbscan_port.py) -> 5x5 closing -> centroid / coverage.
//
// Replaces, per saliency map: unisal/train.py:1270-1274 (a1), sc_threshold
// smartVidCrop.py:1050-1059 (a2), the raw-map sums of sc_compute_mean_sal :1304-1308 (a3), the
// max profiles of sc_border_detection :859-865 (a4), sc_clustering_filt :1062-1161 and the cut
// blend :2369-2373 (a6), sc_compute_cvrg_score :1310-1331 (a7), sc_find_center_of_mass
// :1163-1219 with its emptiness gate :2403 (a8).
//
// One CTA owns one map at a time (persistent CTAs pull map indices from a device work list).
// The map is staged in shared memory with one bulk-async (TMA) copy; after the salient pixels
// are compacted the same shared memory is re-used for the clustering state, and the map is
// rebuilt there for the closing.  Nothing but the input map, a 64-byte result record and (only
// for maps a successor blends with, or on request) the filtered map touches HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvb {

constexpr int kRingTableMax = 640;   // lattice offsets with d^2 <= kRingD2Max
constexpr int kRingCountMax = 128;
constexpr int kRingD2Max = 196;
constexpr int kKeyShift = 13;        // Prim key = (weight << 13) | point index
constexpr uint32_t kKeyIdxMask = (1u << kKeyShift) - 1;
constexpr uint32_t kInTreeCore = 0x1FFFFu;  // larger than any squared distance on a 256x256 lattice
constexpr int kLambdaBits = 46;      // oracle/hdbscan_port.py LAMBDA_FRAC_BITS
constexpr uint16_t kNone16 = 0xFFFFu;

struct RingTable {
	int8_t dy[kRingTableMax];
	int8_t dx[kRingTableMax];
	uint16_t ring_end[kRingCountMax];
	uint16_t ring_d2[kRingCountMax];
	int n_rings;
	int n_offsets;
};

// result record of one map
struct MapOut {
	double cx;            // centroid x (process px) or argmax x
	double cy;
	uint32_t raw_sum;     // sum of the raw uint8 map (mean saliency)
	int32_t n_points;     // non-zero pixels after threshold/blend
	int32_t n_clusters;   // -1: clustering skipped by the gates
	int32_t kept_points;  // non-zero pixels of the final map
	int32_t flags;        // bit0 empty, bit1 capacity overflow, bit2 core fallback used
	int32_t pad;
	double cvrg[8];       // coverage score per ratio
};

constexpr int kFlagEmpty = 1;
constexpr int kFlagOverflow = 2;
constexpr int kFlagCoreFallback = 4;
constexpr int kFlagClusterCapacity = 8;

// byte offsets into dynamic shared memory, computed on the host for (NMAX, H, WPS)
struct SmemLayout {
	int pts;       // u16[NMAX]   (y << 8) | x, row-major order
	int val;       // u8[NMAX]
	int small;     // u8[Hs * WSs] down-scaled map (resize_factor != 1), else unused
	int u_base;    // start of the overlaid region
	// phase 0/1/5 view of the overlaid region
	int map;       // u8[H * WPS]
	// clustering view
	int a4;        // u32[NMAX]  core  -> later Lp1 (u16) + R (u16)
	int order;     // u16[NMAX]
	int wp;        // u32[NMAX]  edge weights in Prim order
	int d4;        // u32[NMAX]  sort keys -> later cL (u16) + cR (u16)
	int rank;      // u16[NMAX]  rank of edge -> later relabel
	int pe;        // u16[NMAX]  parent edge of an edge node
	int pl;        // u16[NMAX]  parent edge of a leaf
	int queue;     // u16[NMAX]  BFS queue -> later cluster of a point
	int pnode;     // u16[NMAX]  node a point fell out at (aliases d4: cL/cR are dead after the BFS)
	int mask;      // u32[H * MW] occupancy bit mask (core distances); aliases d4.. (dead before the sort)
	int cl_stab;   // u64[NCMAX]
	int cl_birth;  // u64[NCMAX] lambda_fix at birth
	int cl_acc;    // u32[NCMAX] per-label weight (sum or max)
	int cl_parent; // u16[NCMAX]
	int cl_ch0;    // u16[NCMAX]
	int cl_ch1;    // u16[NCMAX]
	int cl_label;  // u16[NCMAX] label of a selected cluster / kNone16
	int cl_selanc; // u16[NCMAX] nearest selected ancestor-or-self
	int total;
	int nmax;
	int ncmax;
};

struct MapArgs {
	// input
	const uint8_t *maps_u8;    // [N][H][gstride] or null
	const float *maps_f32;     // [N][H][W] or null
	int H, W, WPS, gstride;
	// work list (persistent CTAs)
	const int *list;
	const int *list_len;
	int *head;
	int *ovf_list;             // maps that did not fit NMAX of this launch
	int *ovf_len;
	// per map
	const int *pred;           // slot (in filt) of the predecessor whose filtered map is blended in, or -1
	const int *store;          // slot (in filt) to write this map's filtered result to, or -1
	const int *map_clip;       // clip index of a map
	const uint8_t *chain_next; // 1: map m+1 blends with this map's result -> the same CTA continues with m+1
	uint8_t *filt;             // [slots][H][fstride]
	int fstride;
	MapOut *out;
	uint32_t *border_prof;     // [n_clips][H + W] max profiles, or null
	const int *cvrg_cfg;       // [n_clips][R][2] = (mode, window) or null
	int n_ratios;
	int32_t *labels_dbg;       // optional: labels of the single map being debugged
	unsigned long long *phase_cycles;  // optional [16]: SM cycles spent per phase, summed over CTAs (profiling aid)
	// resize_factor != 1 (smartVidCrop.py:1078-1084,1158,1184): cluster a down-scaled copy, scale the result back
	int resize_on, resize_type;    // type 1: INTER_LINEAR, 2: INTER_CUBIC, 3: INTER_NEAREST, 4: INTER_AREA by exactly 2 (what OpenCV makes of INTER_LINEAR at fx = 1/2)
	int Hs, Ws, WSs;               // small size and its shared-memory row stride
	double factor;
	const int16_t *rz;             // coefficient tables, offsets below (int16 entries)
	int rz_dx, rz_dy, rz_ux, rz_uy, rz_nx, rz_ny;   // down x/y: idx,a0,a1 ; up x/y: idx,a0,a1 ; nearest x/y: idx
	int rz_cx, rz_cy;                               // cubic down x/y: idx, w0..w3
	// params
	int t_threshold, clust_filt, mcs, min_samples, select_sum, op_close, com_km;
	// split pipeline (front -> prim_kernel -> back): per-point scratch in HBM, addressed by a per-map point offset
	int *cls_lists;            // front: kSplitClasses lists of cls_stride entries, filled by the front kernel
	int *cls_cnt;              // front: their lengths
	int cls_stride;
	int *scr_off;              // [N] point offset of a map's scratch (front writes, prim/back read)
	unsigned long long *scr_top;  // bump allocator (points)
	unsigned int scr_cap;      // points available (< 2^31)
	uint2 *scr_pinfo;          // {core distance, (y << 8) | x}
	uint8_t *scr_val;          // pixel values
	uint32_t *scr_pkey;        // Prim output: (edge weight << 13) | node, in the order the nodes were added
	int force_class;           // front: >= 0 puts every map into this size class (chain sets: few maps, latency matters)
	uint8_t *skip;             // front: 1 = an earlier map of this chain went to a monolithic launch, which walks the rest
	SmemLayout lay;
};

// split pipeline: point-count classes of the Prim / back launches
constexpr int kSplitClasses = 6;
__host__ __device__ constexpr int split_class_cap(int k) { return k == 0 ? 768 : k == 1 ? 1536 : k == 2 ? 2048 : k == 3 ? 3072 : k == 4 ? 4096 : 8192; }
constexpr int kModeMono = 0, kModeFront = 1, kModeBack = 2;

__constant__ RingTable c_rings;

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32addr(const void *p) {
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok = 0;
	while (!ok) {
		asm volatile(
			"{\n"
			".reg .pred p;\n"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
			"selp.u32 %0, 1, 0, p;\n"
			"}\n" : "=r"(ok) : "r"(smem_u32addr(bar)), "r"(parity) : "memory");
	}
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				 ::"r"(smem_u32addr(dst)), "l"(src), "r"(bytes), "r"(smem_u32addr(bar)) : "memory");
}

// shared-memory accesses through precomputed 32-bit addresses (the hot Prim loop must not rebuild generic pointers)
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

template <int NT>
__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t *scratch /*[32]*/) {
	v = __reduce_add_sync(0xffffffffu, v);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	__syncthreads();
	if (lane == 0) scratch[warp] = v;
	__syncthreads();
	uint32_t t = (lane < NT / 32) ? scratch[lane] : 0u;
	return __reduce_add_sync(0xffffffffu, t);
}

// exclusive scan of one value per thread across the block; returns (exclusive prefix, total)
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *scratch /*[32]*/, uint32_t &total) {
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += t;
	}
	__syncthreads();
	if (lane == 31) scratch[warp] = inc;
	__syncthreads();
	uint32_t wsum = (lane < NT / 32) ? scratch[lane] : 0u;
	uint32_t winc = wsum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
		if (lane >= o) winc += t;
	}
	total = __shfl_sync(0xffffffffu, winc, NT / 32 - 1);
	uint32_t wexcl = __shfl_sync(0xffffffffu, winc - wsum, warp);
	return wexcl + inc - v;
}

__device__ __forceinline__ uint64_t lambda_fix(uint32_t w) {
	return (1ull << kLambdaBits) / (uint64_t)w;
}

// ---------------------------------------------------------------------------------------------
// numpy's portable argsort (aquicksort_<double>, npysort/quicksort.cpp) on packed keys
// (weight << kKeyShift | edge).  Only the weight takes part in comparisons, exactly as
// np.argsort(weights) would see them; must reproduce oracle/hdbscan_port.numpy_aquicksort.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool key_less(uint32_t a, uint32_t b) { return (a >> kKeyShift) < (b >> kKeyShift); }

__device__ void np_aheapsort(uint32_t *t /* 0-based start of the range */, int n) {
	uint32_t *a = t - 1;  // 1-based view like the C source
	uint32_t tmp;
	int i, j, l;
	for (l = n >> 1; l > 0; --l) {
		tmp = a[l];
		for (i = l, j = l << 1; j <= n;) {
			if (j < n && key_less(a[j], a[j + 1])) j += 1;
			if (key_less(tmp, a[j])) {
				a[i] = a[j];
				i = j;
				j += j;
			} else {
				break;
			}
		}
		a[i] = tmp;
	}
	for (; n > 1;) {
		tmp = a[n];
		a[n] = a[1];
		n -= 1;
		for (i = 1, j = 2; j <= n;) {
			if (j < n && key_less(a[j], a[j + 1])) j++;
			if (key_less(tmp, a[j])) {
				a[i] = a[j];
				i = j;
				j += j;
			} else {
				break;
			}
		}
		a[i] = tmp;
	}
}

__device__ void np_aquicksort(uint32_t *t, int num) {
	if (num < 2) return;
	constexpr int kSmall = 15;  // partitions of >= 17 elements are split (numpy 2.x, verified against np.argsort)
	int stack_l[64], stack_r[64], stack_d[64];
	int sp = 0;
	int pl = 0, pr = num - 1;
	int cdepth = (31 - __clz(num)) * 2;
	while (true) {
		bool heap = false;
		if (cdepth < 0) {
			np_aheapsort(t + pl, pr - pl + 1);
			heap = true;
		}
		if (!heap) {
			while ((pr - pl) > kSmall) {
				int pm = pl + ((pr - pl) >> 1);
				uint32_t vl = t[pl], vm = t[pm], vr = t[pr];
				if (key_less(vm, vl)) { uint32_t s = vm; vm = vl; vl = s; }
				if (key_less(vr, vm)) { uint32_t s = vr; vr = vm; vm = s; }
				if (key_less(vm, vl)) { uint32_t s = vm; vm = vl; vl = s; }
				t[pl] = vl;
				t[pr] = vr;
				const uint32_t vp = vm;
				int pi = pl;
				int pj = pr - 1;
				// INTP_SWAP(*pm, *pj)
				t[pm] = t[pj];
				t[pj] = vp;
				for (;;) {
					uint32_t vi, vj;
					do { ++pi; vi = t[pi]; } while (key_less(vi, vp));
					do { --pj; vj = t[pj]; } while (key_less(vp, vj));
					if (pi >= pj) break;
					t[pi] = vj;
					t[pj] = vi;
				}
				int pk = pr - 1;
				uint32_t s = t[pi];
				t[pi] = t[pk];
				t[pk] = s;
				if (pi - pl < pr - pi) {
					stack_l[sp] = pi + 1;
					stack_r[sp] = pr;
					pr = pi - 1;
				} else {
					stack_l[sp] = pl;
					stack_r[sp] = pi - 1;
					pl = pi + 1;
				}
				stack_d[sp] = --cdepth;
				++sp;
			}
			for (int pi = pl + 1; pi <= pr; ++pi) {
				uint32_t vi = t[pi];
				int pj = pi;
				while (pj > pl) {
					uint32_t vk = t[pj - 1];
					if (!key_less(vi, vk)) break;
					t[pj] = vk;
					--pj;
				}
				t[pj] = vi;
			}
		}
		if (sp == 0) break;
		--sp;
		pl = stack_l[sp];
		pr = stack_r[sp];
		cdepth = stack_d[sp];
	}
}

// ---------------------------------------------------------------------------------------------
// The same permutation computed by the whole CTA.  The recursion tree of aquicksort_ is processed
// level by level, one warp per range: sub-ranges are independent, and one Hoare partition has a
// closed form -- the scan from the left stops at the positions (ascending) whose key is >= pivot,
// the scan from the right at the positions (descending) whose key is <= pivot, the r-th stops are
// exchanged while the left one is still left of the right one, and the final left pointer is
// min(next left stop, last right stop).  Ranges of <= 16 elements are insertion-sorted by numpy,
// i.e. sorted stably.  Only ranges numpy POPS from its stack check the depth budget (heapsort).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_stable_small_sort(uint32_t *t, int lo, int hi) {
	const int lane = threadIdx.x & 31;
	const int sz = hi - lo + 1;  // 2..16
	uint32_t x = 0;
	int rk = 0;
	if (lane < sz) {
		x = t[lo + lane];
		for (int m = 0; m < sz; ++m) {
			const uint32_t o = t[lo + m];
			rk += (key_less(o, x) || (!key_less(x, o) && m < lane)) ? 1 : 0;
		}
	}
	__syncwarp();
	if (lane < sz) t[lo + rk] = x;
	__syncwarp();
}

// returns the final pivot position
__device__ __forceinline__ int warp_partition(uint32_t *t, int pl, int pr, uint16_t *ilist, uint16_t *jlist) {
	const int lane = threadIdx.x & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	if (lane == 0) {
		const int pm = pl + ((pr - pl) >> 1);
		uint32_t vl = t[pl], vm = t[pm], vr = t[pr];
		if (key_less(vm, vl)) { uint32_t q = vm; vm = vl; vl = q; }
		if (key_less(vr, vm)) { uint32_t q = vr; vr = vm; vm = q; }
		if (key_less(vm, vl)) { uint32_t q = vm; vm = vl; vl = q; }
		t[pl] = vl;
		t[pr] = vr;
		t[pm] = t[pr - 1];
		t[pr - 1] = vm;
	}
	__syncwarp();
	const uint32_t vp = t[pr - 1];
	uint16_t *il = ilist + pl, *jl = jlist + pl;
	int nI = 0, nJ = 0;
	for (int base = pl + 1; base <= pr - 1; base += 32) {
		const int p = base + lane;
		const bool f = (p <= pr - 1) && !key_less(t[p], vp);
		const uint32_t m = __ballot_sync(0xffffffffu, f);
		if (f) il[nI + __popc(m & lt_mask)] = (uint16_t)p;
		nI += __popc(m);
	}
	for (int base = pr - 2; base >= pl; base -= 32) {
		const int p = base - lane;
		const bool f = (p >= pl) && !key_less(vp, t[p]);
		const uint32_t m = __ballot_sync(0xffffffffu, f);
		if (f) jl[nJ + __popc(m & lt_mask)] = (uint16_t)p;
		nJ += __popc(m);
	}
	__syncwarp();
	const int R = min(nI, nJ);
	int S = 0;  // number of exchanges: the prefix of r with il[r] < jl[r]
	for (int base = 0; base < R; base += 32) {
		const int r = base + lane;
		const bool f = (r < R) && (il[r] < jl[r]);
		S += __popc(__ballot_sync(0xffffffffu, f));
	}
	for (int r = lane; r < S; r += 32) {
		const int a = il[r], b = jl[r];
		const uint32_t q = t[a];
		t[a] = t[b];
		t[b] = q;
	}
	__syncwarp();
	const int nextI = (S < nI) ? (int)il[S] : 0x7fffffff;
	const int lastJ = (S >= 1) ? (int)jl[S - 1] : (pr - 1);
	const int pi = min(nextI, lastJ);
	if (lane == 0) {
		const uint32_t q = t[pi];
		t[pi] = t[pr - 1];
		t[pr - 1] = q;
	}
	__syncwarp();
	return pi;
}

template <int NT>
__device__ void np_aquicksort_block(uint32_t *t, int num, uint16_t *ilist, uint16_t *jlist, uint16_t *qbuf, int qcap,
									int *cnt /* shared int[2] */) {
	constexpr int NW = NT / 32;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (num < 2) return;
	if (num <= 16) {
		if (warp == 0) warp_stable_small_sort(t, 0, num - 1);
		__syncthreads();
		return;
	}
	// range lists: entries of 3 u16 = (pl, pr, depth + 256 | popped << 15)
	uint16_t *qa = qbuf, *qb = qbuf + 3 * qcap;
	if (threadIdx.x == 0) {
		qa[0] = 0;
		qa[1] = (uint16_t)(num - 1);
		qa[2] = (uint16_t)(((31 - __clz(num)) * 2 + 256) | 0x8000);
		cnt[0] = 1;
		cnt[1] = 0;
	}
	__syncthreads();
	int cur = 0;
	while (true) {
		const int ncur = cnt[cur];
		if (ncur == 0) break;
		uint16_t *qc = cur ? qb : qa, *qn = cur ? qa : qb;
		for (int e = warp; e < ncur; e += NW) {
			int pl = qc[3 * e], pr = qc[3 * e + 1];
			int depth = (int)(qc[3 * e + 2] & 0x7FFF) - 256;
			const bool popped = (qc[3 * e + 2] & 0x8000) != 0;
			if (popped && depth < 0) {
				if (lane == 0) np_aheapsort(t + pl, pr - pl + 1);
				__syncwarp();
				continue;
			}
			const int pi = warp_partition(t, pl, pr, ilist, jlist);
			--depth;
			// numpy pushes the larger part (popped later, checks the depth) and continues with the smaller
			int lo[2] = {pl, pi + 1}, hi[2] = {pi - 1, pr};
			const bool left_is_pushed = !(pi - pl < pr - pi);
			for (int k = 0; k < 2; ++k) {
				const int sz = hi[k] - lo[k] + 1;
				if (sz > 16) {
					if (lane == 0) {
						const int slot = atomicAdd(&cnt[cur ^ 1], 1);
						const bool pushed = (k == 0) ? left_is_pushed : !left_is_pushed;
						qn[3 * slot] = (uint16_t)lo[k];
						qn[3 * slot + 1] = (uint16_t)hi[k];
						qn[3 * slot + 2] = (uint16_t)((depth + 256) | (pushed ? 0x8000 : 0));
					}
				} else if (sz >= 2) {
					warp_stable_small_sort(t, lo[k], hi[k]);
				}
			}
		}
		__syncthreads();
		if (threadIdx.x == 0) cnt[cur] = 0;
		cur ^= 1;
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------------------------
// separable 5x5 max / min on a byte image held in shared memory (words of 4 pixels).
// OpenCV's default morphology border never wins, i.e. windows are clipped to the image.
// ---------------------------------------------------------------------------------------------
// The passes only visit the rectangle rows [ry0, ry1] x words [xw0, xw1] (the kept pixels' bounding box grown by
// the 2-pixel reach of the 5x5 window): everything outside is zero before and after a closing.
template <int NT, bool IS_MAX>
__device__ __forceinline__ void morph_pass_h(uint32_t *img, int H, int MWS, int ry0, int ry1, int xw0, int xw1) {
	// 1x5 along x, in place: one warp owns a row, loads everything it needs, then stores
	const uint32_t neutral = IS_MAX ? 0u : 0xFFFFFFFFu;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int y = ry0 + warp; y <= ry1; y += NT / 32) {
		uint32_t *row = img + y * MWS;
		uint32_t res[2];
#pragma unroll
		for (int k = 0; k < 2; ++k) {
			const int xw = xw0 + lane + 32 * k;
			uint32_t r = 0;
			if (xw <= xw1) {
				const uint32_t C = row[xw];
				const uint32_t P = (xw > 0) ? row[xw - 1] : neutral;
				const uint32_t N = (xw + 1 < MWS) ? row[xw + 1] : neutral;
				const uint32_t m2 = __funnelshift_l(P, C, 16);  // pixels x-2
				const uint32_t m1 = __funnelshift_l(P, C, 8);   // pixels x-1
				const uint32_t p1 = __funnelshift_r(C, N, 8);   // pixels x+1
				const uint32_t p2 = __funnelshift_r(C, N, 16);  // pixels x+2
				if (IS_MAX) r = __vmaxu4(__vmaxu4(__vmaxu4(m2, m1), __vmaxu4(p1, p2)), C);
				else r = __vminu4(__vminu4(__vminu4(m2, m1), __vminu4(p1, p2)), C);
			}
			res[k] = r;
		}
		__syncwarp();
#pragma unroll
		for (int k = 0; k < 2; ++k) {
			const int xw = xw0 + lane + 32 * k;
			if (xw <= xw1) row[xw] = res[k];
		}
	}
	__syncthreads();
}

template <int NT, bool IS_MAX>
__device__ __forceinline__ void morph_pass_v(uint32_t *img, int H, int MWS, int ry0, int ry1, int xw0, int xw1) {
	// 5x1 along y, in place: a thread owns a vertical segment of one word column; the two rows above
	// and below the segment are read before anybody writes, then a sliding window runs down it
	const uint32_t neutral = IS_MAX ? 0u : 0xFFFFFFFFu;
	const int ncols = xw1 - xw0 + 1, nrows = ry1 - ry0 + 1;
	const int nseg = max(1, NT / ncols);
	const int seg_rows = (nrows + nseg - 1) / nseg;
	const int xw = xw0 + threadIdx.x % ncols, seg = threadIdx.x / ncols;
	const int y0 = ry0 + seg * seg_rows, y1 = min(ry1 + 1, y0 + seg_rows);
	const bool active = (seg < nseg) && (y0 <= ry1);
	auto at = [&](int y) -> uint32_t { return (y >= 0 && y < H) ? img[y * MWS + xw] : neutral; };
	uint32_t a2 = neutral, a1 = neutral, b1 = neutral, b2 = neutral;
	if (active) { a2 = at(y0 - 2); a1 = at(y0 - 1); b1 = at(y1); b2 = at(y1 + 1); }
	__syncthreads();
	if (active) {
		// window registers: m2, m1 (originals of the two previous rows), c, p1, p2
		uint32_t m2 = a2, m1 = a1;
		uint32_t c = img[y0 * MWS + xw];
		uint32_t p1 = (y0 + 1 < y1) ? img[(y0 + 1) * MWS + xw] : ((y0 + 1 == y1) ? b1 : b2);
		for (int y = y0; y < y1; ++y) {
			uint32_t p2;
			const int yy = y + 2;
			if (yy < y1) p2 = img[yy * MWS + xw];
			else if (yy == y1) p2 = b1;
			else if (yy == y1 + 1) p2 = b2;
			else p2 = neutral;
			uint32_t r;
			if (IS_MAX) r = __vmaxu4(__vmaxu4(__vmaxu4(m2, m1), __vmaxu4(p1, p2)), c);
			else r = __vminu4(__vminu4(__vminu4(m2, m1), __vminu4(p1, p2)), c);
			img[y * MWS + xw] = r;
			m2 = m1; m1 = c; c = p1; p1 = p2;
		}
	}
	__syncthreads();
}

// set the padding bytes (x >= W) of every row to `fill`
__device__ __forceinline__ void set_row_padding(uint32_t *img, int ry0, int ry1, int W, int MWS, uint32_t fill_byte, int NT) {
	const int first_pad_word = W >> 2;
	const int pad_words = MWS - first_pad_word;
	for (int i = threadIdx.x; i < (ry1 - ry0 + 1) * pad_words; i += NT) {
		const int y = ry0 + i / pad_words, k = i % pad_words;
		const int xw = first_pad_word + k;
		uint32_t keep_mask = 0u;
		if (xw == first_pad_word) {
			const int valid = W & 3;
			keep_mask = valid ? ((1u << (8 * valid)) - 1u) : 0u;
		}
		const uint32_t fill = fill_byte * 0x01010101u;
		uint32_t v = img[y * MWS + xw];
		img[y * MWS + xw] = (v & keep_mask) | (fill & ~keep_mask);
	}
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
struct MapScalars {
	unsigned long long mbar;
	uint32_t red[32];
	uint32_t wmin[2][32];
	int map_idx;
	int n;
	int ncl;
	int nsel;
	int maxcl;
	int n_clusters;
	int fb_count;
	int scr_slot;   // front: this map's offset in the scratch arrays, or -1
	int err;
	uint32_t rootminw;
	uint32_t root_edge;
	int sort_cnt[2];
	int bb[4];  // bounding box of the kept pixels: ymin, ymax, xmin, xmax
	unsigned long long sx, sy;
	uint32_t cnt, tot;
	uint32_t argmax_key;
};

#define RVB_PHASE(k)                                                              \
	do {                                                                          \
		if (a.phase_cycles != nullptr && threadIdx.x == 0) {                      \
			const long long now_ = clock64();                                     \
			atomicAdd(&a.phase_cycles[k], (unsigned long long)(now_ - phase_t0)); \
			phase_t0 = now_;                                                      \
		}                                                                         \
	} while (0)

// ---------------------------------------------------------------------------------------------
// Prim on the mutual-reachability graph, all pairs.  From point 0, lowest index wins ties (np.argmin) --
// _linkage.pyx:97-112.  key = (min reachability << 13) | index; |dx|,|dy| by one VABSDIFF4, dx^2+dy^2 by
// one IDP.4A, max(d2, core_j, core_u) by one VIMNMX3, one redux + one __syncthreads per added node.
//
// The points OUTSIDE the tree live in registers, K per thread; the run is cut into segments: whenever
// the live points fit into K-1 slots per thread they are compacted through shared memory (their keys
// carry the index) and the next segment runs a loop specialised for K-1 slots, so no instruction is
// spent on points already in the tree and the unrolled update carries no per-slot predicate.
// ---------------------------------------------------------------------------------------------
struct PrimState {
	int step;      // edges found so far
	int live;      // points outside the tree
	int cur;       // the node added last (its update has not been applied yet)
};

// pinfo[j] = { core_j | slot_code << 17, packed (x, y) }: everything a step needs about the node it just added
// comes from one 64-bit shared load.  slot_code = owner thread | register slot << 10.
template <int NT, int K>
__device__ __forceinline__ void prim_segment(uint2 *pinfo, uint16_t *order, uint32_t *wp,
											  uint32_t (*wmin)[32], const uint32_t *lkey_in, uint32_t *lkey_out,
											  int *live_cnt, PrimState &st, int nsteps, bool write_back) {
	constexpr int NW = NT / 32;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint32_t pxy[K], pc[K], key[K], idc[K];
#pragma unroll
	for (int i = 0; i < K; ++i) {
		const int pos = tid + i * NT;
		if (pos < st.live) {
			const uint32_t k = lkey_in[pos];
			const uint32_t j = k & kKeyIdxMask;
			key[i] = k;
			idc[i] = j;
			const uint2 pi = pinfo[j];
			pxy[i] = pi.y;
			pc[i] = pi.x & 0x1FFFFu;
			pinfo[j].x = pc[i] | ((uint32_t)(tid | (i << 10)) << 17);
		} else {
			key[i] = 0xFFFFFFFFu;
			idc[i] = 0u;
			pxy[i] = 0u;
			pc[i] = kInTreeCore;
		}
	}
	if (tid == 0) *live_cnt = 0;
	__syncthreads();
	int cur = st.cur;
	uint32_t cxy = pinfo[cur].y;
	uint32_t cc = pinfo[cur].x & 0x1FFFFu;
	const int step_end = st.step + nsteps;
	// 32-bit shared addresses, computed once
	const uint32_t a_info = smem_u32addr(pinfo), a_order = smem_u32addr(order), a_wp = smem_u32addr(wp);
	const uint32_t a_wm = smem_u32addr(&wmin[0][0]);
	const uint32_t a_wm_mine = a_wm + 4u * (uint32_t)warp, a_wm_lane = a_wm + 4u * (uint32_t)min(lane, NW - 1);
	for (int step = st.step; step < step_end; ++step) {
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
		best = __reduce_min_sync(0xffffffffu, best);
		const uint32_t par = (uint32_t)(step & 1) * 128u;   // wmin[step & 1]
		sts_u32(a_wm_mine + par, best);   // every lane holds the warp minimum: same address, same value, no branch
		__syncthreads();
		// lanes >= NW re-read the last entry: harmless for a minimum
		uint32_t g = lds_u32(a_wm_lane + par);
		g = __reduce_min_sync(0xffffffffu, g);
		const uint32_t cu = g & kKeyIdxMask;
		cur = (int)cu;
		if (warp == 0) {   // warp-uniform: all lanes store the same values
			sts_u16(a_order + 2u * (uint32_t)(step + 1), cu);
			sts_u32(a_wp + 4u * (uint32_t)step, g >> kKeyShift);
		}
		// the owner retires the new node
		uint32_t ix, iy;
		asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ix), "=r"(iy) : "r"(a_info + 8u * cu) : "memory");
		const uint32_t so = ix >> 17;
		cxy = iy;
		cc = ix & 0x1FFFFu;
		if ((int)(so & 0x3FFu) == tid) {
			const int slot = (int)(so >> 10);
#pragma unroll
			for (int i = 0; i < K; ++i)
				if (i == slot) { pc[i] = kInTreeCore; key[i] = 0xFFFFFFFFu; }
		}
	}
	if (write_back) {
		// hand the live points (key carries the index) to the next segment, in any order
#pragma unroll
		for (int i = 0; i < K; ++i) {
			if (pc[i] != kInTreeCore) {
				const int p = atomicAdd(live_cnt, 1);
				lkey_out[p] = key[i];
			}
		}
	}
	st.step = step_end;
	st.live -= nsteps;
	st.cur = cur;
	__syncthreads();
}

template <int NT, int TPT, int K = 1>
__device__ __forceinline__ void prim_dispatch(int ncnt, uint2 *pinfo, uint16_t *order, uint32_t *wp,
											   uint32_t (*wmin)[32], const uint32_t *lin, uint32_t *lout, int *live_cnt,
											   PrimState &st, int nsteps, bool write_back) {
	if constexpr (K >= TPT) {
		prim_segment<NT, TPT>(pinfo, order, wp, wmin, lin, lout, live_cnt, st, nsteps, write_back);
	} else {
		if (ncnt <= K) prim_segment<NT, K>(pinfo, order, wp, wmin, lin, lout, live_cnt, st, nsteps, write_back);
		else prim_dispatch<NT, TPT, K + 1>(ncnt, pinfo, order, wp, wmin, lin, lout, live_cnt, st, nsteps, write_back);
	}
}

// ---------------------------------------------------------------------------------------------
// cv2.resize(..., INTER_LINEAR) for uint8 (OpenCV resize.cpp): horizontal pass in int with weights
// scaled by 2048, vertical pass ((b * (S >> 4)) >> 16), + 2, >> 2, rows clipped.  Tables come from the
// host (float32 source coordinates, weights rounded half-to-even).  src/dst live in shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cv_resize_linear_u8(const uint8_t *src, int sH, int sW, int sstride, uint8_t *dst, int dH,
													 int dW, int dstride, const int16_t *tx, const int16_t *ty, int NT) {
	// tx: idx[dW], a0[dW], a1[dW]; ty: idx[dH], b0[dH], b1[dH]
	for (int i = threadIdx.x; i < dH * dW; i += NT) {
		const int y = i / dW, x = i - y * dW;
		const int xi = tx[x], a0 = tx[dW + x], a1 = tx[2 * dW + x];
		const int yi = ty[y], b0 = ty[dH + y], b1 = ty[2 * dH + y];
		const int x1 = min(xi + 1, sW - 1);
		const int y0 = min(max(yi, 0), sH - 1), y1 = min(max(yi + 1, 0), sH - 1);
		const int r0 = ((int)src[y0 * sstride + xi] * a0 + (int)src[y0 * sstride + x1] * a1) >> 4;
		const int r1 = ((int)src[y1 * sstride + xi] * a0 + (int)src[y1 * sstride + x1] * a1) >> 4;
		dst[y * dstride + x] = (uint8_t)((((b0 * r0) >> 16) + ((b1 * r1) >> 16) + 2) >> 2);
	}
}

// cv2.resize(..., INTER_CUBIC) for uint8 (OpenCV resize.cpp): 4 taps with replicated borders, weights scaled by 2048;
// horizontal pass in int, vertical pass in float32 as VResizeCubicVec_32s8u does it -- S0*b0 + (S1*b1 + (S2*b2 + S3*b3)),
// every product and sum rounded separately (no FMA), b = weight / 2^22 -- then round half to even and saturate.
__device__ __forceinline__ void cv_resize_cubic_u8(const uint8_t *src, int sH, int sW, int sstride, uint8_t *dst, int dH,
													int dW, int dstride, const int16_t *tx, const int16_t *ty, int NT) {
	// tx: idx[dW], w0[dW] .. w3[dW]; ty: idx[dH], w0[dH] .. w3[dH]
	for (int i = threadIdx.x; i < dH * dW; i += NT) {
		const int y = i / dW, x = i - y * dW;
		const int xi = tx[x], yi = ty[y];
		int xs[4], aw[4];
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			xs[k] = min(max(xi - 1 + k, 0), sW - 1);
			aw[k] = tx[(1 + k) * dW + x];
		}
		float acc = 0.f;
#pragma unroll
		for (int k = 3; k >= 0; --k) {
			const uint8_t *row = src + min(max(yi - 1 + k, 0), sH - 1) * sstride;
			const int r = (int)row[xs[0]] * aw[0] + (int)row[xs[1]] * aw[1] + (int)row[xs[2]] * aw[2] + (int)row[xs[3]] * aw[3];
			const float b = __fmul_rn((float)ty[(1 + k) * dH + y], 1.f / 4194304.f);
			const float pr = __fmul_rn((float)r, b);
			acc = (k == 3) ? pr : __fadd_rn(pr, acc);
		}
		const int v = __float2int_rn(acc);
		dst[y * dstride + x] = (uint8_t)min(max(v, 0), 255);
	}
}

// ---------------------------------------------------------------------------------------------
// Streaming variant of the map stage for clust_filt == 0 (smartVidCrop.py:2354: "Skipping clustering"):
// threshold (a2), raw mean saliency (a3) and centre of mass (a8) are one pass over the map, so there
// is nothing to keep on chip.  One warp per map, 128-bit loads straight from HBM, SIMD byte compares,
// dot products for the coordinate sums, five warp reductions, one 96-byte record out.  HBM-bound.
// ---------------------------------------------------------------------------------------------
constexpr int kStreamBatch = 9;   // 140 rows x 16 chunks = 2240 chunks = 8.75 per thread: one batch per map

template <bool ARGMAX>
__global__ void __launch_bounds__(256) map_stream_kernel(const uint8_t *__restrict__ maps, int n_maps, int H, int W, int gstride,
														  int t_threshold, MapOut *__restrict__ out) {
	__shared__ uint32_t part[8][5];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	// a pixel survives the threshold (and is then non-zero) iff value >= max(t, 1)
	const uint32_t thr4 = (uint32_t)min(max(t_threshold, 1), 255) * 0x01010101u;
	const bool none = t_threshold > 255;
	const int cpr = gstride >> 4;                      // 16-byte chunks per row (gstride is a multiple of 16)
	const int n_chunks = H * cpr;
	// thread -> (row offset, chunk in row) is fixed when 256 is a multiple of the chunks per row
	const bool regular = (256 % cpr) == 0;
	const int rows_per_iter = regular ? (256 / cpr) : 0;
	int y_t = regular ? (tid / cpr) : 0;
	const int xb_t = regular ? ((tid % cpr) << 4) : 0;
	uint32_t pmask[4];                                 // row padding (x >= W) is the caller's garbage
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const int valid = min(max(W - (xb_t + 4 * k), 0), 4);
		pmask[k] = (valid >= 4) ? 0xFFFFFFFFu : ((valid > 0) ? ((1u << (8 * valid)) - 1u) : 0u);
	}
	// one CTA per map (persistent, static stride): 256 threads x 16 B keep the whole map in flight
	for (int m = blockIdx.x; m < n_maps; m += gridDim.x) {
		const int4 *src = reinterpret_cast<const int4 *>(maps + (size_t)m * H * gstride);
		uint32_t raw = 0, cnt = 0, sx = 0, sy = 0, amax = 0;
		int y = y_t;
		for (int c0 = tid; c0 < n_chunks; c0 += 256 * kStreamBatch) {
			// issue a whole batch of 128-bit loads before touching any of them
			int4 qb[kStreamBatch];
#pragma unroll
			for (int j = 0; j < kStreamBatch; ++j) {
				const int c = c0 + 256 * j;
				qb[j] = (c < n_chunks) ? __ldcs(src + c) : make_int4(0, 0, 0, 0);   // streamed once: evict-first
			}
#pragma unroll
			for (int j = 0; j < kStreamBatch; ++j) {
				const int c = c0 + 256 * j;
				const int4 q = qb[j];
				int xb = xb_t;
				uint32_t m4[4] = {pmask[0], pmask[1], pmask[2], pmask[3]};
				if (!regular) {
					y = c / cpr;
					xb = (c - y * cpr) << 4;
#pragma unroll
					for (int k = 0; k < 4; ++k) {
						const int valid = min(max(W - (xb + 4 * k), 0), 4);
						m4[k] = (valid >= 4) ? 0xFFFFFFFFu : ((valid > 0) ? ((1u << (8 * valid)) - 1u) : 0u);
					}
				}
				const uint32_t w4[4] = {(uint32_t)q.x & m4[0], (uint32_t)q.y & m4[1], (uint32_t)q.z & m4[2], (uint32_t)q.w & m4[3]};
				uint32_t c4 = 0, xl = 0;
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					raw += __vsadu4(w4[k], 0u);
					const uint32_t nz = none ? 0u : (__vcmpgeu4(w4[k], thr4) & 0x01010101u);
					c4 = __dp4a(nz, 0x01010101u, c4);
					xl = __dp4a(nz, 0x03020100u + 0x04040404u * (uint32_t)k, xl);  // x offsets 4k .. 4k+3 inside the chunk
					if (ARGMAX) {
#pragma unroll
						for (int b = 0; b < 4; ++b) {
							const uint32_t bv = (w4[k] >> (8 * b)) & 0xFFu;
							if ((nz >> (8 * b)) & 1u) amax = max(amax, (bv << 20) | (0xFFFFFu - (uint32_t)(y * W + xb + 4 * k + b)));
						}
					}
				}
				cnt += c4;
				sx += (uint32_t)xb * c4 + xl;
				sy += (uint32_t)y * c4;
				y += rows_per_iter;
			}
		}
		raw = __reduce_add_sync(0xffffffffu, raw);
		cnt = __reduce_add_sync(0xffffffffu, cnt);
		sx = __reduce_add_sync(0xffffffffu, sx);
		sy = __reduce_add_sync(0xffffffffu, sy);
		if (ARGMAX) amax = __reduce_max_sync(0xffffffffu, amax);
		__syncthreads();  // the previous map's partials have been consumed
		if (lane == 0) { part[warp][0] = raw; part[warp][1] = cnt; part[warp][2] = sx; part[warp][3] = sy; part[warp][4] = amax; }
		__syncthreads();
		if (warp == 0) {
			uint32_t v[5];
#pragma unroll
			for (int k = 0; k < 5; ++k) v[k] = (lane < 8) ? part[lane][k] : 0u;
#pragma unroll
			for (int k = 0; k < 4; ++k) v[k] = __reduce_add_sync(0xffffffffu, v[k]);
			v[4] = __reduce_max_sync(0xffffffffu, v[4]);
			if (lane == 0) {
				MapOut r;
				r.cx = 0.0; r.cy = 0.0; r.raw_sum = v[0]; r.n_points = (int)v[1]; r.n_clusters = -1; r.kept_points = (int)v[1];
				r.flags = 0; r.pad = 0;
#pragma unroll
				for (int k = 0; k < 8; ++k) r.cvrg[k] = 0.0;
				if (v[1] == 0u) r.flags = kFlagEmpty;
				else if (!ARGMAX) { r.cx = (double)v[2] / (double)v[1]; r.cy = (double)v[3] / (double)v[1]; }
				else {
					const uint32_t lin = 0xFFFFFu - (v[4] & 0xFFFFFu);
					r.cx = (double)(lin % (uint32_t)W);
					r.cy = (double)(lin / (uint32_t)W);
				}
				out[m] = r;
			}
		}
	}
}

// resident CTAs per SM the register budget must allow, per capacity class (shared memory bounds the same)
template <int NT, int TPT, int MODE>
struct MapKernelCfg {
	// (the monolithic kernel keeps Prim's point registers and the tree code's temporaries alive together: 64 registers)
	static constexpr int kMinBlocks = (MODE == kModeFront) ? (NT <= 256 ? 4 : 2) : (NT * TPT <= 1536) ? (MODE == kModeMono ? 4 : 5) : (NT * TPT <= 2048) ? 4 : (NT * TPT <= 4096) ? 2 : 1;
};

// MODE kModeMono: the whole path for one map after the other (cut-adjacent chains, the largest classes, resize).
// MODE kModeFront: load .. core distances, then the points go to HBM scratch and the map to a Prim class list.
// MODE kModeBack: points and Prim result come back from scratch; sort .. results.
template <int NT, int TPT, int MODE = kModeMono>
__global__ void __launch_bounds__(NT, MapKernelCfg<NT, TPT, MODE>::kMinBlocks) map_kernel(const MapArgs a) {
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ MapScalars S;
	const SmemLayout &L = a.lay;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int NW = NT / 32;
	const int H = a.H, W = a.W, WPS = a.WPS, MWS = WPS >> 2;
	const int n_words = H * MWS;
	// rows are usually 64 words (256 bytes): index -> (row, word) without an integer division
	const int mws_shift = ((MWS & (MWS - 1)) == 0) ? (31 - __clz(MWS)) : -1;

	uint16_t *pts = reinterpret_cast<uint16_t *>(smem + L.pts);
	uint8_t *val = smem + L.val;
	uint8_t *map8 = smem + L.map;
	uint32_t *map32 = reinterpret_cast<uint32_t *>(smem + L.map);

	if (tid == 0) {
		mbar_init(reinterpret_cast<uint64_t *>(&S.mbar), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	uint32_t tma_parity = 0;

	int m = -1;
	while (true) {
		// ---- fetch the next map (or continue a cut-adjacent chain) ---------------------------------
		__syncthreads();
		if (m < 0) {
			if (tid == 0) {
				const int i = atomicAdd(a.head, 1);
				S.map_idx = (i < *a.list_len) ? a.list[i] : -1;
			}
			__syncthreads();
			m = S.map_idx;
			if (m < 0) break;
			if constexpr (MODE == kModeFront) {
				if (a.skip != nullptr && a.skip[m]) { m = -1; continue; }
			}
		}
		long long phase_t0 = clock64();
		MapOut res;
		res.cx = 0.0; res.cy = 0.0; res.raw_sum = 0; res.n_points = 0; res.n_clusters = -1;
		res.kept_points = 0; res.flags = 0; res.pad = 0;
#pragma unroll
		for (int r = 0; r < 8; ++r) res.cvrg[r] = 0.0;

		uint8_t *img8 = map8;
		uint32_t *img32 = map32;
		int LH = H, LW = W, LWPS = WPS;
		int LMWS = LWPS >> 2, ln_words = LH * LMWS;
		bool rz = false;
		int n = 0;
		if constexpr (MODE != kModeBack) {
		// ---- phase 0: stage the map in shared memory --------------------------------------------
		if (a.maps_u8 != nullptr) {
			const uint8_t *src = a.maps_u8 + (size_t)m * H * a.gstride;
			if (a.gstride == WPS) {
				if (tid == 0) {
					// earlier generic-proxy accesses to this shared memory must be ordered before the
					// async-proxy write of the bulk copy
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
					mbar_expect_tx(reinterpret_cast<uint64_t *>(&S.mbar), (uint32_t)(H * WPS));
					tma_bulk_g2s(map8, src, (uint32_t)(H * WPS), reinterpret_cast<uint64_t *>(&S.mbar));
				}
				mbar_wait(reinterpret_cast<uint64_t *>(&S.mbar), tma_parity);
				tma_parity ^= 1u;
			} else {
				for (int i = tid; i < n_words; i += NT) {
					const int y = i / MWS, xw = i - y * MWS;
					uint32_t v = 0;
					if (xw * 4 < a.gstride) v = *reinterpret_cast<const uint32_t *>(src + (size_t)y * a.gstride + xw * 4);
					map32[i] = v;
				}
			}
		} else {
			// a1: u8 = trunc( exp(logp) / max(exp(logp)) * 255 ), fp32 arithmetic, exp correctly rounded
			const float *src = a.maps_f32 + (size_t)m * H * W;
			const int n_px = H * W;
			float mx = -INFINITY;
			for (int i = tid * 4; i < n_px; i += NT * 4) {
				if (i + 3 < n_px) {
					const float4 v = *reinterpret_cast<const float4 *>(src + i);
					mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
				} else {
					for (int k = i; k < n_px; ++k) mx = fmaxf(mx, src[k]);
				}
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
			__syncthreads();
			if (lane == 0) S.red[warp] = __float_as_uint(mx);
			__syncthreads();
			mx = __uint_as_float(S.red[0]);
			for (int wv = 1; wv < NW; ++wv) mx = fmaxf(mx, __uint_as_float(S.red[wv]));
			const float emax = (float)exp((double)mx);
			for (int i = tid; i < n_words; i += NT) map32[i] = 0u;
			__syncthreads();
			for (int i = tid; i < n_px; i += NT) {
				const float e = (float)exp((double)src[i]);
				const float q = __fmul_rn(__fdiv_rn(e, emax), 255.0f);
				const int y = i / W, x = i - y * W;
				map8[y * WPS + x] = (uint8_t)(int)q;
			}
		}
		__syncthreads();

		RVB_PHASE(0);  // load
		// raw statistics, threshold, blend -- one pass over the words
		{
			const int pred = a.pred ? a.pred[m] : -1;
			const uint8_t *pf = (pred >= 0) ? (a.filt + (size_t)pred * H * a.fstride) : nullptr;
			const uint32_t thr4 = (uint32_t)a.t_threshold * 0x01010101u;
			const int first_pad_word = W >> 2;
			uint32_t raw = 0, post = 0;
			for (int i = tid; i < n_words; i += NT) {
				const int y = (mws_shift >= 0) ? (i >> mws_shift) : (i / MWS), xw = i - y * MWS;
				uint32_t v = map32[i];
				if (xw >= first_pad_word) {
					const int valid = (xw == first_pad_word) ? (W & 3) : 0;
					v &= valid ? ((1u << (8 * valid)) - 1u) : 0u;
				}
				raw += __vsadu4(v, 0u);
				if (a.t_threshold > 255) v = 0u;
				else v &= __vcmpgeu4(v, thr4);
				if (pf != nullptr) {
					uint32_t f = 0u;
					if (xw * 4 < a.fstride) f = *reinterpret_cast<const uint32_t *>(pf + (size_t)y * a.fstride + xw * 4);
					if (xw >= first_pad_word) {
						const int valid = (xw == first_pad_word) ? (W & 3) : 0;
						f &= valid ? ((1u << (8 * valid)) - 1u) : 0u;
					}
					// uint8 addition wraps, then /2 truncates (smartVidCrop.py:2371-2373)
					v = (__vadd4(v, f) >> 1) & 0x7F7F7F7Fu;
				}
				map32[i] = v;
				post += __vsadu4(v, 0u);
			}
			// border max profiles of the RAW map are taken before the threshold; they need a second
			// look at the raw values, so they are handled by border_profile_kernel on request.
			res.raw_sum = block_sum_u32<NT>(raw, S.red);
			post = block_sum_u32<NT>(post, S.red);
			if (tid == 0) S.tot = post;
		}
		__syncthreads();

		// ---- resize_factor != 1: the clustering runs on a down-scaled copy (smartVidCrop.py:1078-1084);
		// an all-zero map is returned untouched (:1064)
		// (with clust_filt off sc_clustering_filt is never called, smartVidCrop.py:2354-2366: no resize round trip; the
		// centroid stage below still samples by the same factor)
		rz = a.resize_on && a.clust_filt && (S.tot != 0u);
		if (rz) {
			uint8_t *small8 = smem + L.small;
			for (int i = tid; i < (a.Hs * a.WSs) >> 2; i += NT) reinterpret_cast<uint32_t *>(small8)[i] = 0u;
			__syncthreads();
			if (a.resize_type == 1) {
				cv_resize_linear_u8(map8, H, W, WPS, small8, a.Hs, a.Ws, a.WSs, a.rz + a.rz_dx, a.rz + a.rz_dy, NT);
			} else if (a.resize_type == 2) {
				cv_resize_cubic_u8(map8, H, W, WPS, small8, a.Hs, a.Ws, a.WSs, a.rz + a.rz_cx, a.rz + a.rz_cy, NT);
			} else if (a.resize_type == 4) {
				// resizeAreaFast_ (OpenCV resize.cpp), scale 2: full cells (sum + 2) >> 2; the partial cells of an odd size
				// saturate_cast<uchar>(float(sum) / count) = round half to even
				for (int i = tid; i < a.Hs * a.Ws; i += NT) {
					const int y = i / a.Ws, x = i - y * a.Ws;
					const int y0 = 2 * y, x0 = 2 * x;
					uint32_t v = 0u;
					if (y0 < H && x0 < W) {
						const bool y1 = (y0 + 1) < H, x1 = (x0 + 1) < W;
						uint32_t sum = map8[y0 * WPS + x0];
						if (x1) sum += map8[y0 * WPS + x0 + 1];
						if (y1) sum += map8[(y0 + 1) * WPS + x0];
						if (x1 && y1) sum += map8[(y0 + 1) * WPS + x0 + 1];
						if (x1 && y1) v = (sum + 2u) >> 2;
						else if (x1 || y1) v = (sum >> 1) + ((sum & 1u) & ((sum >> 1) & 1u));   // sum / 2, half to even
						else v = sum;
					}
					small8[y * a.WSs + x] = (uint8_t)v;
				}
			} else {
				const int16_t *nx = a.rz + a.rz_nx, *ny = a.rz + a.rz_ny;
				for (int i = tid; i < a.Hs * a.Ws; i += NT) {
					const int y = i / a.Ws, x = i - y * a.Ws;
					small8[y * a.WSs + x] = map8[ny[y] * WPS + nx[x]];
				}
			}
			__syncthreads();
			img8 = small8;
			img32 = reinterpret_cast<uint32_t *>(small8);
			LH = a.Hs; LW = a.Ws; LWPS = a.WSs;
		}
		LMWS = LWPS >> 2;
		ln_words = LH * LMWS;

		// ---- phase 1: compact the non-zero pixels in row-major order ----------------------------
		{
			const int chunk = (ln_words + NT - 1) / NT;
			const int w0 = tid * chunk, w1 = min(ln_words, w0 + chunk);
			uint32_t c = 0;
			for (int i = w0; i < w1; ++i) c += __popc(__vcmpne4(img32[i], 0u)) >> 3;
			uint32_t total;
			uint32_t base = block_excl_scan<NT>(c, S.red, total);
			n = (int)total;
			if (a.clust_filt && n <= L.nmax) {
				for (int i = w0; i < w1; ++i) {
					uint32_t v = img32[i];
					if (v == 0u) continue;
					const int y = i / LMWS, x0 = (i - y * LMWS) * 4;
#pragma unroll
					for (int b = 0; b < 4; ++b) {
						const uint32_t bv = (v >> (8 * b)) & 0xFFu;
						if (bv) {
							pts[base] = (uint16_t)((y << 8) | (x0 + b));
							val[base] = (uint8_t)bv;
							++base;
						}
					}
				}
			}
		}
		res.n_points = n;
		__syncthreads();
		RVB_PHASE(1);  // threshold + compaction
		if (a.clust_filt && n > L.nmax) {
			// does not fit this launch's capacity: hand it to the next size class
			if (tid == 0) {
				if (a.ovf_list != nullptr) {
					const int k = atomicAdd(a.ovf_len, 1);
					a.ovf_list[k] = m;
					if constexpr (MODE == kModeFront) {
						// the monolithic launch walks the rest of this chain: later front launches leave it alone
						if (a.skip != nullptr) for (int q = m; a.chain_next[q]; ++q) a.skip[q + 1] = 1;
					}
				} else {
					res.flags = kFlagOverflow;
					a.out[m] = res;
				}
			}
			m = -1;  // a chain continues in the next capacity class, starting with this map
			continue;
		}
		} else {
			// ---- back half: the front kernel's record and points come back from scratch --------------
			res = a.out[m];
			n = res.n_points;
			const size_t off = (size_t)a.scr_off[m];
			for (int j = tid; j < n; j += NT) {
				pts[j] = (uint16_t)a.scr_pinfo[off + j].y;
				val[j] = a.scr_val[off + j];
			}
		}

		const bool do_cluster = (MODE == kModeBack) || (a.clust_filt && n > 0 && (n > a.mcs + 1));
		bool rebuilt = false;
		if (do_cluster) {
			uint32_t *core = reinterpret_cast<uint32_t *>(smem + L.a4);
			uint16_t *Lp1 = reinterpret_cast<uint16_t *>(smem + L.a4);
			uint16_t *Rr = Lp1 + L.nmax;
			uint16_t *order = reinterpret_cast<uint16_t *>(smem + L.order);
			uint32_t *wp = reinterpret_cast<uint32_t *>(smem + L.wp);
			uint32_t *skey = reinterpret_cast<uint32_t *>(smem + L.d4);
			uint16_t *cL = reinterpret_cast<uint16_t *>(smem + L.d4);
			uint16_t *cR = cL + L.nmax;
			uint16_t *rank = reinterpret_cast<uint16_t *>(smem + L.rank);
			uint16_t *relabel = rank;
			uint16_t *pe = reinterpret_cast<uint16_t *>(smem + L.pe);
			uint16_t *pl = reinterpret_cast<uint16_t *>(smem + L.pl);
			uint16_t *queue = reinterpret_cast<uint16_t *>(smem + L.queue);
			uint16_t *pcl = queue;
			uint16_t *pnode = reinterpret_cast<uint16_t *>(smem + L.pnode);
			uint32_t *mask = reinterpret_cast<uint32_t *>(smem + L.mask);
			unsigned long long *cl_stab = reinterpret_cast<unsigned long long *>(smem + L.cl_stab);
			unsigned long long *cl_birth = reinterpret_cast<unsigned long long *>(smem + L.cl_birth);
			uint32_t *cl_acc = reinterpret_cast<uint32_t *>(smem + L.cl_acc);
			uint16_t *cl_parent = reinterpret_cast<uint16_t *>(smem + L.cl_parent);
			uint16_t *cl_ch0 = reinterpret_cast<uint16_t *>(smem + L.cl_ch0);
			uint16_t *cl_ch1 = reinterpret_cast<uint16_t *>(smem + L.cl_ch1);
			uint16_t *cl_label = reinterpret_cast<uint16_t *>(smem + L.cl_label);
			uint16_t *cl_selanc = reinterpret_cast<uint16_t *>(smem + L.cl_selanc);
			const int MW = (LW + 31) >> 5;

			if constexpr (MODE != kModeBack) {
			// ---- phase 2: core distances on the lattice -------------------------------------------
			// k-th nearest OTHER salient pixel (hdbscan: sorted-row index min_samples, self at 0).
			int kk = (a.min_samples > 0) ? a.min_samples : a.mcs;
			kk = min(n - 1, kk);
			if (kk == 0) kk = 1;
			for (int i = tid; i < LH * MW; i += NT) mask[i] = 0u;
			if (tid == 0) S.fb_count = 0;
			__syncthreads();
			for (int p = tid; p < n; p += NT) {
				const int y = pts[p] >> 8, x = pts[p] & 0xFF;
				atomicOr(&mask[y * MW + (x >> 5)], 1u << (x & 31));
			}
			__syncthreads();
			for (int p0 = 0; p0 < n; p0 += NT) {
				const int p = p0 + tid;
				bool active = p < n;
				const int y = active ? (pts[p] >> 8) : 0, x = active ? (pts[p] & 0xFF) : 0;
				int cnt = 0, i = 0;
				uint32_t cd = 0;
				for (int r = 0; r < c_rings.n_rings; ++r) {
					const int e = c_rings.ring_end[r];
					if (active) {
						for (; i < e; ++i) {
							const int yy = y + c_rings.dy[i], xx = x + c_rings.dx[i];
							if ((unsigned)yy < (unsigned)LH && (unsigned)xx < (unsigned)LW)
								cnt += (mask[yy * MW + (xx >> 5)] >> (xx & 31)) & 1u;
						}
						if (cnt >= kk) {
							cd = c_rings.ring_d2[r];
							active = false;
						}
					}
					if (!__any_sync(0xffffffffu, active)) break;
				}
				if (p < n) {
					core[p] = cd;
					if (cd == 0) {  // not resolved inside the table radius
						const int k = atomicAdd(&S.fb_count, 1);
						queue[k] = (uint16_t)p;
					}
				}
			}
			__syncthreads();
			if (S.fb_count > 0) {
				// rare: sparse pixels.  One warp per pixel, bisection on the squared radius.
				res.flags |= kFlagCoreFallback;
				const int fbn = S.fb_count;
				for (int f = warp; f < fbn; f += NW) {
					const int p = queue[f];
					const int y = pts[p] >> 8, x = pts[p] & 0xFF;
					int lo = kRingD2Max + 1, hi = LH * LH + LW * LW;
					while (lo < hi) {
						const int mid = (lo + hi) >> 1;
						int c = 0;
						for (int j = lane; j < n; j += 32) {
							const int dy = (pts[j] >> 8) - y, dx = (pts[j] & 0xFF) - x;
							c += (dy * dy + dx * dx <= mid) ? 1 : 0;
						}
						c = __reduce_add_sync(0xffffffffu, c) - 1;  // minus self
						if (c >= kk) hi = mid; else lo = mid + 1;
					}
					if (lane == 0) core[p] = (uint32_t)lo;
				}
				__syncthreads();
			}

			RVB_PHASE(2);  // core distances
			}
			if constexpr (MODE == kModeFront) {
				// ---- front half ends here: the points go to scratch, the map to the Prim list of its size class
				if (tid == 0) {
					const unsigned long long need = (unsigned long long)((n + 15) & ~15);
					const unsigned long long off = atomicAdd(a.scr_top, need);
					S.scr_slot = (off + need <= (unsigned long long)a.scr_cap) ? (int)off : -1;
				}
				__syncthreads();
				const int off = S.scr_slot;
				if (off < 0) {
					// scratch exhausted: a monolithic launch takes the map
					if (tid == 0) {
						const int k = atomicAdd(a.ovf_len, 1);
						a.ovf_list[k] = m;
						if (a.skip != nullptr) for (int q = m; a.chain_next[q]; ++q) a.skip[q + 1] = 1;
					}
				} else {
					for (int j = tid; j < n; j += NT) {
						a.scr_pinfo[(size_t)off + j] = make_uint2(core[j], (uint32_t)pts[j]);
						a.scr_val[(size_t)off + j] = val[j];
					}
					if (tid == 0) {
						a.scr_off[m] = off;
						a.out[m] = res;
						int k = 0;
						while (n > split_class_cap(k)) ++k;
						if (a.force_class >= 0) k = a.force_class;
						const int slot = atomicAdd(&a.cls_cnt[k], 1);
						a.cls_lists[(size_t)k * a.cls_stride + slot] = m;
					}
				}
				m = -1;
				continue;
			}
			// ---- phase 3: Prim on the mutual-reachability graph -------------------------------------
			// (prim_segment above: live points in registers, compacted whenever a slot per thread frees up)
			static_assert(kKeyShift == 13, "prim_segment multiplies by 8192");
			if constexpr (MODE == kModeBack) {
				// done by prim_kernel: nodes in the order they were added, with the weight of the edge that added them
				const uint32_t *pk = a.scr_pkey + (size_t)a.scr_off[m];
				if (tid == 0) order[0] = 0;
				for (int e = tid; e < n - 1; e += NT) {
					const uint32_t g = pk[e];
					order[e + 1] = (uint16_t)(g & kKeyIdxMask);
					wp[e] = g >> kKeyShift;
				}
			} else {
				// storage the sort and the tree use later: pinfo = d4 + rank + pe (8 B/point), live-key lists A = a4
				// (the core distances move into pinfo first), B = pl + queue
				uint2 *pinfo = reinterpret_cast<uint2 *>(skey);
				uint32_t *lkA = core;
				uint32_t *lkB = reinterpret_cast<uint32_t *>(pl);
				for (int j = tid; j < n; j += NT) pinfo[j] = make_uint2(core[j], (uint32_t)pts[j]);
				__syncthreads();
				// every point but the root, "not reached yet": the largest weight field, index in the low bits
				for (int j = tid + 1; j < n; j += NT) lkA[j - 1] = 0xFFFFE000u | (uint32_t)j;
				if (tid == 0) order[0] = 0;
				__syncthreads();
				PrimState ps;
				ps.step = 0; ps.live = n - 1; ps.cur = 0;
				int flip = 0;
				while (ps.step < n - 1) {
					const int kslots = (ps.live + NT - 1) / NT;
					// run until the live points fit into one slot less per thread (or to the end)
					int nsteps = (kslots > 1) ? (ps.live - (kslots - 1) * NT) : ps.live;
					nsteps = min(nsteps, n - 1 - ps.step);
					const bool more = (ps.step + nsteps) < (n - 1);
					prim_dispatch<NT, TPT>(kslots, pinfo, order, wp, S.wmin, flip ? lkB : lkA, flip ? lkA : lkB,
										   &S.sort_cnt[0], ps, nsteps, more);
					flip ^= 1;
				}
			}
			__syncthreads();

			RVB_PHASE(3);  // Prim
			// ---- phase 4a: numpy's unstable argsort of the edge weights ------------------------------
			const int ne = n - 1;
			for (int e = tid; e < ne; e += NT) skey[e] = (wp[e] << kKeyShift) | (uint32_t)e;
			__syncthreads();
			np_aquicksort_block<NT>(skey, ne, reinterpret_cast<uint16_t *>(smem + L.a4),
									reinterpret_cast<uint16_t *>(smem + L.a4) + L.nmax, queue, L.nmax / 6, S.sort_cnt);
			__syncthreads();
			for (int r = tid; r < ne; r += NT) rank[skey[r] & kKeyIdxMask] = (uint16_t)r;
			if (tid == 0) S.root_edge = skey[ne - 1] & kKeyIdxMask;
			__syncthreads();

			RVB_PHASE(4);  // argsort emulation
			// ---- phase 4b: the dendrogram as a Cartesian tree over Prim positions --------------------
			// Edge e joins positions e and e+1 at time rank[e]; at that time its cluster is the maximal
			// interval around e whose edges all have smaller rank (make_single_linkage, _linkage.pyx:226).
			{
				uint16_t *bmax = pl;  // per-32 block maxima of rank[]; pl[] itself is written after this loop
				for (int bb = tid; bb * 32 < ne; bb += NT) {
					uint32_t mx = 0;
					for (int i = bb * 32; i < min(ne, bb * 32 + 32); ++i) mx = max(mx, (uint32_t)rank[i]);
					bmax[bb] = (uint16_t)mx;
				}
				__syncthreads();
				for (int e = tid; e < ne; e += NT) {
					const uint32_t re = rank[e];
					int l = e - 1;
					while (l >= 0 && rank[l] < re) {
						if ((l & 31) == 31 && bmax[l >> 5] < re) l -= 32; else --l;
					}
					int r = e + 1;
					while (r < ne && rank[r] < re) {
						if ((r & 31) == 0 && bmax[r >> 5] < re) r += 32; else ++r;
					}
					if (r > ne) r = ne;
					Lp1[e] = (uint16_t)(l + 1);
					Rr[e] = (uint16_t)r;  // points L+1 .. R  (R == ne means "to the last point")
					uint16_t par;
					if (l < 0 && r >= ne) par = kNone16;
					else if (l < 0) par = (uint16_t)r;
					else if (r >= ne) par = (uint16_t)l;
					else par = (rank[l] < rank[r]) ? (uint16_t)l : (uint16_t)r;
					pe[e] = par;
				}
			}
			__syncthreads();
			// skey is dead from here: its storage becomes cL / cR
			for (int e = tid; e < ne; e += NT) { cL[e] = kNone16; cR[e] = kNone16; }
			__syncthreads();
			for (int e = tid; e < ne; e += NT) {
				const uint16_t p = pe[e];
				if (p != kNone16) {
					if (e < (int)p) cL[p] = (uint16_t)e; else cR[p] = (uint16_t)e;
				}
			}
			for (int q = tid; q < n; q += NT) {
				uint16_t par;
				if (q == 0) par = 0;
				else if (q == n - 1) par = (uint16_t)(ne - 1);
				else par = (rank[q - 1] < rank[q]) ? (uint16_t)(q - 1) : (uint16_t)q;
				pl[q] = par;
			}
			__syncthreads();

			RVB_PHASE(5);  // Cartesian tree
			// ---- phase 4c: condensed tree over the big nodes, breadth first (_condense_tree) ---------
			// rank[] is dead from here: its storage becomes relabel[]
			//
			// The library walks the dendrogram node by node; almost all of that walk runs down "spines" (a node with
			// one child of at least min_cluster_size points passes its label on), which only a true split or a dead end
			// terminates.  Here the spines are contracted in parallel (pointer jumping towards the spine's head), one
			// thread replays the breadth-first order over whole spines -- labels are numbered in the order the library
			// dequeues the splits: by depth, then by queue position -- and the labels are spread back in parallel.
			const int mcs = a.mcs;
			{
				uint16_t *up = queue;       // parent along the spine, after the jumping the spine's head, finally its label
				uint16_t *dist = relabel;   // hops from the head; for a head: index of its spine record
				uint16_t *sp_tail = cl_label, *sp_end = cl_selanc;   // per spine: last node, depth of that node
				uint16_t *sp_lnk = reinterpret_cast<uint16_t *>(cl_acc), *sp_lab = sp_lnk + L.ncmax;
				uint16_t *sp_head = sp_lnk;   // (until the replay starts)
				auto left_size = [&](int e) { return e + 1 - (int)Lp1[e]; };
				auto right_size = [&](int e) { return (int)Rr[e] - e; };
				if (tid == 0) { S.err = 0; S.sort_cnt[0] = 0; }
				for (int e = tid; e < ne; e += NT) {
					const uint16_t p = pe[e];
					bool follow = false;
					if (p != kNone16 && left_size(e) + right_size(e) >= mcs) follow = !(left_size(p) >= mcs && right_size(p) >= mcs);
					up[e] = follow ? p : (uint16_t)e;
					dist[e] = follow ? 1 : 0;
				}
				__syncthreads();
				for (int span = 1; span < ne; span <<= 1) {
					uint16_t nu[TPT], nd[TPT];
#pragma unroll
					for (int i = 0; i < TPT; ++i) {
						const int e = tid + i * NT;
						if (e < ne) {
							const int u = up[e];
							nu[i] = up[u];
							nd[i] = (uint16_t)(dist[e] + dist[u]);
						}
					}
					__syncthreads();
#pragma unroll
					for (int i = 0; i < TPT; ++i) {
						const int e = tid + i * NT;
						if (e < ne) { up[e] = nu[i]; dist[e] = nd[i]; }
					}
					__syncthreads();
				}
				// one record per spine, written by its last node
				for (int e = tid; e < ne; e += NT) {
					const int lc = left_size(e), rc = right_size(e);
					if (lc + rc >= mcs && ((lc >= mcs) == (rc >= mcs))) {
						const int r = atomicAdd(&S.sort_cnt[0], 1);
						if (r < L.ncmax) { sp_head[r] = up[e]; sp_tail[r] = (uint16_t)e; sp_end[r] = dist[e]; }
					}
				}
				__syncthreads();
				const int nrec = S.sort_cnt[0];   // = number of condensed-tree clusters
				if (nrec > L.ncmax) {
					if (tid == 0) { res.flags |= kFlagClusterCapacity; a.out[m] = res; }
					m = -1;
					continue;
				}
				for (int r = tid; r < nrec; r += NT) dist[sp_head[r]] = (uint16_t)r;
				__syncthreads();
				if (tid == 0) {
					int ncl = 1;
					cl_stab[0] = 0ull; cl_birth[0] = 0ull; cl_parent[0] = kNone16; cl_ch0[0] = kNone16; cl_ch1[0] = kNone16;
					uint32_t rootminw = 0xFFFFFFFFu;
					int first = dist[S.root_edge];
					sp_lab[first] = 0;
					sp_lnk[first] = kNone16;
					while (first != kNone16) {
						// the nodes the library dequeues next are the spine ends of the smallest depth, in list order
						int dmin = 0x7fffffff;
						for (int r = first; r != kNone16; r = sp_lnk[r]) dmin = min(dmin, (int)sp_end[r]);
						int prev = kNone16;
						for (int r = first; r != kNone16;) {
							const int nxt = sp_lnk[r];
							if ((int)sp_end[r] != dmin) { prev = r; r = nxt; continue; }
							const int e = sp_tail[r];
							const int c = sp_lab[r];
							const int lc = left_size(e), rc = right_size(e);
							if (lc >= mcs && rc >= mcs) {
								const uint32_t w = wp[e];
								const unsigned long long lam = lambda_fix(w);
								const int ca = ncl++, cb = ncl++;
								cl_stab[ca] = 0ull; cl_stab[cb] = 0ull;
								cl_birth[ca] = lam; cl_birth[cb] = lam;
								cl_parent[ca] = (uint16_t)c; cl_parent[cb] = (uint16_t)c;
								cl_ch0[ca] = kNone16; cl_ch1[ca] = kNone16; cl_ch0[cb] = kNone16; cl_ch1[cb] = kNone16;
								cl_ch0[c] = (uint16_t)ca; cl_ch1[c] = (uint16_t)cb;
								cl_stab[c] += (lam - cl_birth[c]) * (unsigned long long)(lc + rc);
								if (c == 0) rootminw = min(rootminw, w);
								// the two children start spines one level down; they take this spine's place in the order
								const int rl = dist[cL[e]], rr = dist[cR[e]];
								sp_lab[rl] = (uint16_t)ca; sp_lab[rr] = (uint16_t)cb;
								sp_end[rl] = (uint16_t)(sp_end[rl] + dmin + 1); sp_end[rr] = (uint16_t)(sp_end[rr] + dmin + 1);
								sp_lnk[rl] = (uint16_t)rr; sp_lnk[rr] = (uint16_t)nxt;
								if (prev == kNone16) first = rl; else sp_lnk[prev] = (uint16_t)rl;
								prev = rr;
							} else {
								if (prev == kNone16) first = nxt; else sp_lnk[prev] = (uint16_t)nxt;
							}
							r = nxt;
						}
					}
					S.ncl = ncl;
					S.rootminw = rootminw;
				}
				__syncthreads();
				// every big node takes the label of its spine (up[e] is read by the thread that owns e only)
				for (int e = tid; e < ne; e += NT)
					if (left_size(e) + right_size(e) >= mcs) up[e] = sp_lab[dist[up[e]]];
				__syncthreads();
				for (int e = tid; e < ne; e += NT)
					if (left_size(e) + right_size(e) >= mcs) relabel[e] = up[e];
				__syncthreads();
			}

			RVB_PHASE(6);  // condensed-tree BFS
			// ---- phase 4d: where each point falls out (the queue storage becomes pcl[]) --------------
			for (int q = tid; q < n; q += NT) {
				int node = pl[q];
				while (((int)Rr[node] + 1 - (int)Lp1[node]) < mcs) node = pe[node];
				const int c = relabel[node];
				const uint32_t w = wp[node];
				pcl[q] = (uint16_t)c;
				pnode[q] = (uint16_t)node;
				atomicAdd(&cl_stab[c], lambda_fix(w) - cl_birth[c]);
				if (c == 0) atomicMin(&S.rootminw, w);
			}
			__syncthreads();

			RVB_PHASE(7);  // fall-out
			// ---- phase 4e: excess of mass, allow_single_cluster=True (_get_clusters) -----------------
			if (tid == 0) {
				const int ncl = S.ncl;
				// bottom-up (children have larger ids): does the node beat its subtree?
				for (int c = ncl - 1; c >= 0; --c) {
					bool win = true;
					if (cl_ch0[c] != kNone16) {
						const unsigned long long sub = cl_stab[cl_ch0[c]] + cl_stab[cl_ch1[c]];
						if (sub > cl_stab[c]) { win = false; cl_stab[c] = sub; }
					}
					cl_label[c] = win ? 1 : 0;
				}
				// top-down: a winner with a winning ancestor is dropped; number the survivors by id
				int nsel = 0;
				for (int c = 0; c < ncl; ++c) {
					const uint16_t par = cl_parent[c];
					const uint16_t anc = (par == kNone16) ? kNone16 : cl_selanc[par];
					if (anc != kNone16) { cl_selanc[c] = anc; cl_label[c] = kNone16; }
					else if (cl_label[c]) { cl_selanc[c] = (uint16_t)c; cl_label[c] = (uint16_t)nsel++; }
					else { cl_selanc[c] = kNone16; cl_label[c] = kNone16; }
				}
				S.nsel = nsel;
				for (int i = 0; i < nsel; ++i) cl_acc[i] = 0u;
			}
			__syncthreads();

			// ---- phase 4f: labels (_do_labelling) and the dominant cluster (smartVidCrop.py:1107-1121) -
			const uint32_t rootminw = S.rootminw;
			for (int q = tid; q < n; q += NT) {
				const uint16_t s = cl_selanc[pcl[q]];
				uint16_t lab = kNone16;
				if (s != kNone16) {
					if (s != 0) lab = cl_label[s];
					else if (wp[pnode[q]] <= rootminw) lab = cl_label[0];
				}
				pcl[q] = lab;
				if (lab != kNone16) {
					const uint32_t v = val[order[q]];
					if (a.select_sum == 1) atomicAdd(&cl_acc[lab], v); else atomicMax(&cl_acc[lab], v);
				}
			}
			__syncthreads();
			if (tid == 0) {
				// every selected cluster owns at least one labelled point, so n_clusters == nsel
				int best = 0;
				for (int i = 1; i < S.nsel; ++i) if (cl_acc[i] > cl_acc[best]) best = i;
				S.maxcl = best;
				S.n_clusters = S.nsel;
			}
			__syncthreads();
			res.n_clusters = S.n_clusters;
			if (a.labels_dbg != nullptr)
				for (int q = tid; q < n; q += NT) a.labels_dbg[order[q]] = (pcl[q] == kNone16) ? -1 : (int)pcl[q];
			if (S.n_clusters > 0) {
				const uint16_t keep = (uint16_t)S.maxcl;
				for (int q = tid; q < n; q += NT) if (pcl[q] != keep) val[order[q]] = 0;
			}
			__syncthreads();

			RVB_PHASE(8);  // EOM + labels + dominant cluster
			// ---- phase 5: rebuild the map, close it --------------------------------------------------
			for (int i = tid; i < ln_words; i += NT) img32[i] = 0u;
			if (tid == 0) { S.bb[0] = 0x7fffffff; S.bb[1] = -1; S.bb[2] = 0x7fffffff; S.bb[3] = -1; }
			__syncthreads();
			{
				int ymin = 0x7fffffff, ymax = -1, xmin = 0x7fffffff, xmax = -1;
				for (int p = tid; p < n; p += NT) {
					const int y = pts[p] >> 8, x = pts[p] & 0xFF;
					const uint8_t v = val[p];
					img8[y * LWPS + x] = v;
					if (v) { ymin = min(ymin, y); ymax = max(ymax, y); xmin = min(xmin, x); xmax = max(xmax, x); }
				}
				ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
				xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
				if (lane == 0 && ymax >= 0) {
					atomicMin(&S.bb[0], ymin); atomicMax(&S.bb[1], ymax); atomicMin(&S.bb[2], xmin); atomicMax(&S.bb[3], xmax);
				}
			}
			__syncthreads();
			rebuilt = true;
			if (a.op_close && S.n_clusters > 0 && S.bb[1] >= 0) {
				const int ry0 = max(0, S.bb[0] - 2), ry1 = min(LH - 1, S.bb[1] + 2);
				const int xw0 = max(0, (S.bb[2] - 2) >> 2), xw1 = min(LMWS - 1, (S.bb[3] + 2) >> 2);
				morph_pass_h<NT, true>(img32, LH, LMWS, ry0, ry1, xw0, xw1);
				morph_pass_v<NT, true>(img32, LH, LMWS, ry0, ry1, xw0, xw1);
				set_row_padding(img32, ry0, ry1, LW, LMWS, 0xFFu, NT);
				__syncthreads();
				morph_pass_h<NT, false>(img32, LH, LMWS, ry0, ry1, xw0, xw1);
				morph_pass_v<NT, false>(img32, LH, LMWS, ry0, ry1, xw0, xw1);
				set_row_padding(img32, ry0, ry1, LW, LMWS, 0x00u, NT);
				__syncthreads();
			}
		}
		if (rz) {
			// cv2.resize(small, (initW, initH), INTER_LINEAR) -- smartVidCrop.py:1158, also when the gates skipped
			// the clustering.  The full-size buffer was re-used by the clustering state: rebuild it completely.
			__syncthreads();
			for (int i = tid; i < n_words; i += NT) map32[i] = 0u;
			__syncthreads();
			cv_resize_linear_u8(img8, LH, LW, LWPS, map8, H, W, WPS, a.rz + a.rz_ux, a.rz + a.rz_uy, NT);
			__syncthreads();
		}
		(void)rebuilt;

		RVB_PHASE(9);  // rebuild + closing
		// ---- phase 6: results -------------------------------------------------------------------------
		if (tid == 0) { S.sx = 0ull; S.sy = 0ull; S.cnt = 0u; S.tot = 0u; S.argmax_key = 0u; }
		__syncthreads();
		{
			uint32_t cnt = 0, tot = 0, sx = 0, sy = 0, amax = 0;
			for (int i = tid; i < n_words; i += NT) {
				const uint32_t v = map32[i];
				if (v == 0u) continue;
				const int y = i / MWS, x0 = (i - y * MWS) * 4;
				tot += __vsadu4(v, 0u);
#pragma unroll
				for (int b = 0; b < 4; ++b) {
					const uint32_t bv = (v >> (8 * b)) & 0xFFu;
					if (bv) {
						++cnt; sx += x0 + b; sy += y;
						// first maximum in row-major order: larger value, then smaller linear index
						const uint32_t lin = (uint32_t)(y * W + x0 + b);
						const uint32_t k = (bv << 20) | (0xFFFFFu - lin);
						amax = max(amax, k);
					}
				}
			}
			cnt = __reduce_add_sync(0xffffffffu, cnt);
			tot = __reduce_add_sync(0xffffffffu, tot);
			sx = __reduce_add_sync(0xffffffffu, sx);
			sy = __reduce_add_sync(0xffffffffu, sy);
			amax = __reduce_max_sync(0xffffffffu, amax);
			if (lane == 0) {
				atomicAdd(&S.cnt, cnt); atomicAdd(&S.tot, tot);
				atomicAdd(&S.sx, (unsigned long long)sx); atomicAdd(&S.sy, (unsigned long long)sy);
				atomicMax(&S.argmax_key, amax);
			}
		}
		__syncthreads();
		res.kept_points = (int)S.cnt;
		if (S.tot != 0u && a.com_km && a.resize_on) {
			// sc_find_center_of_mass down-samples with INTER_NEAREST by the same factor (smartVidCrop.py:1184)
			// and scales the centroid of the non-zero samples back (:1212-1213)
			const int16_t *nx = a.rz + a.rz_nx, *ny = a.rz + a.rz_ny;
			__syncthreads();
			if (tid == 0) { S.sx = 0ull; S.sy = 0ull; S.cnt = 0u; }
			__syncthreads();
			uint32_t cnt = 0, sx = 0, sy = 0;
			for (int i = tid; i < a.Hs * a.Ws; i += NT) {
				const int y = i / a.Ws, x = i - y * a.Ws;
				if (map8[ny[y] * WPS + nx[x]]) { ++cnt; sx += x; sy += y; }
			}
			cnt = __reduce_add_sync(0xffffffffu, cnt);
			sx = __reduce_add_sync(0xffffffffu, sx);
			sy = __reduce_add_sync(0xffffffffu, sy);
			if (lane == 0) { atomicAdd(&S.cnt, cnt); atomicAdd(&S.sx, (unsigned long long)sx); atomicAdd(&S.sy, (unsigned long long)sy); }
			__syncthreads();
			if (S.cnt == 0u) res.flags |= kFlagEmpty;
			else {
				res.cx = ((double)S.sx / (double)S.cnt) * a.factor;
				res.cy = ((double)S.sy / (double)S.cnt) * a.factor;
			}
		} else if (S.tot == 0u) {
			res.flags |= kFlagEmpty;
		} else if (a.com_km) {
			// KMeans(n_clusters=1) == unweighted centroid of the non-zero pixels
			res.cx = (double)S.sx / (double)S.cnt;
			res.cy = (double)S.sy / (double)S.cnt;
		} else {
			const uint32_t lin = 0xFFFFFu - (S.argmax_key & 0xFFFFFu);
			res.cx = (double)(lin % (uint32_t)W);
			res.cy = (double)(lin / (uint32_t)W);
		}

		// coverage score (a7): max over window positions d in range(L - win) of sum(profile[d:d+win]) / total
		if (a.cvrg_cfg != nullptr) {
			const int clip = a.map_clip[m];
			uint32_t *prof = reinterpret_cast<uint32_t *>(smem + L.pts);  // pts/val are dead now
			for (int r = 0; r < a.n_ratios; ++r) {
				const int mode = a.cvrg_cfg[(clip * a.n_ratios + r) * 2 + 0];
				const int win = a.cvrg_cfg[(clip * a.n_ratios + r) * 2 + 1];
				const int Lp = (mode == 1) ? W : H;
				__syncthreads();
				for (int i = tid; i < Lp; i += NT) {
					uint32_t s = 0;
					if (mode == 1) { for (int y = 0; y < H; ++y) s += map8[y * WPS + i]; }
					else { for (int xw = 0; xw < MWS; ++xw) s += __vsadu4(map32[i * MWS + xw], 0u); }
					prof[i] = s;
				}
				__syncthreads();
				if (tid == 0) S.argmax_key = 0u;
				__syncthreads();
				uint32_t best = 0;
				for (int d = tid; d < Lp - win; d += NT) {
					uint32_t s = 0;
					for (int i = d; i < d + win; ++i) s += prof[i];
					best = max(best, s);
				}
				best = __reduce_max_sync(0xffffffffu, best);
				if (lane == 0) atomicMax(&S.argmax_key, best);
				__syncthreads();
				// t_sum == 0 gives NaN in the reference, which never beats 0.0
				res.cvrg[r] = (S.tot == 0u) ? 0.0 : ((double)S.argmax_key / (double)S.tot);
			}
		}

		if (a.store != nullptr && a.store[m] >= 0) {
			uint8_t *dst = a.filt + (size_t)a.store[m] * H * a.fstride;
			const int fw = a.fstride >> 2;
			for (int i = tid; i < H * fw; i += NT) {
				const int y = i / fw, xw = i - y * fw;
				const uint32_t v = (xw < MWS) ? map32[y * MWS + xw] : 0u;
				*reinterpret_cast<uint32_t *>(dst + (size_t)y * a.fstride + xw * 4) = v;
			}
		}
		if (tid == 0) a.out[m] = res;
		RVB_PHASE(10);  // results
		// the filtered map just stored is the blend source of map m+1 (smartVidCrop.py:2369-2373)
		// (only the monolithic kernel walks a chain; the split pipeline runs chains one depth per launch)
		m = (MODE == kModeMono && a.chain_next != nullptr && a.chain_next[m]) ? (m + 1) : -1;
	}
}

}  // namespace rvb
