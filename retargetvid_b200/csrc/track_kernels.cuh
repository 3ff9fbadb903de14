// Centre track kernels: empty-centre fill, per-shot interpolation, zero-phase Butterworth,
// LOESS / Savitzky-Golay (one output frame per warp), crop boxes, border detection.
// All arithmetic is fp64; the reference runs these stages in numpy/scipy float64.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/retargetvid_b200.h"
#include "map_kernel.cuh"

namespace rvb {

struct ClipDev {
	int n_maps, n_frames, n_shots, h_orig, w_orig;
	int map_offset, frame_offset, shot_offset;
	double fr;
};

struct ShotDev {           // one row of vid_data['segmentation'] / ['segmentation_sel']
	int f0, f1, m0, m1;    // inclusive, clip-relative
	int clip;
	int frame_base;        // absolute index of frame f0 in the packed per-frame arrays
	int map_base;          // absolute index of map m0
	int scratch_base;      // offset (doubles) of this shot's scratch area
};

struct FilterCoef {        // Butterworth low-pass in transfer-function form + lfilter_zi
	int order;             // 0: invalid cutoff -> moving-average fallback (smartVidCrop.py:1611-1625)
	double b[RVB_MAX_LP_ORDER + 1];
	double a[RVB_MAX_LP_ORDER + 1];
	double zi[RVB_MAX_LP_ORDER];
	int chunked_ok;        // host probe (filtfilt_chunked_probe): the parallel-in-time evaluation stays within 2e-9 of the
	                       // sequential one for this filter; otherwise the kernel runs the sequential recurrence
};

// ---------------------------------------------------------------------------------------------
// a9: sc_handle_empty_centers, smartVidCrop.py:1221-1300.  One warp per clip: the lanes copy the per-map centres out of
// the map records in parallel; only a clip that has an empty map (no salient pixel left) runs the reference's
// sequential fill, on one lane.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fill_centres_kernel(const ClipDev *clips, int n_clips, const ShotDev *shots, const MapOut *mo,
															double *dx, double *dy, uint8_t *empty, int *clip_status) {
	const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (c >= n_clips) return;
	const ClipDev cl = clips[c];
	const int N = cl.n_maps, base = cl.map_offset;
	bool cap = false, any_empty = false;
	for (int i = lane; i < N; i += 32) {
		const MapOut &o = mo[base + i];
		const int fl = o.flags;
		const bool e = (fl & kFlagEmpty) != 0;
		cap |= (fl & (kFlagOverflow | kFlagClusterCapacity)) != 0;
		any_empty |= e;
		if (empty) empty[base + i] = e ? 1 : 0;
		dx[base + i] = e ? nan("") : o.cx;
		dy[base + i] = e ? nan("") : o.cy;
	}
	cap = __any_sync(0xffffffffu, cap);
	any_empty = __any_sync(0xffffffffu, any_empty);
	int status = cap ? RVB_ERR_CAPACITY : 0;
	__syncwarp();
	if (any_empty && lane == 0) {
		int i = 0;
		while (i < N) {
			if (!isnan(dx[base + i])) { ++i; continue; }
			int j = i;
			while (j + 1 < N && isnan(dx[base + j + 1])) ++j;  // empty run [i, j]
			int d_start = 0x7fffffff, d_end = 0x7fffffff;
			for (int s = 0; s < cl.n_shots; ++s) {
				const ShotDev &sh = shots[cl.shot_offset + s];
				d_start = min(d_start, abs(sh.m0 - i));
				d_end = min(d_end, abs(sh.m1 - j));
			}
			double xf, yf;
			if (d_start < d_end) {  // closer to a shot start: take the next value
				xf = dx[base + j + 1];
				yf = dy[base + j + 1];
			} else {                // else the previous one (python's dx[-1] when the run starts at 0)
				const int k = (i == 0) ? (N - 1) : (i - 1);
				xf = dx[base + k];
				yf = dy[base + k];
			}
			for (int k = i; k <= j; ++k) { dx[base + k] = xf; dy[base + k] = yf; }
			i = j + 1;
		}
		for (int k = 0; k < N; ++k)
			if (isnan(dx[base + k]) && status == 0) status = RVB_ERR_NO_CENTRES;
	}
	if (lane == 0) clip_status[c] = status;
}

// ---------------------------------------------------------------------------------------------
// focus stability: get_points_on_line smartVidCrop.py:1337-1393, sc_check_for_extra_cuts
// :1395-1455, driver :2424-2473.  jumps[i] = mean saliency of filtered map i under the segment from centre
// i-1 to centre i (255 when the move is shorter than min_d_jump).  `np.int` (:1377,1384) does not exist in
// numpy >= 1.24: there every diagonal move raises inside the reference's try/except and yields 255
// (np_int == 0, the behaviour of this image's numpy); np_int == 1 restates the pinned numpy.
// One thread per map, then one thread per clip for the freeze pass.
// ---------------------------------------------------------------------------------------------
__global__ void focus_jumps_kernel(const ClipDev *clips, const int *map_clip, int n_maps_total, const double *dx,
								   const double *dy, const uint8_t *filt, const int *store, int H, int W, int fstride,
								   double min_d, int np_int, double *jumps) {
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= n_maps_total) return;
	const ClipDev cl = clips[map_clip[m]];
	double out = 255.0;
	if (m > cl.map_offset) {
		const double p1x = dx[m - 1], p1y = dy[m - 1], p2x = dx[m], p2y = dy[m];
		const double dX = p2x - p1x, dY = p2y - p1y, dXa = fabs(dX), dYa = fabs(dY);
		if (!(dXa < min_d && dYa < min_d)) {
			const int np = (int)ceil(fmax(dYa, dXa));
			const bool negY = p1y > p2y, negX = p1x > p2x;
			const uint8_t *img = filt + (size_t)store[m] * H * fstride;
			int mode = 0;  // 1 vertical, 2 horizontal, 3 steep diagonal, 4 shallow diagonal
			if (p1x == p2x) mode = 1;
			else if (p1y == p2y) mode = 2;
			else if (np_int) mode = (dYa > dXa) ? 3 : 4;
			// np.arange(a, a + d) has ceil(d) elements; a length mismatch with the buffer raises -> 255
			const int nlen = (mode == 1 || mode == 3) ? (int)ceil(dYa) : (int)ceil(dXa);
			if (mode != 0 && nlen == np) {
				double sum = 0.0;
				int cnt = 0;
				float slope = 0.f;
				if (mode == 3) slope = (float)dX / (float)dY;
				if (mode == 4) slope = (float)dY / (float)dX;
				for (int k = 1; k <= np; ++k) {
					float fx, fy;
					if (mode == 1 || mode == 3) {
						fy = (float)(negY ? (p1y - (double)k) : (p1y + (double)k));
						if (mode == 1) fx = (float)p1x;
						else fx = (float)((double)(long long)(slope * (fy - (float)p1y)) + p1x);
					} else {
						fx = (float)(negX ? (p1x - (double)k) : (p1x + (double)k));
						if (mode == 2) fy = (float)p1y;
						else fy = (float)((double)(long long)(slope * (fx - (float)p1x)) + p1y);
					}
					if (fx >= 0.f && fy >= 0.f && fx < (float)W && fy < (float)H) {
						++cnt;
						sum += (double)img[(int)floorf(fy) * fstride + (int)floorf(fx)];
					}
				}
				out = (cnt > 0) ? (sum / (double)cnt) : 255.0;
			}
		}
	}
	jumps[m] = out;
}

__global__ void focus_apply_kernel(const ClipDev *clips, int n_clips, const double *jumps, double thr, double max_secs,
								   int skip, double *dx, double *dy) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_clips) return;
	const ClipDev cl = clips[c];
	const int N = cl.n_maps, base = cl.map_offset;
	int prev = -1;  // previous index with jumps < thr
	for (int i = 1; i < N; ++i) {
		if (!(jumps[base + i] < thr)) continue;
		if (prev >= 0) {
			const int start = max(prev - 1, 0), end = min(i + 1, N - 1);
			const double dur = ((double)((end - start) * skip)) / cl.fr;
			if (!(dur > max_secs)) {
				for (int j = 0; j < end - start; ++j) {
					dx[base + start + j] = dx[base + start];
					dy[base + start + j] = dy[base + start];
				}
			}
		}
		prev = i;
	}
}

// ---------------------------------------------------------------------------------------------
// a10: interp_handler / sc_interpolate, smartVidCrop.py:1528-1597.
// scipy.interpolate.interp1d(kind='linear'|'quadratic', fill_value='extrapolate').  The quadratic
// kind is make_interp_spline(k=2): knots x0 x0 x0, midpoints m_1..m_{n-3}, x_{n-1} x3, a banded
// collocation system, de Boor evaluation (extrapolating with the end pieces).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bspl2_basis(const double *t, int mu, double x, double h[3]) {
	// de Boor-Cox recurrence for the three quadratic B-splines that are non-zero on [t_mu, t_mu+1)
	double hh[3];
	h[0] = 1.0;
	for (int j = 1; j <= 2; ++j) {
		for (int i = 0; i < j; ++i) hh[i] = h[i];
		h[0] = 0.0;
		for (int nn = 1; nn <= j; ++nn) {
			const double xb = t[mu + nn], xa = t[mu + nn - j];
			if (xb == xa) { h[nn] = 0.0; continue; }
			const double w = hh[nn - 1] / (xb - xa);
			h[nn - 1] += w * (xb - x);
			h[nn] = w * (x - xa);
		}
	}
}

__device__ __forceinline__ int bspl_interval(const double *t, int nc, double x) {
	// l in [2, nc-1] with t[l] <= x < t[l+1]; clamped for extrapolation
	if (!(x > t[2])) return 2;
	if (x >= t[nc]) return nc - 1;
	int lo = 2, hi = nc - 1;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (t[mid] <= x) lo = mid; else hi = mid - 1;
	}
	return lo;
}

// one warp per shot: build knots and collocation rows (lanes in parallel), solve for x and y together
// (lane 0, sequential banded elimination).  The working set lives in shared memory when it fits
// (smem_doubles), otherwise in the shot's global scratch area; knots and coefficients always end up there.
__global__ void __launch_bounds__(32) spline_setup_kernel(const ShotDev *shots, int n_shots, const int *true_inds,
														   const double *dx, const double *dy, double *scratch, int smem_doubles) {
	extern __shared__ double sp_smem[];
	const int s = blockIdx.x, lane = threadIdx.x;
	if (s >= n_shots) return;
	const ShotDev sh = shots[s];
	const int n = sh.m1 - sh.m0 + 1;
	if (n <= 6) return;
	// layout: t[n+3], cx[n], cy[n], band[n][5]
	double *g = scratch + sh.scratch_base;
	const bool on_chip = (8 * n + 3) <= smem_doubles;
	double *t = on_chip ? sp_smem : g;
	double *cx = t + (n + 3);
	double *cy = cx + n;
	double *band = cy + n;
	const int *ti = true_inds + sh.map_base;
	auto X = [&](int i) { return (double)(ti[i] - ti[0]); };
	for (int i = lane; i < n + 3; i += 32) {
		double v;
		if (i < 3) v = 0.0;
		else if (i >= n) v = X(n - 1);
		else v = (X(i - 2) + X(i - 1)) / 2.0;   // t[2 + k] = (x_k + x_{k+1}) / 2 for k = 1 .. n-3
		t[i] = v;
	}
	__syncwarp();
	for (int i = lane; i < n; i += 32) {
		for (int k = 0; k < 5; ++k) band[i * 5 + k] = 0.0;
		const double x = X(i);
		const int mu = bspl_interval(t, n, x);
		double h[3];
		bspl2_basis(t, mu, x, h);
		// columns mu-2..mu, stored at band[i][col - i + 2]
		for (int k = 0; k < 3; ++k) {
			const int col = mu - 2 + k;
			const int o = col - i + 2;
			if (o >= 0 && o < 5) band[i * 5 + o] = h[k];
		}
		cx[i] = dx[sh.map_base + i];
		cy[i] = dy[sh.map_base + i];
	}
	__syncwarp();
	if (lane == 0) {
		// Gaussian elimination without pivoting (the collocation matrix is totally positive; two sub- and two super-
		// diagonals).  The pivot row and the two rows below it live in registers and the next row is loaded while the
		// current step computes, so the serial chain per row is one division and two dependent FMAs instead of a string
		// of shared-memory round trips.  One division per row: the reciprocal of the pivot is kept in the (dead) sub-
		// diagonal slot band[i][0] for the back substitution.
		struct Row { double l2, l1, d, u1, u2, x, y; };
		auto load_row = [&](int r) -> Row {
			Row q;
			if (r < n) {
				q.l2 = band[r * 5 + 0]; q.l1 = band[r * 5 + 1]; q.d = band[r * 5 + 2]; q.u1 = band[r * 5 + 3]; q.u2 = band[r * 5 + 4];
				q.x = cx[r]; q.y = cy[r];
			} else {
				q.l2 = 0.0; q.l1 = 0.0; q.d = 1.0; q.u1 = 0.0; q.u2 = 0.0; q.x = 0.0; q.y = 0.0;
			}
			return q;
		};
		Row r0 = load_row(0), r1 = load_row(1), r2 = load_row(2);
		for (int i = 0; i < n; ++i) {
			const Row r3 = load_row(i + 3);
			const double inv = 1.0 / r0.d;
			const double f1 = r1.l1 * inv, f2 = r2.l2 * inv;
			r1.d -= f1 * r0.u1; r1.u1 -= f1 * r0.u2; r1.x -= f1 * r0.x; r1.y -= f1 * r0.y;
			r2.l1 -= f2 * r0.u1; r2.d -= f2 * r0.u2; r2.x -= f2 * r0.x; r2.y -= f2 * r0.y;
			band[i * 5 + 0] = inv;
			band[i * 5 + 3] = r0.u1;
			cx[i] = r0.x;
			cy[i] = r0.y;
			r0 = r1; r1 = r2; r2 = r3;
		}
		double x1 = 0.0, x2 = 0.0, y1 = 0.0, y2 = 0.0;
		double u1 = band[(n - 1) * 5 + 3], u2 = band[(n - 1) * 5 + 4], inv = band[(n - 1) * 5 + 0], bx = cx[n - 1], by = cy[n - 1];
		for (int i = n - 1; i >= 0; --i) {
			const double cu1 = (i + 1 < n) ? u1 : 0.0, cu2 = (i + 2 < n) ? u2 : 0.0, cinv = inv, cbx = bx, cby = by;
			if (i > 0) { u1 = band[(i - 1) * 5 + 3]; u2 = band[(i - 1) * 5 + 4]; inv = band[(i - 1) * 5 + 0]; bx = cx[i - 1]; by = cy[i - 1]; }
			double sx = cbx, sy = cby;
			sx -= cu1 * x1; sy -= cu1 * y1;
			sx -= cu2 * x2; sy -= cu2 * y2;
			sx *= cinv; sy *= cinv;
			cx[i] = sx; cy[i] = sy;
			x2 = x1; y2 = y1; x1 = sx; y1 = sy;
		}
	}
	__syncwarp();
	if (on_chip)
		for (int i = lane; i < 3 * n + 3; i += 32) g[i] = t[i];   // knots + both coefficient vectors
}

// one thread per (shot-local frame): evaluate the interpolant
__global__ void interp_eval_kernel(const ShotDev *shots, const int *frame_shot, int n_frames_total,
								   const int *true_inds, const double *dx, const double *dy,
								   const double *scratch, double *dxi, double *dyi) {
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n_frames_total) return;
	const ShotDev sh = shots[frame_shot[f]];
	const int n = sh.m1 - sh.m0 + 1;
	const double x = (double)(f - sh.frame_base);
	const int *ti = true_inds + sh.map_base;
	const double *px = dx + sh.map_base, *py = dy + sh.map_base;
	double ox, oy;
	if (n < 3) {
		ox = px[0];
		oy = py[0];
	} else if (n <= 6) {
		// interp1d._call_linear: searchsorted (left), clip to [1, n-1]
		int idx = 0;
		while (idx < n && (double)(ti[idx] - ti[0]) < x) ++idx;
		idx = max(1, min(n - 1, idx));
		const double xlo = (double)(ti[idx - 1] - ti[0]), xhi = (double)(ti[idx] - ti[0]);
		const double sx = (px[idx] - px[idx - 1]) / (xhi - xlo);
		const double sy = (py[idx] - py[idx - 1]) / (xhi - xlo);
		ox = __dadd_rn(__dmul_rn(sx, x - xlo), px[idx - 1]);
		oy = __dadd_rn(__dmul_rn(sy, x - xlo), py[idx - 1]);
	} else {
		const double *t = scratch + sh.scratch_base;
		const double *cx = t + (n + 3);
		const double *cy = cx + n;
		const int mu = bspl_interval(t, n, x);
		double h[3];
		bspl2_basis(t, mu, x, h);
		ox = 0.0;
		oy = 0.0;
		for (int k = 0; k < 3; ++k) {
			ox += cx[mu - 2 + k] * h[k];
			oy += cy[mu - 2 + k] * h[k];
		}
	}
	dxi[f] = ox;
	dyi[f] = oy;
}

// ---------------------------------------------------------------------------------------------
// a11: sc_butter_lowpass_filter, smartVidCrop.py:1599-1627: scipy.signal.filtfilt (odd padding of
// 3*max(len(a),len(b)) samples, lfilter_zi initial state, direct form II transposed), with the
// reference's moving-average fallback when filtfilt raises (shots of <= padlen frames).
// One warp per (shot, axis).
// ---------------------------------------------------------------------------------------------
// forward and backward pass of filtfilt over buf[0 .. ne), direct form II transposed, by one warp.  The recurrence is
// linear in (state, input): z' = A z + B x.  Each pass is cut into 32 chunks of L samples, one per lane:
//   1. every lane runs its chunk from a ZERO state and keeps the end state e_k (zero-state response);
//   2. the true state at the start of chunk k follows from s_k = A^L s_(k-1) + e_(k-1) (lane 0, 31 small matrix-vector
//      products; the columns of A^L come from lanes 0 .. M-1 running the homogeneous recurrence on the unit vectors);
//   3. every lane runs its chunk again from s_k and stores the outputs -- the same operations in the same order as the
//      sequential filter, only the chunk's start state carries different rounding (relative 1e-16).
// 2 640 serial steps of the longest shot become 3 x 83 + 31.  State and coefficients stay in registers (the order is a
// compile-time constant).
struct FiltScratch { double P[8][8], E[32][8], S[32][8]; };

// The sequential evaluation: scipy's lfilter loop operation for operation, on one lane.  Used when the filter is so ill
// conditioned in transfer-function form (high order, low cut-off) that ANY other order of the same arithmetic moves the
// result by more than the parity tolerance -- scipy's own output carries that error, so only its order reproduces it.
template <int M>
__device__ __forceinline__ void filtfilt_inplace(const FilterCoef &fc, double *buf, int ne) {
	double b[M + 1], a[M + 1], zi[M], z[M];
#pragma unroll
	for (int i = 0; i <= M; ++i) { b[i] = fc.b[i]; a[i] = fc.a[i]; }
#pragma unroll
	for (int i = 0; i < M; ++i) zi[i] = fc.zi[i];
	auto step = [&](double x) -> double {
		const double y = __dadd_rn(z[0], __dmul_rn(b[0], x));
#pragma unroll
		for (int i = 0; i < M - 1; ++i) z[i] = __dsub_rn(__dadd_rn(z[i + 1], __dmul_rn(x, b[i + 1])), __dmul_rn(y, a[i + 1]));
		z[M - 1] = __dsub_rn(__dmul_rn(x, b[M]), __dmul_rn(y, a[M]));
		return y;
	};
	// (the next sample is loaded before the current one enters the serial chain)
	const double e0 = buf[0];
#pragma unroll
	for (int i = 0; i < M; ++i) z[i] = zi[i] * e0;
	double xn = e0;
	for (int i = 0; i < ne; ++i) {
		const double xc = xn;
		if (i + 1 < ne) xn = buf[i + 1];
		buf[i] = step(xc);
	}
	const double y0 = buf[ne - 1];
#pragma unroll
	for (int i = 0; i < M; ++i) z[i] = zi[i] * y0;
	xn = y0;
	for (int i = ne - 1; i >= 0; --i) {
		const double xc = xn;
		if (i > 0) xn = buf[i - 1];
		buf[i] = step(xc);
	}
}

template <int M>
__device__ __forceinline__ void filtfilt_warp(const FilterCoef &fc, double *buf, int ne, int lane, struct FiltScratch &fs);

template <int M>
__device__ __forceinline__ void filtfilt_any(const FilterCoef &fc, double *buf, int ne, int lane, struct FiltScratch &fs) {
	if (fc.chunked_ok) filtfilt_warp<M>(fc, buf, ne, lane, fs);
	else if (lane == 0) filtfilt_inplace<M>(fc, buf, ne);
}



template <int M>
__device__ __forceinline__ void filtfilt_warp(const FilterCoef &fc, double *buf, int ne, int lane, FiltScratch &fs) {
	double b[M + 1], a[M + 1];
#pragma unroll
	for (int i = 0; i <= M; ++i) { b[i] = fc.b[i]; a[i] = fc.a[i]; }
	auto step = [&](double (&z)[M], double x) -> double {
		const double y = __dadd_rn(z[0], __dmul_rn(b[0], x));
#pragma unroll
		for (int i = 0; i < M - 1; ++i) z[i] = __dsub_rn(__dadd_rn(z[i + 1], __dmul_rn(x, b[i + 1])), __dmul_rn(y, a[i + 1]));
		z[M - 1] = __dsub_rn(__dmul_rn(x, b[M]), __dmul_rn(y, a[M]));
		return y;
	};
	const int L = (ne + 31) >> 5;
	if (lane < M) {
		double z[M];
#pragma unroll
		for (int i = 0; i < M; ++i) z[i] = (i == lane) ? 1.0 : 0.0;
		for (int n = 0; n < L; ++n) step(z, 0.0);
#pragma unroll
		for (int r = 0; r < M; ++r) fs.P[r][lane] = z[r];
	}
	const int i0 = min(ne, lane * L), i1 = min(ne, i0 + L);
	for (int pass = 0; pass < 2; ++pass) {
		const int dir = pass ? -1 : 1;
		double *p0 = pass ? (buf + ne - 1 - i0) : (buf + i0);       // first sample of this lane's chunk in pass order
		double z[M];
#pragma unroll
		for (int i = 0; i < M; ++i) z[i] = 0.0;
		{
			// (the next sample is loaded before the current one enters the serial chain)
			const double *q = p0;
			double xn = (i0 < i1) ? *q : 0.0;
			for (int i = i0; i < i1; ++i) {
				const double xc = xn;
				q += dir;
				if (i + 1 < i1) xn = *q;
				step(z, xc);
			}
		}
#pragma unroll
		for (int r = 0; r < M; ++r) fs.E[lane][r] = z[r];
		__syncwarp();
		if (lane == 0) {
			double sv[M];
			const double x0 = pass ? buf[ne - 1] : buf[0];          // filtfilt: lfilter_zi state scaled by the first sample
#pragma unroll
			for (int r = 0; r < M; ++r) { sv[r] = fc.zi[r] * x0; fs.S[0][r] = sv[r]; }
			for (int k = 1; k < 32; ++k) {
				double t[M];
#pragma unroll
				for (int r = 0; r < M; ++r) {
					double acc = fs.E[k - 1][r];
#pragma unroll
					for (int c = 0; c < M; ++c) acc += fs.P[r][c] * sv[c];
					t[r] = acc;
				}
#pragma unroll
				for (int r = 0; r < M; ++r) { sv[r] = t[r]; fs.S[k][r] = t[r]; }
			}
		}
		__syncwarp();
#pragma unroll
		for (int r = 0; r < M; ++r) z[r] = fs.S[lane][r];
		{
			double *q = p0;
			double xn = (i0 < i1) ? *q : 0.0;
			for (int i = i0; i < i1; ++i) {
				const double xc = xn;
				double *cur = q;
				q += dir;
				if (i + 1 < i1) xn = *q;
				*cur = step(z, xc);
			}
		}
		__syncwarp();
	}
}

// also leaves min / max of the low-passed series of every (shot, axis) in shot_minmax[(shot * 2 + axis) * 2 + {0, 1}]:
// the LOESS stage normalises by them (pyloess.py:16-24) and would otherwise rescan the shot for every output frame
__global__ void __launch_bounds__(32) lowpass_kernel(const ShotDev *shots, int n_shots, const ClipDev *clips, const FilterCoef *coefs,
													  const int *clip_coef, const double *dxi, const double *dyi, double *dxl, double *dyl,
													  double *scratch, int lp_filt, int smem_doubles, double *shot_minmax,
													  int loess_filt, double loess_w_secs, int degree) {
	extern __shared__ double lp_smem[];
	__shared__ FiltScratch lp_fs;
	const int id = blockIdx.x, lane = threadIdx.x;
	if (id >= n_shots * 2) return;
	const int s = id >> 1, axis = id & 1;
	const ShotDev sh = shots[s];
	const int cl = sh.f1 - sh.f0 + 1;
	const double *x = (axis ? dyi : dxi) + sh.frame_base;
	double *out = (axis ? dyl : dxl) + sh.frame_base;
	double vmin = INFINITY, vmax = -INFINITY;
	const FilterCoef &fc = coefs[clip_coef[sh.clip]];
	const int m = fc.order;
	const int edge = 3 * (m + 1);
	if (!lp_filt) {
		for (int i = lane; i < cl; i += 32) { const double v = x[i]; out[i] = v; vmin = fmin(vmin, v); vmax = fmax(vmax, v); }
	} else if (m > 0 && cl > edge) {
		// odd extension (lanes in parallel), forward and backward pass in place (filtfilt_warp), copy out
		const int ne = cl + 2 * edge;
		double *buf = (ne <= smem_doubles) ? lp_smem : (scratch + sh.scratch_base + (size_t)axis * ne);
		for (int i = lane; i < ne; i += 32) {
			double v;
			if (i < edge) v = 2.0 * x[0] - x[edge - i];
			else if (i < edge + cl) v = x[i - edge];
			else v = 2.0 * x[cl - 1] - x[cl - 2 - (i - edge - cl)];
			buf[i] = v;
		}
		__syncwarp();
		switch (m) {
		case 1: filtfilt_any<1>(fc, buf, ne, lane, lp_fs); break;
		case 2: filtfilt_any<2>(fc, buf, ne, lane, lp_fs); break;
		case 3: filtfilt_any<3>(fc, buf, ne, lane, lp_fs); break;
		case 4: filtfilt_any<4>(fc, buf, ne, lane, lp_fs); break;
		case 5: filtfilt_any<5>(fc, buf, ne, lane, lp_fs); break;
		case 6: filtfilt_any<6>(fc, buf, ne, lane, lp_fs); break;
		case 7: filtfilt_any<7>(fc, buf, ne, lane, lp_fs); break;
		default: filtfilt_any<8>(fc, buf, ne, lane, lp_fs); break;
		}
		__syncwarp();
		for (int i = lane; i < cl; i += 32) { const double v = buf[edge + i]; out[i] = v; vmin = fmin(vmin, v); vmax = fmax(vmax, v); }
	} else {
		// fallback: 5-tap moving average of the interior, edges untouched (smartVidCrop.py:1611-1615)
		for (int i = lane; i < cl; i += 32) {
			double v = x[i];
			if (cl >= 5 && i >= 2 && i < cl - 2) v = ((((x[i - 2] + x[i - 1]) + x[i]) + x[i + 1]) + x[i + 2]) / 5.0;
			out[i] = v;
			vmin = fmin(vmin, v); vmax = fmax(vmax, v);
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
		vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
	}
	if (lane == 0 && shot_minmax != nullptr) { shot_minmax[(size_t)id * 2] = vmin; shot_minmax[(size_t)id * 2 + 1] = vmax; }
	// The LOESS / Savitzky-Golay estimate of a frame whose window is not clamped by the shot's ends is a fixed linear
	// combination of the window: the weights depend on the offset only, so b0 = sum_k coef[k] * (y[j - h + k] - y[j]).
	// The coefficients of the shot go to the start of its scratch area (the spline data there is dead by now).
	if (axis == 0 && cl >= 10) {
		__syncwarp();
		const double fr = clips[sh.clip].fr;
		int win = min((int)(fr * loess_w_secs), cl - 2);
		if ((win & 1) == 0) win -= 1;
		const int h = (win - 1) >> 1;
		double *coef = scratch + sh.scratch_base;
		if (h >= 1) {
			double s0 = 0, s2 = 0, s4 = 0;
			for (int k = lane; k < win; k += 32) {
				const double u = (double)(k - h) / (double)h;
				double w = 1.0;
				if (loess_filt) { const double r = fabs(u); const double c = 1.0 - r * r * r; w = c * c * c; }
				const double wu2 = w * u * u;
				s0 += w; s2 += wu2; s4 += wu2 * u * u;
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				s0 += __shfl_xor_sync(0xffffffffu, s0, o);
				s2 += __shfl_xor_sync(0xffffffffu, s2, o);
				s4 += __shfl_xor_sync(0xffffffffu, s4, o);
			}
			for (int k = lane; k < win; k += 32) {
				const double u = (double)(k - h) / (double)h;
				double w = 1.0;
				if (loess_filt) { const double r = fabs(u); const double c = 1.0 - r * r * r; w = c * c * c; }
				// symmetric window: s1 = s3 = 0, so Cramer's rule leaves (s2 s4 - s2^2 u^2) / (s0 s2 s4 - s2^3); degree 1: 1 / s0
				coef[k] = (degree >= 2) ? w * (s2 * s4 - s2 * s2 * u * u) / (s0 * s2 * s4 - s2 * s2 * s2) : w / s0;
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// a12: loess_handler + pyloess.Loess.estimate (pyloess.py:61-95), one output frame per thread.
// The reference fits, for every frame j, a tricube-weighted polynomial over the `window` nearest
// frames and evaluates it at j.  Here the fit is done in coordinates centred on j (u = i - j), so the
// estimate is the constant coefficient: moment sums over the window, then the 3x3 (or 2x2) normal equations.  Savitzky-Golay (loess_filt == 0,
// scipy.signal.savgol_filter mode='interp') is the same fit with unit weights.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

__global__ void __launch_bounds__(128) smooth_kernel(const ShotDev *shots, const int *frame_shot, int n_frames_total, const ClipDev *clips,
							  const double *dxl, const double *dyl, double *dxs, double *dys, int loess_filt,
							  double loess_w_secs, int degree, const double *shot_minmax, const double *scratch) {
	// one thread per (frame, axis = blockIdx.y): consecutive threads read consecutive samples; the frames of a shot whose
	// window is clamped (the first and last (window - 1) / 2) sit next to each other, so a warp is almost always all
	// interior (fixed coefficients, one FMA per tap) or all edge (moment sums + 3x3 solve)
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	const int axis = blockIdx.y;
	if (f >= n_frames_total) return;
	const int shot = frame_shot[f];
	const ShotDev sh = shots[shot];
	const int cl = sh.f1 - sh.f0 + 1;
	const int j = f - sh.frame_base;
	const double *y = (axis ? dyl : dxl) + sh.frame_base;
	double *out = (axis ? dys : dxs) + sh.frame_base;
	const double yj = y[j];
	if (cl < 10) {  // loess_handler: short shots bypass smoothing
		out[j] = yj;
		return;
	}
	const double fr = clips[sh.clip].fr;
	int win = min((int)(fr * loess_w_secs), cl - 2);
	if ((win & 1) == 0) win -= 1;
	const int h = (win - 1) >> 1;
	int lo = j - h;
	if (lo < 0) lo = 0;
	if (lo + win > cl) lo = cl - win;
	// normalisation of y (pyloess.py:16-24); a constant series divides by zero -> NaN -> identity.  The shot's min / max
	// come from the low-pass kernel.
	const double ymin = shot_minmax[((size_t)shot * 2 + axis) * 2], ymax = shot_minmax[((size_t)shot * 2 + axis) * 2 + 1];
	if (loess_filt && !(ymax > ymin)) {
		out[j] = yj;
		return;
	}
	if (h >= 1 && j - h >= 0 && j + h <= cl - 1) {
		// window not clamped: fixed coefficients (lowpass_kernel); four partial sums keep the FMA chains short
		const double *coef = scratch + sh.scratch_base;
		const double *yw = y + lo;
		double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
		int i = 0;
		for (; i + 3 < win; i += 4) {
			a0 += coef[i] * (yw[i] - yj);
			a1 += coef[i + 1] * (yw[i + 1] - yj);
			a2 += coef[i + 2] * (yw[i + 2] - yj);
			a3 += coef[i + 3] * (yw[i + 3] - yj);
		}
		for (; i < win; ++i) a0 += coef[i] * (yw[i] - yj);
		out[j] = ((a0 + a1) + (a2 + a3)) + yj;
		return;
	}
	const double dmax = (double)max(j - lo, lo + win - 1 - j);
	double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, t0 = 0, t1 = 0, t2 = 0;
	for (int i = lo; i < lo + win; ++i) {
		const double u = (double)(i - j) / dmax;  // |u| <= 1 keeps the normal equations well scaled
		double w = 1.0;
		if (loess_filt) {
			const double r = fabs(u);
			const double c = 1.0 - r * r * r;
			w = c * c * c;
		}
		const double yy = y[i] - yj;
		const double wu = w * u, wu2 = wu * u;
		s0 += w; s1 += wu; s2 += wu2; s3 += wu2 * u; s4 += wu2 * u * u;
		t0 += w * yy; t1 += wu * yy; t2 += wu2 * yy;
	}
	double b0;
	if (degree >= 2) {
		// solve [s0 s1 s2; s1 s2 s3; s2 s3 s4] b = [t0 t1 t2] for b0 (Cramer, symmetric 3x3)
		const double c00 = s2 * s4 - s3 * s3;
		const double c01 = s1 * s4 - s2 * s3;
		const double c02 = s1 * s3 - s2 * s2;
		const double det = s0 * c00 - s1 * c01 + s2 * c02;
		b0 = (t0 * c00 - t1 * c01 + t2 * c02) / det;
	} else {
		const double det = s0 * s2 - s1 * s1;
		b0 = (t0 * s2 - t1 * s1) / det;
	}
	out[j] = b0 + yj;
}

// ---------------------------------------------------------------------------------------------
// a4 (border detection) -- smartVidCrop.py:842-924.  Profiles are the max over time and rows/cols
// of the RAW maps; one CTA per map accumulates into its clip's profile with atomicMax.
// ---------------------------------------------------------------------------------------------
__global__ void border_profile_kernel(const uint8_t *maps, int H, int W, int gstride, const int *map_clip,
									  uint32_t *prof /* [n_clips][H + W] */) {
	const int m = blockIdx.x;
	const uint8_t *src = maps + (size_t)m * H * gstride;
	uint32_t *p = prof + (size_t)map_clip[m] * (H + W);
	for (int y = threadIdx.x; y < H; y += blockDim.x) {
		uint32_t mx = 0;
		for (int x = 0; x < W; ++x) mx = max(mx, (uint32_t)src[y * gstride + x]);
		atomicMax(&p[y], mx);
	}
	for (int x = threadIdx.x; x < W; x += blockDim.x) {
		uint32_t mx = 0;
		for (int y = 0; y < H; ++y) mx = max(mx, (uint32_t)src[y * gstride + x]);
		atomicMax(&p[H + x], mx);
	}
}

__global__ void border_finish_kernel(const ClipDev *clips, int n_clips, const uint32_t *prof, int H, int W,
									 int t_border, int *borders /* [n_clips][4] t,b,l,r in original px */) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_clips) return;
	int t = 0, b = 0, l = 0, r = 0;
	if (t_border != -1) {
		const uint32_t *fcol = prof + (size_t)c * (H + W);
		const uint32_t *frow = fcol + H;
		for (int i = 0; i < H; ++i) { if ((int)fcol[i] > t_border) break; ++t; }
		for (int i = 0; i < H; ++i) { if ((int)fcol[H - 1 - i] > t_border) break; ++b; }
		for (int i = 0; i < W; ++i) { if ((int)frow[i] > t_border) break; ++l; }
		for (int i = 0; i < W; ++i) { if ((int)frow[W - 1 - i] > t_border) break; ++r; }
		t = min(t, (int)(H * 0.45)); b = min(b, (int)(H * 0.45));
		l = min(l, (int)(W * 0.45)); r = min(r, (int)(W * 0.45));
		const double ho = clips[c].h_orig, wo = clips[c].w_orig;
		t = (int)((ho / H) * t); b = (int)((ho / H) * b);
		l = (int)((wo / W) * l); r = (int)((wo / W) * r);
	}
	borders[c * 4 + 0] = t; borders[c * 4 + 1] = b; borders[c * 4 + 2] = l; borders[c * 4 + 3] = r;
}

// ---------------------------------------------------------------------------------------------
// a14 + a15: sc_compute_bb (smartVidCrop.py:979-1048) and sc_shift_time (:1740-1746).
// One thread per (ratio, frame).
// ---------------------------------------------------------------------------------------------
struct BoxGeom { int fbb_w, fbb_h, hbbw1, hbbw2, hbbh1, hbbh2; };

__device__ __forceinline__ BoxGeom box_geom(int w_final, int h_final, int w_orig, int h_orig, const int *bd) {
	BoxGeom g;
	g.fbb_w = w_final;
	g.fbb_h = h_final;
	if (h_final == h_orig) {
		g.fbb_h = h_final - bd[0] - bd[1];
		g.fbb_w = (int)(((double)g.fbb_h / (double)h_final) * w_final);
	}
	if (w_final == w_orig) {
		g.fbb_w = w_final - bd[2] - bd[3];
		g.fbb_h = (int)(((double)g.fbb_w / (double)w_final) * h_final);
	}
	g.hbbw1 = (int)(g.fbb_w / 2.0);
	g.hbbw2 = g.fbb_w - g.hbbw1;
	g.hbbh1 = (int)(g.fbb_h / 2.0);
	g.hbbh2 = g.fbb_h - g.hbbh1;
	return g;
}

__global__ void boxes_kernel(const ClipDev *clips, const int *frame_clip, int n_frames_total, int n_ratios,
							 const int *clip_final /* [n_clips][R][3] mode, w_final, h_final */,
							 const int *borders, int Hp, int Wp, const double *dxs, const double *dys,
							 int shift, int32_t *boxes /* [R][F][4] */, int32_t *clip_dims /* [n_clips][R][9] or null */) {
	const int id = blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= n_frames_total * n_ratios) return;
	const int r = id / n_frames_total, f = id - r * n_frames_total;
	const int c = frame_clip[f];
	const ClipDev cl = clips[c];
	const int *fin = clip_final + (c * n_ratios + r) * 3;
	const int *bd = borders + c * 4;
	const BoxGeom g = box_geom(fin[1], fin[2], cl.w_orig, cl.h_orig, bd);
	// sc_shift_time (smartVidCrop.py:1740-1746), both loops: first `bbs[-i+1] = bbs[-1]` for
	// i in range(shift) (python indices 1, 0, -1, -2, ...), then bbs[i] = bbs[i + shift].
	const int fl = f - cl.frame_offset;
	int src = fl;
	if (shift > 0) {
		if (fl < cl.n_frames - shift) src = fl + shift;
		bool hit = false;
		for (int k = 0; k < shift; ++k) {
			int idx = 1 - k;
			if (idx < 0) idx += cl.n_frames;
			if (idx == src) hit = true;
		}
		if (hit) src = cl.n_frames - 1;
	}
	const double scale_h = (double)Hp / (double)cl.h_orig;
	const double scale_w = (double)Wp / (double)cl.w_orig;
	const int cx = (int)(dxs[cl.frame_offset + src] / scale_w);
	const int cy = (int)(dys[cl.frame_offset + src] / scale_h);
	int x1 = cx - g.hbbw1, y1 = cy - g.hbbh1, x2 = cx + g.hbbw2, y2 = cy + g.hbbh2;
	if (x1 < bd[2]) { x1 = bd[2]; x2 = x1 + g.fbb_w; }
	if (x2 > cl.w_orig - bd[3]) { x2 = cl.w_orig - bd[3]; x1 = x2 - g.fbb_w; }
	if (y1 < bd[0]) { y1 = bd[0]; y2 = y1 + g.fbb_h; }
	if (y2 > cl.h_orig - bd[1]) { y2 = cl.h_orig - bd[1]; y1 = y2 - g.fbb_h; }
	int32_t *o = boxes + ((size_t)r * n_frames_total + f) * 4;
	o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
	if (clip_dims != nullptr && fl == 0) {
		int32_t *d = clip_dims + (c * n_ratios + r) * 9;
		d[0] = fin[0]; d[1] = fin[1]; d[2] = fin[2]; d[3] = g.fbb_w; d[4] = g.fbb_h;
		d[5] = bd[0]; d[6] = bd[1]; d[7] = bd[2]; d[8] = bd[3];
	}
}

// per-clip scores: mean saliency of the raw maps and mean coverage per ratio.  One warp per clip: the raw sums are exact
// integers (any order); the coverage means are float64 sums in map order as the reference takes them
// (smartVidCrop.py:1326-1329), kept sequential on one lane and only computed when the coverage score is on.
__global__ void __launch_bounds__(128) clip_scores_kernel(const ClipDev *clips, int n_clips, const MapOut *mo, int H, int W, int n_ratios,
														   int with_cvrg, int pinned_python, double *map_scores, double *clip_scores) {
	const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (c >= n_clips) return;
	const ClipDev cl = clips[c];
	unsigned long long tot = 0;
	for (int i = lane; i < cl.n_maps; i += 32) {
		const unsigned int rs = mo[cl.map_offset + i].raw_sum;
		tot += rs;
		if (map_scores) map_scores[cl.map_offset + i] = (double)rs / (double)(H * W);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
	if (lane == 0 && clip_scores) {
		double *out = clip_scores + (size_t)c * (1 + n_ratios);
		out[0] = (double)tot / ((double)H * (double)W * (double)cl.n_maps);
		// sum(cvrg_scores) / len(cvrg_scores) (smartVidCrop.py:1329) as the interpreter computes it: the builtin sum() of
		// Python >= 3.12 adds floats with Neumaier's compensated summation; the Python the reference pins (3.7) adds
		// them left to right (pinned_python, set together with np_int_compat)
		double cv[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		if (with_cvrg)
			for (int i = 0; i < cl.n_maps; ++i)
				for (int r = 0; r < n_ratios; ++r) {
					const double x = mo[cl.map_offset + i].cvrg[r];
					const double t = __dadd_rn(cv[r], x);
					if (!pinned_python)
						cc[r] = __dadd_rn(cc[r], (fabs(cv[r]) >= fabs(x)) ? __dadd_rn(__dsub_rn(cv[r], t), x) : __dadd_rn(__dsub_rn(x, t), cv[r]));
					cv[r] = t;
				}
		for (int r = 0; r < n_ratios; ++r) {
			double sum = cv[r];
			if (cc[r] != 0.0 && isfinite(cc[r])) sum = __dadd_rn(sum, cc[r]);
			out[1 + r] = sum / (double)cl.n_maps;
		}
	}
}

// frame -> shot, frame -> clip and map -> clip tables, from the shot and clip descriptors (blocks 0 .. n_shots - 1: one
// shot each; the following n_clips blocks: the maps of one clip each).  Built on the device: 8 bytes per frame that the
// host neither fills nor sends.
__global__ void __launch_bounds__(128) index_tables_kernel(const ShotDev *__restrict__ shots, int n_shots, const ClipDev *__restrict__ clips, int n_clips,
															 int *__restrict__ frame_shot, int *__restrict__ frame_clip, int *__restrict__ map_clip) {
	const int b = blockIdx.x;
	if (b < n_shots) {
		const ShotDev sh = shots[b];
		const int n = sh.f1 - sh.f0 + 1;
		for (int i = threadIdx.x; i < n; i += blockDim.x) {
			frame_shot[sh.frame_base + i] = b;
			frame_clip[sh.frame_base + i] = sh.clip;
		}
	} else {
		const ClipDev cd = clips[b - n_shots];
		for (int i = threadIdx.x; i < cd.n_maps; i += blockDim.x) map_clip[cd.map_offset + i] = b - n_shots;
	}
}

// [H][W][N] (reference layout, frame index fastest) -> [N][H][WPS], every clip of the batch in one launch.
// The clips' [H][W][n_maps] blocks are packed back to back in src_all (clip i starts at map_offset * H * W).
// One CTA moves a tile of 64 pixels of one image row x 128 consecutive maps of one clip:
//   in : per pixel a run of <= 128 bytes at an arbitrary alignment (the runs of neighbouring pixels follow each other in
//        memory when the tile covers all maps of the clip) -> a warp copies the <= 33 aligned 32-bit words that cover
//        it into one shared row (33 words: conflict-free below);
//   out: a thread takes a 4 pixel x 4 map block: four unaligned words from shared memory (two loads + a funnel shift
//        each), a 4 x 4 byte transposition in registers (8 PRMT), four 32-bit stores -- the 8 threads of a warp that
//        share a map write 32 consecutive bytes of its row.  ~3 instructions per byte moved (the first version scattered
//        single bytes into a transposed tile: 34, issue-bound at 0.80 ms per 726 MB step).
// grid: x = map tile x pixel tile (pixel tile fastest), y = image row, z = clip.
constexpr int kTrMaps = 60;      // (tile of transpose_to_hwn_kernel)
constexpr int kTrStrideW = 65;   // words per tile row there (>= WPS / 4 + 1, odd)
constexpr int kTrTN = 128, kTrTX = 64, kTrRowW = 33;

__global__ void __launch_bounds__(256) transpose_hwn_kernel(const uint8_t *__restrict__ src_all, size_t src_bytes,
															const ClipDev *__restrict__ clips, int H, int W,
															uint8_t *__restrict__ dst, int WPS) {
	__shared__ uint32_t S32[kTrTX * kTrRowW];
	const ClipDev cd = clips[blockIdx.z];
	const int N = cd.n_maps;
	const int xtiles = (W + kTrTX - 1) / kTrTX;
	const int n0 = ((int)blockIdx.x / xtiles) * kTrTN;
	if (n0 >= N) return;
	const int nb = min(kTrTN, N - n0);
	const int x0 = ((int)blockIdx.x % xtiles) * kTrTX;
	const int nx = min(kTrTX, W - x0);
	const int y = blockIdx.y;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uintptr_t lo = reinterpret_cast<uintptr_t>(src_all), hi = lo + src_bytes;
	const uintptr_t run0 = lo + (size_t)cd.map_offset * H * W + ((size_t)y * W + x0) * N + n0;   // run of pixel x0; + N per pixel
	// all offsets below are 32-bit, relative to the aligned word that holds the first byte of the tile
	const uint32_t r0 = (uint32_t)(run0 & 3);
	const uint32_t *base32 = reinterpret_cast<const uint32_t *>(run0 - r0);
	// the aligned words that cover the tile's runs lie inside the source buffer (always, except at its very ends)
	const bool inside = (run0 - r0 >= lo) && (run0 + (size_t)(nx - 1) * N + nb + 3 <= hi);
	// ---- in ----  (every load of a thread is issued before its first store: 9 independent requests in flight)
	if (inside) {
		auto load_word = [&](int xl, int w) -> uint32_t {
			const uint32_t bo = r0 + (uint32_t)(xl * N);          // byte offset of the run from the aligned base
			uint32_t v = 0;
			if (xl < nx && 4u * (uint32_t)w < (bo & 3u) + (uint32_t)nb) v = __ldg(base32 + (bo >> 2) + (uint32_t)w);
			return v;
		};
		uint32_t v8[8];
#pragma unroll
		for (int i = 0; i < 8; ++i) v8[i] = load_word(warp + 8 * i, lane);
		const uint32_t vx = (tid < kTrTX) ? load_word(tid, 32) : 0u;
#pragma unroll
		for (int i = 0; i < 8; ++i) S32[(warp + 8 * i) * kTrRowW + lane] = v8[i];
		if (tid < kTrTX) S32[tid * kTrRowW + 32] = vx;
	} else {
		// a tile at the very start or end of the source buffer: byte loads, every one checked
		for (int i = tid; i < kTrTX * kTrRowW; i += 256) {
			const int xl = i / kTrRowW, w = i - xl * kTrRowW;
			const uint32_t bo = r0 + (uint32_t)(xl * N);
			uint32_t v = 0;
			if (xl < nx && 4u * (uint32_t)w < (bo & 3u) + (uint32_t)nb) {
				const uintptr_t a = reinterpret_cast<uintptr_t>(base32 + (bo >> 2) + (uint32_t)w);
				for (int k = 0; k < 4; ++k)
					if (a + k >= lo && a + k < hi) v |= (uint32_t)__ldg(reinterpret_cast<const uint8_t *>(a + k)) << (8 * k);
			}
			S32[i] = v;
		}
	}
	__syncthreads();
	// ---- out ----
	// 16 pixel quads x 32 map quads; a warp covers 8 pixel quads x 4 map quads
#pragma unroll
	for (int it = 0; it < 2; ++it) {
		const int wt = warp + 8 * it;                   // 0 .. 15
		const int xq = (wt & 1) * 8 + (lane & 7);
		const int n4 = (wt >> 1) * 4 + (lane >> 3);
		if (4 * n4 >= nb || 4 * xq >= nx) continue;
		uint32_t w4[4];
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const int xl = 4 * xq + k;
			uint32_t v = 0;
			if (xl < nx) {
				const int o = (int)((r0 + (uint32_t)(xl * N)) & 3u) + 4 * n4;    // byte offset inside the shared row
				const uint32_t *row = S32 + xl * kTrRowW + (o >> 2);
				v = __funnelshift_r(row[0], row[1], (o & 3) * 8);
			}
			w4[k] = v;
		}
		const uint32_t t0 = __byte_perm(w4[0], w4[1], 0x5140), t1 = __byte_perm(w4[2], w4[3], 0x5140);
		const uint32_t t2 = __byte_perm(w4[0], w4[1], 0x7362), t3 = __byte_perm(w4[2], w4[3], 0x7362);
		const uint32_t o4[4] = {__byte_perm(t0, t1, 0x5410), __byte_perm(t0, t1, 0x7632), __byte_perm(t2, t3, 0x5410), __byte_perm(t2, t3, 0x7632)};
		uint8_t *out = dst + ((size_t)(cd.map_offset + n0 + 4 * n4) * H + y) * WPS + x0 + 4 * xq;
		const size_t map_stride = (size_t)H * WPS;
#pragma unroll
		for (int j = 0; j < 4; ++j)
			if (4 * n4 + j < nb) *reinterpret_cast<uint32_t *>(out + (size_t)j * map_stride) = o4[j];
	}
}

// The way back for the filtered maps a caller wants in the reference layout (vid_data['smaps'] after smart_vid_crop,
// smartVidCrop.py:2366-2373): [N][H][WPS] -> per clip [H][W][n_maps], packed back to back like the input.
//   in : per map the row's W bytes as aligned 32-bit words into the tile (row stride 65 words);
//   out: per pixel a run of <= 60 bytes at an arbitrary alignment -> half a warp writes the <= 16 aligned words that cover
//        it: whole words as 32-bit stores, the two partial words at the ends byte by byte (their other bytes belong to the
//        neighbouring runs, written by other CTAs).
__global__ void __launch_bounds__(256) transpose_to_hwn_kernel(const uint8_t *__restrict__ src, int WPS, const ClipDev *__restrict__ clips,
																int H, int W, uint8_t *__restrict__ dst_all) {
	__shared__ uint32_t tile32[kTrMaps * kTrStrideW];
	const uint8_t *tile = reinterpret_cast<const uint8_t *>(tile32);
	const ClipDev cd = clips[blockIdx.z];
	const int N = cd.n_maps;
	const int n0 = blockIdx.x * kTrMaps;
	if (n0 >= N) return;
	const int nb = min(kTrMaps, N - n0);
	const int y = blockIdx.y;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int half = lane >> 4, hl = lane & 15;
	const int words = (W + 3) >> 2;
	const uint8_t *in = src + ((size_t)(cd.map_offset + n0) * H + y) * WPS;
	const size_t map_stride = (size_t)H * WPS;
	for (int i = tid; i < nb * 64; i += 256) {
		const int r = i >> 6, c = i & 63;
		if (c < words) tile32[r * kTrStrideW + c] = __ldg(reinterpret_cast<const uint32_t *>(in + (size_t)r * map_stride) + c);
	}
	__syncthreads();
	const uintptr_t clip0 = reinterpret_cast<uintptr_t>(dst_all) + (size_t)cd.map_offset * H * W;
	for (int x = warp * 2 + half; x < W; x += 16) {
		const uintptr_t seg = clip0 + ((size_t)y * W + x) * N + n0;
		const uintptr_t a = (seg & ~(uintptr_t)3) + 4u * hl;
		const int nl0 = (int)((long long)a - (long long)seg);       // map index (inside the tile) of byte 0 of this word
		if (nl0 >= nb || nl0 + 3 < 0) continue;
		uint32_t v = 0;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const int nl = nl0 + k;
			if (nl >= 0 && nl < nb) v |= (uint32_t)tile[(size_t)nl * (kTrStrideW * 4) + x] << (8 * k);
		}
		if (nl0 >= 0 && nl0 + 3 < nb) {
			*reinterpret_cast<uint32_t *>(a) = v;
		} else {
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const int nl = nl0 + k;
				if (nl >= 0 && nl < nb) *reinterpret_cast<uint8_t *>(a + k) = (uint8_t)(v >> (8 * k));
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// The renderer's per-frame crop: sc_renderer, smartVidCrop.py:1906-1912,
//   out[f] = frame[f][by1:by2, bx1:bx2, :]
// for interleaved uint8 frames [F][H][W][C].  Pure data movement: every thread produces VEC aligned bytes of the packed
// output; the source row segment starts at an arbitrary byte (bx1 * C), so a thread loads the aligned 32-bit words that
// cover its bytes (neighbouring threads share them through L1) and funnel-shifts them into place.  VEC = 8 or 4 when the
// output row length is a multiple of it, else 1 (byte per thread).
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) crop_frames_kernel(const uint8_t *__restrict__ frames, int n_frames, int H, int W, int C,
														   const int32_t *__restrict__ boxes, int oh, int ow, uint8_t *__restrict__ out) {
	const long long row_bytes = (long long)ow * C;
	const long long vec_per_row = row_bytes / VEC;
	const long long total = (long long)n_frames * oh * vec_per_row;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const long long r = i / vec_per_row;              // output row (frame * oh + y)
		const int v = (int)(i - r * vec_per_row);
		const int f = (int)(r / oh), y = (int)(r - (long long)f * oh);
		const int bx1 = boxes[f * 4 + 0], by1 = boxes[f * 4 + 1];
		const uint8_t *src = frames + (((long long)f * H + (by1 + y)) * W + bx1) * C + (long long)v * VEC;
		uint8_t *dst = out + r * row_bytes + (long long)v * VEC;
		if constexpr (VEC == 1) {
			*dst = *src;
		} else {
			const uintptr_t a = reinterpret_cast<uintptr_t>(src);
			const uint32_t *w0 = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
			const unsigned sh = (unsigned)(a & 3) * 8u;
			uint32_t w[VEC / 4 + 1];
#pragma unroll
			for (int k = 0; k <= VEC / 4; ++k) w[k] = (k < VEC / 4 || sh != 0u) ? __ldg(w0 + k) : 0u;   // (no read past the segment when aligned)
			uint32_t o[VEC / 4];
#pragma unroll
			for (int k = 0; k < VEC / 4; ++k) o[k] = __funnelshift_r(w[k], w[k + 1], sh);
			if constexpr (VEC == 8) *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0], o[1]);
			else *reinterpret_cast<uint32_t *>(dst) = o[0];
		}
	}
}

}  // namespace rvb
