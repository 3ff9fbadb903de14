// C ABI of the B200-native crop-selection path (include/retargetvid_b200.h): context, workspace,
// metadata staging and kernel orchestration.  The arithmetic lives in the *.cuh kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <complex>
#include <string>
#include <mutex>
#include <vector>

#include "../../include/retargetvid_b200.h"
#include "iou_kernel.cuh"
#include <chrono>
#include "map_kernel.cuh"
#include "prim_kernel.cuh"
#include "fprim_kernel.cuh"
#include "track_kernels.cuh"

using namespace rvb;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_err = buf;
	return code;
}

#define CU(call)                                                                                           \
	do {                                                                                                   \
		cudaError_t e_ = (call);                                                                           \
		if (e_ != cudaSuccess) return fail(RVB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

// ---------------------------------------------------------------------------------------------
// growable device / pinned buffers
// ---------------------------------------------------------------------------------------------
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	int ensure(size_t need) {
		if (need <= cap) return RVB_OK;
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
		size_t want = need + need / 4 + 4096;
		CU(cudaMalloc(&p, want));
		cap = want;
		return RVB_OK;
	}
	void release() {
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
	}
};

struct PinBuf {
	void *p = nullptr;
	size_t cap = 0;
	int ensure(size_t need) {
		if (need <= cap) return RVB_OK;
		if (p) cudaFreeHost(p);
		p = nullptr;
		cap = 0;
		size_t want = need + need / 4 + 4096;
		CU(cudaMallocHost(&p, want));
		cap = want;
		return RVB_OK;
	}
	void release() {
		if (p) cudaFreeHost(p);
		p = nullptr;
		cap = 0;
	}
};

struct rvb_ctx {
	int device = 0;
	int n_sm = 0;
	cudaStream_t own_stream = nullptr;
	cudaStream_t side_stream = nullptr;   // cut-adjacent chains (and monolithic fallbacks) run beside the main set
	cudaStream_t cls_stream[kSplitClasses] = {};   // one per size class: Prim + back of the classes run side by side
	cudaEvent_t ev_cls[kSplitClasses] = {};
	cudaEvent_t ev_cls_fork = nullptr;
	bool serial_classes = true;           // the size classes one after the other on the main stream (default: measured faster
	                                      // with several batches in flight); RVB_FAN_CLASSES=1: each class on its own stream
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev_map0 = nullptr, ev_map1 = nullptr, ev_stage = nullptr;
	cudaEvent_t ev_iou0 = nullptr, ev_iou1 = nullptr;
	bool iou_timed = false;
	cudaEvent_t ev_st[3] = {nullptr, nullptr, nullptr};   // split pipeline, main stream: after front / Prim / back of set 0
	bool stages_timed = false;
	bool stage_busy = false;
	bool map_timed = false;
	int map_launches = 0;
	int64_t launches = 0;
	bool phase_on = false;
	DevBuf phase;
	bool chain_levels = true;   // cut-adjacent chains go through the split pipeline, one depth per set, on the side stream;
	                            // RVB_CHAIN_MONO=1: the monolithic kernel walks them instead (the round-1 default)
	struct CachedFilter { int order; double wn; FilterCoef fc; };
	std::vector<CachedFilter> filter_cache;   // Butterworth designs by (order, Wn)
	double host_us[6] = {0, 0, 0, 0, 0, 0};   // host time of the last crop_track call by section (rvb_ctx_last_host_us)
	bool chain_levels_dense = false;   // RVB_CHAIN_LEVELS=1: also with the all-pairs Prim (slower: 3 more sets of 11 small launches)
	int prim_variant[4] = {0, 0, 0, 0};
	bool split = true;   // front -> prim_kernel -> back pipeline (RVB_NO_SPLIT=1 keeps every map in the monolithic kernel)
	DevBuf maps_in, maps_nhw, filt, filt_hwn, meta, mapout, series, scratch, boxes, misc, iou_a, iou_b, iou_c;
	DevBuf scr_pinfo, scr_val, scr_pkey, scr_far, scr_alist;
	bool dense_prim = true;    // the all-pairs prim_kernel (default, the faster one as measured: DESIGN.md 4.1);
	                           // RVB_FRONTIER_PRIM=1: the lattice-local fprim_kernel (exact too, kept for the record)
	const void *nhw_zero_p = nullptr;   // maps_nhw was cleared at this address for this geometry
	size_t nhw_zero_cap = 0;
	int nhw_zero_w = 0, nhw_zero_wps = 0, nhw_zero_h = 0;
	PinBuf stage, stage_out;
};

// ---------------------------------------------------------------------------------------------
// shared-memory layout of the map kernel for a capacity of nmax points
// ---------------------------------------------------------------------------------------------
static const int kPrimVariantDefault[4] = {0, 0, 0, 0};
static int align_up(int v, int a) { return (v + a - 1) / a * a; }
static const int kMaxDynSmem = 227 * 1024 - 1024;  // per-CTA opt-in limit minus the kernel's static shared memory

static SmemLayout make_layout(int nmax, int H, int WPS, int W, int mcs, int small_bytes, bool front = false) {
	SmemLayout L;
	memset(&L, 0, sizeof(L));
	int o = 0;
	L.nmax = nmax;
	// a cluster has >= mcs points and leaf clusters are disjoint: at most 2 * nmax / mcs + 1 clusters
	L.ncmax = std::min(2 * nmax / std::max(mcs, 2) + 3, nmax > 4096 ? 770 : nmax / 2 + 2);
	L.pts = o; o += align_up(std::max(2 * nmax, 4 * std::max(H, W)), 16);
	L.val = o; o += align_up(nmax, 16);
	L.small = o; o += align_up(small_bytes, 16);
	o = align_up(o, 128);
	L.u_base = o;
	L.map = o;
	int c = o;
	if (front) {
		// front half of the split pipeline: only the core distances, the occupancy mask and the fallback queue
		// share the map's storage
		L.a4 = c; c += 4 * nmax;
		L.d4 = L.mask = c; c += align_up(H * ((W + 31) / 32) * 4, 16);
		L.queue = c; c += 2 * nmax;
		L.total = align_up(std::max(c, L.map + H * WPS), 128);
		return L;
	}
	L.a4 = c; c += 4 * nmax;
	L.order = c; c += 2 * nmax;
	L.wp = c; c += 4 * nmax;
	// d4, rank, pe, pl are contiguous: the occupancy mask of phase 2 aliases them
	L.d4 = c; c += 4 * nmax;
	L.rank = c; c += 2 * nmax;
	L.pe = c; c += 2 * nmax;
	L.pl = c; c += 2 * nmax;
	const int mask_bytes = H * ((W + 31) / 32) * 4;
	if (mask_bytes > 10 * nmax) c = L.d4 + align_up(mask_bytes, 16);
	L.mask = L.d4;
	L.queue = c; c += 2 * nmax;
	L.pnode = L.d4;
	c = align_up(c, 16);
	L.cl_stab = c; c += 8 * L.ncmax;
	L.cl_birth = c; c += 8 * L.ncmax;
	L.cl_acc = c; c += 4 * L.ncmax;
	L.cl_parent = c; c += 2 * L.ncmax;
	L.cl_ch0 = c; c += 2 * L.ncmax;
	L.cl_ch1 = c; c += 2 * L.ncmax;
	L.cl_label = c; c += 2 * L.ncmax;
	L.cl_selanc = c; c += 2 * L.ncmax;
	const int cluster_end = c;
	const int map_end = L.map + H * WPS;
	L.total = align_up(std::max(cluster_end, map_end), 128);
	return L;
}

// lattice offsets of the frontier Prim (fprim_kernel.cuh), sorted by squared length, and the weight of each bucket level
static void build_frontier_table(FrOffsetTable &t) {
	struct Off { int d2, dy, dx; };
	std::vector<Off> offs;
	for (int dy = -5; dy <= 5; ++dy)
		for (int dx = -5; dx <= 5; ++dx) {
			const int d2 = dy * dy + dx * dx;
			if (d2 > 0 && d2 <= kFrR0) offs.push_back({d2, dy, dx});
		}
	std::stable_sort(offs.begin(), offs.end(), [](const Off &a, const Off &b) { return a.d2 < b.d2; });
	memset(&t, 0, sizeof(t));
	for (int i = 0; i < kFrIters * 32; ++i) {
		if (i < (int)offs.size()) { t.dy[i] = (int8_t)offs[i].dy; t.dx[i] = (int8_t)offs[i].dx; t.lev[i] = (uint8_t)fr_level_of((uint32_t)offs[i].d2); }
		else { t.dy[i] = 0; t.dx[i] = 0; t.lev[i] = (uint8_t)kFrLevInf; }
	}
	int l = 0;
	for (int w = 0; w < 32; ++w) if (kFrLevelMask & (1u << w)) t.level_w[l++] = (uint32_t)w;
}

static void build_ring_table(RingTable &t) {
	struct Off { int d2, dy, dx; };
	std::vector<Off> offs;
	for (int dy = -14; dy <= 14; ++dy)
		for (int dx = -14; dx <= 14; ++dx) {
			const int d2 = dy * dy + dx * dx;
			if (d2 > 0 && d2 <= kRingD2Max) offs.push_back({d2, dy, dx});
		}
	std::stable_sort(offs.begin(), offs.end(), [](const Off &a, const Off &b) { return a.d2 < b.d2; });
	memset(&t, 0, sizeof(t));
	int nr = 0;
	for (size_t i = 0; i < offs.size(); ++i) {
		t.dy[i] = (int8_t)offs[i].dy;
		t.dx[i] = (int8_t)offs[i].dx;
		if (i + 1 == offs.size() || offs[i + 1].d2 != offs[i].d2) {
			t.ring_end[nr] = (uint16_t)(i + 1);
			t.ring_d2[nr] = (uint16_t)offs[i].d2;
			++nr;
		}
	}
	t.n_rings = nr;
	t.n_offsets = (int)offs.size();
}

// A Butterworth filter of high order and low cut-off is ill conditioned in transfer-function form: its output depends
// on the ORDER of the floating-point operations at the 1e-7 .. 1e-3 level (order 7 at 1 Hz / 30 fps: 4e-4), and scipy's
// own output carries that error.  lowpass_kernel evaluates filtfilt in 32 chunks per pass (filtfilt_warp); this probe
// runs that evaluation and the sequential one on a test signal (two lengths, values in pixel range) and allows the chunked
// form only where the two agree to 2e-9 -- there is a wide gap: order 5 at 2 Hz / 30 fps (the default) 4.5e-10, every
// well-conditioned filter below that; order 5 at 1 Hz 1e-7, order 7 at 2 Hz 1.7e-7.
static bool filtfilt_chunked_probe(const FilterCoef &fc) {
	const int M = fc.order;
	if (M < 1) return true;
	auto step = [&](double *z, double x) -> double {
		const double y = z[0] + fc.b[0] * x;
		for (int i = 0; i < M - 1; ++i) z[i] = (z[i + 1] + x * fc.b[i + 1]) - y * fc.a[i + 1];
		z[M - 1] = x * fc.b[M] - y * fc.a[M];
		return y;
	};
	double worst = 0.0;
	const int lengths[2] = {2700, 330};
	for (int t = 0; t < 2; ++t) {
		const int ne = lengths[t];
		std::vector<double> x(ne), seq(ne), chk(ne);
		for (int i = 0; i < ne; ++i) x[i] = 125.0 + 100.0 * sin(i / 27.5) + 10.0 * sin(i / 1.16) + ((i * 2654435761u >> 16) & 255) / 128.0;
		seq = x; chk = x;
		const int L = (ne + 31) >> 5;
		double P[RVB_MAX_LP_ORDER][RVB_MAX_LP_ORDER];
		for (int c = 0; c < M; ++c) {
			double z[RVB_MAX_LP_ORDER] = {0};
			z[c] = 1.0;
			for (int n = 0; n < L; ++n) step(z, 0.0);
			for (int r = 0; r < M; ++r) P[r][c] = z[r];
		}
		for (int pass = 0; pass < 2; ++pass) {
			auto at = [&](int i) { return pass ? ne - 1 - i : i; };
			// sequential
			double z[RVB_MAX_LP_ORDER];
			for (int i = 0; i < M; ++i) z[i] = fc.zi[i] * seq[at(0)];
			for (int i = 0; i < ne; ++i) seq[at(i)] = step(z, seq[at(i)]);
			// chunked
			double E[32][RVB_MAX_LP_ORDER], S[32][RVB_MAX_LP_ORDER];
			for (int k = 0; k < 32; ++k) {
				double e[RVB_MAX_LP_ORDER] = {0};
				for (int i = std::min(ne, k * L); i < std::min(ne, k * L + L); ++i) step(e, chk[at(i)]);
				for (int r = 0; r < M; ++r) E[k][r] = e[r];
			}
			for (int r = 0; r < M; ++r) S[0][r] = fc.zi[r] * chk[at(0)];
			for (int k = 1; k < 32; ++k)
				for (int r = 0; r < M; ++r) {
					double acc = E[k - 1][r];
					for (int c = 0; c < M; ++c) acc += P[r][c] * S[k - 1][c];
					S[k][r] = acc;
				}
			for (int k = 0; k < 32; ++k) {
				double zz[RVB_MAX_LP_ORDER];
				for (int r = 0; r < M; ++r) zz[r] = S[k][r];
				for (int i = std::min(ne, k * L); i < std::min(ne, k * L + L); ++i) chk[at(i)] = step(zz, chk[at(i)]);
			}
		}
		for (int i = 0; i < ne; ++i) {
			const double d = fabs(seq[i] - chk[i]);
			if (!(d <= worst)) worst = d;      // (NaN counts as a failure)
		}
	}
	return worst <= 2e-9;
}

// ---------------------------------------------------------------------------------------------
// Butterworth low-pass design: scipy.signal.butter(order, Wn, 'lowpass') -> (b, a), and
// scipy.signal.lfilter_zi(b, a) -- the calls made at smartVidCrop.py:1601-1605 (via filtfilt).
// ---------------------------------------------------------------------------------------------
static void butter_design(int order, double wn, FilterCoef &fc) {
	memset(&fc, 0, sizeof(fc));
	if (!(wn > 0.0 && wn < 1.0) || order < 1 || order > RVB_MAX_LP_ORDER) {
		fc.order = 0;  // scipy raises -> the reference falls back to moving averages
		return;
	}
	typedef std::complex<double> cd;
	const double pi = 3.14159265358979323846;
	// buttap: poles of the analog prototype
	std::vector<cd> p(order);
	for (int i = 0; i < order; ++i) {
		const int m = -order + 1 + 2 * i;
		p[i] = -std::exp(cd(0.0, pi * m / (2.0 * order)));
	}
	double k = 1.0;
	// pre-warp (fs = 2), lp2lp_zpk
	const double fs = 2.0;
	const double warped = 2.0 * fs * tan(pi * wn / fs);
	for (auto &v : p) v *= warped;
	k *= pow(warped, (double)order);
	// bilinear_zpk
	const double fs2 = 2.0 * fs;
	cd den(1.0, 0.0);
	std::vector<cd> pz(order);
	for (int i = 0; i < order; ++i) {
		pz[i] = (fs2 + p[i]) / (fs2 - p[i]);
		den *= (fs2 - p[i]);
	}
	const double kz = k * (cd(1.0, 0.0) / den).real();
	// zpk2tf: zeros are all at -1
	std::vector<cd> a(1, cd(1.0, 0.0)), b(1, cd(1.0, 0.0));
	for (int i = 0; i < order; ++i) {
		std::vector<cd> na(a.size() + 1, cd(0.0, 0.0)), nb(b.size() + 1, cd(0.0, 0.0));
		for (size_t j = 0; j < a.size(); ++j) {
			na[j] += a[j];
			na[j + 1] -= a[j] * pz[i];
			nb[j] += b[j];
			nb[j + 1] += b[j];  // (z - (-1))
		}
		a = na;
		b = nb;
	}
	fc.order = order;
	for (int i = 0; i <= order; ++i) {
		fc.a[i] = a[i].real();
		fc.b[i] = kz * b[i].real();
	}
	// lfilter_zi: solve (I - A^T) zi = b[1:] - a[1:] b[0] with the companion matrix of a (a[0] == 1)
	double asum = 1.0, bsum = 0.0;
	for (int i = 1; i <= order; ++i) {
		asum += fc.a[i];
		bsum += fc.b[i] - fc.a[i] * fc.b[0];
	}
	fc.zi[0] = bsum / asum;
	double as = 1.0, cs = 0.0;
	for (int i = 1; i < order; ++i) {
		as += fc.a[i];
		cs += fc.b[i] - fc.a[i] * fc.b[0];
		fc.zi[i] = as * fc.zi[0] - cs;
	}
	fc.chunked_ok = filtfilt_chunked_probe(fc) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// cv2.resize coefficient tables (OpenCV resize.cpp): float32 source coordinate, weights * 2048 rounded
// half-to-even; the horizontal table clamps at the borders, the vertical one keeps its weights (rows are
// clipped by the kernel).  Layout per axis: idx[d], w0[d], w1[d].
// ---------------------------------------------------------------------------------------------
static int cv_round_d(double v) { return (int)nearbyint(v); }

static void linear_table(int dsize, int ssize, double scale, bool vertical, std::vector<int16_t> &out) {
	const size_t base = out.size();
	out.resize(base + 3 * (size_t)dsize);
	for (int d = 0; d < dsize; ++d) {
		float f = (float)((d + 0.5) * scale - 0.5);
		int sidx = (int)floorf(f);
		f -= (float)sidx;
		if (!vertical) {
			if (sidx < 0) { sidx = 0; f = 0.f; }
			if (sidx >= ssize - 1) { sidx = ssize - 1; f = 0.f; }
		}
		out[base + d] = (int16_t)sidx;
		out[base + dsize + d] = (int16_t)nearbyintf((1.f - f) * 2048.f);
		out[base + 2 * (size_t)dsize + d] = (int16_t)nearbyintf(f * 2048.f);
	}
}

// INTER_CUBIC (OpenCV resize.cpp): interpolateCubic with A = -0.75 in float32, weights rounded half-to-even to 1/2048.
// Layout: idx[dsize] (position of tap 1; taps idx - 1 .. idx + 2 are clamped by the kernel), then 4 x w[dsize].
static void cubic_table(int dsize, double scale, std::vector<int16_t> &out) {
	const size_t base = out.size();
	out.resize(base + 5 * (size_t)dsize);
	for (int d = 0; d < dsize; ++d) {
		float x = (float)((d + 0.5) * scale - 0.5);
		const int sidx = (int)floorf(x);
		x -= (float)sidx;
		const float A = -0.75f;
		float cf[4];
		cf[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
		cf[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
		cf[2] = ((A + 2.f) * (1.f - x) - (A + 3.f)) * (1.f - x) * (1.f - x) + 1.f;
		cf[3] = 1.f - cf[0] - cf[1] - cf[2];
		out[base + d] = (int16_t)sidx;
		for (int k = 0; k < 4; ++k) out[base + (size_t)(1 + k) * dsize + d] = (int16_t)nearbyintf(cf[k] * 2048.f);
	}
}

static void nearest_table(int dsize, int ssize, double fx, std::vector<int16_t> &out) {
	const double ifx = 1.0 / fx;
	for (int d = 0; d < dsize; ++d) out.push_back((int16_t)std::min((int)floor(d * ifx), ssize - 1));
}

struct ResizeSetup {
	int on = 0, Hs = 0, Ws = 0, WSs = 0;
	int dx = 0, dy = 0, ux = 0, uy = 0, nx = 0, ny = 0, cx = 0, cy = 0;
	std::vector<int16_t> tab;
};

static void build_resize(double factor, int H, int W, ResizeSetup &r) {
	r.on = 1;
	const double fx = 1.0 / factor;
	r.Ws = cv_round_d(W * fx);
	r.Hs = cv_round_d(H * fx);
	r.WSs = align_up(r.Ws, 16);
	r.dx = (int)r.tab.size(); linear_table(r.Ws, W, 1.0 / fx, false, r.tab);
	r.dy = (int)r.tab.size(); linear_table(r.Hs, H, 1.0 / fx, true, r.tab);
	r.ux = (int)r.tab.size(); linear_table(W, r.Ws, 1.0 / ((double)W / r.Ws), false, r.tab);
	r.uy = (int)r.tab.size(); linear_table(H, r.Hs, 1.0 / ((double)H / r.Hs), true, r.tab);
	r.nx = (int)r.tab.size(); nearest_table(r.Ws, W, fx, r.tab);
	r.ny = (int)r.tab.size(); nearest_table(r.Hs, H, fx, r.tab);
	r.cx = (int)r.tab.size(); cubic_table(r.Ws, 1.0 / fx, r.tab);
	r.cy = (int)r.tab.size(); cubic_table(r.Hs, 1.0 / fx, r.tab);
}

// sc_calc_dest_size -- smartVidCrop.py:946-977
static void calc_dest_size(int w_orig, int h_orig, double tw, double th, int out[3]) {
	const double orig_ratio = (double)w_orig / (double)h_orig;
	const double target_ratio = tw / th;
	if (fabs(orig_ratio - target_ratio) < 0.0000001) {
		out[0] = 0; out[1] = w_orig; out[2] = h_orig;
		return;
	}
	int w_final = (int)floor((tw / th) * h_orig);
	int h_final = h_orig;
	int mode = 1;
	if (w_final > w_orig || h_final > h_orig) {
		w_final = w_orig;
		h_final = (int)floor((th / tw) * w_orig);
		mode = 2;
	}
	out[0] = mode; out[1] = w_final; out[2] = h_final;
}

// ---------------------------------------------------------------------------------------------
// API
// ---------------------------------------------------------------------------------------------
extern "C" const char *rvb_version(void) { return "retargetvid_b200 0.1.0 (smartVidCrop 1.4.0 hot path, sm_100a)"; }
extern "C" const char *rvb_last_error(void) { return g_err.c_str(); }

extern "C" int rvb_ctx_create(int device, rvb_ctx **out) {
	if (!out) return fail(RVB_ERR_INVALID, "rvb_ctx_create: out is NULL");
	*out = nullptr;
	int n = 0;
	CU(cudaGetDeviceCount(&n));
	if (device < 0 || device >= n) return fail(RVB_ERR_INVALID, "rvb_ctx_create: device %d of %d", device, n);
	CU(cudaSetDevice(device));
	rvb_ctx *c = new rvb_ctx();
	c->device = device;
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	c->n_sm = prop.multiProcessorCount;
	CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
	c->stream = c->own_stream;
	CU(cudaEventCreate(&c->ev_iou0));
	CU(cudaEventCreate(&c->ev_iou1));
	CU(cudaEventCreate(&c->ev_map0));
	CU(cudaEventCreate(&c->ev_map1));
	for (int i = 0; i < 3; ++i) CU(cudaEventCreate(&c->ev_st[i]));
	CU(cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming));
	{
		// experiment knobs (profiling only): RVB_SIDE_PRIO = CUDA priority of the side stream, RVB_CHAIN_LEVELS=1
		const char *e = getenv("RVB_SIDE_PRIO");
		int lo = 0, hi = 0;
		cudaDeviceGetStreamPriorityRange(&lo, &hi);
		int prio = e ? atoi(e) : 0;
		prio = std::max(hi, std::min(lo, prio));
		CU(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, prio));
		e = getenv("RVB_CHAIN_MONO");
		c->chain_levels = !(e && e[0] == '1');
		e = getenv("RVB_CHAIN_LEVELS");
		c->chain_levels_dense = e && e[0] == '1';
	}
	for (int k = 0; k < kSplitClasses; ++k) {
		CU(cudaStreamCreateWithFlags(&c->cls_stream[k], cudaStreamNonBlocking));
		CU(cudaEventCreateWithFlags(&c->ev_cls[k], cudaEventDisableTiming));
	}
	CU(cudaEventCreateWithFlags(&c->ev_cls_fork, cudaEventDisableTiming));
	{
		const char *e = getenv("RVB_FAN_CLASSES");
		c->serial_classes = !(e && e[0] == '1');
	}
	CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
	RingTable t;
	build_ring_table(t);
	CU(cudaMemcpyToSymbol(c_rings, &t, sizeof(t)));
	{
		FrOffsetTable ft;
		build_frontier_table(ft);
		CU(cudaMemcpyToSymbol(c_froffs, &ft, sizeof(ft)));
		CU(cudaFuncSetAttribute(fprim_kernel<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
		CU(cudaFuncSetAttribute(fprim_kernel<1536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
		CU(cudaFuncSetAttribute(fprim_kernel<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
		CU(cudaFuncSetAttribute(fprim_kernel<3072>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
		CU(cudaFuncSetAttribute(fprim_kernel<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
		CU(cudaFuncSetAttribute(fprim_kernel<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
		const char *e = getenv("RVB_FRONTIER_PRIM");
		c->dense_prim = !(e && e[0] == '1');
	}
	CU(cudaFuncSetAttribute(map_kernel<256, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CU(cudaFuncSetAttribute(map_kernel<256, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CU(cudaFuncSetAttribute(map_kernel<512, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CU(cudaFuncSetAttribute(map_kernel<512, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CU(cudaFuncSetAttribute(map_kernel<512, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
	CU(cudaFuncSetAttribute(map_kernel<256, 16, kModeFront>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CU(cudaFuncSetAttribute(map_kernel<512, 8, kModeBack>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CU(cudaFuncSetAttribute(map_kernel<512, 16, kModeFront>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CU(cudaFuncSetAttribute(map_kernel<512, 16, kModeBack>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
	CU(cudaFuncSetAttribute(prim_kernel<8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	CU(cudaFuncSetAttribute(map_kernel<256, 6, kModeBack>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CU(cudaFuncSetAttribute(map_kernel<256, 8, kModeBack>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CU(cudaFuncSetAttribute(map_kernel<512, 6, kModeBack>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	CU(cudaFuncSetAttribute(prim_kernel<1, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	CU(cudaFuncSetAttribute(prim_kernel<2, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	CU(cudaFuncSetAttribute(prim_kernel<4, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	{
		const char *e = getenv("RVB_NO_SPLIT");
		c->split = !(e && e[0] == '1');
		// RVB_PRIM_VARIANT="abcd": per Prim size class 0..3, '1' selects the shape with more warps and fewer register
		// slots per thread (<2,12> <4,12> <4,16> <8,12> instead of <1,24> <2,24> <4,24> <4,24>)
		const char *v = getenv("RVB_PRIM_VARIANT");
		for (int k = 0; k < 4; ++k) c->prim_variant[k] = (v && (int)strlen(v) > k && v[k] == '1') ? 1 : kPrimVariantDefault[k];
	}
	*out = c;
	return RVB_OK;
}

extern "C" int rvb_ctx_destroy(rvb_ctx *c) {
	if (!c) return RVB_OK;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	DevBuf *bufs[] = {&c->maps_in, &c->maps_nhw, &c->filt, &c->filt_hwn, &c->meta, &c->mapout, &c->series, &c->scratch,
					  &c->boxes, &c->misc, &c->iou_a, &c->iou_b, &c->iou_c, &c->scr_pinfo, &c->scr_val, &c->scr_pkey,
					  &c->scr_far, &c->scr_alist};
	for (DevBuf *b : bufs) b->release();
	c->stage.release();
	c->stage_out.release();
	if (c->ev_iou0) cudaEventDestroy(c->ev_iou0);
	if (c->ev_iou1) cudaEventDestroy(c->ev_iou1);
	if (c->ev_map0) cudaEventDestroy(c->ev_map0);
	if (c->ev_map1) cudaEventDestroy(c->ev_map1);
	for (int i = 0; i < 3; ++i) if (c->ev_st[i]) cudaEventDestroy(c->ev_st[i]);
	if (c->ev_stage) cudaEventDestroy(c->ev_stage);
	if (c->ev_fork) cudaEventDestroy(c->ev_fork);
	if (c->ev_join) cudaEventDestroy(c->ev_join);
	if (c->side_stream) cudaStreamDestroy(c->side_stream);
	for (int k = 0; k < kSplitClasses; ++k) {
		if (c->ev_cls[k]) cudaEventDestroy(c->ev_cls[k]);
		if (c->cls_stream[k]) cudaStreamDestroy(c->cls_stream[k]);
	}
	if (c->ev_cls_fork) cudaEventDestroy(c->ev_cls_fork);
	if (c->own_stream) cudaStreamDestroy(c->own_stream);
	delete c;
	return RVB_OK;
}

extern "C" int rvb_ctx_set_stream(rvb_ctx *c, void *stream) {
	if (!c) return fail(RVB_ERR_INVALID, "ctx is NULL");
	c->stream = stream ? (cudaStream_t)stream : c->own_stream;
	return RVB_OK;
}

extern "C" int rvb_ctx_synchronize(rvb_ctx *c) {
	if (!c) return fail(RVB_ERR_INVALID, "ctx is NULL");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	return RVB_OK;
}

extern "C" int64_t rvb_ctx_launch_count(const rvb_ctx *c) { return c ? c->launches : 0; }

extern "C" int rvb_ctx_last_map_kernel_ms(rvb_ctx *c, float *ms, int32_t *launches) {
	if (!c) return fail(RVB_ERR_INVALID, "ctx is NULL");
	if (!c->map_timed) return fail(RVB_ERR_INVALID, "no crop_track call has run on this context");
	CU(cudaEventSynchronize(c->ev_map1));
	float t = 0.f;
	CU(cudaEventElapsedTime(&t, c->ev_map0, c->ev_map1));
	if (ms) *ms = t;
	if (launches) *launches = c->map_launches;
	return RVB_OK;
}

extern "C" int rvb_ctx_last_iou_kernel_ms(rvb_ctx *c, float *ms) {
	if (!c || !ms) return fail(RVB_ERR_INVALID, "NULL argument");
	if (!c->iou_timed) return fail(RVB_ERR_INVALID, "no iou_batch call has run on this context");
	CU(cudaEventSynchronize(c->ev_iou1));
	CU(cudaEventElapsedTime(ms, c->ev_iou0, c->ev_iou1));
	return RVB_OK;
}

extern "C" int rvb_ctx_last_stage_ms(rvb_ctx *c, float out[4]) {
	if (!c || !out) return fail(RVB_ERR_INVALID, "NULL argument");
	if (!c->map_timed || !c->stages_timed) return fail(RVB_ERR_INVALID, "the last crop_track call did not run the split pipeline");
	CU(cudaEventSynchronize(c->ev_map1));
	CU(cudaEventElapsedTime(&out[0], c->ev_map0, c->ev_st[0]));
	CU(cudaEventElapsedTime(&out[1], c->ev_st[0], c->ev_st[1]));
	CU(cudaEventElapsedTime(&out[2], c->ev_st[1], c->ev_st[2]));
	CU(cudaEventElapsedTime(&out[3], c->ev_map0, c->ev_map1));
	return RVB_OK;
}

extern "C" int rvb_ctx_last_host_us(rvb_ctx *c, double out[6]) {
	if (!c || !out) return fail(RVB_ERR_INVALID, "NULL argument");
	for (int i = 0; i < 6; ++i) out[i] = c->host_us[i];
	return RVB_OK;
}

extern "C" int rvb_ctx_phase_cycles(rvb_ctx *c, int enable, uint64_t out[16]) {
	if (!c) return fail(RVB_ERR_INVALID, "ctx is NULL");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	if (c->phase.ensure(16 * sizeof(uint64_t))) return RVB_ERR_CUDA;
	if (out) {
		if (c->phase_on) CU(cudaMemcpy(out, c->phase.p, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
		else memset(out, 0, 16 * sizeof(uint64_t));
	}
	CU(cudaMemset(c->phase.p, 0, 16 * sizeof(uint64_t)));
	c->phase_on = enable != 0;
	return RVB_OK;
}

extern "C" int rvb_params_default(rvb_params *p, int use_best_settings) {
	if (!p) return fail(RVB_ERR_INVALID, "params is NULL");
	memset(p, 0, sizeof(*p));
	p->t_threshold = 120; p->clust_filt = 1; p->hdbscan_min = 26; p->hdbscan_min_samples = 0;
	p->select_sum = 2; p->op_close = 1; p->com_km = 1; p->t_border = -1;
	p->loess_filt = 1; p->loess_degree = 2; p->lp_filt = 1; p->lp_order = 5; p->shift_time = 0;
	p->exit_on_low_cvrg = 0; p->cvrg_window = 0; p->resize_type = 1; p->focus_stability = 0; p->min_d_jump = 10;
	p->skip = 6; p->np_int_compat = 0; p->foces_stab_t = 60.0; p->foces_stab_s = 1.5;
	p->loess_w_secs = 2.0; p->lp_cutoff = 2.0; p->resize_factor = 1.0; p->t_cvrg = 0.60;
	if (use_best_settings) {
		p->t_threshold = 90; p->hdbscan_min = 5; p->hdbscan_min_samples = 3; p->resize_factor = 4.0;
		p->select_sum = 1; p->lp_cutoff = 1.0; p->lp_order = 2; p->loess_filt = 0; p->focus_stability = 1; p->min_d_jump = 1;
	}
	return RVB_OK;
}

// one launch of the map kernel family
template <int NT, int TPT, int MODE = kModeMono>
static int launch_map(rvb_ctx *c, MapArgs a, int H, int W, int WPS, int grid, cudaStream_t stream = nullptr) {
	a.lay = make_layout(NT * TPT, H, WPS, W, a.mcs, a.resize_on ? a.Hs * a.WSs : 0, MODE == kModeFront);
	if (a.lay.total > (NT * TPT > 4096 ? kMaxDynSmem : 200 * 1024)) return fail(RVB_ERR_UNSUPPORTED, "process size %dx%d needs %d B of shared memory", H, W, a.lay.total);
	map_kernel<NT, TPT, MODE><<<grid, NT, a.lay.total, stream ? stream : c->stream>>>(a);
	CU(cudaGetLastError());
	c->launches += 1;
	c->map_launches += 1;
	return RVB_OK;
}

// one Prim launch: persistent CTAs of NW warps, as many per SM as the register budget of KMAX slots allows
template <int NW, int KMAX>
static void launch_prim(rvb_ctx *c, const PrimArgs &pa, int n_maps, int smem, cudaStream_t stream) {
	static bool attr_set[64] = {};
	if (c->device >= 0 && c->device < 64 && !attr_set[c->device]) {
		cudaFuncSetAttribute(prim_kernel<NW, KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
		attr_set[c->device] = true;
	}
	const int per_sm = prim_warps_per_sm(KMAX) / NW;
	prim_kernel<NW, KMAX><<<std::min(n_maps, c->n_sm * per_sm), 32 * NW, smem, stream>>>(pa);
}

// one frontier-Prim launch: persistent one-warp CTAs, as many per SM as shared memory and registers allow
template <int CAP>
static void launch_fprim(rvb_ctx *c, FPrimArgs fa, int n_maps, int H, int W, cudaStream_t stream) {
	fa.cap = CAP; fa.H = H; fa.W = W; fa.RS = (W + 31) >> 5;
	const int smem = fprim_smem_bytes(CAP, H, W);
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fprim_kernel<CAP>, 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
	fprim_kernel<CAP><<<std::max(1, std::min(n_maps, c->n_sm * per_sm)), 32, smem, stream>>>(fa);
}

template <int NT, int TPT, int MODE = kModeMono>
static int occupancy_grid(rvb_ctx *c, int smem, int n_items) {
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, map_kernel<NT, TPT, MODE>, NT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
	int g = c->n_sm * per_sm;
	if (n_items >= 0) g = std::min(g, std::max(n_items, 1));
	return std::max(g, 1);
}

namespace {
// Orders the host -> device uploads of all contexts of one device (see rvb_crop_track_batch).
struct H2dGate {
	static constexpr int kMaxDev = 64;
	static std::mutex mu[kMaxDev];
	static cudaEvent_t ev[kMaxDev];
	int dev;
	cudaStream_t st;
	int rc = RVB_OK;
	bool locked = false;
	H2dGate(int device, cudaStream_t stream) : dev(device), st(stream) {
		if (dev < 0 || dev >= kMaxDev) { dev = -1; return; }
		mu[dev].lock();
		locked = true;
		if (ev[dev] == nullptr) {
			if (cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming) != cudaSuccess) { rc = fail(RVB_ERR_CUDA, "cudaEventCreate (upload gate)"); return; }
		} else if (cudaStreamWaitEvent(st, ev[dev], 0) != cudaSuccess) {
			rc = fail(RVB_ERR_CUDA, "cudaStreamWaitEvent (upload gate)");
		}
	}
	int done() {
		int r = RVB_OK;
		if (dev >= 0 && locked) {
			if (cudaEventRecord(ev[dev], st) != cudaSuccess) r = fail(RVB_ERR_CUDA, "cudaEventRecord (upload gate)");
			mu[dev].unlock();
			locked = false;
		}
		return r;
	}
	~H2dGate() { if (locked) mu[dev].unlock(); }
};
std::mutex H2dGate::mu[H2dGate::kMaxDev];
cudaEvent_t H2dGate::ev[H2dGate::kMaxDev] = {};

struct Staging {  // lays the small per-call arrays out in one block: the host arrays are copied straight into pinned memory
	struct Item { size_t off; const void *src; size_t n; };
	std::vector<Item> items;
	size_t size = 0;
	// src == nullptr: device scratch of n bytes (not written by write())
	size_t add(const void *src, size_t n) {
		const size_t off = (size + 255) / 256 * 256;
		size = off + n;
		if (src && n) items.push_back({off, src, n});
		return off;
	}
	// the sources must still be alive; `zero_first`: the first `bytes` bytes start as zeros (gaps and scratch included)
	void write(void *dst, size_t bytes, bool zero_first) const {
		if (zero_first) memset(dst, 0, bytes);
		for (const Item &it : items)
			if (it.off + it.n <= bytes) memcpy((uint8_t *)dst + it.off, it.src, it.n);
	}
};
}  // namespace

extern "C" int rvb_crop_track_batch(rvb_ctx *c, const rvb_params *p, const rvb_batch *b) {
	if (!c || !p || !b) return fail(RVB_ERR_INVALID, "NULL argument");
	if (p->resize_factor != 1.0) {
		if (!(p->resize_factor > 1.0))
			return fail(RVB_ERR_UNSUPPORTED, "resize_factor=%g (up-scaling before the clustering is not built)", p->resize_factor);
		if (p->resize_type < 1 || p->resize_type > 3)
			return fail(RVB_ERR_UNSUPPORTED, "resize_type=%d: bilinear (1), cubic (2) and nearest (3) are built", p->resize_type);
	}
	if (b->n_clips <= 0) return fail(RVB_ERR_INVALID, "n_clips=%d", b->n_clips);
	if (b->n_ratios < 1 || b->n_ratios > RVB_MAX_RATIOS) return fail(RVB_ERR_INVALID, "n_ratios=%d", b->n_ratios);
	const int H = b->h_process, W = b->w_process;
	if (H < 1 || W < 1 || H > 256 || W > 256) return fail(RVB_ERR_UNSUPPORTED, "process size %dx%d (max 256x256)", H, W);
	if (p->hdbscan_min < 2) return fail(RVB_ERR_INVALID, "hdbscan_min=%d (min_cluster_size must be >= 2)", p->hdbscan_min);
	if (p->loess_degree < 1 || p->loess_degree > 2) return fail(RVB_ERR_UNSUPPORTED, "loess_degree=%d", p->loess_degree);
	if (p->lp_filt && p->lp_order > RVB_MAX_LP_ORDER)
		return fail(RVB_ERR_UNSUPPORTED, "lp_order=%d (scipy designs any order; this build stops at %d)", p->lp_order, RVB_MAX_LP_ORDER);
	if (!b->clips || !b->shots || !b->true_inds || (!b->maps && !b->clip_maps) || !b->boxes) return fail(RVB_ERR_INVALID, "NULL array in batch");
	if (p->t_border != -1 && b->maps_kind == RVB_MAPS_F32_NHW)
		return fail(RVB_ERR_UNSUPPORTED, "border detection on float32 maps is not built");
	CU(cudaSetDevice(c->device));
	const int WPS = align_up(W, 16);
	const int R = b->n_ratios;
	const bool host = b->mem_space == RVB_MEM_HOST;
	cudaStream_t st = c->stream;

	// host time by section: 0 tables, 1 packing + upload, 2 map input, 3 map pipeline launches, 4 track launches, 5 results
	auto t_host = std::chrono::steady_clock::now();
	auto lap = [&](int k) {
		const auto now = std::chrono::steady_clock::now();
		c->host_us[k] = std::chrono::duration<double, std::micro>(now - t_host).count();
		t_host = now;
	};
	// ---- metadata --------------------------------------------------------------------------------
	const int nc = b->n_clips;
	long long tot_maps = 0, tot_frames = 0, tot_shots = 0;
	for (int i = 0; i < nc; ++i) {
		const rvb_clip &cl = b->clips[i];
		if (cl.n_maps < 1 || cl.n_frames < 1 || cl.n_shots < 1) return fail(RVB_ERR_INVALID, "clip %d is empty", i);
		if (cl.map_offset != tot_maps || cl.frame_offset != tot_frames || cl.shot_offset != tot_shots)
			return fail(RVB_ERR_INVALID, "clip %d: offsets must be the running sums (packed arrays)", i);
		tot_maps += cl.n_maps; tot_frames += cl.n_frames; tot_shots += cl.n_shots;
	}
	if (tot_maps > 0x7fffffffLL / 64 || tot_frames > 0x3fffffffLL) return fail(RVB_ERR_INVALID, "batch too large");
	const int NM = (int)tot_maps, NF = (int)tot_frames, NS = (int)tot_shots;

	std::vector<ClipDev> clips(nc);
	std::vector<ShotDev> shots(NS);
	std::vector<int> pred(NM, -1), store(NM, -1);
	std::vector<uint8_t> chain_next(NM, 0);
	std::vector<int> clip_final(nc * R * 3), cvrg_cfg(nc * R * 2), clip_coef(nc);
	std::vector<FilterCoef> coefs;
	std::vector<double> coef_fr;
	long long scratch_doubles = 0;
	long long max_spline_doubles = 0, max_lp_doubles = 0;
	int n_slots = 0;
	const bool want_filtered = b->filtered_maps != nullptr;
	const bool keep_all_maps = want_filtered || p->focus_stability;  // focus stability samples every filtered map
	for (int i = 0; i < nc; ++i) {
		const rvb_clip &cl = b->clips[i];
		ClipDev &d = clips[i];
		d.n_maps = cl.n_maps; d.n_frames = cl.n_frames; d.n_shots = cl.n_shots;
		d.h_orig = cl.h_orig; d.w_orig = cl.w_orig; d.fr = cl.fr;
		d.map_offset = (int)cl.map_offset; d.frame_offset = (int)cl.frame_offset; d.shot_offset = (int)cl.shot_offset;
		std::vector<char> is_cut(cl.n_maps + 2, 0);
		for (int s = 0; s < cl.n_shots; ++s) {
			const int32_t *row = b->shots + (cl.shot_offset + s) * 4;
			ShotDev &sh = shots[d.shot_offset + s];
			sh.f0 = row[0]; sh.f1 = row[1]; sh.m0 = row[2]; sh.m1 = row[3]; sh.clip = i;
			if (sh.f0 < 0 || sh.f1 >= cl.n_frames || sh.f1 < sh.f0 || sh.m0 < 0 || sh.m1 >= cl.n_maps || sh.m1 < sh.m0)
				return fail(RVB_ERR_INVALID, "clip %d shot %d: bad range", i, s);
			if (s == 0 ? (sh.f0 != 0) : (sh.f0 != shots[d.shot_offset + s - 1].f1 + 1))
				return fail(RVB_ERR_INVALID, "clip %d shot %d: shots must tile the frames", i, s);
			sh.frame_base = d.frame_offset + sh.f0;
			sh.map_base = d.map_offset + sh.m0;
			const int n = sh.m1 - sh.m0 + 1, clen = sh.f1 - sh.f0 + 1;
			const long long need = std::max<long long>(8LL * n + 3, 2LL * (clen + 6 * (RVB_MAX_LP_ORDER + 1)));
			max_spline_doubles = std::max<long long>(max_spline_doubles, n > 6 ? 8LL * n + 3 : 0);
			max_lp_doubles = std::max<long long>(max_lp_doubles, clen + 6LL * (RVB_MAX_LP_ORDER + 1));
			sh.scratch_base = (int)scratch_doubles;
			scratch_doubles += need;
			is_cut[sh.m0] = 1;                                     // segm_cuts: shot starts ...
			if (s == cl.n_shots - 1) is_cut[sh.m1] = 1;            // ... + the last end (smartVidCrop.py:2324-2327)
		}
		if (shots[d.shot_offset + cl.n_shots - 1].f1 != cl.n_frames - 1)
			return fail(RVB_ERR_INVALID, "clip %d: last shot must end at the last frame", i);
		// cut-adjacent blend (smartVidCrop.py:2369-2373): map k+1 <- (map k+1 + filtered map k) / 2
		// (when every filtered map is kept -- returned to the caller, or sampled by focus stability -- the slots are the map
		// indices, so the maps come back with one strided copy)
		if (keep_all_maps)
			for (int k = 0; k < cl.n_maps; ++k) store[d.map_offset + k] = d.map_offset + k;
		if (p->clust_filt) {
			for (int k = 0; k < cl.n_maps - 2; ++k) {
				const bool hit = (k >= 1 && is_cut[k - 1]) || is_cut[k] || is_cut[k + 1];
				if (hit) {
					const int src = d.map_offset + k, dst = src + 1;
					if (store[src] < 0) store[src] = n_slots++;
					pred[dst] = store[src];
					chain_next[src] = 1;
				}
			}
		}
		for (int r = 0; r < R; ++r) {
			int fin[3];
			calc_dest_size(cl.w_orig, cl.h_orig, b->ratio_w[r], b->ratio_h[r], fin);
			memcpy(&clip_final[(i * R + r) * 3], fin, sizeof(fin));
			int win;
			if (fin[0] == 1) win = p->cvrg_window ? (int)((double)fin[1] * W / cl.w_orig) : W;
			else win = p->cvrg_window ? (int)((double)fin[2] * H / cl.h_orig) : H;
			cvrg_cfg[(i * R + r) * 2 + 0] = fin[0];
			cvrg_cfg[(i * R + r) * 2 + 1] = win;
		}
		// one Butterworth design per distinct frame rate
		int ci = -1;
		for (size_t k = 0; k < coef_fr.size(); ++k) if (coef_fr[k] == cl.fr) ci = (int)k;
		if (ci < 0) {
			// (designs are kept per context: the conditioning probe inside butter_design costs ~0.2 ms)
			const double wn = p->lp_cutoff / (0.5 * cl.fr);
			const FilterCoef *hit = nullptr;
			for (const auto &e : c->filter_cache)
				if (e.order == p->lp_order && e.wn == wn) hit = &e.fc;
			if (!hit) {
				if (c->filter_cache.size() >= 64) c->filter_cache.clear();
				rvb_ctx::CachedFilter e;
				e.order = p->lp_order; e.wn = wn;
				butter_design(p->lp_order, wn, e.fc);
				c->filter_cache.push_back(e);
				hit = &c->filter_cache.back().fc;
			}
			const FilterCoef fc = *hit;
			coefs.push_back(fc);
			coef_fr.push_back(cl.fr);
			ci = (int)coefs.size() - 1;
		}
		clip_coef[i] = ci;
	}
	if (keep_all_maps) n_slots = NM;
	if (scratch_doubles > 0x7fffffffLL) return fail(RVB_ERR_INVALID, "batch too large (scratch)");
	// Monolithic path: every map that does not wait for a predecessor is a work item; a chain is walked by the CTA that
	// took its first map (chain_next), so there are no waves and no inter-CTA waits.
	// Split pipeline (front -> prim_kernel -> back), the default: see the work sets below.
	ResizeSetup rzs;
	if (p->resize_factor != 1.0) build_resize(p->resize_factor, H, W, rzs);
	const bool split = c->split && p->clust_filt && !rzs.on;
	// Work sets of the split pipeline: set 0 = every map outside a chain; set 1 + L = the chain maps at depth L
	// (depth 0 = chain heads).  A set runs front -> Prim -> back; set 1 + L needs the filtered maps of set L.
	constexpr int kMaxChainDepth = 8;
	std::vector<int> depth(NM, 0);
	int max_depth = 0;
	for (int m = 0; m < NM; ++m) {
		depth[m] = (pred[m] < 0) ? 0 : depth[m - 1] + 1;   // the predecessor of map m is map m - 1
		max_depth = std::max(max_depth, depth[m]);
	}
	// (with the all-pairs Prim the chains stay in the monolithic kernel unless RVB_CHAIN_LEVELS=1: measured 21.2 vs 17.8 ms per step)
	const bool split_chains = split && max_depth < kMaxChainDepth && c->chain_levels && (!c->dense_prim || c->chain_levels_dense);
	std::vector<int> work;                    // monolithic launches: chain heads (the CTA walks the chain)
	std::vector<std::vector<int>> sets;       // split pipeline
	if (split) sets.resize(split_chains ? 2 + max_depth : 1);
	work.reserve(NM);
	for (int m = 0; m < NM; ++m) {
		const bool in_chain = pred[m] >= 0 || chain_next[m];
		if (!split) { if (pred[m] < 0) work.push_back(m); }
		else if (!in_chain) sets[0].push_back(m);
		else if (split_chains) sets[1 + depth[m]].push_back(m);
		else if (pred[m] < 0) work.push_back(m);
	}
	if (!split) {
		// chain heads first: the longest sequential dependencies start as early as possible
		std::stable_partition(work.begin(), work.end(), [&](int m) { return chain_next[m] != 0; });
	}
	const int n_sets = (int)sets.size();
	// counters: [0..9] per monolithic capacity class {head, len} of that class's work list; then per split set
	// kSetCounters ints: {front head, front len, class lengths[5], Prim heads[5], back heads[5]}; then the scratch
	// allocator (64 bit)
	constexpr int kSetCounters = 2 + 3 * kSplitClasses + 3;   // ... + {wide-front head, len} + pad
	const int cnt_scr = (10 + n_sets * kSetCounters + 1) & ~1;
	std::vector<int> counters(cnt_scr + 2, 0);
	counters[1] = (int)work.size();
	std::vector<int> set_list_off(n_sets, 0);   // offsets (ints) into the concatenated front lists
	std::vector<int> split_all;
	int n_split = 0;
	for (int k = 0; k < n_sets; ++k) {
		counters[10 + k * kSetCounters + 1] = (int)sets[k].size();
		set_list_off[k] = n_split;
		n_split += (int)sets[k].size();
		split_all.insert(split_all.end(), sets[k].begin(), sets[k].end());
	}

	lap(0);
	Staging sg;
	// host tables first (uploaded), then device scratch (cleared on the device, never sent over the link)
	const size_t o_clips = sg.add(clips.data(), clips.size() * sizeof(ClipDev));
	const size_t o_shots = sg.add(shots.data(), shots.size() * sizeof(ShotDev));
	const size_t o_ti = sg.add(b->true_inds, (size_t)NM * sizeof(int));
	const size_t o_pred = sg.add(pred.data(), (size_t)NM * sizeof(int));
	const size_t o_store = sg.add(store.data(), (size_t)NM * sizeof(int));
	const size_t o_chain = sg.add(chain_next.data(), (size_t)NM);
	const size_t o_final = sg.add(clip_final.data(), clip_final.size() * sizeof(int));
	const size_t o_cvrg = sg.add(cvrg_cfg.data(), cvrg_cfg.size() * sizeof(int));
	const size_t o_ccoef = sg.add(clip_coef.data(), clip_coef.size() * sizeof(int));
	const size_t o_coefs = sg.add(coefs.data(), coefs.size() * sizeof(FilterCoef));
	const size_t o_cnt = sg.add(counters.data(), counters.size() * sizeof(int));
	const size_t o_rz = sg.add(rzs.tab.data(), rzs.tab.size() * sizeof(int16_t));
	const int small_bytes = rzs.on ? rzs.Hs * rzs.WSs : 0;
	const size_t o_work = sg.add(work.data(), work.size() * sizeof(int));
	const size_t o_work2 = sg.add(split_all.data(), split_all.size() * sizeof(int));
	const size_t data_bytes = (sg.size + 255) / 256 * 256;
	const size_t o_fshot = sg.add(nullptr, (size_t)NF * sizeof(int));    // frame -> shot, frame -> clip, map -> clip:
	const size_t o_fclip = sg.add(nullptr, (size_t)NF * sizeof(int));    // written by index_tables_kernel
	const size_t o_mclip = sg.add(nullptr, (size_t)NM * sizeof(int));
	const size_t o_jumps = sg.add(nullptr, (size_t)NM * sizeof(double));
	const size_t o_ovf1 = sg.add(nullptr, (size_t)NM * sizeof(int));
	const size_t o_ovf2 = sg.add(nullptr, (size_t)NM * sizeof(int));
	const size_t o_ovf3 = sg.add(nullptr, (size_t)NM * sizeof(int));
	const size_t o_ovf4 = sg.add(nullptr, (size_t)NM * sizeof(int));
	const size_t o_ovfw = sg.add(nullptr, (size_t)NM * sizeof(int));   // maps the 4096-point front hands to the wide front
	const size_t o_cls = sg.add(nullptr, (size_t)kSplitClasses * n_split * sizeof(int));
	const size_t o_scroff = sg.add(nullptr, (size_t)(split ? NM : 0) * sizeof(int));
	const size_t o_zero0 = (sg.size + 255) / 256 * 256;                  // ---- scratch that must start as zeros
	const size_t o_skip = sg.add(nullptr, (size_t)(split ? NM : 0));
	const size_t o_borders = sg.add(nullptr, (size_t)nc * 4 * sizeof(int));
	const size_t o_status = sg.add(nullptr, (size_t)nc * sizeof(int));
	const size_t o_prof = sg.add(nullptr, (size_t)nc * (H + W) * sizeof(uint32_t));
	const size_t o_dims = sg.add(nullptr, (size_t)nc * R * 9 * sizeof(int));
	const size_t o_cscore = sg.add(nullptr, (size_t)nc * (1 + R) * sizeof(double));
	const size_t o_mscore = sg.add(nullptr, (size_t)NM * sizeof(double));
	const size_t o_minfo = sg.add(nullptr, (size_t)NM * 4 * sizeof(int));
	const size_t o_empty = sg.add(nullptr, (size_t)NM);
	const size_t o_minmax = sg.add(nullptr, (size_t)NS * 4 * sizeof(double));
	const size_t o_nf = sg.add(nullptr, (size_t)((b->centres_nf && host) ? 2 * NM : 0) * sizeof(double));
	const size_t meta_bytes = sg.size;
	if (c->stage_busy) { CU(cudaEventSynchronize(c->ev_stage)); c->stage_busy = false; }
	if (c->stage.ensure(data_bytes)) return RVB_ERR_CUDA;
	if (c->meta.ensure(meta_bytes)) return RVB_ERR_CUDA;
	sg.write(c->stage.p, data_bytes, false);
	CU(cudaMemcpyAsync(c->meta.p, c->stage.p, data_bytes, cudaMemcpyHostToDevice, st));
	CU(cudaEventRecord(c->ev_stage, st));
	c->stage_busy = true;
	CU(cudaMemsetAsync((uint8_t *)c->meta.p + o_zero0, 0, meta_bytes - o_zero0, st));
	index_tables_kernel<<<NS + nc, 128, 0, st>>>((const ShotDev *)((uint8_t *)c->meta.p + o_shots), NS, (const ClipDev *)((uint8_t *)c->meta.p + o_clips), nc,
												  (int *)((uint8_t *)c->meta.p + o_fshot), (int *)((uint8_t *)c->meta.p + o_fclip), (int *)((uint8_t *)c->meta.p + o_mclip));
	c->launches += 1;
	uint8_t *M = (uint8_t *)c->meta.p;
	const ClipDev *d_clips = (const ClipDev *)(M + o_clips);
	const ShotDev *d_shots = (const ShotDev *)(M + o_shots);
	const int *d_ti = (const int *)(M + o_ti);
	const int *d_fshot = (const int *)(M + o_fshot);
	const int *d_fclip = (const int *)(M + o_fclip);
	int *d_cnt = (int *)(M + o_cnt);
	int *d_borders = (int *)(M + o_borders);
	int *d_status = (int *)(M + o_status);
	uint32_t *d_prof = (uint32_t *)(M + o_prof);

	lap(1);
	// ---- maps ------------------------------------------------------------------------------------
	const uint8_t *d_u8 = nullptr;
	const float *d_f32 = nullptr;
	int gstride = WPS;
	size_t bytes_per_map;
	if (b->maps_kind == RVB_MAPS_F32_NHW) bytes_per_map = (size_t)H * W * sizeof(float);
	else if (b->maps_kind == RVB_MAPS_U8_NHW) {
		if (b->row_stride < W || (b->row_stride % 4) != 0) return fail(RVB_ERR_INVALID, "row_stride=%d", b->row_stride);
		gstride = b->row_stride;
		bytes_per_map = (size_t)H * gstride;
	} else if (b->maps_kind == RVB_MAPS_U8_HWN) bytes_per_map = (size_t)H * W;
	else return fail(RVB_ERR_INVALID, "maps_kind=%d", b->maps_kind);
	const void *d_in = b->maps;
	if (host || b->clip_maps) {
		// gather into one device block (H2D from host memory, or D2D from per-clip device blocks)
		if (c->maps_in.ensure((size_t)NM * bytes_per_map)) return RVB_ERR_CUDA;
		const cudaMemcpyKind k = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
		// Host buffers: the big uploads of the contexts of one device go over the link one after the other.  Interleaved
		// they all finish late and every context's kernels start late; in order, the first context computes while the
		// second uploads (measured: 5 steps over 3 contexts, 26.2 -> see profiles/README.md).
		H2dGate gate(host ? c->device : -1, st);
		if (gate.rc) return gate.rc;
		if (b->clip_maps) {
			for (int i = 0; i < nc; ++i) {
				if (!b->clip_maps[i]) return fail(RVB_ERR_INVALID, "clip_maps[%d] is NULL", i);
				CU(cudaMemcpyAsync((uint8_t *)c->maps_in.p + (size_t)clips[i].map_offset * bytes_per_map, b->clip_maps[i],
								   (size_t)clips[i].n_maps * bytes_per_map, k, st));
			}
		} else {
			CU(cudaMemcpyAsync(c->maps_in.p, b->maps, (size_t)NM * bytes_per_map, k, st));
		}
		d_in = c->maps_in.p;
		if (gate.done()) return RVB_ERR_CUDA;
	}
	if (b->maps_kind == RVB_MAPS_F32_NHW) {
		d_f32 = (const float *)d_in;
		if (((uintptr_t)d_f32 & 15) != 0) return fail(RVB_ERR_INVALID, "maps must be 16-byte aligned");
	} else if (b->maps_kind == RVB_MAPS_U8_NHW) {
		d_u8 = (const uint8_t *)d_in;
		if (((uintptr_t)d_u8 & 15) != 0) return fail(RVB_ERR_INVALID, "maps must be 16-byte aligned");
	} else {
		const uint8_t *src = (const uint8_t *)d_in;
		if (c->maps_nhw.ensure((size_t)NM * H * WPS)) return RVB_ERR_CUDA;
		// the padding columns x >= W must read 0; the transposition never writes them, so they are cleared only when
		// the buffer is new or the geometry changed
		if (c->nhw_zero_p != c->maps_nhw.p || c->nhw_zero_cap != c->maps_nhw.cap || c->nhw_zero_w != W || c->nhw_zero_wps != WPS || c->nhw_zero_h != H) {
			CU(cudaMemsetAsync(c->maps_nhw.p, 0, c->maps_nhw.cap, st));
			c->nhw_zero_p = c->maps_nhw.p; c->nhw_zero_cap = c->maps_nhw.cap; c->nhw_zero_w = W; c->nhw_zero_wps = WPS; c->nhw_zero_h = H;
		}
		int max_maps = 1;
		for (int i = 0; i < nc; ++i) max_maps = std::max(max_maps, clips[i].n_maps);
		for (int z0 = 0; z0 < nc; z0 += 65535) {
			const int nz = std::min(nc - z0, 65535);
			dim3 grid(((max_maps + kTrTN - 1) / kTrTN) * ((W + kTrTX - 1) / kTrTX), H, nz);
			transpose_hwn_kernel<<<grid, 256, 0, st>>>(src, (size_t)NM * H * W, d_clips + z0, H, W, (uint8_t *)c->maps_nhw.p, WPS);
			c->launches += 1;
		}
		CU(cudaGetLastError());
		d_u8 = (const uint8_t *)c->maps_nhw.p;
		gstride = WPS;
	}

	lap(2);
	// ---- border profiles ---------------------------------------------------------------------------
	if (p->t_border != -1) {
		border_profile_kernel<<<NM, 128, 0, st>>>(d_u8, H, W, gstride, (const int *)(M + o_mclip), d_prof);
		CU(cudaGetLastError());
		c->launches += 1;
	}
	border_finish_kernel<<<(nc + 127) / 128, 128, 0, st>>>(d_clips, nc, d_prof, H, W, p->t_border, d_borders);
	CU(cudaGetLastError());
	c->launches += 1;

	// ---- the fused map kernel, wave by wave, three capacity classes each ---------------------------
	if (c->mapout.ensure((size_t)NM * sizeof(MapOut))) return RVB_ERR_CUDA;
	if (n_slots > 0 && c->filt.ensure((size_t)n_slots * H * WPS)) return RVB_ERR_CUDA;
	MapArgs a;
	memset(&a, 0, sizeof(a));
	a.force_class = -1;
	a.maps_u8 = d_u8; a.maps_f32 = d_f32; a.H = H; a.W = W; a.WPS = WPS; a.gstride = gstride;
	a.pred = (const int *)(M + o_pred); a.store = (const int *)(M + o_store); a.map_clip = (const int *)(M + o_mclip);
	a.chain_next = (const uint8_t *)(M + o_chain);
	a.filt = (uint8_t *)c->filt.p; a.fstride = WPS; a.out = (MapOut *)c->mapout.p;
	a.border_prof = nullptr;
	a.cvrg_cfg = p->exit_on_low_cvrg ? (const int *)(M + o_cvrg) : nullptr;
	a.n_ratios = R; a.labels_dbg = nullptr;
	a.resize_on = rzs.on; a.resize_type = (p->resize_type == 1 && p->resize_factor == 2.0) ? 4 : p->resize_type; a.Hs = rzs.Hs; a.Ws = rzs.Ws; a.WSs = rzs.WSs; a.factor = p->resize_factor;
	a.rz = (const int16_t *)(M + o_rz); a.rz_dx = rzs.dx; a.rz_dy = rzs.dy; a.rz_ux = rzs.ux; a.rz_uy = rzs.uy; a.rz_nx = rzs.nx; a.rz_ny = rzs.ny; a.rz_cx = rzs.cx; a.rz_cy = rzs.cy;
	a.phase_cycles = c->phase_on ? (unsigned long long *)c->phase.p : nullptr;
	a.t_threshold = p->t_threshold; a.clust_filt = p->clust_filt; a.mcs = p->hdbscan_min;
	a.min_samples = p->hdbscan_min_samples; a.select_sum = p->select_sum; a.op_close = p->op_close; a.com_km = p->com_km;
	c->map_launches = 0;
	c->stages_timed = false;
	CU(cudaEventRecord(c->ev_map0, st));
	const bool streaming = !p->clust_filt && !rzs.on && !p->exit_on_low_cvrg && !keep_all_maps && d_u8 != nullptr &&
						   (gstride % 16) == 0 && (((uintptr_t)d_u8) & 15) == 0;
	if (streaming) {
		// clustering switched off: one streaming pass, one CTA per map (persistent grid, whole waves)
		int per_sm = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, map_stream_kernel<false>, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
		const int blocks = std::min(NM, c->n_sm * per_sm);
		if (p->com_km) map_stream_kernel<false><<<blocks, 256, 0, st>>>(d_u8, NM, H, W, gstride, p->t_threshold, (MapOut *)c->mapout.p);
		else map_stream_kernel<true><<<blocks, 256, 0, st>>>(d_u8, NM, H, W, gstride, p->t_threshold, (MapOut *)c->mapout.p);
		CU(cudaGetLastError());
		c->launches += 1;
		c->map_launches += 1;
	} else {
	// failed / never-reached maps must not look valid
		CU(cudaMemsetAsync(c->mapout.p, 0xFF, (size_t)NM * sizeof(MapOut), st));
		{
			const int nw = (int)work.size();
			int *cnt = d_cnt;
			int *ovf[4] = {(int *)(M + o_ovf1), (int *)(M + o_ovf2), (int *)(M + o_ovf3), (int *)(M + o_ovf4)};
			const int mcs = p->hdbscan_min;
			// capacity classes 1536 / 2048 / 3072 / 4096 / 8192 salient pixels of the monolithic kernel (shared memory
			// per CTA grows with the capacity, so smaller classes keep more maps in flight per SM): a map that does not
			// fit is appended to the next class's list by the kernel itself (no host round trip)
			auto launch_mono = [&](int k_first, cudaStream_t stream) -> int {
				// the largest class whose shared-memory layout fits this process size (a large down-scaled copy, or a process
				// size above 140 x 250, can push the 8192-point class over the per-CTA limit)
				static const int kCaps[5] = {1536, 2048, 3072, 4096, 8192};
				int k_last = -1;
				for (int k = 0; k < 5; ++k)
					if (make_layout(kCaps[k], H, WPS, W, mcs, small_bytes).total <= (kCaps[k] > 4096 ? kMaxDynSmem : 200 * 1024)) k_last = k;
				if (k_last < k_first) {
					if (k_first > 0) return RVB_OK;      // (only a fallback for overflow was asked for)
					return fail(RVB_ERR_UNSUPPORTED, "process size %dx%d does not fit the shared memory of any capacity class", H, W);
				}
				for (int k = k_first; k <= k_last; ++k) {
					a.list = (k == 0) ? (const int *)(M + o_work) : ovf[k - 1];
					a.head = cnt + 2 * k;
					a.list_len = cnt + 2 * k + 1;
					// anything larger than the last class is flagged RVB_ERR_CAPACITY by the kernel
					a.ovf_list = (k < k_last) ? ovf[k] : nullptr;
					a.ovf_len = (k < k_last) ? cnt + 2 * k + 3 : nullptr;
					int rc = RVB_OK;
					const int nwk = (k == 0) ? nw : nw + n_split;   // later classes also receive overflow
					if (k == 0) rc = launch_map<256, 6>(c, a, H, W, WPS, occupancy_grid<256, 6>(c, make_layout(1536, H, WPS, W, mcs, small_bytes).total, nwk), stream);
					if (k == 1) rc = launch_map<256, 8>(c, a, H, W, WPS, occupancy_grid<256, 8>(c, make_layout(2048, H, WPS, W, mcs, small_bytes).total, nwk), stream);
					if (k == 2) rc = launch_map<512, 6>(c, a, H, W, WPS, occupancy_grid<512, 6>(c, make_layout(3072, H, WPS, W, mcs, small_bytes).total, nwk), stream);
					if (k == 3) rc = launch_map<512, 8>(c, a, H, W, WPS, occupancy_grid<512, 8>(c, make_layout(4096, H, WPS, W, mcs, small_bytes).total, nwk), stream);
					if (k == 4) rc = launch_map<512, 16>(c, a, H, W, WPS, occupancy_grid<512, 16>(c, make_layout(8192, H, WPS, W, mcs, small_bytes).total, nwk), stream);
					if (rc) return rc;
				}
				return RVB_OK;
			};
			if (!split) {
				int rc = launch_mono(0, st);
				if (rc) return rc;
			} else {
				// scratch for the points of the split maps: 2048 per map on average; a map that finds it full goes to a
				// monolithic launch instead
				const size_t cap = (size_t)std::min<long long>((long long)n_split * 2048 + 8192, 0x7fffff00LL);
				if (c->scr_pinfo.ensure(cap * sizeof(uint2)) || c->scr_val.ensure(cap) || c->scr_pkey.ensure(cap * sizeof(uint32_t))) return RVB_ERR_CUDA;
				if (!c->dense_prim && (c->scr_far.ensure(cap * sizeof(uint32_t)) || c->scr_alist.ensure(cap * sizeof(uint32_t)))) return RVB_ERR_CUDA;
				a.scr_off = (int *)(M + o_scroff); a.scr_top = (unsigned long long *)(cnt + cnt_scr); a.scr_cap = (unsigned int)cap;
				a.scr_pinfo = (uint2 *)c->scr_pinfo.p; a.scr_val = (uint8_t *)c->scr_val.p; a.scr_pkey = (uint32_t *)c->scr_pkey.p;
				a.skip = M + o_skip;
				PrimArgs pa;
				memset(&pa, 0, sizeof(pa));
				pa.out = (const MapOut *)c->mapout.p; pa.scr_off = a.scr_off; pa.scr_pinfo = a.scr_pinfo; pa.scr_pkey = a.scr_pkey;
				pa.phase_cycles = a.phase_cycles;
				FPrimArgs fa;
				memset(&fa, 0, sizeof(fa));
				fa.out = pa.out; fa.scr_off = pa.scr_off; fa.scr_pinfo = pa.scr_pinfo; fa.scr_pkey = pa.scr_pkey;
				fa.scr_far = (uint32_t *)c->scr_far.p; fa.scr_alist = (uint32_t *)c->scr_alist.p;
				fa.phase_cycles = a.phase_cycles;
				fa.work = a.phase_cycles ? a.phase_cycles + 11 : nullptr;
				// front: load .. core distances.  Its overflow (more than 4096 points, or scratch full) lands in the list
				// of the monolithic class of 8192 points, which also walks the rest of that map's chain.
				auto launch_front = [&](int set, cudaStream_t stream) -> int {
					const int ns = (int)sets[set].size();
					if (ns == 0) return RVB_OK;
					int *sc = cnt + 10 + set * kSetCounters;
					a.cls_lists = (int *)(M + o_cls) + (size_t)kSplitClasses * set_list_off[set]; a.cls_cnt = sc + 2; a.cls_stride = ns;
					a.force_class = -1;
					int rc = RVB_OK;
					if (set == 0) {
						// the bulk of the maps: 4096-point front CTAs (4 per SM); a larger map is handed to the wide front below
						a.list = (const int *)(M + o_work2) + set_list_off[set]; a.head = sc; a.list_len = sc + 1;
						const bool wide = !c->dense_prim;   // (the all-pairs prim_kernel stops at 4096 points)
						int *ovfw = (int *)(M + o_ovfw);
						a.ovf_list = wide ? ovfw : ovf[3]; a.ovf_len = wide ? sc + 2 + 3 * kSplitClasses + 1 : cnt + 2 * 4 + 1;
						rc = launch_map<256, 16, kModeFront>(c, a, H, W, WPS, occupancy_grid<256, 16, kModeFront>(c, make_layout(4096, H, WPS, W, mcs, 0, true).total, ns), stream);
						if (rc || !wide) return rc;
						a.list = ovfw; a.head = sc + 2 + 3 * kSplitClasses; a.list_len = sc + 2 + 3 * kSplitClasses + 1;
					} else {
						// a chain set: few maps, made larger by the blend
						a.list = (const int *)(M + o_work2) + set_list_off[set]; a.head = sc; a.list_len = sc + 1;
						if (c->dense_prim) {
							// the all-pairs prim_kernel stops at 4096 points: larger maps (and the rest of their chains) go to the
							// monolithic class of 8192 points at the end of the side stream
							a.ovf_list = ovf[3]; a.ovf_len = cnt + 2 * 4 + 1;
							return launch_map<256, 16, kModeFront>(c, a, H, W, WPS, occupancy_grid<256, 16, kModeFront>(c, make_layout(4096, H, WPS, W, mcs, 0, true).total, ns), stream);
						}
					}
					// wide front (8192 points).  What it cannot place (more points than that -> flagged by the monolithic
					// launch, or scratch exhausted) goes to the monolithic class of 8192 points, which also walks the rest
					// of that map's chain.
					a.ovf_list = ovf[3]; a.ovf_len = cnt + 2 * 4 + 1;
					return launch_map<512, 16, kModeFront>(c, a, H, W, WPS, occupancy_grid<512, 16, kModeFront>(c, make_layout(8192, H, WPS, W, mcs, 0, true).total, set == 0 ? std::min(ns, c->n_sm) : ns), stream);
				};
				// Prim and back, one launch each per size class
				// Prim and back of one size class: two launches on `stream`
				auto launch_class = [&](int set, int k, cudaStream_t stream) -> int {
					const int ns = (int)sets[set].size();
					int *sc = cnt + 10 + set * kSetCounters;
					int *lists = (int *)(M + o_cls) + (size_t)kSplitClasses * set_list_off[set];
					pa.list = lists + (size_t)k * ns; pa.list_len = sc + 2 + k; pa.head = sc + 2 + kSplitClasses + k;
					pa.cap = split_class_cap(k);
					const int smem = 12 * pa.cap;
					// RVB_DENSE_PRIM=1, the round-1 all-pairs kernel: (warps per map, register slots per thread), 16 warps
					// per SM in every class
					const int v = c->dense_prim ? (k < 4 ? c->prim_variant[k] : 0) : -1;
					if (v >= 0 && k >= 5) return RVB_OK;   // (no map reaches the 8192 class in that mode)
					if (v < 0) {
						fa.list = pa.list; fa.list_len = pa.list_len; fa.head = pa.head;
						if (k == 0) launch_fprim<768>(c, fa, ns, H, W, stream);
						if (k == 1) launch_fprim<1536>(c, fa, ns, H, W, stream);
						if (k == 2) launch_fprim<2048>(c, fa, ns, H, W, stream);
						if (k == 3) launch_fprim<3072>(c, fa, ns, H, W, stream);
						if (k == 4) launch_fprim<4096>(c, fa, ns, H, W, stream);
						if (k == 5) launch_fprim<8192>(c, fa, ns, H, W, stream);
					}
					if (k == 0 && v == 0) launch_prim<1, 24>(c, pa, ns, smem, stream);
					if (k == 0 && v == 1) launch_prim<2, 12>(c, pa, ns, smem, stream);
					if (k == 1 && v == 0) launch_prim<2, 24>(c, pa, ns, smem, stream);
					if (k == 1 && v == 1) launch_prim<4, 12>(c, pa, ns, smem, stream);
					if (k == 2 && v == 0) launch_prim<4, 24>(c, pa, ns, smem, stream);
					if (k == 2 && v == 1) launch_prim<4, 16>(c, pa, ns, smem, stream);
					if (k == 3 && v == 0) launch_prim<4, 24>(c, pa, ns, smem, stream);
					if (k == 3 && v == 1) launch_prim<8, 12>(c, pa, ns, smem, stream);
					if (k >= 4 && v >= 0) launch_prim<8, 16>(c, pa, ns, smem, stream);
					CU(cudaGetLastError());
					c->launches += 1;
					c->map_launches += 1;
					a.ovf_list = nullptr; a.ovf_len = nullptr;
					a.list = lists + (size_t)k * ns; a.list_len = sc + 2 + k; a.head = sc + 2 + 2 * kSplitClasses + k;
					int rc = RVB_OK;
					if (k <= 1) rc = launch_map<256, 6, kModeBack>(c, a, H, W, WPS, occupancy_grid<256, 6, kModeBack>(c, make_layout(1536, H, WPS, W, mcs, 0).total, ns), stream);
					if (k == 2) rc = launch_map<256, 8, kModeBack>(c, a, H, W, WPS, occupancy_grid<256, 8, kModeBack>(c, make_layout(2048, H, WPS, W, mcs, 0).total, ns), stream);
					if (k == 3) rc = launch_map<512, 6, kModeBack>(c, a, H, W, WPS, occupancy_grid<512, 6, kModeBack>(c, make_layout(3072, H, WPS, W, mcs, 0).total, ns), stream);
					if (k == 4) rc = launch_map<512, 8, kModeBack>(c, a, H, W, WPS, occupancy_grid<512, 8, kModeBack>(c, make_layout(4096, H, WPS, W, mcs, 0).total, ns), stream);
					if (k == 5) rc = launch_map<512, 16, kModeBack>(c, a, H, W, WPS, occupancy_grid<512, 16, kModeBack>(c, make_layout(8192, H, WPS, W, mcs, 0).total, ns), stream);
					return rc;
				};
				// Prim and back of every size class of a set.  The classes are independent of each other, so with
				// `fan` each runs on its own stream (forked from / joined into `stream`): a Prim launch is a latency
				// chain per map with few warps, a back launch is barrier bound -- side by side they fill each other's
				// idle issue slots, and the tail of one class (its largest maps) overlaps the other classes.
				auto launch_prim_back = [&](int set, cudaStream_t stream, bool fan) -> int {
					const int ns = (int)sets[set].size();
					if (ns == 0) return RVB_OK;
					if (!fan) {
						for (int k = 0; k < kSplitClasses; ++k) {
							const int rc = launch_class(set, k, stream);
							if (rc) return rc;
						}
						return RVB_OK;
					}
					CU(cudaEventRecord(c->ev_cls_fork, stream));
					for (int k = kSplitClasses - 1; k >= 0; --k) {     // largest maps first: the longest latency chains
						CU(cudaStreamWaitEvent(c->cls_stream[k], c->ev_cls_fork, 0));
						const int rc = launch_class(set, k, c->cls_stream[k]);
						if (rc) return rc;
						CU(cudaEventRecord(c->ev_cls[k], c->cls_stream[k]));
						CU(cudaStreamWaitEvent(stream, c->ev_cls[k], 0));
					}
					return RVB_OK;
				};
				// main stream: the maps outside chains.  Side stream: the chains, one depth after the other (few maps,
				// long dependencies), then whatever overflowed into the monolithic kernel.
				int rc = launch_front(0, st);
				if (rc) return rc;
				CU(cudaEventRecord(c->ev_st[0], st));
				cudaStream_t side = getenv("RVB_NO_SIDE") ? st : c->side_stream;
				CU(cudaEventRecord(c->ev_fork, st));
				CU(cudaStreamWaitEvent(side, c->ev_fork, 0));
				if ((rc = launch_prim_back(0, st, !c->serial_classes))) return rc;
				CU(cudaEventRecord(c->ev_st[1], st));      // (Prim and back of the size classes run interleaved: one stage)
				CU(cudaEventRecord(c->ev_st[2], st));
				c->stages_timed = !sets[0].empty();
				for (int k = 1; k < n_sets; ++k) {
					if ((rc = launch_front(k, side))) return rc;
					if ((rc = launch_prim_back(k, side, !c->serial_classes))) return rc;
				}
				if ((rc = launch_mono(split_chains ? 4 : 0, side))) return rc;
				CU(cudaEventRecord(c->ev_join, side));
				CU(cudaStreamWaitEvent(st, c->ev_join, 0));
			}
		}
	}
	CU(cudaEventRecord(c->ev_map1, st));
	c->map_timed = true;

	lap(3);
	// ---- centre track ------------------------------------------------------------------------------
	// series: dx, dy [NM]; dxi, dyi, dxl, dyl, dxs, dys [NF]
	if (c->series.ensure(((size_t)2 * NM + (size_t)6 * NF) * sizeof(double))) return RVB_ERR_CUDA;
	double *d_dx = (double *)c->series.p, *d_dy = d_dx + NM;
	double *d_ser = d_dy + NM;  // [6][NF]
	double *d_dxi = d_ser, *d_dyi = d_ser + NF, *d_dxl = d_ser + 2 * (size_t)NF, *d_dyl = d_ser + 3 * (size_t)NF;
	double *d_dxs = d_ser + 4 * (size_t)NF, *d_dys = d_ser + 5 * (size_t)NF;
	if (c->scratch.ensure((size_t)std::max<long long>(scratch_doubles, 1) * sizeof(double))) return RVB_ERR_CUDA;
	double *d_scratch = (double *)c->scratch.p;
	if (c->boxes.ensure((size_t)R * NF * 4 * sizeof(int32_t))) return RVB_ERR_CUDA;
	int32_t *d_boxes = host ? (int32_t *)c->boxes.p : b->boxes;
	const MapOut *d_mo = (const MapOut *)c->mapout.p;
	uint8_t *d_empty = M + o_empty;

	fill_centres_kernel<<<(nc * 32 + 127) / 128, 128, 0, st>>>(d_clips, nc, d_shots, d_mo, d_dx, d_dy, d_empty, d_status);
	double *d_jumps = (double *)(M + o_jumps);
	if (b->centres_nf) {  // dxnf, dynf: the centres before focus stability (smartVidCrop.py:2450-2451)
		// (host buffers: kept on the device until the results are copied out, so that the call does not block here)
		CU(cudaMemcpyAsync(host ? (void *)(M + o_nf) : (void *)b->centres_nf, d_dx, (size_t)2 * NM * sizeof(double), cudaMemcpyDeviceToDevice, st));
	}
	if (p->focus_stability) {
		focus_jumps_kernel<<<(NM + 127) / 128, 128, 0, st>>>(d_clips, (const int *)(M + o_mclip), NM, d_dx, d_dy, (const uint8_t *)c->filt.p,
															 (const int *)(M + o_store), H, W, WPS, (double)p->min_d_jump, p->np_int_compat, d_jumps);
		focus_apply_kernel<<<(nc + 63) / 64, 64, 0, st>>>(d_clips, nc, d_jumps, p->foces_stab_t, p->foces_stab_s, p->skip, d_dx, d_dy);
		c->launches += 2;
	}
	if (b->centres_nf && !host && p->focus_stability)
		CU(cudaMemcpyAsync(b->centres_nf + 2 * (size_t)NM, d_jumps, (size_t)NM * sizeof(double), cudaMemcpyDeviceToDevice, st));
	const int sp_doubles = (int)std::min<long long>(max_spline_doubles, 6144);   // up to 48 KB of shared memory per warp
	spline_setup_kernel<<<NS, 32, (size_t)sp_doubles * sizeof(double), st>>>(d_shots, NS, d_ti, d_dx, d_dy, d_scratch, sp_doubles);
	interp_eval_kernel<<<(NF + 255) / 256, 256, 0, st>>>(d_shots, d_fshot, NF, d_ti, d_dx, d_dy, d_scratch, d_dxi, d_dyi);
	const int lp_doubles = (int)std::min<long long>(max_lp_doubles, 6144);
	double *d_minmax = (double *)(M + o_minmax);
	lowpass_kernel<<<2 * NS, 32, (size_t)lp_doubles * sizeof(double), st>>>(d_shots, NS, d_clips, (const FilterCoef *)(M + o_coefs),
													   (const int *)(M + o_ccoef), d_dxi, d_dyi, d_dxl, d_dyl, d_scratch, p->lp_filt, lp_doubles, d_minmax,
													   p->loess_filt, p->loess_w_secs, p->loess_degree);
	{
		smooth_kernel<<<dim3((NF + 127) / 128, 2), 128, 0, st>>>(d_shots, d_fshot, NF, d_clips, d_dxl, d_dyl, d_dxs, d_dys, p->loess_filt,
											  p->loess_w_secs, p->loess_degree, d_minmax, d_scratch);
	}
	int32_t *d_dims = (int32_t *)(M + o_dims);
	boxes_kernel<<<(int)(((long long)NF * R + 255) / 256), 256, 0, st>>>(d_clips, d_fclip, NF, R, (const int *)(M + o_final),
																		d_borders, H, W, d_dxs, d_dys, p->shift_time, d_boxes, d_dims);
	double *d_cscore = (double *)(M + o_cscore), *d_mscore = (double *)(M + o_mscore);
	clip_scores_kernel<<<(nc * 32 + 127) / 128, 128, 0, st>>>(d_clips, nc, d_mo, H, W, R, p->exit_on_low_cvrg, p->np_int_compat, d_mscore, d_cscore);
	CU(cudaGetLastError());
	c->launches += 7;

	lap(4);
	// ---- results -----------------------------------------------------------------------------------
	// Host buffers: a copy into pageable memory is staged by the driver and blocks the calling thread, once per array.
	// So every output whose destination is not page-locked goes through one pinned block of the context (asynchronous
	// copies, one synchronisation at the end, then plain memcpys); page-locked destinations are written directly.
	struct Pending { void *dst; size_t off, n; size_t rows, w, dpitch; };
	std::vector<Pending> pend;
	size_t stage_bytes = 0;
	auto pinned = [&](const void *ptr) -> bool {
		cudaPointerAttributes at;
		if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
		return at.type == cudaMemoryTypeHost;
	};
	struct Out { void *dst; const void *src; size_t n; size_t spitch, w, rows, dpitch; };   // rows == 0: flat copy of n bytes
	std::vector<Out> outs;
	const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
	if (host) outs.push_back({b->boxes, d_boxes, (size_t)R * NF * 4 * sizeof(int32_t), 0, 0, 0, 0});
	if (b->centres) outs.push_back({b->centres, d_dx, (size_t)2 * NM * sizeof(double), 0, 0, 0, 0});
	if (b->centres_nf && host) {
		outs.push_back({b->centres_nf, M + o_nf, (size_t)2 * NM * sizeof(double), 0, 0, 0, 0});
		if (p->focus_stability) outs.push_back({b->centres_nf + 2 * (size_t)NM, d_jumps, (size_t)NM * sizeof(double), 0, 0, 0, 0});
	}
	if (b->empty) outs.push_back({b->empty, d_empty, (size_t)NM, 0, 0, 0, 0});
	if (b->series) outs.push_back({b->series, d_ser, (size_t)6 * NF * sizeof(double), 0, 0, 0, 0});
	if (b->map_scores) outs.push_back({b->map_scores, d_mscore, (size_t)NM * sizeof(double), 0, 0, 0, 0});
	if (b->clip_scores) outs.push_back({b->clip_scores, d_cscore, (size_t)nc * (1 + R) * sizeof(double), 0, 0, 0, 0});
	if (b->clip_dims) outs.push_back({b->clip_dims, d_dims, (size_t)nc * R * 9 * sizeof(int), 0, 0, 0, 0});
	if (b->clip_status) outs.push_back({b->clip_status, d_status, (size_t)nc * sizeof(int), 0, 0, 0, 0});
	// n_points, n_clusters, kept_points, flags are 4 consecutive ints inside MapOut
	if (b->map_info) outs.push_back({b->map_info, (const uint8_t *)d_mo + offsetof(MapOut, n_points), 0, sizeof(MapOut), 4 * sizeof(int), (size_t)NM, 4 * sizeof(int)});
	if (want_filtered) {
		const int so = b->row_stride_out > 0 ? b->row_stride_out : WPS;
		if (so < W) return fail(RVB_ERR_INVALID, "row_stride_out=%d", so);
		if (b->filtered_layout == RVB_FILTERED_HWN) {
			// the reference's own layout, per clip [H][W][n_maps] packed back to back: transposed on the device
			uint8_t *d_hwn = b->filtered_maps;
			if (host) {
				if (c->filt_hwn.ensure((size_t)NM * H * W)) return RVB_ERR_CUDA;
				d_hwn = (uint8_t *)c->filt_hwn.p;
			}
			int max_maps = 1;
			for (int i = 0; i < nc; ++i) max_maps = std::max(max_maps, clips[i].n_maps);
			for (int z0 = 0; z0 < nc; z0 += 65535) {
				dim3 grid((max_maps + kTrMaps - 1) / kTrMaps, H, std::min(nc - z0, 65535));
				transpose_to_hwn_kernel<<<grid, 256, 0, st>>>((const uint8_t *)c->filt.p, WPS, d_clips + z0, H, W, d_hwn);
				c->launches += 1;
			}
			CU(cudaGetLastError());
			if (host) outs.push_back({b->filtered_maps, d_hwn, (size_t)NM * H * W, 0, 0, 0, 0});
		} else if (b->filtered_layout != RVB_FILTERED_NHW) {
			return fail(RVB_ERR_INVALID, "filtered_layout=%d", b->filtered_layout);
		} else {
			// slot m holds map m: all rows of all maps in one strided copy
			outs.push_back({b->filtered_maps, c->filt.p, 0, (size_t)WPS, (size_t)W, (size_t)NM * H, (size_t)so});
		}
	}
	if (host) {
		for (const Out &o : outs) {
			if (pinned(o.dst)) continue;
			stage_bytes = (stage_bytes + 255) / 256 * 256;
			stage_bytes += o.rows ? o.rows * o.w : o.n;
		}
		if (stage_bytes && c->stage_out.ensure(stage_bytes)) return RVB_ERR_CUDA;
	}
	{
		size_t off = 0;
		for (const Out &o : outs) {
			const bool direct = !host || pinned(o.dst);
			uint8_t *dst = (uint8_t *)o.dst;
			size_t dpitch = o.dpitch;
			if (!direct) {
				off = (off + 255) / 256 * 256;
				dst = (uint8_t *)c->stage_out.p + off;
				dpitch = o.w;
				pend.push_back({o.dst, off, o.n, o.rows, o.w, o.dpitch});
				off += o.rows ? o.rows * o.w : o.n;
			}
			if (o.rows) CU(cudaMemcpy2DAsync(dst, dpitch, o.src, o.spitch, o.w, o.rows, kind, st));
			else CU(cudaMemcpyAsync(dst, o.src, o.n, kind, st));
		}
	}
	if (host) {
		std::vector<int> status(nc);
		CU(cudaMemcpyAsync(status.data(), d_status, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		for (const Pending &q : pend) {
			const uint8_t *src = (const uint8_t *)c->stage_out.p + q.off;
			if (q.rows == 0) memcpy(q.dst, src, q.n);
			else if (q.dpitch == q.w) memcpy(q.dst, src, q.rows * q.w);
			else for (size_t r = 0; r < q.rows; ++r) memcpy((uint8_t *)q.dst + r * q.dpitch, src + r * q.w, q.w);
		}
		for (int i = 0; i < nc; ++i)
			if (status[i] != RVB_OK)
				return fail(status[i], "clip %d: %s", i,
							status[i] == RVB_ERR_CAPACITY ? "a map has more salient pixels than RVB_MAX_POINTS"
														  : "no map with a salient pixel (the reference raises TypeError in interp_handler)");
	}
	lap(5);
	return RVB_OK;
}

// ---------------------------------------------------------------------------------------------
// IoU
// ---------------------------------------------------------------------------------------------
extern "C" int rvb_iou_batch_run(rvb_ctx *c, const rvb_iou_batch *b) {
	if (!c || !b) return fail(RVB_ERR_INVALID, "NULL argument");
	if (b->n_videos < 1 || b->n_users < 1 || b->n_users > 64) return fail(RVB_ERR_INVALID, "n_videos=%d n_users=%d", b->n_videos, b->n_users);
	if (!b->frame_offset || (!b->n_eval && !b->n_eval_user) || !b->method_boxes || !b->annot_boxes || !b->acc) return fail(RVB_ERR_INVALID, "NULL array");
	CU(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const bool host = b->mem_space == RVB_MEM_HOST;
	const int V = b->n_videos, U = b->n_users;
	const long long NF = b->frame_offset[V];
	if (NF < 1 || b->frame_offset[0] != 0) return fail(RVB_ERR_INVALID, "frame_offset must start at 0");
	if (NF > 0x7fffffffLL) return fail(RVB_ERR_INVALID, "more than 2^31 frames in one call");
	std::vector<int> first(V + 1), neval((size_t)V * U);
	long long max_frames = 1;
	for (int v = 0; v < V; ++v) {
		const long long f0 = b->frame_offset[v], f1 = b->frame_offset[v + 1];
		if (f1 < f0) return fail(RVB_ERR_INVALID, "frame_offset not monotone");
		for (int u = 0; u < U; ++u) {
			const int ne = b->n_eval_user ? b->n_eval_user[(size_t)v * U + u] : b->n_eval[v];
			if (ne < 1 || ne > f1 - f0) return fail(RVB_ERR_INVALID, "video %d: n_eval=%d of %lld frames", v, ne, f1 - f0);
			neval[(size_t)v * U + u] = ne;
		}
		first[v] = (int)f0;
		max_frames = std::max(max_frames, f1 - f0);
	}
	Staging sg;
	first[V] = (int)NF;
	// work items: every video padded to whole 32-frame warp slots; a chunk = 8 slots; per chunk the video of its first slot,
	// that video's first slot, first frame and frame count
	std::vector<int4> chunk_tab;
	{
		std::vector<long long> pad_first((size_t)V + 1);
		long long slots = 0;
		for (int v = 0; v < V; ++v) {
			pad_first[v] = slots;
			slots += (b->frame_offset[v + 1] - b->frame_offset[v] + 31) / 32;
		}
		if (slots < 1) return fail(RVB_ERR_INVALID, "no frames");
		pad_first[V] = slots;
		chunk_tab.resize((size_t)((slots + 7) / 8));
		int v = 0;
		for (size_t ci = 0; ci < chunk_tab.size(); ++ci) {
			while (v + 1 < V && (long long)ci * 8 >= pad_first[v + 1]) ++v;
			chunk_tab[ci] = make_int4(v, (int)pad_first[v], (int)b->frame_offset[v], (int)(b->frame_offset[v + 1] - b->frame_offset[v]));
		}
	}
	const size_t o_chunks = sg.add(chunk_tab.data(), chunk_tab.size() * sizeof(int4));
	const size_t o_first = sg.add(first.data(), (size_t)(V + 1) * sizeof(int));
	const size_t o_ne = sg.add(neval.data(), (size_t)V * U * sizeof(int));
	const size_t o_acc = sg.add(nullptr, (size_t)V * U * 2 * sizeof(uint64_t));
	const size_t o_acc3 = sg.add(nullptr, (size_t)V * U * 4 * sizeof(uint64_t));
	const size_t o_bad = sg.add(nullptr, sizeof(int));
	if (c->stage_busy) { CU(cudaEventSynchronize(c->ev_stage)); c->stage_busy = false; }
	if (c->stage.ensure(sg.size) || c->iou_a.ensure(sg.size)) return RVB_ERR_CUDA;
	sg.write(c->stage.p, o_acc, false);          // (the tables; the accumulators behind them are device scratch)
	CU(cudaMemcpyAsync(c->iou_a.p, c->stage.p, o_acc, cudaMemcpyHostToDevice, st));
	CU(cudaEventRecord(c->ev_stage, st));
	c->stage_busy = true;
	uint8_t *M = (uint8_t *)c->iou_a.p;
	const int32_t *d_method = b->method_boxes, *d_annot = b->annot_boxes;
	double *d_fiou = b->frame_iou;
	const size_t mb = (size_t)NF * 4 * sizeof(int32_t);
	if (host) {
		if (c->iou_b.ensure(mb * (1 + U))) return RVB_ERR_CUDA;
		CU(cudaMemcpyAsync(c->iou_b.p, b->method_boxes, mb, cudaMemcpyHostToDevice, st));
		CU(cudaMemcpyAsync((uint8_t *)c->iou_b.p + mb, b->annot_boxes, mb * U, cudaMemcpyHostToDevice, st));
		d_method = (const int32_t *)c->iou_b.p;
		d_annot = (const int32_t *)((uint8_t *)c->iou_b.p + mb);
		if (b->frame_iou) {
			if (c->iou_c.ensure((size_t)NF * U * sizeof(double))) return RVB_ERR_CUDA;
			d_fiou = (double *)c->iou_c.p;
		}
	}
	unsigned long long *d_acc = (unsigned long long *)(M + o_acc);
	unsigned long long *d_acc3 = (unsigned long long *)(M + o_acc3);
	CU(cudaMemsetAsync(d_acc3, 0, (size_t)V * U * 4 * sizeof(uint64_t), st));
	int *d_bad = (int *)(M + o_bad);
	CU(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
	CU(cudaEventRecord(c->ev_iou0, st));
	const int n_chunks = (int)chunk_tab.size();
	{
		const int *d_first = (const int *)(M + o_first), *d_ne = (const int *)(M + o_ne);
		const int4 *d_chunks = (const int4 *)(M + o_chunks);
		static const bool generic = getenv("RVB_IOU_GENERIC") != nullptr;
		// persistent grid: exactly the CTAs that are resident at once
#define RVB_IOU_LAUNCH(UT) { static int occ = 0; if (!occ) { CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, iou_kernel<UT, false>, 256, 0)); occ = std::max(occ, 1); } \
		if (d_fiou) iou_kernel<UT, true><<<std::min(n_chunks, c->n_sm * occ), 256, 0, st>>>(d_method, d_annot, d_first, V, d_ne, d_chunks, n_chunks, NF, U, d_fiou, d_acc3, d_bad); \
		else iou_kernel<UT, false><<<std::min(n_chunks, c->n_sm * occ), 256, 0, st>>>(d_method, d_annot, d_first, V, d_ne, d_chunks, n_chunks, NF, U, d_fiou, d_acc3, d_bad); }
		switch (generic ? 0 : U) {
			case 1: RVB_IOU_LAUNCH(1) break;
			case 2: RVB_IOU_LAUNCH(2) break;
			case 3: RVB_IOU_LAUNCH(3) break;
			case 4: RVB_IOU_LAUNCH(4) break;
			case 5: RVB_IOU_LAUNCH(5) break;
			case 6: RVB_IOU_LAUNCH(6) break;      // the RetargetVid annotations: 6 annotators
			case 7: RVB_IOU_LAUNCH(7) break;
			case 8: RVB_IOU_LAUNCH(8) break;
			default: RVB_IOU_LAUNCH(0) break;
		}
#undef RVB_IOU_LAUNCH
	}
	iou_finish_kernel<<<(V * U + 255) / 256, 256, 0, st>>>(d_acc3, V * U, d_acc);
	CU(cudaGetLastError());
	CU(cudaEventRecord(c->ev_iou1, st));
	c->iou_timed = true;
	c->launches += 1;
	c->launches += 1;
	const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
	CU(cudaMemcpyAsync(b->acc, d_acc, (size_t)V * U * 2 * sizeof(uint64_t), kind, st));
	if (b->n_bad) CU(cudaMemcpyAsync(b->n_bad, d_bad, sizeof(int), kind, st));
	if (host) {
		if (b->frame_iou) CU(cudaMemcpyAsync(b->frame_iou, d_fiou, (size_t)NF * U * sizeof(double), kind, st));
		int n_bad = 0;
		CU(cudaMemcpyAsync(&n_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		// malformed boxes (x2 < x1 or an empty union): the reference averages a negative value or raises
		// ZeroDivisionError (retargetvid_eval.py:24-26); neither fits the exact accumulator, so the call says so
		if (n_bad > 0) return fail(RVB_ERR_INVALID, "%d IoU(s) outside [0, 1]: a box with x2 < x1 / y2 < y1 or an empty union", n_bad);
	}
	return RVB_OK;
}

// exactly rounded (round-half-even) acc / (n * 2^80) as a double
extern "C" double rvb_iou_mean_from_acc(const uint64_t acc[2], int64_t n) {
	if (n <= 0) return nan("");
	unsigned __int128 S = ((unsigned __int128)acc[1] << 64) | acc[0];
	if (S == 0) return 0.0;
	const unsigned __int128 d = (unsigned __int128)(uint64_t)n;
	// quotient with 64 extra fractional bits: q = qi + qf / 2^64, sticky = remainder != 0
	const unsigned __int128 qi = S / d;
	const unsigned __int128 r = S % d;
	const unsigned __int128 num2 = r << 64;  // r < n < 2^63
	const unsigned __int128 qf = num2 / d;
	const bool sticky = (num2 % d) != 0;
	// value = (qi * 2^64 + qf [+ sticky]) * 2^-(64 + kIouFracBits); qi < 2^112 so work in 192 bits: hi = qi, lo = qf
	// find the top bit
	int top;  // index of the leading 1 in the 192-bit number (bit 0 = LSB of qf)
	if (qi != 0) {
		int hb = 127;
		while (!((qi >> hb) & 1)) --hb;
		top = 64 + hb;
	} else {
		int hb = 63;
		while (!(((uint64_t)qf >> hb) & 1)) --hb;
		top = hb;
	}
	auto bit = [&](int i) -> int {
		if (i < 0) return 0;
		if (i < 64) return (int)(((uint64_t)qf >> i) & 1);
		return (int)((qi >> (i - 64)) & 1);
	};
	uint64_t mant = 0;  // 53 bits: top .. top-52
	for (int i = 0; i < 53; ++i) mant = (mant << 1) | (uint64_t)bit(top - i);
	const int guard = bit(top - 53);
	bool rest = sticky;
	for (int i = top - 54; i >= 0 && !rest; --i) rest = bit(i) != 0;
	if (guard && (rest || (mant & 1))) mant += 1;  // may carry to 2^53: ldexp handles it
	return ldexp((double)mant, top - 52 - 64 - kIouFracBits);
}

// ---------------------------------------------------------------------------------------------
// renderer crop
// ---------------------------------------------------------------------------------------------
// host-only: the filter the low-pass stage would run for (order, Wn)
extern "C" int rvb_debug_butter(int32_t order, double wn, double *b, double *a, double *zi, int32_t *chunked_ok) {
	if (!b || !a || !zi || !chunked_ok) return fail(RVB_ERR_INVALID, "NULL argument");
	FilterCoef fc;
	butter_design(order, wn, fc);
	if (fc.order == 0) return fail(RVB_ERR_UNSUPPORTED, "order=%d Wn=%g: no filter (scipy raises; the reference falls back to moving averages)", order, wn);
	for (int i = 0; i <= order; ++i) { b[i] = fc.b[i]; a[i] = fc.a[i]; }
	for (int i = 0; i < order; ++i) zi[i] = fc.zi[i];
	*chunked_ok = fc.chunked_ok;
	return RVB_OK;
}

// ---------------------------------------------------------------------------------------------
// a16: result text format (host)
// ---------------------------------------------------------------------------------------------
extern "C" int rvb_format_boxes_txt(const int32_t *boxes, int64_t n_frames, char *out, int64_t cap, int64_t *len_out) {
	if (!boxes || !out || !len_out || n_frames < 0) return fail(RVB_ERR_INVALID, "NULL argument");
	if (cap < n_frames * 48) return fail(RVB_ERR_CAPACITY, "rvb_format_boxes_txt: %lld bytes for %lld frames (48 per frame)", (long long)cap, (long long)n_frames);
	char *q = out;
	for (int64_t f = 0; f < n_frames; ++f) {
		for (int k = 0; k < 4; ++k) {
			long long v = boxes[f * 4 + k];
			if (v < 0) { *q++ = '-'; v = -v; }
			char tmp[12];
			int nd = 0;
			do { tmp[nd++] = (char)('0' + v % 10); v /= 10; } while (v);
			while (nd) *q++ = tmp[--nd];
			*q++ = (k < 3) ? ',' : '\n';
		}
	}
	*len_out = q - out;
	return RVB_OK;
}

extern "C" int rvb_parse_boxes_txt(const char *text, int64_t len, int32_t *boxes, int64_t cap_frames, int64_t *n_frames_out) {
	if (!text || !n_frames_out || len < 0 || (cap_frames > 0 && !boxes)) return fail(RVB_ERR_INVALID, "NULL argument");
	auto is_space = [](char ch) { return ch == ' ' || ch == '\t' || ch == '\v' || ch == '\f'; };
	int64_t n = 0, pos = 0;
	bool overflow = false;
	while (pos < len) {
		// one line: up to '\n', '\r' or '\r\n' (str.splitlines)
		int64_t e = pos;
		while (e < len && text[e] != '\n' && text[e] != '\r') ++e;
		int64_t p = pos;
		int32_t v4[4];
		for (int k = 0; k < 4; ++k) {
			// int(field): optional blanks, optional sign, digits, optional blanks
			while (p < e && is_space(text[p])) ++p;
			bool neg = false;
			if (p < e && (text[p] == '+' || text[p] == '-')) { neg = text[p] == '-'; ++p; }
			long long v = 0;
			int nd = 0;
			while (p < e && text[p] >= '0' && text[p] <= '9') {
				v = v * 10 + (text[p] - '0');
				if (v > 0x7fffffffLL + 1) return fail(RVB_ERR_INVALID, "line %lld: a coordinate does not fit 32 bits", (long long)n + 1);
				++p; ++nd;
			}
			while (p < e && is_space(text[p])) ++p;
			const bool field_end = (p == e) || text[p] == ',';
			if (nd == 0 || !field_end) return fail(RVB_ERR_INVALID, "line %lld: field %d is not an integer", (long long)n + 1, k);
			if (k < 3) {
				if (p == e) return fail(RVB_ERR_INVALID, "line %lld: %d field(s), 4 needed", (long long)n + 1, k + 1);
				++p;   // the comma
			}
			v = neg ? -v : v;
			if (v > 0x7fffffffLL || v < -0x7fffffffLL - 1) return fail(RVB_ERR_INVALID, "line %lld: a coordinate does not fit 32 bits", (long long)n + 1);
			v4[k] = (int32_t)v;
		}
		// (further fields are ignored, as c[4:] is)
		if (n < cap_frames) memcpy(boxes + n * 4, v4, sizeof(v4)); else overflow = true;
		++n;
		pos = e;
		if (pos < len) pos += (text[pos] == '\r' && pos + 1 < len && text[pos + 1] == '\n') ? 2 : 1;
	}
	*n_frames_out = n;
	if (overflow) return fail(RVB_ERR_CAPACITY, "rvb_parse_boxes_txt: %lld lines, room for %lld", (long long)n, (long long)cap_frames);
	return RVB_OK;
}

extern "C" int rvb_crop_frames(rvb_ctx *c, const uint8_t *frames, int32_t n_frames, int32_t h, int32_t w, int32_t channels,
							   const int32_t *boxes, int32_t out_h, int32_t out_w, uint8_t *out, int32_t mem_space) {
	if (!c || !frames || !boxes || !out) return fail(RVB_ERR_INVALID, "NULL argument");
	if (n_frames < 1 || h < 1 || w < 1 || channels < 1 || channels > 4 || out_h < 1 || out_w < 1 || out_h > h || out_w > w)
		return fail(RVB_ERR_INVALID, "bad geometry: %d frames of %dx%dx%d -> %dx%d", n_frames, h, w, channels, out_h, out_w);
	for (int f = 0; f < n_frames; ++f) {
		const int32_t *b = boxes + (size_t)f * 4;
		if (b[2] - b[0] != out_w || b[3] - b[1] != out_h || b[0] < 0 || b[1] < 0 || b[2] > w || b[3] > h)
			return fail(RVB_ERR_INVALID, "frame %d: box %d,%d,%d,%d is not a %dx%d window inside %dx%d", f, b[0], b[1], b[2], b[3], out_w, out_h, w, h);
	}
	CU(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const bool host = mem_space == RVB_MEM_HOST;
	const size_t in_bytes = (size_t)n_frames * h * w * channels, out_bytes = (size_t)n_frames * out_h * out_w * channels;
	const size_t box_bytes = (size_t)n_frames * 4 * sizeof(int32_t);
	if (c->stage_busy) { CU(cudaEventSynchronize(c->ev_stage)); c->stage_busy = false; }
	if (c->stage.ensure(box_bytes) || c->misc.ensure(box_bytes + 256)) return RVB_ERR_CUDA;
	memcpy(c->stage.p, boxes, box_bytes);
	CU(cudaMemcpyAsync(c->misc.p, c->stage.p, box_bytes, cudaMemcpyHostToDevice, st));
	CU(cudaEventRecord(c->ev_stage, st));
	c->stage_busy = true;
	const uint8_t *d_in = frames;
	uint8_t *d_out = out;
	if (host) {
		if (c->maps_in.ensure(in_bytes) || c->filt.ensure(out_bytes)) return RVB_ERR_CUDA;
		CU(cudaMemcpyAsync(c->maps_in.p, frames, in_bytes, cudaMemcpyHostToDevice, st));
		d_in = (const uint8_t *)c->maps_in.p;
		d_out = (uint8_t *)c->filt.p;
	}
	const long long row_bytes = (long long)out_w * channels;
	const bool al4 = ((uintptr_t)d_out & 7) == 0;
	const int vec = (al4 && row_bytes % 8 == 0) ? 8 : (al4 && row_bytes % 4 == 0) ? 4 : 1;
	const long long total = (long long)n_frames * out_h * (row_bytes / vec);
	const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)c->n_sm * 32);
	const int32_t *d_boxes = (const int32_t *)c->misc.p;
	if (vec == 8) crop_frames_kernel<8><<<blocks, 256, 0, st>>>(d_in, n_frames, h, w, channels, d_boxes, out_h, out_w, d_out);
	else if (vec == 4) crop_frames_kernel<4><<<blocks, 256, 0, st>>>(d_in, n_frames, h, w, channels, d_boxes, out_h, out_w, d_out);
	else crop_frames_kernel<1><<<blocks, 256, 0, st>>>(d_in, n_frames, h, w, channels, d_boxes, out_h, out_w, d_out);
	CU(cudaGetLastError());
	c->launches += 1;
	if (host) {
		CU(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
	}
	return RVB_OK;
}

// ---------------------------------------------------------------------------------------------
// stage-level entry points for the parity tests
// ---------------------------------------------------------------------------------------------
extern "C" int rvb_debug_cluster_labels(rvb_ctx *c, const rvb_params *p, const uint8_t *map_hw, int32_t h, int32_t w,
										int32_t *labels_out, int32_t *n_points_out) {
	if (!c || !p || !map_hw || !labels_out || !n_points_out) return fail(RVB_ERR_INVALID, "NULL argument");
	if (h < 1 || w < 1 || h > 256 || w > 256) return fail(RVB_ERR_UNSUPPORTED, "size %dx%d", h, w);
	CU(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const int WPS = align_up(w, 16);
	std::vector<uint8_t> padded((size_t)h * WPS, 0);
	int n = 0;
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) {
			padded[(size_t)y * WPS + x] = map_hw[(size_t)y * w + x];
			n += map_hw[(size_t)y * w + x] != 0;
		}
	if (n > RVB_MAX_POINTS) return fail(RVB_ERR_CAPACITY, "%d salient pixels (max %d)", n, RVB_MAX_POINTS);
	*n_points_out = n;
	const size_t o_map = 0, o_out = align_up((int)padded.size(), 256), o_lab = o_out + 256, o_list = o_lab + (size_t)RVB_MAX_POINTS * 4 + 256;
	const size_t o_cnt = o_list + 256, total = o_cnt + 256;
	if (c->misc.ensure(total)) return RVB_ERR_CUDA;
	uint8_t *D = (uint8_t *)c->misc.p;
	CU(cudaMemsetAsync(D, 0xFF, total, st));
	CU(cudaMemcpyAsync(D + o_map, padded.data(), padded.size(), cudaMemcpyHostToDevice, st));
	const int hdr[4] = {0, 1, 0, 0};  // head, len
	const int zero = 0;
	CU(cudaMemcpyAsync(D + o_cnt, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(D + o_list, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
	MapArgs a;
	memset(&a, 0, sizeof(a));
	a.force_class = -1;
	a.maps_u8 = D + o_map; a.H = h; a.W = w; a.WPS = WPS; a.gstride = WPS;
	a.list = (const int *)(D + o_list); a.head = (int *)(D + o_cnt); a.list_len = (int *)(D + o_cnt) + 1;
	a.out = (MapOut *)(D + o_out); a.labels_dbg = (int32_t *)(D + o_lab);
	a.t_threshold = 0; a.clust_filt = 1; a.mcs = p->hdbscan_min; a.min_samples = p->hdbscan_min_samples;
	a.select_sum = p->select_sum; a.op_close = p->op_close; a.com_km = p->com_km;
	int rc;
	if (n <= 1536) rc = launch_map<256, 6>(c, a, h, w, WPS, 1);
	else if (n <= 2048) rc = launch_map<256, 8>(c, a, h, w, WPS, 1);
	else if (n <= 3072) rc = launch_map<512, 6>(c, a, h, w, WPS, 1);
	else if (n <= 4096) rc = launch_map<512, 8>(c, a, h, w, WPS, 1);
	else rc = launch_map<512, 16>(c, a, h, w, WPS, 1);
	if (rc) return rc;
	CU(cudaMemcpyAsync(labels_out, D + o_lab, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
	MapOut mo;
	CU(cudaMemcpyAsync(&mo, D + o_out, sizeof(mo), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	if (mo.flags & (kFlagOverflow | kFlagClusterCapacity)) return fail(RVB_ERR_CAPACITY, "cluster capacity exceeded");
	if (mo.n_clusters < 0) {  // gates skipped the clustering: the reference would not call fit_predict
		for (int i = 0; i < n; ++i) labels_out[i] = -2;
	}
	return RVB_OK;
}

extern "C" int rvb_debug_smooth_series(rvb_ctx *c, const rvb_params *p, const double *series_in, int32_t n, double fr,
									   double *lowpassed_out, double *smoothed_out) {
	if (!c || !p || !series_in || n < 1) return fail(RVB_ERR_INVALID, "bad argument");
	CU(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	ClipDev cl;
	memset(&cl, 0, sizeof(cl));
	cl.n_maps = 1; cl.n_frames = n; cl.n_shots = 1; cl.h_orig = 360; cl.w_orig = 640; cl.fr = fr;
	ShotDev sh;
	memset(&sh, 0, sizeof(sh));
	sh.f0 = 0; sh.f1 = n - 1; sh.m0 = 0; sh.m1 = 0; sh.clip = 0; sh.frame_base = 0; sh.map_base = 0; sh.scratch_base = 0;
	FilterCoef fc;
	butter_design(p->lp_order, p->lp_cutoff / (0.5 * fr), fc);
	std::vector<int> fshot(n, 0);
	const int zero = 0;
	Staging sg;
	const size_t o_cl = sg.add(&cl, sizeof(cl)), o_sh = sg.add(&sh, sizeof(sh)), o_fc = sg.add(&fc, sizeof(fc));
	const size_t o_fs = sg.add(fshot.data(), (size_t)n * sizeof(int)), o_cc = sg.add(&zero, sizeof(int));
	const size_t o_x = sg.add(series_in, (size_t)n * sizeof(double));
	const size_t o_y = sg.add(series_in, (size_t)n * sizeof(double));
	const size_t o_xl = sg.add(nullptr, (size_t)n * sizeof(double)), o_yl = sg.add(nullptr, (size_t)n * sizeof(double));
	const size_t o_xs = sg.add(nullptr, (size_t)n * sizeof(double)), o_ys = sg.add(nullptr, (size_t)n * sizeof(double));
	const size_t o_scr = sg.add(nullptr, (size_t)(2 * (n + 6 * (RVB_MAX_LP_ORDER + 1)) + 16) * sizeof(double));
	const size_t o_mm = sg.add(nullptr, 4 * sizeof(double));
	if (c->misc.ensure(sg.size)) return RVB_ERR_CUDA;
	uint8_t *D = (uint8_t *)c->misc.p;
	std::vector<uint8_t> blob(sg.size);
	sg.write(blob.data(), sg.size, true);
	CU(cudaMemcpyAsync(D, blob.data(), sg.size, cudaMemcpyHostToDevice, st));
	CU(cudaStreamSynchronize(st));
	const int dbg_lp_doubles = std::min(n + 6 * (RVB_MAX_LP_ORDER + 1), 6144);
	lowpass_kernel<<<2, 32, (size_t)dbg_lp_doubles * sizeof(double), st>>>((const ShotDev *)(D + o_sh), 1, (const ClipDev *)(D + o_cl), (const FilterCoef *)(D + o_fc),
									  (const int *)(D + o_cc), (const double *)(D + o_x), (const double *)(D + o_y),
									  (double *)(D + o_xl), (double *)(D + o_yl), (double *)(D + o_scr), p->lp_filt, dbg_lp_doubles, (double *)(D + o_mm),
									  p->loess_filt, p->loess_w_secs, p->loess_degree);
	smooth_kernel<<<dim3((n + 127) / 128, 2), 128, 0, st>>>((const ShotDev *)(D + o_sh), (const int *)(D + o_fs), n, (const ClipDev *)(D + o_cl),
										  (const double *)(D + o_xl), (const double *)(D + o_yl), (double *)(D + o_xs),
										  (double *)(D + o_ys), p->loess_filt, p->loess_w_secs, p->loess_degree, (const double *)(D + o_mm), (const double *)(D + o_scr));
	CU(cudaGetLastError());
	c->launches += 2;
	if (lowpassed_out) CU(cudaMemcpyAsync(lowpassed_out, D + o_xl, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
	if (smoothed_out) CU(cudaMemcpyAsync(smoothed_out, D + o_xs, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return RVB_OK;
}
