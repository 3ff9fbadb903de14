"""ORACLE (test infrastructure).  numpy restatement of the two cv2.resize calls on the hot path
(smartVidCrop.py:1078-1084, 1158, 1184) for uint8 single-channel images: INTER_LINEAR (OpenCV
modules/imgproc/src/resize.cpp: float32 source coordinate, weights rounded to 1/2048, horizontal pass
in int, vertical pass ((b*(S>>4))>>16, +2, >>2) with clipped rows) and INTER_NEAREST.  Checked against
cv2 itself in tests/test_cpu_host.py.  For fx = fy = 1/2 OpenCV switches INTER_LINEAR to INTER_AREA
(resize.cpp: is_area_fast && iscale == 2): resize_area2_u8 restates that path (full 2x2 cells (sum + 2) >> 2, the
partial cells of an odd size saturate_cast<uchar>(float(sum) / count), i.e. round half to even).
INTER_CUBIC (resize_type 2, smartVidCrop.py:1081-1082): resize_cubic_u8 -- interpolateCubic with A = -0.75 in float32,
weights rounded to 1/2048 (short), horizontal pass in int over 4 taps with replicated borders, vertical pass as
VResizeCubicVec_32s8u computes it: float32 S0*b0 + (S1*b1 + (S2*b2 + S3*b3)) with b = beta / 2^22, rounded half to even,
saturated.  That is what cv2 4.13 as installed here (IPP enabled) returns in every column at the factors tested in
tests/test_cpu_host.py (2, 3, 4, 5, 1.5, 2.5); a build without IPP uses the integer formula (sum + 2^21) >> 22 in the last
width % 8 columns, and IPP's own sampling deviates by +-1 on a few per cent of the pixels at factors such as 1.3 / 1.7."""
import numpy as np


def cv_round(v):
	return int(np.rint(v))  # cvRound: round half to even


def _linear_coeffs(dsize, ssize, scale, vertical):
	idx = np.zeros(dsize, dtype=np.int64)
	a0 = np.zeros(dsize, dtype=np.int64)
	a1 = np.zeros(dsize, dtype=np.int64)
	for d in range(dsize):
		f = np.float32((d + 0.5) * scale - 0.5)
		s = int(np.floor(f))
		f = np.float32(f - np.float32(s))
		if not vertical:
			if s < 0:
				s, f = 0, np.float32(0)
			if s >= ssize - 1:
				s, f = ssize - 1, np.float32(0)
		idx[d] = s
		a0[d] = cv_round(np.float32(np.float32(1.0) - f) * np.float32(2048))
		a1[d] = cv_round(f * np.float32(2048))
	return idx, a0, a1


def down_size(n, factor):
	"""dsize of cv2.resize(src, None, fx=1/factor): saturate_cast<int>(n * fx)"""
	return cv_round(n * (1.0 / factor))


def resize_linear_u8(src, dsize_wh=None, fx=None, fy=None):
	H, W = src.shape
	if dsize_wh is None:
		dw, dh = cv_round(W * fx), cv_round(H * fy)
		sx, sy = 1.0 / fx, 1.0 / fy
	else:
		dw, dh = dsize_wh
		sx, sy = 1.0 / (dw / W), 1.0 / (dh / H)
	xi, xa0, xa1 = _linear_coeffs(dw, W, sx, False)
	yi, yb0, yb1 = _linear_coeffs(dh, H, sy, True)
	s = src.astype(np.int64)
	rows = s[:, xi] * xa0[None, :] + s[:, np.minimum(xi + 1, W - 1)] * xa1[None, :]
	y0 = np.clip(yi, 0, H - 1)
	y1 = np.clip(yi + 1, 0, H - 1)
	out = (((yb0[:, None] * (rows[y0, :] >> 4)) >> 16) + ((yb1[:, None] * (rows[y1, :] >> 4)) >> 16) + 2) >> 2
	return out.astype(np.uint8)


def resize_nearest_u8(src, fx, fy):
	H, W = src.shape
	dw, dh = cv_round(W * fx), cv_round(H * fy)
	xs = np.minimum(np.floor(np.arange(dw) * (1.0 / fx)).astype(np.int64), W - 1)
	ys = np.minimum(np.floor(np.arange(dh) * (1.0 / fy)).astype(np.int64), H - 1)
	return src[ys][:, xs]


def resize_area2_u8(src):
	"""cv2.resize(src, None, fx=0.5, fy=0.5, INTER_LINEAR) -> INTER_AREA fast path (resizeAreaFast_, scale 2)."""
	H, W = src.shape
	dw, dh = cv_round(W * 0.5), cv_round(H * 0.5)
	s = src.astype(np.int64)
	out = np.zeros((dh, dw), dtype=np.uint8)
	for y in range(dh):
		for x in range(dw):
			y0, x0 = 2 * y, 2 * x
			if y0 >= H or x0 >= W:
				continue
			blk = s[y0:min(y0 + 2, H), x0:min(x0 + 2, W)]
			if blk.size == 4:
				out[y, x] = (int(blk.sum()) + 2) >> 2
			else:
				out[y, x] = cv_round(np.float32(blk.sum()) / np.float32(blk.size))
	return out


def _cubic_weights(x):
	x = np.float32(x)
	A, one = np.float32(-0.75), np.float32(1)
	c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
	c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
	c2 = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
	c3 = one - c0 - c1 - c2
	return [np.float32(c0), np.float32(c1), np.float32(c2), np.float32(c3)]


def _cubic_coeffs(dsize, scale):
	idx = np.zeros(dsize, dtype=np.int64)
	a = np.zeros((dsize, 4), dtype=np.int64)
	for d in range(dsize):
		f = np.float32((d + 0.5) * scale - 0.5)
		s = int(np.floor(f))
		f = np.float32(f - np.float32(s))
		idx[d] = s
		a[d] = [cv_round(np.float32(c * np.float32(2048))) for c in _cubic_weights(f)]
	return idx, a


def resize_cubic_u8(src, fx, fy):
	"""cv2.resize(src, None, fx=fx, fy=fy, interpolation=cv2.INTER_CUBIC) for uint8 (see the module docstring)."""
	H, W = src.shape
	dw, dh = cv_round(W * fx), cv_round(H * fy)
	xi, xa = _cubic_coeffs(dw, 1.0 / fx)
	yi, yb = _cubic_coeffs(dh, 1.0 / fy)
	s = src.astype(np.int64)
	rows = np.zeros((H, dw), dtype=np.int64)
	for k in range(4):
		rows += s[:, np.clip(xi - 1 + k, 0, W - 1)] * xa[:, k][None, :]
	b = (yb.astype(np.float32) * (np.float32(1.0) / np.float32(2048 * 2048))).astype(np.float32)
	R = [rows[np.clip(yi - 1 + k, 0, H - 1), :].astype(np.float32) for k in range(4)]
	x = (R[3] * b[:, 3][:, None]).astype(np.float32)
	for k in (2, 1, 0):
		x = ((R[k] * b[:, k][:, None]).astype(np.float32) + x).astype(np.float32)
	return np.clip(np.rint(x), 0, 255).astype(np.uint8)
