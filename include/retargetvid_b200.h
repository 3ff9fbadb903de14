/*
 * retargetvid_b200 -- C ABI of the B200-native crop-selection hot path.
 *
 * The reference (bmezaris/RetargetVid) has no FFI: its boundary for this path is
 * the Python function surface of smartVidCrop.py.  Each entry
 * point below names the reference interface it replaces; INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 (RVB_OK) or a negative status and never
 * throws, prints or blocks on stdin; rvb_last_error() returns a thread-local
 * message.  The caller allocates and frees every buffer; the library owns only
 * its per-context workspace.  One context per device, used by one host thread
 * at a time; different contexts are independent (per-video sharding across
 * GPUs relies on that).  All launches go to the context's stream.
 */
#ifndef RETARGETVID_B200_H
#define RETARGETVID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVB_OK                 0
#define RVB_ERR_INVALID       -1  /* bad argument */
#define RVB_ERR_CUDA          -2  /* CUDA runtime error (message has the detail) */
#define RVB_ERR_UNSUPPORTED   -3  /* parameter combination not built yet (e.g. resize_factor != 1) */
#define RVB_ERR_CAPACITY      -4  /* a map had more salient pixels than RVB_MAX_POINTS */
#define RVB_ERR_NO_CENTRES    -5  /* a clip had no non-empty map: the reference raises TypeError here */

#define RVB_MAX_POINTS      8192  /* salient pixels per map the clustering kernel accepts */
#define RVB_MAX_RATIOS         8
#define RVB_MAX_LP_ORDER       8

/* memory space of the bulk arrays of a call */
#define RVB_MEM_HOST           0
#define RVB_MEM_DEVICE         1

/* layout / type of the saliency maps */
#define RVB_MAPS_U8_NHW        0  /* uint8 [sum N][H][row_stride], device-native (row_stride % 16 == 0) */
#define RVB_MAPS_U8_HWN        1  /* uint8 per clip [H][W][N], the reference's vid_data['smaps'] (smartVidCrop.py:286) */
#define RVB_MAPS_F32_NHW       2  /* float32 log-saliency [sum N][H][W], UNISAL output before unisal/train.py:1270-1274 */

/* layout of the optional filtered-map output */
#define RVB_FILTERED_NHW       0  /* uint8 [sum N][H][row_stride_out] */
#define RVB_FILTERED_HWN       1  /* uint8 per clip [H][W][n_maps], packed back to back (the reference's layout) */

typedef struct rvb_ctx rvb_ctx;

/* Mirrors the crop_params dict of sc_init_crop_params (smartVidCrop.py:132-209);
 * only the keys the hot path reads. */
typedef struct rvb_params {
	int32_t t_threshold;          /* sc_threshold, smartVidCrop.py:1057 */
	int32_t clust_filt;           /* smartVidCrop.py:2354 */
	int32_t hdbscan_min;          /* min_cluster_size, smartVidCrop.py:2341 */
	int32_t hdbscan_min_samples;  /* <= 0 means None (= hdbscan_min), smartVidCrop.py:2342 */
	int32_t select_sum;           /* 1: cluster with max sum, else max value, smartVidCrop.py:1110-1113 */
	int32_t op_close;             /* 5x5 closing, smartVidCrop.py:1126-1128 */
	int32_t com_km;               /* 1: centroid (KMeans(1)), 0: argmax, smartVidCrop.py:1165 */
	int32_t t_border;             /* -1 disables border detection, smartVidCrop.py:844 */
	int32_t loess_filt;           /* 1: LOESS, 0: Savitzky-Golay, smartVidCrop.py:1635-1644 */
	int32_t loess_degree;         /* 1 or 2 */
	int32_t lp_filt;              /* Butterworth + filtfilt, smartVidCrop.py:1688 */
	int32_t lp_order;
	int32_t shift_time;           /* smartVidCrop.py:1740-1746 */
	int32_t exit_on_low_cvrg;     /* compute the coverage score, smartVidCrop.py:2380-2383 */
	int32_t cvrg_window;          /* 0: reference (score == 0.0), 1: crop-sized window */
	int32_t resize_type;          /* 1 bilinear, 2 cubic, 3 nearest: cv2.resize as called at smartVidCrop.py:1079-1084 */
	int32_t focus_stability;      /* smartVidCrop.py:2427-2473 */
	int32_t min_d_jump;
	int32_t skip;                 /* frames between saliency maps, used by the focus-stability duration */
	int32_t np_int_compat;        /* the interpreter the reference runs on.  0: today's (numpy >= 1.24: np.int raises, so diagonal
	                                 focus-stability jumps give 255; Python >= 3.12: the builtin sum() behind mean_cvrg_score is
	                                 compensated); 1: the environment the reference pins (np.int exists; plain left-to-right sum) */
	double  foces_stab_t;
	double  foces_stab_s;
	double  loess_w_secs;
	double  lp_cutoff;
	double  resize_factor;        /* >= 1.0; exactly 2.0 with resize_type 1 takes OpenCV's INTER_AREA path like cv2.resize does */
	double  t_cvrg;
} rvb_params;

/* One video.  Offsets index the packed per-map / per-frame / per-shot arrays. */
typedef struct rvb_clip {
	int32_t n_maps;        /* vid_data['fc_sel'] */
	int32_t n_frames;      /* vid_data['fc'] */
	int32_t n_shots;       /* rows of vid_data['segmentation'] */
	int32_t h_orig;
	int32_t w_orig;
	int32_t reserved0;
	double  fr;            /* vid_data['fr'] */
	int64_t map_offset;
	int64_t frame_offset;
	int64_t shot_offset;
} rvb_clip;

/* A batch of clips sharing one process size (H x W; 140 x 250 for 16:9 input). */
typedef struct rvb_batch {
	int32_t n_clips;
	int32_t h_process;            /* vid_data['h_process'] */
	int32_t w_process;            /* vid_data['w_process'] */
	int32_t row_stride;           /* RVB_MAPS_U8_NHW only; bytes per row */
	int32_t maps_kind;            /* RVB_MAPS_* */
	int32_t mem_space;            /* RVB_MEM_*: where maps and the outputs below live */
	int32_t n_ratios;             /* target aspect ratios evaluated from one pass */
	int32_t reserved0;
	double  ratio_w[RVB_MAX_RATIOS];   /* "a:b" -> ratio_w = a, ratio_h = b (smartVidCrop.py:950-952) */
	double  ratio_h[RVB_MAX_RATIOS];
	const rvb_clip *clips;        /* HOST, n_clips entries */
	const int32_t *shots;         /* HOST, [sum S][4] = frame_start, frame_end, map_start, map_end (inclusive) */
	const int32_t *true_inds;     /* HOST, [sum N], vid_data['true_inds'] */
	const void    *maps;          /* mem_space; for U8_HWN clip c starts at byte H*W*map_offset[c] */
	/* ---- outputs, mem_space, caller allocated; optional ones may be NULL ---- */
	int32_t *boxes;               /* [n_ratios][sum F][4] = x1,y1,x2,y2 (smartVidCrop.py:1046) */
	double  *centres;             /* optional [2][sum N]: dx, dy after sc_handle_empty_centers and focus stability */
	double  *centres_nf;          /* optional [3][sum N]: dxnf, dynf (before focus stability), jumps */
	uint8_t *empty;               /* optional [sum N]: 1 where the filtered map was empty (dx is None) */
	double  *series;              /* optional [6][sum F]: dxi, dyi, dxl, dyl, dxs, dys (dxs/dys before truncation) */
	double  *map_scores;          /* optional [sum N]: mean_sal_scores (smartVidCrop.py:1307) */
	double  *clip_scores;         /* optional [n_clips][1 + n_ratios]: mean_sal_score, mean_cvrg_score per ratio */
	int32_t *clip_dims;           /* optional [n_clips][n_ratios][9]: conversion_mode, w_final, h_final, fbb_w, fbb_h, border t,b,l,r */
	uint8_t *filtered_maps;       /* optional, the maps after clustering/closing/blend: uint8 [sum N][H][row_stride_out]
	                                 (RVB_FILTERED_NHW), or per clip [H][W][n_maps] packed like RVB_MAPS_U8_HWN input
	                                 (RVB_FILTERED_HWN: what smart_vid_crop leaves in vid_data['smaps'], smartVidCrop.py:2366-2373) */
	int32_t row_stride_out;       /* RVB_FILTERED_NHW only; 0 = the library's 256-byte-aligned rows */
	int32_t filtered_layout;      /* RVB_FILTERED_* */
	int32_t *map_info;            /* optional [sum N][4]: n_points, n_clusters, kept_points, flags */
	int32_t *clip_status;         /* optional [n_clips]: RVB_OK or RVB_ERR_* per clip */
	const void *const *clip_maps; /* optional HOST array of n_clips pointers (each in mem_space): per-clip map blocks
	                                 used instead of the single packed `maps` pointer (which may then be NULL) */
} rvb_batch;

/* IoU evaluation of one method run against U annotators
 * (retargetvid_eval.py:10-27,161-194). */
typedef struct rvb_iou_batch {
	int32_t n_videos;
	int32_t n_users;
	int32_t mem_space;
	int32_t reserved0;
	const int64_t *frame_offset;  /* HOST [n_videos + 1] into the packed box arrays */
	const int32_t *n_eval;        /* HOST [n_videos] frames evaluated (frame_counts, clipped to the shorter list) */
	const int32_t *method_boxes;  /* [sum F][4] */
	const int32_t *annot_boxes;   /* [n_users][sum F][4] */
	double   *frame_iou;          /* optional [n_users][sum F] */
	uint64_t *acc;                /* [n_videos][n_users][2]: exact sum of the IoU doubles in units of 2^-80, lo then hi */
	const int32_t *n_eval_user;   /* optional HOST [n_videos][n_users]: frames evaluated per (video, annotator); overrides n_eval
	                                 (the reference breaks out of the frame loop per annotator, retargetvid_eval.py:163-179) */
	int32_t  *n_bad;              /* optional [1]: IoUs outside [0, 1] (malformed boxes: x2 < x1, empty union).  The reference
	                                 averages such negative values or raises ZeroDivisionError (retargetvid_eval.py:24-26);
	                                 with host buffers the call returns RVB_ERR_INVALID when the count is not 0 */
} rvb_iou_batch;

const char *rvb_version(void);
const char *rvb_last_error(void);

int rvb_ctx_create(int device, rvb_ctx **out);
int rvb_ctx_destroy(rvb_ctx *ctx);
/* stream: a cudaStream_t (NULL = the context's own stream) */
int rvb_ctx_set_stream(rvb_ctx *ctx, void *stream);
int rvb_ctx_synchronize(rvb_ctx *ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t rvb_ctx_launch_count(const rvb_ctx *ctx);
/* CUDA-event time (ms) and launches of the dominant kernel (the fused map
 * kernel) summed over the last crop_track call; valid after a synchronise */
int rvb_ctx_last_map_kernel_ms(rvb_ctx *ctx, float *ms, int32_t *launches);
/* split pipeline of the last crop_track call, CUDA-event times (ms) on the launching stream:
 * out = {front launches, Prim + back launches of the maps outside cut-adjacent chains (the size classes one after the
 * other), 0, whole map pipeline incl. the joined side stream with the chains} */
int rvb_ctx_last_stage_ms(rvb_ctx *ctx, float out[4]);

/* CUDA-event time (ms) of the IoU kernel of the last rvb_iou_batch_run call (without the table upload around it) */
int rvb_ctx_last_iou_kernel_ms(rvb_ctx *ctx, float *ms);

/* host time (microseconds) the last crop_track call spent per section: {per-clip tables, packing + upload of the tables,
 * map input (copies / transposition launches), map pipeline launches, track launches, results (host buffers: incl. the
 * wait for the device)} -- profiling aid for latency-bound calls, no reference equivalent */
int rvb_ctx_last_host_us(rvb_ctx *ctx, double out[6]);

/* profiling aid: when enabled, the map kernel accumulates SM cycles per phase (11 phases: load, threshold+compact,
 * core distances, Prim, argsort emulation, Cartesian tree, condensed-tree BFS, fall-out, EOM+labels, rebuild+closing,
 * results), summed over CTAs since the last call of this function */
int rvb_ctx_phase_cycles(rvb_ctx *ctx, int enable, uint64_t out[16]);

/* replaces sc_init_crop_params(use_best_settings) -- smartVidCrop.py:132-209 */
int rvb_params_default(rvb_params *p, int use_best_settings);

/* replaces steps 1-13 of smart_vid_crop (smartVidCrop.py:2293-2521) for a batch
 * of clips: destination size, border detection, mean saliency, threshold,
 * clustering filter + cut blend, coverage score, centre of mass, empty-centre
 * handling, interpolation, low-pass, LOESS / Savitzky-Golay, boxes, time shift. */
int rvb_crop_track_batch(rvb_ctx *ctx, const rvb_params *p, const rvb_batch *b);

/* replaces the frame loop of retargetvid_eval.py:161-194 */
int rvb_iou_batch_run(rvb_ctx *ctx, const rvb_iou_batch *b);
/* exactly rounded mean of n IoU doubles from their exact sum (statistics.mean,
 * retargetvid_eval.py:193): acc = {lo, hi} in units of 2^-80 */
double rvb_iou_mean_from_acc(const uint64_t acc[2], int64_t n);

/* replaces the per-frame crop of sc_renderer (smartVidCrop.py:1906-1912): out[f] = frames[f][by1:by2, bx1:bx2, :].
 * frames: uint8 [n_frames][h][w][channels] and out: uint8 [n_frames][out_h][out_w][channels], both in mem_space;
 * boxes: HOST int32 [n_frames][4] = x1,y1,x2,y2 as rvb_crop_track_batch returns them; every box must measure
 * out_w x out_h and lie inside the frame (RVB_ERR_INVALID otherwise). */
int rvb_crop_frames(rvb_ctx *ctx, const uint8_t *frames, int32_t n_frames, int32_t h, int32_t w, int32_t channels,
                    const int32_t *boxes, int32_t out_h, int32_t out_w, uint8_t *out, int32_t mem_space);

/* a16 and its reader, host code (no context): the text format of the result files -- '%d,%d,%d,%d\n' per frame
 * (smartVidCrop.py:2783-2785), read back as int(c[0]) .. int(c[3]) of line.split(',') (retargetvid_eval.py:152-159).
 * rvb_format_boxes_txt: boxes HOST int32 [n_frames][4] -> text; `out` needs 48 bytes per frame; *len_out = bytes written.
 * rvb_parse_boxes_txt: text of `len` bytes -> boxes (capacity cap_frames); *n_frames_out = lines parsed.  A line that
 * Python's int() / indexing would reject (fewer than 4 fields, an empty or non-integer field) is RVB_ERR_INVALID with
 * the line number in rvb_last_error(); more lines than cap_frames: RVB_ERR_CAPACITY with *n_frames_out = lines present. */
int rvb_format_boxes_txt(const int32_t *boxes, int64_t n_frames, char *out, int64_t cap, int64_t *len_out);
int rvb_parse_boxes_txt(const char *text, int64_t len, int32_t *boxes, int64_t cap_frames, int64_t *n_frames_out);

/* host-only: scipy.signal.butter(order, Wn) and lfilter_zi as the low-pass stage designs them (smartVidCrop.py:1601-1605):
 * b, a: order + 1 doubles, zi: order doubles; *chunked_ok = 1 if the kernel evaluates this filter in parallel chunks, 0 if
 * it is so ill conditioned that only scipy's sequential order of operations reproduces scipy's output */
int rvb_debug_butter(int32_t order, double wn, double *b, double *a, double *zi, int32_t *chunked_ok);

/* stage-level entry points used by the parity tests (device work only) */
/* sc_clustering_filt on one uint8 map in host memory -- smartVidCrop.py:1062-1161 */
int rvb_debug_cluster_labels(rvb_ctx *ctx, const rvb_params *p, const uint8_t *map_hw, int32_t h, int32_t w,
                             int32_t *labels_out /* [n_points] in row-major point order */, int32_t *n_points_out);
/* pyloess.Loess.estimate for every j of one series -- pyloess.py:61-95 via loess_handler */
int rvb_debug_smooth_series(rvb_ctx *ctx, const rvb_params *p, const double *series_in, int32_t n, double fr,
                            double *lowpassed_out, double *smoothed_out);

#ifdef __cplusplus
}
#endif
#endif
