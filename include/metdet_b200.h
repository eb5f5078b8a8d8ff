/*
 * metdet_b200.h -- C ABI of libmetdet_b200.so: the B200 (sm_100a) implementation of MetDetPy's
 * per-frame line-detector hot path.  Plain pointers and sizes only; no torch / numpy types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * repository LilacMeteorObservatory/MetDetPy).  The reference is pure Python, so "the FFI a
 * maintainer would add" is a ctypes binding; see INTEGRATION.md and metdetpy_b200/_lib.py.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MDB_ERR_* code otherwise;
 *     mdb_last_error() returns a thread-local human-readable message for the last failure.
 *   - the caller owns all input and output buffers; a handle owns its device memory and one CUDA
 *     stream.  Handles are independent (no global mutable state): different handles may be driven
 *     from different threads; one handle from one thread at a time (the reference's threading
 *     model: MetDetPy.py:184-227 runs the detector on the main thread only).
 *   - frames are uint8, H rows of W pixels, C-contiguous (what VanillaVideoLoader hands to
 *     detector.update(), videoloader.py:360-388); `on_device` != 0 means the pointer is a CUDA
 *     device pointer on the handle's device, already complete (caller synchronised its producer).
 *   - there is NO CPU fallback: without a CUDA device every entry point that computes fails.
 */
#ifndef METDET_B200_H
#define METDET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDB_OK 0
#define MDB_ERR_INVALID (-1)   /* bad argument / unsupported configuration */
#define MDB_ERR_CUDA (-2)      /* CUDA runtime error (message has the CUDA error string) */
#define MDB_ERR_STATE (-3)     /* call sequence error (e.g. detect before any update) */
#define MDB_ERR_NOMEM (-4)

#define MDB_NUM_LINES_TOOMUCH 500 /* MetLib/Detector.py:30 */
#define MDB_MAX_LINES 512         /* per-frame capacity of the raw / NMS line outputs (> 500) */
#define MDB_MAX_WINDOW 4096       /* longest window n (the streaming kernels serve n <= 128, the generic kernel the rest) */

#define MDB_SENS_LOW 0
#define MDB_SENS_NORMAL 1
#define MDB_SENS_HIGH 2

/* Constructor arguments of M3Detector / LineDetector (MetLib/Detector.py:186-220, :319-322)
 * after the host has evaluated the pure-Python parts (int(window_sec*fps), select_subarea). */
typedef struct mdb_config {
    int32_t width, height;
    int32_t window;          /* n = int(window_sec * fps), 1 <= n <= MDB_MAX_WINDOW  Detector.py:197 */
    int32_t adaptive;        /* cfg.binary.adaptive_bi_thre                     Detector.py:204 */
    int32_t init_value;      /* cfg.binary.init_value (used when !adaptive)     Detector.py:208 */
    int32_t sensitivity;     /* MDB_SENS_*  (cfg.binary.sensitivity)            Detector.py:177-183 */
    int32_t nz_interval;     /* cfg.binary.interval                             Detector.py:202 */
    int32_t roi[4];          /* SNR_SW.std_roi = (r0, c0, r1, c1)               Detector.py:93-122 */
    int32_t hough_threshold; /* cfg.hough_line.threshold                        Detector.py:350 */
    int32_t hough_min_len;   /* cfg.hough_line.min_len                          Detector.py:351 */
    int32_t hough_max_gap;   /* cfg.hough_line.max_gap                          Detector.py:344 */
    int32_t dy_mask;         /* cfg.dynamic.dy_mask                             Detector.py:211 */
    int32_t max_batch;       /* most frames one mdb_detect_batch call may carry (>= 1) */
    int32_t device;          /* CUDA device ordinal */
    int32_t apply_mask;      /* != 0: multiply incoming frames by the mask on the device
                                (Transform.mask_with, MetLib/imgproc.py:96-101) */
    int32_t detector;  /* 0 = M3Detector (Detector.py:302-392), 1 = ClassicDetector (Detector.py:245-299:
                        * window is 4 frames whatever `window` says, dy_mask ignored, fixed maxLineGap, no NMS) */
    int32_t reserved[3];
} mdb_config;

/* Per-frame scalars that M3Detector exposes as attributes (Detector.py:227-229, :342-344, :355,
 * :373; read by visu() :394-448). */
typedef struct mdb_frame_info {
    int64_t timer;             /* SlidingWindow.timer after this frame          utils.py:270 */
    int32_t bi_threshold;      /* LineDetector.bi_threshold                     Detector.py:229 */
    int32_t n_on;              /* number of 255-pixels in dst */
    double bi_threshold_float; /* LineDetector.bi_threshold_float               Detector.py:228 */
    double snr;                /* SNR_SW.snr (noise EMA value)                  Detector.py:124-127 */
    double dst_sum;            /* M3Detector.dst_sum                            Detector.py:342 */
    double gap;                /* maxLineGap handed to HoughLinesP              Detector.py:343-344 */
    int32_t lines_num;         /* M3Detector.lines_num (raw HoughLinesP count)  Detector.py:357 */
    int32_t n_raw;             /* rows valid in raw_lines (0 when lines_num > 500, :358-360) */
    int32_t n_lines;           /* M3Detector.filtered_line_num (after NMS)      Detector.py:373 */
    int32_t len_ties;          /* != 0: two raw segments have the same length; the NMS order among them (numpy's
                                  unstable argsort, utils.py:804) is then the host's business: the library used
                                  "descending index", a caller that wants numpy's order on its machine redoes the frame
                                  with mdb_lineset_nms_ordered */
} mdb_frame_info;

typedef struct mdb_detector *mdb_handle;

const char *mdb_last_error(void);
int mdb_version(void);
/* number of visible CUDA devices (0 if none / no driver); never fails */
int mdb_device_count(void);

/* M3Detector.__init__ (Detector.py:319-322 -> LineDetector.__init__ :186-220 -> SNR_SW.__init__
 * :41-71).  `mask` is the host uint8 {0,1} mask of shape (height,width) (fileio.load_mask). */
int mdb_create(const mdb_config *cfg, const uint8_t *mask, mdb_handle *out);
int mdb_destroy(mdb_handle h);

/* M3Detector.update(new_frame) (Detector.py:225-229): SNR_SW.update (:73-91) incl. the noise
 * sample / EMA / adaptive threshold recurrence. */
int mdb_update(mdb_handle h, const uint8_t *frame, int on_device);

/* M3Detector.detect() (Detector.py:324-392).  lines: up to MDB_MAX_LINES rows [x1,y1,x2,y2] after
 * lineset_nms (utils.py:780-839); nonline_prob: one double per row (cls_pred[:, -1]);
 * raw_lines (optional, may be NULL): linesp_ext, up to MDB_MAX_LINES rows. */
int mdb_detect(mdb_handle h, mdb_frame_info *info, int32_t *lines, double *nonline_prob,
               int32_t *raw_lines);

/* T x { update(frame[t]); detect() } in one call -- the batched form of MetDetPy.py:192-198.
 * infos[T]; lines[T][MDB_MAX_LINES][4]; nonline_prob[T][MDB_MAX_LINES]; raw_lines optional
 * ([T][MDB_MAX_LINES][4] or NULL); dst_out optional ([T][H][W] host or device buffer, or NULL). */
int mdb_detect_batch(mdb_handle h, const uint8_t *frames, int T, int on_device,
                     mdb_frame_info *infos, int32_t *lines, double *nonline_prob,
                     int32_t *raw_lines, uint8_t *dst_out, int dst_on_device);

/* Asynchronous halves of mdb_detect_batch for overlapping the next batch's host->device copy with
 * this batch's kernels: submit enqueues copy + all kernels + the small result copy and returns;
 * collect waits for the OLDEST submitted batch and runs the host NMS.  Up to three batches may be in
 * flight per handle (submit, submit, collect, submit, collect, ...); each keeps its own dst masks. */
int mdb_submit_batch(mdb_handle h, const uint8_t *frames, int T, int on_device);
int mdb_collect_batch(mdb_handle h, mdb_frame_info *infos, int32_t *lines, double *nonline_prob,
                      int32_t *raw_lines, uint8_t *dst_out, int dst_on_device);

/* ---- time-sharded streams (one detector per GPU, each owning a chunk of the frame sequence) -----
 * The reference is a single sequential loop (MetDetPy.py:184-227); these three calls are what lets
 * G detectors reproduce it exactly on G chunks:
 *  mdb_seek        start the frame counter (SlidingWindow.timer, utils.py:270) at a global index, so
 *                  that warm-up rules (length = min(n, timer)) fire only at the true stream start.
 *  mdb_noise_sums  the integer sums behind SNR_SW.update's noise sample (Detector.py:81-91) for every
 *                  sample timer inside `frames` (global index of frames[0] = t0): sums[2i] = sum d,
 *                  sums[2i+1] = sum d^2 for frame i (zero if that timer is no sample or its window
 *                  starts before t0). Stateless. The EMA / threshold recurrence is then replayed on
 *                  the host over the samples of all chunks.
 *  mdb_submit_batch_thr  mdb_submit_batch with per-frame thresholds supplied by the caller instead of
 *                  the handle's own recurrence. */
/* mdb_submit_batch / mdb_submit_batch_thr (thr arrays NULL: the handle's own recurrence) with flags:
 *  MDB_SUBMIT_HALO  the frames are the look-back halo of a time-sharded chunk: they only fill the window and the
 *                   dynamic-mask history (temporal pass + median/close); no mask bytes, no Hough; the collected
 *                   frame infos carry the thresholds, n_on = 0 and no lines.  (After mdb_seek the first frames see
 *                   an empty window and would otherwise produce full-frame masks.) */
#define MDB_SUBMIT_HALO 1
int mdb_submit_batch_ex(mdb_handle h, const uint8_t *frames, int T, int on_device, const int32_t *thr,
                        const double *thr_float, const double *snr, int flags);
int mdb_seek(mdb_handle h, int64_t timer);
int mdb_noise_sums(const uint8_t *frames, int T, int on_device, int64_t t0, int width, int height,
                   int window, int nz_interval, const int32_t *roi, const uint8_t *mask,
                   uint64_t *sums, int device);
int mdb_submit_batch_thr(mdb_handle h, const uint8_t *frames, int T, int on_device, const int32_t *thr,
                         const double *thr_float, const double *snr);

/* M3Detector.dst (Detector.py:371): the binary mask of the most recent detect, (H,W) uint8. */
int mdb_get_dst(mdb_handle h, uint8_t *dst, int on_device);
/* device pointer of the dst masks of the most recent batch ([T][H][W]); valid until the next call */
int mdb_get_dst_device(mdb_handle h, const uint8_t **ptr);

/* SlidingWindow.max / .mean / .sum (utils.py:288-300) of the detector's main window. Any output
 * may be NULL. Host buffers of H*W elements. */
int mdb_get_stack(mdb_handle h, uint8_t *max_out, uint8_t *mean_out, uint32_t *sum_out);
/* SlidingWindow.sliding_window (MetLib/utils.py:263-265): out[n][H][W], the ring in the reference's slot order (slot i =
 * newest frame whose 0-based index is congruent to i modulo n; never-written slots are zero). */
int mdb_get_window(mdb_handle h, uint8_t *out, int on_device);
/* SlidingWindow.std in the uint8 / force_int mode (MetLib/utils.py:309-321): sqrt(mean((sum(x^2) - sum(x)^2 // L) // L))
 * with numpy's uint32 arithmetic per pixel; exact integer total on the device. */
int mdb_get_std(mdb_handle h, double *std_out);
/* Every raw Hough segment of frame `frame` (0-based) of the batch collected last, without the MDB_MAX_LINES cap of the
 * batched outputs: ClassicDetector returns all of them (Detector.py:282-292).  *n_out = the count; with cap == 0 only the
 * count is returned; cap < count is an error. */
int mdb_get_raw_lines(mdb_handle h, int frame, int32_t *out, int cap, int32_t *n_out);

/* The handle's CUDA stream (cudaStream_t) -- for callers that time with CUDA events. */
int mdb_get_stream(mdb_handle h, void **stream);
/* Number of kernels this handle has launched since creation. */
int mdb_get_launch_count(mdb_handle h, int64_t *count);
/* Device time (ms, CUDA events on the handle's stream) the fused mask kernel(s) took in the most
 * recent batch, and how many launches that was. */
int mdb_get_fused_time(mdb_handle h, float *ms, int32_t *launches);
/* Back to the state right after mdb_create (empty window, initial thresholds, timer 0) without giving up the
 * device buffers; mdb_seek may follow.  No batch may be in flight. */
int mdb_reset(mdb_handle h);
/* Noise sums (the integer pair per noise sample that SNR_SW.update, MetLib/Detector.py:73-91, turns into a
 * standard deviation) of the sample timers among DEVICE frames, for nseg segments in one call: segment k = T[k] frames
 * at frames[k], global index of the first one t0[k].  Only samples whose whole window lies inside their segment (or
 * starts at global frame 0) are evaluated; sums[sum of T][2] (host, segments concatenated) is zero elsewhere.  Runs on
 * the handle's own stream and buffers; synchronous.  This is what one rank of a time-sharded run computes for its
 * chunk before the ranks exchange the sums and replay the threshold recurrence. */
int mdb_noise_sums_dev(mdb_handle h, int nseg, const uint8_t *const *frames, const int32_t *T, const int64_t *t0,
                       uint64_t *sums);
/* Host-only: EMA.update (MetLib/utils.py:334-368) over nsamples noise samples (timers ascending, sums[k][2] as
 * produced by mdb_noise_sums*) and LineDetector.update's threshold rule (MetLib/Detector.py:225-229) for frames
 * 0 .. t_end-1; thr / thr_float / snr receive the values of frames t_begin .. t_end-1.  Bit-identical to what a
 * detector handle computes on the device for the same samples.  Every sample timer up to t_end must be present. */
int mdb_replay_thresholds(int nsamples, const int64_t *timers, const uint64_t *sums, int64_t roi_pixels, int window,
                          int nz_interval, int adaptive, int init_value, int sensitivity, int64_t t_begin,
                          int64_t t_end, int32_t *thr, double *thr_float, double *snr);
/* Read-only counters / timings of the most recent batch by name: "temporal_ms" (stack->diff->threshold pass),
 * "spatial_ms" (median + close + dy-mask + mask bytes), "temporal_generation" (which temporal kernel ran: 3 =
 * register ring, 2 = shared-memory ring, 4 = per-frame resident state, 0 = none), "digest_hi" / "digest_lo" (the two
 * 32-bit halves of an FNV-1a digest of the batch finished last: thresholds, on-pixel counts, raw segments), "stream_kernel" (1 if the time-tiled
 * streaming path serves this handle), "hough_tier1a" / "hough_tier1b" / "hough_tier2" / "hough_tier3" (frames of the
 * batch collected last that each PPHT tier resolved; "..._total": since creation).  Unknown names return
 * MDB_ERR_INVALID. */
int mdb_get_info(mdb_handle h, const char *name, double *value);
/* Tuning / test knobs by name (the parity tests force every kernel variant through these): "stream_kernel" (0 = the
 * generic per-frame kernel), "temporal_version" (1, 2, 3), "t3_variant", "temporal_wpt", "temporal_nt", "temporal_kdiv",
 * "force_dense", "force_strip", "sp_rows", "dst_rows", "timeline", "hough_profile".  Not needed in production. */
int mdb_set_option(mdb_handle h, const char *name, int value);
/* debug: event timeline of the batch collected last (after mdb_set_option("timeline", 1)): out[9] ms offsets */
int mdb_debug_timeline(mdb_handle h, float *out);
/* debug: per-frame PPHT phase cycle counters (after mdb_set_option("hough_profile", 1)): out[T][10] */
int mdb_debug_hough_profile(mdb_handle h, long long *out, int T);

/* stacker.max_stacker / MaxImgContainer (MetLib/stacker.py:43-49, :146-175, :197-213) and
 * MergeFunction.max (MetLib/utils.py:203-204): element-wise max over T frames of frame_bytes
 * bytes each (any layout, e.g. H*W*3 colour).  frames/out are host or device per the flags. */
int mdb_max_stack(const uint8_t *frames, int T, size_t frame_bytes, uint8_t *out,
                  int frames_on_device, int out_on_device, int device);

/* lineset_nms (MetLib/utils.py:780-839) on the host: lines_in[n][4] -> lines_out, prob_out;
 * returns the number of kept lines in *n_out.  Ordering: len^2 descending, ties by descending
 * input index (what np.argsort(...)[::-1] yields for the n <= 16 insertion-sort regime). */
int mdb_lineset_nms(const int32_t *lines_in, int n, int32_t *lines_out, double *prob_out,
                    int32_t *n_out);
/* Same greedy pass over a visiting order supplied by the caller (a permutation of 0..n-1, longest first).  The
 * reference orders with np.argsort(len^2)[::-1] (MetLib/utils.py:804), whose order among EQUAL lengths is numpy's
 * business for n > 16; the Python layer passes numpy's own order here whenever lengths tie, so the kept lines are the
 * reference's on the same host. */
int mdb_lineset_nms_ordered(const int32_t *lines_in, int n, const int32_t *order, int32_t *lines_out,
                            double *prob_out, int32_t *n_out);
/* The same for several frames of a finished batch at once (the frames whose mdb_frame_info.len_ties is set): frames[j]
 * indexes infos / raw_lines / lines / nonline_prob as mdb_collect_batch / mdb_detect_batch filled them; orders holds one
 * permutation of 0..n_raw-1 per frame, back to back (order_off[j] .. order_off[j+1]).  Rewrites lines, nonline_prob and
 * infos[].n_lines of those frames. */
int mdb_lineset_nms_frames(int k, const int32_t *frames, const int32_t *orders, const int64_t *order_off,
                           const int32_t *raw_lines, mdb_frame_info *infos, int32_t *lines, double *nonline_prob);

/* FastGaussianContainer.append over T frames (MetLib/stacker.py:52-59; FastGaussianParam.__init__/__add__,
 * MetLib/utils.py:435-452, :485-493): per element sum_out = sum of the frames as uint16 and sq_out = sum of
 * their squares as uint32, both wrapping like numpy's fixed-width adds (n = T is the caller's).  accumulate != 0
 * continues from the values already in sum_out / sq_out.  frames/outputs are host or device per the flags. */
int mdb_gauss_stack(const uint8_t *frames, int T, size_t frame_bytes, uint16_t *sum_out, uint32_t *sq_out,
                    int frames_on_device, int out_on_device, int accumulate, int device);

/* ---- MFNR mix stacker on the device (second half) ---------------------
 * mfnr_mix_stacker, MetLib/stacker.py:296-403, for connect_lines.switch == false and all four background algorithms: "mean"
 * (:339-342), "sigma-clipping" (:333-338, single_sigma_clipping :94-115), "median" / "med-of-med" (:343-349, :62-78).  Frames ((H, W, C) uint8, any C <= 4) are
 * appended as the loader delivers them (what _batch_stacker, :146-175, feeds MaxImgContainer / AllImgContainer /
 * FastGaussianContainer); with keep_frames they stay resident on the device for the clipping pass.  The finishing
 * passes run in float64 like the reference; the two global means are reduced in a fixed order that is not numpy's
 * pairwise one, so the result equals the reference's up to the last ulp of those scalars: a mixed pixel may differ by
 * one grey level where it sits on a rounding boundary (tests hold <= 1 level on <= 1e-4 of the elements).
 * Not built: connect_highlight_area (:239-294). */
typedef struct mdb_mfnr *mdb_mfnr_handle;
typedef struct mdb_mfnr_params {
    double highlight_preserve; /* DenoiseOption.highlight_preserve                      stacker.py:313 */
    int32_t blur_ksize;        /* DenoiseOption.blur_ksize (odd)                        stacker.py:368 */
    int32_t bg_algorithm;      /* 0 = "mean", 1 = "sigma-clipping", 2 = "median", 3 = "med-of-med"   stacker.py:333-349 */
    double blur_sigma;         /* 3 in the reference (sigmaX=3); <= 0: 3                stacker.py:370 */
    double sigma_high;         /* single_sigma_clipping arguments (the reference passes 3.0, 3.0: :335-336) */
    double sigma_low;
    double bg_fix_factor;      /* MFNRDenoiseParam.bg_fix_factor                        stacker.py:353 */
    double gumbel_mean;        /* get_gumbel_mean(n) as the caller computed it, or <= 0 to have it computed here */
    int32_t med_block_size;    /* "med-of-med": int(len(img_stack) ** 0.5) as the caller computed it (stacker.py:68-69), or 0 */
    int32_t reserved;
} mdb_mfnr_params;
int mdb_mfnr_create(int height, int width, int channels, int keep_frames, int device, mdb_mfnr_handle *out);
/* optional: device memory for `frames` more retained frames in one allocation (keep_frames handles only) */
int mdb_mfnr_reserve(mdb_mfnr_handle m, int frames);
int mdb_mfnr_append(mdb_mfnr_handle m, const uint8_t *frames, int T, int on_device);
/* out: (H, W, C) uint8; stats (optional, 4 doubles): est_bg_var, gumbel mean, highlight_avg_diff, count of positive diffs */
int mdb_mfnr_finish(mdb_mfnr_handle m, const mdb_mfnr_params *params, uint8_t *out, int out_on_device, double *stats);
/* the running statistics as the reference's containers hold them (stacker.py:43-59): max (uint8), sum (uint16, wrapping),
 * sum of squares (uint32, wrapping) of the frames appended so far; any output may be NULL; n_frames optional */
int mdb_mfnr_stats(mdb_mfnr_handle m, uint8_t *max_out, uint16_t *sum_out, uint32_t *sq_out, int64_t *n_frames);
int mdb_mfnr_destroy(mdb_mfnr_handle m);

/* ---- loader preprocessing on the device ---------------------------
 * Replaces, for uint8 frames, what the reference's video loader applies to every decoded frame:
 *   Transform.opencv_resize   = cv2.resize(img, dsize, INTER_LINEAR)   MetLib/imgproc.py:82-85
 *   Transform.opencv_BGR2GRAY = cv2.cvtColor(img, COLOR_BGR2GRAY)      MetLib/imgproc.py:87-88 (:90-91 RGB)
 *   Transform.mask_with       = img * mask                             MetLib/imgproc.py:96-101
 *   Transform.exec_transform  (the chain, built at MetLib/videoloader.py:300-308)  imgproc.py:129-139
 *   MergeFunction.max over exp_frame consecutive frames                MetLib/utils.py:203-204, videoloader.py:388
 * Results are bit-exact with cv2 (fixed-point resize and gray).  channels = 1 (gray source: resize and
 * mask only) or 3 (interleaved BGR, or RGB with rgb_order = 1; output is always one channel).
 * mask: host pointer to dst_h*dst_w bytes of {0,1}, or NULL.  The handle owns its stream, the tap
 * tables, the mask and an output buffer of max_out frames. */
typedef struct mdb_preproc *mdb_preproc_handle;
int mdb_preproc_create(int src_w, int src_h, int channels, int rgb_order, int dst_w, int dst_h,
                       const uint8_t *mask, int exp_frame, int max_out, int device,
                       mdb_preproc_handle *out);
/* T source frames ([T][src_h][src_w][channels], host or device) -> ceil(T / exp_frame) output frames
 * ([.][dst_h][dst_w]).  out = NULL keeps the result in the handle's device buffer (see
 * mdb_preproc_output) so that it can be fed to mdb_submit_batch(..., on_device = 1) without leaving
 * the GPU; otherwise it is copied to `out` (host or device per out_on_device).  Synchronous. */
int mdb_preproc_run(mdb_preproc_handle h, const uint8_t *frames, int T, int frames_on_device,
                    uint8_t *out, int out_on_device, int32_t *n_out);
int mdb_preproc_output(mdb_preproc_handle h, const uint8_t **device_ptr);
/* device time (ms, CUDA events) of the most recent run's kernel */
int mdb_preproc_time(mdb_preproc_handle h, float *ms);
int mdb_preproc_destroy(mdb_preproc_handle h);
/* host-only: the source index pair and 11-bit weights cv2's 8-bit INTER_LINEAR uses for every
 * destination coordinate of one axis (x axis: clamp_fraction = 1, y axis: 0) -- what the kernel consumes */
int mdb_preproc_axis_taps(int dst, int src, int clamp_fraction, int32_t *s0, int32_t *s1, int32_t *w0,
                          int32_t *w1);

/* pinned host memory for frame staging (cudaHostAlloc / cudaFreeHost) */
int mdb_alloc_pinned(size_t bytes, void **ptr);
int mdb_free_pinned(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* METDET_B200_H */
