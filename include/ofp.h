/*
 * ofp.h -- C ABI of libofp.so, the B200 (sm_100a) implementation of onset-fingerprinting's
 * data-parallel hot path.  This is the drop-in boundary: every entry point below replaces a
 * native interface or a Python hot loop of the reference (file:line under
 * /root/reference/onset_fingerprinting/ given per function).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  Pointers named *_dev are DEVICE pointers owned by
 *     the caller (e.g. torch allocations); *_host are host pointers.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), no hidden
 *     synchronisation, except the *_host convenience entry points which synchronise before
 *     returning because they hand results back in host memory.
 *   - return value: 0 on success, a negative OFP_E* code otherwise; ofp_last_error() gives a
 *     thread-local message.  The reference's ctypes DLL has no error reporting at all
 *     (detection.py:520-538 sets argtypes only), so this is an extension, not a change.
 *   - the library allocates nothing persistent except the opaque ofp_detector / ofp_ccstream
 *     handles.  A handle must not be used from two streams concurrently (the reference's
 *     follower objects are equally stateful and not thread safe).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     OFP_ECUDA.
 */
#ifndef OFP_H
#define OFP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFP_OK 0
#define OFP_EINVAL (-1)  /* bad argument */
#define OFP_ECUDA (-2)   /* CUDA runtime / driver error, see ofp_last_error() */
#define OFP_ENOMEM (-3)
#define OFP_EUNSUPPORTED (-4)

const char *ofp_last_error(void);
/* "libofp <version> sm_100a" */
const char *ofp_version(void);

/* Strided host <-> device copy of `rows` rows of `width_bytes` (cudaMemcpy2DAsync on `stream`): a time segment
 * of every recording of a [R, N, C] batch in one call.  The reference has no counterpart (its arrays never
 * leave the host); used by the host-buffer entry points (pipeline.HotPath.run_host). */
int ofp_copy2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes, size_t rows,
                     int to_host, void *stream);

/* ---------------------------------------------------------------------------------------
 * K1  amplitude onset detector
 *     replaces AmplitudeOnsetDetector (detection.py:595-840) and the three ctypes calls it
 *     makes per block into envelope_follower.so (ar_envelope x2, minmax_envelope;
 *     envelope_follower.c:6-57), plus ButterworthFilter (detection.py:487-501).
 * ------------------------------------------------------------------------------------- */

/* AmplitudeOnsetDetector.__init__ arguments after the host-side conversions the reference
 * performs (detection.py:493-496, 514-515, 553-555, 683-712). */
typedef struct ofp_detector_params {
    int32_t n_channels;   /* n_signals, 1..32 */
    int32_t block_size;   /* block_size */
    int32_t use_hp;       /* hipass_freq != 0 */
    int32_t manual;       /* on_threshold > 1: absolute thresholds, min/max tracker unused */
    int32_t cooldown;     /* samples */
    float b[5], a[5];     /* float32(scipy.signal.butter(4, hipass_freq, 'high', fs=sr)) */
    float floor_db;       /* floor */
    float fast_att, fast_rel, slow_att, slow_rel; /* float32(1/attack), float32(1/release) */
    float on_thr, off_thr;                        /* on_threshold, off_threshold */
    float alpha_min, alpha_max, minmin;           /* 1e-4, 1e-5, 2 in the reference */
} ofp_detector_params;

typedef struct ofp_detector ofp_detector;

/* One detector per stream/recording: n_streams independent AmplitudeOnsetDetector states
 * (filter delay line, both followers, min/max, FSM) resident in device memory. */
int ofp_detector_create(ofp_detector **out, int64_t n_streams, const ofp_detector_params *p);
int ofp_detector_destroy(ofp_detector *det);
/* Back to the state right after __init__. */
int ofp_detector_reset(ofp_detector *det, void *stream);
/* Copy one field of the per-lane state out of / into the detector (checkpoint, inspection):
 * fields 0..3 z[0..3], 4 yf, 5 ys, 6 min, 7 max, 8 prev (float32); 9 state, 10 debounce (int32).
 * buf_dev: device buffer of n_streams*n_channels 32-bit words, lane = stream*C + channel. */
int ofp_detector_get_state(ofp_detector *det, int field, void *buf_dev, void *stream);
int ofp_detector_set_state(ofp_detector *det, int field, const void *buf_dev, void *stream);

/* detect_onsets_amplitude (detection.py:19-86) for a batch of recordings.
 *   x_dev      [R, n_samples, C] float32, recording r starts at x_dev + r*rec_stride (elements)
 *   warm_n     samples of warm-up (init_minmax_tracker, detection.py:70,827-840); 0 = none
 *   rel_dev    NULL (onsets only) or [R, n_blocks*B, C] float32, recording stride rel_stride
 *   on_channel_dev, on_sample_dev  [R, cap] int32: onsets of recording r in the reference's
 *              order (block-major, channel ascending); on_count_dev [R] int32 is the number
 *              found (may exceed cap; only the first cap are stored)
 * R must equal the detector's n_streams.  State is carried in `det` (continue with another
 * call or with ofp_detect_block). */
int ofp_detect_offline(ofp_detector *det, const float *x_dev, int64_t n_samples, int64_t rec_stride,
                       int64_t warm_n, float *rel_dev, int64_t rel_stride, int32_t *on_channel_dev,
                       int32_t *on_sample_dev, int32_t *on_count_dev, int32_t cap, void *stream);

/* Continue ofp_detect_offline on the NEXT n_samples (whole blocks) of every recording from the state the
 * previous call left in `det` -- the block loop of detect_onsets_amplitude (detection.py:74-84) resumed at
 * block `first_block`.  x_dev / rel_dev hold only this segment ([R, n_samples, C]); onset sample indices are
 * global (first_block * B + ...), and detections are appended after the on_count_dev[r] entries already in
 * on_channel_dev / on_sample_dev (on_count_dev is read and updated). */
int ofp_detect_continue(ofp_detector *det, const float *x_dev, int64_t n_samples, int64_t rec_stride,
                        int64_t first_block, float *rel_dev, int64_t rel_stride, int32_t *on_channel_dev,
                        int32_t *on_sample_dev, int32_t *on_count_dev, int32_t cap, void *stream);

/* AmplitudeOnsetDetector.__call__ (detection.py:727-798) for n_streams concurrent streams:
 *   x_dev [S, B, C] with stream s starting at x_dev + s*stream_stride (elements; B*C when dense);
 *   rel_dev NULL or [S, B, C]; ch_dev/delta_dev [S, C] int32; count_dev [S]. */
int ofp_detect_block(ofp_detector *det, const float *x_dev, int64_t stream_stride, float *rel_dev, int32_t *ch_dev,
                     int32_t *delta_dev, int32_t *count_dev, void *stream);

/* AmplitudeOnsetDetector.init_minmax_tracker (detection.py:827-840): x_dev [S, n, C]. */
int ofp_detect_warmup(ofp_detector *det, const float *x_dev, int64_t n_samples, int64_t rec_stride,
                      void *stream);

/* Host-buffer convenience (the reference-facing call: numpy in, numpy out).  Streams x_host to the device
 * in time segments of all recordings (copy of segment s+1 overlapping the kernel of segment s, which
 * continues from the detector state of segment s-1), copies results back and synchronises.  Same results as
 * ofp_detect_offline on the whole batch.  rel_host may be NULL; pinned host memory makes the copies
 * asynchronous.  The two staging segments, streams and events are kept for the next call (allocating and
 * freeing multi-GB buffers costs more than the pipeline); ofp_host_release() frees them.  Calls are
 * serialised by a mutex. */
int ofp_host_release(void);
int ofp_detect_offline_host(const ofp_detector_params *p, const float *x_host, int64_t n_rec,
                            int64_t n_samples, int64_t warm_n, float *rel_host, int32_t *on_channel_host,
                            int32_t *on_sample_host, int32_t *on_count_host, int32_t cap);

/* backtrack_onsets (detection.py:800-825; C twin envelope_follower.c:59-85) for all onsets of a
 * batch.  rel_dev [R, n_rows, C] is the relative envelope in time order.
 *   streaming = 0: rel_dev is the whole output of ofp_detect_offline, on_sample_dev holds absolute
 *                  samples and is rewritten in place;
 *   streaming = 1: rel_dev holds the last n_rows rows ending with the current block (the reference's
 *                  CircularArray), on_sample_dev holds deltas within that block.
 * alpha = float32(2/(smooth+1)), tol = float32((1-alpha)**buffer_size) as in detection.py:722-725. */
int ofp_backtrack_onsets(const float *rel_dev, int64_t rec_stride, int64_t n_rows, int32_t n_channels,
                         int32_t block_size, int32_t buffer_size, float alpha, float tol, int32_t streaming,
                         const int32_t *on_channel_dev, int32_t *on_sample_dev, const int32_t *on_count_dev,
                         int32_t n_rec, int32_t cap, void *stream);

/* Signature twins of the ctypes DLL (envelope_follower.c:6,27), device pointers.
 *   ar_envelope:  x,y [num_samples, size]; y's last row is the carried state on entry.
 *   minmax_envelope: x [n_samples, n_channels]; min/max [n_channels] in/out. */
int ofp_ar_envelope(const float *x_dev, float *y_dev, float attack, float release, int size,
                    int num_samples, void *stream);
int ofp_minmax_envelope(const float *x_dev, float *min_dev, float *max_dev, float alpha_min,
                        float alpha_max, float minmin, int n_samples, int n_channels, void *stream);

/* ---------------------------------------------------------------------------------------
 * K3  onset grouping -- replaces find_onset_groups (detection.py:131-189)
 * ------------------------------------------------------------------------------------- */

/* One sequential scan per recording over its onsets in detection order (the output of
 * ofp_detect_offline).  groups_dev [R, max_groups, C] int32 (-1 = channel missing),
 * n_groups_dev [R] (may exceed max_groups; only max_groups are stored).
 * close_channel < 0 disables the close-channel filter (detection.py:184-185). */
int ofp_group_onsets(const int32_t *on_channel_dev, const int32_t *on_sample_dev, const int32_t *on_count_dev,
                     int32_t n_rec, int32_t cap, int32_t n_channels, int32_t max_distance, int32_t min_channels,
                     int32_t close_channel, int32_t max_groups, int32_t *groups_dev, int32_t *n_groups_dev,
                     void *stream);
/* Flatten the per-recording groups into a hit list.  offsets_dev [R] int64 = exclusive prefix sum of
 * min(n_groups, max_groups); hit_rec_dev [H] int32, hit_onsets_dev [H, C] int32. */
int ofp_compact_groups(const int32_t *groups_dev, const int32_t *n_groups_dev, const int64_t *offsets_dev,
                       int32_t n_rec, int32_t max_groups, int32_t n_channels, int32_t *hit_rec_dev,
                       int32_t *hit_onsets_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * K4  lag refinement -- replaces fix_onsets (detection.py:373-451) and, inside it,
 *     cross_correlation_lag (195-268) and adjust_onset (299-352)
 * ------------------------------------------------------------------------------------- */

#define OFP_LAG_NONE INT32_MIN /* cross_correlation_lag returned None */
#define OFP_FIX_OK 0
#define OFP_FIX_DEGENERATE 1 /* section start < 0 or too short: the reference wraps / raises (Q6) */
#define OFP_FIX_REF_CRASH 2  /* the reference raises ValueError in adjust_onset (Q10); onsets as of the failing pair */
#define OFP_FIX_TOO_LONG 3   /* section longer than max_section */
#define OFP_FIX_INCOMPLETE 4 /* a channel of the group is missing (-1) */

/* fix_onsets for n_hits onset groups in one launch.
 *   audio_dev [R, n_samples, C] float32 (recording stride rec_stride elements)
 *   hit_rec_dev [H] int32 recording of each hit, or NULL (hit h lives in recording h)
 *   onsets_dev [H, C] int32 sample index per channel
 *   filter_size/d/direction(0 none,1 "up",2 "down")/take_abs/zero_left/cutoff/tol/shift: the keyword
 *   arguments of fix_onsets; max_section: upper bound on (span + 2*(cutoff+tol)) in samples
 *   out_onsets_dev [H, C]; out_lags_dev [H, C] or NULL (lag returned by cross_correlation_lag per
 *   later channel, OFP_LAG_NONE elsewhere); out_status_dev [H] OFP_FIX_*. */
int ofp_fix_onsets(const float *audio_dev, int64_t n_samples, int64_t rec_stride, int32_t n_channels,
                   const int32_t *hit_rec_dev, const int32_t *onsets_dev, int32_t n_hits, int32_t filter_size,
                   int32_t d, int32_t direction, int32_t take_abs, int32_t zero_left, int32_t cutoff, int32_t tol,
                   int32_t shift, int32_t max_section, int32_t *out_onsets_dev, int32_t *out_lags_dev,
                   int32_t *out_status_dev, void *stream);
/* Same with flags: bit 0 = every section runs to the END of its recording instead of stopping
 * lookaround samples after the last onset -- the section Multilaterate3D.locate cuts from its ring
 * buffer (multilateration.py:457-466: rec_audio[-i - 1:]). */
int ofp_fix_onsets_ex(const float *audio_dev, int64_t n_samples, int64_t rec_stride, int32_t n_channels,
                      const int32_t *hit_rec_dev, const int32_t *onsets_dev, int32_t n_hits, int32_t filter_size,
                      int32_t d, int32_t direction, int32_t take_abs, int32_t zero_left, int32_t cutoff,
                      int32_t tol, int32_t shift, int32_t max_section, int32_t flags, int32_t *out_onsets_dev,
                      int32_t *out_lags_dev, int32_t *out_status_dev, void *stream);
int ofp_fix_onsets_smem_bytes(int32_t n_channels, int32_t max_section);
/* Statistics of the float32 lag-window screening inside K4 (csrc/lag_fix.cu:cc_argmax_screen) since
 * the last reset (synchronises the device): stats4_host[0] pairs screened, [1] decided by a single
 * surviving lag, [2] decided by exact recomputation of <= 32 survivors, [3] handed to the exact
 * all-lags path.  The result is identical in every case; this only tells where the time goes. */
int ofp_cc_screen_stats(uint64_t *stats4_host, int32_t reset);

/* cross_correlation_lag for n_pairs pairs of equal length n: x_dev, y_dev [P, n] float32;
 * onsets_or_legal_dev [P, 2] = (onset_x, onset_y) or (legal_lo, legal_hi); lag_dev [P] (OFP_LAG_NONE = None). */
int ofp_cross_correlation_lag(const float *x_dev, const float *y_dev, int32_t n_pairs, int32_t n, int32_t d,
                              int32_t take_abs, int32_t use_legal_lags, int32_t cutoff, int32_t tol,
                              const int32_t *onsets_or_legal_dev, int32_t *lag_dev, void *stream);
/* adjust_onset: out_dev [P, 2] = (change of onset_x, change of onset_y); both OFP_LAG_NONE where the
 * reference raises. */
int ofp_adjust_onset(const float *x_dev, const float *y_dev, int32_t n_pairs, int32_t n, const int32_t *onsets_dev,
                     const int32_t *new_lag_dev, int32_t *out_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * K5  TDOA multilateration -- replaces Multilaterate3D.is_legal / is_legal_3d / trilaterate and
 *     solve_trilateration_3d (multilateration.py:230-316, 397-426, 536-566); per-hit batch form
 * ------------------------------------------------------------------------------------- */

/* status per hit: 0 located, 1 lag beyond max_max_lags, 2 is_legal failed, 3 no seed cell,
 * 4 solver did not converge (fsolve ier != 1), 5 invalid sensors.
 *   sensor_xyz_dev [S, 3] float64 cm; lag_maps_dev [S, S, M, M] float32 (NaN = illegal), entry
 *   [i][j] = Multilaterate3D.lag_maps[i][j]; max/min_lags_dev [S, S] float32; max_max_dev [S];
 *   hit_sensors_dev [H, 3] int32 or NULL (sensors 0,1,2); hit_onsets_dev [H, onset_stride] int32,
 *   the first three entries of a row are used; xy_dev [H, 2] float64 (NaN when not located). */
int ofp_locate_hits(const double *sensor_xyz_dev, int32_t n_sensors, const float *lag_maps_dev,
                    int32_t map_size, const float *max_lags_dev, const float *min_lags_dev,
                    const float *max_max_dev, double radius_cm, double samples_per_cm, double sr,
                    double c_cm_s, const int32_t *hit_sensors_dev, const int32_t *hit_onsets_dev,
                    int32_t onset_stride, int32_t n_hits, double *xy_dev, int32_t *status_dev, void *stream);

/* Model bypass of trilaterate (multilateration.py:553-557: res = self.model.call_np((d_a1, d_b1)) * 100).
 * ofp_locate_hits_lags runs the same legality / seed-cell checks as ofp_locate_hits but, instead of solving,
 * writes the two lags the reference hands to the model (after its sensor rewrite, Q8) to pair_lags_dev [H, 2]
 * float32; status as above (never 4), xy_dev is left NaN.  ofp_fcnn_forward evaluates calibration.FCNN
 * (calibration.py:463-560) in eval mode on rows x_dev [n_rows, widths[0]]: n_layers Linear layers of widths
 * widths_host[0..n_layers] (<= 32), each hidden one followed by the BatchNorm1d inference affine and the
 * activation (0 ReLU, 1 tanh, 2 sigmoid, 3 SiLU, 4 identity).  params_dev per layer: W [out][in], b [out],
 * scale [out], shift [out] (scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale;
 * 1 and 0 without batch norm).  Rows with status_dev[r] != 0 are skipped (status_dev may be NULL).  Outputs are
 * multiplied by out_scale and written to out_f32_dev and/or out_f64_dev [n_rows, widths[n_layers]]. */
int ofp_locate_hits_lags(const double *sensor_xyz_dev, int32_t n_sensors, const float *lag_maps_dev,
                         int32_t map_size, const float *max_lags_dev, const float *min_lags_dev,
                         const float *max_max_dev, double radius_cm, double samples_per_cm, double sr,
                         double c_cm_s, const int32_t *hit_sensors_dev, const int32_t *hit_onsets_dev,
                         int32_t onset_stride, int32_t n_hits, float *pair_lags_dev, double *xy_dev,
                         int32_t *status_dev, void *stream);
int ofp_fcnn_forward(const float *x_dev, int64_t n_rows, int32_t n_layers, const int32_t *widths_host,
                     int32_t activation, const float *params_dev, const int32_t *status_dev, float out_scale,
                     float *out_f32_dev, double *out_f64_dev, void *stream);

/* Streaming locate for n_streams concurrent realtime streams: what PlayRec.detect_hits does per block
 * (realtime/audio.py:62-74) with Multilaterate3D.locate's group state machine
 * (multilateration.py:428-534, rec_audio = None) kept per stream in device memory.
 *   geometry arguments as in ofp_locate_hits; det_*_dev = ofp_detect_block's output for this block
 *   ([S, C], [S, C], [S]); current_index = sample index of the block start;
 *   state_*_dev: the `ongoing` lists (sizes from ofp_stream_locate_state_bytes, zero-initialised by
 *   the caller = empty lists); xy_dev [S, 2] (NaN when nothing located), found_dev [S]: 1 located,
 *   0 nothing, -1 a group list or group outgrew its slot (16 groups x 4 members) and was truncated. */
int ofp_stream_locate_state_bytes(int32_t n_streams, int64_t *count_bytes, int64_t *len_bytes,
                                  int64_t *sensor_bytes, int64_t *onset_bytes);
int ofp_stream_locate(const double *sensor_xyz_dev, int32_t n_sensors, const float *lag_maps_dev,
                      int32_t map_size, const float *max_lags_dev, const float *min_lags_dev,
                      const float *max_max_dev, double radius_cm, double samples_per_cm, double sr,
                      double c_cm_s, int32_t n_streams, int32_t n_channels, const int32_t *det_channel_dev,
                      const int32_t *det_delta_dev, const int32_t *det_count_dev, int64_t current_index,
                      int32_t *state_count_dev, int32_t *state_len_dev, int32_t *state_sensor_dev,
                      int64_t *state_onset_dev, double *xy_dev, int32_t *found_dev, void *stream);

/* solve_trilateration / solve_trilateration_3d (multilateration.py:170-316) with explicit seeds:
 * scipy.optimize.fsolve(xtol, maxfev, fprime) = MINPACK hybrj for n = 2, one problem per row.
 *   problems_dev [P, 11] float64 = sensor_a xyz, sensor_b xyz, sensor_origin xyz, delta_d_a, delta_d_b
 *   (z = 0 for the 2-D form); seeds_dev [P, 2]; xy_dev [P, 2] = fsolve's root (also when not
 *   converged); ier_dev [P] (1 = converged, the reference returns None otherwise); nfev_dev [P] or NULL. */
int ofp_solve_trilateration(const double *problems_dev, const double *seeds_dev, int32_t n_problems, double xtol,
                            int32_t maxfev, double *xy_dev, int32_t *ier_dev, int32_t *nfev_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * Helper twins around the detector / lag path (csrc/onset_tools.cu)
 * ------------------------------------------------------------------------------------- */

/* ButterworthFilter.__call__ (detection.py:487-501) = scipy.signal.lfilter(b, a, x, axis=0, zi): direct
 * form II transposed in float32.  b_host, a_host [order+1] float32 (a[0] == 1); x_dev, y_dev
 * [n_samples, n_channels]; zi_dev [order, n_channels] carried state, updated in place. */
int ofp_lfilter(const float *b_host, const float *a_host, int32_t order, const float *x_dev, float *y_dev,
                float *zi_dev, int32_t n_samples, int32_t n_channels, void *stream);
/* filter_data (detection.py:355-370): direction 1 = "up" (zero where the first difference is negative),
 * 2 = "down"; x_dev, out_dev [n_samples, n_channels], out of place. */
int ofp_filter_data(const float *x_dev, float *out_dev, int64_t n_samples, int32_t n_channels, int32_t direction,
                    void *stream);
/* detect_onset_region (detection.py:454-484) for n_signals rows of audio_dev [n_signals, len]:
 * out_dev [n_signals] = start of the loud region around onsets_dev[i]. */
int ofp_detect_onset_region(const float *audio_dev, int32_t n_signals, int64_t len, const int32_t *onsets_dev,
                            int32_t n, int32_t median_filter_size, float threshold_factor, int32_t *out_dev,
                            void *stream);
/* Peak refinement of the dataset builder (notebooks/refresh.org:262-279): out[h, c] = onsets[h, c] +
 * argmax(audio[onsets[h, c] : onsets[h, c] + tolerance, c]) (first maximum; -1 = missing channel stays -1).
 * audio_dev [R, n_samples, C], hit h in recording hit_rec_dev[h] (NULL: hit h in recording h). */
int ofp_window_argmax(const float *audio_dev, int64_t n_samples, int64_t rec_stride, int32_t n_channels,
                      const int32_t *hit_rec_dev, const int32_t *onsets_dev, int32_t n_hits, int32_t tolerance,
                      int32_t *out_dev, void *stream);
/* RecAnalysis.tempogram (realtime/recording.py:313-327) for the frames first_frame + k*every (k < n_selected) of
 * every row of oe_dev [n_rec, n_frames]: tg_dev [n_rec, n_selected, win_length] = autocorrelation of
 * window * oe[j - win_length + 1 .. j] (zeros before frame 0) divided by (its maximum + 1e-10). */
int ofp_tempogram(const float *oe_dev, int32_t n_rec, int64_t n_frames, const float *window_dev, int32_t win_length,
                  int64_t first_frame, int64_t every, int64_t n_selected, float *tg_dev, void *stream);
/* np.correlate(x, y, "full") as find_lag / find_lag_multi use it (multilateration.py:878-899):
 * x_dev, y_dev [P, n] -> out_dev [P, 2n-1] (double accumulation, rounded once). */
int ofp_correlate_full(const float *x_dev, const float *y_dev, int32_t n_pairs, int32_t n, float *out_dev,
                       void *stream);

/* Onset-window extraction -- FrameExtractor / FastFrameExtractor (data.py:55-192):
 * frames_dev [H, C, F] = sliding_window_view(audio, F, axis=0)[start], start = min over the hit's
 * onsets (use_min_onset) or each channel's own onset, minus (pre_samples - shift[h]); negative starts
 * wrap like numpy indexing, out-of-range ones set status_dev[h] = 1 (numpy raises IndexError).
 *   audio_dev [R, n_samples, C]; hit_rec_dev [H] or NULL (recording 0); onsets_dev [H, C];
 *   shifts_dev [H] or NULL (the random augmentation shift of max_shift). */
int ofp_extract_frames(const float *audio_dev, int64_t n_samples, int64_t rec_stride, int32_t n_channels,
                       const int32_t *hit_rec_dev, const int32_t *onsets_dev, const int32_t *shifts_dev,
                       int32_t n_hits, int32_t frame_length, int32_t pre_samples, int32_t use_min_onset,
                       float *frames_dev, int32_t *status_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * Streaming cross-correlation -- twin of the CPython extension online_cc
 *     (c/cross_corr.c:257-291: CrossCorrelation(n, block_size).update(a, b) -> float32[2n-1])
 * ------------------------------------------------------------------------------------- */
typedef struct ofp_ccstream ofp_ccstream;
/* n_pairs independent stream pairs; rings start at zero like the reference's calloc'd buffers. */
int ofp_ccstream_create(ofp_ccstream **out, int32_t n_pairs, int32_t n, int32_t block_size);
int ofp_ccstream_destroy(ofp_ccstream *h);
/* a_dev, b_dev [P, block_size] float32 -> out_dev [P, 2n-1]: np.correlate(last n of a, last n of b, "full"). */
int ofp_ccstream_update(ofp_ccstream *h, const float *a_dev, const float *b_dev, float *out_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * K2  spectral-flux onset features -- replaces RecAnalysis.fft / onset_strength
 *     (realtime/recording.py:273-311) and the STFT half of detect_onsets_spectral (detection.py:96-110)
 * ------------------------------------------------------------------------------------- */

/* x_dev [R, n_samples, C] float32 -> flux_dev [R, n_frames] float32.  Per frame: channel mean, window,
 * rFFT(n_fft), then mode 0: 10*log10(max(1e-10, |X|^2)) (optionally clamped at frame max - top_db),
 * mode 1: |X| * weight[bin]; flux = mean over the n_fft/2+1 bins of max(0, S_j - S_{j-1}).
 * center = 0: frame j = samples [(j+1)*hop - n_fft, (j+1)*hop) (n_frames = n_samples / hop);
 * center = 1: frame j centred on j*hop, padded with zeros (reflect = 0) or by reflection (1).
 * window_dev [n_fft], weight_dev [n_fft/2+1] or NULL. */
int ofp_spectral_flux(const float *x_dev, int64_t n_rec, int64_t n_samples, int64_t rec_stride, int32_t n_channels,
                      int32_t n_fft, int32_t hop, int32_t center, int32_t reflect, int32_t mode, float top_db,
                      const float *window_dev, const float *weight_dev, int32_t n_frames, float *flux_dev,
                      void *stream);
/* Complex short-time spectra of frames that are already cut -- replaces the per-frame `np.fft.rfft(window * x)` of
 * data.stft / data.stft_frame (data.py:581-654).  frames_dev [n_frames, frame_length] float32; every frame is centred
 * in n_fft points (librosa.util.pad_center, data.py:589-590), multiplied by window_dev [n_fft] float64 and transformed
 * in double; out_dev [n_frames, n_fft/2 + 1] complex64 (interleaved re, im).  n_fft: a power of two up to 4096. */
int ofp_stft_frames(const float *frames_dev, int64_t n_frames, int32_t frame_length, int32_t n_fft,
                    const double *window_dev, float *out_dev, void *stream);

/* librosa.util.peak_pick (detection.py:113-121) per row of oe_dev [R, n_frames]: peaks_dev [R, cap],
 * n_peaks_dev [R]. */
int ofp_peak_pick(const float *oe_dev, int32_t n_rec, int32_t n_frames, int32_t pre_max, int32_t post_max,
                  int32_t pre_avg, int32_t post_avg, float delta, int32_t wait, int32_t *peaks_dev,
                  int32_t *n_peaks_dev, int32_t cap, void *stream);

/* ---------------------------------------------------------------------------------------
 * Realtime session -- PlayRec's per-block callback (realtime/audio.py:62-122: detector -> locate per
 * 128-sample block) for n_streams concurrent streams as one replayed CUDA graph per block.
 * ------------------------------------------------------------------------------------- */

/* ofp_stream_locate with the block's start index kept in device memory (*current_index_dev, advanced by
 * `advance` samples after the block): the form a CUDA graph can replay. */
int ofp_stream_locate_dev(const double *sensor_xyz_dev, int32_t n_sensors, const float *lag_maps_dev,
                          int32_t map_size, const float *max_lags_dev, const float *min_lags_dev,
                          const float *max_max_dev, double radius_cm, double samples_per_cm, double sr,
                          double c_cm_s, int32_t n_streams, int32_t n_channels, const int32_t *det_channel_dev,
                          const int32_t *det_delta_dev, const int32_t *det_count_dev, int64_t *current_index_dev,
                          int32_t advance, int32_t *state_count_dev, int32_t *state_len_dev,
                          int32_t *state_sensor_dev, int64_t *state_onset_dev, double *xy_dev, int32_t *found_dev,
                          void *stream);

/* The same with the ring-buffer refinement of every new (group, detection) pair -- what PlayRec's callback runs
 * (realtime/audio.py:69 passes self.rec_audio; multilateration.py:457-501: median 5 -> diff -> falling flanks ->
 * cross_correlation_lag(tol, cutoff) -> adjust_onset).  ring_dev [n_streams, ring_rows, n_channels] float32 holds
 * the most recent rows of every stream (row = sample index % ring_rows, zeros before the stream started) INCLUDING
 * the current block: call ofp_ring_write first.  One warp per stream.  found_dev additionally reports -2 where the
 * reference would have raised inside adjust_onset. */
int ofp_stream_locate_ring_dev(const double *sensor_xyz_dev, int32_t n_sensors, const float *lag_maps_dev,
                               int32_t map_size, const float *max_lags_dev, const float *min_lags_dev,
                               const float *max_max_dev, double radius_cm, double samples_per_cm, double sr,
                               double c_cm_s, int32_t n_streams, int32_t n_channels, const int32_t *det_channel_dev,
                               const int32_t *det_delta_dev, const int32_t *det_count_dev, int64_t *current_index_dev,
                               int32_t advance, const float *ring_dev, int32_t ring_rows, int32_t block_size,
                               int32_t onset_tolerance, int32_t normalization_cutoff, int32_t *state_count_dev,
                               int32_t *state_len_dev, int32_t *state_sensor_dev, int64_t *state_onset_dev,
                               double *xy_dev, int32_t *found_dev, void *stream);
/* CircularArray.write for every stream (loopmate; realtime/audio.py callback): blocks_dev [S, B, C] (stream s at
 * blocks_dev + s*stream_stride elements, 0 = dense) into ring rows (*current_index_dev + row) % ring_rows. */
int ofp_ring_write(float *ring_dev, int32_t ring_rows, const float *blocks_dev, int64_t stream_stride,
                   int32_t n_streams, int32_t block_size, int32_t n_channels, const int64_t *current_index_dev,
                   void *stream);

typedef struct ofp_rt ofp_rt;
/* The session owns the detector, the locate state, a staging buffer and pinned result buffers; the geometry
 * arrays (as in ofp_locate_hits) stay owned by the caller and must outlive it.  use_graph = 0 issues the same
 * launches eagerly (A/B).  ring_rows > 0: the session also keeps a ring of the last ring_rows audio rows per stream
 * and runs the ring-buffer refinement inside locate (rec_audio passed, as PlayRec does); 0 = locate(rec_audio=None). */
int ofp_rt_create(ofp_rt **out, int32_t n_streams, const ofp_detector_params *p, const double *sensor_xyz_dev,
                  int32_t n_sensors, const float *lag_maps_dev, int32_t map_size, const float *max_lags_dev,
                  const float *min_lags_dev, const float *max_max_dev, double radius_cm, double samples_per_cm,
                  double sr, double c_cm_s, int32_t use_graph, int32_t ring_rows);
int ofp_rt_reset(ofp_rt *rt);
int ofp_rt_destroy(ofp_rt *rt);
/* One block for every stream: blocks [S, B, C] float32 in host (blocks_on_host = 1, pinned for an asynchronous
 * copy) or device memory, stream s at blocks + s*stream_stride elements (0 = dense).  Returns when the results
 * are on the host: xy_host [S, 2] float64 (NaN = nothing located), found_host [S] (1 located, 0 nothing,
 * -1 a group list overflowed); either may be NULL (read them through ofp_rt_results instead). */
int ofp_rt_step(ofp_rt *rt, const float *blocks, int32_t blocks_on_host, int64_t stream_stride, double *xy_host,
                int32_t *found_host);
/* Device-resident blocks are read on the session's private stream: call this first with the stream that
 * produced them (e.g. torch's current stream) so that the copy is ordered after the producer. */
int ofp_rt_wait_stream(ofp_rt *rt, void *producer_stream);
int ofp_rt_results(ofp_rt *rt, const double **xy_host, const int32_t **found_host);

/* ---------------------------------------------------------------------------------------
 * K6  onset-window network inference -- replaces model.CNN.forward in eval mode (model.py:52-120):
 *     n_layers x [Conv1d(kernel_size, padding, dilation, groups; stride 1) + activation (+ BatchNorm1d) (+ MaxPool1d(2))],
 *     flatten (channel-major), Dropout = identity, Linear(flat, out_size).
 * ------------------------------------------------------------------------------------- */

/* Size of the packed parameter buffer and of the flattened feature vector for an architecture. */
int ofp_cnn_param_count(int32_t channels, int32_t input_size, int32_t n_layers, const int32_t *layer_sizes_host,
                        int32_t kernel_size, int32_t padding, int32_t out_size, int64_t *n_params_out,
                        int32_t *flat_out);
/* x_dev [n_windows, channels, input_size] float32 (window w at x_dev + w*win_stride elements, the layout
 * FrameExtractor returns, data.py:55-120) -> out_dev [n_windows, out_size] float32.
 * params_dev, packed float32: per conv layer l the weight transposed to [c_in][k][c_out_padded] followed by
 * the bias [c_out_padded] (c_out padded to a multiple of 8 with zeros), then fc.weight [out_size][flat]
 * and fc.bias [out_size].  activation: 0 SiLU (the reference default), 1 ReLU, 2 tanh, 3 identity. */
int ofp_cnn_forward(const float *x_dev, int64_t n_windows, int64_t win_stride, int32_t channels, int32_t input_size,
                    int32_t n_layers, const int32_t *layer_sizes_host, int32_t kernel_size, int32_t padding,
                    int32_t activation, const float *params_dev, int32_t out_size, float *out_dev, void *stream);

/* The same network with the constructor options the reference's CNN also has (model.py:62-66, 91-108):
 * dilation >= 1; pool != 0: MaxPool1d(kernel_size=2, stride=2) after every layer's activation (and norm);
 * batch_norm != 0: eval-mode BatchNorm1d behind every activation, passed as scale = weight / sqrt(running_var + eps)
 * and shift = bias - running_mean * scale, [c_out_padded] floats each, packed right behind the layer's bias.
 * `groups` needs no argument: pack the grouped weight as a dense [c_in][k][c_out_padded] block with zeros between the
 * groups.  (dilation, pool, batch_norm) = (1, 0, 0) is ofp_cnn_forward / ofp_cnn_param_count. */
int ofp_cnn_param_count_ex(int32_t channels, int32_t input_size, int32_t n_layers, const int32_t *layer_sizes_host,
                           int32_t kernel_size, int32_t padding, int32_t dilation, int32_t pool, int32_t batch_norm,
                           int32_t out_size, int64_t *n_params_out, int32_t *flat_out);
int ofp_cnn_forward_ex(const float *x_dev, int64_t n_windows, int64_t win_stride, int32_t channels, int32_t input_size,
                       int32_t n_layers, const int32_t *layer_sizes_host, int32_t kernel_size, int32_t padding,
                       int32_t dilation, int32_t pool, int32_t batch_norm, int32_t activation, const float *params_dev,
                       int32_t out_size, float *out_dev, void *stream);

/* model.CCCNN.forward in eval mode (model.py:443-538): the conv stack (as above, but with ONE input
 * channel) runs on every sensor channel separately, the K feature maps of a channel are auto-correlated over all
 * 2V-1 lags and summed, soft-maxed over the lags, and the [channels x (2V-1)] probabilities feed Linear.
 * x_dev [n_windows, channels, input_size]; params_dev: conv layers packed as for ofp_cnn_forward (first layer
 * c_in = 1), then fc.weight [out_size][channels * (2V-1)] and fc.bias.  out_size <= 4, the last layer size a
 * multiple of 8 and V a multiple of 16 (the auto-correlation runs as F^T F on the tensor cores).
 * group != 0 is CCCNN(group=True) (model.py:470-484: every conv layer with groups = channels, i.e. each sensor channel
 * runs its own copy of the stack): the packed conv parameters are then `channels` blocks of the layout above, channel
 * 0 first, followed by fc. */
int ofp_cccnn_param_count(int32_t channels, int32_t input_size, int32_t n_layers, const int32_t *layer_sizes_host,
                          int32_t kernel_size, int32_t padding, int32_t out_size, int32_t group, int64_t *n_params_out,
                          int32_t *n_lags_out);
int ofp_cccnn_forward(const float *x_dev, int64_t n_windows, int64_t win_stride, int32_t channels, int32_t input_size,
                      int32_t n_layers, const int32_t *layer_sizes_host, int32_t kernel_size, int32_t padding,
                      int32_t activation, int32_t group, const float *params_dev, int32_t out_size, float *out_dev,
                      void *stream);

/* The same network with the constructor options the reference's CCCNN leaves off by default (model.py:451-457,
 * 485-503): one kernel size and one stride per layer, dilation, pool != 0: MaxPool1d(2, 2) after every layer,
 * group_norm != 0: the GroupNorm(1, K) that `batch_norm=True` builds (model.py:494-498; statistics over the K x L values
 * of one sensor channel's feature maps, eps 1e-5), packed as gamma[c_out_padded] + beta[c_out_padded] right behind each
 * layer's bias.  group together with group_norm is refused (that norm spans all sensor channels of a window).  The
 * last layer size must be a multiple of 8 and the final feature-map length a multiple of 16, as above. */
int ofp_cccnn_param_count_ex(int32_t channels, int32_t input_size, int32_t n_layers, const int32_t *layer_sizes_host,
                             const int32_t *kernel_sizes_host, const int32_t *strides_host, int32_t padding,
                             int32_t dilation, int32_t pool, int32_t group_norm, int32_t out_size, int32_t group,
                             int64_t *n_params_out, int32_t *n_lags_out);
int ofp_cccnn_forward_ex(const float *x_dev, int64_t n_windows, int64_t win_stride, int32_t channels, int32_t input_size,
                         int32_t n_layers, const int32_t *layer_sizes_host, const int32_t *kernel_sizes_host,
                         const int32_t *strides_host, int32_t padding, int32_t dilation, int32_t pool, int32_t group_norm,
                         int32_t activation, int32_t group, const float *params_dev, int32_t out_size, float *out_dev,
                         void *stream);

/* ---------------------------------------------------------------------------------------
 * Benchmark input: seeded synthetic multi-mic drum audio generated on the device
 * (SURVEY.md section 8d signal model; not a reference function).  x_dev [R, N, C] float32;
 * sensors_xyz_host [C, 3] cm (host); rec_offset = global index of recording 0 of this shard;
 * a burst is only emitted if it starts at least tail_guard samples before the end.
 * ------------------------------------------------------------------------------------- */
int ofp_synth_drum(float *x_dev, int64_t n_rec, int64_t n_samples, int32_t n_channels,
                   const float *sensors_xyz_host, float c_cm_s, float sr, float noise, float radius_cm,
                   int64_t first_hit, int64_t hit_period, int32_t tail_guard, uint64_t seed, int64_t rec_offset,
                   void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OFP_H */
