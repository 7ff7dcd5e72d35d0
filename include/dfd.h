/* libdfd -- C-ABI of the B200-native per-frame deepfake-detection hot path.
 *
 * The reference (KrishTanna28/Real-Time-Video-Deepfake-Detection) is pure
 * Python and has no FFI; this header is the boundary a maintainer binds with
 * ctypes to replace, one for one, the Python call sites
 * cited on each entry.  All bulk pointers are DEVICE pointers unless marked
 * HOST; the caller owns every buffer passed in or out; the context owns the
 * weights, the per-stream state and its workspaces.  Every entry returns 0 on
 * success or a negative dfd_status and never throws; dfd_last_error() returns
 * the message.  Calls are asynchronous on the supplied CUDA stream
 * (cudaStream_t passed as void*; NULL = legacy default stream).  A context is
 * bound to one GPU and is not thread-safe (one host thread per context, like
 * the reference's effectively serial detector).  There is no CPU fallback.
 */
#ifndef DFD_H
#define DFD_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFD_ABI_VERSION 2   /* 2: + dfd_face_prep_tta, dfd_face_probability_tta, dfd_set_calibrator, dfd_draw_overlay (additions only) */
#define DFD_TILE 256          /* forensic analysis size, frame_analysis.py:28 */
#define DFD_N_RAW 16          /* raw statistics per frame, order = oracle/forensics.py RAW_NAMES */
#define DFD_N_SIGNALS 6       /* frequency, noise, ela, edge, color, temporal */
#define DFD_CROP 224          /* classifier input size, deepfake_detection.py:383 */

typedef enum {
    DFD_OK = 0,
    DFD_ERR_INVALID = -1,     /* bad argument */
    DFD_ERR_CUDA = -2,        /* CUDA runtime/driver error */
    DFD_ERR_NO_WEIGHTS = -3,  /* classifier called before dfd_load_weights */
    DFD_ERR_CAPACITY = -4,    /* batch / stream id / crop size beyond the configured capacity */
    DFD_ERR_ARCH = -5,        /* device is not sm_100 */
    DFD_ERR_UNSUPPORTED = -6  /* input format outside the supported subset (e.g. a progressive JPEG) */
} dfd_status;

typedef enum { DFD_F32 = 0, DFD_BF16 = 1 } dfd_dtype;
typedef enum { DFD_UNCERTAIN = 0, DFD_REAL = 1, DFD_FAKE = 2 } dfd_verdict;
/* vote input policy: reference code feeds the face probability alone when a
 * face exists (deepfake_detection.py:620-626, backend_server.py:167-171); the
 * README's 70/30 blend is opt-in. */
typedef enum { DFD_BLEND_REFERENCE = 0, DFD_BLEND_README = 1 } dfd_blend_mode;

typedef struct dfd_ctx dfd_ctx;

typedef struct {
    int32_t device;               /* CUDA ordinal */
    int32_t max_streams;          /* per-stream state slots (stream ids 0..max_streams-1) */
    int32_t max_batch;            /* max frames and max face boxes per call */
    int32_t max_crop;             /* max face-box side in pixels */
    int32_t window_size;          /* TemporalTracker(window_size=60)        deepfake_detection.py:99 */
    int32_t voting_window;        /* TemporalTracker(voting_window=10) */
    double detection_threshold;   /* DeepfakeDetector(detection_threshold)  deepfake_detection.py:300 */
    double face_weight;           /* 0.70, used only with DFD_BLEND_README */
    double forensic_weight;       /* 0.30 */
    int32_t blend_mode;           /* dfd_blend_mode */
    int32_t reserved;
} dfd_config;

/* One frame's forensic result; replaces the dict returned by
 * FrameForensicAnalyzer.analyze / analyze_fast (frame_analysis.py:58-126). */
typedef struct {
    double raw[DFD_N_RAW];        /* NaN where the statistic was not computed */
    double scores[DFD_N_SIGNALS]; /* NaN for signals skipped by analyze_fast */
    double fake_probability;
    int32_t frame_number;         /* analyzer.frame_count after this frame */
    int32_t full;                 /* 1 = analyze, 0 = analyze_fast */
} dfd_forensic_result;

/* One stream's vote record; replaces TemporalTracker.update + get_confidence_level +
 * get_voting_stats + get_temporal_average + get_stability_score
 * (deepfake_detection.py:120-268).  72 bytes, gathered across GPUs as is. */
typedef struct {
    int32_t stream_id;
    int32_t verdict;              /* dfd_verdict */
    int32_t fake_count, real_count;
    int32_t history_len;          /* len(score_history) */
    int32_t frame_count;          /* detector.frame_count of the stream */
    int32_t last_vote;            /* this frame's classification: 1 FAKE, 0 REAL, -1 none (update(None)) */
    int32_t reserved;
    double vote_input;            /* NaN = nothing fed this frame */
    double temporal_average;
    double stability_score;
    double face_probability;      /* NaN if no face */
    double forensic_probability;
} dfd_vote_record;

void dfd_default_config(dfd_config* cfg);
int dfd_abi_version(void);
int dfd_create(const dfd_config* cfg, dfd_ctx** out);
void dfd_destroy(dfd_ctx* ctx);
const char* dfd_last_error(dfd_ctx* ctx);   /* ctx may be NULL for dfd_create failures */

/* Classifier weights.  HOST blob of float32: BN-folded parameters in the order
 * produced by dfd_b200.weights.pack_state_dict() from the reference
 * checkpoint's `net.*` state_dict (model.py:36-61, deepfake_detection.py:44-51).
 * n_floats must equal dfd_weights_blob_floats(). */
size_t dfd_weights_blob_floats(void);
int dfd_load_weights(dfd_ctx* ctx, const float* blob_host, size_t n_floats);

/* FrameForensicAnalyzer.analyze / analyze_fast for a batch of frames, one per
 * stream (frame_analysis.py:58-126; cadence chosen by the caller as in
 * deepfake_detection.py:504-515).  frames: n images of H x W BGR u8, image i at
 * frames + i*frame_stride, rows row_pitch bytes apart.  stream_ids[n], full[n]
 * (1 = analyze, 0 = analyze_fast) are device arrays.  A stream id may appear at
 * most once per call.  Preconditions checked on the device: a stream id outside [0, max_streams) touches no
 * state and yields a record with frame_number = -1 and fake_probability = NaN (vote record: verdict = -1). */
int dfd_forensics_batch(dfd_ctx* ctx, const uint8_t* frames, int n, int H, int W, size_t frame_stride,
                        int row_pitch, const int32_t* stream_ids, const uint8_t* full,
                        dfd_forensic_result* results, void* stream);

/* preprocess_face_quality + _single_prediction preprocessing (deepfake_detection.py:357-389)
 * for m face boxes: crop -> LAB CLAHE -> RGB -> PIL-bilinear 160^2 -> bilinear 224^2 -> /255 ->
 * ImageNet normalise.  boxes[m*4] = x,y,w,h (device int32); frame_idx[m] selects the frame of each
 * box.  out: m x 224 x 224 x 3 (NHWC) of dtype.
 * Boxes are clamped to the H x W frame exactly as the reference's numpy slicing frame[y:y+h, x:x+w] does
 * (deepfake_detection.py:612-619).  A box that is empty after clamping, whose frame index is outside
 * [0, n_frames) or whose clamped side exceeds max_crop is REJECTED: its output tensor is that of a dummy 8 x 8 crop and
 * dfd_face_probability / dfd_analyze_batch report its probability as NaN ("no face": analyze_face -> (None, None, None),
 * deepfake_detection.py:545-550).  Nothing is ever read outside the frames or written outside the workspaces. */
int dfd_face_prep_batch(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, size_t frame_stride,
                        int row_pitch, const int32_t* boxes, const int32_t* frame_idx, int m, void* out_nhwc,
                        int dtype, void* stream);

/* DeepfakeEfficientNet.forward (model.py:63-72): in m x 224 x 224 x 3 NHWC -> logits[m] (float32). */
int dfd_effnet_forward(dfd_ctx* ctx, const void* in_nhwc, int m, int dtype, float* logits, void* stream);

/* sigmoid + apply_heuristics (deepfake_detection.py:398,489-502): prob[i] = clip(sigmoid(logit) + 0.10*(w<80||h<80)).
 * When m equals the box count of the context's last dfd_face_prep_batch call, boxes that call rejected get NaN. */
int dfd_face_probability(dfd_ctx* ctx, const float* logits, const int32_t* boxes, int m, double* prob, void* stream);

/* Test-time augmentation, analyze_face_with_tta (deepfake_detection.py:408-443): n_pred predictions per face -- the
 * preprocessed (CLAHE'd) crop itself and n_pred - 1 augmented copies of it: cv2.flip(.., 1) when `flip`,
 * cv2.convertScaleAbs(alpha = brightness), cv2.warpAffine with cv2.getRotationMatrix2D((w/2, h/2), angle, 1.0) -- each then
 * resized / normalised like dfd_face_prep_batch's crops.  All pixel work is on the device and bit-exact with OpenCV; the
 * caller draws the random parameters (the reference uses Python's global `random`: see dfd_b200/tta.py, which consumes it in
 * the reference's order) and supplies, per augmentation, the INVERSE of the 2 x 3 rotation matrix (the inversion
 * cv2.warpAffine performs before its loops; a dozen double operations, done on the host so that cos / sin come from the same
 * libm as the reference's).  augs: DEVICE array of m * (n_pred - 1) records, box-major.  out: m * n_pred crops, crop
 * i * n_pred + j = prediction j of box i (j = 0 un-augmented).  Requires m * n_pred <= max_batch and n_pred <= DFD_TTA_MAX_PRED;
 * the matrices must be built from the box sizes AFTER clamping to the frame (what face_region.shape is in the reference). */
#define DFD_TTA_MAX_PRED 16
typedef struct {
    int32_t flip;                 /* 1 = cv2.flip(aug, 1)                              deepfake_detection.py:422-423 */
    float brightness;             /* float32(random.uniform(0.9, 1.1))                 :426-427 */
    double im[6];                 /* inverse of getRotationMatrix2D((w/2, h/2), angle) :430-433 */
} dfd_tta_aug;
int dfd_face_prep_tta(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, size_t frame_stride, int row_pitch,
                      const int32_t* boxes, const int32_t* frame_idx, int m, int n_pred, const dfd_tta_aug* augs,
                      void* out_nhwc, int dtype, void* stream);
/* np.mean of the n_pred sigmoids of each box (float64, NumPy's summation order) + apply_calibration + apply_heuristics
 * (deepfake_detection.py:441, 445-455, 489-502): logits[m * n_pred] -> prob[m]. */
int dfd_face_probability_tta(dfd_ctx* ctx, const float* logits, const int32_t* boxes, int m, int n_pred, double* prob,
                             void* stream);

/* apply_calibration (deepfake_detection.py:445-455; weights/calibrator.pkl, not shipped): the probability map applied
 * between the classifier and the heuristics by dfd_face_probability(_tta) and dfd_analyze_batch.
 *   DFD_CALIB_NONE              identity (the reference without calibrator.pkl)
 *   DFD_CALIB_LOGISTIC          n = 1: expit(xs[0] * p + ys[0])  (Platt scaling / sklearn LogisticRegression on the raw probability)
 *   DFD_CALIB_PIECEWISE_LINEAR  n knots (xs strictly increasing): np.interp(p, xs, ys)  (sklearn IsotonicRegression)
 * xs / ys are HOST arrays, copied before the call returns. */
typedef enum { DFD_CALIB_NONE = 0, DFD_CALIB_LOGISTIC = 1, DFD_CALIB_PIECEWISE_LINEAR = 2 } dfd_calibrator_kind;
int dfd_set_calibrator(dfd_ctx* ctx, int kind, int n, const double* xs_host, const double* ys_host, void* stream);

/* TemporalTracker.update for n streams (deepfake_detection.py:120-196).  vote_input[i] NaN = update(None).
 * np_flags (nullable, device u8[n]): 1 where the reference's value would be a numpy scalar (the face
 * probability returned by np.clip, deepfake_detection.py:502) rather than a Python float; it selects how
 * Python's sum() rounds temporal_average / stability_score (compensated vs plain) and nothing else. */
int dfd_vote_update(dfd_ctx* ctx, const int32_t* stream_ids, const double* vote_input, const uint8_t* np_flags, int n,
                    dfd_vote_record* records, void* stream);

/* The whole per-frame path for a batch: forensics + face prep + classifier + probability + vote input
 * selection + vote (backend_server.py:148-174).  One frame per stream; box_frame[m] maps each box to its
 * frame (only the first box of a frame feeds the vote, like faces[0]); full[n] as above.
 * records[n] (device) receives one record per frame, in frame order. */
int dfd_analyze_batch(dfd_ctx* ctx, const uint8_t* frames, int n, int H, int W, size_t frame_stride, int row_pitch,
                      const int32_t* stream_ids, const uint8_t* full, const int32_t* boxes, const int32_t* box_frame,
                      int m, int dtype, dfd_forensic_result* forensic_out /* may be NULL */,
                      double* face_prob_out /* m, may be NULL */, dfd_vote_record* records, void* stream);

/* Frame ingest: cv2.imdecode(np.frombuffer(image_bytes), cv2.IMREAD_COLOR) (backend_server.py:140-142) for n frames of the
 * /analyze wire format (JPEG, canvas.toDataURL('image/jpeg', 0.85), extension/content.js:106), decoded ON THE DEVICE so that
 * only the compressed stream crosses PCIe.  Bit-exact with OpenCV's libjpeg-turbo decoder.
 *   bytes_host   HOST: the n streams back to back (pinned memory makes the copy asynchronous)
 *   offsets_host HOST: n + 1 byte offsets into bytes_host (stream i = [offsets[i], offsets[i+1]))
 *   H, W         every stream must decode to H x W (use dfd_jpeg_info to peek); frames_out: n frames of H x W BGR u8 (DEVICE),
 *                frame i at frames_out + i * frame_stride, rows row_pitch bytes apart -- the layout dfd_analyze_batch takes
 *   status_dev   DEVICE int32[n]: 0, or -4 when a stream's entropy-coded data is corrupt (its frame content is undefined)
 * Supported: baseline sequential Huffman JPEG, 8 bit, gray or YCbCr 4:4:4 / 4:2:2 / 4:2:0, one interleaved scan, no restart
 * markers (what browsers and cv2.imencode emit).  Anything else returns DFD_ERR_UNSUPPORTED (progressive, arithmetic, 12-bit,
 * CMYK, restart intervals) or DFD_ERR_INVALID (not a JPEG, wrong size) and decodes nothing -- there is no CPU fallback. */
int dfd_decode_jpeg_batch(dfd_ctx* ctx, const uint8_t* bytes_host, const int64_t* offsets_host, int n, int H, int W,
                          uint8_t* frames_out, size_t frame_stride, int row_pitch, int32_t* status_dev, void* stream);
/* HOST-only header peek (no context): info = {H, W, components, luma h sampling, luma v sampling}; returns 0 if the stream is
 * decodable by dfd_decode_jpeg_batch, DFD_ERR_UNSUPPORTED / DFD_ERR_INVALID otherwise (info is filled when SOF was seen). */
int dfd_jpeg_info(const uint8_t* bytes_host, size_t n_bytes, int32_t* info);

/* Result annotation: draw_detection_overlay / _draw_frame_analysis_overlay (deepfake_detection.py:559-586, 688-726) on the
 * DEVICE copy of one frame, in place, bit-exact with the OpenCV calls the reference makes.  cmds_host[n_cmds] (HOST, applied in
 * order like successive OpenCV calls; n_cmds <= DFD_DRAW_MAX_CMDS):
 *   DFD_DRAW_OUTLINE  cv2.rectangle(img, (x0,y0), (x1,y1), color, thickness >= 1)
 *   DFD_DRAW_FILL     cv2.rectangle(.., thickness = -1)
 *   DFD_DRAW_BLEND    overlay = img.copy(); cv2.rectangle(overlay, (x0,y0), (x1,y1), color, -1);
 *                     cv2.addWeighted(overlay, alpha, img, beta, 0, img)
 *   DFD_DRAW_MASK     cv2.putText: masks_host[mask_off .. + mask_w*mask_h) is the string's stroke mask (nonzero = ink), its
 *                     top-left corner at (x0, y0); ink pixels take `color`.  The caller rasterises the string with OpenCV's
 *                     Hershey font (dfd_b200/overlay.py): glyph outlines are font data, every frame pixel is written here.
 * color is B, G, R.  Coordinates may lie outside the frame (clipped like OpenCV).  The host arrays are free on return. */
#define DFD_DRAW_MAX_CMDS 64
typedef enum { DFD_DRAW_OUTLINE = 1, DFD_DRAW_FILL = 2, DFD_DRAW_BLEND = 3, DFD_DRAW_MASK = 4 } dfd_draw_op;
typedef struct {
    int32_t op, x0, y0, x1, y1, thickness;
    uint8_t color[4];             /* B, G, R, unused */
    float alpha, beta;            /* DFD_DRAW_BLEND only */
    int32_t mask_off, mask_w, mask_h, reserved;
} dfd_draw_cmd;
int dfd_draw_overlay(dfd_ctx* ctx, uint8_t* frame, int H, int W, int row_pitch, const dfd_draw_cmd* cmds_host, int n_cmds,
                     const uint8_t* masks_host, size_t mask_bytes, void* stream);

/* DeepfakeDetector.reset / FrameForensicAnalyzer.reset / TemporalTracker.reset
 * (deepfake_detection.py:344-355, 270-289; frame_analysis.py:391-395).  stream_id < 0 resets all. */
int dfd_reset_stream(dfd_ctx* ctx, int stream_id, void* stream);
/* what: 1 = FrameForensicAnalyzer state only, 2 = TemporalTracker state + frame_count only, 3 = both. */
int dfd_reset_stream_part(dfd_ctx* ctx, int stream_id, int what, void* stream);

/* Per-stream tracker parameters (TemporalTracker(window_size, voting_window, detection_threshold),
 * deepfake_detection.py:99); streams start with the context defaults.  Resets the stream's vote state. */
int dfd_configure_stream(dfd_ctx* ctx, int stream_id, int window_size, int voting_window, double detection_threshold,
                         void* stream);

/* Hang diagnosis: with DFD_FLIGHT=1 in the environment every launch is followed by a stream-ordered marker written to
 * mapped host memory; the report names the last completed kernel and the pending ones (HOST buffer). */
int dfd_flight_report(dfd_ctx* ctx, char* buf_host, size_t buf_bytes);

/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
int64_t dfd_launch_count(dfd_ctx* ctx);

/* Per-kernel device timing for bench.py's roofline: between start and stop an event is recorded after
 * every launch; stop writes "kernel:label,launches,total_ms" lines into buf (HOST). */
int dfd_profile_start(dfd_ctx* ctx, void* stream);
int dfd_profile_stop(dfd_ctx* ctx, char* buf_host, size_t buf_bytes, void* stream);

/* ---- diagnostics used by the parity tests (stage outputs) ------------------------------------------ */
/* tile (n x 256 x 256 x 3 BGR u8) and gray (n x 256 x 256 u8) of the last dfd_forensics_batch call. */
int dfd_dbg_tiles(dfd_ctx* ctx, uint8_t* tile_out, uint8_t* gray_out, int n, void* stream);
/* JPEG Q90 round trip of n tiles (256 x 256 x 3 BGR u8) -> same shape. */
int dfd_dbg_jpeg_roundtrip(dfd_ctx* ctx, const uint8_t* tiles, uint8_t* out, int n, void* stream);
/* Canny(50,150) edge map (0/255) of n gray tiles (256 x 256). */
int dfd_dbg_canny(dfd_ctx* ctx, const uint8_t* gray, uint8_t* edges, int n, void* stream);
/* tcgen05 GEMM self-test: C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual) on random bf16 data against a
 * CUDA-core reference; writes max |err| to *max_err_host (HOST double) and returns 0 if the kernel ran. */
int dfd_gemm_selftest(dfd_ctx* ctx, int M, int N, int K, int act, int with_residual, double* max_err_host, void* stream);
/* Mean milliseconds per launch of one GEMM shape over `iters` launches (kernel tuning; flags: 1 = no stores, 2 = no epilogue
 * math, 4 = SE-gated A, 8 = residual). */
int dfd_gemm_bench(dfd_ctx* ctx, int M, int N, int K, int act, int flags, int iters, double* ms_host, void* stream);
/* 3xTF32 GEMM self-test (fp32 accuracy mode, gemm_tf32x3.cu): C = act(A . W^T + bias) (+ residual) on random fp32 data against a
 * CUDA-core reference with fp64 accumulation; mode bit 0 = residual, bit 1 = SE-gated A (hw = 49).  Writes max relative error to
 * *max_err_host; iters > 0 also times the launch (mean ms -> *ms_host). */
int dfd_gemm_tf32_selftest(dfd_ctx* ctx, int M, int N, int K, int act, int mode, int iters, double* max_err_host,
                           double* ms_host, void* stream);
/* After dfd_face_prep_batch: the 160 x 160 x 3 RGB u8 image of box i. */
int dfd_dbg_face160(dfd_ctx* ctx, int i, uint8_t* out_dev, void* stream);
/* After dfd_face_prep_batch: the CLAHE'd crop of box i as w*h*3 BGR u8 (tight). */
int dfd_dbg_face_clahe(dfd_ctx* ctx, const uint8_t* frames, int H, int W, size_t frame_stride, int row_pitch,
                       const int32_t* boxes, const int32_t* frame_idx, int i, uint8_t* out_dev, void* stream);
/* Activation tap: dfd_dbg_set_tap(name) before dfd_effnet_forward records that layer's output
 * (name = "stem", "b<i>.expand", "b<i>.dw", "b<i>.out", "features"; NULL/"" = off);
 * dfd_dbg_activation then copies min(n_floats, size) float32 values (NHWC, converted from bf16 if
 * needed) and returns the element count or a negative status. */
int dfd_dbg_set_tap(dfd_ctx* ctx, const char* name);
int64_t dfd_dbg_activation(dfd_ctx* ctx, const char* name, float* out_dev, int64_t n_floats, void* stream);
/* A/B switches for the parity tests: "no_fuse" (1 = run the expand 1x1 GEMM and the depthwise kernel separately
 * instead of the fused mbconv_fused.cu kernel), "se_mode" (bf16 SE excite: 0 = two kernels, 1 = one kernel, 2 = one kernel on 8-CTA clusters), "no_gated_w" (1 = blocks 0-4 gate the
 * project GEMM's A operand instead of using per-image gated weights), "pdl" (0 = no programmatic dependent launch),
 * "no_overlap" (1 = forensic kernels on the caller's stream), "fp32_simt" (1 = fp32 mode on the CUDA-core kernels instead of
 * the 3xTF32 tensor-core path), "dual_chain" (0 = the classifier runs a large batch as ONE chain of kernels instead of two
 * concurrent half-batch chains).
 * Threading / devices: one context per GPU; every entry point makes the context's device current for its duration and
 * restores the caller's, so one process may drive several contexts on different GPUs (from one thread at a time each). */
int dfd_dbg_set_option(dfd_ctx* ctx, const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* DFD_H */
