// Internal context shared by the translation units of libdfd.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <unordered_map>
#include "../../include/dfd.h"
#include "px_color.h"

#define DFD_MAX_VOTES 64
#define DFD_MAX_SCORES 128
#define DFD_RING 30                 // FrameForensicAnalyzer.temporal_diffs deque(maxlen=30)
#define DFD_POOL_FLOATS 32768         // per-image SE squeeze partials: n_parts * C floats
#define DFD_NBLK 64                 // 8x8 blocks of 32x32 on the 256^2 tile
#define DFD_FFT_GROUPS 17           // 129 half-spectrum columns in groups of 8

// Per-stream state (device).  Replaces the Python objects' fields:
//   FrameForensicAnalyzer.{prev_frame_gray, temporal_diffs, frame_count}  frame_analysis.py:35-37
//   TemporalTracker.{score_history, frame_classifications, current_verdict} deepfake_detection.py:111-118
//   DeepfakeDetector.frame_count                                           deepfake_detection.py:324
struct DfdStreamState {
    int32_t has_prev;
    int32_t analyzer_frames;
    int32_t ring_n, ring_head;
    float ring[DFD_RING];
    int32_t score_n, score_head;
    int32_t vote_n, vote_head;
    int32_t verdict;
    int32_t detector_frames;
    int32_t window_size, voting_window;     // per-stream TemporalTracker parameters
    double threshold;
    uint8_t votes[DFD_MAX_VOTES];
    uint8_t score_is_np[DFD_MAX_SCORES];    // 1 = the score was a numpy scalar in the reference (see vote.cu py_sum)
    double scores[DFD_MAX_SCORES];
};

// Per-frame partial statistics written by the tile kernels and reduced by the finalize kernel.
struct DfdFramePartials {
    long long noise_sx[DFD_NBLK], noise_sxx[DFD_NBLK];      // x = 256*gray - 256*blur (exact integers)
    long long lap_s[DFD_NBLK], lap_ss[DFD_NBLK];
    unsigned long long sat_s[DFD_NBLK], sat_ss[DFD_NBLK], val_s[DFD_NBLK], val_ss[DFD_NBLK];
    unsigned int hue_bits[DFD_NBLK][6];
    int tdiff[DFD_NBLK];
    int ela_sum[DFD_NBLK];
    int canny_count;
    int pad;
    double fft[DFD_FFT_GROUPS][8];   // low_sum, mid_sum, mid_sumsq, high_sum, n_low, n_mid, n_high, -
};

struct DfdBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct dfd_ctx {
    dfd_config cfg;
    int sm_count = 0;
    std::string err;
    int64_t launches = 0;
    // tables + state
    DfdColorTables* d_tables = nullptr;
    struct RsEntry { int h, w; void* p; };
    std::vector<RsEntry> rs_cache;        // cv2.resize tap tables per frame size (forensics.cu)
    float2* d_twiddle = nullptr;          // 128 twiddles of the 256-point FFT
    DfdStreamState* d_state = nullptr;
    uint8_t* d_prev_gray = nullptr;       // [max_streams][256*256]
    // forensic workspaces (max_batch)
    uint8_t* d_tile = nullptr;            // [n][256][256][3]
    uint8_t* d_gray = nullptr;            // [n][256][256]
    float2* d_fft = nullptr;              // [n][129][256]
    DfdFramePartials* d_part = nullptr;   // [n]
    dfd_forensic_result* d_fres = nullptr;  // [n] internal results for analyze_batch
    // face-prep workspaces
    uint8_t* d_luts = nullptr;            // [m][64][256]
    int* d_pil = nullptr;                 // [m][2][160][2+KMAX]
    uint8_t* d_hpass = nullptr;           // [m][max_crop][160][3]
    uint8_t* d_face160 = nullptr;         // [m][160][160][3]
    int32_t* d_boxes_ok = nullptr;        // [m][4] boxes clamped to the frame (k_box_sanitize)
    int32_t* d_fidx_ok = nullptr;         // [m]
    uint8_t* d_box_bad = nullptr;         // [m] 1 = box rejected (empty after clamping / larger than max_crop): probability NaN
    int box_flags_m = 0;                  // number of boxes the flags describe (last face-prep call)
    DfdBuf tta_base;                      // [m][max_crop*max_crop*3] CLAHE'd crops, the base images of the test-time augmentations (allocated on first use)
    int calib_kind = 0;                   // probability calibrator (dfd_set_calibrator): 0 none, 1 logistic, 2 piecewise linear
    int calib_n = 0;
    DfdBuf calib;                         // [2][calib_n] doubles: xs then ys (kind 2) or {coef, intercept} (kind 1)
    DfdBuf draw_buf;                      // overlay.cu: command list + text masks of one dfd_draw_overlay call
    void* draw_host = nullptr;            //   pinned staging copy
    size_t draw_host_bytes = 0;
    // classifier
    bool has_weights = false;
    float* d_wf32 = nullptr;              // packed folded fp32 parameters
    size_t w_floats = 0;
    __nv_bfloat16* d_wbf16 = nullptr;     // bf16 copies of the GEMM weights (same offsets)
    __nv_bfloat16* d_stem_wg = nullptr;   // stem weights as a [32][32] K-major GEMM operand (27 taps + zero pad)
    float* d_wtf_hi = nullptr;            // fp32 accuracy mode on the tensor cores (gemm_tf32x3.cu): tf32 hi / lo planes of the
    float* d_wtf_lo = nullptr;            //   parameter blob (same offsets), W = hi + lo to ~2^-22
    float* d_stem_wtf = nullptr;          // stem [32][32] K-major operand, hi plane then lo plane
    bool no_subbatch = true;              // fp32 mode: L2-resident expand -> depthwise sub-batches are OFF (measured slower, see effnet.cu);
    int l2_budget = 64 << 20;             //   DFD_L2_BUDGET_MB=n / "no_subbatch" = 0 turns them on with n MB of expanded tensor per sub-batch
    bool fp32_simt = false;               // "fp32_simt" / DFD_FP32_SIMT=1: run the fp32 mode on the CUDA-core kernels (k_pw / k_dw / k_stem; A/B testing)
    DfdBuf act[3];                        // activation ping-pong + expanded buffer
    DfdBuf act_b[3];                      // the same for the second half-batch chain (effnet.cu dfd_effnet_launch)
    bool dual_chain = false;              // batches >= dual_min run as two concurrent half-batch chains ("dual_chain" option / DFD_DUAL_CHAIN=1): MEASURED AND OFF -- fp32 5.27 -> 5.47 ms per 256 crops, bf16 unchanged: the persistent GEMM CTAs (140-200 KB of shared memory, 61 k registers) leave no room for a second chain's CTAs, and half-size kernels lose efficiency
    int dual_min = 64;
    cudaStream_t aux2 = nullptr;
    cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
    DfdBuf face_in;                       // prepared crops for analyze_batch
    float* d_pool = nullptr;              // [m][n_parts][C] SE squeeze partial sums (<= DFD_POOL_FLOATS per image)
    float* d_sescale = nullptr;           // [m][1152]
    __nv_bfloat16* d_wgated = nullptr;    // [m][112][672] (sized for block 10) per-image SE-gated project weights of blocks 0-4
    __nv_bfloat16* d_wgated_fold = nullptr;   // [m][32][64] block 0: block-diagonal weights of the 2-pixel folded GEMM (off-diagonal zeros)
    float* d_bias_fold = nullptr;         // [32] block 0 project bias, repeated
    bool no_fold = false;                 // "no_fold": block 0 project GEMM with one pixel per row
    float* d_front_aux = nullptr;         // mbconv_fused.cu: per block / chunk packed depthwise weights + biases
    size_t front_aux_off[16] = {0};
    float* d_se_r = nullptr;              // [m][64] squeezed activations between the two SE kernels
    float* d_wxt = nullptr;               // transposed SE expand weights, all blocks
    float* d_feat = nullptr;              // [m][1280]
    float* d_fc_h1 = nullptr;             // [m][512] classifier hidden layers
    float* d_fc_h2 = nullptr;             // [m][256]
    float* d_logits = nullptr;            // [m]
    double* d_faceprob = nullptr;         // [m]
    double* d_voteinput = nullptr;        // [n]
    // device JPEG ingest (jpegdec.cu)
    void* jpg_host = nullptr;             // host-side staging (pinned header / meta arrays)
    cudaEvent_t jpg_ev = nullptr;         // the previous call's copies out of the staging arrays
    DfdBuf jpg_raw, jpg_words, jpg_sub, jpg_coef, jpg_dc, jpg_planes, jpg_hdr;
    // diagnostics
    std::string tap_name;
    DfdBuf tap;
    int64_t tap_elems = 0;
    void* tmaps = nullptr;                // host-side cache of TMA descriptors (effnet_bf16.cu)
    // analyze_batch runs the forensic kernels on a second stream, concurrently with face prep + classifier
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // per-launch profiling (bench.py roofline): an event after every launch, labelled
    bool profiling = false;
    bool flight = false;                  // DFD_FLIGHT=1: stream-ordered completion markers in mapped host memory (hang diagnosis)
    volatile unsigned long long* h_mark = nullptr;
    unsigned long long* d_mark = nullptr;
    unsigned long long flight_seq = 0;
    std::vector<std::string> flight_names;
    bool trace = false;                   // DFD_TRACE=1: synchronise after every launch and log it (debugging)
    int se_mode = 2;                      // bf16 SE excite: 0 = k_se_reduce + k_se_expand, 1 = k_se_excite (CTA per 4 images), 2 = k_se_cluster (8-CTA clusters, DSMEM)
    bool pdl = true;                      // programmatic dependent launch between the tcgen05 / SE kernels (DFD_NO_PDL=1 or "pdl" option = 0 disables)
    int gated_w_max = 4;                  // last block whose project conv uses per-image gated weights (DFD_GATED_W_MAX)
    bool no_gated_w = false;              // "no_gated_w": blocks 0-4 apply the SE gate to the A operand (A_SCALE) instead of per-image weights
    bool no_fuse = false;                 // DFD_NO_FUSE=1: expand GEMM + depthwise as two kernels (A/B testing of mbconv_fused.cu)
    bool no_overlap = false;              // DFD_NO_OVERLAP=1: run the forensic kernels on the caller's stream
    const char* label = "";               // set by the launch code before each kernel
    // Function attributes (dynamic shared-memory limit, carve-out) are PER DEVICE: they are tracked per context, never in
    // function-local statics, so a second context on another GPU of the same process sets them again for its device.
    std::unordered_map<const void*, size_t> func_smem;
    std::vector<cudaEvent_t> prof_events;
    std::vector<std::string> prof_labels;
    size_t prof_used = 0;
};

void dfd_prof_mark(dfd_ctx* ctx, const char* kernel, cudaStream_t st);
void dfd_trace(dfd_ctx* ctx, const char* kernel, cudaStream_t st);
void dfd_flight_mark(dfd_ctx* ctx, const char* kernel, cudaStream_t st);

#define DFD_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                  \
            return DFD_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define DFD_LAUNCH_CHECK(kname, st_)                                                        \
    do {                                                                                    \
        ctx->launches++;                                                                    \
        if (ctx->profiling) dfd_prof_mark(ctx, kname, st_);                                 \
        if (ctx->trace) dfd_trace(ctx, kname, st_);                                         \
        if (ctx->flight) dfd_flight_mark(ctx, kname, st_);                                  \
        cudaError_t e_ = cudaPeekAtLastError();                                             \
        if (e_ != cudaSuccess) {                                                            \
            ctx->err = std::string("kernel launch at ") + __FILE__ + ":" + std::to_string(__LINE__) + ": " + \
                       cudaGetErrorString(e_);                                              \
            return DFD_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define DFD_REQUIRE(cond, code, msg)                                                        \
    do {                                                                                    \
        if (!(cond)) { ctx->err = msg; return code; }                                       \
    } while (0)

int dfd_ensure(dfd_ctx* ctx, DfdBuf& b, size_t bytes);

// Raises the kernel's dynamic shared-memory limit on this context's device to at least `bytes`.  The attribute belongs to the
// (device, function) pair and cudaFuncSetAttribute REPLACES it, so the largest value any context of this process asked for is
// kept in a process-wide table (dfd_api.cu): a second context with smaller workspaces must never lower the limit under a
// context that needs more (two Engines with different max_crop on one GPU).  ctx->func_smem caches what this context knows.
int dfd_func_smem_raise(dfd_ctx* ctx, const void* fn, size_t bytes, bool full_carveout);
template <typename F>
static inline int dfd_func_smem(dfd_ctx* ctx, F* fn, size_t bytes, bool full_carveout = false) {
    size_t& have = ctx->func_smem[(const void*)fn];
    if (bytes <= have) return DFD_OK;
    int rc = dfd_func_smem_raise(ctx, (const void*)fn, bytes, full_carveout);
    if (rc == DFD_OK) have = bytes;
    return rc;
}

// Makes the context's device current for the duration of a C entry point (a process may hold one context per GPU).
struct DfdDeviceGuard {
    int prev = -1; bool switched = false;
    explicit DfdDeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DfdDeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// Programmatic dependent launch (PDL): the kernel may become resident while its predecessor in the stream is still
// draining, run its prologue (barrier init, TMEM allocation, descriptor prefetch, weight loads) and block in
// griddepcontrol.wait until the predecessor's memory is visible.  EVERY kernel launched this way executes
// pdl_wait() before it reads or writes anything another kernel touches -- that keeps completion transitive along the
// stream -- and pdl_trigger() right after its prologue.
template <typename... KArgs, typename... Args>
static inline cudaError_t dfd_launch(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// x * sigmoid(x) for the fp32 accuracy mode in ~10 instructions and ~3 ulp (expf + IEEE division: ~35 instructions, which made
// the epilogues of the expand convs ALU-bound).  exp(-x) = 2^t with t = -x * log2(e) carried as hi + lo (the rounding error
// of the product would otherwise scale with |x|), 2^t from MUFU.EX2 (2 ulp), corrected by (1 + lo * ln 2); division by
// 1 + e with the 2-ulp approximate divide.  t is clamped so that e stays finite (swish(-87) is 0 to fp32 anyway).
__device__ __forceinline__ float swish_f32(float x) {
    const float nx = -x;
    float t = nx * 1.44269502162933349609375f;
    float tl = fmaf(nx, 1.44269502162933349609375f, -t);
    tl = fmaf(nx, 1.92596299112661746e-8f, tl);
    t = fminf(t, 126.0f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    e = fmaf(e * tl, 0.693147182464599609375f, e);
    return __fdividef(x, 1.0f + e);
}
// The same on two values with packed fp32 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2): 6 packed + 2 scalar + 4 MUFU
// instructions per PAIR instead of ~13 per value -- the epilogues of the fp32 expand convs and the depthwise kernel are
// instruction-issue-bound.  u = -(lo part of the exponent); 2^(t - u) = 2^t (1 - u ln 2).
__device__ __forceinline__ uint64_t f2_pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t swish_f32x2(uint64_t x) {
    const uint64_t L2E = f2_pack(1.44269502162933349609375f, 1.44269502162933349609375f);
    uint64_t t = f2_mul(x, f2_pack(-1.44269502162933349609375f, -1.44269502162933349609375f));
    uint64_t u = f2_fma(x, L2E, t);                                       // x * log2e + t, exact: minus the rounding error of t
    u = f2_fma(x, f2_pack(1.92596299112661746e-8f, 1.92596299112661746e-8f), u);
    float t0, t1, e0, e1;
    f2_unpack(t, t0, t1);
    t0 = fminf(t0, 126.0f); t1 = fminf(t1, 126.0f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
    uint64_t e = f2_pack(e0, e1);
    e = f2_fma(f2_mul(e, u), f2_pack(-0.693147182464599609375f, -0.693147182464599609375f), e);
    const uint64_t d = f2_add(e, f2_pack(1.0f, 1.0f));
    float d0, d1, r0, r1;
    f2_unpack(d, d0, d1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
    return f2_mul(x, f2_pack(r0, r1));
}
#endif

// forensics.cu
int dfd_forensics_launch(dfd_ctx* ctx, const uint8_t* frames, int n, int H, int W, size_t frame_stride, int row_pitch,
                         const int32_t* stream_ids, const uint8_t* full, dfd_forensic_result* results, cudaStream_t st);
// faceprep.cu
int dfd_faceprep_launch(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, size_t frame_stride,
                        int row_pitch, const int32_t* boxes, const int32_t* frame_idx, int m, void* out, int dtype,
                        cudaStream_t st);
int dfd_faceprep_tta_launch(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, size_t frame_stride,
                            int row_pitch, const int32_t* boxes, const int32_t* frame_idx, int m, int n_pred,
                            const dfd_tta_aug* augs, void* out, int dtype, cudaStream_t st);
// overlay.cu
int dfd_overlay_launch(dfd_ctx* ctx, uint8_t* frame, int H, int W, int row_pitch, const dfd_draw_cmd* cmds_host, int n_cmds,
                       const uint8_t* masks_host, size_t mask_bytes, cudaStream_t st);
// jpegdec.cu
int dfd_jpeg_decode_launch(dfd_ctx* ctx, const uint8_t* bytes_host, const int64_t* offsets_host, int n, int H, int W,
                           uint8_t* frames_out, size_t frame_stride, int row_pitch, int32_t* status_dev, cudaStream_t st);
void dfd_jpeg_free(dfd_ctx* ctx);
// effnet.cu
int dfd_effnet_launch(dfd_ctx* ctx, const void* in, int m, int dtype, float* logits, cudaStream_t st);
int dfd_effnet_upload(dfd_ctx* ctx, const float* blob, size_t n);
size_t dfd_effnet_blob_floats();
// vote.cu
int dfd_faceprob_launch(dfd_ctx* ctx, const float* logits, const int32_t* boxes, int m, int n_pred, double* prob, cudaStream_t st);
int dfd_vote_launch(dfd_ctx* ctx, const int32_t* stream_ids, const double* vote_input, const uint8_t* np_flags, int n,
                    dfd_vote_record* rec, cudaStream_t st);
int dfd_select_vote_launch(dfd_ctx* ctx, int n, int m, const int32_t* box_frame, const double* face_prob,
                           const dfd_forensic_result* fres, const int32_t* stream_ids, dfd_vote_record* rec,
                           cudaStream_t st);
int dfd_reset_launch(dfd_ctx* ctx, int stream_id, int what, cudaStream_t st);
int dfd_configure_launch(dfd_ctx* ctx, int stream_id, int window_size, int voting_window, double thr, cudaStream_t st);
