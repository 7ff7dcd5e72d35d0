// C-ABI of libdfd.so (include/dfd.h): context, workspaces, entry points.
#include "dfd_internal.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>

int dfd_forensics_init(dfd_ctx* ctx);
int dfd_dbg_jpeg_launch(dfd_ctx* ctx, const uint8_t* tiles, uint8_t* out, int n, cudaStream_t st);
int dfd_dbg_canny_launch(dfd_ctx* ctx, const uint8_t* gray, uint8_t* edges, int n, cudaStream_t st);
int dfd_dbg_clahe_launch(dfd_ctx* ctx, const uint8_t* frames, size_t frame_stride, int row_pitch, const int32_t* boxes,
                         const int32_t* frame_idx, int i, uint8_t* out, cudaStream_t st);
void dfd_gemm_free(dfd_ctx* ctx);

static thread_local std::string g_create_err;

int dfd_ensure(dfd_ctx* ctx, DfdBuf& b, size_t bytes) {
    if (b.bytes >= bytes) return DFD_OK;
    if (b.p) { DFD_CUDA(cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
    DFD_CUDA(cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
    return DFD_OK;
}

void dfd_prof_mark(dfd_ctx* ctx, const char* kernel, cudaStream_t st) {
    if (ctx->prof_used >= ctx->prof_events.size()) return;
    cudaEventRecord(ctx->prof_events[ctx->prof_used], st);
    ctx->prof_labels[ctx->prof_used] = std::string(kernel) + (ctx->label[0] ? std::string(":") + ctx->label : std::string());
    ctx->prof_used++;
}

__global__ void k_mark(unsigned long long* p, unsigned long long v) { *p = v; __threadfence_system(); }

void dfd_flight_mark(dfd_ctx* ctx, const char* kernel, cudaStream_t st) {
    if (ctx->flight_names.empty()) ctx->flight_names.resize(4096);
    ctx->flight_seq++;
    ctx->flight_names[ctx->flight_seq % 4096] = std::string(kernel) + ":" + ctx->label;
    k_mark<<<1, 1, 0, st>>>(ctx->d_mark, ctx->flight_seq);     // completes after kernel #flight_seq on this stream
}

void dfd_trace(dfd_ctx* ctx, const char* kernel, cudaStream_t st) {
    fprintf(stderr, "[dfd] %s:%s ...", kernel, ctx->label);
    fflush(stderr);
    cudaError_t e = cudaStreamSynchronize(st);
    fprintf(stderr, " %s\n", e == cudaSuccess ? "ok" : cudaGetErrorString(e));
    fflush(stderr);
}

#include <map>
#include <mutex>
int dfd_func_smem_raise(dfd_ctx* ctx, const void* fn, size_t bytes, bool full_carveout) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> limit;          // (device, function) -> bytes set so far
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = limit[{ctx->cfg.device, fn}];
    if (bytes > cur) {
        DFD_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        if (full_carveout) DFD_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        cur = bytes;
    }
    return DFD_OK;
}

extern "C" {

int dfd_profile_start(dfd_ctx* ctx, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    const size_t cap = 8192;
    if (ctx->prof_events.empty()) {
        ctx->prof_events.resize(cap);
        ctx->prof_labels.resize(cap);
        for (size_t i = 0; i < cap; i++) DFD_CUDA(cudaEventCreate(&ctx->prof_events[i]));
    }
    ctx->prof_used = 0;
    ctx->label = "";
    ctx->profiling = true;
    dfd_prof_mark(ctx, "begin", (cudaStream_t)stream);
    return DFD_OK;
}

// Stops profiling and writes "kernel:label,launches,total_ms" lines (sorted by first appearance) into buf.
int dfd_profile_stop(dfd_ctx* ctx, char* buf, size_t buf_bytes, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    ctx->profiling = false;
    DFD_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    std::vector<std::string> order;
    std::vector<double> ms;
    std::vector<int> cnt;
    for (size_t i = 1; i < ctx->prof_used; i++) {
        float t = 0.f;
        DFD_CUDA(cudaEventElapsedTime(&t, ctx->prof_events[i - 1], ctx->prof_events[i]));
        size_t k = 0;
        for (; k < order.size(); k++) if (order[k] == ctx->prof_labels[i]) break;
        if (k == order.size()) { order.push_back(ctx->prof_labels[i]); ms.push_back(0); cnt.push_back(0); }
        ms[k] += t; cnt[k]++;
    }
    std::string out;
    for (size_t k = 0; k < order.size(); k++) {
        char line[256];
        snprintf(line, sizeof line, "%s,%d,%.6f\n", order[k].c_str(), cnt[k], ms[k]);
        out += line;
    }
    if (buf && buf_bytes) {
        size_t n = out.size() < buf_bytes - 1 ? out.size() : buf_bytes - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return DFD_OK;
}

void dfd_default_config(dfd_config* c) {
    memset(c, 0, sizeof(*c));
    c->device = 0; c->max_streams = 256; c->max_batch = 256; c->max_crop = 1024;
    c->window_size = 60; c->voting_window = 10; c->detection_threshold = 0.5;
    c->face_weight = 0.70; c->forensic_weight = 0.30; c->blend_mode = DFD_BLEND_REFERENCE;
}

int dfd_abi_version(void) { return DFD_ABI_VERSION; }

const char* dfd_last_error(dfd_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

static int create_impl(dfd_ctx* ctx) {
    const dfd_config& c = ctx->cfg;
    DFD_REQUIRE(c.max_streams > 0 && c.max_batch > 0 && c.max_crop >= 8, DFD_ERR_INVALID, "create: bad capacities");
    DFD_REQUIRE(c.max_crop <= 31 * 160 && c.max_crop * 3 <= 48 * 1024, DFD_ERR_INVALID, "create: max_crop too large");
    DFD_REQUIRE(c.window_size >= 10 && c.window_size <= DFD_MAX_SCORES, DFD_ERR_INVALID, "create: window_size must be 10..128");
    DFD_REQUIRE(c.voting_window >= 1 && c.voting_window <= DFD_MAX_VOTES, DFD_ERR_INVALID, "create: voting_window must be 1..64");
    DFD_CUDA(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    DFD_CUDA(cudaGetDeviceProperties(&prop, c.device));
    DFD_REQUIRE(prop.major == 10, DFD_ERR_ARCH, "create: libdfd is built for sm_100a (B200) only");
    ctx->sm_count = prop.multiProcessorCount;
    // tables
    DfdColorTables* T = new DfdColorTables;
    dfd_build_color_tables(T);
    DFD_CUDA(cudaMalloc(&ctx->d_tables, sizeof(DfdColorTables)));
    DFD_CUDA(cudaMemcpy(ctx->d_tables, T, sizeof(DfdColorTables), cudaMemcpyHostToDevice));
    delete T;
    float2 tw[128];
    for (int j = 0; j < 128; j++) {
        double a = -2.0 * M_PI * (double)j / 256.0;
        tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    DFD_CUDA(cudaMalloc(&ctx->d_twiddle, sizeof(tw)));
    DFD_CUDA(cudaMemcpy(ctx->d_twiddle, tw, sizeof(tw), cudaMemcpyHostToDevice));
    // per-stream state
    DFD_CUDA(cudaMalloc(&ctx->d_state, sizeof(DfdStreamState) * c.max_streams));
    DFD_CUDA(cudaMemset(ctx->d_state, 0, sizeof(DfdStreamState) * c.max_streams));
    DFD_CUDA(cudaMalloc(&ctx->d_prev_gray, (size_t)c.max_streams * 65536));
    DFD_CUDA(cudaMemset(ctx->d_prev_gray, 0, (size_t)c.max_streams * 65536));
    const size_t nb = c.max_batch;
    DFD_CUDA(cudaMalloc(&ctx->d_tile, nb * 65536 * 3));
    DFD_CUDA(cudaMalloc(&ctx->d_gray, nb * 65536));
    DFD_CUDA(cudaMalloc(&ctx->d_fft, nb * 129 * 256 * sizeof(float2)));
    DFD_CUDA(cudaMalloc(&ctx->d_part, nb * sizeof(DfdFramePartials)));
    DFD_CUDA(cudaMemset(ctx->d_part, 0, nb * sizeof(DfdFramePartials)));
    DFD_CUDA(cudaMalloc(&ctx->d_fres, nb * sizeof(dfd_forensic_result)));
    DFD_CUDA(cudaMalloc(&ctx->d_luts, nb * 64 * 256));
    DFD_CUDA(cudaMalloc(&ctx->d_pil, nb * 2 * 160 * (2 + 64) * sizeof(int)));
    DFD_CUDA(cudaMalloc(&ctx->d_hpass, nb * (size_t)c.max_crop * 480));
    DFD_CUDA(cudaMalloc(&ctx->d_face160, nb * 160 * 480));
    DFD_CUDA(cudaMalloc(&ctx->d_boxes_ok, nb * 4 * sizeof(int32_t)));
    DFD_CUDA(cudaMalloc(&ctx->d_fidx_ok, nb * sizeof(int32_t)));
    DFD_CUDA(cudaMalloc(&ctx->d_box_bad, nb));
    DFD_CUDA(cudaMemset(ctx->d_box_bad, 0, nb));
    DFD_CUDA(cudaMalloc(&ctx->d_pool, nb * DFD_POOL_FLOATS * sizeof(float)));
    DFD_CUDA(cudaMalloc(&ctx->d_sescale, nb * 1152 * sizeof(float)));
    DFD_CUDA(cudaMalloc(&ctx->d_se_r, nb * 64 * sizeof(float)));
    DFD_CUDA(cudaMalloc(&ctx->d_wgated, nb * 112 * 672 * sizeof(__nv_bfloat16)));
    DFD_CUDA(cudaMalloc(&ctx->d_wgated_fold, nb * 32 * 64 * sizeof(__nv_bfloat16)));
    DFD_CUDA(cudaMemset(ctx->d_wgated_fold, 0, nb * 32 * 64 * sizeof(__nv_bfloat16)));
    DFD_CUDA(cudaMalloc(&ctx->d_feat, nb * 1280 * sizeof(float)));
    DFD_CUDA(cudaMalloc(&ctx->d_fc_h1, nb * 512 * sizeof(float)));
    DFD_CUDA(cudaMalloc(&ctx->d_fc_h2, nb * 256 * sizeof(float)));
    DFD_CUDA(cudaMalloc(&ctx->d_logits, nb * sizeof(float)));
    DFD_CUDA(cudaMalloc(&ctx->d_faceprob, nb * sizeof(double)));
    DFD_CUDA(cudaMalloc(&ctx->d_voteinput, nb * sizeof(double)));
    if (ctx->flight) {
        DFD_CUDA(cudaHostAlloc((void**)&ctx->h_mark, 8, cudaHostAllocMapped));
        *ctx->h_mark = 0;
        DFD_CUDA(cudaHostGetDevicePointer((void**)&ctx->d_mark, (void*)ctx->h_mark, 0));
    }
    DFD_CUDA(cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking));
    DFD_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    DFD_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    int rc = dfd_forensics_init(ctx);
    if (rc) return rc;
    if ((rc = dfd_configure_launch(ctx, -1, c.window_size, c.voting_window, c.detection_threshold, 0))) return rc;
    DFD_CUDA(cudaDeviceSynchronize());
    return DFD_OK;
}

int dfd_create(const dfd_config* cfg, dfd_ctx** out) {
    if (!cfg || !out) { g_create_err = "create: null argument"; return DFD_ERR_INVALID; }
    dfd_ctx* ctx = new dfd_ctx;
    ctx->cfg = *cfg;
    ctx->trace = getenv("DFD_TRACE") != nullptr;
    ctx->flight = getenv("DFD_FLIGHT") != nullptr;
    ctx->no_overlap = getenv("DFD_NO_OVERLAP") != nullptr;
    ctx->no_fuse = getenv("DFD_NO_FUSE") != nullptr;
    ctx->pdl = getenv("DFD_NO_PDL") == nullptr;
    ctx->no_gated_w = getenv("DFD_NO_GATED_W") != nullptr;
    if (const char* e = getenv("DFD_GATED_W_MAX")) { ctx->gated_w_max = atoi(e); if (ctx->gated_w_max > 10) ctx->gated_w_max = 10; }
    ctx->no_fold = getenv("DFD_NO_FOLD") != nullptr;
    ctx->fp32_simt = getenv("DFD_FP32_SIMT") != nullptr;
    if (getenv("DFD_DUAL_CHAIN")) ctx->dual_chain = atoi(getenv("DFD_DUAL_CHAIN")) != 0;
    if (getenv("DFD_DUAL_MIN")) ctx->dual_min = atoi(getenv("DFD_DUAL_MIN"));
    if (const char* e = getenv("DFD_L2_BUDGET_MB")) { ctx->l2_budget = atoi(e) << 20; ctx->no_subbatch = ctx->l2_budget <= 0; }
    if (getenv("DFD_SE_MODE")) ctx->se_mode = atoi(getenv("DFD_SE_MODE"));
    int rc = create_impl(ctx);
    if (rc) { g_create_err = ctx->err; dfd_destroy(ctx); *out = nullptr; return rc; }
    *out = ctx;
    return DFD_OK;
}

void dfd_destroy(dfd_ctx* ctx) {
    if (!ctx) return;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    void* ptrs[] = {ctx->d_tables, ctx->d_twiddle, ctx->d_state, ctx->d_prev_gray, ctx->d_tile, ctx->d_gray, ctx->d_fft,
                    ctx->d_part, ctx->d_fres, ctx->d_luts, ctx->d_pil, ctx->d_hpass, ctx->d_face160, ctx->d_boxes_ok, ctx->d_fidx_ok, ctx->d_box_bad, ctx->d_wf32,
                    ctx->d_wbf16, ctx->d_stem_wg, ctx->d_wtf_hi, ctx->d_wtf_lo, ctx->d_stem_wtf, ctx->act[0].p, ctx->act[1].p, ctx->act[2].p, ctx->face_in.p, ctx->d_pool,
                    ctx->d_sescale, ctx->d_se_r, ctx->d_front_aux, ctx->d_wgated, ctx->d_wgated_fold, ctx->d_bias_fold, ctx->d_wxt, ctx->d_feat, ctx->d_fc_h1, ctx->d_fc_h2, ctx->d_logits, ctx->d_faceprob, ctx->d_voteinput, ctx->tap.p};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& e : ctx->rs_cache) if (e.p) cudaFree(e.p);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->aux2) cudaStreamDestroy(ctx->aux2);
    if (ctx->ev_fork2) cudaEventDestroy(ctx->ev_fork2);
    if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
    for (DfdBuf& b : ctx->act_b) if (b.p) cudaFree(b.p);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    dfd_gemm_free(ctx);
    dfd_jpeg_free(ctx);
    if (ctx->draw_host) cudaFreeHost(ctx->draw_host);
    for (DfdBuf* b : {&ctx->jpg_raw, &ctx->jpg_words, &ctx->jpg_sub, &ctx->jpg_coef, &ctx->jpg_dc, &ctx->jpg_planes, &ctx->jpg_hdr, &ctx->tta_base,
                      &ctx->calib, &ctx->draw_buf})
        if (b->p) cudaFree(b->p);
    delete ctx;
}

size_t dfd_weights_blob_floats(void) { return dfd_effnet_blob_floats(); }

int dfd_load_weights(dfd_ctx* ctx, const float* blob_host, size_t n_floats) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(blob_host != nullptr, DFD_ERR_INVALID, "load_weights: null blob");
    return dfd_effnet_upload(ctx, blob_host, n_floats);
}

int dfd_forensics_batch(dfd_ctx* ctx, const uint8_t* frames, int n, int H, int W, size_t frame_stride, int row_pitch,
                        const int32_t* stream_ids, const uint8_t* full, dfd_forensic_result* results, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(frames && stream_ids && full && results, DFD_ERR_INVALID, "forensics_batch: null pointer");
    return dfd_forensics_launch(ctx, frames, n, H, W, frame_stride, row_pitch, stream_ids, full, results, (cudaStream_t)stream);
}

int dfd_face_prep_batch(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, size_t frame_stride, int row_pitch,
                        const int32_t* boxes, const int32_t* frame_idx, int m, void* out_nhwc, int dtype, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(frames && boxes && frame_idx && out_nhwc, DFD_ERR_INVALID, "face_prep_batch: null pointer");
    return dfd_faceprep_launch(ctx, frames, n_frames, H, W, frame_stride, row_pitch, boxes, frame_idx, m, out_nhwc, dtype,
                               (cudaStream_t)stream);
}

int dfd_effnet_forward(dfd_ctx* ctx, const void* in_nhwc, int m, int dtype, float* logits, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(in_nhwc && logits, DFD_ERR_INVALID, "effnet_forward: null pointer");
    return dfd_effnet_launch(ctx, in_nhwc, m, dtype, logits, (cudaStream_t)stream);
}

int dfd_face_probability(dfd_ctx* ctx, const float* logits, const int32_t* boxes, int m, double* prob, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(logits && boxes && prob && m > 0, DFD_ERR_INVALID, "face_probability: bad argument");
    return dfd_faceprob_launch(ctx, logits, boxes, m, 1, prob, (cudaStream_t)stream);
}

int dfd_face_prep_tta(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, size_t frame_stride, int row_pitch,
                      const int32_t* boxes, const int32_t* frame_idx, int m, int n_pred, const dfd_tta_aug* augs,
                      void* out_nhwc, int dtype, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(frames && boxes && frame_idx && out_nhwc, DFD_ERR_INVALID, "face_prep_tta: null pointer");
    return dfd_faceprep_tta_launch(ctx, frames, n_frames, H, W, frame_stride, row_pitch, boxes, frame_idx, m, n_pred, augs, out_nhwc,
                                   dtype, (cudaStream_t)stream);
}

int dfd_face_probability_tta(dfd_ctx* ctx, const float* logits, const int32_t* boxes, int m, int n_pred, double* prob,
                             void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(logits && boxes && prob && m > 0 && n_pred >= 1 && n_pred <= DFD_TTA_MAX_PRED, DFD_ERR_INVALID,
                "face_probability_tta: bad argument");
    return dfd_faceprob_launch(ctx, logits, boxes, m, n_pred, prob, (cudaStream_t)stream);
}

int dfd_draw_overlay(dfd_ctx* ctx, uint8_t* frame, int H, int W, int row_pitch, const dfd_draw_cmd* cmds_host, int n_cmds,
                     const uint8_t* masks_host, size_t mask_bytes, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(frame && (n_cmds == 0 || cmds_host) && (mask_bytes == 0 || masks_host), DFD_ERR_INVALID, "draw_overlay: null pointer");
    return dfd_overlay_launch(ctx, frame, H, W, row_pitch, cmds_host, n_cmds, masks_host, mask_bytes, (cudaStream_t)stream);
}

int dfd_set_calibrator(dfd_ctx* ctx, int kind, int n, const double* xs_host, const double* ys_host, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(kind >= DFD_CALIB_NONE && kind <= DFD_CALIB_PIECEWISE_LINEAR, DFD_ERR_INVALID, "set_calibrator: unknown kind");
    if (kind == DFD_CALIB_NONE) { ctx->calib_kind = 0; ctx->calib_n = 0; return DFD_OK; }
    DFD_REQUIRE(xs_host && ys_host && n >= 1 && n <= (1 << 20), DFD_ERR_INVALID, "set_calibrator: bad table");
    DFD_REQUIRE(kind != DFD_CALIB_LOGISTIC || n == 1, DFD_ERR_INVALID, "set_calibrator: the logistic kind takes one (coef, intercept) pair");
    for (int i = 1; i < n; i++)
        DFD_REQUIRE(xs_host[i] > xs_host[i - 1], DFD_ERR_INVALID, "set_calibrator: xs must be strictly increasing");
    cudaStream_t st = (cudaStream_t)stream;
    DFD_CUDA(cudaStreamSynchronize(st));                        // a running k_faceprob may still read the old table
    int rc = dfd_ensure(ctx, ctx->calib, (size_t)2 * n * sizeof(double));
    if (rc) return rc;
    DFD_CUDA(cudaMemcpyAsync(ctx->calib.p, xs_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    DFD_CUDA(cudaMemcpyAsync((double*)ctx->calib.p + n, ys_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    DFD_CUDA(cudaStreamSynchronize(st));                        // the host arrays belong to the caller
    ctx->calib_kind = kind; ctx->calib_n = n;
    return DFD_OK;
}

int dfd_vote_update(dfd_ctx* ctx, const int32_t* stream_ids, const double* vote_input, const uint8_t* np_flags, int n,
                    dfd_vote_record* records, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(stream_ids && vote_input && records && n > 0, DFD_ERR_INVALID, "vote_update: bad argument");
    return dfd_vote_launch(ctx, stream_ids, vote_input, np_flags, n, records, (cudaStream_t)stream);
}

int dfd_analyze_batch(dfd_ctx* ctx, const uint8_t* frames, int n, int H, int W, size_t frame_stride, int row_pitch,
                      const int32_t* stream_ids, const uint8_t* full, const int32_t* boxes, const int32_t* box_frame, int m,
                      int dtype, dfd_forensic_result* forensic_out, double* face_prob_out, dfd_vote_record* records,
                      void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(frames && stream_ids && full && records, DFD_ERR_INVALID, "analyze_batch: null pointer");
    DFD_REQUIRE(m == 0 || (boxes && box_frame), DFD_ERR_INVALID, "analyze_batch: boxes missing");
    cudaStream_t st = (cudaStream_t)stream;
    dfd_forensic_result* fres = forensic_out ? forensic_out : ctx->d_fres;
    // The forensic signals and the face path are independent until the vote: fork the (ALU/latency-bound)
    // forensic kernels onto the auxiliary stream so they overlap the (HBM-bound) classifier.
    const bool overlap = m > 0 && !ctx->profiling && !ctx->trace && !ctx->no_overlap;
    cudaStream_t fst = overlap ? ctx->aux : st;
    if (overlap) {
        DFD_CUDA(cudaEventRecord(ctx->ev_fork, st));
        DFD_CUDA(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
    }
    int rc = dfd_forensics_launch(ctx, frames, n, H, W, frame_stride, row_pitch, stream_ids, full, fres, fst);
    if (rc) return rc;
    if (overlap) DFD_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux));
    double* fprob = face_prob_out ? face_prob_out : ctx->d_faceprob;
    if (m > 0) {
        size_t esz = dtype == DFD_BF16 ? 2 : 4;
        if ((rc = dfd_ensure(ctx, ctx->face_in, (size_t)m * 224 * 224 * 3 * esz))) return rc;
        if ((rc = dfd_faceprep_launch(ctx, frames, n, H, W, frame_stride, row_pitch, boxes, box_frame, m, ctx->face_in.p, dtype, st))) return rc;
        if ((rc = dfd_effnet_launch(ctx, ctx->face_in.p, m, dtype, ctx->d_logits, st))) return rc;
        if ((rc = dfd_faceprob_launch(ctx, ctx->d_logits, ctx->d_boxes_ok, m, 1, fprob, st))) return rc;   // heuristics see the clamped crop, as face_bgr.shape does
    }
    if (overlap) DFD_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    return dfd_select_vote_launch(ctx, n, m, box_frame, fprob, fres, stream_ids, records, st);
}

int dfd_decode_jpeg_batch(dfd_ctx* ctx, const uint8_t* bytes_host, const int64_t* offsets_host, int n, int H, int W,
                          uint8_t* frames_out, size_t frame_stride, int row_pitch, int32_t* status_dev, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(bytes_host && offsets_host && frames_out && status_dev, DFD_ERR_INVALID, "decode_jpeg_batch: null pointer");
    return dfd_jpeg_decode_launch(ctx, bytes_host, offsets_host, n, H, W, frames_out, frame_stride, row_pitch, status_dev, (cudaStream_t)stream);
}

int dfd_reset_stream(dfd_ctx* ctx, int stream_id, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(stream_id < ctx->cfg.max_streams, DFD_ERR_CAPACITY, "reset_stream: id beyond max_streams");
    return dfd_reset_launch(ctx, stream_id, 3, (cudaStream_t)stream);
}

int dfd_reset_stream_part(dfd_ctx* ctx, int stream_id, int what, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(stream_id < ctx->cfg.max_streams, DFD_ERR_CAPACITY, "reset_stream: id beyond max_streams");
    DFD_REQUIRE(what >= 1 && what <= 3, DFD_ERR_INVALID, "reset_stream_part: what must be 1, 2 or 3");
    return dfd_reset_launch(ctx, stream_id, what, (cudaStream_t)stream);
}

int dfd_configure_stream(dfd_ctx* ctx, int stream_id, int window_size, int voting_window, double detection_threshold,
                         void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(stream_id < ctx->cfg.max_streams, DFD_ERR_CAPACITY, "configure_stream: id beyond max_streams");
    DFD_REQUIRE(window_size >= 10 && window_size <= DFD_MAX_SCORES, DFD_ERR_INVALID, "configure_stream: window_size must be 10..128");
    DFD_REQUIRE(voting_window >= 1 && voting_window <= DFD_MAX_VOTES, DFD_ERR_INVALID, "configure_stream: voting_window must be 1..64");
    return dfd_configure_launch(ctx, stream_id, window_size, voting_window, detection_threshold, (cudaStream_t)stream);
}

/* Hang diagnosis (DFD_FLIGHT=1): the last kernel whose completion marker reached host memory and the ones after it. */
int dfd_flight_report(dfd_ctx* ctx, char* buf, size_t n) {
    if (!ctx || !ctx->flight || !buf || n == 0) return DFD_ERR_INVALID;
    unsigned long long done = *ctx->h_mark, issued = ctx->flight_seq;
    std::string out = "completed #" + std::to_string(done) + " of " + std::to_string(issued) + " issued;";
    for (unsigned long long q = done > 2 ? done - 2 : 1; q <= issued && q <= done + 4; q++)
        out += " [" + std::to_string(q) + (q <= done ? " done " : " PENDING ") + ctx->flight_names[q % 4096] + "]";
    size_t k = out.size() < n - 1 ? out.size() : n - 1;
    memcpy(buf, out.data(), k); buf[k] = 0;
    return DFD_OK;
}

int64_t dfd_launch_count(dfd_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- diagnostics ---------------------------------------------------------------------------------
int dfd_dbg_tiles(dfd_ctx* ctx, uint8_t* tile_out, uint8_t* gray_out, int n, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(n > 0 && n <= ctx->cfg.max_batch, DFD_ERR_CAPACITY, "dbg_tiles: bad n");
    if (tile_out) DFD_CUDA(cudaMemcpyAsync(tile_out, ctx->d_tile, (size_t)n * 65536 * 3, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (gray_out) DFD_CUDA(cudaMemcpyAsync(gray_out, ctx->d_gray, (size_t)n * 65536, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DFD_OK;
}

int dfd_dbg_jpeg_roundtrip(dfd_ctx* ctx, const uint8_t* tiles, uint8_t* out, int n, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    return dfd_dbg_jpeg_launch(ctx, tiles, out, n, (cudaStream_t)stream);
}

int dfd_dbg_canny(dfd_ctx* ctx, const uint8_t* gray, uint8_t* edges, int n, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    return dfd_dbg_canny_launch(ctx, gray, edges, n, (cudaStream_t)stream);
}

int dfd_dbg_face160(dfd_ctx* ctx, int i, uint8_t* out_dev, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(i >= 0 && i < ctx->cfg.max_batch, DFD_ERR_CAPACITY, "dbg_face160: bad index");
    DFD_CUDA(cudaMemcpyAsync(out_dev, ctx->d_face160 + (size_t)i * 160 * 480, 160 * 480, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DFD_OK;
}

int dfd_dbg_face_clahe(dfd_ctx* ctx, const uint8_t* frames, int H, int W, size_t frame_stride, int row_pitch,
                       const int32_t* boxes, const int32_t* frame_idx, int i, uint8_t* out_dev, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    return dfd_dbg_clahe_launch(ctx, frames, frame_stride, row_pitch, boxes, frame_idx, i, out_dev, (cudaStream_t)stream);
}

int dfd_dbg_set_tap(dfd_ctx* ctx, const char* name) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    ctx->tap_name = name ? name : "";
    ctx->tap_elems = 0;
    return DFD_OK;
}

int dfd_dbg_set_option(dfd_ctx* ctx, const char* name, int value) {
    if (!ctx || !name) return DFD_ERR_INVALID;
    const std::string n = name;
    if (n == "no_fuse") ctx->no_fuse = value != 0;
    else if (n == "pdl") ctx->pdl = value != 0;
    else if (n == "no_gated_w") ctx->no_gated_w = value != 0;
    else if (n == "no_fold") ctx->no_fold = value != 0;
    else if (n == "se_mode") ctx->se_mode = value;
    else if (n == "no_overlap") ctx->no_overlap = value != 0;
    else if (n == "fp32_simt") ctx->fp32_simt = value != 0;
    else if (n == "dual_chain") ctx->dual_chain = value != 0;
    else if (n == "no_subbatch") ctx->no_subbatch = value != 0;
    else { ctx->err = "dbg_set_option: unknown option " + n; return DFD_ERR_INVALID; }
    return DFD_OK;
}

int64_t dfd_dbg_activation(dfd_ctx* ctx, const char* name, float* out_dev, int64_t n_floats, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    DFD_REQUIRE(name && ctx->tap_name == name && ctx->tap_elems > 0, DFD_ERR_INVALID,
                "dbg_activation: call dfd_dbg_set_tap(name) before the forward pass");
    int64_t n = n_floats < ctx->tap_elems ? n_floats : ctx->tap_elems;
    if (out_dev && n > 0)
        DFD_CUDA(cudaMemcpyAsync(out_dev, ctx->tap.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return ctx->tap_elems;
}

}  // extern "C"
