// Score selection + 10-frame temporal vote as device-side per-stream ring buffers.
//
// Replaces TemporalTracker (reference deepfake_detection.py:93-289) and the
// vote-input selection of /analyze (backend_server.py:160-174,206): all
// arithmetic is IEEE double in deque insertion order so verdicts and counts are
// bit-identical with the Python objects (hard part 6).
#include "dfd_internal.cuh"

struct VoteCfg {
    int window_size, voting_window, blend_mode;
    double threshold, face_w, forensic_w;
};

__device__ void tracker_update(DfdStreamState& S, const VoteCfg& c, double p, dfd_vote_record& r) {
    if (p == p) {                                   // update(None) is ignored (:123-124)
        if (S.score_n < c.window_size) { S.scores[(S.score_head + S.score_n) % c.window_size] = p; S.score_n++; }
        else { S.scores[S.score_head] = p; S.score_head = (S.score_head + 1) % c.window_size; }
        uint8_t cls = p > c.threshold ? 1 : 0;      // strict > (:135)
        if (S.vote_n < c.voting_window) { S.votes[(S.vote_head + S.vote_n) % c.voting_window] = cls; S.vote_n++; }
        else { S.votes[S.vote_head] = cls; S.vote_head = (S.vote_head + 1) % c.voting_window; }
        int fake = 0;
        for (int k = 0; k < S.vote_n; k++) fake += S.votes[k];
        if (S.vote_n < c.voting_window) S.verdict = DFD_UNCERTAIN;          // :158-160
        else S.verdict = fake > S.vote_n - fake ? DFD_FAKE : DFD_REAL;      // tie -> REAL (:175-178)
    }
    int fake = 0;
    for (int k = 0; k < S.vote_n; k++) fake += S.votes[k];
    r.verdict = S.verdict;
    r.fake_count = fake;
    r.real_count = S.vote_n - fake;
    r.history_len = S.score_n;
    r.frame_count = S.detector_frames;
    r.vote_input = p;
    double sum = 0.0;                               // sum(deque)/len (:198-202)
    for (int k = 0; k < S.score_n; k++) sum = __dadd_rn(sum, S.scores[(S.score_head + k) % c.window_size]);
    r.temporal_average = S.score_n ? __ddiv_rn(sum, (double)S.score_n) : 0.0;
    if (S.score_n < 10) r.stability_score = 0.0;    // :214-221
    else {
        double mean = r.temporal_average, v = 0.0;
        for (int k = 0; k < S.score_n; k++) {
            double d = __dsub_rn(S.scores[(S.score_head + k) % c.window_size], mean);
            v = __dadd_rn(v, __dmul_rn(d, d));
        }
        v = __ddiv_rn(v, (double)S.score_n);
        double m4 = __dmul_rn(v, 4.0);
        r.stability_score = __dsub_rn(1.0, m4 < 1.0 ? m4 : 1.0);
    }
}

__global__ void k_vote(int n, const int32_t* __restrict__ stream_ids, const double* __restrict__ vote_input,
                       DfdStreamState* __restrict__ state, VoteCfg cfg, dfd_vote_record* __restrict__ rec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dfd_vote_record r;
    r.stream_id = stream_ids[i];
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    r.face_probability = NaN; r.forensic_probability = NaN;
    tracker_update(state[r.stream_id], cfg, vote_input[i], r);
    rec[i] = r;
}

// analyze_batch tail: choose the vote input per frame, bump detector.frame_count, update the tracker.
__global__ void k_select_vote(int n, int m, const int32_t* __restrict__ box_frame, const double* __restrict__ face_prob,
                              const dfd_forensic_result* __restrict__ fres, const int32_t* __restrict__ stream_ids,
                              DfdStreamState* __restrict__ state, VoteCfg cfg, dfd_vote_record* __restrict__ rec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    double fp = NaN;
    for (int j = 0; j < m; j++)
        if (box_frame[j] == i) { fp = face_prob[j]; break; }       // faces[0] (backend_server.py:160)
    double forensic = fres[i].fake_probability;
    double p;
    if (fp == fp) p = cfg.blend_mode == DFD_BLEND_README ? cfg.face_w * fp + cfg.forensic_w * forensic : fp;
    else p = forensic;                                               // backend_server.py:206
    dfd_vote_record r;
    r.stream_id = stream_ids[i];
    DfdStreamState& S = state[r.stream_id];
    S.detector_frames += 1;                                          // backend_server.py:156
    r.face_probability = fp;
    r.forensic_probability = forensic;
    tracker_update(S, cfg, p, r);
    rec[i] = r;
}

// sigmoid + apply_heuristics (deepfake_detection.py:398,489-502)
__global__ void k_faceprob(int m, const float* __restrict__ logits, const int32_t* __restrict__ boxes,
                           double* __restrict__ prob) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    float z = logits[i];
    float s = 1.0f / (1.0f + expf(-z));                              // torch.sigmoid in float32
    double p = (double)s;                                            // .item() -> Python float
    int w = boxes[i * 4 + 2], h = boxes[i * 4 + 3];
    double adj = (h < 80 || w < 80) ? 0.10 : 0.0;
    p = p + adj;
    prob[i] = p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p);
}

__global__ void k_reset(DfdStreamState* state, int first, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    DfdStreamState& S = state[first + i];
    S.has_prev = 0; S.analyzer_frames = 0; S.ring_n = 0; S.ring_head = 0;
    S.score_n = 0; S.score_head = 0; S.vote_n = 0; S.vote_head = 0;
    S.verdict = DFD_UNCERTAIN; S.detector_frames = 0;
}

static VoteCfg make_cfg(const dfd_ctx* ctx) {
    VoteCfg c;
    c.window_size = ctx->cfg.window_size; c.voting_window = ctx->cfg.voting_window; c.blend_mode = ctx->cfg.blend_mode;
    c.threshold = ctx->cfg.detection_threshold; c.face_w = ctx->cfg.face_weight; c.forensic_w = ctx->cfg.forensic_weight;
    return c;
}

int dfd_faceprob_launch(dfd_ctx* ctx, const float* logits, const int32_t* boxes, int m, double* prob, cudaStream_t st) {
    k_faceprob<<<(m + 127) / 128, 128, 0, st>>>(m, logits, boxes, prob);
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

int dfd_vote_launch(dfd_ctx* ctx, const int32_t* stream_ids, const double* vote_input, int n, dfd_vote_record* rec,
                    cudaStream_t st) {
    k_vote<<<(n + 63) / 64, 64, 0, st>>>(n, stream_ids, vote_input, ctx->d_state, make_cfg(ctx), rec);
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

int dfd_select_vote_launch(dfd_ctx* ctx, int n, int m, const int32_t* box_frame, const double* face_prob,
                           const dfd_forensic_result* fres, const int32_t* stream_ids, dfd_vote_record* rec,
                           cudaStream_t st) {
    k_select_vote<<<(n + 63) / 64, 64, 0, st>>>(n, m, box_frame, face_prob, fres, stream_ids, ctx->d_state, make_cfg(ctx), rec);
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

int dfd_reset_launch(dfd_ctx* ctx, int stream_id, cudaStream_t st) {
    int first = stream_id < 0 ? 0 : stream_id, count = stream_id < 0 ? ctx->cfg.max_streams : 1;
    k_reset<<<(count + 127) / 128, 128, 0, st>>>(ctx->d_state, first, count);
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}
