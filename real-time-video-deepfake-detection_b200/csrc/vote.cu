// Score selection + 10-frame temporal vote as device-side per-stream ring buffers.
//
// Replaces TemporalTracker (reference deepfake_detection.py:93-289) and the
// vote-input selection of /analyze (backend_server.py:160-174,206): all
// arithmetic is IEEE double in deque insertion order so verdicts and counts are
// bit-identical with the Python objects (hard part 6).
#include "dfd_internal.cuh"

struct VoteCfg {
    int blend_mode;
    double face_w, forensic_w;
};

struct TrackerParams { int window_size, voting_window; double threshold; };

// Python's builtin sum() over the score deque, bit for bit (CPython >= 3.12, Python/bltinmodule.c):
// the running total starts as int 0; while the items are exact Python floats the total is a
// Neumaier-compensated (hi, lo) pair; the first item that is not an exact float (np.float64, which is
// what np.clip returns in apply_heuristics, deepfake_detection.py:502) collapses the pair and every
// later addition is a plain double add.  is_np[k] != 0 marks such items.
// The serial arithmetic runs on a copy of the score deque in shared memory, in deque order (ScoreList), which the
// stream's warp stages with coalesced loads: walking the ring in global memory costs one dependent L2 round trip per item.
struct ScoreList {
    const double* v; const uint8_t* is_np;
    __device__ double val(int k) const { return v[k]; }
    __device__ bool np(int k) const { return is_np[k] != 0; }
};

template <class Item>
__device__ double py_sum(int n, Item item, bool any_np_forces_plain, const ScoreList& R) {
    if (n == 0) return 0.0;
    double hi = item(0), lo = 0.0;
    int k = 1;
    bool exact = any_np_forces_plain ? false : !R.np(0);
    if (exact) {
        for (; k < n; k++) {
            if (R.np(k)) break;
            double x = item(k);
            double t = __dadd_rn(hi, x);
            if (fabs(hi) >= fabs(x)) lo = __dadd_rn(lo, __dadd_rn(__dsub_rn(hi, t), x));
            else lo = __dadd_rn(lo, __dadd_rn(__dsub_rn(x, t), hi));
            hi = t;
        }
        if (lo != 0.0 && isfinite(lo)) hi = __dadd_rn(hi, lo);
    }
    for (; k < n; k++) hi = __dadd_rn(hi, item(k));
    return hi;
}

#define VOTE_WARPS 8                                 // streams per CTA: one warp each

// Warp-cooperative TemporalTracker.update + statistics: every lane of the stream's warp calls; lane 0 owns the state
// updates and the order-sensitive double arithmetic, all lanes stage the score deque.  sv / snp: this warp's
// DFD_MAX_SCORES-entry staging arrays in shared memory.  r is meaningful in lane 0.
__device__ void tracker_update(DfdStreamState& S, const VoteCfg&, double p, int p_is_np, dfd_vote_record& r,
                               double* sv, uint8_t* snp) {
    const int lane = threadIdx.x & 31;
    TrackerParams c{S.window_size, S.voting_window, S.threshold};
    if (lane == 0) {
        r.last_vote = -1; r.reserved = 0;
        if (p == p) {                                   // update(None) is ignored (:123-124)
            int slot = S.score_n < c.window_size ? (S.score_head + S.score_n) % c.window_size : S.score_head;
            S.scores[slot] = p; S.score_is_np[slot] = (uint8_t)(p_is_np != 0);
            if (S.score_n < c.window_size) S.score_n++; else S.score_head = (S.score_head + 1) % c.window_size;
            uint8_t cls = p > c.threshold ? 1 : 0;      // strict > (:135)
            r.last_vote = cls;
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
    }
    __syncwarp();                                       // lane 0's ring insertion is visible to the warp
    const int n = S.score_n, head = S.score_head;
    for (int k = lane; k < n; k += 32) {
        const int idx = (head + k) % c.window_size;
        sv[k] = S.scores[idx];
        snp[k] = S.score_is_np[idx];
    }
    __syncwarp();
    if (lane != 0) return;
    ScoreList R{sv, snp};
    bool any_np = false;
    for (int k = 0; k < n; k++) any_np |= R.np(k);
    double sum = py_sum(n, [&](int k) { return R.val(k); }, false, R);                    // sum(deque)/len (:198-202)
    r.temporal_average = n ? __ddiv_rn(sum, (double)n) : 0.0;
    if (n < 10) r.stability_score = 0.0;            // :214-221
    else {
        const double mean = r.temporal_average;
        // (x - mean) ** 2 items are np.float64 as soon as the mean is (any np score) -> plain summation
        double v = py_sum(n, [&](int k) { double d = __dsub_rn(R.val(k), mean); return __dmul_rn(d, d); }, any_np, R);
        v = __ddiv_rn(v, (double)n);
        double m4 = __dmul_rn(v, 4.0);
        r.stability_score = __dsub_rn(1.0, m4 < 1.0 ? m4 : 1.0);
    }
}

// record of a frame whose stream id is outside [0, max_streams): verdict -1, counts 0
__device__ void bad_record(dfd_vote_record& r, double p) {
    r.verdict = -1; r.fake_count = 0; r.real_count = 0; r.history_len = 0; r.frame_count = 0; r.last_vote = -1; r.reserved = 0;
    r.vote_input = p; r.temporal_average = 0.0; r.stability_score = 0.0;
}

// CTA = VOTE_WARPS streams, one warp per stream.
__global__ void __launch_bounds__(32 * VOTE_WARPS)
k_vote(int n, const int32_t* __restrict__ stream_ids, const double* __restrict__ vote_input,
       const uint8_t* __restrict__ np_flags, DfdStreamState* __restrict__ state, VoteCfg cfg,
       dfd_vote_record* __restrict__ rec, int max_streams) {
    __shared__ double s_v[VOTE_WARPS][DFD_MAX_SCORES];
    __shared__ uint8_t s_np[VOTE_WARPS][DFD_MAX_SCORES];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * VOTE_WARPS + w;
    if (i >= n) return;
    dfd_vote_record r;
    r.stream_id = stream_ids[i];
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    r.face_probability = NaN; r.forensic_probability = NaN;
    if ((unsigned)r.stream_id >= (unsigned)max_streams) {        // caller error: flagged in the record, no state touched
        if (lane == 0) { bad_record(r, vote_input[i]); rec[i] = r; }
        return;
    }
    tracker_update(state[r.stream_id], cfg, vote_input[i], np_flags ? np_flags[i] : 0, r, s_v[w], s_np[w]);
    if (lane == 0) rec[i] = r;
}

// analyze_batch tail: choose the vote input per frame, bump detector.frame_count, update the tracker.
// CTA = VOTE_WARPS frames, one warp per frame (= per stream).
__global__ void __launch_bounds__(32 * VOTE_WARPS)
k_select_vote(int n, int m, const int32_t* __restrict__ box_frame, const double* __restrict__ face_prob,
              const dfd_forensic_result* __restrict__ fres, const int32_t* __restrict__ stream_ids,
              DfdStreamState* __restrict__ state, VoteCfg cfg, dfd_vote_record* __restrict__ rec, int max_streams) {
    // first box of every frame of this CTA (faces[0], backend_server.py:160): all threads sweep the box list once and
    // keep the smallest box index per frame in shared memory (a per-thread scan of all m boxes was O(m) dependent loads)
    __shared__ int s_first[VOTE_WARPS];
    __shared__ double s_v[VOTE_WARPS][DFD_MAX_SCORES];
    __shared__ uint8_t s_np[VOTE_WARPS][DFD_MAX_SCORES];
    const int base = blockIdx.x * VOTE_WARPS;
    if (threadIdx.x < VOTE_WARPS) s_first[threadIdx.x] = 0x7fffffff;
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const int f = box_frame[j] - base;
        if (f >= 0 && f < VOTE_WARPS) atomicMin(&s_first[f], j);
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = base + w;
    if (i >= n) return;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    double fp = NaN;
    if (s_first[w] != 0x7fffffff) fp = face_prob[s_first[w]];
    double forensic = fres[i].fake_probability;
    double p;
    // README blend (opt-in): face_w * face + forensic_w * forensic in Python float arithmetic (two rounded products,
    // one rounded sum -- no FMA contraction)
    if (fp == fp) p = cfg.blend_mode == DFD_BLEND_README ? __dadd_rn(__dmul_rn(cfg.face_w, fp), __dmul_rn(cfg.forensic_w, forensic)) : fp;
    else p = forensic;                                               // backend_server.py:206
    dfd_vote_record r;
    r.stream_id = stream_ids[i];
    r.face_probability = fp;
    r.forensic_probability = forensic;
    if ((unsigned)r.stream_id >= (unsigned)max_streams) {
        if (lane == 0) { bad_record(r, p); rec[i] = r; }
        return;
    }
    DfdStreamState& S = state[r.stream_id];
    if (lane == 0) S.detector_frames += 1;                           // backend_server.py:156
    r.face_probability = fp;
    r.forensic_probability = forensic;
    tracker_update(S, cfg, p, (fp == fp) ? 1 : 0, r, s_v[w], s_np[w]);   // face prob is np.float64 (np.clip), forensic prob a Python float
    if (lane == 0) rec[i] = r;
}

// np.mean of a list of n Python floats (analyze_face_with_tta, deepfake_detection.py:441): float64 pairwise summation --
// n < 8 a plain running sum, otherwise eight accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and the tail
// added sequentially (n <= DFD_TTA_MAX_PRED < 128: no recursion) -- divided by n.
__device__ double np_mean_f64(const double* a, int n) {
    double res;
    if (n < 8) {
        res = 0.0;
        for (int i = 0; i < n; i++) res = __dadd_rn(res, a[i]);
    } else {
        double r[8];
        for (int j = 0; j < 8; j++) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], a[i + j]);
        res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; i++) res = __dadd_rn(res, a[i]);
    }
    return __ddiv_rn(res, (double)n);
}

// apply_calibration (deepfake_detection.py:445-455): calibrator.predict_proba([[p]])[0][1] for the two calibrator families
// a pickled scikit-learn object reduces to -- a logistic (Platt) map expit(coef * p + intercept), or a monotone
// piecewise-linear table (isotonic regression; np.interp semantics: clamped at the ends, slope * (p - x_j) + y_j inside).
__device__ double calibrate(double p, int kind, int n, const double* __restrict__ tab) {
    if (kind == DFD_CALIB_LOGISTIC) {
        const double z = __dadd_rn(__dmul_rn(tab[0], p), tab[1]);
        // scipy.special.expit: 1 / (1 + exp(-z)) for z >= 0, exp(z) / (1 + exp(z)) below
        if (z >= 0) return 1.0 / (1.0 + exp(-z));
        const double e = exp(z);
        return e / (1.0 + e);
    }
    if (kind == DFD_CALIB_PIECEWISE_LINEAR) {
        const double* xs = tab; const double* ys = tab + n;
        if (!(p > xs[0])) return ys[0];
        if (!(p < xs[n - 1])) return ys[n - 1];
        int lo = 0, hi = n - 1;                                  // xs[lo] <= p < xs[hi]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (xs[mid] <= p) lo = mid; else hi = mid; }
        const double slope = __ddiv_rn(__dsub_rn(ys[lo + 1], ys[lo]), __dsub_rn(xs[lo + 1], xs[lo]));
        return __dadd_rn(__dmul_rn(slope, __dsub_rn(p, xs[lo])), ys[lo]);
    }
    return p;
}

// sigmoid (+ mean over the n_pred test-time predictions of a box) + apply_calibration + apply_heuristics
// (deepfake_detection.py:398, 441, 445-455, 489-502); logits[i * n_pred + j] = prediction j of box i
__global__ void k_faceprob(int m, int n_pred, const float* __restrict__ logits, const int32_t* __restrict__ boxes,
                           const uint8_t* __restrict__ bad, int calib_kind, int calib_n, const double* __restrict__ calib,
                           double* __restrict__ prob) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (bad && bad[i]) { prob[i] = __longlong_as_double(0x7ff8000000000000LL); return; }   // rejected box: "no face"
    double p;
    if (n_pred == 1) {
        float z = logits[i];
        float s = 1.0f / (1.0f + expf(-z));                          // torch.sigmoid in float32
        p = (double)s;                                               // .item() -> Python float
    } else {
        double preds[DFD_TTA_MAX_PRED];
        for (int j = 0; j < n_pred; j++) {
            float z = logits[(size_t)i * n_pred + j];
            preds[j] = (double)(1.0f / (1.0f + expf(-z)));
        }
        p = np_mean_f64(preds, n_pred);
    }
    p = calibrate(p, calib_kind, calib_n, calib);
    int w = boxes[i * 4 + 2], h = boxes[i * 4 + 3];
    double adj = (h < 80 || w < 80) ? 0.10 : 0.0;
    p = p + adj;
    prob[i] = p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p);
}

// what: bit 0 = forensic analyzer state, bit 1 = tracker + detector.frame_count; cfg != 0 also sets parameters
__global__ void k_reset(DfdStreamState* state, int first, int count, int what, int set_cfg, int window_size,
                        int voting_window, double thr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    DfdStreamState& S = state[first + i];
    if (what & 1) { S.has_prev = 0; S.analyzer_frames = 0; S.ring_n = 0; S.ring_head = 0; }
    if (what & 2) {
        S.score_n = 0; S.score_head = 0; S.vote_n = 0; S.vote_head = 0;
        S.verdict = DFD_UNCERTAIN; S.detector_frames = 0;
    }
    if (set_cfg) { S.window_size = window_size; S.voting_window = voting_window; S.threshold = thr; }
}

static VoteCfg make_cfg(const dfd_ctx* ctx) {
    VoteCfg c;
    c.blend_mode = ctx->cfg.blend_mode;
    c.face_w = ctx->cfg.face_weight; c.forensic_w = ctx->cfg.forensic_weight;
    return c;
}

int dfd_faceprob_launch(dfd_ctx* ctx, const float* logits, const int32_t* boxes, int m, int n_pred, double* prob, cudaStream_t st) {
    // boxes the last face-prep call rejected (k_box_sanitize) get NaN; the flags describe exactly its m boxes
    const uint8_t* bad = ctx->box_flags_m == m ? ctx->d_box_bad : nullptr;
    k_faceprob<<<(m + 127) / 128, 128, 0, st>>>(m, n_pred, logits, boxes, bad, ctx->calib_kind, ctx->calib_n, (const double*)ctx->calib.p, prob);
    DFD_LAUNCH_CHECK("k_faceprob", st);
    return DFD_OK;
}

int dfd_vote_launch(dfd_ctx* ctx, const int32_t* stream_ids, const double* vote_input, const uint8_t* np_flags, int n,
                    dfd_vote_record* rec, cudaStream_t st) {
    k_vote<<<(n + VOTE_WARPS - 1) / VOTE_WARPS, 32 * VOTE_WARPS, 0, st>>>(n, stream_ids, vote_input, np_flags, ctx->d_state, make_cfg(ctx), rec, ctx->cfg.max_streams);
    DFD_LAUNCH_CHECK("k_vote", st);
    return DFD_OK;
}

int dfd_select_vote_launch(dfd_ctx* ctx, int n, int m, const int32_t* box_frame, const double* face_prob,
                           const dfd_forensic_result* fres, const int32_t* stream_ids, dfd_vote_record* rec,
                           cudaStream_t st) {
    k_select_vote<<<(n + VOTE_WARPS - 1) / VOTE_WARPS, 32 * VOTE_WARPS, 0, st>>>(n, m, box_frame, face_prob, fres, stream_ids, ctx->d_state, make_cfg(ctx), rec, ctx->cfg.max_streams);
    DFD_LAUNCH_CHECK("k_select_vote", st);
    return DFD_OK;
}

int dfd_reset_launch(dfd_ctx* ctx, int stream_id, int what, cudaStream_t st) {
    int first = stream_id < 0 ? 0 : stream_id, count = stream_id < 0 ? ctx->cfg.max_streams : 1;
    k_reset<<<(count + 127) / 128, 128, 0, st>>>(ctx->d_state, first, count, what, 0, 0, 0, 0.0);
    DFD_LAUNCH_CHECK("k_reset", st);
    return DFD_OK;
}

int dfd_configure_launch(dfd_ctx* ctx, int stream_id, int window_size, int voting_window, double thr, cudaStream_t st) {
    int first = stream_id < 0 ? 0 : stream_id, count = stream_id < 0 ? ctx->cfg.max_streams : 1;
    k_reset<<<(count + 127) / 128, 128, 0, st>>>(ctx->d_state, first, count, 2, 1, window_size, voting_window, thr);
    DFD_LAUNCH_CHECK("k_reset", st);
    return DFD_OK;
}
