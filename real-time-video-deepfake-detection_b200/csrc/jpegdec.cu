// Frame ingest on the device: cv2.imdecode(IMREAD_COLOR) of the /analyze wire format (baseline JPEG; reference
// backend_server.py:140-142, extension/content.js:86-109) for a batch of frames, so that the host-to-device copy carries the
// ~100-400 KB JPEG stream instead of the 2.76 MB raw 720p frame (raw frames make `e2e` PCIe-bound).
// Bit-exact with OpenCV / libjpeg-turbo (islow IDCT, fancy up-sampling, fixed-point YCbCr): the pixel functions are those of
// px_jpeg.h / px_jpegdec.h, which tests/hostcheck runs on the CPU against cv2.imdecode.
//
//   host            dfd_jpeg_parse: markers -> DfdJpegHeader (quantisers, Huffman lookup tables, geometry); no pixel work
//   k_jpeg_unstuff  CTA per frame: removes the FF 00 byte stuffing of the entropy-coded segment (block scan + scatter)
//   k_jpeg_huffman  CTA per frame: SELF-SYNCHRONISING parallel Huffman decode.  The bit stream is cut into subsequences of
//                   1024 bits, one per thread (strided).  (1) every thread decodes its subsequence blindly from the state
//                   "a block starts here"; Huffman streams resynchronise, so most end states are already right.  (2) rounds:
//                   thread i re-decodes subsequence i from the END state of subsequence i-1 whenever that changed, until a
//                   whole round changes nothing -- by induction from subsequence 0 (whose start is known) every state is then
//                   exact.  (3) a block scan of the per-subsequence block counts gives every subsequence its first block
//                   number; (4) a last pass writes the coefficients (AC values in natural order, DC differences in
//                   prediction-chain order); (5) the DC prediction chains are prefix sums per component.
//   k_jpeg_idct     thread per 8 x 8 block: dequantise, jidctint (columns, rows), +128, clamp -> component planes
//   k_jpeg_color    fancy h2v2 / h2v1 chroma up-sampling + YCbCr -> BGR into the caller's frame buffer
#include "dfd_internal.cuh"
#include <string.h>
#include <stdlib.h>
#include "px_jpegdec.h"

#define JPG_SUB_BITS 1024
#define JPG_THREADS 512

struct JpgMeta {
    long long raw_off;        // first byte of the stream in the raw buffer
    long long words_off;      // first 32-bit word of the frame's clean bit stream
    long long sub_off;        // first subsequence slot of the frame
    int ecs_bytes;            // raw size of the entropy-coded segment
    int pad;
};

// exclusive scan of one int per thread over the CTA (JPG_THREADS threads); returns the exclusive prefix, *total = CTA sum
__device__ int cta_exscan(int v, int* total, int* s_warp /* [JPG_THREADS / 32 + 1] */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __syncthreads();                                           // s_warp may still be read from the previous call
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        int t = lane < JPG_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
        if (lane < JPG_THREADS / 32) s_warp[lane] = t;         // inclusive warp totals
    }
    __syncthreads();
    const int base = w ? s_warp[w - 1] : 0;
    *total = s_warp[JPG_THREADS / 32 - 1];
    return base + x - v;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JPG_THREADS)
k_jpeg_unstuff(const uint8_t* __restrict__ raw, const JpgMeta* __restrict__ meta, const DfdJpegHeader* __restrict__ hdr,
               uint32_t* __restrict__ words, uint32_t* __restrict__ nbits_out) {
    __shared__ int s_warp[JPG_THREADS / 32 + 1];
    __shared__ int s_base;
    const int f = blockIdx.x;
    const JpgMeta M = meta[f];
    const uint8_t* src = raw + M.raw_off + hdr[f].ecs_begin;
    uint8_t* dst = (uint8_t*)(words + M.words_off);
    const int n = M.ecs_bytes;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += JPG_THREADS * 8) {
        const int i0 = c0 + threadIdx.x * 8;
        uint8_t b[8];
        int keep = 0;
        uint8_t prev = (i0 > 0 && i0 - 1 < n) ? src[i0 - 1] : 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int i = i0 + j;
            b[j] = i < n ? src[i] : 0;
            const bool stuffed = b[j] == 0 && prev == 0xFF;
            if (i < n && !stuffed) keep |= 1 << j;
            prev = b[j];
        }
        int total;
        int off = s_base + cta_exscan(__popc(keep), &total, s_warp);
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (keep >> j & 1) { dst[(off & ~3) + (3 - (off & 3))] = b[j]; off++; }        // MSB-first inside each 32-bit word
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
    // zero the tail of the last word and two guard words (dfd_peek16 reads one word ahead)
    const int nb = s_base;
    if (threadIdx.x < 12) {
        const int o = nb + threadIdx.x;
        if (o < ((nb + 3) & ~3) + 8) dst[(o & ~3) + (3 - (o & 3))] = 0;
    }
    if (threadIdx.x == 0) nbits_out[f] = (uint32_t)nb * 8u;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JPG_THREADS)
k_jpeg_huffman(const JpgMeta* __restrict__ meta, const DfdJpegHeader* __restrict__ hdr, const uint32_t* __restrict__ words_all,
               const uint32_t* __restrict__ nbits_in, unsigned long long* __restrict__ E_all, unsigned long long* __restrict__ used_all,
               int* __restrict__ cnt_all, int* __restrict__ blk0_all, int16_t* __restrict__ coef_all, int32_t* __restrict__ dc_all,
               long long blocks_stride, int32_t* __restrict__ status) {
    __shared__ DfdJpegHeader h;
    __shared__ int s_warp[JPG_THREADS / 32 + 1];
    __shared__ int s_carry;
    const int f = blockIdx.x, tid = threadIdx.x;
    {
        const uint32_t* src = (const uint32_t*)(hdr + f);
        uint32_t* d = (uint32_t*)&h;
        for (int i = tid; i < (int)(sizeof(DfdJpegHeader) / 4); i += JPG_THREADS) d[i] = src[i];
    }
    __syncthreads();
    const JpgMeta M = meta[f];
    const uint32_t* words = words_all + M.words_off;
    const uint32_t nbits = nbits_in[f];
    const uint32_t nwords = (nbits + 31) / 32 + 2;
    const int nsub = (int)((nbits + JPG_SUB_BITS - 1) / JPG_SUB_BITS);
    volatile unsigned long long* E = E_all + M.sub_off;
    volatile unsigned long long* used = used_all + M.sub_off;
    int* cnt = cnt_all + M.sub_off;
    int* blk0 = blk0_all + M.sub_off;
    int16_t* coef = coef_all + (size_t)f * blocks_stride * 64;
    int32_t* dcd = dc_all + (size_t)f * blocks_stride;
    int32_t dc_off[3] = {0, 0, 0};
    { int a = 0; for (int c = 0; c < h.ncomp; c++) { dc_off[c] = a; a += h.comp_bw[c] * h.comp_bh[c]; } }
    int nerr_total = 0;

    // (1) blind pass
    for (int i = tid; i < nsub; i += JPG_THREADS) {
        DfdJpegState s0; s0.p = (uint32_t)i * JPG_SUB_BITS; s0.c = 0; s0.z = 0;
        const unsigned long long st = dfd_jpeg_pack_state(s0);
        const uint32_t lim = min((uint32_t)(i + 1) * JPG_SUB_BITS, nbits);
        int nb, ne;
        const unsigned long long e = dfd_jpeg_decode_sub<false>(&h, words, nwords, st, lim, &nb, &ne, 0, nullptr, nullptr, nullptr);
        used[i] = st; E[i] = e; cnt[i] = nb;
    }
    // (2) synchronisation rounds: subsequence i restarts from the end state of i-1 until nothing changes.  States are single
    // 64-bit words, so a concurrent update is seen whole or not at all; a round that changes no state is a consistent
    // snapshot in which every used[i] equals E[i-1], and subsequence 0 starts from the true state: by induction all are exact.
    for (int round = 0; round <= nsub + 1; round++) {
        __syncthreads();
        int changed = 0;
        for (int i = tid; i < nsub; i += JPG_THREADS) {
            if (i == 0) continue;
            const unsigned long long st = E[i - 1];
            if (st == used[i]) continue;
            const uint32_t lim = min((uint32_t)(i + 1) * JPG_SUB_BITS, nbits);
            int nb, ne;
            const unsigned long long e = dfd_jpeg_decode_sub<false>(&h, words, nwords, st, lim, &nb, &ne, 0, nullptr, nullptr, nullptr);
            if (e != E[i]) changed = 1;
            used[i] = st; E[i] = e; cnt[i] = nb;
        }
        if (!__syncthreads_or(changed)) break;
    }
    // (3) first block number of every subsequence: exclusive scan of the block counts, in index order
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < nsub; c0 += JPG_THREADS) {
        const int i = c0 + tid;
        const int v = i < nsub ? cnt[i] : 0;
        int total;
        const int ex = cta_exscan(v, &total, s_warp);
        if (i < nsub) blk0[i] = s_carry + ex;
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    const int total_blocks = s_carry;
    // (4) write pass
    for (int i = tid; i < nsub; i += JPG_THREADS) {
        const uint32_t lim = min((uint32_t)(i + 1) * JPG_SUB_BITS, nbits);
        int nb, ne;
        dfd_jpeg_decode_sub<true>(&h, words, nwords, used[i], lim, &nb, &ne, blk0[i], coef, dcd, dc_off);
        nerr_total += ne;
    }
    const int any_err = __syncthreads_or(nerr_total != 0 && false);      // (invalid codes in the padding after the last block are legal)
    (void)any_err;
    // (5) DC prediction chains: inclusive prefix sum per component, in chain order
    for (int c = 0; c < h.ncomp; c++) {
        const int n = h.comp_bw[c] * h.comp_bh[c];
        int32_t* d = dcd + dc_off[c];
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (int c0 = 0; c0 < n; c0 += JPG_THREADS) {
            const int i = c0 + tid;
            const int v = i < n ? d[i] : 0;
            int total;
            const int ex = cta_exscan(v, &total, s_warp);
            if (i < n) d[i] = s_carry + ex + v;
            __syncthreads();
            if (tid == 0) s_carry += total;
            __syncthreads();
        }
    }
    if (tid == 0) status[f] = total_blocks >= h.total_blocks ? DFD_JPEG_OK : DFD_JPEG_ERR_DATA;
}

// ---------------------------------------------------------------------------------------------
// thread = one 8x8 block.  Planes of frame f: component c at plane_base + plane_off[c], pitch comp_bw[c] * 8.
__global__ void __launch_bounds__(128)
k_jpeg_idct(const DfdJpegHeader* __restrict__ hdr, const int16_t* __restrict__ coef_all, const int32_t* __restrict__ dc_all,
            long long blocks_stride, uint8_t* __restrict__ planes_all, long long plane_stride) {
    const int f = blockIdx.y;
    const DfdJpegHeader* h = hdr + f;
    const int blk = blockIdx.x * 128 + threadIdx.x;
    if (blk >= h->total_blocks) return;
    int c = 0;
    if (h->ncomp == 3) c = blk >= h->comp_blk0[2] ? 2 : (blk >= h->comp_blk0[1] ? 1 : 0);
    const int j = blk - h->comp_blk0[c];
    const int by = j / h->comp_bw[c], bx = j - by * h->comp_bw[c];
    int dc_off = 0;
    for (int q = 0; q < c; q++) dc_off += h->comp_bw[q] * h->comp_bh[q];
    const int16_t* cf = coef_all + ((size_t)f * blocks_stride + blk) * 64;
    const int dc = dc_all[(size_t)f * blocks_stride + dc_off + dfd_jpeg_dc_seq(h, c, bx, by)];
    int b[64];
    const uint4* cv = (const uint4*)cf;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint4 v = cv[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            b[i * 8 + 2 * k] = (int)(short)(w[k] & 0xffffu) * (int)h->qt[c][i * 8 + 2 * k];
            b[i * 8 + 2 * k + 1] = (int)(short)(w[k] >> 16) * (int)h->qt[c][i * 8 + 2 * k + 1];
        }
    }
    b[0] = dc * (int)h->qt[c][0];
#pragma unroll
    for (int q = 0; q < 8; q++) dfd_idct8(b + q, 8, 1);
#pragma unroll
    for (int r = 0; r < 8; r++) dfd_idct8(b + 8 * r, 1, 0);
    size_t poff = 0;
    for (int q = 0; q < c; q++) poff += (size_t)h->comp_bw[q] * h->comp_bh[q] * 64;
    const int pitch = h->comp_bw[c] * 8;
    uint8_t* out = planes_all + (size_t)f * plane_stride + poff + (size_t)(by * 8) * pitch + bx * 8;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            lo |= (uint32_t)dfd_sat_u8(b[r * 8 + k] + 128) << (8 * k);
            hi |= (uint32_t)dfd_sat_u8(b[r * 8 + 4 + k] + 128) << (8 * k);
        }
        *(uint2*)(out + (size_t)r * pitch) = make_uint2(lo, hi);
    }
}

// thread = 4 horizontally adjacent output pixels
__global__ void __launch_bounds__(256)
k_jpeg_color(const DfdJpegHeader* __restrict__ hdr, const uint8_t* __restrict__ planes_all, long long plane_stride,
             uint8_t* __restrict__ frames, size_t frame_stride, int row_pitch, int H, int W) {
    const int f = blockIdx.z;
    const DfdJpegHeader* h = hdr + f;
    const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (y >= H || x0 >= W) return;
    const uint8_t* P = planes_all + (size_t)f * plane_stride;
    const int pitch0 = h->comp_bw[0] * 8;
    uint8_t px[12];
    const uint8_t* p1 = nullptr;
    const uint8_t* p2 = nullptr;
    int pitch1 = 0, cw = 0, chh = 0;
    if (h->ncomp == 3) {
        p1 = P + (size_t)h->comp_bw[0] * h->comp_bh[0] * 64;
        p2 = p1 + (size_t)h->comp_bw[1] * h->comp_bh[1] * 64;
        pitch1 = h->comp_bw[1] * 8;
        cw = (W * h->hs[1] + h->hmax - 1) / h->hmax; chh = (H * h->vs[1] + h->vmax - 1) / h->vmax;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = x0 + i;
        int r = 0, g = 0, b = 0;
        if (x < W) {
            const int Y = P[(size_t)y * pitch0 + x];
            r = g = b = Y;
            if (h->ncomp == 3) {
                const int cb = dfd_jpeg_chroma_at(p1, pitch1, cw, chh, h->hs[1], h->vs[1], h->hmax, h->vmax, x, y);
                const int cr = dfd_jpeg_chroma_at(p2, pitch1, cw, chh, h->hs[2], h->vs[2], h->hmax, h->vmax, x, y);
                dfd_jpeg_ycc2rgb(Y, cb, cr, &r, &g, &b);
            }
        }
        px[3 * i] = (uint8_t)b; px[3 * i + 1] = (uint8_t)g; px[3 * i + 2] = (uint8_t)r;
    }
    uint8_t* o = frames + (size_t)f * frame_stride + (size_t)y * row_pitch + (size_t)x0 * 3;
    if (x0 + 4 <= W && ((uintptr_t)o & 3) == 0) {
        uint32_t w[3];
#pragma unroll
        for (int k = 0; k < 3; k++) w[k] = px[4 * k] | (px[4 * k + 1] << 8) | (px[4 * k + 2] << 16) | ((uint32_t)px[4 * k + 3] << 24);
        ((uint32_t*)o)[0] = w[0]; ((uint32_t*)o)[1] = w[1]; ((uint32_t*)o)[2] = w[2];
    } else {
        for (int i = 0; i < 4 && x0 + i < W; i++) { o[3 * i] = px[3 * i]; o[3 * i + 1] = px[3 * i + 1]; o[3 * i + 2] = px[3 * i + 2]; }
    }
}

// ---------------------------------------------------------------------------------------------
struct JpgHost {
    std::vector<DfdJpegHeader> hdr;
    std::vector<JpgMeta> meta;
    DfdJpegHeader* h_hdr_pinned = nullptr;
    JpgMeta* h_meta_pinned = nullptr;
    size_t pinned_n = 0;
};

int dfd_jpeg_decode_launch(dfd_ctx* ctx, const uint8_t* bytes_host, const int64_t* offsets_host, int n, int H, int W,
                           uint8_t* frames_out, size_t frame_stride, int row_pitch, int32_t* status_dev, cudaStream_t st) {
    DFD_REQUIRE(n > 0 && n <= ctx->cfg.max_batch, DFD_ERR_CAPACITY, "decode_jpeg: batch exceeds max_batch");
    DFD_REQUIRE(H >= 1 && W >= 1 && H <= 16384 && W <= 16384 && row_pitch >= 3 * W, DFD_ERR_INVALID, "decode_jpeg: bad frame geometry");
    if (!ctx->jpg_host) ctx->jpg_host = new JpgHost;
    JpgHost* J = (JpgHost*)ctx->jpg_host;
    if (J->pinned_n < (size_t)n) {
        if (J->h_hdr_pinned) { cudaFreeHost(J->h_hdr_pinned); cudaFreeHost(J->h_meta_pinned); }
        DFD_CUDA(cudaHostAlloc((void**)&J->h_hdr_pinned, sizeof(DfdJpegHeader) * ctx->cfg.max_batch, cudaHostAllocDefault));
        DFD_CUDA(cudaHostAlloc((void**)&J->h_meta_pinned, sizeof(JpgMeta) * ctx->cfg.max_batch, cudaHostAllocDefault));
        J->pinned_n = ctx->cfg.max_batch;
    }
    // the staging arrays are reused by the next call: the previous call's copies must have left them
    if (ctx->jpg_ev) DFD_CUDA(cudaEventSynchronize(ctx->jpg_ev));
    else DFD_CUDA(cudaEventCreateWithFlags(&ctx->jpg_ev, cudaEventDisableTiming));
    // ---- host: headers only ----
    long long words = 0, subs = 0;
    const long long total_bytes = offsets_host[n] - offsets_host[0];
    for (int i = 0; i < n; i++) {
        const long long b0 = offsets_host[i], b1 = offsets_host[i + 1];
        DFD_REQUIRE(b1 > b0, DFD_ERR_INVALID, "decode_jpeg: empty stream");
        DfdJpegHeader* h = &J->h_hdr_pinned[i];
        const int rc = dfd_jpeg_parse(bytes_host + b0, (size_t)(b1 - b0), h);
        if (rc == DFD_JPEG_ERR_UNSUPPORTED) {
            ctx->err = "decode_jpeg: frame " + std::to_string(i) + " is not a baseline 8-bit 4:4:4 / 4:2:2 / 4:2:0 / gray JPEG without restart markers";
            return DFD_ERR_UNSUPPORTED;
        }
        if (rc != DFD_JPEG_OK) { ctx->err = "decode_jpeg: frame " + std::to_string(i) + " is not a valid JPEG stream"; return DFD_ERR_INVALID; }
        if (h->height != H || h->width != W) {
            ctx->err = "decode_jpeg: frame " + std::to_string(i) + " is " + std::to_string(h->width) + "x" + std::to_string(h->height) + ", the batch is " +
                       std::to_string(W) + "x" + std::to_string(H);
            return DFD_ERR_INVALID;
        }
        JpgMeta& m = J->h_meta_pinned[i];
        m.raw_off = b0 - offsets_host[0];
        m.ecs_bytes = h->ecs_end - h->ecs_begin;
        m.words_off = words;
        m.sub_off = subs;
        m.pad = 0;
        words += (m.ecs_bytes + 3) / 4 + 4;
        subs += ((long long)m.ecs_bytes * 8 + JPG_SUB_BITS - 1) / JPG_SUB_BITS + 1;
    }
    // workspaces (worst case 4:4:4 with 16-pixel MCU padding: 3 components of ceil16(H) x ceil16(W))
    const long long bw = (W + 15) / 16 * 2, bh = (H + 15) / 16 * 2;
    const long long blocks_stride = 3 * bw * bh;
    const long long plane_stride = blocks_stride * 64;
    int rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_raw, (size_t)total_bytes + 16))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_words, (size_t)words * 4 + 64))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_sub, (size_t)subs * (8 + 8 + 4 + 4) + 64))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_coef, (size_t)n * blocks_stride * 64 * sizeof(int16_t)))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_dc, (size_t)n * blocks_stride * sizeof(int32_t)))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_planes, (size_t)n * plane_stride))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_hdr, (size_t)ctx->cfg.max_batch * (sizeof(DfdJpegHeader) + sizeof(JpgMeta) + 8)))) return rc;
    DfdJpegHeader* d_hdr = (DfdJpegHeader*)ctx->jpg_hdr.p;
    JpgMeta* d_meta = (JpgMeta*)((uint8_t*)ctx->jpg_hdr.p + (size_t)ctx->cfg.max_batch * sizeof(DfdJpegHeader));
    uint32_t* d_nbits = (uint32_t*)((uint8_t*)d_meta + (size_t)ctx->cfg.max_batch * sizeof(JpgMeta));
    unsigned long long* d_E = (unsigned long long*)ctx->jpg_sub.p;
    unsigned long long* d_used = d_E + subs;
    int* d_cnt = (int*)(d_used + subs);
    int* d_blk0 = d_cnt + subs;
    // ---- device ----
    DFD_CUDA(cudaMemcpyAsync(ctx->jpg_raw.p, bytes_host + offsets_host[0], (size_t)total_bytes, cudaMemcpyHostToDevice, st));
    DFD_CUDA(cudaMemcpyAsync(d_hdr, J->h_hdr_pinned, sizeof(DfdJpegHeader) * n, cudaMemcpyHostToDevice, st));
    DFD_CUDA(cudaMemcpyAsync(d_meta, J->h_meta_pinned, sizeof(JpgMeta) * n, cudaMemcpyHostToDevice, st));
    DFD_CUDA(cudaEventRecord(ctx->jpg_ev, st));
    DFD_CUDA(cudaMemsetAsync(ctx->jpg_coef.p, 0, (size_t)n * blocks_stride * 64 * sizeof(int16_t), st));
    DFD_CUDA(cudaMemsetAsync(ctx->jpg_dc.p, 0, (size_t)n * blocks_stride * sizeof(int32_t), st));
    k_jpeg_unstuff<<<n, JPG_THREADS, 0, st>>>((const uint8_t*)ctx->jpg_raw.p, d_meta, d_hdr, (uint32_t*)ctx->jpg_words.p, d_nbits);
    DFD_LAUNCH_CHECK("k_jpeg_unstuff", st);
    k_jpeg_huffman<<<n, JPG_THREADS, 0, st>>>(d_meta, d_hdr, (const uint32_t*)ctx->jpg_words.p, d_nbits, d_E, d_used, d_cnt, d_blk0,
                                              (int16_t*)ctx->jpg_coef.p, (int32_t*)ctx->jpg_dc.p, blocks_stride, status_dev);
    DFD_LAUNCH_CHECK("k_jpeg_huffman", st);
    k_jpeg_idct<<<dim3((unsigned)((blocks_stride + 127) / 128), n), 128, 0, st>>>(d_hdr, (const int16_t*)ctx->jpg_coef.p, (const int32_t*)ctx->jpg_dc.p,
                                                                                  blocks_stride, (uint8_t*)ctx->jpg_planes.p, plane_stride);
    DFD_LAUNCH_CHECK("k_jpeg_idct", st);
    k_jpeg_color<<<dim3((W + 255) / 256, (H + 3) / 4, n), 256, 0, st>>>(d_hdr, (const uint8_t*)ctx->jpg_planes.p, plane_stride, frames_out,
                                                                        frame_stride, row_pitch, H, W);
    DFD_LAUNCH_CHECK("k_jpeg_color", st);
    return DFD_OK;
}

void dfd_jpeg_free(dfd_ctx* ctx) {
    JpgHost* J = (JpgHost*)ctx->jpg_host;
    if (J) {
        if (J->h_hdr_pinned) { cudaFreeHost(J->h_hdr_pinned); cudaFreeHost(J->h_meta_pinned); }
        delete J;
        ctx->jpg_host = nullptr;
    }
    if (ctx->jpg_ev) { cudaEventDestroy(ctx->jpg_ev); ctx->jpg_ev = nullptr; }
}

// host-only header peek (no context, no device): dims and sampling of a stream, or a negative DFD_JPEG_* code
extern "C" int dfd_jpeg_info(const uint8_t* bytes_host, size_t n, int32_t* info /* H, W, components, luma h, luma v */) {
    if (!bytes_host || !info) return DFD_ERR_INVALID;
    DfdJpegHeader* h = new DfdJpegHeader;
    const int rc = dfd_jpeg_parse(bytes_host, n, h);
    info[0] = h->height; info[1] = h->width; info[2] = h->ncomp; info[3] = h->hs[0]; info[4] = h->vs[0];
    delete h;
    return rc;
}
