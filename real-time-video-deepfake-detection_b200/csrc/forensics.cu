// Six frame-forensic signals as batched sm_100a kernels.
//
// Replaces FrameForensicAnalyzer.analyze / analyze_fast (reference
// frame_analysis.py:58-126) for a batch of frames, one per stream:
//
//   k_resize256   cv2.resize(frame,(256,256),INTER_LINEAR) + BGR2GRAY          :71,111,136
//   k_tile_stats  noise residual (:188-202), Laplacian sums (:292), HSV sums (:318-341),
//                 temporal |gray - prev| (:356-366), per 32x32 block, exact integers
//   k_canny       cv2.Canny(gray,50,150) edge count (:288-289): NMS + bit-packed hysteresis in smem
//   k_ela         JPEG Q90 4:2:0 round trip + absdiff + gray block sums (:234-253), all in smem
//   k_fft_rows / k_fft_cols   np.fft.fft2 -> fftshift -> log1p|.| band sums (:139-152,168-170)
//   k_finalize    thresholds, weighted sum, per-stream temporal ring update (:154-180,204-225,...)
//
// The integer stages are bit-exact with cv2 (same px_*.h functions that tests/hostcheck checks on
// the CPU); float statistics agree to ~1e-6 relative (double accumulation vs NumPy float32 pairwise).
#include "dfd_internal.cuh"
#include <stdlib.h>
#include "px_resize.h"
#include "px_jpeg.h"
#include "px_canny.h"
#include "px_numpy.h"

#define T 256

// ---------------------------------------------------------------------------------------------
// CTA = one output row of one frame.  The fixed-point tap positions / weights depend only on (H, W): they are tabulated on
// the host once per frame size (dfd_resize_tables), so the kernel has no double-precision math.  When the two horizontal
// taps are adjacent pixels (always, when down-scaling) their 6 bytes are fetched as three aligned 32-bit words per row
// instead of six byte loads.  (Staging the source rows in shared memory -- 16-byte loads or cp.async.bulk, 1 or 4 output
// rows per CTA -- was measured and is not faster: the kernel streams 2 of every 2.8 frame rows at 2.6-2.8 TB/s either way.)
struct RsTap { int s0, s1, a0, a1; };

__device__ __forceinline__ void rs_fetch6(const uint8_t* row, int byte_off, int* p0, int* p1) {
    const uint32_t* w = (const uint32_t*)(row + (byte_off & ~3));
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const int sh = (byte_off & 3) * 8;
    const uint32_t lo = __funnelshift_r(w0, w1, sh);       // bytes 0..3 of the window
    const uint32_t hi = __funnelshift_r(w1, w2, sh);       // bytes 4..7
    p0[0] = lo & 255; p0[1] = (lo >> 8) & 255; p0[2] = (lo >> 16) & 255;
    p1[0] = lo >> 24; p1[1] = hi & 255; p1[2] = (hi >> 8) & 255;
}

__global__ void __launch_bounds__(256) k_resize256(const uint8_t* __restrict__ frames, int H, int W, size_t fstride,
                                                   int pitch, const RsTap* __restrict__ tab, int wide_ok,
                                                   uint8_t* __restrict__ tile, uint8_t* __restrict__ gray) {
    const int n = blockIdx.y, y = blockIdx.x, x = threadIdx.x;
    const RsTap ty = tab[y], tx = tab[T + x];
    const uint8_t* f = frames + (size_t)n * fstride;
    const uint8_t* r0 = f + (size_t)ty.s0 * pitch;
    const uint8_t* r1 = f + (size_t)ty.s1 * pitch;
    int p00[3], p01[3], p10[3], p11[3];
    // the 12-byte window of rs_fetch6 must stay inside the row (last pixels: byte loads)
    if (wide_ok && tx.s1 == tx.s0 + 1 && ((tx.s0 * 3) & ~3) + 12 <= W * 3) {
        rs_fetch6(r0, tx.s0 * 3, p00, p01);
        rs_fetch6(r1, tx.s0 * 3, p10, p11);
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            p00[c] = __ldg(r0 + tx.s0 * 3 + c); p01[c] = __ldg(r0 + tx.s1 * 3 + c);
            p10[c] = __ldg(r1 + tx.s0 * 3 + c); p11[c] = __ldg(r1 + tx.s1 * 3 + c);
        }
    }
    int px[3];
#pragma unroll
    for (int c = 0; c < 3; c++) px[c] = dfd_cvresize_px(p00[c], p01[c], p10[c], p11[c], tx.a0, tx.a1, ty.a0, ty.a1);
    size_t o = ((size_t)n * T + y) * T + x;
    tile[o * 3 + 0] = (uint8_t)px[0];
    tile[o * 3 + 1] = (uint8_t)px[1];
    tile[o * 3 + 2] = (uint8_t)px[2];
    gray[o] = (uint8_t)dfd_bgr2gray(px[0], px[1], px[2]);
}

// host: tap table for an H x W frame, [0..255] vertical, [256..511] horizontal.  Tables are cached per frame size and never
// rewritten while cached (a kernel of an earlier call may still be reading one); the first call with a new size uploads
// synchronously, so CUDA-graph capture needs one warm-up call per frame size (Engine.capture_step does that).
static int dfd_resize_tables(dfd_ctx* ctx, int H, int W, const RsTap** out) {
    for (auto& e : ctx->rs_cache)
        if (e.h == H && e.w == W) { *out = (const RsTap*)e.p; return DFD_OK; }
    RsTap h[2 * T];
    for (int i = 0; i < T; i++) {
        dfd_cvresize_coef(i, H, T, 0, &h[i].s0, &h[i].s1, &h[i].a0, &h[i].a1);
        dfd_cvresize_coef(i, W, T, 1, &h[T + i].s0, &h[T + i].s1, &h[T + i].a0, &h[T + i].a1);
    }
    dfd_ctx::RsEntry e;
    e.h = H; e.w = W; e.p = nullptr;
    if (ctx->rs_cache.size() >= 32) {                       // recycle the oldest table once nothing can be using it
        DFD_CUDA(cudaDeviceSynchronize());
        e.p = ctx->rs_cache.front().p;
        ctx->rs_cache.erase(ctx->rs_cache.begin());
    } else {
        DFD_CUDA(cudaMalloc(&e.p, sizeof h));
    }
    DFD_CUDA(cudaMemcpy(e.p, h, sizeof h, cudaMemcpyHostToDevice));
    ctx->rs_cache.push_back(e);
    *out = (const RsTap*)e.p;
    return DFD_OK;
}

// ---------------------------------------------------------------------------------------------
template <typename V>
__device__ __forceinline__ V warp_sum(V v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block = 256 threads (8 warps); every thread calls; result valid in thread 0
template <typename V>
__device__ __forceinline__ V block_sum_256(V v, V* sh /* [8] */) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    V r = 0;
    if (threadIdx.x == 0)
        for (int i = 0; i < 8; i++) r += sh[i];
    return r;
}

// CTA = one 32x32 block of one frame.  A thread owns 4 vertically adjacent pixels of one column, so the 5x5 Gaussian runs
// separably on an 8x5 register window (40 shared loads for 4 pixels instead of 100) and the Laplacian reads the same
// registers.  Every sum is an exact integer: warp sums use redux.sync on 32-bit pieces (the one 64-bit quantity, the sum
// of squared noise residuals, is reduced as 16-bit halves), one barrier, then 11 threads add the 8 warp partials.
__global__ void __launch_bounds__(256, 6) k_tile_stats(const uint8_t* __restrict__ tile, const uint8_t* __restrict__ gray,
                                                    const int32_t* __restrict__ stream_ids,
                                                    const uint8_t* __restrict__ full, const DfdColorTables* __restrict__ tab,
                                                    const DfdStreamState* __restrict__ state, uint8_t* __restrict__ prev_gray,
                                                    DfdFramePartials* __restrict__ part, int max_streams) {
    const int n = blockIdx.y, blk = blockIdx.x;
    const int by = blk >> 3, bx = blk & 7;
    const bool is_full = full[n] != 0;
    // an id outside [0, max_streams) is a caller error: k_finalize flags the record; here it must only not fault
    const int sid = min(max(stream_ids[n], 0), max_streams - 1);
    __shared__ uint8_t sg[36][40];
    __shared__ int wpart[8][12];
    __shared__ unsigned int shue[6];
    __shared__ __align__(16) int s_hsv[512];                // sdiv[256], hdiv180[256]: the first 2 KB of DfdColorTables
    if (is_full)
        for (int i = threadIdx.x; i < 128; i += 256) ((uint4*)s_hsv)[i] = ((const uint4*)tab)[i];
    const DfdColorTables* stab = (const DfdColorTables*)s_hsv;     // dfd_bgr2hsv touches sdiv / hdiv180 only
    // The CTA's work is a few hundred instructions; what it costs is its chain of dependent global round trips.  Every
    // global load of the thread (halo bytes, previous gray, colour pixels) is therefore issued up front, before the first
    // shared-memory store and the barrier: {full, stream id} -> {halo, prev, tile} -> compute, two round trips instead of four.
    const uint8_t* g = gray + (size_t)n * T * T;
    uint8_t* prev = prev_gray + (size_t)sid * T * T;
    const int lx = threadIdx.x & 31, ly0 = (threadIdx.x >> 5) * 4, warp = threadIdx.x >> 5;
    const size_t g0 = (size_t)(by * 32 + ly0) * T + bx * 32 + lx;
    uint8_t hv[6];
#pragma unroll
    for (int u = 0; u < 6; u++) {
        const int i = threadIdx.x + u * 256;
        hv[u] = 0;
        if (i < 36 * 36) {
            const int hy = i / 36, hx = i - hy * 36;
            const int gy = dfd_reflect101(by * 32 + hy - 2, T), gx = dfd_reflect101(bx * 32 + hx - 2, T);
            hv[u] = g[gy * T + gx];
        }
    }
    int pv[4]; uint8_t t3[4][3];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        pv[k] = prev[g0 + (size_t)k * T];
        t3[k][0] = t3[k][1] = t3[k][2] = 0;
        if (is_full) {
            const uint8_t* tp = tile + ((size_t)n * T * T + g0 + (size_t)k * T) * 3;
            t3[k][0] = tp[0]; t3[k][1] = tp[1]; t3[k][2] = tp[2];
        }
    }
#pragma unroll
    for (int u = 0; u < 6; u++) {
        const int i = threadIdx.x + u * 256;
        if (i < 36 * 36) (&sg[0][0])[(i / 36) * 40 + (i % 36)] = hv[u];
    }
    if (threadIdx.x < 6) shue[threadIdx.x] = 0;
    __syncthreads();
    int ls = 0, lss = 0, td = 0, nsx = 0, ss = 0, sss = 0, vs = 0, vss = 0;
    unsigned nsxx_lo = 0, nsxx_hi = 0;
    if (is_full) {
        int ctr[8][3];                              // columns lx+1..lx+3 of window rows 0..7 (window row j = block row ly0 + j - 2)
        int hrow[8];                                // horizontal [1 4 6 4 1] sums of the 8 window rows
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int a0 = sg[ly0 + j][lx], a1 = sg[ly0 + j][lx + 1], a2 = sg[ly0 + j][lx + 2], a3 = sg[ly0 + j][lx + 3], a4 = sg[ly0 + j][lx + 4];
            ctr[j][0] = a1; ctr[j][1] = a2; ctr[j][2] = a3;
            hrow[j] = a0 + 4 * a1 + 6 * a2 + 4 * a3 + a4;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = ctr[k + 2][1];
            const int lap = ctr[k + 1][1] + ctr[k + 3][1] + ctr[k + 2][0] + ctr[k + 2][2] - 4 * c;
            ls += lap; lss += lap * lap;
            td += dfd_absi(c - pv[k]);
            prev[g0 + (size_t)k * T] = (uint8_t)c;
            const int acc = hrow[k] + 4 * hrow[k + 1] + 6 * hrow[k + 2] + 4 * hrow[k + 3] + hrow[k + 4];
            const int x = 256 * c - acc;            // |x| <= 65280
            nsx += x;
            const unsigned x2 = (unsigned)(x * x);  // < 2^32
            nsxx_lo += x2 & 0xffffu; nsxx_hi += x2 >> 16;
            int h, sv, v;
            dfd_bgr2hsv(stab, t3[k][0], t3[k][1], t3[k][2], &h, &sv, &v);
            ss += sv; sss += sv * sv; vs += v; vss += v * v;
            atomicOr(&shue[h >> 5], 1u << (h & 31));
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ly = ly0 + k;
            const int c = sg[ly + 2][lx + 2];
            const int lap = sg[ly + 1][lx + 2] + sg[ly + 3][lx + 2] + sg[ly + 2][lx + 1] + sg[ly + 2][lx + 3] - 4 * c;
            ls += lap; lss += lap * lap;
            td += dfd_absi(c - pv[k]);
            prev[g0 + (size_t)k * T] = (uint8_t)c;
        }
    }
    int q[11] = {ls, lss, td, nsx, (int)nsxx_lo, (int)nsxx_hi, ss, sss, vs, vss, 0};
#pragma unroll
    for (int i = 0; i < 10; i++) {
        if (i < 3 || is_full) {
            const int r = __reduce_add_sync(0xffffffffu, q[i]);
            if (lx == 0) wpart[warp][i] = r;
        }
    }
    __syncthreads();
    DfdFramePartials* P = part + n;
    if (threadIdx.x < 10 && (threadIdx.x < 3 || is_full)) {
        long long r = 0;
        if (threadIdx.x == 4 || threadIdx.x == 5) {                // unsigned halves of the squared residuals
            for (int w = 0; w < 8; w++) r += (unsigned)wpart[w][threadIdx.x];
        } else {
            for (int w = 0; w < 8; w++) r += wpart[w][threadIdx.x];
        }
        switch (threadIdx.x) {
            case 0: P->lap_s[blk] = r; break;
            case 1: P->lap_ss[blk] = r; break;
            case 2: P->tdiff[blk] = (int)r; break;
            case 3: P->noise_sx[blk] = r; break;
            case 4: wpart[0][10] = 0; P->noise_sxx[blk] = r; break;          // low halves; the high halves are added below
            case 6: P->sat_s[blk] = (unsigned long long)r; break;
            case 7: P->sat_ss[blk] = (unsigned long long)r; break;
            case 8: P->val_s[blk] = (unsigned long long)r; break;
            case 9: P->val_ss[blk] = (unsigned long long)r; break;
            default: break;
        }
    }
    if (is_full) {
        __syncthreads();
        if (threadIdx.x == 5) {
            long long hi = 0;
            for (int w = 0; w < 8; w++) hi += (unsigned)wpart[w][5];
            P->noise_sxx[blk] += hi << 16;
        }
        if (threadIdx.x < 6) P->hue_bits[blk][threadIdx.x] = shue[threadIdx.x];
    }
}

// ---------------------------------------------------------------------------------------------
// Canny: one CTA per frame, two CTAs per SM.  smem: gray 64 KB + candidate bits 8 KB + strong bits 8 KB + one band of packed
// (magnitude | direction) codes.  The frame is processed in bands of CN_BR rows: phase A computes the Sobel magnitude and the
// gradient direction class of every pixel of the band (+ one row above and below) ONCE into the band buffer, phase B does the
// non-maximum suppression from the codes (own code + the two neighbours along the gradient).  The buffer carries a zero column
// on either side and zero rows outside the frame: cv2 treats the magnitude outside the image as 0.
#define CN_BR 32
#define CN_PITCH (T + 2)
#define CN_SMEM (T * T + 2 * 2048 * 4 + (CN_BR + 2) * CN_PITCH * 2)

__global__ void __launch_bounds__(1024, 2) k_canny(const uint8_t* __restrict__ gray, int* __restrict__ count_out,
                                                   size_t count_stride_bytes, uint8_t* __restrict__ edges_out) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* sg = smem;
    uint32_t* W = (uint32_t*)(smem + T * T);
    uint32_t* S = W + 2048;
    uint16_t* mb = (uint16_t*)(S + 2048);
    const int n = blockIdx.x;
    const uint4* src = (const uint4*)(gray + (size_t)n * T * T);
    for (int i = threadIdx.x; i < T * T / 16; i += 1024) ((uint4*)sg)[i] = src[i];
    if (threadIdx.x < 2 * (CN_BR + 2)) mb[(threadIdx.x >> 1) * CN_PITCH + (threadIdx.x & 1) * (T + 1)] = 0;   // guard columns
    __syncthreads();
    for (int y0 = 0; y0 < T; y0 += CN_BR) {
        for (int i = threadIdx.x; i < (CN_BR + 2) * T; i += 1024) {
            const int r = i >> 8, x = i & 255, y = y0 - 1 + r;
            unsigned c = 0;
            if (y >= 0 && y < T) {
                int dx, dy;
                dfd_sobel3(sg, T, T, x, y, &dx, &dy);
                c = dfd_canny_pack(dx, dy);
            }
            mb[r * CN_PITCH + x + 1] = (uint16_t)c;
        }
        __syncthreads();
#pragma unroll 2
        for (int i = threadIdx.x; i < CN_BR * T; i += 1024) {
            const int r = i >> 8, x = i & 255;
            const uint16_t* q = mb + (r + 1) * CN_PITCH + x + 1;
            const unsigned c = *q;
            int st = 0;
            if ((c & 2047u) > DFD_CANNY_LOW) {
                const int off = dfd_canny_first_off(c >> 11, CN_PITCH);
                st = dfd_canny_nms_packed(c, q[off] & 2047, q[-off] & 2047);
            }
            const uint32_t wb = __ballot_sync(0xffffffffu, st >= 1);
            const uint32_t sb = __ballot_sync(0xffffffffu, st == 2);
            const int p = (y0 + r) * T + x;
            if ((threadIdx.x & 31) == 0) { W[p >> 5] = wb; S[p >> 5] = sb; }
        }
        __syncthreads();
    }
    // hysteresis: S <- W & dilate3x3(S) until stable (monotone, in place)
    while (true) {
        int changed = 0;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            int w = threadIdx.x + k * 1024;
            int r = w >> 3, c = w & 7;
            uint32_t cand = W[w], cur = S[w];
            if (cand == cur) continue;
            uint32_t dil = 0;
#pragma unroll
            for (int dr = -1; dr <= 1; dr++) {
                int rr = r + dr;
                if (rr < 0 || rr >= T) continue;
                uint32_t mid = S[rr * 8 + c];
                uint32_t lft = c > 0 ? S[rr * 8 + c - 1] : 0u;
                uint32_t rgt = c < 7 ? S[rr * 8 + c + 1] : 0u;
                dil |= mid | (mid << 1) | (mid >> 1) | (lft >> 31) | (rgt << 31);
            }
            uint32_t nw = cand & dil;
            // run the in-word horizontal propagation to a fixed point
            uint32_t prevv;
            do { prevv = nw; nw |= cand & ((nw << 1) | (nw >> 1)); } while (nw != prevv);
            nw |= cur;
            if (nw != cur) { S[w] = nw; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    int cnt = __popc(S[threadIdx.x]) + __popc(S[threadIdx.x + 1024]);
    cnt = warp_sum(cnt);
    __shared__ int sc[32];
    if ((threadIdx.x & 31) == 0) sc[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int i = 0; i < 32; i++) tot += sc[i];
        if (count_out) *(int*)((char*)count_out + (size_t)n * count_stride_bytes) = tot;
    }
    if (edges_out) {
        for (int p = threadIdx.x; p < T * T; p += 1024)
            edges_out[(size_t)n * T * T + p] = (S[p >> 5] >> (p & 31)) & 1 ? 255 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// ELA: JPEG Q90 4:2:0 round trip in shared memory; one CTA (512 threads) per frame.
__global__ void __launch_bounds__(512, 2) k_ela(const uint8_t* __restrict__ tile, const uint8_t* __restrict__ full,
                                             DfdFramePartials* __restrict__ part, uint8_t* __restrict__ recon_out) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* Y = smem;                  // 256 x 256
    uint8_t* Cb = smem + T * T;         // 128 x 128
    uint8_t* Cr = Cb + 128 * 128;
    __shared__ int sums[DFD_NBLK];
    const int n = blockIdx.x;
    if (full && !full[n]) return;
    const uint8_t* src = tile + (size_t)n * T * T * 3;
    if (threadIdx.x < DFD_NBLK) sums[threadIdx.x] = 0;
    for (int q = threadIdx.x; q < 128 * 128; q += 512) {
        int qy = q >> 7, qx = q & 127;
        int cb = 0, cr = 0;
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int i = 0; i < 2; i++) {
                int y = 2 * qy + j, x = 2 * qx + i;
                const uint8_t* p = src + (y * T + x) * 3;
                int b = p[0], g = p[1], r = p[2];
                Y[y * T + x] = (uint8_t)dfd_jpeg_y(r, g, b);
                cb += dfd_jpeg_cb(r, g, b);
                cr += dfd_jpeg_cr(r, g, b);
            }
        int bias = (qx & 1) ? 2 : 1;
        Cb[q] = (uint8_t)((cb + bias) >> 2);
        Cr[q] = (uint8_t)((cr + bias) >> 2);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 1536; b += 512) {
        uint8_t* plane; int pw, bi, chroma;
        if (b < 1024) { plane = Y; pw = T; bi = b; chroma = 0; }
        else if (b < 1280) { plane = Cb; pw = 128; bi = b - 1024; chroma = 1; }
        else { plane = Cr; pw = 128; bi = b - 1280; chroma = 1; }
        int nbx = pw >> 3;
        uint8_t* base = plane + ((bi / nbx) * 8) * pw + (bi % nbx) * 8;
        int blk[64];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            uint2 v = *(const uint2*)(base + r * pw);
            blk[r * 8 + 0] = v.x & 255; blk[r * 8 + 1] = (v.x >> 8) & 255; blk[r * 8 + 2] = (v.x >> 16) & 255; blk[r * 8 + 3] = v.x >> 24;
            blk[r * 8 + 4] = v.y & 255; blk[r * 8 + 5] = (v.y >> 8) & 255; blk[r * 8 + 6] = (v.y >> 16) & 255; blk[r * 8 + 7] = v.y >> 24;
        }
        dfd_jpeg_block_roundtrip(blk, chroma);
#pragma unroll
        for (int r = 0; r < 8; r++) {
            uint2 v;
            v.x = blk[r * 8 + 0] | (blk[r * 8 + 1] << 8) | (blk[r * 8 + 2] << 16) | (blk[r * 8 + 3] << 24);
            v.y = blk[r * 8 + 4] | (blk[r * 8 + 5] << 8) | (blk[r * 8 + 6] << 16) | (blk[r * 8 + 7] << 24);
            *(uint2*)(base + r * pw) = v;
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < T * T; p += 512) {
        int y = p >> 8, x = p & 255;
        int r, g, b;
        dfd_jpeg_ycc2rgb(Y[p], dfd_jpeg_fancy_up(Cb, 128, 128, x, y), dfd_jpeg_fancy_up(Cr, 128, 128, x, y), &r, &g, &b);
        const uint8_t* o = src + p * 3;
        int d = dfd_bgr2gray(dfd_absi(o[0] - b), dfd_absi(o[1] - g), dfd_absi(o[2] - r));
        if (recon_out) {
            uint8_t* ro = recon_out + ((size_t)n * T * T + p) * 3;
            ro[0] = (uint8_t)b; ro[1] = (uint8_t)g; ro[2] = (uint8_t)r;
        }
        d = warp_sum(d);                                   // a warp covers 32 consecutive x of one row
        if ((threadIdx.x & 31) == 0) atomicAdd(&sums[(y >> 5) * 8 + (x >> 5)], d);
    }
    __syncthreads();
    if (part && threadIdx.x < DFD_NBLK) part[n].ela_sum[threadIdx.x] = sums[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// 256-point FFT as 16 x 16 (Cooley-Tukey, decimation in time) by 16 threads: a thread owns 16 points in registers, runs a
// 16-point DFT (4 x 4, constant twiddles), applies the W256 twiddles, the 16 threads transpose through shared memory
// (one __syncwarp -- they are half a warp) and run the second 16-point DFT.  Output index k = k1 + 16 * k2.
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmuli_neg(float2 a) { return make_float2(a.y, -a.x); }          // a * (-i)

// 4-point forward DFT of (a0..a3) -> (y0..y3)
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = cmuli_neg(csub(a1, a3));
    a0 = cadd(s02, s13); a2 = csub(s02, s13); a1 = cadd(d02, d13); a3 = csub(d02, d13);
}

// 16-point forward DFT in place: v[n] (n = 4*n1 + n2) -> v[k] (k = k1 + 4*k2), natural order in and out
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
    // step A: for each n2, 4-point DFT over n1 of v[4*n1 + n2]  -> a[n2][k1] stored at v[4*k1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    // step B: twiddle a[n2][k1] *= W16^(n2*k1)
    v[4 + 1] = cmulf(v[4 + 1], make_float2(C1, -S1));      // k1=1,n2=1 : W^1
    v[4 + 2] = cmulf(v[4 + 2], make_float2(R, -R));        // W^2
    v[4 + 3] = cmulf(v[4 + 3], make_float2(S1, -C1));      // W^3
    v[8 + 1] = cmulf(v[8 + 1], make_float2(R, -R));        // k1=2,n2=1 : W^2
    v[8 + 2] = cmuli_neg(v[8 + 2]);                        // W^4 = -i
    v[8 + 3] = cmulf(v[8 + 3], make_float2(-R, -R));       // W^6
    v[12 + 1] = cmulf(v[12 + 1], make_float2(S1, -C1));    // k1=3,n2=1 : W^3
    v[12 + 2] = cmulf(v[12 + 2], make_float2(-R, -R));     // W^6
    v[12 + 3] = cmulf(v[12 + 3], make_float2(-C1, S1));    // W^9
    // step C: for each k1, 4-point DFT over n2 of v[4*k1 + n2] -> X[k1 + 4*k2] stored at v[4*k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // reorder: X[k1 + 4*k2] sits at v[4*k1 + k2]  ->  natural order
    float2 t[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) t[k1 + 4 * k2] = v[4 * k1 + k2];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = t[i];
}

// x[16*n1 + t] in v[n1] on entry (t = lane within the 16-thread group); on exit v[k2] = X[t + 16*k2].
// S: this group's shared scratch, 16 x 17 float2.
__device__ __forceinline__ void fft256_16x16(float2 (&v)[16], float2* S, const float2* __restrict__ tw, int t, unsigned gmask) {
    dft16(v);                                              // v[k1] = sum_n1 x[16 n1 + t] W16^(n1 k1)
#pragma unroll
    for (int k1 = 1; k1 < 16; k1++) {                      // * W256^(t * k1)
        const int j = t * k1;                              // < 226
        float2 w = tw[j & 127];
        if (j & 128) { w.x = -w.x; w.y = -w.y; }
        v[k1] = cmulf(v[k1], w);
    }
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) S[k1 * 17 + t] = v[k1];
    __syncwarp(gmask);
#pragma unroll
    for (int n2 = 0; n2 < 16; n2++) v[n2] = S[t * 17 + n2];     // now t plays k1: v[n2] = A[n2][k1 = t]
    __syncwarp(gmask);
    dft16(v);                                              // v[k2] = X[t + 16 k2]
}

// rows: a 16-thread group transforms one row PAIR (real-pair trick); CTA = 8 groups = 16 rows; grid (16, n).
// Output half-spectrum, transposed: fft[n][k][row].
__global__ void __launch_bounds__(128) k_fft_rows(const uint8_t* __restrict__ gray, const float2* __restrict__ tw,
                                                  float2* __restrict__ out) {
    __shared__ float2 S[8][16 * 17];                       // transpose scratch, then the group's spectrum Z[256]
    __shared__ float2 os[129][16];
    const int n = blockIdx.y, grp = blockIdx.x;
    const int f = threadIdx.x >> 4, t = threadIdx.x & 15;
    const unsigned gmask = 0xffffu << (16 * ((threadIdx.x >> 4) & 1));
    const int row0 = grp * 16 + 2 * f;
    const uint8_t* g = gray + (size_t)n * T * T;
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) v[n1] = make_float2((float)g[row0 * T + 16 * n1 + t], (float)g[(row0 + 1) * T + 16 * n1 + t]);
    fft256_16x16(v, S[f], tw, t, gmask);
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) S[f][t + 16 * k2] = v[k2];     // (fft256_16x16 ended with a group barrier after its last read of S)
    __syncwarp(gmask);
    for (int k = t; k <= 128; k += 16) {
        const float2 zk = S[f][k], zn = S[f][(256 - k) & 255];
        // F1 = (Zk + conj(Zn))/2 ; F2 = (Zk - conj(Zn))/(2i)
        os[k][2 * f] = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        os[k][2 * f + 1] = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
    }
    __syncthreads();
    // the 16 rows of this CTA are 128 contiguous bytes of every output column k: write full 128-byte segments
    float2* o = out + (size_t)n * 129 * 256 + grp * 16;
    for (int i = threadIdx.x; i < 129 * 16; i += 128) o[(size_t)(i >> 4) * 256 + (i & 15)] = os[i >> 4][i & 15];
}

// columns: a 16-thread group transforms one half-spectrum column; CTA = 8 columns; grid (17, n).
__global__ void __launch_bounds__(128) k_fft_cols(const float2* __restrict__ in, const float2* __restrict__ tw,
                                                  DfdFramePartials* __restrict__ part) {
    __shared__ float2 S[8][16 * 17];
    __shared__ double red[4][7];
    const int n = blockIdx.y, grp = blockIdx.x;
    const int f = threadIdx.x >> 4, t = threadIdx.x & 15;
    const unsigned gmask = 0xffffu << (16 * ((threadIdx.x >> 4) & 1));
    const int k = grp * 8 + f;
    const bool active = k <= 128;
    float2 v[16];
    const float2* c = in + ((size_t)n * 129 + (active ? k : 0)) * 256;
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) v[n1] = c[16 * n1 + t];
    fft256_16x16(v, S[f], tw, t, gmask);                  // inactive columns transform column 0 again; results unused
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    if (active) {
        const int fv = k < 128 ? k : -128;
        const double wgt = (k == 0 || k == 128) ? 1.0 : 2.0;
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const int u = t + 16 * k2;
            const int fu = u < 128 ? u : u - 256;
            const int d2 = fu * fu + fv * fv;
            const float2 z = v[k2];
            const double m = (double)log1pf(hypotf(z.x, z.y));
            if (d2 <= 32 * 32) { acc[0] += wgt * m; acc[4] += wgt; }
            else if (d2 <= 64 * 64) { acc[1] += wgt * m; acc[2] += wgt * m * m; acc[5] += wgt; }
            else if (d2 <= 128 * 128) { acc[3] += wgt * m; acc[6] += wgt; }
        }
    }
#pragma unroll
    for (int i = 0; i < 7; i++) acc[i] = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < 7; i++) red[threadIdx.x >> 5][i] = acc[i];
    __syncthreads();
    if (threadIdx.x < 7) {
        double r = 0;
        for (int w = 0; w < 4; w++) r += red[w][threadIdx.x];
        part[n].fft[grp][threadIdx.x] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// Score tables + state update; one WARP per frame (8 frames per CTA).  The 64-block / 30-entry statistics are computed
// lane-parallel, the float32 means / stds in NumPy's exact pairwise order (px_numpy.h): lanes 0-7 carry the eight partial
// sums r[j] of the unrolled loop, lane 0 combines them ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and adds the tail.  Lane 0 owns
// the thresholds, the per-stream ring and the weighted sum.  (One thread per frame walked every 64-element loop
// serially with fp64 sqrt / div in it: 105 us per step on 4 CTAs.)
__device__ double clip01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }

#define FIN_WARPS 8

// np.sum of a float32 array a[0..n) (n <= 64) held in shared memory; result valid in every lane
__device__ float warp_np_sum_f32(const float* a, int n, int lane) {
    float res;
    if (n < 8) {
        res = 0.f;
        if (lane == 0) for (int i = 0; i < n; i++) res = __fadd_rn(res, a[i]);
    } else {
        const int n8 = n - (n % 8);
        float r = 0.f;
        if (lane < 8) {
            r = a[lane];
            for (int i = 8 + lane; i < n8; i += 8) r = __fadd_rn(r, a[i]);
        }
        const float r1 = __shfl_down_sync(0xffffffffu, r, 1);
        const float s01 = __fadd_rn(r, r1);                                    // lanes 0,2,4,6: r[j] + r[j+1]
        const float s23 = __shfl_down_sync(0xffffffffu, s01, 2);
        const float q = __fadd_rn(s01, s23);                                   // lanes 0,4: (r0+r1)+(r2+r3), (r4+r5)+(r6+r7)
        const float q4 = __shfl_down_sync(0xffffffffu, q, 4);
        res = __fadd_rn(q, q4);
        if (lane == 0) for (int i = n8; i < n; i++) res = __fadd_rn(res, a[i]);
    }
    return __shfl_sync(0xffffffffu, res, 0);
}
__device__ float warp_np_mean_f32(const float* a, int n, int lane) { return __fdiv_rn(warp_np_sum_f32(a, n, lane), (float)n); }
// np.std (population); tmp: n floats of scratch in shared memory
__device__ float warp_np_std_f32(const float* a, int n, float* tmp, int lane) {
    const float mean = warp_np_mean_f32(a, n, lane);
    for (int i = lane; i < n; i += 32) { const float d = __fsub_rn(a[i], mean); tmp[i] = __fmul_rn(d, d); }
    __syncwarp();
    const float s = warp_np_sum_f32(tmp, n, lane);
    __syncwarp();
    return __fsqrt_rn(__fdiv_rn(s, (float)n));
}
template <typename V> __device__ __forceinline__ V warp_isum(V v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(32 * FIN_WARPS)
k_finalize(int n, int max_streams, const int32_t* __restrict__ stream_ids, const uint8_t* __restrict__ full,
           const DfdFramePartials* __restrict__ part, DfdStreamState* __restrict__ state,
           dfd_forensic_result* __restrict__ results) {
    __shared__ float s_a[FIN_WARPS][DFD_NBLK], s_b[FIN_WARPS][DFD_NBLK];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * FIN_WARPS + w;
    if (i >= n) return;
    const DfdFramePartials& P = part[i];
    const int sid = stream_ids[i];
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    float* tmp = s_a[w];
    float* tmp2 = s_b[w];
    dfd_forensic_result R;                                      // assembled by lane 0
    for (int k = 0; k < DFD_N_RAW; k++) R.raw[k] = NaN;
    for (int k = 0; k < DFD_N_SIGNALS; k++) R.scores[k] = NaN;
    if ((unsigned)sid >= (unsigned)max_streams) {               // caller error: no state is touched, the record says so
        if (lane == 0) { R.fake_probability = NaN; R.frame_number = -1; R.full = 0; results[i] = R; }
        return;
    }
    DfdStreamState& S = state[sid];
    const bool is_full = full[i] != 0;
    int frames_now = S.analyzer_frames + 1;                     // frame_analysis.py:68,110
    R.frame_number = frames_now;
    R.full = is_full ? 1 : 0;

    // ---- frequency (frame_analysis.py:150-180): lane k sums column k of the 17 group partials in group order ----
    {
        double a = 0.0;
        if (lane < 7) for (int g = 0; g < DFD_FFT_GROUPS; g++) a += P.fft[g][lane];
        const double a0 = __shfl_sync(0xffffffffu, a, 0), a1 = __shfl_sync(0xffffffffu, a, 1), a2 = __shfl_sync(0xffffffffu, a, 2);
        const double a3 = __shfl_sync(0xffffffffu, a, 3), a4 = __shfl_sync(0xffffffffu, a, 4), a5 = __shfl_sync(0xffffffffu, a, 5);
        const double a6 = __shfl_sync(0xffffffffu, a, 6);
        if (lane == 0) {
            float low = (float)(a0 / a4), mid = (float)(a1 / a5), high = (float)(a3 / a6);
            float total = __fadd_rn(__fadd_rn(__fadd_rn(low, mid), high), 1e-10f);
            float hr = __fdiv_rn(high, total), mr = __fdiv_rn(mid, total);
            double mmean = a1 / a5;
            double var = a2 / a5 - mmean * mmean;
            float mstd = (float)sqrt(var > 0 ? var : 0.0);
            float mcv = __fdiv_rn(mstd, __fadd_rn(mid, 1e-10f));
            R.raw[0] = hr; R.raw[1] = mr; R.raw[2] = mcv;
            double s = 0.0;
            if (hr < 0.18f) s += 0.4; else if (hr < 0.22f) s += 0.2;
            if (mcv > 0.6f) s += 0.25; else if (mcv > 0.45f) s += 0.1;
            if (mr > 0.45f && hr < 0.2f) s += 0.15;
            R.scores[0] = clip01(s);
        }
    }
    if (is_full) {
        // ---- noise (:194-225) ----
        for (int b = lane; b < DFD_NBLK; b += 32) {
            double sx = (double)P.noise_sx[b], sxx = (double)P.noise_sxx[b];
            double var = (sxx - sx * sx / 1024.0) / 1024.0;
            tmp[b] = (float)(sqrt(var > 0 ? var : 0.0) / 256.0);
        }
        __syncwarp();
        float mean_noise = warp_np_mean_f32(tmp, DFD_NBLK, lane);
        float ncv = __fdiv_rn(warp_np_std_f32(tmp, DFD_NBLK, tmp2, lane), __fadd_rn(mean_noise, 1e-10f));
        if (lane == 0) {
            R.raw[3] = ncv; R.raw[4] = mean_noise;
            double s = 0.0;
            if (ncv > 0.7f) s += 0.5; else if (ncv > 0.5f) s += 0.25;
            if (mean_noise < 1.0f) s += 0.3; else if (mean_noise < 2.0f) s += 0.1;
            R.scores[1] = clip01(s);
        }
        __syncwarp();
        // ---- ELA (:245-276) ----
        for (int b = lane; b < DFD_NBLK; b += 32) tmp[b] = __fdiv_rn((float)P.ela_sum[b], 1024.0f);
        __syncwarp();
        float emean = warp_np_mean_f32(tmp, DFD_NBLK, lane);
        float ecv = __fdiv_rn(warp_np_std_f32(tmp, DFD_NBLK, tmp2, lane), __fadd_rn(emean, 1e-10f));
        if (lane == 0) {
            R.raw[5] = ecv; R.raw[6] = emean;
            double s = 0.0;
            if (ecv > 0.9f) s += 0.5; else if (ecv > 0.6f) s += 0.2;
            if (emean > 15.0f) s += 0.2; else if (emean > 10.0f) s += 0.1;
            R.scores[2] = clip01(s);
        }
        __syncwarp();
    }
    // ---- edges (:288-309): exact integer sums, any order ----
    {
        long long ls = 0, lss = 0;
        for (int b = lane; b < DFD_NBLK; b += 32) { ls += P.lap_s[b]; lss += P.lap_ss[b]; }
        ls = warp_isum(ls); lss = warp_isum(lss);
        if (lane == 0) {
            double density = (double)P.canny_count / 65536.0;
            double mean = (double)ls / 65536.0;
            double var = (double)lss / 65536.0 - mean * mean;
            R.raw[7] = density; R.raw[8] = var;
            double s = 0.0;
            if (density < 0.02) s += 0.35; else if (density < 0.04) s += 0.15;
            if (var < 50.0) s += 0.3; else if (var < 100.0) s += 0.1;
            R.scores[3] = clip01(s);
        }
    }
    if (is_full) {
        // ---- colour (:318-347) ----
        unsigned long long ss = 0, sss = 0, vs = 0, vss = 0;
        unsigned int hb[6] = {0, 0, 0, 0, 0, 0};
        for (int b = lane; b < DFD_NBLK; b += 32) {
            ss += P.sat_s[b]; sss += P.sat_ss[b]; vs += P.val_s[b]; vss += P.val_ss[b];
#pragma unroll
            for (int k = 0; k < 6; k++) hb[k] |= P.hue_bits[b][k];
        }
        ss = warp_isum(ss); sss = warp_isum(sss); vs = warp_isum(vs); vss = warp_isum(vss);
#pragma unroll
        for (int k = 0; k < 6; k++) hb[k] = __reduce_or_sync(0xffffffffu, hb[k]);
        if (lane == 0) {
            double sm = (double)ss / 65536.0, vm = (double)vs / 65536.0;
            double svar = (double)sss / 65536.0 - sm * sm, vvar = (double)vss / 65536.0 - vm * vm;
            float sstd = (float)sqrt(svar > 0 ? svar : 0.0), vstd = (float)sqrt(vvar > 0 ? vvar : 0.0);
            int hues = 0;
            for (int k = 0; k < 6; k++) hues += __popc(hb[k]);
            R.raw[9] = sstd; R.raw[10] = vstd; R.raw[11] = hues;
            double s = 0.0;
            if (sstd < 15.0f) s += 0.3; else if (sstd < 25.0f) s += 0.1;
            if (vstd < 15.0f) s += 0.25; else if (vstd < 25.0f) s += 0.1;
            if (hues < 30) s += 0.25; else if (hues < 50) s += 0.1;
            R.scores[4] = clip01(s);
        }
    }
    // ---- temporal (:356-389) ----
    {
        double s = 0.0;
        const int has_prev = S.has_prev;
        int td = 0;
        for (int b = lane; b < DFD_NBLK; b += 32) td += P.tdiff[b];
        td = warp_isum(td);
        int ring_n = S.ring_n, ring_head = S.ring_head;
        const float mean_diff = __fdiv_rn((float)td, 65536.0f);       // exact: integer sum < 2^24
        if (has_prev) {
            if (lane == 0) {
                if (ring_n < DFD_RING) S.ring[(ring_head + ring_n) % DFD_RING] = mean_diff;
                else S.ring[ring_head] = mean_diff;
            }
            if (ring_n < DFD_RING) ring_n++; else ring_head = (ring_head + 1) % DFD_RING;
            __syncwarp();
        }
        float cv = 0.f;
        if (has_prev && ring_n >= 5) {
            if (lane < ring_n) tmp[lane] = S.ring[(ring_head + lane) % DFD_RING];
            __syncwarp();
            const float md = warp_np_mean_f32(tmp, ring_n, lane);
            cv = __fdiv_rn(warp_np_std_f32(tmp, ring_n, tmp2, lane), __fadd_rn(md, 1e-10f));
        }
        if (lane == 0) {
            if (!has_prev) {
                S.has_prev = 1;
                R.raw[14] = 0;
            } else {
                S.ring_n = ring_n; S.ring_head = ring_head;
                R.raw[13] = mean_diff; R.raw[14] = ring_n;
                if (ring_n >= 5) {
                    R.raw[12] = cv;
                    if (cv > 1.5f) s += 0.4; else if (cv > 1.0f) s += 0.2;
                    if (mean_diff < 0.3f && frames_now > 10) s += 0.3;
                    else if (mean_diff < 0.8f && frames_now > 10) s += 0.1;
                }
            }
            S.analyzer_frames = frames_now;
            R.scores[5] = clip01(s);
        }
    }
    if (lane != 0) return;
    // ---- weighted sum in the reference's dict order with Python's compensated sum() (:94,119) ----
    double c;
    if (is_full) {
        const double sc[6] = {R.scores[0], R.scores[1], R.scores[2], R.scores[3], R.scores[4], R.scores[5]};
        const double wt[6] = {0.25, 0.20, 0.20, 0.15, 0.10, 0.10};
        c = dfd_py_sum_products(sc, wt, 6);
    } else {
        const double sc[3] = {R.scores[0], R.scores[5], R.scores[3]};
        const double wt[3] = {0.45, 0.25, 0.30};
        c = dfd_py_sum_products(sc, wt, 3);
    }
    R.fake_probability = clip01(c);
    results[i] = R;
}

// ---------------------------------------------------------------------------------------------
int dfd_forensics_launch(dfd_ctx* ctx, const uint8_t* frames, int n, int H, int W, size_t frame_stride, int row_pitch,
                         const int32_t* stream_ids, const uint8_t* full, dfd_forensic_result* results, cudaStream_t st) {
    DFD_REQUIRE(n > 0 && n <= ctx->cfg.max_batch, DFD_ERR_CAPACITY, "forensics: batch exceeds max_batch");
    DFD_REQUIRE(H >= 1 && W >= 1 && row_pitch >= 3 * W, DFD_ERR_INVALID, "forensics: bad frame geometry");
    const RsTap* rs_tab = nullptr;
    { int rc = dfd_resize_tables(ctx, H, W, &rs_tab); if (rc) return rc; }
    // aligned 32-bit tap loads need 4-byte aligned rows
    const int wide_ok = ((uintptr_t)frames % 4 == 0) && (frame_stride % 4 == 0) && (row_pitch % 4 == 0);
    k_resize256<<<dim3(T, n), 256, 0, st>>>(frames, H, W, frame_stride, row_pitch, rs_tab, wide_ok, ctx->d_tile, ctx->d_gray);
    DFD_LAUNCH_CHECK("k_resize256", st);
    k_tile_stats<<<dim3(DFD_NBLK, n), 256, 0, st>>>(ctx->d_tile, ctx->d_gray, stream_ids, full, ctx->d_tables, ctx->d_state,
                                                     ctx->d_prev_gray, ctx->d_part, ctx->cfg.max_streams);
    DFD_LAUNCH_CHECK("k_tile_stats", st);
    k_canny<<<n, 1024, CN_SMEM, st>>>(ctx->d_gray, &ctx->d_part[0].canny_count, sizeof(DfdFramePartials), nullptr);
    DFD_LAUNCH_CHECK("k_canny", st);
    k_ela<<<n, 512, T * T + 2 * 128 * 128, st>>>(ctx->d_tile, full, ctx->d_part, nullptr);
    DFD_LAUNCH_CHECK("k_ela", st);
    k_fft_rows<<<dim3(16, n), 128, 0, st>>>(ctx->d_gray, ctx->d_twiddle, ctx->d_fft);
    DFD_LAUNCH_CHECK("k_fft_rows", st);
    k_fft_cols<<<dim3(DFD_FFT_GROUPS, n), 128, 0, st>>>(ctx->d_fft, ctx->d_twiddle, ctx->d_part);
    DFD_LAUNCH_CHECK("k_fft_cols", st);
    k_finalize<<<(n + FIN_WARPS - 1) / FIN_WARPS, 32 * FIN_WARPS, 0, st>>>(n, ctx->cfg.max_streams, stream_ids, full, ctx->d_part, ctx->d_state, results);
    DFD_LAUNCH_CHECK("k_finalize", st);
    return DFD_OK;
}

int dfd_forensics_init(dfd_ctx* ctx) {
    int rc;
    if ((rc = dfd_func_smem(ctx, k_canny, CN_SMEM))) return rc;
    if ((rc = dfd_func_smem(ctx, k_ela, T * T + 2 * 128 * 128))) return rc;
    return DFD_OK;
}

int dfd_dbg_jpeg_launch(dfd_ctx* ctx, const uint8_t* tiles, uint8_t* out, int n, cudaStream_t st) {
    k_ela<<<n, 512, T * T + 2 * 128 * 128, st>>>(tiles, nullptr, nullptr, out);
    DFD_LAUNCH_CHECK("k_ela", st);
    return DFD_OK;
}

int dfd_dbg_canny_launch(dfd_ctx* ctx, const uint8_t* gray, uint8_t* edges, int n, cudaStream_t st) {
    k_canny<<<n, 1024, CN_SMEM, st>>>(gray, nullptr, 0, edges);
    DFD_LAUNCH_CHECK("k_canny", st);
    return DFD_OK;
}
