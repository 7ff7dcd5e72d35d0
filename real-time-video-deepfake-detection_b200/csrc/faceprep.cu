// Face-crop preparation as batched sm_100a kernels.
//
// Replaces, for m face boxes at once, the reference's per-face host pipeline
// (deepfake_detection.py:357-389) in the reference's ORDER:
//   crop -> BGR2LAB -> CLAHE(2.0, 8x8) on L at native crop resolution -> LAB2BGR -> RGB
//        -> Pillow BILINEAR (antialiased, two u8 passes) to 160x160     [MTCNN extract_face step]
//        -> F.interpolate bilinear 224x224 -> /255 -> ImageNet normalise -> NHWC f32 | bf16
// All u8 stages are bit-exact (px_*.h, checked on the CPU against cv2 / PIL by tests/hostcheck).
//
//   k_pil_coeffs        Pillow resample coefficients for both axes of every box (double math)
//   k_clahe_lut         per (box, row of 8 CLAHE tiles), warp per tile: L histogram in smem -> clipped, redistributed LUT
//   k_clahe_hpass       per (box, 16 crop rows), warp per row: LAB/CLAHE/LAB2BGR row into smem, Pillow horizontal pass
//   k_vpass_up_norm     per (box, 56-row band): Pillow vertical pass into smem, bilinear 224, normalise
//   k_tta_hpass         test-time augmentation (deepfake_detection.py:408-443): flip / convertScaleAbs / warpAffine of the CLAHE'd
//                       crop computed per pixel of an augmented row (px_warp.h) and fed straight into the Pillow horizontal pass
#include "dfd_internal.cuh"
#include "px_resize.h"
#include "px_clahe.h"
#include "px_numpy.h"
#include "px_warp.h"

#define PIL_STRIDE (2 + DFD_PIL_KMAX)

__global__ void k_pil_coeffs(const int32_t* __restrict__ boxes, int* __restrict__ pil) {
    const int m = blockIdx.x, t = threadIdx.x;       // 320 threads: axis = t / 160
    const int axis = t / 160, xx = t % 160;
    const int in_size = boxes[m * 4 + 2 + axis];
    int* o = pil + (((size_t)m * 2 + axis) * 160 + xx) * PIL_STRIDE;
    int xmin, k[DFD_PIL_KMAX];
    int cnt = dfd_pil_coeffs(xx, in_size, 160, &xmin, k);
    o[0] = xmin; o[1] = cnt;
    for (int i = 0; i < cnt; i++) o[2 + i] = k[i];
}

// CTA = one row of 8 CLAHE tiles of one box; warp = one tile.  The L channel alone is needed here (3 gamma + 1 cube-root
// table look-ups per pixel, tables staged in shared memory); bins are counted with warp-aggregated shared atomics and
// the clipped / redistributed / cumulative LUT is built by the warp itself (dfd_clahe_lut_warp).
__global__ void __launch_bounds__(256) k_clahe_lut(const uint8_t* __restrict__ frames, size_t fstride, int pitch,
                                                   const int32_t* __restrict__ boxes, const int32_t* __restrict__ frame_idx,
                                                   const DfdColorTables* __restrict__ tab, uint8_t* __restrict__ luts) {
    __shared__ int hist[8][256];
    __shared__ __align__(16) uint16_t s_gamma[256];
    __shared__ __align__(16) uint16_t s_cbrt[3072];
    // blockDim.x / 32 tiles per CTA (8 at large batch, 2 at small batch so one box still fills the chip)
    const int m = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nthr = blockDim.x;
    const int tile = blockIdx.x * (nthr >> 5) + warp, ty = tile >> 3, tx = tile & 7;
    const int bx = boxes[m * 4], by = boxes[m * 4 + 1], bw = boxes[m * 4 + 2], bh = boxes[m * 4 + 3];
    const DfdClaheGeom g = dfd_clahe_geom(bw, bh);
    for (int i = threadIdx.x; i < 8 * 256; i += nthr) (&hist[0][0])[i] = 0;
    for (int i = threadIdx.x; i < 256 / 8; i += nthr) ((uint4*)s_gamma)[i] = ((const uint4*)tab->gamma)[i];
    for (int i = threadIdx.x; i < 3072 / 8; i += nthr) ((uint4*)s_cbrt)[i] = ((const uint4*)tab->cbrt)[i];
    __syncthreads();
    const uint8_t* f = frames + (size_t)frame_idx[m] * fstride;
    const int area = g.tw * g.th;
    const unsigned magic = g.tw > 1 ? 0xffffffffu / (unsigned)g.tw + 1u : 0u;   // p / tw == umulhi(p, magic) for p * tw < 2^32
    int* h = hist[warp];
    // four pixels per lane per trip: the 12 byte loads of a trip are issued before the first conversion, so a warp has four
    // DRAM round trips in flight instead of one (the crop is read here for the first time: the loop was latency-bound)
    for (int p0 = 0; p0 < area; p0 += 128) {
        int b[4], gq[4], r[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int p = p0 + u * 32 + lane;
            ok[u] = p < area;
            b[u] = gq[u] = r[u] = 0;
            if (ok[u]) {
                const int yy = g.tw > 1 ? (int)__umulhi((unsigned)p, magic) : p, xx = p - yy * g.tw;
                int y = ty * g.th + yy, x = tx * g.tw + xx;
                if (x >= bw) x = dfd_reflect101(x, bw);
                if (y >= bh) y = dfd_reflect101(y, bh);
                const uint8_t* px = f + (size_t)(by + y) * pitch + (size_t)(bx + x) * 3;
                b[u] = __ldg(px); gq[u] = __ldg(px + 1); r[u] = __ldg(px + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (p0 + u * 32 >= area) break;                     // warp-uniform
            const int L = ok[u] ? dfd_bgr2lab_L(s_gamma, s_cbrt, b[u], gq[u], r[u]) : 0;
            // warp-aggregated histogram update: one shared atomic per distinct value
            const unsigned act = __ballot_sync(0xffffffffu, ok[u]);
            if (ok[u]) {
                const unsigned peers = __match_any_sync(act, L);
                if ((int)(__ffs(peers) - 1) == lane) atomicAdd(&h[L], __popc(peers));
            }
        }
    }
    __syncwarp();
    dfd_clahe_lut_warp(h, g.clip, g.lut_scale, luts + ((size_t)m * 64 + ty * 8 + tx) * 256, lane);
}

// Pillow's horizontal resampling pass of one RGB row (in shared memory) to 160 pixels: a lane resamples one output pixel
// (3 channels share the tap positions and weights); k = this box's coefficient table, kstride ints per output pixel.
__device__ __forceinline__ void pil_hpass_row(const uint8_t* row, uint8_t* out, const int* ktab, int kstride, int lane) {
    for (int xx = lane; xx < 160; xx += 32) {
        const int* k = ktab + xx * kstride;
        const int xmin = k[0], cnt = k[1];
        int a0 = 1 << (DFD_PIL_PRECISION - 1), a1 = a0, a2 = a0;
        const uint8_t* rp = row + xmin * 3;
        for (int t = 0; t < cnt; t++) {
            const int w = k[2 + t];
            a0 += rp[0] * w; a1 += rp[1] * w; a2 += rp[2] * w;
            rp += 3;
        }
        out[xx * 3] = (uint8_t)dfd_pil_clip8(a0);
        out[xx * 3 + 1] = (uint8_t)dfd_pil_clip8(a1);
        out[xx * 3 + 2] = (uint8_t)dfd_pil_clip8(a2);
    }
}

// CTA = HP_ROWS consecutive crop rows of one box (32: staging the 34 KB of tables costs more than 16 rows of pixels), warp = one row at a time: BGR -> LAB -> CLAHE(L) -> LAB2BGR -> RGB row
// in shared memory, then Pillow's horizontal resampling pass of that row to 160 pixels.  The colour tables and the
// box's 64 CLAHE LUTs are staged in shared memory once per CTA (11 table gathers + 4 LUT gathers per pixel).
#define HP_ROWS 32
#define DFD_PIL_KMAX_STAGED 15      // Pillow tap counts up to 15 (crops up to 1120 px) are staged in shared memory
__global__ void __launch_bounds__(256) k_clahe_hpass(const uint8_t* __restrict__ frames, size_t fstride, int pitch,
                                                     const int32_t* __restrict__ boxes, const int32_t* __restrict__ frame_idx,
                                                     const DfdColorTables* __restrict__ tab, const uint8_t* __restrict__ luts,
                                                     const int* __restrict__ pil, uint8_t* __restrict__ hpass, int max_crop,
                                                     uint8_t* __restrict__ clahe_out, size_t clahe_stride, int clahe_box,
                                                     int slot_mul, int rows_per_cta) {
    extern __shared__ __align__(16) uint8_t hp_smem[];
    DfdColorTables* s_tab = (DfdColorTables*)hp_smem;                               // sizeof is a multiple of 16
    uint8_t* s_lut = hp_smem + sizeof(DfdColorTables);                              // [64][256]
    int* s_pc = (int*)(s_lut + 64 * 256);                                           // [160][2 + ksize] Pillow horizontal taps of this box
    uint8_t* s_rows = (uint8_t*)(s_pc + 160 * (2 + DFD_PIL_KMAX_STAGED));           // [8 warps][row_stride]
    const int m = blockIdx.y, y0 = blockIdx.x * rows_per_cta;
    const int bx = boxes[m * 4], by = boxes[m * 4 + 1], bw = boxes[m * 4 + 2], bh = boxes[m * 4 + 3];
    if (y0 >= bh) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (int)(sizeof(DfdColorTables) / 16); i += 256) ((uint4*)s_tab)[i] = ((const uint4*)tab)[i];
    for (int i = threadIdx.x; i < 64 * 256 / 16; i += 256) ((uint4*)s_lut)[i] = ((const uint4*)(luts + (size_t)m * 64 * 256))[i];
    // Pillow's horizontal coefficients (the same for every row of the box) are read ~7 times per output value: keep them in
    // shared memory, compacted to the box's real tap count
    const int* pg = pil + ((size_t)m * 2 + 0) * 160 * PIL_STRIDE;
    const int ks_real = dfd_pil_ksize(bw, 160);
    const bool staged = ks_real <= DFD_PIL_KMAX_STAGED;    // very large crops (> 1120 px wide) read the table from global memory
    const int pcs = 2 + (staged ? ks_real : 0);
    if (staged) {
        for (int i = threadIdx.x; i < 160 * pcs; i += 256) {
            const int xx = i / pcs, j = i - xx * pcs;
            s_pc[i] = pg[xx * PIL_STRIDE + j];
        }
    }
    __syncthreads();
    const DfdClaheGeom g = dfd_clahe_geom(bw, bh);
    const bool words_ok = ((uintptr_t)frames % 4 == 0) && (fstride % 4 == 0) && (pitch % 4 == 0);
    const int row_stride = (max_crop * 3 + 8 + 15) & ~15;
    uint8_t* row = s_rows + (size_t)warp * row_stride;
    for (int y = y0 + warp; y < y0 + rows_per_cta && y < bh; y += 8) {
        const uint8_t* f = frames + (size_t)frame_idx[m] * fstride + (size_t)(by + y) * pitch + (size_t)bx * 3;
        // the crop row is fetched with aligned, coalesced 32-bit loads into the row buffer and converted in place (the three
        // byte loads per pixel straight from the frame were the kernel's main stall: 55 % long-scoreboard)
        int off = 0;
        const uint8_t* src = f;
        if (words_ok) {
            off = (int)((uintptr_t)f & 3);
            const uint32_t* fw = (const uint32_t*)(f - off);
            const int nwords = (off + bw * 3 + 3) >> 2;
            for (int i = lane; i < nwords; i += 32) ((uint32_t*)row)[i] = __ldg(fw + i);
            __syncwarp();
            src = row + off;
        }
        for (int x0 = 0; x0 < bw; x0 += 32) {
            const int x = x0 + lane;
            int L = 0, A = 0, B = 0, ob = 0, og = 0, orr = 0;
            if (x < bw) {
                dfd_bgr2lab(s_tab, src[x * 3], src[x * 3 + 1], src[x * 3 + 2], &L, &A, &B);
                L = dfd_clahe_apply(s_lut, g, x, y, L);
                dfd_lab2bgr(s_tab, L, A, B, &ob, &og, &orr);
            }
            __syncwarp();                                    // every lane has read its pixel before anyone overwrites the buffer
            if (x < bw) {
                row[x * 3] = (uint8_t)orr; row[x * 3 + 1] = (uint8_t)og; row[x * 3 + 2] = (uint8_t)ob;   // RGB order (:376)
                // the CLAHE'd crop itself (tight w*h*3 BGR): box `clahe_box` alone (diagnostics) or, with clahe_box < 0, every box
                // at its own clahe_stride slot (the base image of the test-time augmentations)
                if (clahe_out && (clahe_box < 0 || m == clahe_box)) {
                    uint8_t* d = clahe_out + (clahe_box < 0 ? (size_t)m * clahe_stride : 0) + ((size_t)y * bw + x) * 3;
                    d[0] = (uint8_t)ob; d[1] = (uint8_t)og; d[2] = (uint8_t)orr;
                }
            }
        }
        __syncwarp();
        if (hpass) pil_hpass_row(row, hpass + (((size_t)m * slot_mul * max_crop + y) * 160) * 3, staged ? s_pc : pg, staged ? pcs : PIL_STRIDE, lane);
        __syncwarp();
    }
}

// Test-time augmentation (deepfake_detection.py:417-434): CTA = rows_per_cta rows of ONE augmented copy of one box, warp = one
// row at a time.  The augmented image is never materialised: a lane computes its pixel of the row -- cv2.warpAffine's
// fixed-point source position, four taps fetched from the CLAHE'd base crop (written by k_clahe_hpass, L2-resident) through
// the flip and the brightness table (px_warp.h) -- into the shared-memory row buffer in RGB order, and the row goes straight
// through Pillow's horizontal pass into h-pass slot i * n_pred + 1 + a; k_vpass_up_norm finishes it like any other crop.
struct DfdTtaAug { int32_t flip; float brightness; double im[6]; };
static_assert(sizeof(DfdTtaAug) == sizeof(dfd_tta_aug), "dfd_tta_aug layout");
__global__ void __launch_bounds__(256) k_tta_hpass(const uint8_t* __restrict__ base, size_t base_stride,
                                                   const int32_t* __restrict__ boxes, const DfdTtaAug* __restrict__ augs,
                                                   const int* __restrict__ pil, uint8_t* __restrict__ hpass, int max_crop,
                                                   int n_pred, int rows_per_cta) {
    extern __shared__ __align__(16) uint8_t hp_smem[];
    uint8_t* s_lut = hp_smem;                                                       // [256] brightness table
    int* s_pc = (int*)(hp_smem + 256);                                              // [160][2 + ksize]
    uint8_t* s_rows = (uint8_t*)(s_pc + 160 * (2 + DFD_PIL_KMAX_STAGED));           // [8 warps][row_stride]
    const int n_aug = n_pred - 1;
    const int i = blockIdx.y / n_aug, a = blockIdx.y - i * n_aug, y0 = blockIdx.x * rows_per_cta;
    const int bw = boxes[i * 4 + 2], bh = boxes[i * 4 + 3];
    if (y0 >= bh) return;
    const DfdTtaAug A = augs[(size_t)i * n_aug + a];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    s_lut[threadIdx.x] = (uint8_t)dfd_scale_abs_u8(threadIdx.x, A.brightness);
    const int* pg = pil + ((size_t)i * 2 + 0) * 160 * PIL_STRIDE;
    const int ks_real = dfd_pil_ksize(bw, 160);
    const bool staged = ks_real <= DFD_PIL_KMAX_STAGED;
    const int pcs = 2 + (staged ? ks_real : 0);
    if (staged) {
        for (int t = threadIdx.x; t < 160 * pcs; t += 256) {
            const int xx = t / pcs, j = t - xx * pcs;
            s_pc[t] = pg[xx * PIL_STRIDE + j];
        }
    }
    __syncthreads();
    const int row_stride = (max_crop * 3 + 8 + 15) & ~15;
    uint8_t* row = s_rows + (size_t)warp * row_stride;
    const uint8_t* src = base + (size_t)i * base_stride;
    const size_t slot = (size_t)i * n_pred + 1 + a;
    for (int y = y0 + warp; y < y0 + rows_per_cta && y < bh; y += 8) {
        for (int x = lane; x < bw; x += 32) {
            int ob, og, orr;
            dfd_tta_pixel(src, bw * 3, bw, bh, A.flip, s_lut, A.im, x, y, &ob, &og, &orr);
            row[x * 3] = (uint8_t)orr; row[x * 3 + 1] = (uint8_t)og; row[x * 3 + 2] = (uint8_t)ob;     // RGB order (:376)
        }
        __syncwarp();
        pil_hpass_row(row, hpass + ((slot * max_crop + y) * 160) * 3, staged ? s_pc : pg, staged ? pcs : PIL_STRIDE, lane);
        __syncwarp();
    }
}

template <typename OutT>
__device__ __forceinline__ void store8(OutT* o, const float* v);
template <>
__device__ __forceinline__ void store8<float>(float* o, const float* v) {
    ((float4*)o)[0] = make_float4(v[0], v[1], v[2], v[3]);
    ((float4*)o)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* o, const float* v) {
    uint4 t;
    __nv_bfloat162* h = (__nv_bfloat162*)&t;
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *(uint4*)o = t;
}

template <typename OutT>
__global__ void __launch_bounds__(512) k_vpass_up_norm(const int32_t* __restrict__ boxes, const int* __restrict__ pil,
                                                       const uint8_t* __restrict__ hpass, int max_crop,
                                                       uint8_t* __restrict__ face160, OutT* __restrict__ out, int rep) {
    __shared__ __align__(16) uint8_t s160[44][480];
    const int m = blockIdx.y, band = blockIdx.x;           // 4 bands of 56 output rows; crop m belongs to box m / rep (TTA copies)
    int r_first, r_last, t0, t1; float l0, l1;
    dfd_torch_bilinear_coef(band * 56, 160, 224, &r_first, &t1, &l0, &l1);
    dfd_torch_bilinear_coef(band * 56 + 55, 160, 224, &t0, &r_last, &l0, &l1);
    const int nrows = r_last - r_first + 1;                // <= 42
    const int* pc = pil + ((size_t)(m / rep) * 2 + 1) * 160 * PIL_STRIDE;
    const uint8_t* hp = hpass + (size_t)m * max_crop * 480;
    // a thread resamples 4 consecutive bytes of an output row: one 32-bit load of the h-pass image per tap
    for (int o = threadIdx.x; o < nrows * 120; o += 512) {
        const int r = o / 120, xw = o - r * 120;
        const int* k = pc + (r_first + r) * PIL_STRIDE;
        const int ymin = k[0], cnt = k[1];
        int a0 = 1 << (DFD_PIL_PRECISION - 1), a1 = a0, a2 = a0, a3 = a0;
        const uint32_t* src = (const uint32_t*)(hp + (size_t)ymin * 480) + xw;
        // four taps per trip, loads first: the h-pass image comes from L2 and a one-tap loop pays that latency per tap
        for (int t = 0; t < cnt; t += 4) {
            int w[4]; uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const bool ok = t + u < cnt;
                w[u] = ok ? k[2 + t + u] : 0;
                v[u] = ok ? src[(size_t)(t + u) * 120] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                a0 += (int)(v[u] & 255u) * w[u]; a1 += (int)((v[u] >> 8) & 255u) * w[u];
                a2 += (int)((v[u] >> 16) & 255u) * w[u]; a3 += (int)(v[u] >> 24) * w[u];
            }
        }
        const uint32_t pk = (uint32_t)dfd_pil_clip8(a0) | ((uint32_t)dfd_pil_clip8(a1) << 8) | ((uint32_t)dfd_pil_clip8(a2) << 16) |
                            ((uint32_t)dfd_pil_clip8(a3) << 24);
        *(uint32_t*)&s160[r][xw * 4] = pk;
        if (face160) *(uint32_t*)(face160 + ((size_t)m * 160 + r_first + r) * 480 + xw * 4) = pk;
    }
    // bilinear source columns / weights of the 224 output columns: computed once per CTA instead of once per output value
    // (entry of column x at [(x % 8) * 28 + x / 8]: the lanes of a warp read column 8 * lane + px, i.e. consecutive entries)
    __shared__ __align__(16) float4 s_xc[224];             // {w0, w1, bits(x0 * 3), bits(x1 * 3)}
    for (int x = threadIdx.x; x < 224; x += 512) {
        int x0, x1; float w0, w1;
        dfd_torch_bilinear_coef(x, 160, 224, &x0, &x1, &w0, &w1);
        s_xc[(x & 7) * 28 + (x >> 3)] = make_float4(w0, w1, __int_as_float(x0 * 3), __int_as_float(x1 * 3));
    }
    __syncthreads();
    // (x / 255 - mean) / std with both divisions as exact 3-instruction constant divisions (dfd_div_const).
    // A thread produces 8 consecutive pixels of an output row (24 values: the channel of every value is a compile-time
    // constant, so the normalisation constants are immediates) and stores them as 16-byte vectors.
    constexpr float R255 = 1.0f / 255.0f;
    for (int o = threadIdx.x; o < 56 * 28; o += 512) {
        const int yr = o / 28, g8 = o - yr * 28;
        const int y = band * 56 + yr;
        int y0, y1; float h0, h1;
        dfd_torch_bilinear_coef(y, 160, 224, &y0, &y1, &h0, &h1);
        const uint8_t* r0 = s160[y0 - r_first];
        const uint8_t* r1 = s160[y1 - r_first];
        OutT* dst = out + (((size_t)m * 224 + y) * 224 + g8 * 8) * 3;
#pragma unroll
        for (int q = 0; q < 3; q++) {                      // 3 x 8 values
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int e = q * 8 + j, px = e / 3, c = e - px * 3;      // compile-time after unrolling
                const float4 xc = s_xc[px * 28 + g8];
                const int i0 = __float_as_int(xc.z) + c, i1 = __float_as_int(xc.w) + c;
                const float p00 = r0[i0], p01 = r0[i1], p10 = r1[i0], p11 = r1[i1];
                float t = DFD_FADD(DFD_FMUL(h0, DFD_FADD(DFD_FMUL(xc.x, p00), DFD_FMUL(xc.y, p01))),
                                   DFD_FMUL(h1, DFD_FADD(DFD_FMUL(xc.x, p10), DFD_FMUL(xc.y, p11))));
                t = dfd_div_const(t, 255.0f, R255);
                const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
                const float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
                const float rs = c == 0 ? 1.0f / 0.229f : (c == 1 ? 1.0f / 0.224f : 1.0f / 0.225f);
                v[j] = dfd_div_const(DFD_FSUB(t, mean), sd, rs);
            }
            store8<OutT>(dst + q * 8, v);
        }
    }
}

// Box sanitiser: the reference crops with numpy slicing (frame[y:y+h, x:x+w], deepfake_detection.py:612-619), which
// clamps a box to the frame; a box that is empty after clamping gives `analyze_face` -> (None, None, None).  Here:
// boxes are clamped to the H x W frame, frame indices to [0, n_frames); a box that is then empty, or whose side exceeds
// max_crop (the workspaces are sized for it), is replaced by a harmless 8 x 8 box and FLAGGED: k_faceprob turns its
// probability into NaN (= "no face" for the vote), so an oversized or out-of-frame box can never index out of bounds.
__global__ void k_box_sanitize(const int32_t* __restrict__ boxes, const int32_t* __restrict__ frame_idx, int m, int n_frames,
                               int H, int W, int max_crop, int32_t* __restrict__ boxes_ok, int32_t* __restrict__ fidx_ok,
                               uint8_t* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int x = boxes[i * 4], y = boxes[i * 4 + 1], w = boxes[i * 4 + 2], h = boxes[i * 4 + 3];
    int f = frame_idx[i];
    bool is_bad = f < 0 || f >= n_frames || w <= 0 || h <= 0;
    long long x1 = (long long)x + w, y1 = (long long)y + h;
    if (x < 0) x = 0;
    if (y < 0) y = 0;
    if (x1 > W) x1 = W;
    if (y1 > H) y1 = H;
    w = (int)(x1 - x); h = (int)(y1 - y);
    is_bad = is_bad || w <= 0 || h <= 0 || w > max_crop || h > max_crop;
    if (is_bad) { x = 0; y = 0; w = W < 8 ? W : 8; h = H < 8 ? H : 8; f = f < 0 || f >= n_frames ? 0 : f; }
    boxes_ok[i * 4] = x; boxes_ok[i * 4 + 1] = y; boxes_ok[i * 4 + 2] = w; boxes_ok[i * 4 + 3] = h;
    fidx_ok[i] = f;
    bad[i] = is_bad ? 1 : 0;
}

// n_pred = 1: plain face prep.  n_pred > 1 (test-time augmentation): crop q = i * n_pred + j of `out` is prediction j of box i
// (j = 0 the un-augmented crop, j >= 1 augmentation augs[i * (n_pred - 1) + j - 1]).
static int faceprep_launch(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, size_t frame_stride,
                           int row_pitch, const int32_t* boxes_in, const int32_t* frame_idx_in, int m, int n_pred,
                           const dfd_tta_aug* augs, void* out, int dtype, cudaStream_t st) {
    DFD_REQUIRE(m > 0 && n_pred >= 1 && (long long)m * n_pred <= ctx->cfg.max_batch, DFD_ERR_CAPACITY,
                "face_prep: box count (x predictions per box) exceeds max_batch");
    DFD_REQUIRE(dtype == DFD_F32 || dtype == DFD_BF16, DFD_ERR_INVALID, "face_prep: bad dtype");
    DFD_REQUIRE(n_frames > 0 && H >= 1 && W >= 1 && row_pitch >= 3 * W, DFD_ERR_INVALID, "face_prep: bad frame geometry");
    DFD_REQUIRE(n_pred == 1 || augs != nullptr, DFD_ERR_INVALID, "face_prep_tta: augmentation parameters missing");
    const int mc = ctx->cfg.max_crop;
    const size_t base_stride = (size_t)mc * mc * 3;
    uint8_t* base = nullptr;
    if (n_pred > 1) {
        int rc = dfd_ensure(ctx, ctx->tta_base, (size_t)m * base_stride);
        if (rc) return rc;
        base = (uint8_t*)ctx->tta_base.p;
    }
    k_box_sanitize<<<(m + 127) / 128, 128, 0, st>>>(boxes_in, frame_idx_in, m, n_frames, H, W, mc, ctx->d_boxes_ok, ctx->d_fidx_ok,
                                                    ctx->d_box_bad);
    DFD_LAUNCH_CHECK("k_box_sanitize", st);
    ctx->box_flags_m = m;
    const int32_t* boxes = ctx->d_boxes_ok;
    const int32_t* frame_idx = ctx->d_fidx_ok;
    k_pil_coeffs<<<m, 320, 0, st>>>(boxes, ctx->d_pil);
    DFD_LAUNCH_CHECK("k_pil_coeffs", st);
    const int lut_warps = m >= 16 ? 8 : 2;
    k_clahe_lut<<<dim3(64 / lut_warps, m), 32 * lut_warps, 0, st>>>(frames, frame_stride, row_pitch, boxes, frame_idx, ctx->d_tables, ctx->d_luts);
    DFD_LAUNCH_CHECK("k_clahe_lut", st);
    const size_t hp_smem = sizeof(DfdColorTables) + 64 * 256 + 160 * (2 + DFD_PIL_KMAX_STAGED) * sizeof(int) + (size_t)8 * ((mc * 3 + 8 + 15) & ~15);
    { int rc = dfd_func_smem(ctx, k_clahe_hpass, hp_smem); if (rc) return rc; }
    const int hp_rows = m >= 16 ? HP_ROWS : 8;
    k_clahe_hpass<<<dim3((mc + hp_rows - 1) / hp_rows, m), 256, hp_smem, st>>>(frames, frame_stride, row_pitch, boxes, frame_idx,
                                                                               ctx->d_tables, ctx->d_luts, ctx->d_pil, ctx->d_hpass, mc,
                                                                               base, base_stride, -1, n_pred, hp_rows);
    DFD_LAUNCH_CHECK("k_clahe_hpass", st);
    if (n_pred > 1) {
        const size_t tta_smem = 256 + 160 * (2 + DFD_PIL_KMAX_STAGED) * sizeof(int) + (size_t)8 * ((mc * 3 + 8 + 15) & ~15);
        { int rc = dfd_func_smem(ctx, k_tta_hpass, tta_smem); if (rc) return rc; }
        k_tta_hpass<<<dim3((mc + hp_rows - 1) / hp_rows, m * (n_pred - 1)), 256, tta_smem, st>>>(
            base, base_stride, boxes, (const DfdTtaAug*)augs, ctx->d_pil, ctx->d_hpass, mc, n_pred, hp_rows);
        DFD_LAUNCH_CHECK("k_tta_hpass", st);
    }
    if (dtype == DFD_F32)
        k_vpass_up_norm<float><<<dim3(4, m * n_pred), 512, 0, st>>>(boxes, ctx->d_pil, ctx->d_hpass, mc, ctx->d_face160, (float*)out, n_pred);
    else
        k_vpass_up_norm<__nv_bfloat16><<<dim3(4, m * n_pred), 512, 0, st>>>(boxes, ctx->d_pil, ctx->d_hpass, mc, ctx->d_face160,
                                                                            (__nv_bfloat16*)out, n_pred);
    DFD_LAUNCH_CHECK("k_vpass_up_norm", st);
    return DFD_OK;
}

int dfd_faceprep_launch(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, size_t frame_stride,
                        int row_pitch, const int32_t* boxes_in, const int32_t* frame_idx_in, int m, void* out, int dtype,
                        cudaStream_t st) {
    return faceprep_launch(ctx, frames, n_frames, H, W, frame_stride, row_pitch, boxes_in, frame_idx_in, m, 1, nullptr, out, dtype, st);
}

int dfd_faceprep_tta_launch(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, size_t frame_stride,
                            int row_pitch, const int32_t* boxes_in, const int32_t* frame_idx_in, int m, int n_pred,
                            const dfd_tta_aug* augs, void* out, int dtype, cudaStream_t st) {
    DFD_REQUIRE(n_pred >= 1 && n_pred <= DFD_TTA_MAX_PRED, DFD_ERR_INVALID, "face_prep_tta: n_pred outside 1..DFD_TTA_MAX_PRED");
    return faceprep_launch(ctx, frames, n_frames, H, W, frame_stride, row_pitch, boxes_in, frame_idx_in, m, n_pred, augs, out, dtype, st);
}

int dfd_dbg_clahe_launch(dfd_ctx* ctx, const uint8_t* frames, size_t frame_stride, int row_pitch, const int32_t* boxes,
                         const int32_t* frame_idx, int i, uint8_t* out, cudaStream_t st) {
    // luts of the last dfd_face_prep_batch call are reused; only box i is written
    const int mc = ctx->cfg.max_crop;
    const size_t hp_smem = sizeof(DfdColorTables) + 64 * 256 + 160 * (2 + DFD_PIL_KMAX_STAGED) * sizeof(int) + (size_t)8 * ((mc * 3 + 8 + 15) & ~15);
    { int rc = dfd_func_smem(ctx, k_clahe_hpass, hp_smem); if (rc) return rc; }
    // (the sanitised boxes of that call are used as well: see k_box_sanitize)
    (void)boxes; (void)frame_idx;
    k_clahe_hpass<<<dim3((mc + HP_ROWS - 1) / HP_ROWS, i + 1), 256, hp_smem, st>>>(frames, frame_stride, row_pitch, ctx->d_boxes_ok, ctx->d_fidx_ok,
                                                                                   ctx->d_tables, ctx->d_luts, ctx->d_pil, nullptr, mc, out, 0, i, 1, HP_ROWS);
    DFD_LAUNCH_CHECK("k_clahe_hpass", st);
    return DFD_OK;
}
