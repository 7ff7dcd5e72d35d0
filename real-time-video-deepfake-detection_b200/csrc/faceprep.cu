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
//   k_clahe_lut         per (box, CLAHE tile): L histogram in smem -> clipped, redistributed LUT
//   k_clahe_hpass       per (box, crop row): LAB/CLAHE/LAB2BGR row into smem, Pillow horizontal pass
//   k_vpass_up_norm     per (box, 56-row band): Pillow vertical pass into smem, bilinear 224, normalise
#include "dfd_internal.cuh"
#include "px_resize.h"
#include "px_clahe.h"
#include "px_numpy.h"

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

__global__ void __launch_bounds__(256) k_clahe_lut(const uint8_t* __restrict__ frames, size_t fstride, int pitch,
                                                   const int32_t* __restrict__ boxes, const int32_t* __restrict__ frame_idx,
                                                   const DfdColorTables* __restrict__ tab, uint8_t* __restrict__ luts) {
    const int m = blockIdx.y, tile = blockIdx.x;
    const int ty = tile >> 3, tx = tile & 7;
    const int bx = boxes[m * 4], by = boxes[m * 4 + 1], bw = boxes[m * 4 + 2], bh = boxes[m * 4 + 3];
    const DfdClaheGeom g = dfd_clahe_geom(bw, bh);
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint8_t* f = frames + (size_t)frame_idx[m] * fstride;
    const int area = g.tw * g.th;
    for (int p = threadIdx.x; p < area; p += 256) {
        int y = ty * g.th + p / g.tw, x = tx * g.tw + p % g.tw;
        if (x >= bw) x = dfd_reflect101(x, bw);
        if (y >= bh) y = dfd_reflect101(y, bh);
        const uint8_t* px = f + (size_t)(by + y) * pitch + (size_t)(bx + x) * 3;
        int L, A, B;
        dfd_bgr2lab(tab, px[0], px[1], px[2], &L, &A, &B);
        atomicAdd(&hist[L], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) dfd_clahe_lut(hist, g.clip, g.lut_scale, luts + ((size_t)m * 64 + tile) * 256);
}

__global__ void __launch_bounds__(256) k_clahe_hpass(const uint8_t* __restrict__ frames, size_t fstride, int pitch,
                                                     const int32_t* __restrict__ boxes, const int32_t* __restrict__ frame_idx,
                                                     const DfdColorTables* __restrict__ tab, const uint8_t* __restrict__ luts,
                                                     const int* __restrict__ pil, uint8_t* __restrict__ hpass, int max_crop,
                                                     uint8_t* __restrict__ dbg_clahe, int dbg_box) {
    extern __shared__ __align__(16) uint8_t row[];           // bw * 3 RGB
    const int m = blockIdx.y, y = blockIdx.x;
    const int bx = boxes[m * 4], by = boxes[m * 4 + 1], bw = boxes[m * 4 + 2], bh = boxes[m * 4 + 3];
    if (y >= bh) return;
    const DfdClaheGeom g = dfd_clahe_geom(bw, bh);
    const uint8_t* f = frames + (size_t)frame_idx[m] * fstride + (size_t)(by + y) * pitch + (size_t)bx * 3;
    const uint8_t* lut = luts + (size_t)m * 64 * 256;
    for (int x = threadIdx.x; x < bw; x += 256) {
        int L, A, B, ob, og, orr;
        dfd_bgr2lab(tab, f[x * 3], f[x * 3 + 1], f[x * 3 + 2], &L, &A, &B);
        L = dfd_clahe_apply(lut, g, x, y, L);
        dfd_lab2bgr(tab, L, A, B, &ob, &og, &orr);
        row[x * 3] = (uint8_t)orr; row[x * 3 + 1] = (uint8_t)og; row[x * 3 + 2] = (uint8_t)ob;   // RGB order (:376)
        if (dbg_clahe && m == dbg_box) {
            uint8_t* d = dbg_clahe + ((size_t)y * bw + x) * 3;
            d[0] = (uint8_t)ob; d[1] = (uint8_t)og; d[2] = (uint8_t)orr;
        }
    }
    __syncthreads();
    if (!hpass) return;
    const int* pc = pil + ((size_t)m * 2 + 0) * 160 * PIL_STRIDE;
    uint8_t* out = hpass + (((size_t)m * max_crop + y) * 160) * 3;
    for (int o = threadIdx.x; o < 480; o += 256) {
        int xx = o / 3, c = o % 3;
        const int* k = pc + xx * PIL_STRIDE;
        int xmin = k[0], cnt = k[1];
        int acc = 1 << (DFD_PIL_PRECISION - 1);
        for (int t = 0; t < cnt; t++) acc += row[(xmin + t) * 3 + c] * k[2 + t];
        out[o] = (uint8_t)dfd_pil_clip8(acc);
    }
}

template <typename OutT>
__device__ __forceinline__ void store_px(OutT* o, float a, float b, float c);
template <>
__device__ __forceinline__ void store_px<float>(float* o, float a, float b, float c) { o[0] = a; o[1] = b; o[2] = c; }
template <>
__device__ __forceinline__ void store_px<__nv_bfloat16>(__nv_bfloat16* o, float a, float b, float c) {
    o[0] = __float2bfloat16_rn(a); o[1] = __float2bfloat16_rn(b); o[2] = __float2bfloat16_rn(c);
}

template <typename OutT>
__global__ void __launch_bounds__(512) k_vpass_up_norm(const int32_t* __restrict__ boxes, const int* __restrict__ pil,
                                                       const uint8_t* __restrict__ hpass, int max_crop,
                                                       uint8_t* __restrict__ face160, OutT* __restrict__ out) {
    __shared__ uint8_t s160[44][480];
    const int m = blockIdx.y, band = blockIdx.x;           // 4 bands of 56 output rows
    int r_first, r_last, t0, t1; float l0, l1;
    dfd_torch_bilinear_coef(band * 56, 160, 224, &r_first, &t1, &l0, &l1);
    dfd_torch_bilinear_coef(band * 56 + 55, 160, 224, &t0, &r_last, &l0, &l1);
    const int nrows = r_last - r_first + 1;                // <= 42
    const int* pc = pil + ((size_t)m * 2 + 1) * 160 * PIL_STRIDE;
    const uint8_t* hp = hpass + (size_t)m * max_crop * 480;
    for (int o = threadIdx.x; o < nrows * 480; o += 512) {
        int r = o / 480, xc = o % 480;
        const int* k = pc + (r_first + r) * PIL_STRIDE;
        int ymin = k[0], cnt = k[1];
        int acc = 1 << (DFD_PIL_PRECISION - 1);
        for (int t = 0; t < cnt; t++) acc += hp[(size_t)(ymin + t) * 480 + xc] * k[2 + t];
        uint8_t v = (uint8_t)dfd_pil_clip8(acc);
        s160[r][xc] = v;
        if (face160) face160[((size_t)m * 160 + r_first + r) * 480 + xc] = v;
    }
    __syncthreads();
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (int o = threadIdx.x; o < 56 * 224; o += 512) {
        int y = band * 56 + o / 224, x = o % 224;
        int y0, y1, x0, x1; float h0, h1, w0, w1;
        dfd_torch_bilinear_coef(y, 160, 224, &y0, &y1, &h0, &h1);
        dfd_torch_bilinear_coef(x, 160, 224, &x0, &x1, &w0, &w1);
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float p00 = s160[y0 - r_first][x0 * 3 + c], p01 = s160[y0 - r_first][x1 * 3 + c];
            float p10 = s160[y1 - r_first][x0 * 3 + c], p11 = s160[y1 - r_first][x1 * 3 + c];
            float t = DFD_FADD(DFD_FMUL(h0, DFD_FADD(DFD_FMUL(w0, p00), DFD_FMUL(w1, p01))),
                               DFD_FMUL(h1, DFD_FADD(DFD_FMUL(w0, p10), DFD_FMUL(w1, p11))));
            t = DFD_FDIV(t, 255.0f);
            v[c] = DFD_FDIV(DFD_FSUB(t, mean[c]), stdv[c]);
        }
        store_px<OutT>(out + (((size_t)m * 224 + y) * 224 + x) * 3, v[0], v[1], v[2]);
    }
}

int dfd_faceprep_launch(dfd_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, size_t frame_stride,
                        int row_pitch, const int32_t* boxes, const int32_t* frame_idx, int m, void* out, int dtype,
                        cudaStream_t st) {
    DFD_REQUIRE(m > 0 && m <= ctx->cfg.max_batch, DFD_ERR_CAPACITY, "face_prep: box count exceeds max_batch");
    DFD_REQUIRE(dtype == DFD_F32 || dtype == DFD_BF16, DFD_ERR_INVALID, "face_prep: bad dtype");
    const int mc = ctx->cfg.max_crop;
    k_pil_coeffs<<<m, 320, 0, st>>>(boxes, ctx->d_pil);
    DFD_LAUNCH_CHECK("k_pil_coeffs", st);
    k_clahe_lut<<<dim3(64, m), 256, 0, st>>>(frames, frame_stride, row_pitch, boxes, frame_idx, ctx->d_tables, ctx->d_luts);
    DFD_LAUNCH_CHECK("k_clahe_lut", st);
    k_clahe_hpass<<<dim3(mc, m), 256, mc * 3, st>>>(frames, frame_stride, row_pitch, boxes, frame_idx, ctx->d_tables,
                                                    ctx->d_luts, ctx->d_pil, ctx->d_hpass, mc, nullptr, -1);
    DFD_LAUNCH_CHECK("k_clahe_hpass", st);
    if (dtype == DFD_F32)
        k_vpass_up_norm<float><<<dim3(4, m), 512, 0, st>>>(boxes, ctx->d_pil, ctx->d_hpass, mc, ctx->d_face160, (float*)out);
    else
        k_vpass_up_norm<__nv_bfloat16><<<dim3(4, m), 512, 0, st>>>(boxes, ctx->d_pil, ctx->d_hpass, mc, ctx->d_face160,
                                                                   (__nv_bfloat16*)out);
    DFD_LAUNCH_CHECK("k_vpass_up_norm", st);
    return DFD_OK;
}

int dfd_dbg_clahe_launch(dfd_ctx* ctx, const uint8_t* frames, size_t frame_stride, int row_pitch, const int32_t* boxes,
                         const int32_t* frame_idx, int i, uint8_t* out, cudaStream_t st) {
    // luts of the last dfd_face_prep_batch call are reused; only box i is written
    const int mc = ctx->cfg.max_crop;
    k_clahe_hpass<<<dim3(mc, i + 1), 256, mc * 3, st>>>(frames, frame_stride, row_pitch, boxes, frame_idx, ctx->d_tables,
                                                       ctx->d_luts, ctx->d_pil, nullptr, mc, out, i);
    DFD_LAUNCH_CHECK("k_clahe_hpass", st);
    return DFD_OK;
}
