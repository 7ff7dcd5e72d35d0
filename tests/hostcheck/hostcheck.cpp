// TEST INFRASTRUCTURE: compiles the product's __host__ __device__ pixel math
// (csrc/px_*.h) for the CPU so tests can compare it with cv2 / PIL without a
// GPU.  Not linked into libdfd.so; never used by the product.
#include "px_color.h"
#include <cstring>

extern "C" {

static DfdColorTables g_T;
void hc_init(int cbrt_mode, int inv_div) { dfd_build_color_tables(&g_T, cbrt_mode, inv_div); }

void hc_bgr2gray(const uint8_t* src, uint8_t* dst, long n) {
    for (long i = 0; i < n; i++) dst[i] = (uint8_t)dfd_bgr2gray(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
}
void hc_bgr2hsv(const uint8_t* src, uint8_t* dst, long n) {
    for (long i = 0; i < n; i++) {
        int h, s, v;
        dfd_bgr2hsv(&g_T, src[3 * i], src[3 * i + 1], src[3 * i + 2], &h, &s, &v);
        dst[3 * i] = h; dst[3 * i + 1] = s; dst[3 * i + 2] = v;
    }
}
void hc_bgr2lab(const uint8_t* src, uint8_t* dst, long n) {
    for (long i = 0; i < n; i++) {
        int l, a, b;
        dfd_bgr2lab(&g_T, src[3 * i], src[3 * i + 1], src[3 * i + 2], &l, &a, &b);
        dst[3 * i] = l; dst[3 * i + 1] = a; dst[3 * i + 2] = b;
    }
}
void hc_lab2bgr(const uint8_t* src, uint8_t* dst, long n) {
    for (long i = 0; i < n; i++) {
        int b, g, r;
        dfd_lab2bgr(&g_T, src[3 * i], src[3 * i + 1], src[3 * i + 2], &b, &g, &r);
        dst[3 * i] = b; dst[3 * i + 1] = g; dst[3 * i + 2] = r;
    }
}
}

#include "px_resize.h"
#include "px_clahe.h"
#include "px_jpeg.h"
#include "px_canny.h"
#include "px_numpy.h"
#include <vector>

extern "C" {

// cv2.resize(src (h,w,3), (dw,dh), INTER_LINEAR)
void hc_cvresize(const uint8_t* src, int h, int w, uint8_t* dst, int dh, int dw) {
    for (int y = 0; y < dh; y++) {
        int sy0, sy1, b0, b1;
        dfd_cvresize_coef(y, h, dh, 0, &sy0, &sy1, &b0, &b1);
        for (int x = 0; x < dw; x++) {
            int sx0, sx1, a0, a1;
            dfd_cvresize_coef(x, w, dw, 1, &sx0, &sx1, &a0, &a1);
            for (int c = 0; c < 3; c++) {
                int p00 = src[(sy0 * w + sx0) * 3 + c], p01 = src[(sy0 * w + sx1) * 3 + c];
                int p10 = src[(sy1 * w + sx0) * 3 + c], p11 = src[(sy1 * w + sx1) * 3 + c];
                dst[(y * dw + x) * 3 + c] = (uint8_t)dfd_cvresize_px(p00, p01, p10, p11, a0, a1, b0, b1);
            }
        }
    }
}

// CLAHE on a single u8 plane
void hc_clahe(const uint8_t* src, int h, int w, uint8_t* dst) {
    DfdClaheGeom g = dfd_clahe_geom(w, h);
    std::vector<uint8_t> luts(64 * 256);
    for (int ty = 0; ty < 8; ty++)
        for (int tx = 0; tx < 8; tx++) {
            int hist[256] = {0};
            for (int y = ty * g.th; y < (ty + 1) * g.th; y++)
                for (int x = tx * g.tw; x < (tx + 1) * g.tw; x++) {
                    int sx = x < w ? x : dfd_reflect101(x, w), sy = y < h ? y : dfd_reflect101(y, h);
                    hist[src[sy * w + sx]]++;
                }
            dfd_clahe_lut(hist, g.clip, g.lut_scale, &luts[(ty * 8 + tx) * 256]);
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) dst[y * w + x] = (uint8_t)dfd_clahe_apply(luts.data(), g, x, y, src[y * w + x]);
}

// Pillow resize of an RGB u8 image to (out,out) BILINEAR: horizontal then vertical pass
void hc_pil_resize(const uint8_t* src, int h, int w, uint8_t* dst, int out) {
    std::vector<uint8_t> tmp((size_t)h * out * 3);
    int k[DFD_PIL_KMAX];
    for (int xx = 0; xx < out; xx++) {
        int xmin, n = dfd_pil_coeffs(xx, w, out, &xmin, k);
        for (int y = 0; y < h; y++)
            for (int c = 0; c < 3; c++) {
                int acc = 1 << (DFD_PIL_PRECISION - 1);
                for (int t = 0; t < n; t++) acc += src[(y * w + xmin + t) * 3 + c] * k[t];
                tmp[((size_t)y * out + xx) * 3 + c] = (uint8_t)dfd_pil_clip8(acc);
            }
    }
    for (int yy = 0; yy < out; yy++) {
        int ymin, n = dfd_pil_coeffs(yy, h, out, &ymin, k);
        for (int x = 0; x < out; x++)
            for (int c = 0; c < 3; c++) {
                int acc = 1 << (DFD_PIL_PRECISION - 1);
                for (int t = 0; t < n; t++) acc += tmp[((size_t)(ymin + t) * out + x) * 3 + c] * k[t];
                dst[(yy * out + x) * 3 + c] = (uint8_t)dfd_pil_clip8(acc);
            }
    }
}

// F.interpolate(160->224) + /255 + normalise: src RGB u8 (in,in,3) -> dst float CHW (3,out,out)
void hc_torch_up_norm(const uint8_t* src, int in, float* dst, int out) {
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (int y = 0; y < out; y++) {
        int y0, y1; float h0, h1;
        dfd_torch_bilinear_coef(y, in, out, &y0, &y1, &h0, &h1);
        for (int x = 0; x < out; x++) {
            int x0, x1; float w0, w1;
            dfd_torch_bilinear_coef(x, in, out, &x0, &x1, &w0, &w1);
            for (int c = 0; c < 3; c++) {
                float p00 = src[(y0 * in + x0) * 3 + c], p01 = src[(y0 * in + x1) * 3 + c];
                float p10 = src[(y1 * in + x0) * 3 + c], p11 = src[(y1 * in + x1) * 3 + c];
                float v = DFD_FADD(DFD_FMUL(h0, DFD_FADD(DFD_FMUL(w0, p00), DFD_FMUL(w1, p01))),
                                   DFD_FMUL(h1, DFD_FADD(DFD_FMUL(w0, p10), DFD_FMUL(w1, p11))));
                v = v / 255.0f;
                dst[(c * out + y) * out + x] = (v - mean[c]) / stdv[c];
            }
        }
    }
}

// JPEG Q90 4:2:0 round trip of a BGR image with h,w multiples of 16
void hc_jpeg_roundtrip(const uint8_t* src, int h, int w, uint8_t* dst) {
    int cw = w / 2, ch = h / 2;
    std::vector<uint8_t> Y(h * w), Cb(h * w), Cr(h * w), cb(cw * ch), cr(cw * ch);
    for (int i = 0; i < h * w; i++) {
        int b = src[3 * i], g = src[3 * i + 1], r = src[3 * i + 2];
        Y[i] = dfd_jpeg_y(r, g, b); Cb[i] = dfd_jpeg_cb(r, g, b); Cr[i] = dfd_jpeg_cr(r, g, b);
    }
    for (int y = 0; y < ch; y++)
        for (int x = 0; x < cw; x++) {
            int bias = (x & 1) ? 2 : 1;
            int i0 = (2 * y) * w + 2 * x, i1 = i0 + w;
            cb[y * cw + x] = (Cb[i0] + Cb[i0 + 1] + Cb[i1] + Cb[i1 + 1] + bias) >> 2;
            cr[y * cw + x] = (Cr[i0] + Cr[i0 + 1] + Cr[i1] + Cr[i1 + 1] + bias) >> 2;
        }
    auto plane = [](std::vector<uint8_t>& p, int pw, int ph, int chroma) {
        int blk[64];
        for (int by = 0; by < ph; by += 8)
            for (int bx = 0; bx < pw; bx += 8) {
                for (int i = 0; i < 64; i++) blk[i] = p[(by + i / 8) * pw + bx + i % 8];
                dfd_jpeg_block_roundtrip(blk, chroma);
                for (int i = 0; i < 64; i++) p[(by + i / 8) * pw + bx + i % 8] = (uint8_t)blk[i];
            }
    };
    plane(Y, w, h, 0); plane(cb, cw, ch, 1); plane(cr, cw, ch, 1);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int r, g, b;
            dfd_jpeg_ycc2rgb(Y[y * w + x], dfd_jpeg_fancy_up(cb.data(), cw, ch, x, y),
                             dfd_jpeg_fancy_up(cr.data(), cw, ch, x, y), &r, &g, &b);
            dst[(y * w + x) * 3] = b; dst[(y * w + x) * 3 + 1] = g; dst[(y * w + x) * 3 + 2] = r;
        }
}

// Canny(50,150) on a gray image -> 0/255 map
void hc_canny(const uint8_t* g, int h, int w, uint8_t* dst) {
    std::vector<int> mag(h * w), dxs(h * w), dys(h * w);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int dx, dy; dfd_sobel3(g, w, h, x, y, &dx, &dy);
            dxs[y * w + x] = dx; dys[y * w + x] = dy; mag[y * w + x] = dfd_absi(dx) + dfd_absi(dy);
        }
    auto M = [&](int x, int y) { return (x < 0 || y < 0 || x >= w || y >= h) ? 0 : mag[y * w + x]; };
    std::vector<uint8_t> st(h * w);
    std::vector<int> stack;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            st[y * w + x] = dfd_canny_nms(dxs[y * w + x], dys[y * w + x], mag[y * w + x], x, y, M);
            // the packed form the CUDA kernel uses must give the same state
            const unsigned c = dfd_canny_pack(dxs[y * w + x], dys[y * w + x]);
            const int off = dfd_canny_first_off(c >> 11, w);
            const int ox = off == -1 ? -1 : off == -w ? 0 : off == -w - 1 ? -1 : 1, oy = off == -1 ? 0 : -1;
            if (st[y * w + x] != dfd_canny_nms_packed(c, M(x + ox, y + oy), M(x - ox, y - oy))) st[y * w + x] = 77;
            if (st[y * w + x] == 2) stack.push_back(y * w + x);
        }
    while (!stack.empty()) {
        int p = stack.back(); stack.pop_back();
        int px = p % w, py = p / w;
        for (int j = -1; j <= 1; j++)
            for (int i = -1; i <= 1; i++) {
                int x = px + i, y = py + j;
                if (x < 0 || y < 0 || x >= w || y >= h) continue;
                if (st[y * w + x] == 1) { st[y * w + x] = 2; stack.push_back(y * w + x); }
            }
    }
    for (int i = 0; i < h * w; i++) dst[i] = st[i] == 2 ? 255 : 0;
}

void hc_laplacian(const uint8_t* g, int h, int w, double* dst) {
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) dst[y * w + x] = dfd_laplacian(g, w, h, x, y);
}
void hc_gauss_resid(const uint8_t* g, int h, int w, float* dst) {    // gray - blur
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
        dst[y * w + x] = (float)g[y * w + x] - (float)dfd_gauss5_x256(g, w, h, x, y) / 256.0f;
}
float hc_np_mean(const float* a, int n) { return dfd_np_mean_f32(a, n); }
double hc_py_sum_products(const double* s, const double* w, int n) { return dfd_py_sum_products(s, w, n); }
void hc_py_sum_products_batch(const double* s, const double* w, int n, long count, double* out) {
    for (long i = 0; i < count; i++) out[i] = dfd_py_sum_products(s + i * n, w, n);
}
float hc_np_std(const float* a, int n) { std::vector<float> t(n); return dfd_np_std_f32(a, n, t.data()); }
}

// ---------------------------------------------------------------------------------------------
// Baseline JPEG decode with the product's functions (csrc/px_jpegdec.h), on the CPU: sequentially (sub_bits = 0) or as a
// faithful simulation of the GPU kernel's self-synchronising parallel schedule (sub_bits = subsequence size in bits).
#include "px_jpegdec.h"
#include <string.h>
extern "C" {
int hc_jpeg_info(const uint8_t* data, long n, int* info /* H, W, ncomp, hs0, vs0, ecs bytes, total blocks */) {
    DfdJpegHeader* h = new DfdJpegHeader;
    int rc = dfd_jpeg_parse(data, (size_t)n, h);
    info[0] = h->height; info[1] = h->width; info[2] = h->ncomp; info[3] = h->hs[0]; info[4] = h->vs[0];
    info[5] = h->ecs_end - h->ecs_begin; info[6] = h->total_blocks;
    delete h;
    return rc;
}

int hc_jpeg_decode(const uint8_t* data, long n, uint8_t* out, int sub_bits, int* rounds_out) {
    DfdJpegHeader* h = new DfdJpegHeader;
    int rc = dfd_jpeg_parse(data, (size_t)n, h);
    if (rc) { delete h; return rc; }
    // 1. remove byte stuffing, pack MSB-first words
    std::vector<uint8_t> clean;
    for (int i = h->ecs_begin; i < h->ecs_end; i++) {
        clean.push_back(data[i]);
        if (data[i] == 0xFF && i + 1 < h->ecs_end && data[i + 1] == 0x00) i++;
    }
    const uint32_t nbits = (uint32_t)clean.size() * 8;
    while (clean.size() % 4) clean.push_back(0);
    std::vector<uint32_t> words(clean.size() / 4 + 2, 0);
    for (size_t w = 0; w < clean.size() / 4; w++)
        words[w] = ((uint32_t)clean[4 * w] << 24) | ((uint32_t)clean[4 * w + 1] << 16) | ((uint32_t)clean[4 * w + 2] << 8) | clean[4 * w + 3];
    const uint32_t nwords = (uint32_t)words.size();
    // 2. Huffman decode
    std::vector<int16_t> coef((size_t)h->total_blocks * 64, 0);
    int32_t dc_off[3] = {0, 0, 0};
    int ndc = 0;
    for (int c = 0; c < h->ncomp; c++) { dc_off[c] = ndc; ndc += h->comp_bw[c] * h->comp_bh[c]; }
    std::vector<int32_t> dcdiff(ndc, 0);
    int rounds = 0, total = 0, nb, ne;
    if (sub_bits <= 0) {
        DfdJpegState s0 = {0, 0, 0};
        dfd_jpeg_decode_sub<true>(h, words.data(), nwords, dfd_jpeg_pack_state(s0), nbits, &nb, &ne, 0, coef.data(), dcdiff.data(), dc_off);
        total = nb;
    } else {
        const int nsub = (int)((nbits + sub_bits - 1) / sub_bits);
        std::vector<uint64_t> E(nsub), used(nsub);
        std::vector<int> cnt(nsub);
        for (int i = 0; i < nsub; i++) {                           // blind pass
            DfdJpegState s = {(uint32_t)i * (uint32_t)sub_bits, 0, 0};
            used[i] = dfd_jpeg_pack_state(s);
            const uint32_t lim = (uint32_t)std::min<uint64_t>((uint64_t)(i + 1) * sub_bits, nbits);
            E[i] = dfd_jpeg_decode_sub<false>(h, words.data(), nwords, used[i], lim, &cnt[i], &ne, 0, nullptr, nullptr, nullptr);
        }
        bool changed = true;
        while (changed) {                                          // synchronisation rounds
            changed = false;
            rounds++;
            std::vector<uint64_t> snap = E;                        // (a round reads the previous round's states)
            for (int i = 1; i < nsub; i++) {
                if (used[i] == snap[i - 1]) continue;
                used[i] = snap[i - 1];
                const uint32_t lim = (uint32_t)std::min<uint64_t>((uint64_t)(i + 1) * sub_bits, nbits);
                const uint64_t e = dfd_jpeg_decode_sub<false>(h, words.data(), nwords, used[i], lim, &cnt[i], &ne, 0, nullptr, nullptr, nullptr);
                if (e != E[i]) changed = true;                     // its successor must look at it again
                E[i] = e;
            }
            if (rounds > nsub + 2) break;
        }
        std::vector<int> blk0(nsub + 1, 0);
        for (int i = 0; i < nsub; i++) blk0[i + 1] = blk0[i] + cnt[i];
        total = blk0[nsub];
        for (int i = 0; i < nsub; i++) {                           // write pass
            const uint32_t lim = (uint32_t)std::min<uint64_t>((uint64_t)(i + 1) * sub_bits, nbits);
            dfd_jpeg_decode_sub<true>(h, words.data(), nwords, used[i], lim, &nb, &ne, blk0[i], coef.data(), dcdiff.data(), dc_off);
        }
    }
    if (rounds_out) *rounds_out = rounds;
    if (total < h->total_blocks) { delete h; return DFD_JPEG_ERR_DATA; }
    // 3. DC prediction chains
    for (int c = 0; c < h->ncomp; c++) {
        int acc = 0;
        const int nbk = h->comp_bw[c] * h->comp_bh[c];
        for (int i = 0; i < nbk; i++) { acc += dcdiff[dc_off[c] + i]; dcdiff[dc_off[c] + i] = acc; }
    }
    // 4. IDCT into component planes
    std::vector<std::vector<uint8_t>> plane(h->ncomp);
    for (int c = 0; c < h->ncomp; c++) {
        const int pw = h->comp_bw[c] * 8, ph = h->comp_bh[c] * 8;
        plane[c].assign((size_t)pw * ph, 0);
        for (int by = 0; by < h->comp_bh[c]; by++)
            for (int bx = 0; bx < h->comp_bw[c]; bx++) {
                uint8_t o[64];
                const int index = h->comp_blk0[c] + by * h->comp_bw[c] + bx;
                dfd_jpeg_idct_block(&coef[(size_t)index * 64], dcdiff[dc_off[c] + dfd_jpeg_dc_seq(h, c, bx, by)], h->qt[c], o);
                for (int i = 0; i < 64; i++) plane[c][(size_t)(by * 8 + i / 8) * pw + bx * 8 + i % 8] = o[i];
            }
    }
    // 5. up-sampling + colour conversion
    const int W = h->width, H = h->height;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int Y = plane[0][(size_t)y * h->comp_bw[0] * 8 + x];
            int r = Y, g = Y, b = Y;
            if (h->ncomp == 3) {
                int cc[2];
                for (int c = 1; c < 3; c++) {
                    const int cw = (W * h->hs[c] + h->hmax - 1) / h->hmax, chh = (H * h->vs[c] + h->vmax - 1) / h->vmax;
                    cc[c - 1] = dfd_jpeg_chroma_at(plane[c].data(), h->comp_bw[c] * 8, cw, chh, h->hs[c], h->vs[c], h->hmax, h->vmax, x, y);
                }
                dfd_jpeg_ycc2rgb(Y, cc[0], cc[1], &r, &g, &b);
            }
            out[((size_t)y * W + x) * 3] = (uint8_t)b; out[((size_t)y * W + x) * 3 + 1] = (uint8_t)g; out[((size_t)y * W + x) * 3 + 2] = (uint8_t)r;
        }
    delete h;
    return 0;
}
}

// ---------------------------------------------------------------------------------------------
// Test-time augmentation (csrc/px_warp.h): flip -> convertScaleAbs -> warpAffine of a whole w x h x 3 image.
#include "px_warp.h"
extern "C" {
void hc_tta_augment(const uint8_t* src, int h, int w, int flip, float alpha, const double* im, uint8_t* dst) {
    uint8_t lut[256];
    for (int v = 0; v < 256; v++) lut[v] = (uint8_t)dfd_scale_abs_u8(v, alpha);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int o0, o1, o2;
            dfd_tta_pixel(src, w * 3, w, h, flip, lut, im, x, y, &o0, &o1, &o2);
            uint8_t* d = dst + ((size_t)y * w + x) * 3;
            d[0] = (uint8_t)o0; d[1] = (uint8_t)o1; d[2] = (uint8_t)o2;
        }
}
}
