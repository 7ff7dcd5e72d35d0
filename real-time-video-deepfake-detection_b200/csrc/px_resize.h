// Bit-exact resampling coefficients.
//   cv2.resize(frame,(256,256),INTER_LINEAR) on 8UC3   frame_analysis.py:71,111   
//   PIL Image.resize((160,160), BILINEAR) on RGB        MTCNN extract_face step of
//                                                       deepfake_detection.py:377   
//   F.interpolate(224, bilinear, align_corners=False)   deepfake_detection.py:383
#pragma once
#include "px_common.h"

// ---- OpenCV fixed-point bilinear (11-bit coefficients) ----------------------
// Horizontal taps zero the fractional weight at the borders; vertical taps
// only clamp the row index.
DFD_HD void dfd_cvresize_coef(int d, int src, int dst, int horizontal, int* s0, int* s1, int* a0, int* a1) {
    double inv_scale = (double)dst / (double)src;
    double scale = 1.0 / inv_scale;
    float f = (float)DFD_DSUB(DFD_DMUL((double)d + 0.5, scale), 0.5);
    int s = (int)floorf(f);
    float fr = DFD_FSUB(f, (float)s);
    if (horizontal) {
        if (s < 0) { s = 0; fr = 0.f; }
        if (s >= src - 1) { s = src - 1; fr = 0.f; }
        *s0 = s; *s1 = s + 1 < src ? s + 1 : src - 1;
    } else {
        *s0 = dfd_clampi(s, 0, src - 1);
        *s1 = dfd_clampi(s + 1, 0, src - 1);
    }
    *a0 = DFD_RINTF(DFD_FMUL(DFD_FSUB(1.f, fr), 2048.f));
    *a1 = DFD_RINTF(DFD_FMUL(fr, 2048.f));
}

// p00,p01 = taps of row s0 ; p10,p11 = taps of row s1
DFD_HD int dfd_cvresize_px(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1) {
    int r0 = p00 * a0 + p01 * a1;
    int r1 = p10 * a0 + p11 * a1;
    return (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
}

// ---- Pillow antialiased triangle filter (Resample.c precompute_coeffs) ------
#define DFD_PIL_KMAX 64              // ksize = 2*ceil(in/out)+1  -> in <= 31*out
#define DFD_PIL_PRECISION 22

DFD_HD int dfd_pil_ksize(int in_size, int out_size) {
    double scale = DFD_DDIV((double)in_size, (double)out_size);
    double support = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(support) * 2 + 1;
}

// Coefficients for output index xx; returns tap count, writes first tap to *xmin and
// fixed-point weights to k[0..count).
DFD_HD int dfd_pil_coeffs(int xx, int in_size, int out_size, int* xmin_out, int* k) {
    double scale = DFD_DDIV((double)in_size, (double)out_size);
    double filterscale = scale < 1.0 ? 1.0 : scale;
    double support = filterscale;                     // bilinear support 1.0 * filterscale
    double center = DFD_DMUL((double)xx + 0.5, scale);
    double ss = DFD_DDIV(1.0, filterscale);
    int xmin = (int)DFD_DADD(DFD_DSUB(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)DFD_DADD(DFD_DADD(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double w[DFD_PIL_KMAX];
    double ww = 0.0;
    for (int x = 0; x < xmax; x++) {
        double t = DFD_DMUL(DFD_DADD(DFD_DSUB((double)(x + xmin), center), 0.5), ss);
        if (t < 0.0) t = -t;
        double v = t < 1.0 ? DFD_DSUB(1.0, t) : 0.0;
        w[x] = v;
        ww = DFD_DADD(ww, v);
    }
    for (int x = 0; x < xmax; x++) {
        double v = w[x];
        if (ww != 0.0) v = DFD_DDIV(v, ww);
        double sc = DFD_DMUL(v, (double)(1 << DFD_PIL_PRECISION));
        k[x] = v < 0.0 ? (int)DFD_DADD(-0.5, sc) : (int)DFD_DADD(0.5, sc);
    }
    *xmin_out = xmin;
    return xmax;
}

DFD_HD int dfd_pil_clip8(int acc) { return dfd_sat_u8(acc >> DFD_PIL_PRECISION); }

// ---- torch upsample_bilinear2d source index (align_corners=False) -----------
DFD_HD void dfd_torch_bilinear_coef(int d, int in_size, int out_size, int* i0, int* i1, float* l0, float* l1) {
    float scale = (float)in_size / (float)out_size;
    // torch's vectorised CPU kernel contracts scale*(d+0.5)-0.5 into one FMA (checked against
    // F.interpolate in this image: the unfused form is off by up to 2e-3 on a 0..255 scale).
    float src = DFD_FFMA(scale, DFD_FADD((float)d, 0.5f), -0.5f);
    if (src < 0.f) src = 0.f;
    int i = (int)src;
    if (i > in_size - 1) i = in_size - 1;
    *i0 = i; *i1 = i + (i < in_size - 1 ? 1 : 0);
    float lam = DFD_FSUB(src, (float)i);
    if (lam < 0.f) lam = 0.f; if (lam > 1.f) lam = 1.f;
    *l1 = lam; *l0 = DFD_FSUB(1.f, lam);
}
