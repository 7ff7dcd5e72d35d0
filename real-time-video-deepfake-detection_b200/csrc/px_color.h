// Bit-exact restatement of the OpenCV 8-bit colour conversions the reference
// calls on the hot path (opencv-python, imgproc/color_{rgb,hsv,lab}):
//   cv2.COLOR_BGR2GRAY  frame_analysis.py:136,188,243,285,356
//   cv2.COLOR_BGR2HSV   frame_analysis.py:318
//   cv2.COLOR_BGR2LAB / COLOR_LAB2BGR   deepfake_detection.py:363,368
// Recipes: SURVEY.md Appendix B.2-B.5.  Checked exhaustively (2^24 inputs)
// against cv2 by tests/test_hostcheck.py.
#pragma once
#include "px_common.h"

struct DfdColorTables {
    int32_t sdiv[256];        // HSV: rint((255<<12)/v)
    int32_t hdiv180[256];     // HSV: rint((180<<12)/(6 d))
    uint16_t gamma[256];      // Lab: rint(2040 * srgb_to_linear(i/255))
    uint16_t cbrt[3072];      // Lab: rint(32768 * f(i/2040))
    uint16_t lab_y[256];      // Lab->RGB: y(L)
    uint16_t lab_ify[256];    // Lab->RGB: fy(L)
    uint8_t inv_gamma[4096];  // Lab->RGB: rint(255 * linear_to_srgb(i/4095|4096))
};

// ---- BGR -> GRAY (B.2) -----------------------------------------------------
DFD_HD int dfd_bgr2gray(int b, int g, int r) {
    return (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15;
}

// ---- BGR -> HSV, H in [0,180) (B.3) ----------------------------------------
DFD_HD void dfd_bgr2hsv(const DfdColorTables* T, int b, int g, int r, int* H, int* S, int* V) {
    int v = b > g ? b : g; v = v > r ? v : r;
    int mn = b < g ? b : g; mn = mn < r ? mn : r;
    int d = v - mn;
    int s = (d * T->sdiv[v] + 2048) >> 12;
    int h;
    if (v == r) h = g - b;
    else if (v == g) h = b - r + 2 * d;
    else h = r - g + 4 * d;
    h = (h * T->hdiv180[d] + 2048) >> 12;
    if (h < 0) h += 180;
    *H = dfd_sat_u8(h); *S = s; *V = v;
}

// ---- BGR -> Lab (sRGB, D65) (B.4) -------------------------------------------
DFD_HD void dfd_bgr2lab(const DfdColorTables* T, int b, int g, int r, int* L, int* A, int* B) {
    int R = T->gamma[r], G = T->gamma[g], Bl = T->gamma[b];
    int fX = T->cbrt[(R * 1777 + G * 1541 + Bl * 778 + 2048) >> 12];
    int fY = T->cbrt[(R * 871 + G * 2929 + Bl * 296 + 2048) >> 12];
    int fZ = T->cbrt[(R * 73 + G * 448 + Bl * 3575 + 2048) >> 12];
    int l = (296 * fY - 1336934 + 16384) >> 15;
    int a = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
    int bb = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
    *L = dfd_sat_u8(l); *A = dfd_sat_u8(a); *B = dfd_sat_u8(bb);
}

// L channel alone (the CLAHE histogram pass needs nothing else): same arithmetic as dfd_bgr2lab, fY only.
DFD_HD int dfd_bgr2lab_L(const uint16_t* gamma, const uint16_t* cbrt, int b, int g, int r) {
    int R = gamma[r], G = gamma[g], Bl = gamma[b];
    int fY = cbrt[(R * 871 + G * 2929 + Bl * 296 + 2048) >> 12];
    return dfd_sat_u8((296 * fY - 1336934 + 16384) >> 15);
}

// ---- Lab -> BGR (Lab2RGBinteger) (B.5) --------------------------------------
DFD_HD int dfd_ab_to_xz(int i) {            // abToXZ_b[i - minABvalue]
    const int BASE = 16384;
    if (i <= 3390) return i * 108 / 841 - BASE * 16 / 116 * 108 / 841;   // C truncating division
    return i * i / BASE * i / BASE;
}

DFD_HD void dfd_lab2bgr(const DfdColorTables* T, int L, int a, int b, int* ob, int* og, int* orr) {
    const int BASE = 16384;
    int y = T->lab_y[L];
    int ify = T->lab_ify[L];
    int adiv = ((5 * a * 53687 + 128) >> 13) - 128 * BASE / 500;
    int bdiv = ((b * 41943 + 16) >> 9) - 128 * BASE / 200 + 1;
    int x = dfd_ab_to_xz(ify + adiv);
    int z = dfd_ab_to_xz(ify - bdiv);
    int ro = (12615 * x - 6296 * y - 2223 * z + 8192) >> 14;
    int go = (-3773 * x + 7684 * y + 185 * z + 8192) >> 14;
    int bo = (217 * x - 836 * y + 4715 * z + 8192) >> 14;
    ro = dfd_clampi(ro, 0, 4095); go = dfd_clampi(go, 0, 4095); bo = dfd_clampi(bo, 0, 4095);
    *orr = T->inv_gamma[ro]; *og = T->inv_gamma[go]; *ob = T->inv_gamma[bo];
}

// ---- host-side table construction ------------------------------------------
#include <cmath>
static inline int dfd_cvround_d(double v) { return (int)std::nearbyint(v); }   // FE_TONEAREST: half-to-even
static inline int dfd_cvround_f(float v) { return (int)std::nearbyintf(v); }

static inline float dfd_cv_cuberoot(float value) {
    // OpenCV's quartic-rational cube root (cv::cubeRoot / softfloat cbrt), error < 2^-24.
    union { float f; int32_t i; } v, m;
    v.f = value;
    int ix = v.i & 0x7fffffff, s = v.i & 0x80000000;
    int ex = (ix >> 23) - 127;
    int shx = ex % 3;
    shx -= shx >= 0 ? 3 : 0;
    ex = (ex - shx) / 3;
    v.i = (ix & ((1 << 23) - 1)) | ((shx + 127) << 23);
    double fr = v.f;
    fr = ((((45.2548339756803022511987494 * fr + 192.2798368355061050458134625) * fr +
            119.1654824285581628956914143) * fr + 13.43250139086239872172837314) * fr +
          0.1636161226585754240958355063) /
         ((((14.80884093219134573786480845 * fr + 151.9714051044435648658557668) * fr +
            168.5254414101568283957668343) * fr + 33.9905941350215598754191872) * fr + 1.0);
    m.f = value;
    v.f = (float)fr;
    v.i = (v.i + (ex << 23) + s) & (m.i * 2 != 0 ? -1 : 0);
    return v.f;
}

static inline void dfd_build_color_tables(DfdColorTables* T, int cbrt_mode = 0, int inv_div = 4096) {
    T->sdiv[0] = T->hdiv180[0] = 0;
    for (int i = 1; i < 256; i++) {
        T->sdiv[i] = dfd_cvround_d((255 << 12) / (1.0 * i));
        T->hdiv180[i] = dfd_cvround_d((180 << 12) / (6.0 * i));
    }
    for (int i = 0; i < 256; i++) {
        float x = (float)i / 255.0f;
        double xd = x;
        double gd = xd <= 809.0 / 20000.0 ? xd / (323.0 / 25.0)
                                          : std::pow((xd + 11.0 / 200.0) / (1.0 + 11.0 / 200.0), 12.0 / 5.0);
        float gf = (float)gd;
        T->gamma[i] = (uint16_t)dfd_cvround_f(2040.0f * gf);
    }
    const float lthresh = 216.0f / 24389.0f, lscale = 841.0f / 108.0f, lbias = 16.0f / 116.0f;
    for (int i = 0; i < 3072; i++) {
        float x = (float)i / 2040.0f;
        float f;
        if (x < lthresh) f = (float)std::fma((double)x, (double)lscale, (double)lbias);   // single rounding
        else f = cbrt_mode == 0 ? dfd_cv_cuberoot(x) : (float)std::cbrt((double)x);
        T->cbrt[i] = (uint16_t)dfd_cvround_f(32768.0f * f);
    }
    // OpenCV builds this table with its softfloat cbrt, whose result for x = 324/2040 is one ulp
    // below the correctly rounded value; 32768*cbrt(x) = 17745.4992 sits on a rounding tie in
    // float32, so that single entry differs.  Pinned by the exhaustive 2^24 check against cv2.
    T->cbrt[324] = 17745;
    const int BASE = 16384;
    for (int i = 0; i < 256; i++) {
        int y, ify;
        if (i <= 20) {
            y = dfd_cvround_f((float)(i * BASE * 20 * 9) / (float)(17 * 29 * 29 * 29));
            ify = dfd_cvround_f((float)BASE * (16.0f / 116.0f + (float)(i * 5) / (float)(3 * 17 * 29)));
        } else {
            float fy = (float)(i * 100 * BASE) / (float)(255 * 116) + (float)(16 * BASE) / 116.0f;
            ify = dfd_cvround_f(fy);
            y = dfd_cvround_f(fy * fy * fy / (float)(BASE * BASE));
        }
        T->lab_y[i] = (uint16_t)y;
        T->lab_ify[i] = (uint16_t)ify;
    }
    for (int i = 0; i < 4096; i++) {
        float x = (float)i / (float)inv_div;
        double xd = x;
        double gd = xd <= 7827.0 / 2500000.0 ? xd * (323.0 / 25.0)
                                             : std::pow(xd, 5.0 / 12.0) * (1.0 + 11.0 / 200.0) - 11.0 / 200.0;
        float gf = (float)gd;
        int v = dfd_cvround_f(255.0f * gf);
        T->inv_gamma[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
}
