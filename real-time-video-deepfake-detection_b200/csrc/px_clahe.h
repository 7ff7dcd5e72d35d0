// Bit-exact restatement of cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8,8)).apply(L)
// as the reference calls it on the LAB L channel of the native-resolution face
// crop (deepfake_detection.py:363-368).  SURVEY.md Appendix B.6.
#pragma once
#include "px_common.h"

#define DFD_CLAHE_TILES 8

struct DfdClaheGeom {
    int w, h;        // crop size
    int ew, eh;      // extended (padded) size, multiples of 8
    int tw, th;      // tile size
    int clip;        // integer clip limit
    float lut_scale; // 255 / tile_area
};

DFD_HD DfdClaheGeom dfd_clahe_geom(int w, int h) {
    DfdClaheGeom g;
    g.w = w; g.h = h;
    if (w % DFD_CLAHE_TILES == 0 && h % DFD_CLAHE_TILES == 0) { g.ew = w; g.eh = h; }
    else {   // OpenCV pads BOTH axes whenever either is not divisible (a divisible axis grows by 8)
        g.ew = w + (DFD_CLAHE_TILES - (w % DFD_CLAHE_TILES));
        g.eh = h + (DFD_CLAHE_TILES - (h % DFD_CLAHE_TILES));
    }
    g.tw = g.ew / DFD_CLAHE_TILES; g.th = g.eh / DFD_CLAHE_TILES;
    int area = g.tw * g.th;
    int clip = (int)(2.0 * (double)area / 256.0);
    g.clip = clip < 1 ? 1 : clip;
    g.lut_scale = 255.0f / (float)area;
    return g;
}

// hist[256] -> lut[256]; sequential (one thread per tile).
DFD_HD void dfd_clahe_lut(int* hist, int clip, float lut_scale, uint8_t* lut) {
    int clipped = 0;
    for (int i = 0; i < 256; i++)
        if (hist[i] > clip) { clipped += hist[i] - clip; hist[i] = clip; }
    int batch = clipped / 256;
    int residual = clipped - batch * 256;
    for (int i = 0; i < 256; i++) hist[i] += batch;
    if (residual != 0) {
        int step = 256 / residual; if (step < 1) step = 1;
        for (int i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++;
    }
    int sum = 0;
    for (int i = 0; i < 256; i++) {
        sum += hist[i];
        lut[i] = (uint8_t)dfd_sat_u8(DFD_RINTF(DFD_FMUL((float)sum, lut_scale)));
    }
}

#if defined(__CUDACC__)
// The same LUT built by one WARP (lane = 8 consecutive bins): clip + excess by warp reduction, the residual
// "every step-th bin" increment in closed form, the cumulative sum by a shuffle scan.  hist is in shared memory.
__device__ __forceinline__ void dfd_clahe_lut_warp(const int* hist, int clip, float lut_scale, uint8_t* lut, int lane) {
    int h[8];
    int excess = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int v = hist[lane * 8 + j];
        if (v > clip) { excess += v - clip; v = clip; }
        h[j] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) excess += __shfl_xor_sync(0xffffffffu, excess, o);
    const int batch = excess / 256;
    const int residual = excess - batch * 256;
    int step = residual ? 256 / residual : 1;
    if (step < 1) step = 1;
    int run = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int i = lane * 8 + j;
        h[j] += batch;
        if (residual != 0 && i % step == 0 && i / step < residual) h[j]++;     // bins 0, step, 2*step, ... (residual of them)
        run += h[j];
        h[j] = run;                                                             // inclusive prefix inside the lane
    }
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int base = incl - run;                                                // exclusive prefix of the lanes before
    uint32_t w0 = 0, w1 = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t v = (uint32_t)dfd_sat_u8(DFD_RINTF(DFD_FMUL((float)(base + h[j]), lut_scale)));
        if (j < 4) w0 |= v << (8 * j); else w1 |= v << (8 * (j - 4));
    }
    *(uint2*)(lut + lane * 8) = make_uint2(w0, w1);
}
#endif

// luts: [8][8][256] for this crop.  (x,y) in the ORIGINAL crop.
DFD_HD int dfd_clahe_apply(const uint8_t* luts, const DfdClaheGeom& g, int x, int y, int val) {
    float inv_tw = 1.0f / (float)g.tw, inv_th = 1.0f / (float)g.th;
    float txf = DFD_FSUB(DFD_FMUL((float)x, inv_tw), 0.5f);
    int tx1 = (int)floorf(txf); int tx2 = tx1 + 1;
    float xa = DFD_FSUB(txf, (float)tx1), xa1 = DFD_FSUB(1.0f, xa);
    if (tx1 < 0) tx1 = 0;
    if (tx2 > DFD_CLAHE_TILES - 1) tx2 = DFD_CLAHE_TILES - 1;
    float tyf = DFD_FSUB(DFD_FMUL((float)y, inv_th), 0.5f);
    int ty1 = (int)floorf(tyf); int ty2 = ty1 + 1;
    float ya = DFD_FSUB(tyf, (float)ty1), ya1 = DFD_FSUB(1.0f, ya);
    if (ty1 < 0) ty1 = 0;
    if (ty2 > DFD_CLAHE_TILES - 1) ty2 = DFD_CLAHE_TILES - 1;
    float l11 = (float)luts[(ty1 * 8 + tx1) * 256 + val], l12 = (float)luts[(ty1 * 8 + tx2) * 256 + val];
    float l21 = (float)luts[(ty2 * 8 + tx1) * 256 + val], l22 = (float)luts[(ty2 * 8 + tx2) * 256 + val];
    float top = DFD_FADD(DFD_FMUL(l11, xa1), DFD_FMUL(l12, xa));
    float bot = DFD_FADD(DFD_FMUL(l21, xa1), DFD_FMUL(l22, xa));
    float res = DFD_FADD(DFD_FMUL(top, ya1), DFD_FMUL(bot, ya));
    return dfd_sat_u8(DFD_RINTF(res));
}
