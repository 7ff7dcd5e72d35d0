// Test-time augmentation pixel math (deepfake_detection.py:417-434), bit-exact with OpenCV 4.x:
//   cv2.flip(img, 1)                                   column mirror
//   cv2.convertScaleAbs(img, alpha=brightness, beta=0) saturate_cast<uchar>(|float(v) * float(alpha)|), round-half-even
//   cv2.warpAffine(img, M, (w, h))                     INTER_LINEAR, BORDER_CONSTANT 0: the fixed-point path of imgwarp.cpp
//       inverse matrix im (computed by the caller in double, as OpenCV does before the loops);
//       X0 = round((im[1]*y + im[2]) * 1024) + 16,  adelta(x) = round(im[0]*x * 1024)   (AB_BITS 10, round_delta 16)
//       X  = (X0 + adelta) >> 5:  source column X >> 5, fraction (X & 31) / 32;  rows likewise with im[3..5]
//       value = (sum of 4 taps * 15-bit weights + 2^14) >> 15; the 32 x 32 weight table of OpenCV is
//       round((1-fy)(1-fx) * 32768) etc. with fx, fy multiples of 1/32: exact integers (32-ax)(32-ay)*32, no fix-up needed
// Checked on the CPU against cv2 by tests/hostcheck (hc_tta_augment).
#pragma once
#include "px_common.h"

#if defined(__CUDA_ARCH__)
#define DFD_D2I_RN(x) __double2int_rn(x)
#else
#define DFD_D2I_RN(x) ((int)lrint(x))
#endif

DFD_HD int dfd_scale_abs_u8(int v, float alpha) {            // cv2.convertScaleAbs, beta = 0
    const float f = DFD_FMUL((float)v, alpha);
    return dfd_sat_u8(DFD_RINTF(f < 0.f ? -f : f));
}

// Source position of destination pixel (x, y): integer part (sx, sy) and 5-bit fractions (ax, ay).
DFD_HD void dfd_warp_src(const double* im, int x, int y, int* sx, int* sy, int* ax, int* ay) {
    const int X0 = DFD_D2I_RN(DFD_DMUL(DFD_DADD(DFD_DMUL(im[1], (double)y), im[2]), 1024.0)) + 16;
    const int Y0 = DFD_D2I_RN(DFD_DMUL(DFD_DADD(DFD_DMUL(im[4], (double)y), im[5]), 1024.0)) + 16;
    const int ad = DFD_D2I_RN(DFD_DMUL(DFD_DMUL(im[0], (double)x), 1024.0));
    const int bd = DFD_D2I_RN(DFD_DMUL(DFD_DMUL(im[3], (double)x), 1024.0));
    const int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
    *sx = dfd_clampi(X >> 5, -32768, 32767);                  // saturate_cast<short>
    *sy = dfd_clampi(Y >> 5, -32768, 32767);
    *ax = X & 31; *ay = Y & 31;
}

DFD_HD int dfd_warp_blend(int v00, int v01, int v10, int v11, int ax, int ay) {
    const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
    return dfd_sat_u8((v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15);
}

// One augmented pixel, three channels: taps are read from the un-augmented w x h image `src` (3 bytes per pixel, rows
// `pitch` bytes apart) through the flip and the brightness table `lut` (256 entries of dfd_scale_abs_u8).
template <typename LutT>
DFD_HD void dfd_tta_pixel(const uint8_t* src, int pitch, int w, int h, int flip, const LutT* lut, const double* im, int x, int y,
                          int* o0, int* o1, int* o2) {
    int sx, sy, ax, ay;
    dfd_warp_src(im, x, y, &sx, &sy, &ax, &ay);
    int v[4][3];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int xx = sx + (t & 1), yy = sy + (t >> 1);
        if ((unsigned)xx < (unsigned)w && (unsigned)yy < (unsigned)h) {
            const uint8_t* p = src + (size_t)yy * pitch + (size_t)(flip ? w - 1 - xx : xx) * 3;
            v[t][0] = lut[p[0]]; v[t][1] = lut[p[1]]; v[t][2] = lut[p[2]];
        } else {
            v[t][0] = v[t][1] = v[t][2] = 0;                  // borderValue
        }
    }
    *o0 = dfd_warp_blend(v[0][0], v[1][0], v[2][0], v[3][0], ax, ay);
    *o1 = dfd_warp_blend(v[0][1], v[1][1], v[2][1], v[3][1], ax, ay);
    *o2 = dfd_warp_blend(v[0][2], v[1][2], v[2][2], v[3][2], ax, ay);
}
