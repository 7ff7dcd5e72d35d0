// Bit-exact restatement of cv2.Canny(gray, 50, 150) (aperture 3, L1 gradient)
// and cv2.Laplacian(gray, CV_64F) as the reference's edge signal calls them
// (frame_analysis.py:285-293).  SURVEY.md Appendix B.8.
#pragma once
#include "px_common.h"

#define DFD_CANNY_LOW 50
#define DFD_CANNY_HIGH 150

// Sobel 3x3 with BORDER_REPLICATE on a w x h u8 image.
DFD_HD void dfd_sobel3(const uint8_t* g, int w, int h, int x, int y, int* dx, int* dy) {
    int xm = x > 0 ? x - 1 : 0, xp = x < w - 1 ? x + 1 : w - 1;
    int ym = y > 0 ? y - 1 : 0, yp = y < h - 1 ? y + 1 : h - 1;
    int a = g[ym * w + xm], b = g[ym * w + x], c = g[ym * w + xp];
    int d = g[y * w + xm], f = g[y * w + xp];
    int p = g[yp * w + xm], q = g[yp * w + x], r = g[yp * w + xp];
    *dx = (c + 2 * f + r) - (a + 2 * d + p);
    *dy = (p + 2 * q + r) - (a + 2 * b + c);
}

// Non-maximum suppression.  mag(xx,yy) must return 0 outside the image.
// Returns 0 = not an edge, 1 = weak candidate (> low), 2 = strong (> high).
template <class MagFn>
DFD_HD int dfd_canny_nms(int dxv, int dyv, int m, int x, int y, MagFn mag) {
    if (m <= DFD_CANNY_LOW) return 0;
    const int TG22 = 13573;
    int ax = dfd_absi(dxv), ay = dfd_absi(dyv) << 15;
    int tg22x = ax * TG22;
    bool keep;
    if (ay < tg22x) {
        keep = m > mag(x - 1, y) && m >= mag(x + 1, y);
    } else {
        int tg67x = tg22x + (ax << 16);
        if (ay > tg67x) {
            keep = m > mag(x, y - 1) && m >= mag(x, y + 1);
        } else {
            int s = (dxv ^ dyv) < 0 ? -1 : 1;
            keep = m > mag(x - s, y - 1) && m > mag(x + s, y + 1);
        }
    }
    if (!keep) return 0;
    return m > DFD_CANNY_HIGH ? 2 : 1;
}

// Packed form of the same rule for kernels that compute every magnitude once: code = magnitude (<= 2040, 11 bits)
// | direction << 11, direction 0 = horizontal, 1 = vertical, 2 = diagonal with s = +1, 3 = diagonal with s = -1.
DFD_HD unsigned dfd_canny_pack(int dxv, int dyv) {
    const int TG22 = 13573;
    int ax = dfd_absi(dxv), ayv = dfd_absi(dyv), ay = ayv << 15;
    int tg22x = ax * TG22;
    unsigned dir;
    if (ay < tg22x) dir = 0;
    else if (ay > tg22x + (ax << 16)) dir = 1;
    else dir = (dxv ^ dyv) < 0 ? 3 : 2;
    return (unsigned)(ax + ayv) | (dir << 11);
}
// offset of the FIRST neighbour of direction `dir` in a magnitude plane of the given pitch (the second is its negation)
DFD_HD int dfd_canny_first_off(unsigned dir, int pitch) {
    return dir == 0 ? -1 : dir == 1 ? -pitch : dir == 2 ? -pitch - 1 : -pitch + 1;
}
// c = this pixel's code, a / b = magnitudes of the first / second neighbour (0 outside the image)
DFD_HD int dfd_canny_nms_packed(unsigned c, int a, int b) {
    const int m = (int)(c & 2047u);
    if (m <= DFD_CANNY_LOW) return 0;
    const bool keep = m > a && ((c >> 11) < 2 ? m >= b : m > b);
    if (!keep) return 0;
    return m > DFD_CANNY_HIGH ? 2 : 1;
}

// Laplacian ksize=1: [0 1 0; 1 -4 1; 0 1 0], BORDER_REFLECT_101.
DFD_HD int dfd_laplacian(const uint8_t* g, int w, int h, int x, int y) {
    int xm = dfd_reflect101(x - 1, w), xp = dfd_reflect101(x + 1, w);
    int ym = dfd_reflect101(y - 1, h), yp = dfd_reflect101(y + 1, h);
    return g[ym * w + x] + g[yp * w + x] + g[y * w + xm] + g[y * w + xp] - 4 * g[y * w + x];
}

// GaussianBlur 5x5 sigma 0 = [1 4 6 4 1]/16 separable, REFLECT_101: returns 256*blur (exact integer).
DFD_HD int dfd_gauss5_x256(const uint8_t* g, int w, int h, int x, int y) {
    const int k[5] = {1, 4, 6, 4, 1};
    int acc = 0;
    for (int j = -2; j <= 2; j++) {
        int yy = dfd_reflect101(y + j, h);
        int row = 0;
        for (int i = -2; i <= 2; i++) row += k[i + 2] * g[yy * w + dfd_reflect101(x + i, w)];
        acc += k[j + 2] * row;
    }
    return acc;
}
