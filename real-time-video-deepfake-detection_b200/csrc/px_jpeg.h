// Bit-exact restatement of the JPEG quality-90 round trip the reference's ELA
// signal performs with cv2.imencode('.jpg', Q=90) -> cv2.imdecode
// (frame_analysis.py:234-236): libjpeg-turbo baseline, 4:2:0, integer "islow"
// DCTs, fancy chroma up-sampling.  Entropy coding is lossless and skipped.
// SURVEY.md Appendix B.7 / D.
#pragma once
#include "px_common.h"

#define DFD_F0_298 2446
#define DFD_F0_390 3196
#define DFD_F0_541 4433
#define DFD_F0_765 6270
#define DFD_F0_899 7373
#define DFD_F1_175 9633
#define DFD_F1_501 12299
#define DFD_F1_847 15137
#define DFD_F1_961 16069
#define DFD_F2_053 16819
#define DFD_F2_562 20995
#define DFD_F3_072 25172

DFD_HD int dfd_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// RGB -> YCbCr (jccolor.c, 16-bit fixed point)
DFD_HD int dfd_jpeg_y(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
DFD_HD int dfd_jpeg_cb(int r, int g, int b) { return (-11059 * r - 21709 * g + 32768 * b + 8388608 + 32767) >> 16; }
DFD_HD int dfd_jpeg_cr(int r, int g, int b) { return (32768 * r - 27439 * g - 5329 * b + 8388608 + 32767) >> 16; }

// YCbCr -> RGB (jdcolor.c)
DFD_HD void dfd_jpeg_ycc2rgb(int y, int cb, int cr, int* r, int* g, int* b) {
    cb -= 128; cr -= 128;
    *r = dfd_sat_u8(y + ((91881 * cr + 32768) >> 16));
    *b = dfd_sat_u8(y + ((116130 * cb + 32768) >> 16));
    *g = dfd_sat_u8(y + ((-22554 * cb - 46802 * cr + 32768) >> 16));
}

// Annex-K base tables scaled for quality 90 (scale factor 20), natural order.
DFD_HD int dfd_jpeg_q90(int chroma, int idx) {
    const unsigned char luma[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
                                    14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                                    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                                    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    const unsigned char chr[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                   24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                   99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                   99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
    int base = chroma ? chr[idx] : luma[idx];
    int q = (base * 20 + 50) / 100;
    return q < 1 ? 1 : (q > 255 ? 255 : q);
}

// one 8-point forward pass (jfdctint.c).  d has stride `st`.
DFD_HD void dfd_fdct8(int* d, int st, int first) {
    int d0 = d[0], d1 = d[st], d2 = d[2 * st], d3 = d[3 * st], d4 = d[4 * st], d5 = d[5 * st], d6 = d[6 * st], d7 = d[7 * st];
    int t0 = d0 + d7, t7 = d0 - d7, t1 = d1 + d6, t6 = d1 - d6, t2 = d2 + d5, t5 = d2 - d5, t3 = d3 + d4, t4 = d3 - d4;
    int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    int sh = first ? 11 : 15;
    d[0] = first ? (t10 + t11) << 2 : dfd_descale(t10 + t11, 2);
    d[4 * st] = first ? (t10 - t11) << 2 : dfd_descale(t10 - t11, 2);
    int z1 = (t12 + t13) * DFD_F0_541;
    d[2 * st] = dfd_descale(z1 + t13 * DFD_F0_765, sh);
    d[6 * st] = dfd_descale(z1 - t12 * DFD_F1_847, sh);
    z1 = t4 + t7; int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7, z5 = (z3 + z4) * DFD_F1_175;
    t4 *= DFD_F0_298; t5 *= DFD_F2_053; t6 *= DFD_F3_072; t7 *= DFD_F1_501;
    z1 *= -DFD_F0_899; z2 *= -DFD_F2_562; z3 = z3 * -DFD_F1_961 + z5; z4 = z4 * -DFD_F0_390 + z5;
    d[7 * st] = dfd_descale(t4 + z1 + z3, sh);
    d[5 * st] = dfd_descale(t5 + z2 + z4, sh);
    d[3 * st] = dfd_descale(t6 + z2 + z3, sh);
    d[1 * st] = dfd_descale(t7 + z1 + z4, sh);
}

// one 8-point inverse pass (jidctint.c).
DFD_HD void dfd_idct8(int* c, int st, int first) {
    int c0 = c[0], c1 = c[st], c2 = c[2 * st], c3 = c[3 * st], c4 = c[4 * st], c5 = c[5 * st], c6 = c[6 * st], c7 = c[7 * st];
    int z1 = (c2 + c6) * DFD_F0_541;
    int t2 = z1 - c6 * DFD_F1_847, t3 = z1 + c2 * DFD_F0_765;
    int t0 = (c0 + c4) << 13, t1 = (c0 - c4) << 13;
    int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    int a0 = c7, a1 = c5, a2 = c3, a3 = c1;
    z1 = a0 + a3; int z2 = a1 + a2, z3 = a0 + a2, z4 = a1 + a3, z5 = (z3 + z4) * DFD_F1_175;
    a0 *= DFD_F0_298; a1 *= DFD_F2_053; a2 *= DFD_F3_072; a3 *= DFD_F1_501;
    z1 *= -DFD_F0_899; z2 *= -DFD_F2_562; z3 = z3 * -DFD_F1_961 + z5; z4 = z4 * -DFD_F0_390 + z5;
    a0 += z1 + z3; a1 += z2 + z4; a2 += z2 + z3; a3 += z1 + z4;
    int sh = first ? 11 : 18;
    c[0] = dfd_descale(t10 + a3, sh); c[7 * st] = dfd_descale(t10 - a3, sh);
    c[st] = dfd_descale(t11 + a2, sh); c[6 * st] = dfd_descale(t11 - a2, sh);
    c[2 * st] = dfd_descale(t12 + a1, sh); c[5 * st] = dfd_descale(t12 - a1, sh);
    c[3 * st] = dfd_descale(t13 + a0, sh); c[4 * st] = dfd_descale(t13 - a0, sh);
}

// Full round trip of one 8x8 block held as ints (samples 0..255 in, 0..255 out).  Fully unrolled with the component as
// a template parameter: the block stays in registers and every quantiser is a compile-time constant, so the 64
// divisions become multiply-shifts.
#if defined(__CUDACC__)
#define DFD_UNROLL _Pragma("unroll")
#else
#define DFD_UNROLL
#endif
template <int CHROMA>
DFD_HD void dfd_jpeg_block_roundtrip_t(int* blk) {
    DFD_UNROLL
    for (int i = 0; i < 64; i++) blk[i] -= 128;
    DFD_UNROLL
    for (int r = 0; r < 8; r++) dfd_fdct8(blk + 8 * r, 1, 1);
    DFD_UNROLL
    for (int c = 0; c < 8; c++) dfd_fdct8(blk + c, 8, 0);
    DFD_UNROLL
    for (int i = 0; i < 64; i++) {
        const int q = dfd_jpeg_q90(CHROMA, i);
        int v = blk[i];
        int a = v < 0 ? -v : v;
        a = (a + 4 * q) / (8 * q);
        blk[i] = (v < 0 ? -a : a) * q;
    }
    DFD_UNROLL
    for (int c = 0; c < 8; c++) dfd_idct8(blk + c, 8, 1);
    DFD_UNROLL
    for (int r = 0; r < 8; r++) dfd_idct8(blk + 8 * r, 1, 0);
    DFD_UNROLL
    for (int i = 0; i < 64; i++) blk[i] = dfd_sat_u8(blk[i] + 128);
}
DFD_HD void dfd_jpeg_block_roundtrip(int* blk, int chroma) {
    if (chroma) dfd_jpeg_block_roundtrip_t<1>(blk); else dfd_jpeg_block_roundtrip_t<0>(blk);
}

// h2v2 fancy up-sampling of one chroma sample position (jdsample.c).
// plane: cw x ch chroma plane; returns the value at full-res (X, Y).
DFD_HD int dfd_jpeg_fancy_up(const uint8_t* plane, int cw, int ch, int X, int Y) {
    int cy = Y >> 1, cx = X >> 1;
    int ny = (Y & 1) ? cy + 1 : cy - 1;               // nearer neighbour row
    ny = dfd_clampi(ny, 0, ch - 1);
    const uint8_t* r0 = plane + cy * cw;
    const uint8_t* r1 = plane + ny * cw;
    int cur = 3 * r0[cx] + r1[cx];
    if (X & 1) {
        if (cx == cw - 1) return (cur * 4 + 7) >> 4;
        int nxt = 3 * r0[cx + 1] + r1[cx + 1];
        return (cur * 3 + nxt + 7) >> 4;
    } else {
        if (cx == 0) return (cur * 4 + 8) >> 4;
        int prv = 3 * r0[cx - 1] + r1[cx - 1];
        return (cur * 3 + prv + 8) >> 4;
    }
}
