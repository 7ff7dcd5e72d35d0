// Shared helpers for the bit-exact pixel math (host + device).
//
// Every integer pipeline on the hot path (OpenCV fixed-point resize / colour
// conversions / CLAHE / Canny, libjpeg-turbo round trip, Pillow resample) is
// written once as __host__ __device__ inline functions so that the SAME code
// that runs inside the CUDA kernels can be compiled for the CPU by
// tests/hostcheck (test infrastructure) and compared against cv2 / PIL without
// a GPU.  The product never runs these on the CPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define DFD_HD __host__ __device__ __forceinline__
#else
#define DFD_HD inline
#endif

// IEEE single ops that must not be contracted into FMAs (the reference's CPU
// libraries do not fuse them).
#if defined(__CUDA_ARCH__)
#define DFD_FMUL(a, b) __fmul_rn((a), (b))
#define DFD_FADD(a, b) __fadd_rn((a), (b))
#define DFD_FSUB(a, b) __fsub_rn((a), (b))
#define DFD_DMUL(a, b) __dmul_rn((a), (b))
#define DFD_DADD(a, b) __dadd_rn((a), (b))
#define DFD_DSUB(a, b) __dsub_rn((a), (b))
#define DFD_DDIV(a, b) __ddiv_rn((a), (b))
#define DFD_RINTF(x) __float2int_rn(x)
#define DFD_FFMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define DFD_FMUL(a, b) ((float)(a) * (float)(b))
#define DFD_FADD(a, b) ((float)(a) + (float)(b))
#define DFD_FSUB(a, b) ((float)(a) - (float)(b))
#define DFD_DMUL(a, b) ((double)(a) * (double)(b))
#define DFD_DADD(a, b) ((double)(a) + (double)(b))
#define DFD_DSUB(a, b) ((double)(a) - (double)(b))
#define DFD_DDIV(a, b) ((double)(a) / (double)(b))
#define DFD_RINTF(x) ((int)lrintf(x))
#define DFD_FFMA(a, b, c) fmaf((a), (b), (c))
#endif

DFD_HD int dfd_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
DFD_HD int dfd_sat_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
DFD_HD int dfd_absi(int v) { return v < 0 ? -v : v; }
DFD_HD int dfd_reflect101(int p, int n) {          // cv::BORDER_REFLECT_101
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}

// t / d for a compile-time constant d, correctly rounded, in three instructions: q = t * r, q' = fma(fma(-d, q, t), r, q)
// with r = RN(1 / d).  Proven equal to the IEEE quotient by exhaustion for the operand ranges of the face-prep
// normalisation (tests/hostcheck/divconst_check.c); NOT valid for denormal quotients.
DFD_HD float dfd_div_const(float t, float d, float r) {
    const float q = DFD_FMUL(t, r);
    return DFD_FFMA(DFD_FFMA(-d, q, t), r, q);
}
