// float32 reductions in NumPy's order, for the SMALL arrays whose statistics the
// reference thresholds (64 block statistics, <=30 temporal differences):
// np.mean / np.std on a contiguous float32 array use pairwise summation with
// an 8-way unrolled inner loop for n <= 128 (numpy/_core/src/umath/loops_utils.h.src)
// and do every step in float32.
#pragma once
#include "px_common.h"

#if defined(__CUDA_ARCH__)
#define DFD_FDIV(a, b) __fdiv_rn((a), (b))
#define DFD_FSQRT(a) __fsqrt_rn((a))
#else
#define DFD_FDIV(a, b) ((float)(a) / (float)(b))
#define DFD_FSQRT(a) sqrtf((a))
#endif

DFD_HD float dfd_np_sum_f32(const float* a, int n) {      // n <= 128
    if (n < 8) {
        float res = 0.f;                                  // numpy starts from -0.0; identical for sums
        for (int i = 0; i < n; i++) res = DFD_FADD(res, a[i]);
        return res;
    }
    float r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] = DFD_FADD(r[j], a[i + j]);
    float res = DFD_FADD(DFD_FADD(DFD_FADD(r[0], r[1]), DFD_FADD(r[2], r[3])),
                         DFD_FADD(DFD_FADD(r[4], r[5]), DFD_FADD(r[6], r[7])));
    for (; i < n; i++) res = DFD_FADD(res, a[i]);
    return res;
}

DFD_HD float dfd_np_mean_f32(const float* a, int n) { return DFD_FDIV(dfd_np_sum_f32(a, n), (float)n); }

// np.std (population): sqrt(sum((a-mean)^2)/n), tmp must hold n floats.
DFD_HD float dfd_np_std_f32(const float* a, int n, float* tmp) {
    float mean = dfd_np_mean_f32(a, n);
    for (int i = 0; i < n; i++) { float d = DFD_FSUB(a[i], mean); tmp[i] = DFD_FMUL(d, d); }
    return DFD_FSQRT(DFD_FDIV(dfd_np_sum_f32(tmp, n), (float)n));
}
