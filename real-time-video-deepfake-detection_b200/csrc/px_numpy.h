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

// Python's builtin sum() over exact floats (CPython >= 3.12, Python/bltinmodule.c: Neumaier-compensated (hi, lo) pair, the
// correction added once at the end) of the products s[k] * w[k] -- the reference's combined forensic score,
// sum(scores[k] * weights[k] for k in weights), frame_analysis.py:94,119.  A plain running sum differs in the last bit for
// a third of the reachable score combinations and flips the strict `p > 0.5` vote for a few hundred of them.
DFD_HD double dfd_py_sum_products(const double* s, const double* w, int n) {
    double hi = 0.0, lo = 0.0;
    for (int k = 0; k < n; k++) {
        const double x = DFD_DMUL(s[k], w[k]);                                 // no FMA: Python rounds the product
        const double t = DFD_DADD(hi, x);
        if (fabs(hi) >= fabs(x)) lo = DFD_DADD(lo, DFD_DADD(DFD_DSUB(hi, t), x));
        else lo = DFD_DADD(lo, DFD_DADD(DFD_DSUB(x, t), hi));
        hi = t;
    }
    if (lo != 0.0 && isfinite(lo)) hi = DFD_DADD(hi, lo);
    return hi;
}
