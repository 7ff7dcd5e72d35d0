// Test infrastructure (not product code): exhaustive proof that the 3-instruction division by a constant used by
// k_vpass_up_norm (q = t * r; q' = fma(fma(-d, q, t), r, q), r = RN(1/d)) equals the IEEE quotient t / d for EVERY float
// the kernel can feed it: all of [0, 256] for d = 255, and all of [-0.5, 0.6] with |t| >= 2^-100 or t == +0 for the three
// ImageNet std constants (differences of floats in [0, 1] and a mean >= 0.4 are never smaller than 2^-26 unless exactly +0).
// A second argument samples every n-th float (tests/test_hostcheck.py runs it with 97); the exit status is 1 on any mismatch.
// Build and run:  gcc -O2 -mfma -ffp-contract=off divconst_check.c -o divconst_check -lm && for i in 0 1 2 3; do ./divconst_check $i; done
// Result on 2026-10-18: 0 mismatches in all four runs (1.13e9 + 3 x 2.12e9 values; the only difference outside the stated
// domain is the sign of zero for t = -0 and results in the denormal range).
// exhaustive check: q' = fma(fma(-d, q, t), r, q) with q = t * r, r = RN(1/d)  ==  t / d  (IEEE) over a float range
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static uint32_t STRIDE = 1;
static uint64_t run(float d, float lo, float hi, const char* name) {
    const float r = 1.0f / d;
    uint64_t bad = 0, n = 0;
    // positives [max(lo,0), hi]
    if (hi >= 0) {
        uint32_t a = f2u(lo > 0 ? lo : 0.0f), b = f2u(hi);
        for (uint64_t u = a; u <= b; u += STRIDE) {
            volatile float t = u2f((uint32_t)u);
            float q = t * r; float rem = fmaf(-d, q, t); float q2 = fmaf(rem, r, q);
            float ref = t / d;
            if (f2u(q2) != f2u(ref) && (fabsf(t) >= 0x1p-100f || f2u(t) == 0u)) { if (bad < 5) printf("  mismatch %s t=%a q2=%a ref=%a\n", name, t, q2, ref); bad++; }
            n++;
        }
    }
    if (lo < 0) {
        uint32_t a = f2u(-0.0f), b = f2u(lo);
        for (uint64_t u = a; u <= b; u += STRIDE) {
            volatile float t = u2f((uint32_t)u);
            float q = t * r; float rem = fmaf(-d, q, t); float q2 = fmaf(rem, r, q);
            float ref = t / d;
            if (f2u(q2) != f2u(ref) && (fabsf(t) >= 0x1p-100f || f2u(t) == 0u)) { if (bad < 5) printf("  mismatch %s t=%a q2=%a ref=%a\n", name, t, q2, ref); bad++; }
            n++;
        }
    }
    printf("%s d=%a: %llu values, %llu mismatches\n", name, d, (unsigned long long)n, (unsigned long long)bad);
    return bad;
}
int main(int argc, char** argv) {
    int which = argc > 1 ? argv[1][0] - '0' : 0;
    if (argc > 2) STRIDE = (uint32_t)atoi(argv[2]);      // sampled run for the CPU test-suite (every STRIDE-th float)
    uint64_t bad = 0;
    if (which == 0) bad += run(255.0f, 0.0f, 256.0f, "div255");
    if (which == 1) bad += run(0.229f, -0.5f, 0.6f, "std0");
    if (which == 2) bad += run(0.224f, -0.5f, 0.6f, "std1");
    if (which == 3) bad += run(0.225f, -0.5f, 0.6f, "std2");
    return bad ? 1 : 0;
}
