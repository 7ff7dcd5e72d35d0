// Squeeze-excite excite stage fused into the tail of the depthwise kernels (bf16 path).
//
// lukemelas MBConvBlock: x_se = avg_pool(x); x_se = expand(swish(reduce(x_se))); x = sigmoid(x_se) * x.
// The depthwise kernels leave one squeeze partial per (image, tile, channel).  The LAST CTA of an image to finish
// (device-wide atomic counter) sums the partials in a fixed order -- so the result does not depend on which CTA is
// last -- and runs the two tiny FCs for that image, writing the gates the project GEMM applies to its A operand.
// This removes two launches per block (32 per forward pass) whose cost was launch latency, not work.
#pragma once
#include <cuda_runtime.h>

struct SeTail {
    const float* Wr;       // [se][C]   reduce weights
    const float* br;       // [se]
    const float* WxT;      // [se][C]   expand weights, transposed
    const float* bx;       // [C]
    float* scale;          // [m][C]    gates out
    int* counter;          // [m]       CTAs finished per image (self-resetting)
    int se;
    int ctas_per_image;
    float inv_hw;
};

// Call with ALL threads of the CTA after this CTA's partials for image b were written to pool[(b*n_parts+q)*C+c].
// smem: >= C + 64 floats of shared memory no longer in use.  Returns after the gates are written (last CTA only).
__device__ __forceinline__ void se_tail_run(const SeTail& t, const float* pool, int n_parts, int C, int b, float* smem) {
    __shared__ int s_last;
    const int tid = threadIdx.x, nthreads = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
    // Release/acquire on the counter by ONE thread: the CTA barrier orders every thread's partial-sum stores before
    // thread 0's release (PTX memory model: causality order is cumulative), and the last CTA's acquire + barrier orders
    // them before its reads.  No per-thread __threadfence(), so finishing CTAs do not wait for their output stores.
    __syncthreads();
    if (tid == 0) {
        int prev;
        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(prev) : "l"(t.counter + b) : "memory");
        s_last = prev == t.ctas_per_image - 1;
    }
    __syncthreads();
    if (!s_last) return;
    float* mean = smem;
    float* r = smem + C;
    for (int c = tid; c < C; c += nthreads) {
        float a = 0.f;
        for (int q = 0; q < n_parts; q++) a += __ldcg(pool + ((size_t)b * n_parts + q) * C + c);
        mean[c] = a * t.inv_hw;
    }
    __syncthreads();
    for (int j = warp; j < t.se; j += nwarps) {
        const float4* w4 = (const float4*)(t.Wr + (size_t)j * C);
        float a = 0.f;
        for (int c4 = lane; c4 < (C >> 2); c4 += 32) {
            const float4 w = __ldg(w4 + c4);
            const float4 x = *(const float4*)(mean + c4 * 4);
            a = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, a))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) { a += t.br[j]; r[j] = a / (1.0f + expf(-a)); }
    }
    __syncthreads();
    for (int c = tid; c < C; c += nthreads) {
        float a = t.bx[c];
        for (int j = 0; j < t.se; j++) a = fmaf(__ldg(t.WxT + (size_t)j * C + c), r[j], a);
        t.scale[(size_t)b * C + c] = 1.0f / (1.0f + expf(-a));
    }
    if (tid == 0) t.counter[b] = 0;                    // ready for the next launch (stream-ordered)
}
