// EfficientNet-B0 forward (reference model.py:63-72 -> lukemelas EfficientNet + custom _fc) on NHWC
// activations.  This translation unit holds the launch plan and the CUDA-core kernels:
//
//   k_stem      dense 3x3 s2 conv 3->32, TF-SAME pad (0,1), folded BN, swish
//   k_pw        1x1 conv as an fp32-FMA tiled GEMM with fused bias / swish / SE-scale / residual
//               (fp32 accuracy mode; bf16 mode uses the tcgen05 GEMM in gemm_tcgen05.cu instead)
//   k_dw        depthwise kxk stride s conv + folded BN + swish with the SE squeeze (per image,
//               per channel sums) fused in: register accumulation -> smem -> one atomic per channel
//   k_se        SE excite: mean -> reduce FC -> swish -> expand FC -> sigmoid  (one CTA per image)
//   k_scale     x * se (bf16 mode only; fp32 mode applies the scale while loading A in k_pw)
//   k_pool      global average pool of the head output
//   k_fc_layer  custom classifier 1280->512->256->1 (BN1d folded, ReLU): one launch per layer, 16 images x 32 outputs per CTA
//               at large batch (8 x 8 at small batch)
//
// Activations are float (fp32 mode: true fp32 FMA everywhere, parity 1e-4) or bf16 (storage only;
// all accumulation in fp32).
#include "dfd_internal.cuh"
#include "effnet_plan.h"
#include <string.h>
#include <type_traits>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

int dfd_gemm_bf16(dfd_ctx* ctx, const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias,
                  const __nv_bfloat16* residual, __nv_bfloat16* C, int M, int N, int K, int act, cudaStream_t st);
int dfd_gemm_bf16_ex(dfd_ctx* ctx, int a_mode, const __nv_bfloat16* A, const float* se, int hw, const __nv_bfloat16* W,
                     const float* bias, const __nv_bfloat16* residual, __nv_bfloat16* C, int M, int N, int K, int act,
                     cudaStream_t st);
bool dfd_gemm_bf16_enabled();
int dfd_gemm_bf16_img(dfd_ctx* ctx, const __nv_bfloat16* A, const __nv_bfloat16* Wg, const float* bias,
                      const __nv_bfloat16* residual, __nv_bfloat16* C, int n_img, int hw, int N, int K, int act, cudaStream_t st);
int dfd_dw_bf16(dfd_ctx* ctx, const EffBlock& b, const __nv_bfloat16* in, const float* W, const float* bias,
                __nv_bfloat16* out, int m, int* n_parts, cudaStream_t st);
int dfd_mbconv_front_bf16(dfd_ctx* ctx, int blk, const __nv_bfloat16* x, const __nv_bfloat16* We, __nv_bfloat16* out, int m,
                          int* n_parts, cudaStream_t st);
int dfd_front_pack(dfd_ctx* ctx, const float* blob);
int dfd_gemm_tf32x3(dfd_ctx* ctx, int a_mode, const float* A, const float* se, int hw, const float* W_hi, const float* W_lo,
                    const float* bias, const float* residual, float* C, int M, int N, int K, int act, cudaStream_t st);
void dfd_tf32_split_host(const float* w, size_t n, float* hi, float* lo);
int dfd_dw_f32(dfd_ctx* ctx, const EffBlock& b, const float* in, const float* W, const float* bias, float* out, int m,
               int* n_parts, cudaStream_t st, int img0);

template <typename T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float swishf(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_stem(const T* __restrict__ in, const float* __restrict__ W,
                                              const float* __restrict__ bias, T* __restrict__ out, int total) {
    __shared__ float sw[27 * 32];
    __shared__ float sb[32];
    for (int i = threadIdx.x; i < 27 * 32; i += 128) sw[i] = W[i];
    if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    int o = blockIdx.x * 128 + threadIdx.x;
    if (o >= total) return;
    int ox = o % 112, oy = (o / 112) % 112, b = o / (112 * 112);
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; c++) acc[c] = sb[c];
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
        int iy = 2 * oy + ky;
        if (iy >= 224) continue;
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            int ix = 2 * ox + kx;
            if (ix >= 224) continue;
            const T* p = in + (((size_t)b * 224 + iy) * 224 + ix) * 3;
#pragma unroll
            for (int ci = 0; ci < 3; ci++) {
                float v = ld1<T>(p + ci);
                const float* w = sw + ((ky * 3 + kx) * 3 + ci) * 32;
#pragma unroll
                for (int c = 0; c < 32; c++) acc[c] = fmaf(v, w[c], acc[c]);
            }
        }
    }
    T* q = out + (size_t)o * 32;
#pragma unroll
    for (int c = 0; c < 32; c++) st1<T>(q + c, swishf(acc[c]));
}

// ---------------------------------------------------------------------------------------------
// C[M,N] = act(A[M,K] (* se[img][k]) . W[N,K]^T + bias[n]) (+ residual[M,N]).  BM x 64 x 16 tiles, 4x4 outputs per thread.
// fp32 path (T = float): 16-byte global loads along K, the SE gate applied as the A tile is loaded (image index per row
// computed once, no division in the loop), and the next k-block's loads issued before the current block's FMAs.
// BM = 32 (128 threads) when the 64-row grid would not fill the chip (7x7 / 14x14 stages at small batch).
template <typename T, int BM>
__global__ void __launch_bounds__(BM * 4) k_pw(const T* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                                               const float* __restrict__ se, int hw, const T* __restrict__ residual,
                                               T* __restrict__ C, int M, int N, int K, int act) {
    constexpr int NT = BM * 4;                                // threads: (BM / 4) x 16
    __shared__ float sa[16][BM + 4];
    __shared__ float sb[16][64 + 4];
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    if constexpr (std::is_same<T, float>::value) {
        // loader mapping: one float4 (4 consecutive k) of one row per thread and tile
        const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;          // A: rows 0..BM-1
        const int am = m0 + lr;
        const bool a_ok = am < M;
        const float* a_row = A + (size_t)(a_ok ? am : 0) * K;
        const float* g_row = se ? se + (size_t)((a_ok ? am : 0) / hw) * K : nullptr;
        constexpr int BPT = 64 * 4 / NT;                                        // W float4 per thread: 1 (BM = 64) or 2 (BM = 32)
        const float* w_row[BPT];
        bool w_ok[BPT];
#pragma unroll
        for (int q = 0; q < BPT; q++) {
            const int n = n0 + lr + q * (NT / 4);
            w_ok[q] = n < N;
            w_row[q] = W + (size_t)(w_ok[q] ? n : 0) * K;
        }
        const bool k4 = (K & 3) == 0;
        auto load_a = [&](int k0) -> float4 {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = k0 + lk;
            if (a_ok && k < K) {
                if (k4) {
                    v = *(const float4*)(a_row + k);
                    if (g_row) { const float4 g = __ldg((const float4*)(g_row + k)); v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w; }
                } else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int j = 0; j < 4 && k + j < K; j++) t[j] = a_row[k + j] * (g_row ? g_row[k + j] : 1.0f);
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            return v;
        };
        auto load_b = [&](int q, int k0) -> float4 {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = k0 + lk;
            if (w_ok[q] && k < K) {
                if (k4) v = __ldg((const float4*)(w_row[q] + k));
                else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int j = 0; j < 4 && k + j < K; j++) t[j] = w_row[q][k + j];
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            return v;
        };
        float4 ra = load_a(0), rb[BPT];
#pragma unroll
        for (int q = 0; q < BPT; q++) rb[q] = load_b(q, 0);
        for (int k0 = 0; k0 < K; k0 += 16) {
            sa[lk][lr] = ra.x; sa[lk + 1][lr] = ra.y; sa[lk + 2][lr] = ra.z; sa[lk + 3][lr] = ra.w;
#pragma unroll
            for (int q = 0; q < BPT; q++) {
                const int r = lr + q * (NT / 4);
                sb[lk][r] = rb[q].x; sb[lk + 1][r] = rb[q].y; sb[lk + 2][r] = rb[q].z; sb[lk + 3][r] = rb[q].w;
            }
            __syncthreads();
            if (k0 + 16 < K) {                                                  // next block's loads fly during the FMAs
                ra = load_a(k0 + 16);
#pragma unroll
                for (int q = 0; q < BPT; q++) rb[q] = load_b(q, k0 + 16);
            }
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                const float4 a4 = *(const float4*)&sa[kk][ty * 4];
                const float4 b4 = *(const float4*)&sb[kk][tx * 4];
                const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bq[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
            }
            __syncthreads();
        }
    } else {
        for (int k0 = 0; k0 < K; k0 += 16) {
            for (int e = threadIdx.x; e < 64 * 16; e += NT) {
                int r = e >> 4, kk = e & 15;
                int m = m0 + r, k = k0 + kk;
                if (r < BM) {
                    float v = 0.f;
                    if (m < M && k < K) {
                        v = ld1<T>(A + (size_t)m * K + k);
                        if (se) v *= se[(size_t)(m / hw) * K + k];
                    }
                    sa[kk][r] = v;
                }
                int n = n0 + r;
                sb[kk][r] = (n < N && k < K) ? W[(size_t)n * K + k] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                float a[4], bq[4];
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = sa[kk][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; j++) bq[j] = sb[kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + bias[n];
            if (act) v = swishf(v);
            if (residual) v += ld1<T>(residual + (size_t)m * N + n);
            st1<T>(C + (size_t)m * N + n, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Depthwise conv.  Block = CG channel-groups x PS pixel lanes; each thread owns VEC channels and walks
// the pixels of its tile, so the SE squeeze accumulates in registers.
template <typename T, int VEC> struct VecIO;
template <> struct VecIO<float, 4> {
    static __device__ __forceinline__ void load(const float* p, float* v) { float4 t = *(const float4*)p; v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void store(float* p, const float* v) { *(float4*)p = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct VecIO<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
        uint4 t = *(const uint4*)p;
        const __nv_bfloat162* h = (const __nv_bfloat162*)&t;
#pragma unroll
        for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
        uint4 t;
        __nv_bfloat162* h = (__nv_bfloat162*)&t;
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *(uint4*)p = t;
    }
};

template <typename T, int VEC, int KS>
__global__ void __launch_bounds__(320) k_dw(const T* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                                            T* __restrict__ out, float* __restrict__ pool, int C, int hin, int hout,
                                            int stride, int pad, int pix_per_block) {
    extern __shared__ float spool[];                      // [PS][C]
    const int CG = C / VEC;
    const int PS = blockDim.x / CG;
    const int cg = threadIdx.x % CG, ps = threadIdx.x / CG;
    const int b = blockIdx.y;
    const int npix = hout * hout;
    const int p0 = blockIdx.x * pix_per_block;
    float psum[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) psum[v] = 0.f;
    if (ps < PS) {
        const int c0 = cg * VEC;
        float bv[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) bv[v] = bias[c0 + v];
        const int pend = min(p0 + pix_per_block, npix);
        for (int p = p0 + ps; p < pend; p += PS) {
            int oy = p / hout, ox = p % hout;
            float acc[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] = bv[v];
#pragma unroll
            for (int ky = 0; ky < KS; ky++) {
                int iy = oy * stride + ky - pad;
                if (iy < 0 || iy >= hin) continue;
#pragma unroll
                for (int kx = 0; kx < KS; kx++) {
                    int ix = ox * stride + kx - pad;
                    if (ix < 0 || ix >= hin) continue;
                    float xv[VEC], wv[VEC];
                    VecIO<T, VEC>::load(in + (((size_t)b * hin + iy) * hin + ix) * C + c0, xv);
                    const float* wp = W + (size_t)(ky * KS + kx) * C + c0;
#pragma unroll
                    for (int v = 0; v < VEC; v += 4) { float4 t = __ldg((const float4*)(wp + v)); wv[v] = t.x; wv[v + 1] = t.y; wv[v + 2] = t.z; wv[v + 3] = t.w; }
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[v] = fmaf(xv[v], wv[v], acc[v]);
                }
            }
#pragma unroll
            for (int v = 0; v < VEC; v++) { acc[v] = swishf(acc[v]); psum[v] += acc[v]; }
            VecIO<T, VEC>::store(out + ((size_t)b * npix + p) * C + c0, acc);
        }
#pragma unroll
        for (int v = 0; v < VEC; v++) spool[ps * C + c0 + v] = psum[v];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {      // deterministic: one partial per CTA, summed in order by k_se
        float s = 0.f;
        for (int q = 0; q < PS; q++) s += spool[q * C + c];
        pool[((size_t)b * gridDim.x + blockIdx.x) * C + c] = s;
    }
}

// SE excite as two small kernels with full-chip parallelism and coalesced, unrolled weight reads.
//   k_se_reduce : r[b][j] = swish(Wr[j] . mean[b] + br[j])      CTA = (8 squeeze channels, 8 images); warp = one j
//   k_se_expand : g[b][c] = sigmoid(WxT[:, c] . r[b] + bx[c])   CTA = (256 channels, 8 images); thread = one c
// The per-CTA squeeze partials of the depthwise kernel are summed here in a fixed order (deterministic).
#define SE_IPC 8
__global__ void __launch_bounds__(256) k_se_reduce(const float* __restrict__ pool, int n_parts, const float* __restrict__ Wr,
                                                   const float* __restrict__ br, float* __restrict__ rbuf, int C, int se,
                                                   float inv_hw, int m) {
    __shared__ __align__(16) float s[SE_IPC][1152];
    const int b0 = blockIdx.y * SE_IPC;
    for (int e = threadIdx.x; e < SE_IPC * C; e += 256) {
        const int i = e / C, c = e - i * C, b = b0 + i;
        float a = 0.f;
        if (b < m) {
#pragma unroll 4
            for (int q = 0; q < n_parts; q++) a += pool[((size_t)b * n_parts + q) * C + c];
        }
        s[i][c] = a * inv_hw;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (j >= se) return;
    float a[SE_IPC];
#pragma unroll
    for (int i = 0; i < SE_IPC; i++) a[i] = 0.f;
    const float4* w4 = (const float4*)(Wr + (size_t)j * C);
#pragma unroll 3
    for (int c4 = lane; c4 < (C >> 2); c4 += 32) {
        const float4 w = __ldg(w4 + c4);
#pragma unroll
        for (int i = 0; i < SE_IPC; i++) {
            const float4 x = *(const float4*)&s[i][c4 * 4];
            a[i] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, a[i]))));
        }
    }
#pragma unroll
    for (int i = 0; i < SE_IPC; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
        if (lane == 0 && b0 + i < m) rbuf[(size_t)(b0 + i) * 64 + j] = swishf(a[i] + br[j]);
    }
}

__global__ void __launch_bounds__(256) k_se_expand(const float* __restrict__ rbuf, const float* __restrict__ WxT,
                                                   const float* __restrict__ bx, float* __restrict__ scale, int C, int se, int m) {
    __shared__ float r[SE_IPC][64];
    const int b0 = blockIdx.y * SE_IPC;
    for (int e = threadIdx.x; e < SE_IPC * 64; e += 256) {
        const int i = e >> 6, j = e & 63;
        r[i][j] = (b0 + i < m && j < se) ? rbuf[(size_t)(b0 + i) * 64 + j] : 0.f;
    }
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    float a[SE_IPC];
    const float bc = bx[c];
#pragma unroll
    for (int i = 0; i < SE_IPC; i++) a[i] = bc;
#pragma unroll 4
    for (int j = 0; j < se; j++) {
        const float w = __ldg(WxT + (size_t)j * C + c);
#pragma unroll
        for (int i = 0; i < SE_IPC; i++) a[i] = fmaf(w, r[i][j], a[i]);
    }
#pragma unroll
    for (int i = 0; i < SE_IPC; i++)
        if (b0 + i < m) scale[(size_t)(b0 + i) * C + c] = sigmoidf(a[i]);
}

// SE excite in ONE launch (bf16 path): CTA = SE_X_IPC images; squeeze partials -> mean -> reduce FC -> swish -> expand FC
// -> sigmoid.  The two FCs are latency-bound chains of L2 weight reads, so every loop keeps 8-16 independent
// 16-byte / 4-byte loads in flight per lane and the weights are shared by the CTA's images.
#define SE_X_IPC 4
#define SE_X_THREADS 512
__global__ void __launch_bounds__(SE_X_THREADS) k_se_excite(const float* __restrict__ pool, int n_parts, const float* __restrict__ Wr,
                                                            const float* __restrict__ br, const float* __restrict__ WxT,
                                                            const float* __restrict__ bx, float* __restrict__ scale, int C, int se,
                                                            float inv_hw, int m) {
    __shared__ __align__(16) float mean[SE_X_IPC][1152];
    __shared__ float r[SE_X_IPC][64];
    const int b0 = blockIdx.x * SE_X_IPC;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < SE_X_IPC * C; e += SE_X_THREADS) {
        const int i = e / C, c = e - i * C, b = b0 + i;
        float a = 0.f;
        if (b < m) {
            const float* pp = pool + (size_t)b * n_parts * C + c;
            int q = 0;
            for (; q + 8 <= n_parts; q += 8) {                 // fixed order, 8 loads in flight
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = pp[(size_t)(q + u) * C];
#pragma unroll
                for (int u = 0; u < 8; u++) a += v[u];
            }
            for (; q < n_parts; q++) a += pp[(size_t)q * C];
        }
        mean[i][c] = a * inv_hw;
    }
    __syncthreads();
    // reduce FC: warp per squeeze channel j; a lane reads C/128 float4 of the weight row (<= 9), all issued up front
    for (int j = warp; j < se; j += SE_X_THREADS / 32) {
        const float4* w4 = (const float4*)(Wr + (size_t)j * C);
        const int n4 = C >> 2;
        float4 w[9];
#pragma unroll
        for (int u = 0; u < 9; u++) { const int c4 = lane + 32 * u; w[u] = c4 < n4 ? __ldg(w4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f); }
        float a[SE_X_IPC];
#pragma unroll
        for (int i = 0; i < SE_X_IPC; i++) a[i] = 0.f;
#pragma unroll
        for (int u = 0; u < 9; u++) {
            const int c4 = lane + 32 * u;
            if (c4 < n4) {
#pragma unroll
                for (int i = 0; i < SE_X_IPC; i++) {
                    const float4 x = *(const float4*)&mean[i][c4 * 4];
                    a[i] = fmaf(w[u].x, x.x, fmaf(w[u].y, x.y, fmaf(w[u].z, x.z, fmaf(w[u].w, x.w, a[i]))));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < SE_X_IPC; i++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
            if (lane == 0) r[i][j] = swishf(a[i] + br[j]);
        }
    }
    __syncthreads();
    // expand FC: a thread owns up to three channels (C <= 1152 < 3 * 512) and walks the squeeze channels 8 at a time
    // with all 24 weight loads issued before the FMAs
    {
        float a[3][SE_X_IPC];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int c = tid + SE_X_THREADS * k;
            const float bc = c < C ? bx[c] : 0.f;
#pragma unroll
            for (int i = 0; i < SE_X_IPC; i++) a[k][i] = bc;
        }
        for (int j0 = 0; j0 < se; j0 += 8) {
            float w[3][8];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int c = tid + SE_X_THREADS * k;
#pragma unroll
                for (int u = 0; u < 8; u++) w[k][u] = (c < C && j0 + u < se) ? __ldg(WxT + (size_t)(j0 + u) * C + c) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (j0 + u < se) {
#pragma unroll
                    for (int i = 0; i < SE_X_IPC; i++) {
                        const float rv = r[i][j0 + u];
#pragma unroll
                        for (int k = 0; k < 3; k++) a[k][i] = fmaf(w[k][u], rv, a[k][i]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int c = tid + SE_X_THREADS * k;
            if (c < C) {
#pragma unroll
                for (int i = 0; i < SE_X_IPC; i++)
                    if (b0 + i < m) scale[(size_t)(b0 + i) * C + c] = sigmoidf(a[k][i]);
            }
        }
    }
}

// SE excite on a thread-block CLUSTER (bf16 path, default): 8 CTAs share 8 images.  Each CTA owns 1/8 of the
// channels for the squeeze mean and the expand FC and 1/8 of the squeeze channels for the reduce FC, so a CTA reads only
// 1/8 of each weight matrix; the means and the reduced activations are exchanged through distributed shared memory
// (remote stores + cluster.sync).  One launch per block, latency of a few microseconds at batch 1 and 256 alike.
#define SE_CL 8
__global__ void __cluster_dims__(SE_CL, 1, 1) __launch_bounds__(256)
k_se_cluster(const float* __restrict__ pool, int n_parts, const float* __restrict__ Wr, const float* __restrict__ br,
             const float* __restrict__ WxT, const float* __restrict__ bx, float* __restrict__ scale, int C, int se,
             float inv_hw, int m, const float* __restrict__ Wp, __nv_bfloat16* __restrict__ Wg, int N, int fold) {
    __shared__ __align__(16) float mean[SE_CL][1152];
    __shared__ float r[SE_CL][64];
    __shared__ float gsm[SE_CL][144];                           // this CTA's gates (8 images x channel slice)
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / SE_CL) * SE_CL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Cs = C / SE_CL;                                   // every expanded width is a multiple of 8
    // The FC weights are static: this warp's reduce-FC row and this thread's first 16 expand-FC weights are fetched BEFORE the
    // programmatic-dependent-launch wait, i.e. while the producer of the squeeze partials is still running, which takes two of
    // the kernel's dependent L2 round trips off its critical path.
    float4 wr_pre[9];
    {
        const int j = rank + SE_CL * warp;
        const float4* w4 = (const float4*)(Wr + (size_t)(j < se ? j : 0) * C);
        const int n4 = C >> 2;
#pragma unroll
        for (int u = 0; u < 9; u++) { const int c4 = lane + 32 * u; wr_pre[u] = (j < se && c4 < n4) ? __ldg(w4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f); }
    }
    float wx_pre[16];
    float bx_pre = 0.f;
    if (tid < 2 * Cs) {
        const int half = tid / Cs, c = rank * Cs + (tid - half * Cs);
#pragma unroll
        for (int u = 0; u < 16; u++) wx_pre[u] = u < se ? __ldg(WxT + (size_t)u * C + c) : 0.f;
        bx_pre = bx[c];
    } else {
#pragma unroll
        for (int u = 0; u < 16; u++) wx_pre[u] = 0.f;
    }
    pdl_trigger();
    pdl_wait();
    float* mean_rem[SE_CL];
    float* r_rem[SE_CL];
#pragma unroll
    for (int d = 0; d < SE_CL; d++) {
        mean_rem[d] = cluster.map_shared_rank(&mean[0][0], d);
        r_rem[d] = cluster.map_shared_rank(&r[0][0], d);
    }
    // phase 1: squeeze means of this CTA's channel slice for the 8 images -> every CTA of the cluster
    for (int e = tid; e < SE_CL * Cs; e += 256) {
        const int i = e / Cs, c = rank * Cs + (e - i * Cs), b = b0 + i;
        float a = 0.f;
        if (b < m) {
            const float* pp = pool + (size_t)b * n_parts * C + c;
            int q = 0;
            for (; q + 8 <= n_parts; q += 8) {                 // fixed order, 8 loads in flight
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = pp[(size_t)(q + u) * C];
#pragma unroll
                for (int u = 0; u < 8; u++) a += v[u];
            }
            for (; q < n_parts; q++) a += pp[(size_t)q * C];
        }
        a *= inv_hw;
#pragma unroll
        for (int d = 0; d < SE_CL; d++) mean_rem[d][i * 1152 + c] = a;
    }
    cluster.sync();
    // phase 2: reduce FC for squeeze channel j = rank + 8 * warp, all 8 images
    {
        const int j = rank + SE_CL * warp;
        if (j < se) {
            const int n4 = C >> 2;
            const float4 (&w)[9] = wr_pre;
            float a[SE_CL];
#pragma unroll
            for (int i = 0; i < SE_CL; i++) a[i] = 0.f;
#pragma unroll
            for (int u = 0; u < 9; u++) {
                const int c4 = lane + 32 * u;
                if (c4 < n4) {
#pragma unroll
                    for (int i = 0; i < SE_CL; i++) {
                        const float4 x = *(const float4*)&mean[i][c4 * 4];
                        a[i] = fmaf(w[u].x, x.x, fmaf(w[u].y, x.y, fmaf(w[u].z, x.z, fmaf(w[u].w, x.w, a[i]))));
                    }
                }
            }
            const float bj = br[j];
#pragma unroll
            for (int i = 0; i < SE_CL; i++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
            }
            if (lane < SE_CL) {                                 // lane d publishes the 8 values to CTA d
#pragma unroll
                for (int i = 0; i < SE_CL; i++) r_rem[lane][i * 64 + j] = swishf(a[i] + bj);
            }
        }
    }
    cluster.sync();
    // phase 3: expand FC + sigmoid for this CTA's channel slice
    for (int e = tid; e < 2 * Cs; e += 256) {                   // two threads per channel: images 0-3 / 4-7
        const int half = e / Cs, c = rank * Cs + (e - half * Cs);
        float a[4];
        const float bc = e == tid ? bx_pre : bx[c];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = bc;
        for (int j0 = 0; j0 < se; j0 += 16) {
            float w[16];
            if (e == tid && j0 == 0) {
#pragma unroll
                for (int u = 0; u < 16; u++) w[u] = wx_pre[u];
            } else {
#pragma unroll
                for (int u = 0; u < 16; u++) w[u] = j0 + u < se ? __ldg(WxT + (size_t)(j0 + u) * C + c) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 16; u++) {
                if (j0 + u < se) {
#pragma unroll
                    for (int i = 0; i < 4; i++) a[i] = fmaf(w[u], r[half * 4 + i][j0 + u], a[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float g = sigmoidf(a[i]);
            gsm[half * 4 + i][c - rank * Cs] = g;
            if (b0 + half * 4 + i < m) scale[(size_t)(b0 + half * 4 + i) * C + c] = g;
        }
    }
    // phase 4 (high-resolution blocks): fold the gate into a per-image copy of the project weights,
    // Wg[b][n][k] = bf16(Wp[n][k] * g[b][k]), so the project GEMM needs no pass over its A operand (gemm A_IMG)
    if (Wg) {
        __syncthreads();
        // thread = one (output channel n, channel k of this CTA's slice): the weight is read once and gated for the 8 images
        // (no per-element division: n = e / Cs by multiplication with a precomputed reciprocal, valid for e < 2^16)
        const unsigned magic = 0xffffffffu / (unsigned)Cs + 1u;
        for (int e = tid; e < N * Cs; e += 256) {
            const int n = (int)__umulhi((unsigned)e, magic), cl = e - n * Cs;
            const int k = rank * Cs + cl;
            const float wv = __ldg(Wp + (size_t)n * C + k);
#pragma unroll
            for (int i = 0; i < SE_CL; i++) {
                if (b0 + i < m) {
                    const __nv_bfloat16 w = __float2bfloat16_rn(wv * gsm[i][cl]);
                    // fold > 1: `fold` consecutive pixels form one GEMM row (K' = fold * C, N' = fold * N) and the weight matrix is
                    // block-diagonal: copy f of W sits at rows f*N.., columns f*C.. (the off-diagonal blocks were zeroed once)
                    if (fold <= 1) Wg[((size_t)(b0 + i) * N + n) * C + k] = w;
                    else
                        for (int fd = 0; fd < fold; fd++)
                            Wg[((size_t)(b0 + i) * fold * N + fd * N + n) * (fold * C) + fd * C + k] = w;
                }
            }
        }
    }
}

// x *= se[img][c]   (bf16 mode: produces the A operand of the project GEMM)
__global__ void __launch_bounds__(256) k_scale(__nv_bfloat16* __restrict__ x, const float* __restrict__ se, int C, int hw,
                                               size_t total_vec) {
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total_vec) return;
    size_t e = i * 8;
    int c = (int)(e % C);
    size_t img = e / ((size_t)C * hw);
    float v[8];
    VecIO<__nv_bfloat16, 8>::load(x + e, v);
    const float* s = se + img * C + c;
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] *= s[k];
    VecIO<__nv_bfloat16, 8>::store(x + e, v);
}

template <typename T>
__global__ void k_pool(const T* __restrict__ x, float* __restrict__ feat, int hw, int C) {
    int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int p = 0; p < hw; p++) s += ld1<T>(x + ((size_t)b * hw + p) * C + c);
    feat[(size_t)b * C + c] = s / (float)hw;
}

// One layer of the custom classifier (model.py:50-61, BatchNorm1d folded, Dropout = identity in eval):
//   out[b][j] = act(W[j] . in[b] + bias[j]);  CTA = IPC images x 8 * JR outputs; a warp computes JR outputs at once for all
// IPC images, so an input float4 read from shared memory feeds JR FMAs x 4 (with JR = 1 the layer is bound by shared-memory
// bandwidth: every (image, output) pair streams the whole input vector) and a weight row is fetched once per IPC images.
// Large batches use <16, 4>, small ones <8, 1> (more CTAs, lower latency).  The per-image summation order does not depend
// on IPC / JR (batch invariance).
#define FC_IPC 8
template <int IN, int IPC, int JR>
__global__ void __launch_bounds__(256) k_fc_layer(const float* __restrict__ in, const float* __restrict__ W,
                                                  const float* __restrict__ bias, float* __restrict__ out, int J, int m, int relu) {
    extern __shared__ __align__(16) float s_in[];           // [IPC][IN]
    const int b0 = blockIdx.y * IPC;
    for (int e = threadIdx.x; e < IPC * (IN / 4); e += 256) {
        const int i = e / (IN / 4), c4 = e - i * (IN / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b0 + i < m) v = *(const float4*)(in + (size_t)(b0 + i) * IN + c4 * 4);
        *(float4*)(s_in + i * IN + c4 * 4) = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j0 = (blockIdx.x * 8 + warp) * JR;
    if (j0 >= J) return;
    float a[JR][IPC];
#pragma unroll
    for (int q = 0; q < JR; q++)
#pragma unroll
        for (int i = 0; i < IPC; i++) a[q][i] = 0.f;
    const float4* w4[JR];
#pragma unroll
    for (int q = 0; q < JR; q++) w4[q] = (const float4*)(W + (size_t)min(j0 + q, J - 1) * IN);
    float4 wn[JR];                                          // next iteration's weights, in flight during the FMAs
#pragma unroll
    for (int q = 0; q < JR; q++) wn[q] = __ldg(w4[q] + lane);
#pragma unroll 1
    for (int c4 = lane; c4 < IN / 4; c4 += 32) {
        float4 w[JR];
#pragma unroll
        for (int q = 0; q < JR; q++) w[q] = wn[q];
        if (c4 + 32 < IN / 4) {
#pragma unroll
            for (int q = 0; q < JR; q++) wn[q] = __ldg(w4[q] + c4 + 32);
        }
#pragma unroll
        for (int i = 0; i < IPC; i++) {
            const float4 x = *(const float4*)(s_in + i * IN + c4 * 4);
#pragma unroll
            for (int q = 0; q < JR; q++)
                a[q][i] = fmaf(w[q].x, x.x, fmaf(w[q].y, x.y, fmaf(w[q].z, x.z, fmaf(w[q].w, x.w, a[q][i]))));
        }
    }
#pragma unroll
    for (int q = 0; q < JR; q++) {
        if (j0 + q >= J) break;
        const float bj = bias[j0 + q];
#pragma unroll
        for (int i = 0; i < IPC; i++) {
            float v = a[q][i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && b0 + i < m) { v += bj; out[(size_t)(b0 + i) * J + j0 + q] = relu ? fmaxf(v, 0.f) : v; }
        }
    }
}

template <int IN, int IPC, int JR>
static int fc_launch(dfd_ctx* ctx, const float* in, const float* W, const float* bias, float* out, int J, int m, int relu, cudaStream_t st) {
    { int rc = dfd_func_smem(ctx, k_fc_layer<IN, IPC, JR>, (size_t)IPC * IN * 4); if (rc) return rc; }
    k_fc_layer<IN, IPC, JR><<<dim3((J + 8 * JR - 1) / (8 * JR), (m + IPC - 1) / IPC), 256, IPC * IN * 4, st>>>(in, W, bias, out, J, m, relu);
    return DFD_OK;
}

template <typename T>
__global__ void k_to_f32(const T* __restrict__ x, float* __restrict__ o, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = ld1<T>(x + i);
}

// ---------------------------------------------------------------------------------------------
size_t dfd_effnet_blob_floats() { return eff_offsets().total; }

int dfd_effnet_upload(dfd_ctx* ctx, const float* blob, size_t n) {
    EffOffsets o = eff_offsets();
    DFD_REQUIRE(n == o.total, DFD_ERR_INVALID, "load_weights: blob size does not match dfd_weights_blob_floats()");
    if (!ctx->d_wf32) {
        DFD_CUDA(cudaMalloc(&ctx->d_wf32, o.total * sizeof(float)));
        DFD_CUDA(cudaMalloc(&ctx->d_wbf16, o.total * sizeof(__nv_bfloat16)));
        DFD_CUDA(cudaMalloc(&ctx->d_stem_wg, 32 * 32 * sizeof(__nv_bfloat16)));
    }
    DFD_CUDA(cudaMemcpy(ctx->d_wf32, blob, o.total * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<__nv_bfloat16> h(o.total);
    for (size_t i = 0; i < o.total; i++) h[i] = __float2bfloat16_rn(blob[i]);
    DFD_CUDA(cudaMemcpy(ctx->d_wbf16, h.data(), o.total * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    {   // transposed SE expand weights WxT[j][c] (coalesced reads in k_se_expand)
        size_t tot = 0;
        for (int i = 0; i < 16; i++) tot += (size_t)EFF_BLOCKS[i].cexp * EFF_BLOCKS[i].se;
        std::vector<float> wt(tot);
        size_t at = 0;
        for (int i = 0; i < 16; i++) {
            const EffBlock& b = EFF_BLOCKS[i];
            for (int j = 0; j < b.se; j++)
                for (int c = 0; c < b.cexp; c++) wt[at + (size_t)j * b.cexp + c] = blob[o.blk[i].wx + (size_t)c * b.se + j];
            at += (size_t)b.cexp * b.se;
        }
        if (!ctx->d_wxt) DFD_CUDA(cudaMalloc(&ctx->d_wxt, tot * sizeof(float)));
        DFD_CUDA(cudaMemcpy(ctx->d_wxt, wt.data(), tot * sizeof(float), cudaMemcpyHostToDevice));
    }
    std::vector<__nv_bfloat16> wg(32 * 32);
    for (int n = 0; n < 32; n++)
        for (int k = 0; k < 32; k++) wg[n * 32 + k] = __float2bfloat16_rn(k < 27 ? blob[o.stem_w + (size_t)k * 32 + n] : 0.f);
    DFD_CUDA(cudaMemcpy(ctx->d_stem_wg, wg.data(), wg.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    {   // tf32 hi / lo planes for the fp32 accuracy mode (gemm_tf32x3.cu)
        if (!ctx->d_wtf_hi) {
            DFD_CUDA(cudaMalloc(&ctx->d_wtf_hi, o.total * sizeof(float)));
            DFD_CUDA(cudaMalloc(&ctx->d_wtf_lo, o.total * sizeof(float)));
            DFD_CUDA(cudaMalloc(&ctx->d_stem_wtf, 2 * 32 * 32 * sizeof(float)));
        }
        std::vector<float> hi(o.total), lo(o.total);
        dfd_tf32_split_host(blob, o.total, hi.data(), lo.data());
        DFD_CUDA(cudaMemcpy(ctx->d_wtf_hi, hi.data(), o.total * sizeof(float), cudaMemcpyHostToDevice));
        DFD_CUDA(cudaMemcpy(ctx->d_wtf_lo, lo.data(), o.total * sizeof(float), cudaMemcpyHostToDevice));
        std::vector<float> sw(32 * 32), sh(2 * 32 * 32);
        for (int n = 0; n < 32; n++)
            for (int k = 0; k < 32; k++) sw[n * 32 + k] = k < 27 ? blob[o.stem_w + (size_t)k * 32 + n] : 0.f;
        dfd_tf32_split_host(sw.data(), 32 * 32, sh.data(), sh.data() + 32 * 32);
        DFD_CUDA(cudaMemcpy(ctx->d_stem_wtf, sh.data(), sh.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    { int rc = dfd_front_pack(ctx, blob); if (rc) return rc; }
    {   // block 0 project bias repeated for the 2-pixel folded GEMM rows
        float hb[32];
        for (int j = 0; j < 32; j++) hb[j] = blob[o.blk[0].bp + (j % 16)];
        if (!ctx->d_bias_fold) DFD_CUDA(cudaMalloc(&ctx->d_bias_fold, sizeof hb));
        DFD_CUDA(cudaMemcpy(ctx->d_bias_fold, hb, sizeof hb, cudaMemcpyHostToDevice));
    }
    ctx->w_floats = o.total;
    ctx->has_weights = true;
    return DFD_OK;
}

template <typename T>
static int tap(dfd_ctx* ctx, const char* name, const T* x, size_t n, cudaStream_t st) {
    if (ctx->tap_name.empty() || ctx->tap_name != name) return DFD_OK;
    int rc = dfd_ensure(ctx, ctx->tap, n * sizeof(float));
    if (rc) return rc;
    k_to_f32<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, (float*)ctx->tap.p, n);
    DFD_LAUNCH_CHECK("k_to_f32", st);
    ctx->tap_elems = (int64_t)n;
    return DFD_OK;
}

template <typename T, int VEC>
static int launch_dw(dfd_ctx* ctx, const EffBlock& b, const T* in, const float* W, const float* bias, T* out, int m,
                     int* n_parts, cudaStream_t st) {
    const int CG = b.cexp / VEC;
    int threads = CG >= 256 ? CG : (256 / CG) * CG;
    if (threads > 1024) threads = CG;               // CG <= 288
    const int PS = threads / CG;
    const int npix = b.hout * b.hout;
    // enough CTAs to fill the chip: aim for >= 4 waves but at least 8 pixels per thread-lane
    int ppb = PS * 8;
    if (ppb > npix) ppb = npix;
    const int gx = (npix + ppb - 1) / ppb;
    *n_parts = gx;
    if ((size_t)gx * b.cexp > DFD_POOL_FLOATS) { ctx->err = "internal: squeeze partial buffer too small"; return DFD_ERR_CAPACITY; }
    size_t smem = (size_t)PS * b.cexp * sizeof(float);
#define DW_CASE(KS)                                                                                         \
    k_dw<T, VEC, KS><<<dim3(gx, m), threads, smem, st>>>(in, W, bias, out, ctx->d_pool, b.cexp, b.hin, b.hout, \
                                                         b.s, b.pad, ppb)
    if (b.k == 3) DW_CASE(3); else DW_CASE(5);
#undef DW_CASE
    DFD_LAUNCH_CHECK("k_dw", st);
    return DFD_OK;
}

template <typename T>
static int forward_t(dfd_ctx* ctx, const T* in, int m, float* logits, cudaStream_t st) {
    const EffOffsets o = eff_offsets();
    const float* Wf = ctx->d_wf32;
    constexpr bool BF = sizeof(T) == 2;
    constexpr int VEC = BF ? 8 : 4;
    const bool tc = BF && dfd_gemm_bf16_enabled();
    // fp32 accuracy mode: 3xTF32 tensor-core GEMMs + tiled fp32 depthwise (default); "fp32_simt" keeps the CUDA-core kernels
    const bool tc32 = !BF && !ctx->fp32_simt;
    // activation buffers: x (block input / output ping-pong in act[0], act[1]) and e (expanded, act[2])
    const size_t max_io = (size_t)m * 112 * 112 * 32;          // stem out
    const size_t max_e = (size_t)m * (112 * 112 * 96 + 56 * 56 * 96);   // block 1: expand out + depthwise out
    int rc;
    if ((rc = dfd_ensure(ctx, ctx->act[0], max_io * sizeof(T)))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->act[1], max_io * sizeof(T)))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->act[2], max_e * sizeof(T)))) return rc;
    T* x = (T*)ctx->act[0].p;
    T* y = (T*)ctx->act[1].p;
    T* e = (T*)ctx->act[2].p;

    static const char* L_EXP[16] = {"b0.expand", "b1.expand", "b2.expand", "b3.expand", "b4.expand", "b5.expand", "b6.expand", "b7.expand", "b8.expand", "b9.expand", "b10.expand", "b11.expand", "b12.expand", "b13.expand", "b14.expand", "b15.expand"};
    static const char* L_DW[16] = {"b0.dw", "b1.dw", "b2.dw", "b3.dw", "b4.dw", "b5.dw", "b6.dw", "b7.dw", "b8.dw", "b9.dw", "b10.dw", "b11.dw", "b12.dw", "b13.dw", "b14.dw", "b15.dw"};
    static const char* L_FRONT[16] = {"b0.front", "b1.front", "b2.front", "b3.front", "b4.front", "b5.front", "b6.front", "b7.front", "b8.front", "b9.front", "b10.front", "b11.front", "b12.front", "b13.front", "b14.front", "b15.front"};
    static const char* L_SE[16] = {"b0.se", "b1.se", "b2.se", "b3.se", "b4.se", "b5.se", "b6.se", "b7.se", "b8.se", "b9.se", "b10.se", "b11.se", "b12.se", "b13.se", "b14.se", "b15.se"};
    static const char* L_PROJ[16] = {"b0.project", "b1.project", "b2.project", "b3.project", "b4.project", "b5.project", "b6.project", "b7.project", "b8.project", "b9.project", "b10.project", "b11.project", "b12.project", "b13.project", "b14.project", "b15.project"};
    ctx->label = "stem";
    {
        int total = m * 112 * 112;
        if (tc) {
            if ((rc = dfd_gemm_bf16_ex(ctx, 2, (const __nv_bfloat16*)in, nullptr, 0, ctx->d_stem_wg, Wf + o.stem_b, nullptr,
                                       (__nv_bfloat16*)x, total, 32, 32, 1, st))) return rc;
        } else if (tc32) {
            if ((rc = dfd_gemm_tf32x3(ctx, 2, (const float*)in, nullptr, 0, ctx->d_stem_wtf, ctx->d_stem_wtf + 32 * 32, Wf + o.stem_b,
                                      nullptr, (float*)x, total, 32, 32, 1, st))) return rc;
        } else {
            k_stem<T><<<(total + 127) / 128, 128, 0, st>>>(in, Wf + o.stem_w, Wf + o.stem_b, x, total);
            DFD_LAUNCH_CHECK("k_stem", st);
        }
        if ((rc = tap<T>(ctx, "stem", x, (size_t)total * 32, st))) return rc;
    }
    auto pw = [&](const T* A, size_t w_off, size_t b_off, const float* se, int hw, const T* res, T* C, int M, int N, int K,
                  int act) -> int {
        if (tc) {
            return dfd_gemm_bf16(ctx, (const __nv_bfloat16*)A, ctx->d_wbf16 + w_off, Wf + b_off, (const __nv_bfloat16*)res,
                                 (__nv_bfloat16*)C, M, N, K, act, st);
        }
        if (tc32)       // (a gate is applied while the A tile is staged: the gated tensor never exists in HBM)
            return dfd_gemm_tf32x3(ctx, se ? 1 : 0, (const float*)A, se, hw, ctx->d_wtf_hi + w_off, ctx->d_wtf_lo + w_off, Wf + b_off,
                                   (const float*)res, (float*)C, M, N, K, act, st);
        // 32-row tiles when 64-row tiles would leave SMs idle
        if ((size_t)((M + 63) / 64) * ((N + 63) / 64) < (size_t)2 * ctx->sm_count)
            k_pw<T, 32><<<dim3((M + 31) / 32, (N + 63) / 64), 128, 0, st>>>(A, Wf + w_off, Wf + b_off, se, hw, res, C, M, N, K, act);
        else
            k_pw<T, 64><<<dim3((M + 63) / 64, (N + 63) / 64), 256, 0, st>>>(A, Wf + w_off, Wf + b_off, se, hw, res, C, M, N, K, act);
        DFD_LAUNCH_CHECK("k_pw", st);
        return DFD_OK;
    };
    char nm[32];
    size_t wxt_off[16];
    { size_t acc = 0; for (int i = 0; i < 16; i++) { wxt_off[i] = acc; acc += (size_t)EFF_BLOCKS[i].cexp * EFF_BLOCKS[i].se; } }
    for (int i = 0; i < 16; i++) {
        const EffBlock& b = EFF_BLOCKS[i];
        const EffBlockOff& f = o.blk[i];
        const int Min = m * b.hin * b.hin, Mout = m * b.hout * b.hout;
        const T* dw_in = x;
        T* dw_out;
        int n_parts = 1;
        // bf16: expand + depthwise run as ONE kernel (mbconv_fused.cu) and the expanded tensor never reaches HBM;
        // the two-kernel path remains for the ".expand" diagnostic tap and DFD_NO_FUSE=1 (A/B testing)
        snprintf(nm, sizeof nm, "b%d.expand", i);
        const bool fused = tc && b.cexp != b.cin && !ctx->no_fuse && ctx->tap_name != nm;
        if (fused) {
            if constexpr (BF) {
                dw_out = e;
                ctx->label = L_FRONT[i];
                if ((rc = dfd_mbconv_front_bf16(ctx, i, (const __nv_bfloat16*)x, ctx->d_wbf16 + f.we, (__nv_bfloat16*)dw_out, m,
                                                &n_parts, st))) return rc;
            }
        } else {
            // fp32 accuracy mode: expand + depthwise run as two kernels, so the expanded tensor (6x the block input, fp32) would
            // make a round trip through HBM: 1.23 GB written and read back for block 1 at batch 256.  It is produced and consumed
            // in SUB-BATCHES whose expanded tensor fits the 126 MB L2 (same buffer every time): the depthwise kernel then reads it
            // from L2 and the dirty lines are overwritten before they are ever evicted.  MEASURED AND OFF BY DEFAULT: 6.33 ms per
            // 256 crops without, 6.72 / 6.95 / 7.20 ms with 96 / 64 / 48 MB sub-batches (b1.expand 0.38 -> 0.62 ms over 20 launches):
            // neither kernel is HBM-bound (epilogue ALU / shared-memory issue), so the shorter launches only add tails.
            int sub = m;
            if (tc32 && b.cexp != b.cin && ctx->tap_name != nm && !ctx->no_subbatch) {
                const size_t per_img = (size_t)b.hin * b.hin * b.cexp * sizeof(T);
                const size_t fit = (size_t)ctx->l2_budget / per_img;
                if (fit < (size_t)m) sub = fit < 8 ? 8 : (int)fit;
                if (sub > m) sub = m;
            }
            if (sub < m) {
                T* dwo = e + (size_t)sub * b.hin * b.hin * b.cexp;          // depthwise output of the whole batch behind ONE sub-batch of expanded data
                if (((size_t)sub * b.hin * b.hin + (size_t)Mout) * b.cexp * sizeof(T) > ctx->act[2].bytes) { ctx->err = "internal: expanded buffer too small"; return DFD_ERR_CAPACITY; }
                for (int m0 = 0; m0 < m; m0 += sub) {
                    const int mc = m - m0 < sub ? m - m0 : sub;
                    ctx->label = L_EXP[i];
                    if ((rc = pw(x + (size_t)m0 * b.hin * b.hin * b.cin, f.we, f.be, nullptr, 0, nullptr, e, mc * b.hin * b.hin, b.cexp, b.cin, 1))) return rc;
                    ctx->label = L_DW[i];
                    if ((rc = dfd_dw_f32(ctx, b, (const float*)e, Wf + f.wd, Wf + f.bd, (float*)(dwo + (size_t)m0 * b.hout * b.hout * b.cexp), mc,
                                         &n_parts, st, m0))) return rc;
                }
                dw_out = dwo;
                goto dw_done;
            }
            if (b.cexp != b.cin) {
                ctx->label = L_EXP[i];
                if ((rc = pw(x, f.we, f.be, nullptr, 0, nullptr, e, Min, b.cexp, b.cin, 1))) return rc;
                if ((rc = tap<T>(ctx, nm, e, (size_t)Min * b.cexp, st))) return rc;
                dw_in = e;
            }
            // depthwise output: behind the expand output inside e (sized for block 1), or y for block 0
            if (b.cexp != b.cin) {
                dw_out = e + (size_t)Min * b.cexp;
                size_t need = ((size_t)Min + Mout) * b.cexp * sizeof(T);
                if (need > ctx->act[2].bytes) { ctx->err = "internal: expanded buffer too small"; return DFD_ERR_CAPACITY; }
            } else dw_out = y;
            ctx->label = L_DW[i];
            if constexpr (BF) {
                if ((rc = dfd_dw_bf16(ctx, b, (const __nv_bfloat16*)dw_in, Wf + f.wd, Wf + f.bd, (__nv_bfloat16*)dw_out, m, &n_parts, st))) return rc;
            } else if (tc32) {
                if ((rc = dfd_dw_f32(ctx, b, (const float*)dw_in, Wf + f.wd, Wf + f.bd, (float*)dw_out, m, &n_parts, st, 0))) return rc;
            } else {
                if ((rc = launch_dw<T, VEC>(ctx, b, dw_in, Wf + f.wd, Wf + f.bd, dw_out, m, &n_parts, st))) return rc;
            }
        }
    dw_done:
        snprintf(nm, sizeof nm, "b%d.dw", i);
        if ((rc = tap<T>(ctx, nm, dw_out, (size_t)Mout * b.cexp, st))) return rc;
        ctx->label = L_SE[i];
        // blocks 0-4 (>= 784 rows per image): the SE gate is folded into per-image project weights by the SE kernel
        // (ctx->gated_w_max: last block that does so; 4 = the >= 784-row blocks, 10 = also the 14x14 stage)
        const bool gated_w = BF && tc && i <= ctx->gated_w_max && ctx->se_mode == 2 && !ctx->no_gated_w;
        // block 0 (K = 32): two pixels per GEMM row, so the TMA moves full 128-byte rows (its row rate, not bytes, is the limit)
        const int fold = (gated_w && i == 0 && !ctx->no_fold) ? 2 : 1;
        __nv_bfloat16* wg_buf = fold > 1 ? ctx->d_wgated_fold : ctx->d_wgated;
        if (BF && ctx->se_mode == 3) {
            // timing experiment only (DFD_SE_MODE=3): no SE excite at all, gates stay whatever they were
        } else if ((BF || tc32) && ctx->se_mode == 2) {
            DFD_CUDA(dfd_launch(ctx->pdl, k_se_cluster, dim3((m + SE_CL - 1) / SE_CL * SE_CL), dim3(256), 0, st, (const float*)ctx->d_pool, n_parts,
                                (const float*)(Wf + f.wr), (const float*)(Wf + f.br), (const float*)(ctx->d_wxt + wxt_off[i]),
                                (const float*)(Wf + f.bx), ctx->d_sescale, b.cexp, b.se, 1.0f / (float)(b.hout * b.hout), m,
                                (const float*)(gated_w ? Wf + f.wp : nullptr), gated_w ? wg_buf : (__nv_bfloat16*)nullptr, b.cout, fold));
            DFD_LAUNCH_CHECK("k_se_cluster", st);
        } else if (BF && ctx->se_mode != 0) {
            k_se_excite<<<(m + SE_X_IPC - 1) / SE_X_IPC, SE_X_THREADS, 0, st>>>(ctx->d_pool, n_parts, Wf + f.wr, Wf + f.br, ctx->d_wxt + wxt_off[i],
                                                                        Wf + f.bx, ctx->d_sescale, b.cexp, b.se,
                                                                        1.0f / (float)(b.hout * b.hout), m);
            DFD_LAUNCH_CHECK("k_se_excite", st);
        } else {
            const int mg = (m + SE_IPC - 1) / SE_IPC;
            k_se_reduce<<<dim3((b.se + 7) / 8, mg), 256, 0, st>>>(ctx->d_pool, n_parts, Wf + f.wr, Wf + f.br, ctx->d_se_r, b.cexp,
                                                                  b.se, 1.0f / (float)(b.hout * b.hout), m);
            DFD_LAUNCH_CHECK("k_se_reduce", st);
            k_se_expand<<<dim3((b.cexp + 255) / 256, mg), 256, 0, st>>>(ctx->d_se_r, ctx->d_wxt + wxt_off[i], Wf + f.bx,
                                                                        ctx->d_sescale, b.cexp, b.se, m);
            DFD_LAUNCH_CHECK("k_se_expand", st);
        }
        const bool skip = b.s == 1 && b.cin == b.cout;
        T* outp = (dw_out == y) ? x : y;       // block 0 wrote dw into y; its project output goes to x (input is dead, no skip)
        ctx->label = L_PROJ[i];
        if (gated_w) {
            if ((rc = dfd_gemm_bf16_img(ctx, (const __nv_bfloat16*)dw_out, wg_buf, fold > 1 ? ctx->d_bias_fold : Wf + f.bp,
                                        (const __nv_bfloat16*)(skip ? x : nullptr), (__nv_bfloat16*)outp, m, b.hout * b.hout / fold,
                                        b.cout * fold, b.cexp * fold, 0, st))) return rc;
        } else if (tc) {      // the SE gate is applied while the A tile is staged (A_SCALE)
            if ((rc = dfd_gemm_bf16_ex(ctx, 1, (const __nv_bfloat16*)dw_out, ctx->d_sescale, b.hout * b.hout,
                                       ctx->d_wbf16 + f.wp, Wf + f.bp, (const __nv_bfloat16*)(skip ? x : nullptr),
                                       (__nv_bfloat16*)outp, Mout, b.cout, b.cexp, 0, st))) return rc;
        } else {
            if ((rc = pw(dw_out, f.wp, f.bp, ctx->d_sescale, b.hout * b.hout, skip ? x : nullptr, outp, Mout, b.cout, b.cexp, 0))) return rc;
        }
        snprintf(nm, sizeof nm, "b%d.out", i);
        if ((rc = tap<T>(ctx, nm, outp, (size_t)Mout * b.cout, st))) return rc;
        if (outp == y) { T* t = x; x = y; y = t; }
    }
    // head 1x1 320->1280 + swish, global average pool, classifier
    {
        const int M = m * 49;
        ctx->label = "head";
        if ((rc = pw(x, o.head_w, o.head_b, nullptr, 0, nullptr, e, M, 1280, 320, 1))) return rc;
        ctx->label = "pool";
        k_pool<T><<<dim3(1280 / 128, m), 128, 0, st>>>(e, ctx->d_feat, 49, 1280);
        DFD_LAUNCH_CHECK("k_pool", st);
        if (!ctx->tap_name.empty() && ctx->tap_name == "features") {
            if ((rc = tap<float>(ctx, "features", ctx->d_feat, (size_t)m * 1280, st))) return rc;
        }
        ctx->label = "fc";
        {
            if (m >= 64) {
                if ((rc = fc_launch<1280, 16, 4>(ctx, ctx->d_feat, Wf + o.fc1_w, Wf + o.fc1_b, ctx->d_fc_h1, 512, m, 1, st))) return rc;
                DFD_LAUNCH_CHECK("k_fc", st);
                if ((rc = fc_launch<512, 16, 4>(ctx, ctx->d_fc_h1, Wf + o.fc2_w, Wf + o.fc2_b, ctx->d_fc_h2, 256, m, 1, st))) return rc;
            } else {
                if ((rc = fc_launch<1280, FC_IPC, 1>(ctx, ctx->d_feat, Wf + o.fc1_w, Wf + o.fc1_b, ctx->d_fc_h1, 512, m, 1, st))) return rc;
                DFD_LAUNCH_CHECK("k_fc", st);
                if ((rc = fc_launch<512, FC_IPC, 1>(ctx, ctx->d_fc_h1, Wf + o.fc2_w, Wf + o.fc2_b, ctx->d_fc_h2, 256, m, 1, st))) return rc;
            }
            DFD_LAUNCH_CHECK("k_fc", st);
            if ((rc = fc_launch<256, FC_IPC, 1>(ctx, ctx->d_fc_h2, Wf + o.fc3_w, Wf + o.fc3_b, logits, 1, m, 0, st))) return rc;
        }
        DFD_LAUNCH_CHECK("k_fc", st);
        ctx->label = "";
    }
    return DFD_OK;
}

static int effnet_forward_any(dfd_ctx* ctx, const void* in, int m, int dtype, float* logits, cudaStream_t st) {
    if (dtype == DFD_F32) return forward_t<float>(ctx, (const float*)in, m, logits, st);
    return forward_t<__nv_bfloat16>(ctx, (const __nv_bfloat16*)in, m, logits, st);
}

// Large batches run as TWO independent half-batch chains on two streams (the caller's and ctx->aux2).  A single chain leaves
// SMs idle INSIDE its kernels -- the 16 squeeze-excite launches are latency chains that occupy a fraction of the chip, the
// deep layers (M = batch x 49 rows) have fewer tiles than SMs, every kernel ends in a partial wave -- and nothing else is
// runnable, because each layer depends on the previous one.  The second chain is: its CTAs fill those holes.  The halves share
// nothing but the (read-only) weights: the second one gets its own activation buffers and its slice of every per-image
// scratch array.  Results are bit-identical to the single chain (every kernel is batch-invariant: tests/test_gpu_effnet.py).
int dfd_effnet_launch(dfd_ctx* ctx, const void* in, int m, int dtype, float* logits, cudaStream_t st) {
    DFD_REQUIRE(ctx->has_weights, DFD_ERR_NO_WEIGHTS, "effnet_forward: call dfd_load_weights first");
    DFD_REQUIRE(m > 0 && m <= ctx->cfg.max_batch, DFD_ERR_CAPACITY, "effnet_forward: batch exceeds max_batch");
    DFD_REQUIRE(dtype == DFD_F32 || dtype == DFD_BF16, DFD_ERR_INVALID, "effnet_forward: bad dtype");
    const bool dual = ctx->dual_chain && m >= ctx->dual_min && !ctx->profiling && !ctx->trace && ctx->tap_name.empty() &&
                      !(dtype == DFD_F32 && ctx->fp32_simt);
    if (!dual) return effnet_forward_any(ctx, in, m, dtype, logits, st);
    if (!ctx->aux2) {
        DFD_CUDA(cudaStreamCreateWithFlags(&ctx->aux2, cudaStreamNonBlocking));
        DFD_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming));
        DFD_CUDA(cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming));
    }
    const int m0 = ((m + 1) / 2 + 7) & ~7, m1 = m - m0;       // (gated tiles / clusters like multiples of 8 images)
    const size_t esz = dtype == DFD_BF16 ? 2 : 4;
    DFD_CUDA(cudaEventRecord(ctx->ev_fork2, st));
    DFD_CUDA(cudaStreamWaitEvent(ctx->aux2, ctx->ev_fork2, 0));
    int rc = effnet_forward_any(ctx, in, m0, dtype, logits, st);
    if (rc) return rc;
    // second half: its own activation buffers, its slice of the per-image scratch (pointers are launch arguments: swapping them
    // on the host while the first half's kernels are queued is safe)
    struct Saved { DfdBuf act[3]; float *pool, *sescale, *se_r, *feat, *h1, *h2; __nv_bfloat16 *wg, *wgf; } sv;
    for (int i = 0; i < 3; i++) { sv.act[i] = ctx->act[i]; ctx->act[i] = ctx->act_b[i]; }
    sv.pool = ctx->d_pool; sv.sescale = ctx->d_sescale; sv.se_r = ctx->d_se_r; sv.feat = ctx->d_feat; sv.h1 = ctx->d_fc_h1; sv.h2 = ctx->d_fc_h2;
    sv.wg = ctx->d_wgated; sv.wgf = ctx->d_wgated_fold;
    ctx->d_pool += (size_t)m0 * DFD_POOL_FLOATS; ctx->d_sescale += (size_t)m0 * 1152; ctx->d_se_r += (size_t)m0 * 64;
    ctx->d_feat += (size_t)m0 * 1280; ctx->d_fc_h1 += (size_t)m0 * 512; ctx->d_fc_h2 += (size_t)m0 * 256;
    ctx->d_wgated += (size_t)m0 * 112 * 672; ctx->d_wgated_fold += (size_t)m0 * 32 * 64;
    rc = effnet_forward_any(ctx, (const uint8_t*)in + (size_t)m0 * 224 * 224 * 3 * esz, m1, dtype, logits + m0, ctx->aux2);
    for (int i = 0; i < 3; i++) { ctx->act_b[i] = ctx->act[i]; ctx->act[i] = sv.act[i]; }
    ctx->d_pool = sv.pool; ctx->d_sescale = sv.sescale; ctx->d_se_r = sv.se_r; ctx->d_feat = sv.feat; ctx->d_fc_h1 = sv.h1; ctx->d_fc_h2 = sv.h2;
    ctx->d_wgated = sv.wg; ctx->d_wgated_fold = sv.wgf;
    if (rc) return rc;
    DFD_CUDA(cudaEventRecord(ctx->ev_join2, ctx->aux2));
    DFD_CUDA(cudaStreamWaitEvent(st, ctx->ev_join2, 0));
    return DFD_OK;
}
