// Depthwise kxk convolution + folded BN + swish + fused SE squeeze for bf16 NHWC activations
// (the _depthwise_conv/_bn1/swish/avg-pool stage of every MBConv block).
//
// HBM-bound stage (12.2 MB/img of the 27.4 MB/img layer-granular traffic).  One CTA owns an output tile of
// TH x TW pixels x 64 channels of one image:
//   1. the (TH-1)*S+K by (TW-1)*S+K input patch is staged in shared memory with 16-byte cp.async
//      (zero-fill outside the image = TF-SAME padding, and beyond C for ragged channel chunks);
//      a pixel's 64 channels are one 128-byte row, so every later warp access is conflict-free;
//   2. lane = channel pair, warp = output row: the warp slides along its row keeping TW fp32
//      accumulator pairs in registers, so each staged input word is read once per kernel row
//      and reused for up to K outputs;
//   3. swish, bf16 pack, 128-byte coalesced stores; per-channel sums of the swish outputs (the SE
//      squeeze) are reduced across the CTA's warps in a fixed order and written as one partial per
//      (image, tile, channel) -- no atomics, so the forward pass is bit-reproducible; k_se adds the partials.
#include "dfd_internal.cuh"
#include "effnet_plan.h"

#define DW_WARPS 8

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ float dw_swish(float x) {
    // x * sigmoid(x) = h + h*tanh(h), h = x/2 : one MUFU op
    float h = 0.5f * x, t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
// two fp32 FMAs in one instruction (Blackwell FFMA2): d.xy += a.xy * b.xy
__device__ __forceinline__ void ffma2(uint64_t& d, uint64_t a, uint64_t b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// CC = channels per CTA (64: lane = channel pair, one output row per warp; 32: half-warp = channel pairs,
// two output rows per warp).
template <int K, int S, int TW, int TH, int CC>
__global__ void __launch_bounds__(DW_WARPS * 32)
k_dw_tile(const __nv_bfloat16* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
          __nv_bfloat16* __restrict__ out, float* __restrict__ pool, int C, int hin, int hout, int pad, int tiles_x) {
    constexpr int PH = (TH - 1) * S + K, PW = (TW - 1) * S + K;
    constexpr int LP = CC / 2;                     // lanes (32-bit words) per pixel
    constexpr int RW = 32 / LP;                    // output rows per warp pass
    constexpr int PWP = (CC == 32 && (PW % 2 == 0)) ? PW + 1 : PW;   // odd pixel pitch keeps the two half-warps on different banks
    constexpr int CHUNKS = CC / 8;                 // 16-byte chunks per pixel
    extern __shared__ __align__(16) uint32_t smem_dw[];
    uint32_t* patch = smem_dw;                               // [PH][PWP][LP] bf16x2
    float* sw = (float*)(patch + PH * PWP * LP);             // [K*K][CC]
    float* spool = sw + K * K * CC;                          // [DW_WARPS * RW][CC]
    const int tile = blockIdx.x, chunk = blockIdx.y, b = blockIdx.z;
    const int ty = tile / tiles_x, tx = tile % tiles_x;
    const int oy0 = ty * TH, ox0 = tx * TW;
    const int c0 = chunk * CC;
    const int iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- stage the input patch with 16-byte cp.async (zero fill = TF-SAME padding / ragged channel chunk) ----
    const uint32_t patch_s = (uint32_t)__cvta_generic_to_shared(patch);
    const __nv_bfloat16* img = in + (size_t)b * hin * hin * C;
    for (int idx = tid; idx < PH * PW * CHUNKS; idx += DW_WARPS * 32) {
        const int pix = idx / CHUNKS, part = idx - pix * CHUNKS;
        const int py = pix / PW, px = pix - py * PW;
        const int iy = iy0 + py, ix = ix0 + px, c = c0 + part * 8;
        const bool ok = iy >= 0 && iy < hin && ix >= 0 && ix < hin && c < C;
        const __nv_bfloat16* src = ok ? img + ((size_t)iy * hin + ix) * C + c : in;
        cp_async16(patch_s + (uint32_t)(((py * PWP + px) * LP) * 4 + part * 16), src, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = tid; i < K * K * CC; i += DW_WARPS * 32) {
        const int c = c0 + (i % CC);
        sw[i] = c < C ? W[(size_t)(i / CC) * C + c] : 0.f;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int sub = lane / LP, pl = lane % LP;               // row within the warp pass, channel pair
    const int ch = c0 + 2 * pl;
    const bool ch_ok = ch < C;                               // C is even: a pair is valid or not as a whole
    const uint64_t bias2 = ch_ok ? pack2(bias[ch], bias[ch + 1]) : pack2(0.f, 0.f);
    float ps0 = 0.f, ps1 = 0.f;
    for (int r = warp * RW + sub; r < TH; r += DW_WARPS * RW) {
        const int oy = oy0 + r;
        if (oy >= hout) break;
        uint64_t acc[TW];
#pragma unroll
        for (int i = 0; i < TW; i++) acc[i] = bias2;
#pragma unroll
        for (int ky = 0; ky < K; ky++) {
            uint64_t w[K];
#pragma unroll
            for (int kx = 0; kx < K; kx++) w[kx] = *(const uint64_t*)(sw + (ky * K + kx) * CC + 2 * pl);
            const uint32_t* prow = patch + (size_t)((r * S + ky) * PWP) * LP + pl;
#pragma unroll
            for (int ix = 0; ix < PW; ix++) {
                const uint32_t v = prow[ix * LP];
                const uint64_t x = pack2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
#pragma unroll
                for (int kx = 0; kx < K; kx++)
                    if ((ix - kx) % S == 0 && (ix - kx) >= 0 && (ix - kx) / S < TW) ffma2(acc[(ix - kx) / S], x, w[kx]);
            }
        }
        __nv_bfloat16* orow = out + (((size_t)b * hout + oy) * hout + ox0) * C + ch;
#pragma unroll
        for (int i = 0; i < TW; i++) {
            if (ox0 + i < hout && ch_ok) {
                // swish on the pair with packed fp32 math (h = 0.5 x exactly, y = h + h * tanh(h): same values as dw_swish)
                uint64_t h2 = 0, y2;
                ffma2(h2, acc[i], pack2(0.5f, 0.5f));
                float h0, h1;
                unpack2(h2, h0, h1);
                float t0, t1;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
                y2 = h2;
                ffma2(y2, h2, pack2(t0, t1));
                float y0, y1;
                unpack2(y2, y0, y1);
                ps0 += y0; ps1 += y1;
                *(__nv_bfloat162*)(orow + (size_t)i * C) = __floats2bfloat162_rn(y0, y1);
            }
        }
    }
    spool[(warp * RW + sub) * CC + 2 * pl] = ps0; spool[(warp * RW + sub) * CC + 2 * pl + 1] = ps1;
    __syncthreads();
    if (tid < CC && c0 + tid < C) {                          // deterministic: fixed-order sum, one partial per tile
        float sacc = 0.f;
#pragma unroll
        for (int wv = 0; wv < DW_WARPS * RW; wv++) sacc += spool[wv * CC + tid];
        pool[((size_t)b * gridDim.x + tile) * C + c0 + tid] = sacc;
    }
}

template <int K, int S, int TW, int TH, int CC>
static int launch(dfd_ctx* ctx, const EffBlock& b, const __nv_bfloat16* in, const float* W, const float* bias,
                  __nv_bfloat16* out, int m, int* n_parts, cudaStream_t st) {
    constexpr int PH = (TH - 1) * S + K, PW = (TW - 1) * S + K;
    constexpr int PWP = (CC == 32 && (PW % 2 == 0)) ? PW + 1 : PW;
    constexpr int RW = 64 / CC;
    const size_t smem = (size_t)PH * PWP * CC * 2 + (size_t)K * K * CC * 4 + (size_t)DW_WARPS * RW * CC * 4;
    { int rc = dfd_func_smem(ctx, k_dw_tile<K, S, TW, TH, CC>, smem); if (rc) return rc; }
    const int tiles_x = (b.hout + TW - 1) / TW, tiles_y = (b.hout + TH - 1) / TH;
    dim3 grid(tiles_x * tiles_y, (b.cexp + CC - 1) / CC, m);
    *n_parts = tiles_x * tiles_y;
    if ((size_t)grid.x * b.cexp > DFD_POOL_FLOATS) { ctx->err = "internal: squeeze partial buffer too small"; return DFD_ERR_CAPACITY; }
    k_dw_tile<K, S, TW, TH, CC><<<grid, DW_WARPS * 32, smem, st>>>(in, W, bias, out, ctx->d_pool, b.cexp, b.hin, b.hout, b.pad, tiles_x);
    DFD_LAUNCH_CHECK("k_dw_tile", st);
    return DFD_OK;
}

int dfd_dw_bf16(dfd_ctx* ctx, const EffBlock& b, const __nv_bfloat16* in, const float* W, const float* bias,
                __nv_bfloat16* out, int m, int* n_parts, cudaStream_t st) {
    // tile shapes per output size: 112 (C=32) -> 16x16 with 32-channel CTAs, 56 -> 8x14, 28 / 14 -> 7x14, 7 -> 7x7
    if (b.k == 3 && b.s == 1 && b.hout == 112) return launch<3, 1, 16, 16, 32>(ctx, b, in, W, bias, out, m, n_parts, st);
    if (b.k == 3 && b.s == 2 && b.hout == 56) return launch<3, 2, 14, 8, 64>(ctx, b, in, W, bias, out, m, n_parts, st);
    if (b.k == 3 && b.s == 1 && b.hout == 56) return launch<3, 1, 14, 8, 64>(ctx, b, in, W, bias, out, m, n_parts, st);
    if (b.k == 5 && b.s == 2 && b.hout == 28) return launch<5, 2, 14, 7, 64>(ctx, b, in, W, bias, out, m, n_parts, st);
    if (b.k == 5 && b.s == 1 && b.hout == 28) return launch<5, 1, 14, 7, 64>(ctx, b, in, W, bias, out, m, n_parts, st);
    if (b.k == 3 && b.s == 2 && b.hout == 14) return launch<3, 2, 14, 7, 64>(ctx, b, in, W, bias, out, m, n_parts, st);
    if (b.k == 3 && b.s == 1 && b.hout == 14) return launch<3, 1, 14, 7, 64>(ctx, b, in, W, bias, out, m, n_parts, st);
    if (b.k == 5 && b.s == 1 && b.hout == 14) return launch<5, 1, 14, 7, 64>(ctx, b, in, W, bias, out, m, n_parts, st);
    if (b.k == 5 && b.s == 2 && b.hout == 7) return launch<5, 2, 7, 7, 64>(ctx, b, in, W, bias, out, m, n_parts, st);
    if (b.k == 5 && b.s == 1 && b.hout == 7) return launch<5, 1, 7, 7, 64>(ctx, b, in, W, bias, out, m, n_parts, st);
    if (b.k == 3 && b.s == 1 && b.hout == 7) return launch<3, 1, 7, 7, 64>(ctx, b, in, W, bias, out, m, n_parts, st);
    ctx->err = "dw_bf16: no tile configuration for this layer";
    return DFD_ERR_INVALID;
}
