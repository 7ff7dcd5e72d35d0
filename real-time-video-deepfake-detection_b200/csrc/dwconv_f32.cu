// Depthwise kxk convolution + folded BN + swish + fused SE squeeze on fp32 NHWC activations: the fp32 accuracy mode of the
// classifier (reference model.py:63-72 in fp32).  Same tiled design as dwconv_bf16.cu:
// CTA = TH x TW output pixels x 32 channels of one image (a pixel's 32 channels are one 128-byte shared-memory row),
// lane = channel, warp = output row(s), plain fp32 FMAs in the reference's tap order (ky, kx ascending, like the fp32 oracle's
// conv), swish_f32 (~3 ulp; not the approximate tanh of the bf16 path, which is 2^-11: the fp32 gate is 1e-4 on the
// probability).  Squeeze partials: fixed order, no atomics.
//
// The input patch is fetched by ONE 4-D TMA box [32 ch][PW][PH][1 image] (fp32, no swizzle: the box lands as [PH][PW][32]):
// out-of-image coordinates and channels beyond C are zero-filled by the TMA, which is exactly the TF-"SAME" padding and the
// ragged last channel chunk.  ncu showed the cp.async version at 70-82 % issue-active with a quarter of its instructions
// spent on per-chunk index arithmetic for the staging loop; the TMA version issues none.
#include "dfd_internal.cuh"
#include "effnet_plan.h"
#include <cuda.h>
#include "tc_ptx.cuh"

#define DWF_WARPS 8

int dfd_tmap_encode(dfd_ctx* ctx, CUtensorMap* m, int dtype_f32, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

__device__ __forceinline__ void dwf_tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

template <int K, int S, int TW, int TH>
__global__ void __launch_bounds__(DWF_WARPS * 32)
k_dw_tile_f32(const __grid_constant__ CUtensorMap map_in, const float* __restrict__ W, const float* __restrict__ bias,
              float* __restrict__ out, float* __restrict__ pool, int C, int hout, int pad, int tiles_x) {
    constexpr int PH = (TH - 1) * S + K, PW = (TW - 1) * S + K;
    constexpr int CC = 32;
    extern __shared__ __align__(128) uint8_t smem_dwf[];
    __shared__ __align__(8) uint64_t bar;
    float* patch = (float*)smem_dwf;                         // [PH][PW][32]
    float* sw = patch + PH * PW * CC;                        // [K*K][32]
    float* spool = sw + K * K * CC;                          // [DWF_WARPS][32]
    const int tile = blockIdx.x, chunk = blockIdx.y, b = blockIdx.z;
    const int ty = tile / tiles_x, tx = tile % tiles_x;
    const int oy0 = ty * TH, ox0 = tx * TW;
    const int c0 = chunk * CC;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_s = smem_u32(&bar);
    if (tid == 0) {
        mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_s, (uint32_t)(PH * PW * CC * 4));
        dwf_tma_load_4d(smem_u32(patch), &map_in, c0, ox0 * S - pad, oy0 * S - pad, b, bar_s);
    }
    for (int i = tid; i < K * K * CC; i += DWF_WARPS * 32) {
        const int c = c0 + (i % CC);
        sw[i] = c < C ? W[(size_t)(i / CC) * C + c] : 0.f;
    }
    const int ch = c0 + lane;
    const bool ch_ok = ch < C;
    const float bv = ch_ok ? bias[ch] : 0.f;
    __syncthreads();                                         // weights staged, barrier initialised
    mbar_wait(bar_s, 0);                                     // patch landed

    float ps = 0.f;
    for (int r = warp; r < TH; r += DWF_WARPS) {
        const int oy = oy0 + r;
        if (oy >= hout) break;
        float acc[TW];
#pragma unroll
        for (int i = 0; i < TW; i++) acc[i] = bv;
#pragma unroll
        for (int ky = 0; ky < K; ky++) {
            float w[K];
#pragma unroll
            for (int kx = 0; kx < K; kx++) w[kx] = sw[(ky * K + kx) * CC + lane];
            const float* prow = patch + (size_t)((r * S + ky) * PW) * CC + lane;
#pragma unroll
            for (int ix = 0; ix < PW; ix++) {
                const float x = prow[ix * CC];
#pragma unroll
                for (int kx = 0; kx < K; kx++)
                    if ((ix - kx) % S == 0 && (ix - kx) >= 0 && (ix - kx) / S < TW) acc[(ix - kx) / S] = fmaf(x, w[kx], acc[(ix - kx) / S]);
            }
        }
        // every tile shape divides its layer's output width exactly (launch table below): no per-column bounds test
        float* orow = out + (((size_t)b * hout + oy) * hout + ox0) * C + ch;
        float y[TW];
#pragma unroll
        for (int i = 0; i + 1 < TW; i += 2) {
            f2_unpack(swish_f32x2(f2_pack(acc[i], acc[i + 1])), y[i], y[i + 1]);
            ps += y[i]; ps += y[i + 1];
        }
        if (TW & 1) { y[TW - 1] = swish_f32(acc[TW - 1]); ps += y[TW - 1]; }
        if (ch_ok) {
#pragma unroll
            for (int i = 0; i < TW; i++) orow[(size_t)i * C] = y[i];
        }
    }
    spool[warp * CC + lane] = ch_ok ? ps : 0.f;
    __syncthreads();
    if (tid < CC && c0 + tid < C) {
        float sacc = 0.f;
#pragma unroll
        for (int wv = 0; wv < DWF_WARPS; wv++) sacc += spool[wv * CC + tid];
        pool[((size_t)b * gridDim.x + tile) * C + c0 + tid] = sacc;
    }
}

template <int K, int S, int TW, int TH>
static int launch_f32(dfd_ctx* ctx, const EffBlock& b, const float* in, const float* W, const float* bias, float* out, int m,
                      int* n_parts, cudaStream_t st, int img0) {
    constexpr int PH = (TH - 1) * S + K, PW = (TW - 1) * S + K;
    const size_t smem = ((size_t)PH * PW * 32 + (size_t)K * K * 32 + (size_t)DWF_WARPS * 32) * 4 + 128;
    DFD_REQUIRE(b.hout % TW == 0, DFD_ERR_INVALID, "dw_f32: tile width must divide the output width");
    { int rc = dfd_func_smem(ctx, k_dw_tile_f32<K, S, TW, TH>, smem); if (rc) return rc; }
    const int tiles_x = b.hout / TW, tiles_y = (b.hout + TH - 1) / TH;
    dim3 grid(tiles_x * tiles_y, (b.cexp + 31) / 32, m);
    *n_parts = tiles_x * tiles_y;
    if ((size_t)grid.x * b.cexp > DFD_POOL_FLOATS) { ctx->err = "internal: squeeze partial buffer too small"; return DFD_ERR_CAPACITY; }
    CUtensorMap mp;
    {
        const uint64_t dims[4] = {(uint64_t)b.cexp, (uint64_t)b.hin, (uint64_t)b.hin, (uint64_t)m};
        const uint64_t str[3] = {(uint64_t)b.cexp * 4, (uint64_t)b.hin * b.cexp * 4, (uint64_t)b.hin * b.hin * b.cexp * 4};
        const uint32_t box[4] = {32, (uint32_t)PW, (uint32_t)PH, 1};
        int rc = dfd_tmap_encode(ctx, &mp, 1, in, 4, dims, str, box, 0);
        if (rc) return rc;
    }
    // img0: first image of a sub-batch (in / out already point at it): its squeeze partials go to the images' own slots
    k_dw_tile_f32<K, S, TW, TH><<<grid, DWF_WARPS * 32, smem, st>>>(mp, W, bias, out, ctx->d_pool + (size_t)img0 * grid.x * b.cexp, b.cexp,
                                                                    b.hout, b.pad, tiles_x);
    DFD_LAUNCH_CHECK("k_dw_tile_f32", st);
    return DFD_OK;
}

int dfd_dw_f32(dfd_ctx* ctx, const EffBlock& b, const float* in, const float* W, const float* bias, float* out, int m,
               int* n_parts, cudaStream_t st, int img0) {
    if (b.k == 3 && b.s == 1 && b.hout == 112) return launch_f32<3, 1, 16, 16>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    if (b.k == 3 && b.s == 2 && b.hout == 56) return launch_f32<3, 2, 14, 8>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    if (b.k == 3 && b.s == 1 && b.hout == 56) return launch_f32<3, 1, 14, 16>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    if (b.k == 5 && b.s == 2 && b.hout == 28) return launch_f32<5, 2, 14, 7>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    if (b.k == 5 && b.s == 1 && b.hout == 28) return launch_f32<5, 1, 14, 14>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    if (b.k == 3 && b.s == 2 && b.hout == 14) return launch_f32<3, 2, 14, 7>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    if (b.k == 3 && b.s == 1 && b.hout == 14) return launch_f32<3, 1, 14, 14>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    if (b.k == 5 && b.s == 1 && b.hout == 14) return launch_f32<5, 1, 14, 14>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    if (b.k == 5 && b.s == 2 && b.hout == 7) return launch_f32<5, 2, 7, 7>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    if (b.k == 5 && b.s == 1 && b.hout == 7) return launch_f32<5, 1, 7, 7>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    if (b.k == 3 && b.s == 1 && b.hout == 7) return launch_f32<3, 1, 7, 7>(ctx, b, in, W, bias, out, m, n_parts, st, img0);
    ctx->err = "dw_f32: no tile configuration for this layer";
    return DFD_ERR_INVALID;
}
