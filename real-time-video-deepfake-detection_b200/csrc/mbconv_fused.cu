// Fused front half of an MBConv block for bf16 NHWC activations (reference model.py:63-72 ->
// lukemelas MBConvBlock: _expand_conv/_bn0/swish -> _depthwise_conv/_bn1/swish -> avg-pool of the SE squeeze):
//
//   x [m][hin][hin][cin]  --1x1 expand (tcgen05, fp32 accum in TMEM) + bias + swish-->  e (bf16, SHARED MEMORY ONLY)
//                         --depthwise kxk stride s (fp32 FMA) + bias + swish-->          out [m][hout][hout][cexp]
//                                                                                         + SE squeeze partials
//
// The expanded tensor is 6x the block input and, layer by layer, was written once and read once: 14.4 MB of the
// 27.4 MB/img layer-granular traffic.  Here it never leaves the SM: a CTA owns (image, TH x TW output tile,
// CC-channel chunk), fetches the input patch with ONE 4-D TMA box per 64 input channels (zero fill outside the
// image), multiplies it by the chunk's expand weights with tcgen05.mma (M = 128 patch pixels per instruction,
// N = CC, K = cin), reads the accumulators back with tcgen05.ld, applies bias + swish, forces TF-"SAME" padding
// pixels to zero (the reference pads the EXPANDED tensor) and leaves the bf16 patch in shared memory, where the
// depthwise stage of dwconv_bf16.cu runs on it unchanged (lane = channel pair, warp = output row).
// The halo is recomputed per tile; results are bit-identical to the unfused expand GEMM + k_dw_tile pair.
//
// WHOLE = the tile is the whole output image (14x14 and 7x7 stages): only the hin x hin real pixels go through
// the MMA and the padding ring of the patch is zero-filled directly.
#include "dfd_internal.cuh"
#include "effnet_plan.h"
#include "tc_ptx.cuh"
#include "se_tail.cuh"

#define MF_THREADS 256
#define MF_WARPS 8
#define MF_PITCH_W 36            // 32-bit words per patch pixel: 128 B of channels + 16 B pad (conflict-free 16-byte stores)

int dfd_tmap_bf16(dfd_ctx* ctx, CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

struct FrontParams {
    const float* be;             // expand bias [C]
    const float* Wd;             // depthwise weights [K*K][C]
    const float* bd;             // depthwise bias [C]
    __nv_bfloat16* out;          // [m][hout][hout][C]
    float* pool;                 // [m][tiles][C] SE squeeze partials
    int C, cin, hin, hout, pad, tiles_x, num_kb;
    SeTail se;                   // SE excite run by the last CTA of each image
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void mf_ffma2(uint64_t& d, uint64_t a, uint64_t b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ uint64_t mf_pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void mf_unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint32_t mf_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *(uint32_t*)&h;
}

constexpr int mf_pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int K, int S, int TW, int TH, int CC, int HIN, bool WHOLE>
struct MfGeom {
    static constexpr int PH = (TH - 1) * S + K, PW = (TW - 1) * S + K;
    static constexpr int NPIX = PH * PW;                         // patch pixels
    static constexpr int BOX_W = WHOLE ? HIN : PW, BOX_H = WHOLE ? HIN : PH;
    static constexpr int ROWS = BOX_W * BOX_H;                   // MMA rows that carry data
    static constexpr int N_MB = (ROWS + 127) / 128;
    static constexpr int TM_COLS = mf_pow2_cols(N_MB * CC);
    static constexpr int A_KB_BYTES = N_MB * 128 * 128;          // one 64-channel k-block of the A operand
    static constexpr int B_KB_BYTES = CC * 128;
    static constexpr int PATCH_BYTES = NPIX * MF_PITCH_W * 4;
    static constexpr int TAIL_FLOATS = K * K * CC + CC + MF_WARPS * CC;    // dw weights, expand bias, squeeze partials
    static int region_bytes(int num_kb) {
        int a = num_kb * A_KB_BYTES;
        int r = a > PATCH_BYTES ? a : PATCH_BYTES;
        return (r + 1023) & ~1023;
    }
    static size_t smem_bytes(int num_kb) { return 1024 + (size_t)region_bytes(num_kb) + (size_t)num_kb * B_KB_BYTES + (size_t)TAIL_FLOATS * 4; }
};

template <int K, int S, int TW, int TH, int CC, int HIN, bool WHOLE>
__global__ void __launch_bounds__(MF_THREADS)
k_mbconv_front(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const FrontParams p) {
    using G = MfGeom<K, S, TW, TH, CC, HIN, WHOLE>;
    constexpr int PW = G::PW, NPIX = G::NPIX, ROWS = G::ROWS, N_MB = G::N_MB;
    extern __shared__ __align__(1024) uint8_t smem_mf[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, chunk = blockIdx.y, b = blockIdx.z;
    const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    const int oy0 = ty * TH, ox0 = tx * TW;
    const int c0 = chunk * CC;
    const int iy0 = oy0 * S - p.pad, ix0 = ox0 * S - p.pad;

    const uint32_t region = (smem_u32(smem_mf) + 1023u) & ~1023u;
    int rb = p.num_kb * G::A_KB_BYTES; if (rb < G::PATCH_BYTES) rb = G::PATCH_BYTES; rb = (rb + 1023) & ~1023;
    const uint32_t bsm = region + (uint32_t)rb;
    uint8_t* gen_base = smem_mf + (region - smem_u32(smem_mf));
    uint32_t* patch = (uint32_t*)gen_base;                                  // aliases the A operand once the MMAs have retired
    float* sw = (float*)(gen_base + rb + p.num_kb * G::B_KB_BYTES);         // [K*K][CC]
    float* sbe = sw + K * K * CC;                                           // [CC]
    float* spool = sbe + CC;                                                // [MF_WARPS][CC]
    const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

    if (tid == 0) {                                  // the input patch + weight chunk are requested before anything else
        mbar_init(bar_tma, 1); mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_tma, (uint32_t)p.num_kb * (uint32_t)(ROWS * 128 + G::B_KB_BYTES));
        for (int kb = 0; kb < p.num_kb; kb++) {
            tma_load_4d(region + kb * G::A_KB_BYTES, &map_x, kb * 64, WHOLE ? 0 : ix0, WHOLE ? 0 : iy0, b, bar_tma);
            tma_load_2d(bsm + kb * G::B_KB_BYTES, &map_w, kb * 64, c0, bar_tma);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(G::TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < K * K * CC; i += MF_THREADS) {
        const int c = c0 + (i % CC);
        sw[i] = c < p.C ? p.Wd[(size_t)(i / CC) * p.C + c] : 0.f;
    }
    if (tid < CC) sbe[tid] = c0 + tid < p.C ? p.be[c0 + tid] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    // ---- one thread: wait for the TMA data, then issue every MMA of the tile ----
    if (warp == 0) {
        if (lane == 0) {
            mbar_wait(bar_tma, 0);
            tc_fence_after();
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#pragma unroll 1
            for (int mb = 0; mb < N_MB; mb++) {
                for (int kb = 0; kb < p.num_kb; kb++) {
                    const uint64_t adesc = make_smem_desc(region + kb * G::A_KB_BYTES + mb * 16384);
                    const uint64_t bdesc = make_smem_desc(bsm + kb * G::B_KB_BYTES);
                    const int krem = p.cin - kb * 64;
                    const int ksteps = krem >= 64 ? 4 : (krem + 15) / 16;
                    for (int k = 0; k < ksteps; k++)
                        tc_mma_bf16(tmem_base + (uint32_t)(mb * CC), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                }
            }
            tc_commit(bar_mma);
        }
        __syncwarp();
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after();

    // ---- TMEM -> bias + swish -> bf16 patch in shared memory (padding pixels = 0) ----
    const uint32_t patch_s = region;
    if (WHOLE) {                                     // padding ring of the patch
        for (int pix = tid; pix < NPIX; pix += MF_THREADS) {
            const int py = pix / PW, px = pix - py * PW;
            const int iy = py - p.pad, ix = px - p.pad;
            if (iy < 0 || iy >= HIN || ix < 0 || ix >= HIN) {
#pragma unroll
                for (int j = 0; j < CC / 8; j++) sts128(patch_s + (uint32_t)(pix * (MF_PITCH_W * 4) + j * 16), make_uint4(0, 0, 0, 0));
            }
        }
    }
    {
        const int q = warp & 3, pair = warp >> 2;
        constexpr int NG = CC / 16;
        for (int u = pair; u < N_MB * NG; u += 2) {
            const int mb = u / NG, g = u - mb * NG;
            if (mb * 128 + q * 32 >= ROWS) continue;                  // warp-uniform: no data rows in this lane quarter
            const int r = mb * 128 + q * 32 + lane;
            uint32_t v[16];
            tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * CC + g * 16), v);
            tc_ld_wait();
            bool inimg; int pp;
            if (WHOLE) {
                const int iy = r / HIN, ix = r - iy * HIN;
                inimg = true;
                pp = (iy + p.pad) * PW + ix + p.pad;
            } else {
                const int py = r / PW, px = r - py * PW;
                const int iy = iy0 + py, ix = ix0 + px;
                inimg = iy >= 0 && iy < p.hin && ix >= 0 && ix < p.hin;
                pp = r;
            }
            if (r < ROWS) {
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float2 bb = *(const float2*)(sbe + g * 16 + 2 * j);
                    const float a0 = swish_fast(__uint_as_float(v[2 * j]) + bb.x);
                    const float a1 = swish_fast(__uint_as_float(v[2 * j + 1]) + bb.y);
                    o[j] = inimg ? mf_bf16x2(a0, a1) : 0u;
                }
                const uint32_t dst = patch_s + (uint32_t)(pp * (MF_PITCH_W * 4) + g * 32);
                sts128(dst, make_uint4(o[0], o[1], o[2], o[3]));
                sts128(dst + 16, make_uint4(o[4], o[5], o[6], o[7]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {                                 // the accumulators are drained: free TMEM for the next CTA on this SM
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G::TM_COLS) : "memory");
    }

    // ---- depthwise kxk + bias + swish + SE squeeze partials (lane = channel pair, warp = output row) ----
    const int pl = lane < CC / 2 ? lane : CC / 2 - 1;
    const int ch = c0 + 2 * pl;
    const bool ch_ok = lane < CC / 2 && ch < p.C;
    const uint64_t bias2 = ch_ok ? mf_pack2(p.bd[ch], p.bd[ch + 1]) : mf_pack2(0.f, 0.f);
    float ps0 = 0.f, ps1 = 0.f;
    for (int r = warp; r < TH; r += MF_WARPS) {
        const int oy = oy0 + r;
        if (oy >= p.hout) break;
        uint64_t acc[TW];
#pragma unroll
        for (int i = 0; i < TW; i++) acc[i] = bias2;
#pragma unroll
        for (int ky = 0; ky < K; ky++) {
            uint64_t w[K];
#pragma unroll
            for (int kx = 0; kx < K; kx++) w[kx] = *(const uint64_t*)(sw + (ky * K + kx) * CC + 2 * pl);
            const uint32_t* prow = patch + (size_t)((r * S + ky) * PW) * MF_PITCH_W + pl;
#pragma unroll
            for (int ix = 0; ix < PW; ix++) {
                const uint32_t v = prow[ix * MF_PITCH_W];
                const uint64_t x = mf_pack2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
#pragma unroll
                for (int kx = 0; kx < K; kx++)
                    if ((ix - kx) % S == 0 && (ix - kx) >= 0 && (ix - kx) / S < TW) mf_ffma2(acc[(ix - kx) / S], x, w[kx]);
            }
        }
        __nv_bfloat16* orow = p.out + (((size_t)b * p.hout + oy) * p.hout + ox0) * p.C + ch;
#pragma unroll
        for (int i = 0; i < TW; i++) {
            if (ox0 + i < p.hout && ch_ok) {
                float y0, y1;
                mf_unpack2(acc[i], y0, y1);
                y0 = swish_fast(y0); y1 = swish_fast(y1);
                ps0 += y0; ps1 += y1;
                *(__nv_bfloat162*)(orow + (size_t)i * p.C) = __floats2bfloat162_rn(y0, y1);
            }
        }
    }
    if (lane < CC / 2) { spool[warp * CC + 2 * lane] = ps0; spool[warp * CC + 2 * lane + 1] = ps1; }
    __syncthreads();
    if (tid < CC && c0 + tid < p.C) {                // deterministic: fixed-order sum, one partial per (image, tile, channel)
        float sacc = 0.f;
#pragma unroll
        for (int wv = 0; wv < MF_WARPS; wv++) sacc += spool[wv * CC + tid];
        p.pool[((size_t)b * gridDim.x + tile) * p.C + c0 + tid] = sacc;
    }
    if (p.se.counter) se_tail_run(p.se, p.pool, gridDim.x, p.C, b, (float*)patch);
}

template <int K, int S, int TW, int TH, int CC, int HIN, bool WHOLE>
static int launch_front(dfd_ctx* ctx, const EffBlock& b, const __nv_bfloat16* x, const __nv_bfloat16* We, const float* be,
                        const float* Wd, const float* bd, __nv_bfloat16* out, int m, int* n_parts, const SeTail& se, cudaStream_t st) {
    using G = MfGeom<K, S, TW, TH, CC, HIN, WHOLE>;
    const int num_kb = (b.cin + 63) / 64;
    const size_t smem = G::smem_bytes(num_kb);
    static size_t attr_bytes = 0;
    if (smem > attr_bytes) {
        DFD_CUDA(cudaFuncSetAttribute(k_mbconv_front<K, S, TW, TH, CC, HIN, WHOLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_bytes = smem;
    }
    CUtensorMap mx, mw;
    int rc;
    {
        const uint64_t dims[4] = {(uint64_t)b.cin, (uint64_t)b.hin, (uint64_t)b.hin, (uint64_t)m};
        const uint64_t str[3] = {(uint64_t)b.cin * 2, (uint64_t)b.hin * b.cin * 2, (uint64_t)b.hin * b.hin * b.cin * 2};
        const uint32_t box[4] = {64, (uint32_t)G::BOX_W, (uint32_t)G::BOX_H, 1};
        if ((rc = dfd_tmap_bf16(ctx, &mx, x, 4, dims, str, box))) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)b.cin, (uint64_t)b.cexp};
        const uint64_t str[1] = {(uint64_t)b.cin * 2};
        const uint32_t box[2] = {64, (uint32_t)CC};
        if ((rc = dfd_tmap_bf16(ctx, &mw, We, 2, dims, str, box))) return rc;
    }
    FrontParams p;
    p.be = be; p.Wd = Wd; p.bd = bd; p.out = out; p.pool = ctx->d_pool;
    p.C = b.cexp; p.cin = b.cin; p.hin = b.hin; p.hout = b.hout; p.pad = b.pad; p.num_kb = num_kb;
    const int tiles_x = (b.hout + TW - 1) / TW, tiles_y = (b.hout + TH - 1) / TH;
    p.tiles_x = tiles_x;
    dim3 grid(tiles_x * tiles_y, (b.cexp + CC - 1) / CC, m);
    *n_parts = tiles_x * tiles_y;
    if ((size_t)grid.x * b.cexp > DFD_POOL_FLOATS) { ctx->err = "internal: squeeze partial buffer too small"; return DFD_ERR_CAPACITY; }
    p.se = se; p.se.ctas_per_image = (int)(grid.x * grid.y);
    k_mbconv_front<K, S, TW, TH, CC, HIN, WHOLE><<<grid, MF_THREADS, smem, st>>>(mx, mw, p);
    DFD_LAUNCH_CHECK("k_mbconv_front", st);
    return DFD_OK;
}

// Expand 1x1 + depthwise of block `b` (cexp != cin) in one kernel.  x: block input, We: bf16 expand weights [cexp][cin].
int dfd_mbconv_front_bf16(dfd_ctx* ctx, const EffBlock& b, const __nv_bfloat16* x, const __nv_bfloat16* We, const float* be,
                          const float* Wd, const float* bd, __nv_bfloat16* out, int m, int* n_parts, const SeTail& se, cudaStream_t st) {
#define MF_ARGS ctx, b, x, We, be, Wd, bd, out, m, n_parts, se, st
    if (b.k == 3 && b.s == 2 && b.hin == 112) return launch_front<3, 2, 7, 8, 48, 112, false>(MF_ARGS);
    if (b.k == 3 && b.s == 1 && b.hin == 56) return launch_front<3, 1, 14, 14, 48, 56, false>(MF_ARGS);
    if (b.k == 5 && b.s == 2 && b.hin == 56) return launch_front<5, 2, 7, 7, 48, 56, false>(MF_ARGS);
    if (b.k == 5 && b.s == 1 && b.hin == 28) return launch_front<5, 1, 14, 14, 64, 28, false>(MF_ARGS);
    if (b.k == 3 && b.s == 2 && b.hin == 28) return launch_front<3, 2, 7, 7, 64, 28, false>(MF_ARGS);
    if (b.k == 3 && b.s == 1 && b.hin == 14) return launch_front<3, 1, 14, 14, 64, 14, true>(MF_ARGS);
    if (b.k == 5 && b.s == 1 && b.hin == 14) return launch_front<5, 1, 14, 14, 64, 14, true>(MF_ARGS);
    if (b.k == 5 && b.s == 2 && b.hin == 14) return launch_front<5, 2, 7, 7, 64, 14, true>(MF_ARGS);
    if (b.k == 5 && b.s == 1 && b.hin == 7) return launch_front<5, 1, 7, 7, 64, 7, true>(MF_ARGS);
    if (b.k == 3 && b.s == 1 && b.hin == 7) return launch_front<3, 1, 7, 7, 64, 7, true>(MF_ARGS);
#undef MF_ARGS
    ctx->err = "mbconv_front: no tile configuration for this layer";
    return DFD_ERR_INVALID;
}
