// bf16 convolution-as-GEMM on the 5th-generation tensor cores (sm_100a).
//
//   C[M,N] = act( A'[M,K] . W[N,K]^T + bias[N] ) (+ residual[M,N])     A, W, C, residual bf16; accumulate fp32 in TMEM
//
// A is the NHWC activation matrix (M = batch*H*W rows, K-major) and W the BN-folded weight matrix [N][K]
// (K-major) of a convolution of the reference's EfficientNet-B0 (model.py:63-72):
//   A_TMA    expand / head 1x1 convs: A' = A, tiles fetched by TMA
//   A_SCALE  project 1x1 convs: A' = A * se[image][k] -- the squeeze-excite gate is applied while the
//            tile is staged, so the gated tensor never exists in HBM
//   A_STEM   stem 3x3 stride-2 conv 3->32: A' = im2col rows (27 taps, zero padded to 32) gathered on the fly
//   A_IMG    project 1x1 convs of the high-resolution blocks: the SE gate is folded into a PER-IMAGE weight matrix
//            W_b = W * diag(g_b) (written by the SE kernel), tiles never straddle images (3-D tensor maps
//            [image][row][k]) and A goes from TMA straight to the tensor core -- no staging pass over A at all
//
// One persistent CTA per SM, warp-specialised (640 threads), no cluster:
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads (SWIZZLE_128B) of the W tile (N x 64) and (A_TMA,
//               A_SCALE) the A tile (128 x 64) into a multi-stage smem ring; out-of-bounds rows/columns are
//               zero-filled by TMA so ragged M, K = 16..1152 and N = 16..256 need no padding copies.
//   warps 12-19 two groups of 4 staging warps working on alternate k-blocks:
//               A_SCALE: wait for the raw TMA tile, multiply it in place in shared memory by the SE gates (gates fetched
//                        one k-block ahead, split into bf16 hi + lo and applied as fma(v, hi, v * lo) on packed bf16 pairs;
//                        ld.shared -> 2 HFMA2.BF16 per pair -> st.shared at the swizzled address), fence.proxy.async;
//               A_STEM:  gather the im2col row of each output pixel (15 aligned 4-byte loads) into the same
//                        128-byte-swizzled K-major layout, fence.proxy.async.
//   warp 1      MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N, K=16) per K step
//               from smem descriptors; tcgen05.commit frees the smem stage and publishes the accumulator.
//   warp 2      allocates / frees TMEM (two accumulators of N fp32 columns).
//   warps 4-7 / 8-11  two epilogue sets in ping-pong (set s owns accumulator s and the tiles of its parity):
//               tcgen05.ld (32 lanes x 32 columns) -> + bias (smem), swish (one MUFU tanh), + residual (prefetched one
//               TMEM read ahead; the kernel is instantiated per <residual, swish, dense-store> variant) -> bf16 ->
//               128-byte-swizzled smem staging (two 16 KB buffers per set) -> TMA tensor store of each
//               64-column block (clips ragged M / N), so every global write is a full 128-byte line and the
//               latency of one tile's epilogue hides behind the other set's.
// These layers are HBM-bound (arithmetic intensity 10-250 flop/B): the design streams A once and writes C
// once at full line width; the tensor pipe has >10x headroom.
#include "dfd_internal.cuh"
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "tc_ptx.cuh"

#define BLOCK_M 128
#define BLOCK_K 64
#define GEMM_THREADS 640
#define A_STAGE_BYTES (BLOCK_M * BLOCK_K * 2)
#define STAGING_BLOCK_BYTES (BLOCK_M * 128)
#define MAX_BIAS 1280

enum { A_TMA = 0, A_SCALE = 1, A_STEM = 2, A_IMG = 3 };

struct GemmParams {
    int M, N, K;
    int n_pad;            // UMMA N (multiple of 16, <= 256)
    int n_blocks;         // ceil(N / n_pad)
    int num_tiles;        // m_blocks * n_blocks
    int stages;
    int act;
    int a_mode;
    int hw;               // A_SCALE / A_IMG: rows per image
    int tiles_per_img;    // A_IMG: ceil(hw / 128)
    int b_resident;       // W (all k-blocks) stays in shared memory for the whole kernel; the ring holds A only
    int debug;            // diagnostics (dfd_gemm_bench): bit 0 = skip the TMA stores, bit 1 = skip the epilogue math
    const float* bias;
    const __nv_bfloat16* residual;
    const __nv_bfloat16* A;    // A_STEM: NHWC input [B,224,224,3]
    __nv_bfloat16* C;          // output matrix (dense epilogue: the 128 x N tile is one contiguous block of C)
    int n_acc;                 // TMEM accumulator ring depth (even, 2..8)
    int epi_db;                // staging modes: 1 = two store-staging buffers per epilogue set, 0 = one (frees 32 KB for pipeline stages)
    int dense_c;               // N <= 64 and one N block: stage the tile densely and write it with ONE bulk copy
    const float* se;           // A_SCALE: [images][K] gates
};

// RES: the layer has a residual input (its epilogue prefetches the residual one TMEM read ahead; layers without one keep the
// shorter epilogue -- the narrow early layers are bound by per-tile epilogue latency and pay for every extra instruction)
// ACT / DENSE: swish epilogue / dense bulk-store epilogue (p.act, p.dense_c) resolved at compile time for the same reason.
template <bool RES, bool ACT, bool DENSE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tcgen05(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[5 * 8 + 1];     // full[8], empty[8], raw[8], tmem_full[8], tmem_empty[8], bfull
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float sbias[MAX_BIAS];
    __shared__ __align__(16) float sgate[8][4][64];       // A_SCALE: per staging warp, the current k-block's gates of the tile's <= 4 images

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool plain = p.a_mode == A_TMA || p.a_mode == A_IMG;     // A goes from TMA straight to the MMA
    const uint32_t b_stage_bytes = (uint32_t)p.n_pad * BLOCK_K * 2;
    const uint32_t stage_bytes = A_STAGE_BYTES + (p.b_resident ? 0u : b_stage_bytes);
    const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
    const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    const uint32_t b_region = smem_base + (uint32_t)p.stages * stage_bytes;    // resident W: num_kb x (n_pad x 128 B)
    const uint32_t staging = b_region + (p.b_resident ? (uint32_t)num_kb * b_stage_bytes : 0u);     // 4 x 16 KB
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), raw0 = smem_u32(&bars[16]);
    const uint32_t tfull0 = smem_u32(&bars[24]), tempty0 = smem_u32(&bars[32]), bfull = smem_u32(&bars[40]);
    // two accumulators of n_pad columns; the epilogue reads 32 columns at a time, so the last read of the second
    // accumulator may extend to the next multiple of 32 -- it must stay inside the allocation
    // p.n_acc accumulators of n_pad columns (a ring: the MMA issuer runs up to n_acc tiles ahead of the epilogue, which
    // hides the commit -> wait -> tcgen05.ld -> arrive round trip of the accumulator hand-over; with only two accumulators
    // that round trip, ~1 us, capped narrow layers at two tiles per trip).  The epilogue reads 32 columns at a time, so
    // the last read of the last accumulator may extend to the next multiple of 32 -- it must stay inside the allocation.
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(p.n_acc - 1) * (uint32_t)p.n_pad + (((uint32_t)p.n_pad + 31u) & ~31u)) tmem_cols <<= 1;

    for (int i = threadIdx.x; i < p.n_pad * p.n_blocks && i < MAX_BIAS; i += GEMM_THREADS) sbias[i] = p.bias[i];
    if (warp == 0 && lane == 0) {
        if (p.a_mode != A_STEM) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1 && lane == 0) {
        // full[s]: A_TMA = the TMA thread; A_SCALE = 4 fix-up warps; A_STEM = TMA thread (W) + 4 gather warps
        const uint32_t full_count = plain ? 1u : (p.a_mode == A_SCALE ? 4u : (p.b_resident ? 4u : 5u));
        mbar_init(bfull, 1);
        for (int s = 0; s < p.stages; s++) {
            mbar_init(full0 + 8 * s, full_count); mbar_init(empty0 + 8 * s, 1); mbar_init(raw0 + 8 * s, 1);
        }
        // A_TMA: the 8 staging warps double as a second pair of epilogue groups (column halves) -> 8 arrivals
        for (int a = 0; a < p.n_acc; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, plain ? 8 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    // PDL: everything above (and the resident weight load below: static data) overlaps the predecessor's tail
    if (warp == 0 && lane == 0 && p.b_resident) {                    // every tile reuses the same W: load it once
        mbar_expect_tx(bfull, (uint32_t)num_kb * b_stage_bytes);
        for (int kb = 0; kb < num_kb; kb++) tma_load_2d(b_region + kb * b_stage_bytes, &map_b, kb * BLOCK_K, 0, bfull);
    }
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const bool load_a = p.a_mode != A_STEM, load_b = !p.b_resident;
            const uint32_t tx = (load_a ? A_STAGE_BYTES : 0u) + (load_b ? b_stage_bytes : 0u);
            const uint32_t sig0 = p.a_mode == A_SCALE ? raw0 : full0;    // A_SCALE: the fix-up warps publish full[]
            if (tx != 0) {
                for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                    const int m_blk = tile / p.n_blocks, n_blk = tile % p.n_blocks;
                    for (int kb = 0; kb < num_kb; kb++) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + A_STAGE_BYTES;
                        const bool skip_b = (p.debug & 32) != 0;      // diagnostics only
                        mbar_expect_tx(sig0 + 8 * stage, skip_b ? tx - b_stage_bytes : tx);
                        if (p.a_mode == A_IMG) {
                            const int img = tile / p.tiles_per_img, mb = tile - img * p.tiles_per_img;
                            tma_load_3d(sa, &map_a, kb * BLOCK_K, mb * BLOCK_M, img, sig0 + 8 * stage);
                            tma_load_3d(sb, &map_b, kb * BLOCK_K, 0, img, sig0 + 8 * stage);
                        } else {
                            if (load_a) tma_load_2d(sa, &map_a, kb * BLOCK_K, m_blk * BLOCK_M, sig0 + 8 * stage);
                            if (load_b && !skip_b) tma_load_2d(sb, &map_b, kb * BLOCK_K, n_blk * p.n_pad, sig0 + 8 * stage);
                        }
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            if (p.b_resident) mbar_wait(bfull, 0);
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_pad);
                for (int kb = 0; kb < num_kb; kb++) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint32_t sb = p.b_resident ? b_region + kb * b_stage_bytes : sa + A_STAGE_BYTES;
                    const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
                    const int krem = p.K - kb * BLOCK_K;
                    const int ksteps = krem >= BLOCK_K ? BLOCK_K / 16 : (krem + 15) / 16;
                    for (int k = 0; k < ksteps; k++)
                        tc_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    tc_commit(empty0 + 8 * stage);                 // frees the smem stage when the MMAs retire
                    if (kb == num_kb - 1) tc_commit(tfull0 + 8 * acc);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (++acc == p.n_acc) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 12 && !plain) {
        // ===== staging warps: two groups of 128 threads on alternate k-blocks =====
        {
            const int g = (warp - 12) >> 2;
            const int t = threadIdx.x - (12 + 4 * g) * 32;         // 0..127
            if (p.a_mode == A_STEM) {
                // stem im2col (one k-block per tile): row = output pixel; 3 kernel rows x 9 contiguous bf16 (3 px x 3 ch) -> 27 taps
                // + 5 zeros.  Software-pipelined: the 15 loads of the group's NEXT tile are in flight while it waits for the
                // smem slot of the current one and writes it, so the gather latency is paid once, not per tile.
                auto gather = [&](int tile, uint32_t (&a)[3][5]) {
                    const int m = tile * BLOCK_M + t;
#pragma unroll
                    for (int ky = 0; ky < 3; ky++)
#pragma unroll
                        for (int i = 0; i < 5; i++) a[ky][i] = 0;
                    if (m < p.M) {
                        const int ox = m % 112, oy = (m / 112) % 112, b = m / (112 * 112);
#pragma unroll
                        for (int ky = 0; ky < 3; ky++) {
                            const int iy = 2 * oy + ky;
                            if (iy < 224) {
                                const uint32_t* src = (const uint32_t*)(p.A + (((size_t)b * 224 + iy) * 224 + 2 * ox) * 3);
                                a[ky][0] = __ldg(src); a[ky][1] = __ldg(src + 1); a[ky][2] = __ldg(src + 2);
                                if (ox < 111) { a[ky][3] = __ldg(src + 3); a[ky][4] = __ldg(src + 4) & 0xffffu; }   // third pixel is padding at the right edge
                            }
                        }
                    }
                };
                uint32_t cur[3][5], nxt[3][5];
                int it = g;
                int tile = blockIdx.x + it * (int)gridDim.x;
                if (tile < p.num_tiles) gather(tile, cur);
                for (; tile < p.num_tiles; it += 2) {
                    const int ntile = blockIdx.x + (it + 2) * (int)gridDim.x;
                    if (ntile < p.num_tiles) gather(ntile, nxt);
                    const int stage = it % p.stages;
                    const uint32_t phase = (uint32_t)(it / p.stages) & 1u;
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    uint32_t w[16];
                    w[0] = cur[0][0]; w[1] = cur[0][1]; w[2] = cur[0][2]; w[3] = cur[0][3];
                    w[4] = cur[0][4] | (cur[1][0] << 16);
                    w[5] = (cur[1][0] >> 16) | (cur[1][1] << 16);
                    w[6] = (cur[1][1] >> 16) | (cur[1][2] << 16);
                    w[7] = (cur[1][2] >> 16) | (cur[1][3] << 16);
                    w[8] = (cur[1][3] >> 16) | (cur[1][4] << 16);
                    w[9] = cur[2][0]; w[10] = cur[2][1]; w[11] = cur[2][2]; w[12] = cur[2][3]; w[13] = cur[2][4];
                    w[14] = 0; w[15] = 0;
                    const int row = t;
#pragma unroll
                    for (int c = 0; c < 4; c++)
                        sts128(sa + (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4)), make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full0 + 8 * stage);
#pragma unroll
                    for (int ky = 0; ky < 3; ky++)
#pragma unroll
                        for (int i = 0; i < 5; i++) cur[ky][i] = nxt[ky][i];
                    tile = ntile;
                }
            } else {
                // A_SCALE.  The gates of a k-block (at most 4 images x 64 channels for one 128-row tile, hw >= 49) are fetched one
                // k-block AHEAD into registers (lane = image x 8-channel chunk), then published to the warp through a 1 KB
                // shared-memory table, so the L2 latency of the gate loads is not paid between "raw tile landed" and
                // "tile ready for the tensor core" (it was: twice per k-block, and it bounded the deep-K 7x7 / 14x14 layers).
                // The multiply itself is two packed bf16 instructions per pair: g = hi + lo (both bf16, |lo| <= 2^-9 |g|), and
                // v * g = fma(v, hi, v * lo) -- the inner product is a 2^-9-relative correction term, the fma rounds once, so
                // the result is the correctly rounded bf16 of v * g up to ~2^-17 relative (unpack / fp32 multiply / re-pack was
                // 5 instructions per pair and made this fix-up, not the tensor core or the loads, the pace of the layer).
                uint4* wg_hi = (uint4*)&sgate[warp - 12][0][0];       // [4 images][8 chunks] x 8 bf16
                uint4* wg_lo = wg_hi + 32;
                const int c = t & 7;
                const int gl_img = lane >> 3, gl_c = lane & 7;
                const int n_img = (p.M + p.hw - 1) / p.hw;
                int tile = blockIdx.x, kb = g;                        // this group's next k-block (p.stages is even: stage parity == group)
                auto norm = [&]() { while (kb >= num_kb && tile < p.num_tiles) { kb -= num_kb; tile += gridDim.x; } };
                norm();
                float4 q0, q1;
                auto prefetch = [&]() {
                    q0 = make_float4(0.f, 0.f, 0.f, 0.f); q1 = q0;
                    if (tile < p.num_tiles) {
                        int img = ((tile / p.n_blocks) * BLOCK_M) / p.hw + gl_img;
                        if (img > n_img - 1) img = n_img - 1;
                        const int k = kb * BLOCK_K + gl_c * 8;
                        if (k < p.K) {
                            const float* sp = p.se + (size_t)img * p.K + k;
                            q0 = __ldg((const float4*)sp); q1 = __ldg((const float4*)(sp + 4));
                        }
                    }
                };
                auto split = [](float a, float b, uint32_t& hi, uint32_t& lo) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                    const float2 hf = __bfloat1622float2(h);
                    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
                    hi = *(const uint32_t*)&h; lo = *(const uint32_t*)&l;
                };
                prefetch();
                for (int j = g; tile < p.num_tiles; j += 2) {
                    const int m0 = (tile / p.n_blocks) * BLOCK_M;
                    const int stage = j % p.stages;
                    const uint32_t phase = (uint32_t)(j / p.stages) & 1u;
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    {
                        uint4 h, l;
                        split(q0.x, q0.y, h.x, l.x); split(q0.z, q0.w, h.y, l.y);
                        split(q1.x, q1.y, h.z, l.z); split(q1.z, q1.w, h.w, l.w);
                        wg_hi[gl_img * 8 + gl_c] = h; wg_lo[gl_img * 8 + gl_c] = l;
                    }
                    __syncwarp();
                    const int k = kb * BLOCK_K + c * 8;
                    kb += 2;
                    norm();
                    prefetch();                                    // next k-block's gates fly during the wait and the fix-up
                    mbar_wait(raw0 + 8 * stage, phase);            // raw A (and W) tile landed
                    if (k < p.K) {
                        // image of a row without a division per row: one division per tile, then the (at most 3, hw >= 49)
                        // image boundaries inside the 128-row tile by comparison
                        const int img0 = m0 / p.hw, rem0 = m0 - img0 * p.hw;
                        const int last = p.M - 1 - m0;                         // rows beyond M are clamped to the last row
                        uint4 v[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int row = i * 16 + (t >> 3);
                            v[i] = lds128(sa + (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4)));
                        }
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int row = i * 16 + (t >> 3);
                            const int rr = rem0 + (row < last ? row : last);
                            const int rel = (rr >= p.hw) + (rr >= 2 * p.hw) + (rr >= 3 * p.hw);
                            const uint4 gh = wg_hi[rel * 8 + c], gl = wg_lo[rel * 8 + c];
                            __nv_bfloat162* h = (__nv_bfloat162*)&v[i];
                            const __nv_bfloat162* ph = (const __nv_bfloat162*)&gh;
                            const __nv_bfloat162* pl = (const __nv_bfloat162*)&gl;
#pragma unroll
                            for (int e = 0; e < 4; e++) h[e] = __hfma2(h[e], ph[e], __hmul2(h[e], pl[e]));
                            sts128(sa + (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4)), v[i]);
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full0 + 8 * stage);
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue groups of 4 warps (ping-pong over the two accumulators) =====
        // A_SCALE / A_STEM: groups 0,1 (warps 4-11), each drains a whole accumulator.
        // A_TMA: groups 0-3 (warps 4-19); group g drains accumulator g&1 and the 64-column blocks of parity g>>1.
        const bool split = plain;
        const int grp = (warp - 4) >> 2;
        const int set = grp & 1, half = grp >> 1;
        const int q = warp & 3;                                    // TMEM lane quarter (= warp % 4)
        const int row = q * 32 + lane;
        const bool issuer = q == 0 && lane == 0;
        // staging: 4 x 16 KB; unsplit groups own two buffers (double-buffered), split groups own one
        const uint32_t my_staging = staging + (uint32_t)(split ? grp : (p.epi_db ? 2 * set : set)) * STAGING_BLOCK_BYTES;
        const int nblk64 = (p.n_pad + 63) >> 6;
        uint32_t blk_count = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++) {
            if ((it & 1) != set) continue;
            int m_blk = tile / p.n_blocks, n_blk = tile % p.n_blocks;
            int m = m_blk * BLOCK_M + row;
            bool row_ok = m < p.M;
            int img = 0;
            if (p.a_mode == A_IMG) {                               // tile = (image, 128-row block inside the image)
                img = tile / p.tiles_per_img; m_blk = tile - img * p.tiles_per_img; n_blk = 0;
                row_ok = m_blk * BLOCK_M + row < p.hw;
                m = img * p.hw + m_blk * BLOCK_M + row;
            }
            const int n_base = n_blk * p.n_pad;
            const int acc = it % p.n_acc;                          // n_acc is even: an accumulator always belongs to the same set
            const __nv_bfloat16* rrow = (RES && row_ok) ? p.residual + (size_t)m * p.N : nullptr;
            // residual: the 32 columns of a TMEM read are fetched one read AHEAD (the first before the accumulator is even
            // complete), so the L2 round trip of these row-strided loads is not paid once per 32 columns
            uint4 rnx[4];
            auto load_res = [&](int jb_, int c32_) {
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const int col = jb_ * 64 + c32_ + h * 8, n = n_base + col;
                    rnx[h] = (rrow && col < p.n_pad && n + 8 <= p.N) ? __ldg((const uint4*)(rrow + n)) : make_uint4(0u, 0u, 0u, 0u);
                }
            };
            if (RES) load_res(split ? half : 0, 0);
            mbar_wait(tfull0 + 8 * acc, (uint32_t)(it / p.n_acc) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.n_pad);
            for (int jb = split ? half : 0; jb < ((p.debug & 16) ? 0 : nblk64); jb += split ? 2 : 1, blk_count++) {
                const uint32_t buf = my_staging + ((split || !p.epi_db) ? 0u : (blk_count & 1u) * STAGING_BLOCK_BYTES);
                // One staging buffer: its previous store must have been read before anyone writes -> wait + barrier here.
                // Two buffers (dbl): the issuer instead confirms, just BEFORE the barrier that ends a block, that the store issued
                // one block earlier has been read; every thread that writes buffer b in block k has passed the barrier of block
                // k - 1, which followed that confirmation for store k - 2 (the last user of b).  One barrier per block, not two.
                const bool dbl = !split && p.epi_db;
                if (!dbl) {
                    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    set_bar_sync(1 + grp);
                }
                const int cols_here = min(64, p.n_pad - jb * 64);
                for (int c32 = 0; c32 < cols_here; c32 += 32) {
                    uint32_t r[32];
                    tc_ld32(taddr + jb * 64 + c32, r);            // columns beyond n_pad read the other accumulator's TMEM: ignored below
                    uint4 rcur[4];
                    if (RES) {                                     // next read's residual columns (next 32 columns or next 64-column block)
#pragma unroll
                        for (int h = 0; h < 4; h++) rcur[h] = rnx[h];
                        if (c32 + 32 < cols_here) load_res(jb, c32 + 32);
                        else load_res(jb + (split ? 2 : 1), 0);
                    }
                    tc_ld_wait();
                    if (jb + (split ? 2 : 1) >= nblk64 && c32 + 32 >= cols_here) {     // last TMEM read of this group for this tile:
                        tc_fence_before();                                              // hand the accumulator back before the math
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
                    }
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const int col = jb * 64 + c32 + h * 8;     // column inside the tile
                        if (col < p.n_pad) {
                            const int n = n_base + col;
                            float v[8];
                            const float4 b0 = *(const float4*)(sbias + n), b1 = *(const float4*)(sbias + n + 4);
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                            if (ACT && !(p.debug & 2)) {
                                // swish on pairs with packed fp32 math: h = 0.5*(acc + b) as ONE fma (scaling by 0.5 is exact, so this
                                // rounds exactly like (acc + b) * 0.5), y = h + h * tanh(h): 2 FFMA2 + 2 MUFU per pair
#pragma unroll
                                for (int jj = 0; jj < 4; jj++) {
                                    uint64_t x2, hb2, half2, h2, t2, y2;
                                    asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "r"(r[h * 8 + 2 * jj]), "r"(r[h * 8 + 2 * jj + 1]));
                                    asm("mov.b64 %0, {%1, %2};" : "=l"(hb2) : "f"(0.5f * bb[2 * jj]), "f"(0.5f * bb[2 * jj + 1]));
                                    asm("mov.b64 %0, {%1, %2};" : "=l"(half2) : "f"(0.5f), "f"(0.5f));
                                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(h2) : "l"(x2), "l"(half2), "l"(hb2));
                                    float h0, h1, t0, t1;
                                    asm("mov.b64 {%0, %1}, %2;" : "=f"(h0), "=f"(h1) : "l"(h2));
                                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
                                    asm("mov.b64 %0, {%1, %2};" : "=l"(t2) : "f"(t0), "f"(t1));
                                    asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(y2) : "l"(h2), "l"(t2));
                                    asm("mov.b64 {%0, %1}, %2;" : "=f"(v[2 * jj]), "=f"(v[2 * jj + 1]) : "l"(y2));
                                }
                            } else {
#pragma unroll
                                for (int jj = 0; jj < 8; jj++) v[jj] = __uint_as_float(r[h * 8 + jj]) + bb[jj];
                            }
                            if (RES && rrow && n + 8 <= p.N) {
                                const __nv_bfloat162* rh = (const __nv_bfloat162*)&rcur[h];
#pragma unroll
                                for (int jj = 0; jj < 4; jj++) { float2 f = __bfloat1622float2(rh[jj]); v[2 * jj] += f.x; v[2 * jj + 1] += f.y; }
                            }
                            uint4 o;
                            __nv_bfloat162* oh = (__nv_bfloat162*)&o;
#pragma unroll
                            for (int jj = 0; jj < 4; jj++) oh[jj] = __floats2bfloat162_rn(v[2 * jj], v[2 * jj + 1]);
                            if (DENSE) { if (col < p.N) sts128(buf + (uint32_t)(row * (p.N * 2) + (col >> 3) * 16), o); }
                            else sts128(buf + (uint32_t)(row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4)), o);
                        }
                    }
                }
                fence_async_smem();
                if (dbl && issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                set_bar_sync(1 + grp);
                if (issuer) {
                    if (n_base + jb * 64 < p.N && !(p.debug & 1)) {
                        if (DENSE) {
                            // rows of a narrow C tile are 32-128 bytes: 128 separate row writes through the tensor path cost
                            // more than the tile's math; the tile is contiguous in C, so it goes out as ONE bulk copy
                            const int lim = p.a_mode == A_IMG ? p.hw : p.M;
                            const int rows = min(BLOCK_M, lim - m_blk * BLOCK_M);
                            const size_t g_row = (size_t)(p.a_mode == A_IMG ? img * p.hw : 0) + (size_t)m_blk * BLOCK_M;
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                         ::"l"(p.C + g_row * p.N), "r"(buf), "r"((uint32_t)(rows * p.N * 2)) : "memory");
                        } else if (p.a_mode == A_IMG) tma_store_3d(&map_c, n_base + jb * 64, m_blk * BLOCK_M, img, buf);
                        else tma_store_2d(&map_c, n_base + jb * 64, m_blk * BLOCK_M, buf);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if ((split && half >= nblk64) || (p.debug & 16)) {     // no column block for this group: still release the accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int get_encode(dfd_ctx* ctx) {
    if (g_encode) return DFD_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DFD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    DFD_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, DFD_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode = (PFN_encodeTiled)fn;
    return DFD_OK;
}

// 2D bf16 row-major [rows][cols] tensor, box [box_rows][64], 128-byte swizzle
static int make_map(dfd_ctx* ctx, CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return DFD_ERR_CUDA; }
    return DFD_OK;
}

// generic tiled map (rank <= 4; bf16 or fp32 elements), 32/64/128-byte swizzle, zero OOB fill -- used by mbconv_fused.cu for
// NHWC patches and by gemm_tf32x3.cu for the fp32 operands
int dfd_tmap_encode(dfd_ctx* ctx, CUtensorMap* m, int dtype_f32, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    int rc = get_encode(ctx);
    if (rc) return rc;
    cuuint64_t d[4], s[3];
    cuuint32_t b[4], e[4] = {1, 1, 1, 1};
    for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; i++) s[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = g_encode(m, dtype_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, (void*)base, d, s, b, e,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->err = "cuTensorMapEncodeTiled (rank " + std::to_string(rank) + ") failed (" + std::to_string((int)r) + ")"; return DFD_ERR_CUDA; }
    return DFD_OK;
}
int dfd_tmap_bf16(dfd_ctx* ctx, CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_bytes) {
    return dfd_tmap_encode(ctx, m, 0, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

// accumulator ring depth: as many n_pad-column accumulators as fit 512 TMEM columns (the last one rounded up to 32), even, <= 8
static int gemm_n_acc(int n_pad) {
    if (getenv("DFD_GEMM_ACC2")) return 2;
    int n = 8;
    while (n > 2 && (n - 1) * n_pad + ((n_pad + 31) & ~31) > 512) n -= 2;
    return n;
}
typedef void (*GemmKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmParams);
static GemmKernel gemm_kernel(const GemmParams& p) {
    const int sel = (p.residual ? 4 : 0) | (p.act ? 2 : 0) | (p.dense_c ? 1 : 0);
    switch (sel) {
        case 0: return k_gemm_tcgen05<false, false, false>;
        case 1: return k_gemm_tcgen05<false, false, true>;
        case 2: return k_gemm_tcgen05<false, true, false>;
        case 3: return k_gemm_tcgen05<false, true, true>;
        case 4: return k_gemm_tcgen05<true, false, false>;
        case 5: return k_gemm_tcgen05<true, false, true>;
        case 6: return k_gemm_tcgen05<true, true, false>;
        default: return k_gemm_tcgen05<true, true, true>;
    }
}
static int gemm_set_attrs(dfd_ctx* ctx) {
    GemmParams q;
    memset(&q, 0, sizeof q);
    for (int sel = 0; sel < 8; sel++) {
        q.residual = (sel & 4) ? (const __nv_bfloat16*)1 : nullptr; q.act = (sel & 2) ? 1 : 0; q.dense_c = sel & 1;
        int rc = dfd_func_smem(ctx, gemm_kernel(q), 210 * 1024);        // per context (= per device), see dfd_func_smem
        if (rc) return rc;
    }
    return DFD_OK;
}

static bool g_no_dense = getenv("DFD_NO_DENSE_C") != nullptr;     // A/B switch for the dense bulk-store epilogue
static bool g_enabled = true;
static int g_debug = 0;
bool dfd_gemm_bf16_enabled() { return g_enabled; }
void dfd_gemm_free(dfd_ctx*) {}

// a_mode A_TMA: A = [M][K] matrix.  A_SCALE: A = [M][K] matrix gated by se[m / hw][k].  A_STEM: A = NHWC input
// [M / 12544][224][224][3], K must be 32 (27 taps + zero pad), W = [32][32].
int dfd_gemm_bf16_ex(dfd_ctx* ctx, int a_mode, const __nv_bfloat16* A, const float* se, int hw, const __nv_bfloat16* W,
                     const float* bias, const __nv_bfloat16* residual, __nv_bfloat16* C, int M, int N, int K, int act,
                     cudaStream_t st) {
    int rc = get_encode(ctx);
    if (rc) return rc;
    DFD_REQUIRE(K % 8 == 0 && N % 8 == 0 && N <= MAX_BIAS, DFD_ERR_INVALID, "gemm: K and N must be multiples of 8, N <= 1280");
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.act = act; p.bias = bias; p.residual = residual;
    p.a_mode = a_mode; p.A = A; p.se = se; p.hw = hw > 0 ? hw : 1; p.debug = g_debug; p.tiles_per_img = 1;
    p.C = C; p.dense_c = (N <= 64 && !g_no_dense) ? 1 : 0;
    // N tiling: the smallest number of equal UMMA-N blocks (multiples of 16, <= 256) covering N
    // (with several N blocks the block width is a multiple of 64 so the 64-column TMA stores of one block
    // never touch its neighbour's columns)
    int nb = (N + 255) / 256;
    int n_pad = nb == 1 ? (N + 15) / 16 * 16 : ((N + nb - 1) / nb + 63) / 64 * 64;
    p.n_pad = n_pad; p.n_blocks = (N + n_pad - 1) / n_pad;
    p.n_acc = gemm_n_acc(n_pad);
    const int m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
    p.num_tiles = m_blocks * p.n_blocks;
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    int staging_bytes = 4 * STAGING_BLOCK_BYTES;             // two 16 KB buffers per epilogue set
    p.epi_db = 1;
    // W stays resident in shared memory when it fits next to >= 3 A stages: every tile then loads only A
    // (measured: 148 CTAs re-fetching the same few KB of W per tile serialise on one L2 slice, 1-2 us per tile)
    const int b_bytes = num_kb * n_pad * BLOCK_K * 2;
    p.b_resident = (p.n_blocks == 1 && b_bytes + staging_bytes + 3 * A_STAGE_BYTES <= 204 * 1024) ? 1 : 0;
    const int stage_bytes = A_STAGE_BYTES + (p.b_resident ? 0 : n_pad * BLOCK_K * 2);
    int stages = (204 * 1024 - staging_bytes - (p.b_resident ? b_bytes : 0)) / stage_bytes;
    if (stages > 8) stages = 8;
    if (!p.b_resident && stages > 2 * num_kb && stages > 4) stages = 2 * num_kb > 4 ? 2 * num_kb : 4;
    // The two staging groups take alternate k-blocks.  With an even stage count a stage is always served by the same
    // group, so every waiter of a stage's mbarriers observes every phase; with an odd count a group would skip every
    // other phase of a stage and a parity wait could alias with the phase two uses earlier (seen as a rare hang of
    // b7.project with 3 stages).
    if (a_mode != A_TMA) stages &= ~1;
    if (a_mode != A_TMA && stages < 6 && !getenv("DFD_GEMM_EPI_DB")) {
        // deep-K layers with a wide, non-resident W: the k-block pipeline, not the epilogue, is the limit; give the
        // 32 KB of the second store-staging buffers to pipeline stages instead
        const int sb2 = 2 * STAGING_BLOCK_BYTES;
        int st2 = (204 * 1024 - sb2 - (p.b_resident ? b_bytes : 0)) / stage_bytes;
        if (st2 > 8) st2 = 8;
        st2 &= ~1;
        if (st2 > stages) { stages = st2; staging_bytes = sb2; p.epi_db = 0; }
    }
    DFD_REQUIRE(stages >= 2, DFD_ERR_INVALID, "gemm: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + (p.b_resident ? b_bytes : 0) + staging_bytes + 1024;
    if ((rc = gemm_set_attrs(ctx))) return rc;
    CUtensorMap ma, mb, mc;
    if (a_mode != A_STEM) { if ((rc = make_map(ctx, &ma, A, (uint64_t)M, (uint64_t)K, BLOCK_M))) return rc; }
    else memset(&ma, 0, sizeof ma);
    if ((rc = make_map(ctx, &mb, W, (uint64_t)N, (uint64_t)K, (uint32_t)n_pad))) return rc;
    if ((rc = make_map(ctx, &mc, C, (uint64_t)M, (uint64_t)N, BLOCK_M))) return rc;
    int grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
    DFD_CUDA(dfd_launch(ctx->pdl, gemm_kernel(p), dim3(grid), dim3(GEMM_THREADS), smem, st, ma, mb, mc, p));
    DFD_LAUNCH_CHECK("k_gemm_tcgen05", st);
    return DFD_OK;
}

// Project 1x1 conv with per-image gated weights: C[img] = A[img] (hw x K) . Wg[img] (N x K)^T + bias (+ residual).
int dfd_gemm_bf16_img(dfd_ctx* ctx, const __nv_bfloat16* A, const __nv_bfloat16* Wg, const float* bias,
                      const __nv_bfloat16* residual, __nv_bfloat16* C, int n_img, int hw, int N, int K, int act, cudaStream_t st) {
    int rc = get_encode(ctx);
    if (rc) return rc;
    DFD_REQUIRE(K % 8 == 0 && N % 8 == 0 && N <= 256, DFD_ERR_INVALID, "gemm_img: K and N must be multiples of 8, N <= 256");
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.M = n_img * hw; p.N = N; p.K = K; p.act = act; p.bias = bias; p.residual = residual;
    p.a_mode = A_IMG; p.A = A; p.se = nullptr; p.hw = hw; p.debug = 0;
    p.C = C; p.dense_c = (N <= 64 && !g_no_dense) ? 1 : 0; p.epi_db = 1;
    p.n_pad = (N + 15) / 16 * 16; p.n_blocks = 1;
    p.n_acc = gemm_n_acc(p.n_pad);
    p.tiles_per_img = (hw + BLOCK_M - 1) / BLOCK_M;
    p.num_tiles = n_img * p.tiles_per_img;
    p.b_resident = 0;
    const int staging_bytes = 4 * STAGING_BLOCK_BYTES;
    const int stage_bytes = A_STAGE_BYTES + p.n_pad * BLOCK_K * 2;
    int stages = (204 * 1024 - staging_bytes) / stage_bytes;
    if (stages > 8) stages = 8;
    DFD_REQUIRE(stages >= 2, DFD_ERR_INVALID, "gemm_img: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + staging_bytes + 1024;
    if ((rc = gemm_set_attrs(ctx))) return rc;
    CUtensorMap ma, mb, mc;
    {
        const uint64_t d[3] = {(uint64_t)K, (uint64_t)hw, (uint64_t)n_img}, s[2] = {(uint64_t)K * 2, (uint64_t)hw * K * 2};
        const uint32_t b[3] = {BLOCK_K, BLOCK_M, 1};
        if ((rc = dfd_tmap_bf16(ctx, &ma, A, 3, d, s, b, 128))) return rc;
    }
    {
        const uint64_t d[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)n_img}, s[2] = {(uint64_t)K * 2, (uint64_t)N * K * 2};
        const uint32_t b[3] = {BLOCK_K, (uint32_t)p.n_pad, 1};
        if ((rc = dfd_tmap_bf16(ctx, &mb, Wg, 3, d, s, b, 128))) return rc;
    }
    {
        const uint64_t d[3] = {(uint64_t)N, (uint64_t)hw, (uint64_t)n_img}, s[2] = {(uint64_t)N * 2, (uint64_t)hw * N * 2};
        const uint32_t b[3] = {BLOCK_K, BLOCK_M, 1};
        if ((rc = dfd_tmap_bf16(ctx, &mc, C, 3, d, s, b, 128))) return rc;
    }
    int grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
    DFD_CUDA(dfd_launch(ctx->pdl, gemm_kernel(p), dim3(grid), dim3(GEMM_THREADS), smem, st, ma, mb, mc, p));
    DFD_LAUNCH_CHECK("k_gemm_tcgen05", st);
    return DFD_OK;
}

int dfd_gemm_bf16(dfd_ctx* ctx, const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias,
                  const __nv_bfloat16* residual, __nv_bfloat16* C, int M, int N, int K, int act, cudaStream_t st) {
    return dfd_gemm_bf16_ex(ctx, A_TMA, A, nullptr, 0, W, bias, residual, C, M, N, K, act, st);
}

// ---------------------------------------------------------------------------------------------
// self-test against a CUDA-core reference (tests/test_gpu_gemm.py through dfd_gemm_selftest)
__global__ void k_gemm_ref(const __nv_bfloat16* A, const float* se, int hw, const __nv_bfloat16* W, const float* bias,
                           const __nv_bfloat16* res, float* C, int M, int N, int K, int act) {
    int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (n >= N || m >= M) return;
    float acc = 0.f;
    for (int k = 0; k < K; k++) {
        float a = __bfloat162float(A[(size_t)m * K + k]);
        if (se) a = __bfloat162float(__float2bfloat16_rn(a * se[(size_t)(m / hw) * K + k]));
        acc = fmaf(a, __bfloat162float(W[(size_t)n * K + k]), acc);
    }
    acc += bias[n];
    if (act) acc = acc / (1.0f + expf(-acc));
    if (res) acc += __bfloat162float(res[(size_t)m * N + n]);
    C[(size_t)m * N + n] = acc;
}

__global__ void k_fill_bf16(__nv_bfloat16* x, size_t n, uint32_t seed, float scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    x[i] = __float2bfloat16_rn(((float)(h & 0xffff) / 32768.0f - 1.0f) * scale);
}

__global__ void k_fill_f32(float* x, size_t n, uint32_t seed, float lo, float hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = (uint32_t)i * 2246822519u ^ seed;
    h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13;
    x[i] = lo + (hi - lo) * (float)(h & 0xffff) / 65536.0f;
}

__global__ void k_maxerr(const __nv_bfloat16* c, const float* ref, size_t n, float* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float e = fabsf(__bfloat162float(c[i]) - ref[i]) / fmaxf(1.0f, fabsf(ref[i]));
    if (!(e == e)) e = 1e30f;
    atomicMax((int*)out, __float_as_int(e));        // non-negative floats order like ints
}

// with_residual bit 0: residual; bit 1: SE-gated A (A_SCALE) with hw = 49
extern "C" int dfd_gemm_selftest(dfd_ctx* ctx, int M, int N, int K, int act, int with_residual, double* max_err_host,
                                 void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    cudaStream_t st = (cudaStream_t)stream;
    const bool use_res = with_residual & 1, use_se = (with_residual & 2) != 0;
    const int hw = 49, n_img = (M + hw - 1) / hw;
    __nv_bfloat16 *A, *W, *R, *C;
    float *bias, *ref, *err, *se;
    DFD_CUDA(cudaMalloc(&A, (size_t)M * K * 2));
    DFD_CUDA(cudaMalloc(&W, (size_t)N * K * 2 + 4096));
    DFD_CUDA(cudaMalloc(&R, (size_t)M * N * 2));
    DFD_CUDA(cudaMalloc(&C, (size_t)M * N * 2));
    DFD_CUDA(cudaMalloc(&bias, (size_t)(N + 512) * 4));
    DFD_CUDA(cudaMalloc(&se, (size_t)n_img * K * 4));
    DFD_CUDA(cudaMalloc(&ref, (size_t)M * N * 4));
    DFD_CUDA(cudaMalloc(&err, 4));
    DFD_CUDA(cudaMemsetAsync(err, 0, 4, st));
    DFD_CUDA(cudaMemsetAsync(C, 0xff, (size_t)M * N * 2, st));
    DFD_CUDA(cudaMemsetAsync(bias, 0, (size_t)(N + 512) * 4, st));
    k_fill_bf16<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, st>>>(A, (size_t)M * K, 11u, 1.0f);
    k_fill_bf16<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, st>>>(W, (size_t)N * K, 22u, 0.25f);
    k_fill_bf16<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, st>>>(R, (size_t)M * N, 33u, 1.0f);
    k_fill_f32<<<(N + 255) / 256, 256, 0, st>>>(bias, (size_t)N, 44u, -0.5f, 0.5f);
    k_fill_f32<<<(unsigned)(((size_t)n_img * K + 255) / 256), 256, 0, st>>>(se, (size_t)n_img * K, 55u, 0.05f, 1.0f);
    int rc = dfd_gemm_bf16_ex(ctx, use_se ? A_SCALE : A_TMA, A, use_se ? se : nullptr, hw, W, bias, use_res ? R : nullptr, C, M,
                              N, K, act, st);
    if (rc == DFD_OK) {
        k_gemm_ref<<<dim3((N + 127) / 128, M), 128, 0, st>>>(A, use_se ? se : nullptr, hw, W, bias, use_res ? R : nullptr, ref, M, N, K, act);
        k_maxerr<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, st>>>(C, ref, (size_t)M * N, err);
        float e = 0.f;
        cudaError_t ce = cudaMemcpyAsync(&e, err, 4, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { ctx->err = std::string("gemm selftest: ") + cudaGetErrorString(ce); rc = DFD_ERR_CUDA; }
        if (max_err_host) *max_err_host = (double)e;
    }
    cudaFree(A); cudaFree(W); cudaFree(R); cudaFree(C); cudaFree(bias); cudaFree(ref); cudaFree(err); cudaFree(se);
    return rc;
}

// Times `iters` launches of one GEMM shape (diagnostics for kernel tuning; flags as GemmParams::debug, bit 2 = SE-gated A,
// bit 3 = residual).  Writes the mean milliseconds per launch to *ms_host.
extern "C" int dfd_gemm_bench(dfd_ctx* ctx, int M, int N, int K, int act, int flags, int iters, double* ms_host, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = 49, n_img = (M + hw - 1) / hw;
    __nv_bfloat16 *A, *W, *R, *C;
    float *bias, *se;
    DFD_CUDA(cudaMalloc(&A, (size_t)M * K * 2));
    DFD_CUDA(cudaMalloc(&W, (size_t)N * K * 2 + 4096));
    DFD_CUDA(cudaMalloc(&R, (size_t)M * N * 2));
    DFD_CUDA(cudaMalloc(&C, (size_t)M * N * 2));
    DFD_CUDA(cudaMalloc(&bias, (size_t)(N + 512) * 4));
    DFD_CUDA(cudaMalloc(&se, (size_t)n_img * K * 4));
    DFD_CUDA(cudaMemsetAsync(bias, 0, (size_t)(N + 512) * 4, st));
    k_fill_bf16<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, st>>>(A, (size_t)M * K, 11u, 1.0f);
    k_fill_bf16<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, st>>>(W, (size_t)N * K, 22u, 0.25f);
    k_fill_bf16<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, st>>>(R, (size_t)M * N, 33u, 1.0f);
    k_fill_f32<<<(unsigned)(((size_t)n_img * K + 255) / 256), 256, 0, st>>>(se, (size_t)n_img * K, 55u, 0.05f, 1.0f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    g_debug = flags & (3 | 16 | 32);
    int rc = DFD_OK;
    for (int i = 0; i < iters + 2 && rc == DFD_OK; i++) {
        if (i == 2) cudaEventRecord(e0, st);
        rc = dfd_gemm_bf16_ex(ctx, (flags & 4) ? A_SCALE : A_TMA, A, (flags & 4) ? se : nullptr, hw, W, bias, (flags & 8) ? R : nullptr, C,
                              M, N, K, act, st);
    }
    g_debug = 0;
    cudaEventRecord(e1, st);
    cudaError_t ce = cudaStreamSynchronize(st);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ce != cudaSuccess) { ctx->err = std::string("gemm bench: ") + cudaGetErrorString(ce); rc = DFD_ERR_CUDA; }
    if (ms_host) *ms_host = (double)ms / (iters > 0 ? iters : 1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(A); cudaFree(W); cudaFree(R); cudaFree(C); cudaFree(bias); cudaFree(se);
    return rc;
}
