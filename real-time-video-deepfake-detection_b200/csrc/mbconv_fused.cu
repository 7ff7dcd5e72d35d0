// Fused front half of an MBConv block for bf16 NHWC activations (reference model.py:63-72 ->
// lukemelas MBConvBlock: _expand_conv/_bn0/swish -> _depthwise_conv/_bn1/swish -> avg-pool of the SE squeeze):
//
//   x [m][hin][hin][cin]  --1x1 expand (tcgen05, fp32 accum in TMEM) + bias + swish-->  e (bf16, SHARED MEMORY ONLY)
//                         --depthwise kxk stride s (fp32 FMA) + bias + swish-->          out [m][hout][hout][cexp]
//                                                                                         + SE squeeze partials
//
// The expanded tensor is 6x the block input and, layer by layer, was written once and read once: 14.4 MB of the
// 27.4 MB/img layer-granular traffic.  Here it never leaves the SM.
//
// Persistent, warp-specialised CTAs (320 threads, 2 per SM).  Work item = (image, TH x TW output tile); inside an
// item the CC-wide channel chunks of the expanded tensor are produced and consumed one after the other:
//   warp 0      producer: one 4-D TMA box per 64 input channels fetches the item's input patch (zero fill outside the
//               image) into an A slot; per chunk a 2-D TMA box fetches the expand weights [CC][cin] and one bulk copy
//               the chunk's packed depthwise weights + both biases into a 2-deep B ring.
//   warp 1      MMA issuer: per chunk and per 128-pixel block of the patch, tcgen05.mma (M=128, N=CC, K=cin) into a
//               4-deep ring of TMEM accumulators; tcgen05.commit publishes accumulators and frees smem slots.
//   warps 2-9   consumers: tcgen05.ld (lane quarter = warp % 4, column half = warp group) -> + bias, swish, padding
//               pixels forced to zero (the reference pads the EXPANDED tensor) -> bf16 patch in shared memory; then the
//               depthwise stage of dwconv_bf16.cu on that patch (lane = channel pair, warp = output row, fp32x2 FMAs),
//               swish, bf16 stores, per-channel squeeze partials (fixed order: deterministic).
// So the TMA latency, the MMA and the TMEM drain of chunk c+1 hide behind the depthwise math of chunk c.
// The halo is recomputed per tile.  The depthwise output is bit-identical to the unfused expand GEMM + k_dw_tile pair.
//
// WHOLE = the tile is the whole output image (14x14 and 7x7 stages): only the hin x hin real pixels go through the
// MMA, the patch has no padding ring and the depthwise loops skip out-of-image taps instead.
#include "dfd_internal.cuh"
#include "effnet_plan.h"
#include "tc_ptx.cuh"

#define MF_THREADS 320
#define MF_CWARPS 8               // consumer warps
#define MF_NB 2                   // B ring depth (chunks)

int dfd_tmap_bf16(dfd_ctx* ctx, CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_bytes);

struct FrontParams {
    const float* aux;            // packed per chunk: Wd[K*K][CC], 0.5*be[CC], bd[CC]
    __nv_bfloat16* out;          // [m][hout][hout][C]
    float* pool;                 // [m][tiles][C] SE squeeze partials
    int C, cin, hin, hout, tiles_x, tiles, num_kb, n_chunks, n_items, na;
    int csplit, cpi;             // small batches: an (image, tile) is split into csplit items of cpi chunks each
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
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
__device__ __forceinline__ uint64_t mf_fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mf_add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float mf_tanh(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
}
// swish of two values with packed fp32 math: h = 0.5*x (+ hb = 0.5*bias), y = h + h*tanh(h).  Bit-identical to
// swish_fast(x + bias): scaling by 0.5 is exact, so fma(x, 0.5, 0.5*b) rounds once exactly like (x + b) * 0.5.
__device__ __forceinline__ uint64_t mf_swish2(uint64_t x, uint64_t half_bias) {
    const uint64_t h = mf_fma2(x, mf_pack2(0.5f, 0.5f), half_bias);
    float h0, h1;
    mf_unpack2(h, h0, h1);
    const uint64_t t = mf_pack2(mf_tanh(h0), mf_tanh(h1));
    return mf_fma2(h, t, h);
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int K, int S, int TW, int TH, int CC, int HIN, bool WHOLE, int KW, int NT>
struct MfGeom {
    static constexpr int HOUT = (HIN + S - 1) / S;
    static constexpr int PAD_TOTAL = (HOUT - 1) * S + K - HIN > 0 ? (HOUT - 1) * S + K - HIN : 0;
    static constexpr int PAD = PAD_TOTAL / 2;                    // TF-"SAME": the smaller half goes in front
    static constexpr int PH = WHOLE ? HIN : (TH - 1) * S + K, PW = WHOLE ? HIN : (TW - 1) * S + K;
    static constexpr int NPIX = PH * PW;                         // patch pixels = MMA rows that carry data
    static constexpr int N_MB = (NPIX + 127) / 128;
    static constexpr int ROWB = KW * 2;                          // bytes per operand row = swizzle width (32 / 64 / 128)
    static constexpr int A_KB_BYTES = (NPIX * ROWB + 1023) & ~1023;  // one KW-channel k-block of the A operand (tight)
    static constexpr int AUX_FLOATS = (K * K + 2) * CC;          // Wd chunk, be chunk, bd chunk
    static constexpr int AUX_BYTES = (AUX_FLOATS * 4 + 127) & ~127;
    static constexpr int PITCH = CC * 2 + 16;                    // patch bytes per pixel (+16: conflict-free 16-byte stores)
    static constexpr int PATCH_BYTES = (NPIX * PITCH + 127) & ~127;
    static constexpr int TM_COLS = NT * CC <= 128 ? 128 : 256;
    static int b_stage_bytes(int num_kb) { return (num_kb * CC * ROWB + AUX_BYTES + 1023) & ~1023; }
    // (the MMA of the last 128-row block reads N_MB*128 rows from each k-block base; the overrun past the tight
    //  A slots lands in the B ring / patch that follow them inside the same allocation: garbage rows, never used)
    static size_t smem_bytes(int num_kb, int na) {
        return 1024 + (size_t)na * num_kb * A_KB_BYTES + (size_t)MF_NB * b_stage_bytes(num_kb) + PATCH_BYTES + MF_CWARPS * CC * 4 + 256;
    }
};

template <int K, int S, int TW, int TH, int CC, int HIN, bool WHOLE, int KW, int NT, int MINB>
__global__ void __launch_bounds__(MF_THREADS, MINB)
k_mbconv_front(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const FrontParams p) {
    using G = MfGeom<K, S, TW, TH, CC, HIN, WHOLE, KW, NT>;
    constexpr int PW = G::PW, NPIX = G::NPIX, N_MB = G::N_MB, PITCH = G::PITCH, PITCH_W = G::PITCH / 4, PAD = G::PAD, ROWB = G::ROWB;
    constexpr int MF_NT = NT;
    extern __shared__ __align__(1024) uint8_t smem_mf[];
    __shared__ __align__(8) uint64_t bars[2 + 2 + MF_NB * 2 + MF_NT * 2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t s0 = smem_u32(smem_mf);
    const uint32_t a_base = (s0 + 1023u) & ~1023u;
    const int a_slot_bytes = p.num_kb * G::A_KB_BYTES;
    const int b_stage = (p.num_kb * CC * ROWB + G::AUX_BYTES + 1023) & ~1023;
    const uint32_t b_base = a_base + (uint32_t)(p.na * a_slot_bytes);
    const uint32_t patch_s = b_base + (uint32_t)(MF_NB * b_stage);
    uint8_t* gen = smem_mf + (a_base - s0);
    const uint8_t* b_gen = gen + p.na * a_slot_bytes;
    uint32_t* patch = (uint32_t*)(gen + p.na * a_slot_bytes + MF_NB * b_stage);
    float* spool = (float*)((uint8_t*)patch + G::PATCH_BYTES);           // [MF_CWARPS][CC]
    const int aux_off = p.num_kb * CC * ROWB;                             // aux block inside a B stage

    const uint32_t bar0 = smem_u32(&bars[0]);
    const uint32_t a_full = bar0, a_empty = bar0 + 16, b_full = bar0 + 32, b_empty = b_full + 8 * MF_NB;
    const uint32_t t_full = b_empty + 8 * MF_NB, t_empty = t_full + 8 * MF_NT;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
        for (int i = 0; i < MF_NB; i++) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 2); }
        for (int i = 0; i < MF_NT; i++) { mbar_init(t_full + 8 * i, 1); mbar_init(t_empty + 8 * i, MF_CWARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(G::TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_trigger();                                             // PDL: the prologue above overlapped the predecessor's tail
    pdl_wait();

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int it = item / p.csplit, cs = item - it * p.csplit;
                const int ch_lo = cs * p.cpi, ch_hi = min(p.n_chunks, ch_lo + p.cpi);
                const int b = it / p.tiles, tile = it - b * p.tiles;
                const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
                const int iy0 = WHOLE ? 0 : ty * TH * S - PAD, ix0 = WHOLE ? 0 : tx * TW * S - PAD;
                mbar_wait_backoff(a_empty + 8 * as, aph ^ 1, 256);
                mbar_expect_tx(a_full + 8 * as, (uint32_t)(p.num_kb * NPIX * ROWB));
                for (int kb = 0; kb < p.num_kb; kb++)
                    tma_load_4d(a_base + as * a_slot_bytes + kb * G::A_KB_BYTES, &map_x, kb * KW, ix0, iy0, b, a_full + 8 * as);
                if (++as == p.na) { as = 0; aph ^= 1; }
                for (int ch = ch_lo; ch < ch_hi; ch++) {
                    mbar_wait_backoff(b_empty + 8 * bs, bph ^ 1, 256);
                    const uint32_t dst = b_base + bs * b_stage;
                    mbar_expect_tx(b_full + 8 * bs, (uint32_t)(p.num_kb * CC * ROWB + G::AUX_FLOATS * 4));
                    for (int kb = 0; kb < p.num_kb; kb++) tma_load_2d(dst + kb * CC * ROWB, &map_w, kb * KW, ch * CC, b_full + 8 * bs);
                    bulk_load(dst + aux_off, p.aux + (size_t)ch * G::AUX_FLOATS, G::AUX_FLOATS * 4, b_full + 8 * bs);
                    if (++bs == MF_NB) { bs = 0; bph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; int ts = 0; uint32_t tph = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int cs = item % p.csplit;
                const int ch_lo = cs * p.cpi, ch_hi = min(p.n_chunks, ch_lo + p.cpi);
                mbar_wait_backoff(a_full + 8 * as, aph, 128);
                tc_fence_after();
                const uint32_t a_slot = a_base + as * a_slot_bytes;
                for (int ch = ch_lo; ch < ch_hi; ch++) {
                    mbar_wait_backoff(b_full + 8 * bs, bph, 128);
                    tc_fence_after();
                    const uint32_t bsm = b_base + bs * b_stage;
#pragma unroll 1
                    for (int mb = 0; mb < N_MB; mb++) {
                        mbar_wait_backoff(t_empty + 8 * ts, tph ^ 1, 96);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + (uint32_t)(ts * CC);
                        for (int kb = 0; kb < p.num_kb; kb++) {
                            const uint64_t adesc = make_smem_desc_rows<ROWB>(a_slot + kb * G::A_KB_BYTES + mb * 128 * ROWB);
                            const uint64_t bdesc = make_smem_desc_rows<ROWB>(bsm + kb * CC * ROWB);
                            const int krem = p.cin - kb * KW;
                            const int ksteps = krem >= KW ? KW / 16 : (krem + 15) / 16;
                            for (int k = 0; k < ksteps; k++)
                                tc_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                        }
                        tc_commit(t_full + 8 * ts);
                        if (++ts == MF_NT) { ts = 0; tph ^= 1; }
                    }
                    tc_commit(b_empty + 8 * bs);               // weights of this chunk consumed by the tensor core
                    if (++bs == MF_NB) { bs = 0; bph ^= 1; }
                }
                tc_commit(a_empty + 8 * as);                   // every MMA reading this patch has retired
                if (++as == p.na) { as = 0; aph ^= 1; }
            }
        }
    } else {
        // ===== consumers: TMEM -> patch, depthwise, stores =====
        const int cw = warp - 2;                               // 0..7
        const int ctid = tid - 64;                             // 0..255
        const int q = warp & 3, hsel = cw >> 2;                // TMEM lane quarter (hardware: warp % 4), column half
        constexpr int HC = CC / 2, NP = HC / 8;                // columns per warp, x8 pieces
        const int pl = lane < CC / 2 ? lane : CC / 2 - 1;
        int bs = 0; uint32_t bph = 0; int ts = 0; uint32_t tph = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int it = item / p.csplit, cs = item - it * p.csplit;
            const int ch_lo = cs * p.cpi, ch_hi = min(p.n_chunks, ch_lo + p.cpi);
            const int b = it / p.tiles, tile = it - b * p.tiles;
            const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
            const int oy0 = ty * TH, ox0 = tx * TW;
            const int iy0 = oy0 * S - PAD, ix0 = ox0 * S - PAD;
            for (int chn = ch_lo; chn < ch_hi; chn++) {
                const int c0 = chn * CC;
                mbar_wait(b_full + 8 * bs, bph);               // aux block (dw weights + biases) landed
                const float* aux = (const float*)(b_gen + bs * b_stage + aux_off);
                const float* sw = aux;                         // [K*K][CC]
                const float* sbe = aux + K * K * CC;           // [CC] 0.5 * expand bias
                const float* sbd = sbe + CC;                   // [CC]
                // ---- epilogue of the expand GEMM: accumulator blocks -> bf16 patch ----
#pragma unroll 1
                for (int mb = 0; mb < N_MB; mb++) {
                    mbar_wait(t_full + 8 * ts, tph);
                    tc_fence_after();
                    const int rowbase = mb * 128 + q * 32;
                    const bool active = rowbase < NPIX;        // warp-uniform
                    uint32_t v[NP][8];
                    if (active) {
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ts * CC + hsel * HC);
#pragma unroll
                        for (int i = 0; i < NP; i++) tc_ld8(taddr + i * 8, v[i]);
                        tc_ld_wait();
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_empty + 8 * ts);          // accumulator drained: hand it back before the math
                    if (++ts == MF_NT) { ts = 0; tph ^= 1; }
                    const int r = rowbase + lane;
                    if (active && r < NPIX) {
                        bool inimg = true;
                        if (!WHOLE) {
                            const int py = r / PW, px = r - py * PW;
                            const int iy = iy0 + py, ix = ix0 + px;
                            inimg = iy >= 0 && iy < HIN && ix >= 0 && ix < HIN;
                        }
                        const uint32_t dst = patch_s + (uint32_t)(r * PITCH + hsel * HC * 2);
#pragma unroll
                        for (int i = 0; i < NP; i++) {
                            uint32_t o[4];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const uint64_t hb = *(const uint64_t*)(sbe + hsel * HC + i * 8 + 2 * j);      // 0.5 * bias pair
                                const uint64_t y = mf_swish2(mf_pack2(__uint_as_float(v[i][2 * j]), __uint_as_float(v[i][2 * j + 1])), hb);
                                float a0, a1;
                                mf_unpack2(y, a0, a1);
                                o[j] = inimg ? mf_bf16x2(a0, a1) : 0u;
                            }
                            sts128(dst + i * 16, make_uint4(o[0], o[1], o[2], o[3]));
                        }
                    }
                }
                consumer_sync();
                // ---- depthwise kxk + bias + swish + squeeze partials (lane = channel pair, warp = output row) ----
                const int ch = c0 + 2 * pl;
                const bool ch_ok = lane < CC / 2 && ch < p.C;
                const uint64_t bias2 = *(const uint64_t*)(sbd + 2 * pl);
                uint64_t ps = 0ull;
                for (int r = cw; r < TH; r += MF_CWARPS) {
                    const int oy = oy0 + r;
                    if (oy >= p.hout) break;
                    uint64_t acc[TW];
#pragma unroll
                    for (int i = 0; i < TW; i++) acc[i] = bias2;
#pragma unroll
                    for (int ky = 0; ky < K; ky++) {
                        const int prow_i = WHOLE ? r * S + ky - PAD : r * S + ky;       // patch row
                        if (WHOLE && (prow_i < 0 || prow_i >= HIN)) continue;           // warp-uniform: padding row
                        uint64_t w[K];
#pragma unroll
                        for (int kx = 0; kx < K; kx++) w[kx] = *(const uint64_t*)(sw + (ky * K + kx) * CC + 2 * pl);
                        const uint32_t* prow = patch + (size_t)(prow_i * PW) * PITCH_W + pl;
#pragma unroll
                        for (int ix = 0; ix < PW; ix++) {
                            const uint32_t v = prow[ix * PITCH_W];
                            const uint64_t x = mf_pack2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
                            // patch column ix feeds output column ox through tap kx when ox*S + kx - OFF == ix
                            // (OFF = PAD when the patch is the bare image, 0 when the patch carries its own halo)
                            constexpr int OFF = WHOLE ? PAD : 0;
#pragma unroll
                            for (int kx = 0; kx < K; kx++) {
                                const int t = ix + OFF - kx;
                                if (t >= 0 && t % S == 0 && t / S < TW) mf_ffma2(acc[t / S], x, w[kx]);
                            }
                        }
                    }
                    __nv_bfloat16* orow = p.out + (((size_t)b * p.hout + oy) * p.hout + ox0) * p.C + ch;
#pragma unroll
                    for (int i = 0; i < TW; i++) {
                        if (ox0 + i < p.hout && ch_ok) {
                            const uint64_t y = mf_swish2(acc[i], 0ull);
                            ps = mf_add2(ps, y);
                            float y0, y1;
                            mf_unpack2(y, y0, y1);
                            *(__nv_bfloat162*)(orow + (size_t)i * p.C) = __floats2bfloat162_rn(y0, y1);
                        }
                    }
                }
                float ps0, ps1;
                mf_unpack2(ps, ps0, ps1);
                if (lane < CC / 2) { spool[cw * CC + 2 * lane] = ps0; spool[cw * CC + 2 * lane + 1] = ps1; }
                consumer_sync();                               // patch + aux reads done, squeeze partials complete
                if (ctid < CC && c0 + ctid < p.C) {            // deterministic: fixed-order sum, one partial per (image, tile, channel)
                    float sacc = 0.f;
#pragma unroll
                    for (int wv = 0; wv < MF_CWARPS; wv++) sacc += spool[wv * CC + ctid];
                    p.pool[((size_t)b * p.tiles + tile) * p.C + c0 + ctid] = sacc;
                }
                if (ctid == 0) mbar_arrive(b_empty + 8 * bs);  // (the tensor core's commit is the other arrival)
                if (++bs == MF_NB) { bs = 0; bph ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G::TM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// chunk width per block: 48 divides every expanded width (96 .. 672) exactly, but fills only 24 of the 32 lanes of the
// depthwise stage (lane = channel pair).  64-wide chunks where the padding of the last chunk costs less than that:
// 240 = 3.75 x 64, 480 = 7.5 x 64, 1152 = 18 x 64 (96, 144 and 672 stay at 48: 672 does not fit two CTAs per SM at 64).
static int front_cc(int blk) {
    const int c = EFF_BLOCKS[blk].cexp;
    return (c == 1152 || c == 480 || c == 240) ? 64 : 48;
}

// Packs, per block and per chunk, the depthwise weights and both biases into one contiguous bulk-copy source.
int dfd_front_pack(dfd_ctx* ctx, const float* blob) {
    const EffOffsets o = eff_offsets();
    size_t tot = 0;
    for (int i = 1; i < 16; i++) {
        const EffBlock& b = EFF_BLOCKS[i];
        const int cc = front_cc(i), nch = (b.cexp + cc - 1) / cc;
        ctx->front_aux_off[i] = tot;
        tot += (size_t)nch * (b.k * b.k + 2) * cc;
        tot = (tot + 31) & ~(size_t)31;               // 128-byte aligned blocks
    }
    std::vector<float> h(tot, 0.f);
    for (int i = 1; i < 16; i++) {
        const EffBlock& b = EFF_BLOCKS[i];
        const EffBlockOff& f = o.blk[i];
        const int cc = front_cc(i), nch = (b.cexp + cc - 1) / cc, kk = b.k * b.k;
        for (int ch = 0; ch < nch; ch++) {
            float* dst = h.data() + ctx->front_aux_off[i] + (size_t)ch * (kk + 2) * cc;
            for (int j = 0; j < cc; j++) {
                const int c = ch * cc + j;
                if (c >= b.cexp) continue;
                for (int t = 0; t < kk; t++) dst[t * cc + j] = blob[f.wd + (size_t)t * b.cexp + c];
                dst[kk * cc + j] = 0.5f * blob[f.be + c];          // the epilogue's swish works on h = 0.5 * (acc + bias)
                dst[(kk + 1) * cc + j] = blob[f.bd + c];
            }
        }
    }
    if (!ctx->d_front_aux) DFD_CUDA(cudaMalloc(&ctx->d_front_aux, tot * sizeof(float)));
    DFD_CUDA(cudaMemcpy(ctx->d_front_aux, h.data(), tot * sizeof(float), cudaMemcpyHostToDevice));
    return DFD_OK;
}

template <int K, int S, int TW, int TH, int CC, int HIN, bool WHOLE, int KW = 64, int NT = 4, int MINB = 2>
static int launch_front(dfd_ctx* ctx, int blk, const __nv_bfloat16* x, const __nv_bfloat16* We, __nv_bfloat16* out, int m,
                        int* n_parts, cudaStream_t st) {
    using G = MfGeom<K, S, TW, TH, CC, HIN, WHOLE, KW, NT>;
    const EffBlock& b = EFF_BLOCKS[blk];
    DFD_REQUIRE(G::PAD == b.pad && G::HOUT == b.hout && front_cc(blk) == CC, DFD_ERR_INVALID, "mbconv_front: geometry mismatch");
    const int num_kb = (b.cin + KW - 1) / KW;
    // two A slots when MINB CTAs still fit one SM, else one
    const size_t budget = (size_t)(227 * 1024) / MINB - 1536;
    const int na = G::smem_bytes(num_kb, 2) <= budget ? 2 : 1;
    const size_t smem = G::smem_bytes(num_kb, na);
    // (two CTAs per SM need the full shared-memory carve-out: the driver's default picks a smaller one)
    { int rc0 = dfd_func_smem(ctx, k_mbconv_front<K, S, TW, TH, CC, HIN, WHOLE, KW, NT, MINB>, smem, true); if (rc0) return rc0; }
    CUtensorMap mx, mw;
    int rc;
    {
        const uint64_t dims[4] = {(uint64_t)b.cin, (uint64_t)b.hin, (uint64_t)b.hin, (uint64_t)m};
        const uint64_t str[3] = {(uint64_t)b.cin * 2, (uint64_t)b.hin * b.cin * 2, (uint64_t)b.hin * b.hin * b.cin * 2};
        const uint32_t box[4] = {(uint32_t)KW, (uint32_t)G::PW, (uint32_t)G::PH, 1};
        if ((rc = dfd_tmap_bf16(ctx, &mx, x, 4, dims, str, box, G::ROWB))) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)b.cin, (uint64_t)b.cexp};
        const uint64_t str[1] = {(uint64_t)b.cin * 2};
        const uint32_t box[2] = {(uint32_t)KW, (uint32_t)CC};
        if ((rc = dfd_tmap_bf16(ctx, &mw, We, 2, dims, str, box, G::ROWB))) return rc;
    }
    FrontParams p;
    p.aux = ctx->d_front_aux + ctx->front_aux_off[blk];
    p.out = out; p.pool = ctx->d_pool;
    p.C = b.cexp; p.cin = b.cin; p.hin = b.hin; p.hout = b.hout; p.num_kb = num_kb; p.na = na;
    const int tiles_x = (b.hout + TW - 1) / TW, tiles_y = (b.hout + TH - 1) / TH;
    p.tiles_x = tiles_x; p.tiles = tiles_x * tiles_y;
    p.n_chunks = (b.cexp + CC - 1) / CC;
    // enough items to occupy every CTA slot at small batch: split the chunk loop of an (image, tile) across items
    p.csplit = 1;
    while (m * p.tiles * p.csplit < MINB * ctx->sm_count && p.csplit < p.n_chunks) p.csplit++;
    p.cpi = (p.n_chunks + p.csplit - 1) / p.csplit;
    p.csplit = (p.n_chunks + p.cpi - 1) / p.cpi;              // no empty items
    p.n_items = m * p.tiles * p.csplit;
    *n_parts = p.tiles;
    if ((size_t)p.tiles * b.cexp > DFD_POOL_FLOATS) { ctx->err = "internal: squeeze partial buffer too small"; return DFD_ERR_CAPACITY; }
    int grid = MINB * ctx->sm_count;
    if (grid > p.n_items) grid = p.n_items;
    DFD_CUDA(dfd_launch(ctx->pdl, k_mbconv_front<K, S, TW, TH, CC, HIN, WHOLE, KW, NT, MINB>, dim3(grid), dim3(MF_THREADS), smem, st, mx, mw, p));
    DFD_LAUNCH_CHECK("k_mbconv_front", st);
    return DFD_OK;
}

// Expand 1x1 + depthwise of block `blk` (cexp != cin) in one kernel.  x: block input, We: bf16 expand weights [cexp][cin].
int dfd_mbconv_front_bf16(dfd_ctx* ctx, int blk, const __nv_bfloat16* x, const __nv_bfloat16* We, __nv_bfloat16* out, int m,
                          int* n_parts, cudaStream_t st) {
    const EffBlock& b = EFF_BLOCKS[blk];
#define MF_ARGS ctx, blk, x, We, out, m, n_parts, st
    if (b.k == 3 && b.s == 2 && b.hin == 112) return launch_front<3, 2, 7, 8, 48, 112, false, 16, 2, 3>(MF_ARGS);      // 16 input channels: 32-byte operand rows
    if (b.k == 3 && b.s == 1 && b.hin == 56) return launch_front<3, 1, 14, 14, 48, 56, false, 32, 2, 3>(MF_ARGS);    // 24 input channels: 64-byte operand rows, 3 CTAs/SM (-4 %)
    if (b.k == 5 && b.s == 2 && b.hin == 56) return launch_front<5, 2, 7, 7, 48, 56, false, 32, 4, 2>(MF_ARGS);        // (3 CTAs/SM at 64 registers spills the 5x5 window: +17 %)
    // 64-wide chunks with 32-byte operand rows (40 / 80 input channels = 3 / 5 exact 16-channel k-blocks: no zero-filled
    // operand columns, which is what makes the wider patch fit two CTAs per SM)
    if (b.k == 5 && b.s == 1 && b.hin == 28) return launch_front<5, 1, 14, 14, 64, 28, false, 16>(MF_ARGS);
    if (b.k == 3 && b.s == 2 && b.hin == 28) return launch_front<3, 2, 7, 7, 64, 28, false, 16>(MF_ARGS);
    if (b.k == 3 && b.s == 1 && b.hin == 14) return launch_front<3, 1, 14, 14, 64, 14, true, 16>(MF_ARGS);
    if (b.k == 5 && b.s == 1 && b.hin == 14 && b.cexp == 480) return launch_front<5, 1, 14, 14, 64, 14, true, 16>(MF_ARGS);
    if (b.k == 5 && b.s == 1 && b.hin == 14) return launch_front<5, 1, 14, 14, 48, 14, true>(MF_ARGS);
    if (b.k == 5 && b.s == 2 && b.hin == 14) return launch_front<5, 2, 7, 7, 48, 14, true>(MF_ARGS);
    if (b.k == 5 && b.s == 1 && b.hin == 7) return launch_front<5, 1, 7, 7, 64, 7, true>(MF_ARGS);
    if (b.k == 3 && b.s == 1 && b.hin == 7) return launch_front<3, 1, 7, 7, 64, 7, true>(MF_ARGS);
#undef MF_ARGS
    ctx->err = "mbconv_front: no tile configuration for this layer";
    return DFD_ERR_INVALID;
}
