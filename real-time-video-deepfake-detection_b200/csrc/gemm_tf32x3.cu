// fp32-ACCURATE convolution-as-GEMM on the 5th-generation tensor cores (sm_100a): the classifier's accuracy mode
// (dtype = DFD_F32, north_star: probability within 1e-4 of the reference's fp32 forward, model.py:63-72).
//
//   C[M,N] = act( A'[M,K] . W[N,K]^T + bias[N] ) (+ residual[M,N])      A, C, residual fp32 in HBM; accumulate fp32 in TMEM
//
// The tensor core has no fp32 operand type; kind::tf32 reads 19 of the 32 bits.  Every operand is therefore split into
// two tf32 terms, x = hi + lo with hi = rna_tf32(x) and lo = x - hi (exact in fp32, |lo| <= 2^-11 |x|), and the product is
// accumulated as three MMAs into the SAME fp32 accumulator:
//
//       A.W  ~=  A_hi.W_hi + A_lo.W_hi + A_hi.W_lo            (the dropped A_lo.W_lo term is <= 2^-22 relative)
//
// which restores ~22 bits per product ("3xTF32").  That alone is NOT fp32 accuracy: the tensor core adds into its fp32
// accumulator with ROUND-TOWARD-ZERO, one truncation per MMA instruction (K = 8), so a single accumulator drifts
// systematically by ~1 ulp per instruction -- measured here: relative error 4e-8 x K, i.e. 4.5e-5 at K = 1152, and
// |dp| = 4.7e-4 on the whole network (the fp32 CUDA-core path: 9e-6).  Three measures:
//  (1) the small correction products (A_lo.W_hi + A_hi.W_lo) and the main product (A_hi.W_hi) go to TWO different TMEM
//      accumulators of the ring -- added to a large accumulator under round-toward-zero, a tiny addend of the opposite sign
//      costs a whole ulp -- and the epilogue adds the pair;
//  (2) layers with K > 192 accumulate in CHUNKS of six k-blocks (K = 192): every chunk starts a fresh accumulator pair, and
//      the epilogue warps -- idle during the k-loop anyway -- drain every accumulator as it completes and add it, with
//      round-to-nearest FADDs, into a running tile in shared memory (the store-staging buffer in its final layout).  The
//      truncation of a chunk is relative to the chunk's own small magnitude and follows the chunk's sign, so across chunks it
//      averages out instead of accumulating (Ootomo & Yokota's observation for Ampere mma.sync, restated for TMEM);
//  (3) what remains is, to 70 %, a pure SCALE factor: truncation toward zero shrinks a main-term accumulator by an expected
//      2.51e-8 * n^0.87 after n accumulations (tools/tf32_bias.py, profiles/tf32_bias_r02.txt: 8.4e-8 / 1.5e-7 / 2.8e-7 at
//      n = 4 / 8 / 16), and the epilogue multiplies every drained main accumulator by 1 + that expectation, which turns
//      round-toward-zero into an unbiased rounding.  The same law holds for the single pair of the shallow (K <= 192) layers.
//      (Chunk length and the plain / chunked boundary were re-swept at the end of round 2 -- 4 / 160, 6 / 192, 8 / 256, 12 / 384:
//      5.36 / 5.21 / 5.19 / 5.17 ms per 256 crops, largest per-layer error 1.7 / 2.4 / 3.4 / 4.4 e-6, whole network max |dp|
//      8.0 / 5.5 / 6.0 / 7.2 e-6: six k-blocks take most of the time; the largest per-layer error over every test shape is 4.1e-6.)
// Measured: rms relative error 0.6e-7 .. 1.9e-7 for every layer shape with a residual scale bias below 1.3e-8 (one accumulator
// at K = 1152: 4.5e-5); whole network max |dp| 4.7e-4 -> 8e-6 (fp32 CUDA-core path: 9e-6).
// The weights are split once on the host (W_hi / W_lo planes).  The activations are split ON THE FLY, tile by tile, in
// shared memory, so HBM holds plain fp32 tensors:
//
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads (fp32, SWIZZLE_128B: 32 floats = one 128-byte row) of the raw
//               A tile (128 x 32) and -- unless W is resident -- the W_hi / W_lo tiles (n_pad x 32 each) of the k-block.
//   warps 12-19 two groups of 4 staging warps on alternate k-blocks (a stage is served by whichever group its k-block falls to): wait for the raw tile, multiply by the squeeze-excite
//               gates (A_SCALE: project convs, the gated tensor never exists in HBM), split: hi written in place, lo into
//               the stage's second A buffer at the same swizzled offset (the pass is address-agnostic), fence.proxy.async.
//               A_STEM: gather the im2col row of the 3x3 stride-2 stem (27 taps + zero pad = 32 floats) instead.
//   warp 1      MMA issuer: per k-block 3 x 4 tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = n_pad, K = 8).
//   warp 2      TMEM allocator (ring of n_acc accumulators of n_pad fp32 columns).
//   warps 4-11  two epilogue groups in ping-pong: tcgen05.ld 32x32b.x32 -> + bias, swish_f32 (~3 ulp; NOT tanh.approx: that
//               approximation is 2^-11), + fp32 residual -> 128-byte-swizzled staging -> TMA store of each 32-column
//               block (or one bulk copy of the whole tile when N <= 32).
//               Shallow, wide layers (the first expand convs; epi_groups = 3): warps 16-19 are a THIRD epilogue group and warps
//               12-15 stage every k-block -- there the epilogue is the critical path and the second staging group only spins.
// Shared memory: as many pipeline stages as fit (2-8, odd counts included; W resident only when that costs no stage) -- every
// layer shape is bound by the latency of the TMA -> split -> MMA -> release chain, and ring depth is what hides it.
// Same skeleton (barrier protocol, accumulator ring, PDL) as the bf16 kernel in gemm_tcgen05.cu; byte geometry is
// identical (128-byte operand rows), only the element type, the MMA kind and the split pass differ.
#include "dfd_internal.cuh"
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "tc_ptx.cuh"

#define TBLOCK_M 128
#define TBLOCK_K 32                            // floats per k-block = one 128-byte swizzled row
#define TGEMM_THREADS 640
#define TA_BYTES (TBLOCK_M * 128)              // one A buffer (hi or lo) of a stage
#define TSTAGING_BLOCK_BYTES (TBLOCK_M * 128)  // 128 rows x 32 fp32 columns
#define TMAX_BIAS 1280

enum { TA_PLAIN = 0, TA_SCALE = 1, TA_STEM = 2 };

struct TGemmParams {
    int M, N, K;
    int n_pad, n_blocks, num_tiles, stages, act, a_mode, hw, b_resident, n_acc, epi_db, dense_c;
    uint32_t stg_stride;       // plain epilogue: bytes per store-staging block (128 rows x 128 B; dense N <= 32 tiles: 128 rows x N x 4 B -- the
                               //    difference buys the huge-M project layers another pipeline stage)
    int raw_hi;                // plain A (no gate, no im2col): the MMA reads the RAW fp32 tile as the hi operand (the tensor core uses the top 19 bits =
                               //    hi truncated; lo = x - trunc(x) stays exact), so two of the three MMA groups start when the TMA lands instead of
                               //    after the split pass, and the pass writes one plane instead of two
    int pf;                    // A tiles prefetched into L2 this many k-blocks ahead of their TMA load (0 = off)
    uint32_t bias_off;         // byte offset of the bias table from the aligned base of dynamic shared memory
    int epi_groups;            // 2: warps 4-11 drain accumulators, warps 12-19 stage A (two groups on alternate k-blocks).  3 (shallow plain layers with a
                               //    wide N: the epilogue is the critical path and one staging group keeps up): warps 16-19 are a THIRD epilogue group
    int ch;                    // k-blocks per accumulation chunk
    float beta_instr;          // expected relative shrink of a main-term accumulator after n MMA accumulations (round toward zero) = beta_instr * n^0.87, compensated in the epilogue
    float beta_plain;          // the same expectation for the one main-term accumulator of a plain (K <= 192) layer
    int chunked;               // every chunk uses TWO ring accumulators (correction terms, main term).  1: several chunks per tile, the
                               //    epilogue adds them into a running tile with round-to-nearest; 0 (K <= 192): one chunk per tile,
                               //    plain ping-pong epilogue (main * (1 + beta) + corrections)
    const float* bias;
    const float* residual;
    const float* A;            // A_STEM: NHWC input [B,224,224,3]
    float* C;
    const float* se;           // A_SCALE: [images][K] gates
};

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// x ~= hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi): |x - hi - lo| <= 2^-22 |x|
#ifndef DFD_TF32_CVT
// rna_tf32 with two full-rate integer instructions, (bits + 0x1000) & ~0x1fff: round to nearest, ties away from zero -- what
// cvt.rna.tf32.f32 computes for every finite value below the overflow threshold (activations), and what the host uses for the
// weight planes (tf32_rna_host).  cvt is a conversion-pipe instruction (16 / clk / SM): 64 of them per thread and k-block were
// the staging warps' longest dependency chain.
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    const float d = x - hi;                                  // exact
    lo = __uint_as_float((__float_as_uint(d) + 0x1000u) & 0xffffe000u);   // rounded, not left to the tensor core's truncation: one more bit
}
#else
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    const float d = x - hi;                                  // exact
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(d));      // rounded, not left to the tensor core's truncation: one more bit
    lo = __uint_as_float(h);
}
#endif
__device__ __forceinline__ void tf32_split4(const float4 v, uint4& hi, uint4& lo) {
    float h, l;
    tf32_split(v.x, h, l); hi.x = __float_as_uint(h); lo.x = __float_as_uint(l);
    tf32_split(v.y, h, l); hi.y = __float_as_uint(h); lo.y = __float_as_uint(l);
    tf32_split(v.z, h, l); hi.z = __float_as_uint(h); lo.z = __float_as_uint(l);
    tf32_split(v.w, h, l); hi.w = __float_as_uint(h); lo.w = __float_as_uint(l);
}
__device__ __forceinline__ float swish_exact(float x) { return swish_f32(x); }
__device__ __forceinline__ bool jb_none(int grp, int nblk32) { return grp >= nblk32; }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

// Pipeline stage and mbarrier parities of the CTA's j-th k-block: stage = j % S, ring parity (j / S) & 1.
// The "raw tile landed" barrier needs more care.  Two staging groups take alternate k-blocks and a group waits only for ITS
// k-blocks; with an odd S a stage alternates between the groups, so on a per-stage barrier a group would skip every other
// phase -- and a parity wait can only tell the current phase from the one before it (if TMA j + 1 lands before TMA j, the group
// would fall through its wait for j + 3 and split stale data; the bf16 kernel saw this as a rare hang with 3 stages).  With an
// odd S every stage therefore has TWO raw barriers, one per group (rsel = j & 1): barrier (stage, group) is used once every
// 2 S k-blocks, by that group alone and in consecutive phases.  (Even S, or one staging group: a stage belongs to one group.)
// Kept as running counters (the producer and the MMA issuer are single threads on the critical path: divisions by a run-time
// stage count there cost a quarter of a microsecond per k-block, 6 % of the layer table).
struct TStageCursor {
    int stage, c2, two_s, S, split;          // c2 = j % (2 S)
    uint32_t phase, rph;
    __device__ __forceinline__ void init(int n_sg, int S_) { S = S_; two_s = 2 * S_; split = (n_sg == 2 && (S_ & 1)) ? 1 : 0; stage = 0; c2 = 0; phase = 0; rph = 0; }
    __device__ __forceinline__ void advance() {
        if (++stage == S) { stage = 0; phase ^= 1u; }
        if (++c2 == two_s) { c2 = 0; rph ^= 1u; }
    }
    __device__ __forceinline__ int rsel() const { return split ? (c2 & 1) : 0; }          // (2 S is even: c2 & 1 == j & 1)
    __device__ __forceinline__ uint32_t rphase() const { return split ? rph : phase; }
};

template <bool RES, bool ACT, bool DENSE>
__global__ void __launch_bounds__(TGEMM_THREADS, 1)
k_gemm_tf32x3(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bh,
              const __grid_constant__ CUtensorMap map_bl, const __grid_constant__ CUtensorMap map_c, const TGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[6 * 8 + 1];     // full[8], empty[8], raw[8], tmem_full[8], tmem_empty[8], bfull, raw of group 1 [8]
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float sgate[8][4][TBLOCK_K];  // A_SCALE: per staging warp, the k-block's gates of the tile's <= 4 images

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b_plane_bytes = (uint32_t)p.n_pad * 128u;                 // W_hi or W_lo tile of one k-block
    const uint32_t stage_bytes = 2u * TA_BYTES + (p.b_resident ? 0u : 2u * b_plane_bytes);
    const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
    const int num_kb = (p.K + TBLOCK_K - 1) / TBLOCK_K;
    const uint32_t b_region = smem_base + (uint32_t)p.stages * stage_bytes;  // resident W: num_kb x (hi, lo)
    const uint32_t staging = b_region + (p.b_resident ? (uint32_t)num_kb * 2u * b_plane_bytes : 0u);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), raw0 = smem_u32(&bars[16]);
    const uint32_t tfull0 = smem_u32(&bars[24]), tempty0 = smem_u32(&bars[32]), bfull = smem_u32(&bars[40]);
    const uint32_t raw1 = smem_u32(&bars[41]);               // (TStageCursor: odd stage counts)
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(p.n_acc - 1) * (uint32_t)p.n_pad + (((uint32_t)p.n_pad + 31u) & ~31u)) tmem_cols <<= 1;

    // bias of every column of the layer (n_pad * n_blocks floats) behind the staging blocks, in dynamic shared memory: as a static
    // 5 KB array it cost the deep layers their third pipeline stage
    float* sbias = (float*)(smem + (smem_base - smem_u32(smem)) + p.bias_off);
    for (int i = threadIdx.x; i < p.n_pad * p.n_blocks; i += TGEMM_THREADS) sbias[i] = i < p.N ? p.bias[i] : 0.f;
    if (warp == 0 && lane == 0) {
        if (p.a_mode != TA_STEM) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(bfull, 1);
        for (int s = 0; s < p.stages; s++) {
            // full[s]: the 4 staging warps of the stage's group (+ the TMA thread's expect_tx for a non-resident stem W: never, W is 8 KB)
            mbar_init(full0 + 8 * s, 4); mbar_init(empty0 + 8 * s, 1); mbar_init(raw0 + 8 * s, 1); mbar_init(raw1 + 8 * s, 1);
        }
        // tempty[a]: one epilogue group drains an accumulator (plain), or both do (chunked accumulation)
        for (int a = 0; a < 8; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, p.chunked ? 8 : 4); }
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
    if (warp == 0 && lane == 0 && p.b_resident) {
        mbar_expect_tx(bfull, (uint32_t)num_kb * 2u * b_plane_bytes);
        for (int kb = 0; kb < num_kb; kb++) {
            tma_load_2d(b_region + (uint32_t)kb * 2u * b_plane_bytes, &map_bh, kb * TBLOCK_K, 0, bfull);
            tma_load_2d(b_region + (uint32_t)kb * 2u * b_plane_bytes + b_plane_bytes, &map_bl, kb * TBLOCK_K, 0, bfull);
        }
    }
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            TStageCursor cur;
            cur.init(p.epi_groups == 3 ? 1 : 2, p.stages);
            const bool load_a = p.a_mode != TA_STEM, load_b = !p.b_resident;
            const uint32_t tx = (load_a ? (uint32_t)TA_BYTES : 0u) + (load_b ? 2u * b_plane_bytes : 0u);
            if (tx != 0) {
                // L2 prefetch cursor, p.pf k-blocks ahead of the loads (cp.async.bulk.prefetch.tensor needs no shared memory: the tile
                // is pulled into L2 early and the real load pays L2 latency instead of DRAM latency).  MEASURED AND OFF (DFD_TF32_PF=n):
                // 4 / 8 / 16 / 32 k-blocks ahead change the layer table by +0.1 / +0.7 / +1.9 / +3.9 % -- DRAM latency is not what
                // the pipeline waits for; its own hand-over chain (TMA -> split -> MMA -> release, four barrier hops) is.
                int pf_tile = blockIdx.x, pf_kb = 0;
                auto pf_issue = [&]() {
                    if (!load_a || pf_tile >= p.num_tiles) return;
                    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                                 ::"l"(&map_a), "r"(pf_kb * TBLOCK_K), "r"((pf_tile / p.n_blocks) * TBLOCK_M) : "memory");
                    if (++pf_kb == num_kb) { pf_kb = 0; pf_tile += gridDim.x; }
                };
                for (int i = 0; i < p.pf; i++) pf_issue();
                for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                    const int m_blk = tile / p.n_blocks, n_blk = tile % p.n_blocks;
                    for (int kb = 0; kb < num_kb; kb++) {
                        if (p.pf > 0) pf_issue();
                        const int stage = cur.stage; const uint32_t phase = cur.phase;
                        const uint32_t rawb = (cur.rsel() ? raw1 : raw0) + 8 * stage;
                        cur.advance();
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + 2u * TA_BYTES;
                        mbar_expect_tx(rawb, tx);
                        if (load_a) tma_load_2d(sa, &map_a, kb * TBLOCK_K, m_blk * TBLOCK_M, rawb);
                        if (load_b) {
                            tma_load_2d(sb, &map_bh, kb * TBLOCK_K, n_blk * p.n_pad, rawb);
                            tma_load_2d(sb + b_plane_bytes, &map_bl, kb * TBLOCK_K, n_blk * p.n_pad, rawb);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // idesc: D = F32 (bit 4), A = B = TF32 (2 << 7, 2 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(TBLOCK_M >> 4) << 24);
            TStageCursor cur;
            cur.init(p.epi_groups == 3 ? 1 : 2, p.stages);
            int acc = 0; uint32_t acc_phase = 0;
            int tile_it = 0;
            if (p.b_resident) mbar_wait(bfull, 0);
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                uint32_t d_main = 0, d_corr = 0;
                for (int kb = 0; kb < num_kb; kb++) {
                    const int kc = kb % p.ch;                      // (plain layers: ch = num_kb, one chunk per tile)
                    if (kc == 0) {                                 // a new chunk: a fresh PAIR of accumulators of the ring
                        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                        mbar_wait(tempty0 + 8 * (acc + 1), acc_phase ^ 1);
                        tc_fence_after();
                        d_corr = tmem_base + (uint32_t)(acc * p.n_pad);
                        d_main = d_corr + (uint32_t)p.n_pad;
                    }
                    const int stage = cur.stage, rsel = cur.rsel(); const uint32_t phase = cur.phase, rphase = cur.rphase();
                    cur.advance();
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint32_t sb = p.b_resident ? b_region + (uint32_t)kb * 2u * b_plane_bytes : sa + 2u * TA_BYTES;
                    const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + TA_BYTES);
                    const uint64_t b_hi = make_smem_desc(sb), b_lo = make_smem_desc(sb + b_plane_bytes);
                    const int krem = p.K - kb * TBLOCK_K;
                    const int ksteps = krem >= TBLOCK_K ? TBLOCK_K / 8 : (krem + 7) / 8;
                    // correction terms into their own accumulator (added to a large accumulator, a tiny addend of the opposite sign
                    // costs a whole ulp under round-toward-zero), and the dominant hi.hi product into its own
                    if (p.raw_hi) {
                        // hi = the raw tile as the TMA wrote it: these two groups do not wait for the split pass
                        mbar_wait((rsel ? raw1 : raw0) + 8 * stage, rphase);
                        tc_fence_after();
                        for (int k = 0; k < ksteps; k++)
                            tc_mma_tf32(d_corr, a_hi + (uint64_t)(k * 2), b_lo + (uint64_t)(k * 2), idesc, (kc | k) != 0);
                        for (int k = 0; k < ksteps; k++)
                            tc_mma_tf32(d_main, a_hi + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), idesc, (uint32_t)((kc | k) != 0));
                        mbar_wait(full0 + 8 * stage, phase);       // lo plane written
                        tc_fence_after();
                        for (int k = 0; k < ksteps; k++)
                            tc_mma_tf32(d_corr, a_lo + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), idesc, 1u);
                    } else {
                        mbar_wait(full0 + 8 * stage, phase);
                        tc_fence_after();
                        for (int k = 0; k < ksteps; k++)
                            tc_mma_tf32(d_corr, a_lo + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), idesc, (kc | k) != 0);
                        for (int k = 0; k < ksteps; k++)
                            tc_mma_tf32(d_corr, a_hi + (uint64_t)(k * 2), b_lo + (uint64_t)(k * 2), idesc, 1u);
                        for (int k = 0; k < ksteps; k++)
                            tc_mma_tf32(d_main, a_hi + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), idesc, (uint32_t)((kc | k) != 0));
                    }
                    tc_commit(empty0 + 8 * stage);                 // frees the smem stage when the MMAs retire
                    const bool chunk_end = kb == num_kb - 1 || kc == p.ch - 1;
                    if (chunk_end) {                               // publish the pair, move on in the ring
                        if (p.epi_groups == 3) {                   // (plain layers: one chunk per tile) see the epilogue's wait
                            tc_commit(tfull0 + 8 * (tile_it % 6));
                            tile_it++;
                        } else {
                            tc_commit(tfull0 + 8 * acc);
                            tc_commit(tfull0 + 8 * (acc + 1));
                        }
                        acc += 2;
                        if (acc == p.n_acc) { acc = 0; acc_phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= 12 && !(p.epi_groups == 3 && warp >= 16)) {
        const int n_sg = p.epi_groups == 3 ? 1 : 2;            // staging groups
        // ===== staging warps: two groups of 128 threads on alternate k-blocks (one group when epi_groups == 3); stage / barrier assignment: TStageCursor =====
        const int g = (warp - 12) >> 2;
        const int t = threadIdx.x - (12 + 4 * g) * 32;         // 0..127
        TStageCursor sc;                                      // positioned on this group's first k-block (j = g)
        sc.init(n_sg, p.stages);
        if (g == 1) sc.advance();
        if (p.a_mode == TA_STEM) {
            // stem im2col (one k-block per tile): row = output pixel; 3 kernel rows x 9 contiguous floats (3 px x 3 ch) -> 27 taps
            // + 5 zeros.  The 15 loads of the group's NEXT tile are in flight while it waits for the smem slot of the current one.
            auto gather = [&](int tile, float (&a)[3][9]) {
                const int m = tile * TBLOCK_M + t;
#pragma unroll
                for (int ky = 0; ky < 3; ky++)
#pragma unroll
                    for (int i = 0; i < 9; i++) a[ky][i] = 0.f;
                if (m < p.M) {
                    const int ox = m % 112, oy = (m / 112) % 112, b = m / (112 * 112);
#pragma unroll
                    for (int ky = 0; ky < 3; ky++) {
                        const int iy = 2 * oy + ky;
                        if (iy < 224) {
                            // (iy*224 + 2*ox) * 3 floats = a multiple of 6 floats: 8-byte aligned
                            const float2* src = (const float2*)(p.A + (((size_t)b * 224 + iy) * 224 + 2 * ox) * 3);
                            const float2 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
                            a[ky][0] = v0.x; a[ky][1] = v0.y; a[ky][2] = v1.x; a[ky][3] = v1.y; a[ky][4] = v2.x; a[ky][5] = v2.y;
                            if (ox < 111) {                      // the third pixel is padding at the right edge
                                const float2 v3 = __ldg(src + 3);
                                a[ky][6] = v3.x; a[ky][7] = v3.y; a[ky][8] = __ldg((const float*)(src + 4));
                            }
                        }
                    }
                }
            };
            float cur[3][9], nxt[3][9];
            int it = g;
            int tile = blockIdx.x + it * (int)gridDim.x;
            if (tile < p.num_tiles) gather(tile, cur);
            for (; tile < p.num_tiles; it += 2) {
                const int ntile = blockIdx.x + (it + 2) * (int)gridDim.x;
                if (ntile < p.num_tiles) gather(ntile, nxt);
                const int stage = sc.stage; const uint32_t phase = sc.phase;       // (the stem has no TMA loads: no raw barrier)
                sc.advance(); sc.advance();
                const uint32_t sa = smem_base + stage * stage_bytes;
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const int row = t;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int e = c * 4 + j;                 // compile-time: element 0..31 of the im2col row
                        v[j] = e < 27 ? cur[e / 9][e % 9] : 0.f;
                    }
                    uint4 hi, lo;
                    tf32_split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
                    const uint32_t off = (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4));
                    sts128(sa + off, hi);
                    sts128(sa + TA_BYTES + off, lo);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full0 + 8 * stage);
#pragma unroll
                for (int ky = 0; ky < 3; ky++)
#pragma unroll
                    for (int i = 0; i < 9; i++) cur[ky][i] = nxt[ky][i];
                tile = ntile;
            }
        } else {
            // TA_PLAIN / TA_SCALE: [gate,] split, in shared memory.  The gates of a k-block (<= 4 images x 32 channels for one
            // 128-row tile, hw >= 49) are fetched one k-block AHEAD (lane = image x 4-channel chunk) and published to the warp
            // through a 512-byte table, so their L2 latency is not paid between "raw tile landed" and "tile ready".
            const bool gated = p.a_mode == TA_SCALE;
            float4* wg = (float4*)&sgate[warp - 12][0][0];         // [4 images][8 chunks] x float4
            const int c = t & 7;
            const int gl_img = lane >> 3, gl_c = lane & 7;
            const int n_img = gated ? (p.M + p.hw - 1) / p.hw : 1;
            int tile = blockIdx.x, kb = g;                         // this group's next k-block
            auto norm = [&]() { while (kb >= num_kb && tile < p.num_tiles) { kb -= num_kb; tile += gridDim.x; } };
            norm();
            float4 q = make_float4(1.f, 1.f, 1.f, 1.f);
            auto prefetch = [&]() {
                if (!gated) return;
                q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tile < p.num_tiles) {
                    int img = ((tile / p.n_blocks) * TBLOCK_M) / p.hw + gl_img;
                    if (img > n_img - 1) img = n_img - 1;
                    const int k = kb * TBLOCK_K + gl_c * 4;
                    if (k < p.K) q = __ldg((const float4*)(p.se + (size_t)img * p.K + k));
                }
            };
            prefetch();
            for (int j = g; tile < p.num_tiles; j += n_sg) {
                const int m0 = (tile / p.n_blocks) * TBLOCK_M;
                const int stage = sc.stage, rsel = sc.rsel(); const uint32_t rphase = sc.rphase();
                sc.advance(); if (n_sg == 2) sc.advance();
                const uint32_t sa = smem_base + stage * stage_bytes;
                if (gated) { wg[gl_img * 8 + gl_c] = q; __syncwarp(); }
                const int k = kb * TBLOCK_K + c * 4;
                kb += n_sg;
                norm();
                prefetch();                                        // next k-block's gates fly during the wait and the pass
                mbar_wait((rsel ? raw1 : raw0) + 8 * stage, rphase);       // raw A (and W) tile landed
                const int img0 = gated ? m0 / p.hw : 0, rem0 = gated ? m0 - img0 * p.hw : 0;
                const int last = p.M - 1 - m0;                     // rows beyond M are clamped to the last row (their data is zero fill)
                uint4 v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int row = i * 16 + (t >> 3);
                    v[i] = lds128(sa + (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4)));
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int row = i * 16 + (t >> 3);
                    float4 f = make_float4(__uint_as_float(v[i].x), __uint_as_float(v[i].y), __uint_as_float(v[i].z), __uint_as_float(v[i].w));
                    if (gated && k < p.K) {
                        const int rr = rem0 + (row < last ? row : last);
                        const int rel = (rr >= p.hw) + (rr >= 2 * p.hw) + (rr >= 3 * p.hw);
                        const float4 gq = wg[rel * 8 + c];
                        f.x *= gq.x; f.y *= gq.y; f.z *= gq.z; f.w *= gq.w;     // fp32 product rounded once, like torch's x * gate
                    }
                    const uint32_t off = (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4));
                    if (p.raw_hi) {
                        // the raw tile stays as it is (the MMA reads its top 19 bits = hi truncated); lo = rna_tf32(x - trunc(x))
                        uint4 lo;
                        lo.x = (__float_as_uint(f.x - __uint_as_float(__float_as_uint(f.x) & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
                        lo.y = (__float_as_uint(f.y - __uint_as_float(__float_as_uint(f.y) & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
                        lo.z = (__float_as_uint(f.z - __uint_as_float(__float_as_uint(f.z) & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
                        lo.w = (__float_as_uint(f.w - __uint_as_float(__float_as_uint(f.w) & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
                        sts128(sa + TA_BYTES + off, lo);
                    } else {
                        uint4 hi, lo;
                        tf32_split4(f, hi, lo);
                        sts128(sa + off, hi);
                        sts128(sa + TA_BYTES + off, lo);
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full0 + 8 * stage);
            }
        }
    } else if (warp >= 4 && warp < 12 && p.chunked) {
        // ===== chunked accumulation: both epilogue groups serve the SAME tile (group = parity of the 32-column block) =====
        // Running tile in shared memory: nblk32 blocks of 128 rows x 128 bytes in the 128-byte-swizzled layout the TMA store
        // reads (dense layers: row-major rows of N floats), so the last chunk finishes the tile in place.
        const int n_chunks = 2 * ((num_kb + p.ch - 1) / p.ch);     // ring entries per tile: (corrections, main) per chunk
        const int grp = (warp - 4) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool issuer = q == 0 && lane == 0;
        const int nblk32 = (p.n_pad + 31) >> 5;
        uint32_t cc = 0;                                           // chunk counter of this CTA: accumulator = cc % n_acc
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int m_blk = tile / p.n_blocks, n_blk = tile % p.n_blocks;
            const int m = m_blk * TBLOCK_M + row;
            const bool row_ok = m < p.M;
            const int n_base = n_blk * p.n_pad;
            const float* rrow = (RES && row_ok) ? p.residual + (size_t)m * p.N : nullptr;
            // the previous tile's stores must have READ the running tile before its first chunk overwrites it
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            set_bar_sync(1 + grp);
            for (int ch = 0; ch < n_chunks; ch++, cc++) {
                const int acc = (int)(cc % (uint32_t)p.n_acc);
                const bool last = ch == n_chunks - 1;
                // main-term entries (odd): undo the EXPECTED truncation of their n MMA accumulations (header comment)
                float comp = 0.f;
                if (ch & 1) {
                    const int k0 = (ch >> 1) * p.ch * TBLOCK_K;
                    const int k1 = min(p.K, k0 + p.ch * TBLOCK_K);
                    comp = p.beta_instr * __powf((float)((k1 - k0 + 7) / 8), 0.87f);
                }
                mbar_wait(tfull0 + 8 * acc, (cc / (uint32_t)p.n_acc) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.n_pad);
                for (int jb = grp; jb < nblk32; jb += 2) {
                    const uint32_t buf = staging + (uint32_t)jb * TSTAGING_BLOCK_BYTES;
                    uint32_t r[32];
                    tc_ld32(taddr + jb * 32, r);
                    float4 rcur[8];
                    if (RES && last) {
#pragma unroll
                        for (int h = 0; h < 8; h++) {
                            const int col = jb * 32 + h * 4, n = n_base + col;
                            rcur[h] = (rrow && col < p.n_pad && n + 4 <= p.N) ? __ldg((const float4*)(rrow + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    tc_ld_wait();
                    if (jb + 2 >= nblk32) {                        // this group's last TMEM read of the chunk: hand the accumulator back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
                    }
#pragma unroll
                    for (int h = 0; h < 8; h++) {
                        const int col = jb * 32 + h * 4;
                        if (col < p.n_pad && (!DENSE || col < p.N)) {
                            const uint32_t addr = DENSE ? buf + (uint32_t)(row * (p.N * 4) + h * 16)
                                                        : buf + (uint32_t)(row * 128 + ((h ^ (row & 7)) << 4));
                            float v0 = __uint_as_float(r[h * 4]), v1 = __uint_as_float(r[h * 4 + 1]);
                            float v2 = __uint_as_float(r[h * 4 + 2]), v3 = __uint_as_float(r[h * 4 + 3]);
                            v0 = fmaf(v0, comp, v0); v1 = fmaf(v1, comp, v1); v2 = fmaf(v2, comp, v2); v3 = fmaf(v3, comp, v3);
                            if (ch != 0) {                         // running sum, round-to-nearest
                                const uint4 o = lds128(addr);
                                v0 += __uint_as_float(o.x); v1 += __uint_as_float(o.y); v2 += __uint_as_float(o.z); v3 += __uint_as_float(o.w);
                            }
                            if (last) {
                                const float4 bq = *(const float4*)(sbias + n_base + col);
                                v0 += bq.x; v1 += bq.y; v2 += bq.z; v3 += bq.w;
                                if (ACT) {
                                    const uint64_t s01 = swish_f32x2(f2_pack(v0, v1)), s23 = swish_f32x2(f2_pack(v2, v3));
                                    f2_unpack(s01, v0, v1); f2_unpack(s23, v2, v3);
                                }
                                if (RES && rrow && n_base + col + 4 <= p.N) { v0 += rcur[h].x; v1 += rcur[h].y; v2 += rcur[h].z; v3 += rcur[h].w; }
                            }
                            sts128(addr, make_uint4(__float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3)));
                        }
                    }
                }
                if (jb_none(grp, nblk32)) {                        // no column block for this group (n_pad <= 32): still release the accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
                }
                if (last) {
                    fence_async_smem();
                    set_bar_sync(1 + grp);
                    if (issuer) {
                        if (DENSE) {
                            if (grp == 0) {
                                const int rows = min(TBLOCK_M, p.M - m_blk * TBLOCK_M);
                                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                             ::"l"(p.C + (size_t)m_blk * TBLOCK_M * p.N), "r"(staging), "r"((uint32_t)(rows * p.N * 4)) : "memory");
                            }
                        } else {
                            for (int jb = grp; jb < nblk32; jb += 2)
                                if (n_base + jb * 32 < p.N) tma_store_2d(&map_c, n_base + jb * 32, m_blk * TBLOCK_M, staging + (uint32_t)jb * TSTAGING_BLOCK_BYTES);
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (warp >= 4) {
        // ===== plain epilogue (one chunk per tile): two groups of 4 warps in ping-pong over the accumulator ring =====
        const int set = warp >= 16 ? 2 : (warp - 4) >> 2;          // (warps 16-19: the third group of epi_groups == 3)
        const int q = warp & 3;                                    // TMEM lane quarter (= warp % 4)
        const int row = q * 32 + lane;
        const bool issuer = q == 0 && lane == 0;
        const uint32_t my_staging = staging + (uint32_t)(p.epi_db ? 2 * set : set) * p.stg_stride;
        const int nblk32 = (p.n_pad + 31) >> 5;
        uint32_t blk_count = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++) {
            if ((it % p.epi_groups) != set) continue;
            const int m_blk = tile / p.n_blocks, n_blk = tile % p.n_blocks;
            const int m = m_blk * TBLOCK_M + row;
            const bool row_ok = m < p.M;
            const int n_base = n_blk * p.n_pad;
            const int npairs = p.n_acc >> 1;
            const int acc = 2 * (it % npairs);                     // corrections in accumulator acc, main term in acc + 1
            const float* rrow = (RES && row_ok) ? p.residual + (size_t)m * p.N : nullptr;
            // Three groups on a ring of two accumulator pairs: "accumulator full" cannot be signalled on a barrier per pair -- a
            // group would visit a pair whose previous phase belongs to another group, and a parity wait can only tell the current
            // phase from the one before it (a waiter two phases ahead falls straight through, one that waits for an old phase
            // blocks for ever).  The MMA warp signals on barrier (tile % 6) instead: each of the six is used by ONE group, in order.
            if (p.epi_groups == 3) {
                mbar_wait(tfull0 + 8 * (it % 6), (uint32_t)(it / 6) & 1u);
            } else {
                mbar_wait(tfull0 + 8 * acc, (uint32_t)(it / npairs) & 1u);
                mbar_wait(tfull0 + 8 * (acc + 1), (uint32_t)(it / npairs) & 1u);
            }
            tc_fence_after();
            const uint32_t taddr_c = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.n_pad);
            const uint32_t taddr = taddr_c + (uint32_t)p.n_pad;
            for (int jb = 0; jb < nblk32; jb++, blk_count++) {
                const uint32_t buf = my_staging + (p.epi_db ? (blk_count & 1u) * p.stg_stride : 0u);
                // one buffer: its previous store must have been read before anyone writes -> wait + barrier here; two buffers:
                // the issuer confirms before the barrier that ENDS a block that the store issued one block earlier has been read
                if (!p.epi_db) {
                    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    set_bar_sync(1 + set);
                }
                uint32_t r[32];
                tc_ld32(taddr + jb * 32, r);                       // columns beyond n_pad read the next accumulator's TMEM: ignored below
                // residual: this block's 32 columns are fetched while the TMEM read is in flight (a one-block-ahead prefetch, as in
                // the bf16 kernel, does not fit the 102-register budget of 640 threads next to 32 fp32 accumulator values)
                float4 rcur[8];
                if (RES) {
#pragma unroll
                    for (int h = 0; h < 8; h++) {
                        const int col = jb * 32 + h * 4, n = n_base + col;
                        rcur[h] = (rrow && col < p.n_pad && n + 4 <= p.N) ? __ldg((const float4*)(rrow + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                tc_ld_wait();
                {   // main * (1 + beta) + corrections, 16 columns of the correction accumulator at a time (register budget)
                    const float cp = p.beta_plain;
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        uint32_t qc[16];
                        tc_ld16(taddr_c + jb * 32 + hh * 16, qc);
                        tc_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const float mv = __uint_as_float(r[hh * 16 + j]);
                            r[hh * 16 + j] = __float_as_uint(fmaf(mv, cp, mv) + __uint_as_float(qc[j]));
                        }
                    }
                }
                if (jb + 1 >= nblk32) {                            // last TMEM read of this tile: hand the accumulators back before the math
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(tempty0 + 8 * acc); mbar_arrive(tempty0 + 8 * (acc + 1)); }
                }
#pragma unroll
                for (int h = 0; h < 8; h++) {
                    const int col = jb * 32 + h * 4;               // column inside the tile
                    if (col < p.n_pad) {
                        const int n = n_base + col;
                        const float4 bq = *(const float4*)(sbias + n);
                        // packed fp32 pairs: + bias, swish_f32x2, + residual
                        const uint64_t r01 = f2_pack(__uint_as_float(r[h * 4]), __uint_as_float(r[h * 4 + 1]));
                        const uint64_t r23 = f2_pack(__uint_as_float(r[h * 4 + 2]), __uint_as_float(r[h * 4 + 3]));
                        uint64_t a01 = f2_add(r01, f2_pack(bq.x, bq.y));
                        uint64_t a23 = f2_add(r23, f2_pack(bq.z, bq.w));
                        if (ACT) { a01 = swish_f32x2(a01); a23 = swish_f32x2(a23); }
                        if (RES && rrow && n + 4 <= p.N) { a01 = f2_add(a01, f2_pack(rcur[h].x, rcur[h].y)); a23 = f2_add(a23, f2_pack(rcur[h].z, rcur[h].w)); }
                        float v0, v1, v2, v3;
                        f2_unpack(a01, v0, v1); f2_unpack(a23, v2, v3);
                        const uint4 o = make_uint4(__float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
                        if (DENSE) { if (col < p.N) sts128(buf + (uint32_t)(row * (p.N * 4) + h * 16), o); }
                        else sts128(buf + (uint32_t)(row * 128 + ((h ^ (row & 7)) << 4)), o);
                    }
                }
                fence_async_smem();
                if (p.epi_db && issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                set_bar_sync(1 + set);
                if (issuer) {
                    if (n_base + jb * 32 < p.N) {
                        if (DENSE) {
                            // N <= 32: the 128 x N tile is one contiguous block of C -> ONE bulk copy instead of 128 row writes
                            const int rows = min(TBLOCK_M, p.M - m_blk * TBLOCK_M);
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                         ::"l"(p.C + (size_t)m_blk * TBLOCK_M * p.N), "r"(buf), "r"((uint32_t)(rows * p.N * 4)) : "memory");
                        } else tma_store_2d(&map_c, n_base + jb * 32, m_blk * TBLOCK_M, buf);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
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
int dfd_tmap_encode(dfd_ctx* ctx, CUtensorMap* m, int dtype_f32, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// 2D fp32 row-major [rows][cols] tensor, box [box_rows][32 floats], 128-byte swizzle, zero fill outside
static int make_map_f32(dfd_ctx* ctx, CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    const uint64_t dims[2] = {cols, rows}, str[1] = {cols * 4};
    const uint32_t box[2] = {TBLOCK_K, box_rows};
    return dfd_tmap_encode(ctx, m, 1, base, 2, dims, str, box, 128);
}

static int tgemm_n_acc(int n_pad) {
    int n = 8;
    while (n > 2 && (n - 1) * n_pad + ((n_pad + 31) & ~31) > 512) n -= 2;
    return n;
}
typedef void (*TGemmKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TGemmParams);
static TGemmKernel tgemm_kernel(bool res, bool act, bool dense) {
    const int sel = (res ? 4 : 0) | (act ? 2 : 0) | (dense ? 1 : 0);
    switch (sel) {
        case 0: return k_gemm_tf32x3<false, false, false>;
        case 1: return k_gemm_tf32x3<false, false, true>;
        case 2: return k_gemm_tf32x3<false, true, false>;
        case 3: return k_gemm_tf32x3<false, true, true>;
        case 4: return k_gemm_tf32x3<true, false, false>;
        case 5: return k_gemm_tf32x3<true, false, true>;
        case 6: return k_gemm_tf32x3<true, true, false>;
        default: return k_gemm_tf32x3<true, true, true>;
    }
}

// a_mode 0: A = [M][K].  1: A = [M][K] gated by se[m / hw][k].  2: A = NHWC input [M / 12544][224][224][3], K = 32 (27 taps
// + zero pad), W = [32][32].  W_hi / W_lo: [N][K] fp32 planes, W_hi = rna_tf32(W), W_lo = rna_tf32(W - W_hi).
int dfd_gemm_tf32x3(dfd_ctx* ctx, int a_mode, const float* A, const float* se, int hw, const float* W_hi, const float* W_lo,
                    const float* bias, const float* residual, float* C, int M, int N, int K, int act, cudaStream_t st) {
    DFD_REQUIRE(K % 4 == 0 && N % 8 == 0 && N <= TMAX_BIAS, DFD_ERR_INVALID, "gemm_tf32x3: K % 4, N % 8, N <= 1280");
    DFD_REQUIRE(a_mode != TA_SCALE || hw >= 49, DFD_ERR_INVALID, "gemm_tf32x3: gated tiles span at most 4 images (hw >= 49)");
    TGemmParams p;
    memset(&p, 0, sizeof p);
    p.M = M; p.N = N; p.K = K; p.act = act; p.bias = bias; p.residual = residual;
    p.a_mode = a_mode; p.A = A; p.se = se; p.hw = hw > 0 ? hw : 1; p.C = C;
    p.dense_c = N <= 32 ? 1 : 0;
    // N tiling: equal UMMA-N blocks (multiples of 16, <= 128 so that a stage with both W planes stays <= 64 KB); with several
    // blocks the width is a multiple of 32 so that the 32-column TMA stores of one block never touch its neighbour's columns
    static const int n_max = getenv("DFD_TF32_NMAX") ? atoi(getenv("DFD_TF32_NMAX")) : 128;
    const int nb = (N + n_max - 1) / n_max;
    const int n_pad = nb == 1 ? (N + 15) / 16 * 16 : ((N + nb - 1) / nb + 31) / 32 * 32;
    p.n_pad = n_pad; p.n_blocks = (N + n_pad - 1) / n_pad;
    p.n_acc = tgemm_n_acc(n_pad);
    const int m_blocks = (M + TBLOCK_M - 1) / TBLOCK_M;
    p.num_tiles = m_blocks * p.n_blocks;
    const int num_kb = (K + TBLOCK_K - 1) / TBLOCK_K;
    // accumulation chunks of six k-blocks (see the header: the tensor core's accumulator truncates); K <= 192 is one chunk and
    // keeps the plain ping-pong epilogue
    static const int force_ch = getenv("DFD_TF32_CHUNK") ? atoi(getenv("DFD_TF32_CHUNK")) : 0;
    p.ch = 6;                                                // K = 192 per chunk: 24 main-term accumulations per accumulator
    if (force_ch > 0) p.ch = force_ch;
    static const int plain_kb = getenv("DFD_TF32_PLAIN_KB") ? atoi(getenv("DFD_TF32_PLAIN_KB")) : 6;
    if (num_kb <= plain_kb) p.ch = num_kb;                   // plain layers: one chunk (= one accumulator pair) per tile
    // K <= 192 (six k-blocks): one accumulator pair per tile and the ping-pong epilogue -- the shallow layers are huge-M and
    // epilogue-bound (b2.project: 196 us plain, 310 us chunked), their truncation bias is small and compensated as a whole
    const bool chunked = num_kb > plain_kb;
    p.chunked = chunked ? 1 : 0;
    // plain layers: the two epilogue groups take alternate tiles and an accumulator pair must always be drained by the same group
    // (the parity-wait notes at TStageCursor apply to the "accumulator full" barriers too), so their ring holds an EVEN number of
    // pairs.  No EfficientNet-B0 shape is affected (n_pad = 80 layers are chunked: both groups drain every chunk there).
    if (!chunked && p.n_acc == 6) p.n_acc = 4;
    // Expected-value compensation of the tensor core's round-toward-zero accumulation (measured with tools/tf32_bias.py: the
    // GEMM's error is, to 70 %, a pure scale factor 1 - beta; profiles/tf32_bias_r02.txt).  Chunked layers: beta per main-term
    // instruction of a chunk; single-accumulator layers (all 12 x K / 32 instructions in one accumulator): beta per layer depth.
    static const float beta_env = getenv("DFD_TF32_BETA") ? (float)atof(getenv("DFD_TF32_BETA")) : -1.f;
    // measured shrink of a main-term accumulator: 8.4e-8 / 1.5e-7 / 2.8e-7 after 4 / 8 / 16 accumulations = 2.51e-8 * n^0.87
    p.beta_instr = beta_env >= 0.f ? beta_env : 2.51e-8f;
    p.beta_plain = p.beta_instr * powf((float)((K + 7) / 8), 0.87f);     // the main-term accumulator of a plain layer: same law
    const int nblk32 = (n_pad + 31) / 32;
    // Shallow plain layers with a wide N (the expand convs: one to six k-blocks, 3-5 column blocks of epilogue per tile) are bound by
    // their epilogue warps -- ncu source counters: the two epilogue groups run flat out at ~0.2 IPC per warp (TMEM read, MUFU and
    // shared-memory latencies in one dependent chain per block) while the eight staging warps spin on their barriers 30 % of all
    // samples.  They get a THIRD epilogue group out of the second staging group (DFD_TF32_EPI3=0 switches it off).
    static const bool epi3_off = getenv("DFD_TF32_EPI3") && atoi(getenv("DFD_TF32_EPI3")) == 0;
    p.epi_groups = (!chunked && a_mode == TA_PLAIN && nblk32 >= 3 && !epi3_off) ? 3 : 2;
    const int b_bytes = num_kb * 2 * n_pad * 128;
    p.epi_db = 1;
    // Shared-memory plan.  227 KB per CTA minus the static part (barriers, gate tables: ~5 KB) and the alignment slack; the pipeline
    // wants DEPTH -- every layer shape is bound by the latency of its TMA -> split -> MMA -> release chain, not by bytes or flops:
    // b1 / b2.project went from 171 / 280 us to 136 / 218 us with three stages instead of two -- so: as many stages as fit (odd
    // counts included: a stage is then served by the two staging groups in turn), no cap by the layer's own depth (the ring runs
    // across tiles), and W resident only if that does not cost a stage.
    const int bias_bytes = (n_pad * p.n_blocks * 4 + 127) & ~127;
    const int avail = 227 * 1024 - 6 * 1024 - 1024 - bias_bytes;
    static const bool even_only = getenv("DFD_TF32_EVEN") != nullptr, cap_depth = getenv("DFD_TF32_CAP") != nullptr;
    int staging_bytes = 0, stage_bytes = 0, stages = 0;
    for (;;) {
        p.stg_stride = (!chunked && p.dense_c) ? (uint32_t)((TBLOCK_M * N * 4 + 1023) & ~1023) : (uint32_t)TSTAGING_BLOCK_BYTES;
        staging_bytes = chunked ? nblk32 * TSTAGING_BLOCK_BYTES : 2 * p.epi_groups * (int)p.stg_stride;
        const int st_res = p.n_blocks == 1 ? (avail - staging_bytes - b_bytes) / (2 * TA_BYTES) : 0;
        const int st_str = (avail - staging_bytes) / (2 * TA_BYTES + 2 * n_pad * 128);
        p.b_resident = (st_res >= 2 && st_res >= st_str) ? 1 : 0;
        stage_bytes = 2 * TA_BYTES + (p.b_resident ? 0 : 2 * n_pad * 128);
        stages = p.b_resident ? st_res : st_str;
        if (stages > 8) stages = 8;
        // the stem is bound by its im2col gather, not by pipeline depth: 2 / 3 / 4 stages measured 218 / 218 / 219 us
        static const int stem_st = getenv("DFD_TF32_STEM_STAGES") ? atoi(getenv("DFD_TF32_STEM_STAGES")) : 2;
        if (a_mode == TA_STEM && stages > stem_st) stages = stem_st;
        if (even_only) stages &= ~1;
        if (cap_depth && stages > 2 * ((num_kb + 1) / 2) && stages > 2) { stages = 2 * ((num_kb + 1) / 2); if (stages < 2) stages = 2; }
        if (stages >= 2 || p.epi_groups == 2) break;
        p.epi_groups = 2;                                    // the third group's staging blocks do not fit next to two pipeline stages
    }
    DFD_REQUIRE(stages >= 2, DFD_ERR_INVALID, "gemm_tf32x3: tile does not fit shared memory");
    p.stages = stages;
    static const int pf_env = getenv("DFD_TF32_PF") ? atoi(getenv("DFD_TF32_PF")) : 0;
    p.pf = a_mode == TA_STEM ? 0 : pf_env;
    // OFF by default (DFD_TF32_RAWHI=1): 1.3 % faster (same-box A/B, 4.949 vs 5.012 ms per 256 crops), but the truncated hi leaves a
    // lo twice as large and the whole network's max |dp| goes from 5.5e-6 to 1.2e-5 -- inside the 1e-4 gate, not worth the margin
    static const bool raw_hi_on = getenv("DFD_TF32_RAWHI") && atoi(getenv("DFD_TF32_RAWHI")) != 0;
    p.raw_hi = (a_mode == TA_PLAIN && raw_hi_on) ? 1 : 0;
    p.bias_off = (uint32_t)(stages * stage_bytes + (p.b_resident ? b_bytes : 0) + staging_bytes);
    const size_t smem = (size_t)p.bias_off + bias_bytes + 1024;
    TGemmKernel kern = tgemm_kernel(residual != nullptr, act != 0, p.dense_c != 0);
    int rc;
    if ((rc = dfd_func_smem(ctx, kern, 221 * 1024))) return rc;
    CUtensorMap ma, mbh, mbl, mc;
    if (a_mode != TA_STEM) { if ((rc = make_map_f32(ctx, &ma, A, (uint64_t)M, (uint64_t)K, TBLOCK_M))) return rc; }
    else memset(&ma, 0, sizeof ma);
    if ((rc = make_map_f32(ctx, &mbh, W_hi, (uint64_t)N, (uint64_t)K, (uint32_t)n_pad))) return rc;
    if ((rc = make_map_f32(ctx, &mbl, W_lo, (uint64_t)N, (uint64_t)K, (uint32_t)n_pad))) return rc;
    if ((rc = make_map_f32(ctx, &mc, C, (uint64_t)M, (uint64_t)N, TBLOCK_M))) return rc;
    int grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
    DFD_CUDA(dfd_launch(ctx->pdl, kern, dim3(grid), dim3(TGEMM_THREADS), smem, st, ma, mbh, mbl, mc, p));
    DFD_LAUNCH_CHECK("k_gemm_tf32x3", st);
    return DFD_OK;
}

// host: the two tf32 planes of a weight tensor (round to nearest, ties away -- what cvt.rna.tf32.f32 does)
static inline float tf32_rna_host(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;          // inf / nan
    u = (u + 0x1000u) & 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}
void dfd_tf32_split_host(const float* w, size_t n, float* hi, float* lo) {
    for (size_t i = 0; i < n; i++) {
        const float h = tf32_rna_host(w[i]);
        hi[i] = h;
        lo[i] = tf32_rna_host(w[i] - h);
    }
}

// ---------------------------------------------------------------------------------------------
// self-test against an fp32-FMA CUDA-core reference with fp64 accumulation (tests/test_gpu_gemm.py)
__global__ void k_tgemm_ref(const float* A, const float* se, int hw, const float* W, const float* bias, const float* res,
                            float* C, int M, int N, int K, int act) {
    int n = blockIdx.y * blockDim.x + threadIdx.x, m = blockIdx.x;       // m on grid.x: M reaches 3.2 M rows
    if (n >= N || m >= M) return;
    double acc = 0.0;
    for (int k = 0; k < K; k++) {
        float a = A[(size_t)m * K + k];
        if (se) a = a * se[(size_t)(m / hw) * K + k];
        acc += (double)a * (double)W[(size_t)n * K + k];
    }
    float v = (float)acc + bias[n];
    if (act) v = v / (1.0f + expf(-v));
    if (res) v += res[(size_t)m * N + n];
    C[(size_t)m * N + n] = v;
}
__global__ void k_tfill(float* x, size_t n, uint32_t seed, float lo, float hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    x[i] = lo + (hi - lo) * (float)(h & 0xffffff) / 16777216.0f;
}
__global__ void k_tsplit(const float* w, size_t n, float* hi, float* lo) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float h, l, l2, dummy;
    tf32_split(w[i], h, l);
    tf32_split(l, l2, dummy);
    hi[i] = h; lo[i] = l2;
}
__global__ void k_tmaxerr(const float* c, const float* ref, size_t n, float* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float e = fabsf(c[i] - ref[i]) / fmaxf(1.0f, fabsf(ref[i]));
    if (!(e == e)) e = 1e30f;
    atomicMax((int*)out, __float_as_int(e));
}

// signed error statistics: out[0] += (c - ref) * ref, out[1] += ref^2, out[2] += (c - ref)^2  (diagnostics: is the error a scale bias?)
__global__ void k_tstats(const float* c, const float* ref, size_t n, double* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double a = 0, b = 0, d = 0;
    if (i < n) { const double e = (double)c[i] - (double)ref[i]; a = e * ref[i]; b = (double)ref[i] * ref[i]; d = e * e; }
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); d += __shfl_xor_sync(0xffffffffu, d, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); atomicAdd(out + 2, d); }
}

// mode bit 0: residual; bit 1: SE-gated A with hw = 49; bit 2: *ms_host receives the SCALE BIAS sum((c-ref) ref) / sum(ref^2) and
// *max_err_host the RMS relative error sqrt(sum (c-ref)^2 / sum ref^2) instead (diagnostics).  iters > 0: also time `iters` launches (mean ms -> *ms_host).
extern "C" int dfd_gemm_tf32_selftest(dfd_ctx* ctx, int M, int N, int K, int act, int mode, int iters, double* max_err_host,
                                      double* ms_host, void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    DfdDeviceGuard dev_guard(ctx->cfg.device);
    cudaStream_t st = (cudaStream_t)stream;
    const bool use_res = mode & 1, use_se = (mode & 2) != 0;
    const int hw = 49, n_img = (M + hw - 1) / hw;
    float *A, *W, *Wh, *Wl, *R, *C, *bias, *ref, *err, *se;
    DFD_CUDA(cudaMalloc(&A, (size_t)M * K * 4));
    DFD_CUDA(cudaMalloc(&W, (size_t)N * K * 4));
    DFD_CUDA(cudaMalloc(&Wh, (size_t)N * K * 4));
    DFD_CUDA(cudaMalloc(&Wl, (size_t)N * K * 4));
    DFD_CUDA(cudaMalloc(&R, (size_t)M * N * 4));
    DFD_CUDA(cudaMalloc(&C, (size_t)M * N * 4));
    DFD_CUDA(cudaMalloc(&bias, (size_t)(N + 512) * 4));
    DFD_CUDA(cudaMalloc(&se, (size_t)n_img * K * 4));
    DFD_CUDA(cudaMalloc(&ref, (size_t)M * N * 4));
    DFD_CUDA(cudaMalloc(&err, 4));
    DFD_CUDA(cudaMemsetAsync(err, 0, 4, st));
    DFD_CUDA(cudaMemsetAsync(C, 0xff, (size_t)M * N * 4, st));
    DFD_CUDA(cudaMemsetAsync(bias, 0, (size_t)(N + 512) * 4, st));
    k_tfill<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, st>>>(A, (size_t)M * K, 11u, -1.0f, 1.0f);
    k_tfill<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, st>>>(W, (size_t)N * K, 22u, -0.25f, 0.25f);
    k_tfill<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, st>>>(R, (size_t)M * N, 33u, -1.0f, 1.0f);
    k_tfill<<<(N + 255) / 256, 256, 0, st>>>(bias, (size_t)N, 44u, -0.5f, 0.5f);
    k_tfill<<<(unsigned)(((size_t)n_img * K + 255) / 256), 256, 0, st>>>(se, (size_t)n_img * K, 55u, 0.05f, 1.0f);
    k_tsplit<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, st>>>(W, (size_t)N * K, Wh, Wl);
    int rc = dfd_gemm_tf32x3(ctx, use_se ? TA_SCALE : TA_PLAIN, A, use_se ? se : nullptr, hw, Wh, Wl, bias, use_res ? R : nullptr, C, M, N, K, act, st);
    if (rc == DFD_OK) {
        k_tgemm_ref<<<dim3(M, (N + 127) / 128), 128, 0, st>>>(A, use_se ? se : nullptr, hw, W, bias, use_res ? R : nullptr, ref, M, N, K, act);
        k_tmaxerr<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, st>>>(C, ref, (size_t)M * N, err);
        float e = 0.f;
        cudaError_t ce = cudaMemcpyAsync(&e, err, 4, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { ctx->err = std::string("gemm tf32 selftest: ") + cudaGetErrorString(ce); rc = DFD_ERR_CUDA; }
        if (max_err_host) *max_err_host = (double)e;
        if (mode & 4) {
            double* dst;
            double hs[3] = {0, 0, 0};
            if (cudaMalloc(&dst, 24) == cudaSuccess) {
                cudaMemsetAsync(dst, 0, 24, st);
                k_tstats<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, st>>>(C, ref, (size_t)M * N, dst);
                cudaMemcpyAsync(hs, dst, 24, cudaMemcpyDeviceToHost, st);
                cudaStreamSynchronize(st);
                cudaFree(dst);
                if (ms_host) *ms_host = hs[0] / (hs[1] > 0 ? hs[1] : 1);
                if (max_err_host) *max_err_host = sqrt(hs[2] / (hs[1] > 0 ? hs[1] : 1));
            }
            iters = 0;
        }
    }
    if (rc == DFD_OK && iters > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < iters + 2 && rc == DFD_OK; i++) {
            if (i == 2) cudaEventRecord(e0, st);
            rc = dfd_gemm_tf32x3(ctx, use_se ? TA_SCALE : TA_PLAIN, A, use_se ? se : nullptr, hw, Wh, Wl, bias, use_res ? R : nullptr, C, M, N, K, act, st);
        }
        cudaEventRecord(e1, st);
        cudaError_t ce = cudaStreamSynchronize(st);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ce != cudaSuccess) { ctx->err = std::string("gemm tf32 bench: ") + cudaGetErrorString(ce); rc = DFD_ERR_CUDA; }
        if (ms_host) *ms_host = (double)ms / iters;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaFree(A); cudaFree(W); cudaFree(Wh); cudaFree(Wl); cudaFree(R); cudaFree(C); cudaFree(bias); cudaFree(ref); cudaFree(err); cudaFree(se);
    return rc;
}
