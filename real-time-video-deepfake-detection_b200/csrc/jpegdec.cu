// Frame ingest on the device: cv2.imdecode(IMREAD_COLOR) of the /analyze wire format (baseline JPEG; reference
// backend_server.py:140-142, extension/content.js:86-109) for a batch of frames, so that the host-to-device copy carries the
// ~100-400 KB JPEG stream instead of the 2.76 MB raw 720p frame (raw frames make `e2e` PCIe-bound).
// Bit-exact with OpenCV / libjpeg-turbo (islow IDCT, fancy up-sampling, fixed-point YCbCr): the pixel functions are those of
// px_jpeg.h / px_jpegdec.h, which tests/hostcheck runs on the CPU against cv2.imdecode.
//
//   host            dfd_jpeg_parse: markers -> DfdJpegHeader (quantisers, Huffman lookup tables, geometry); no pixel work
//   k_ju_count / k_ju_scan / k_ju_scatter   removal of the FF 00 byte stuffing of the entropy-coded segments, flat over
//                   2 KB chunks of all frames (count, per-frame scan, scatter into MSB-first 32-bit words)
//   SELF-SYNCHRONISING parallel Huffman decode, flat over the 2048-bit subsequences of ALL frames (one thread each), so a
//   700 KB noise frame and a 10 KB flat frame in the same batch load the chip evenly:
//     k_jh_blind    every thread decodes its subsequence from the state "a block starts here"; Huffman streams
//                   resynchronise, so most END states are already right
//     k_jh_round    x 4 (DFD_JH_ROUNDS): thread i re-decodes subsequence i from the END state of subsequence i-1 whenever that
//                   changed (a frame whose previous round changed nothing exits at once).  When a whole round changes
//                   nothing, by induction from subsequence 0 (whose start is known) every state is exact.
//     k_jh_finish   CTA per frame, only for frames still changing after the flat rounds (pathological streams): rounds in a
//                   loop with CTA barriers until nothing changes -- correctness never depends on the round count
//     k_jh_scan     per frame: exclusive scan of the per-subsequence block counts = first block number of each subsequence
//     k_jh_write    last pass: coefficients (AC values in natural order, DC differences in prediction-chain order)
//     k_jh_dc       DC prediction chains = prefix sums per component
//   k_jpeg_idct     thread per 8 x 8 block: dequantise, jidctint (columns, rows), +128, clamp -> component planes
//   k_jpeg_color    fancy h2v2 / h2v1 chroma up-sampling + YCbCr -> BGR into the caller's frame buffer
// The device bit reader (64-bit register window over the clean words) and symbol loop below are a faster restatement of
// dfd_jpeg_step / dfd_jpeg_decode_sub (px_jpegdec.h), which tests/hostcheck checks on the CPU; the GPU tests compare the
// kernels' output with cv2.imdecode bit for bit.
#include "dfd_internal.cuh"
#include <string.h>
#include <stdlib.h>
#include "px_jpegdec.h"

#define JPG_SUB_BITS 2048
#define JPG_THREADS 512
#define JH_THREADS 128
#define JH_ROUNDS 12                  // upper bound of the flat rounds (array sizing)
#define JH_ROUNDS_DEFAULT 4           // measured on 256 x 720p: rounds + finish 2.75 ms at 12, 2.51 at 6, 2.43 at 3 (bench mix); natural frames 1.21 -> 1.16 ms at 4
#define JU_CHUNK 2048                 // raw bytes per unstuffing chunk = 128 threads x 16 bytes

struct JpgMeta {
    long long raw_off;        // first byte of the stream in the raw buffer
    long long words_off;      // first 32-bit word of the frame's clean bit stream
    long long sub_off;        // first subsequence slot of the frame
    int ecs_bytes;            // raw size of the entropy-coded segment
    int chunk_off;            // first unstuffing-chunk slot of the frame
};

// exclusive scan of one int per thread over the CTA (JPG_THREADS threads); returns the exclusive prefix, *total = CTA sum
__device__ int cta_exscan(int v, int* total, int* s_warp /* [JPG_THREADS / 32 + 1] */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __syncthreads();                                           // s_warp may still be read from the previous call
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        int t = lane < JPG_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
        if (lane < JPG_THREADS / 32) s_warp[lane] = t;         // inclusive warp totals
    }
    __syncthreads();
    const int base = w ? s_warp[w - 1] : 0;
    *total = s_warp[JPG_THREADS / 32 - 1];
    return base + x - v;
}

// ---------------------------------------------------------------------------------------------
// Byte-stuffing removal, flat over JU_CHUNK-byte chunks: grid (chunks, frames), 128 threads x 16 bytes.
__device__ __forceinline__ int ju_keep_mask(const uint8_t* __restrict__ src, int n, int i0, uint8_t (&b)[16]) {
    int keep = 0;
    uint8_t prev = (i0 > 0 && i0 - 1 < n) ? __ldg(src + i0 - 1) : 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int i = i0 + j;
        b[j] = i < n ? __ldg(src + i) : 0;
        const bool stuffed = b[j] == 0 && prev == 0xFF;        // a zero that follows FF carries no data
        if (i < n && !stuffed) keep |= 1 << j;
        prev = b[j];
    }
    return keep;
}

__global__ void __launch_bounds__(128)
k_ju_count(const uint8_t* __restrict__ raw, const JpgMeta* __restrict__ meta, const DfdJpegHeader* __restrict__ hdr, int* __restrict__ chunk_cnt) {
    const int f = blockIdx.y;
    const JpgMeta M = meta[f];
    const int c0 = blockIdx.x * JU_CHUNK;
    if (c0 >= M.ecs_bytes) return;
    const uint8_t* src = raw + M.raw_off + hdr[f].ecs_begin;
    uint8_t b[16];
    int cnt = __popc(ju_keep_mask(src, M.ecs_bytes, c0 + threadIdx.x * 16, b));
    __shared__ int s[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) chunk_cnt[M.chunk_off + blockIdx.x] = s[0] + s[1] + s[2] + s[3];
}

// per frame: chunk counts -> chunk bases (exclusive scan), clean length, zeroed tail + guard words
__global__ void __launch_bounds__(JPG_THREADS)
k_ju_scan(const JpgMeta* __restrict__ meta, int* __restrict__ chunk_cnt, uint32_t* __restrict__ words, uint32_t* __restrict__ nbits_out) {
    __shared__ int s_warp[JPG_THREADS / 32 + 1];
    __shared__ int s_carry;
    const int f = blockIdx.x;
    const JpgMeta M = meta[f];
    const int nch = (M.ecs_bytes + JU_CHUNK - 1) / JU_CHUNK;
    int* cc = chunk_cnt + M.chunk_off;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < nch; c0 += JPG_THREADS) {
        const int i = c0 + threadIdx.x;
        const int v = i < nch ? cc[i] : 0;
        int total;
        const int ex = cta_exscan(v, &total, s_warp);
        if (i < nch) cc[i] = s_carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    const int nb = s_carry;
    uint8_t* dst = (uint8_t*)(words + M.words_off);
    if (threadIdx.x < 20) {                                    // the bit reader runs up to three words ahead
        const int o = nb + threadIdx.x;
        if (o < ((nb + 3) & ~3) + 16) dst[(o & ~3) + (3 - (o & 3))] = 0;
    }
    if (threadIdx.x == 0) nbits_out[f] = (uint32_t)nb * 8u;
}

__global__ void __launch_bounds__(128)
k_ju_scatter(const uint8_t* __restrict__ raw, const JpgMeta* __restrict__ meta, const DfdJpegHeader* __restrict__ hdr,
             const int* __restrict__ chunk_base, uint32_t* __restrict__ words) {
    const int f = blockIdx.y;
    const JpgMeta M = meta[f];
    const int c0 = blockIdx.x * JU_CHUNK;
    if (c0 >= M.ecs_bytes) return;
    const uint8_t* src = raw + M.raw_off + hdr[f].ecs_begin;
    uint8_t* dst = (uint8_t*)(words + M.words_off);
    uint8_t b[16];
    const int keep = ju_keep_mask(src, M.ecs_bytes, c0 + threadIdx.x * 16, b);
    const int cnt = __popc(keep), lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __shared__ int s[4];
    if (lane == 31) s[w] = x;
    __syncthreads();
    int off = chunk_base[M.chunk_off + blockIdx.x] + x - cnt;
    for (int q = 0; q < w; q++) off += s[q];
#pragma unroll
    for (int j = 0; j < 16; j++)
        if (keep >> j & 1) { dst[(off & ~3) + (3 - (off & 3))] = b[j]; off++; }            // MSB-first inside each 32-bit word
}

// ---------------------------------------------------------------------------------------------
// Device bit reader: a 64-bit window (next bit = bit 63) over the clean MSB-first words.  Guard words follow the data.
struct JhBits { uint64_t buf; int nb; uint32_t wi; };
__device__ __forceinline__ void jh_init(JhBits& r, const uint32_t* __restrict__ w, uint32_t p) {
    const uint32_t i = p >> 5;
    r.buf = (((uint64_t)__ldg(w + i) << 32) | __ldg(w + i + 1)) << (p & 31u);
    r.nb = 64 - (int)(p & 31u);
    r.wi = i + 2;
}
__device__ __forceinline__ void jh_refill(JhBits& r, const uint32_t* __restrict__ w) {
    if (r.nb <= 32) { r.buf |= (uint64_t)__ldg(w + r.wi) << (32 - r.nb); r.wi++; r.nb += 32; }
}

// Per-CTA decode tables in shared memory (a CTA of the flat kernels works on ONE frame): the four 8-bit look-ahead tables and,
// per block-in-MCU index, which DC / AC table its component uses.
struct JhTabs {
    uint16_t look[4][256];                // dc0, dc1, ac0, ac1
    uint8_t sel[16];                      // bit 0: DC table, bit 1: AC table of block-in-MCU c
    uint8_t zz[64];
    // write pass: where block-in-MCU j of MCU (mx, my) lives.  coefficient block index = base[j] + my * ystep[j] + mx * xstep[j],
    // DC chain slot = dcb[j] + (my * mcus_x + mx) * nbk[j]   (dfd_jpeg_block_pos without its divisions)
    int32_t base[DFD_JPEG_MAX_BPM], ystep[DFD_JPEG_MAX_BPM], xstep[DFD_JPEG_MAX_BPM], dcb[DFD_JPEG_MAX_BPM], nbk[DFD_JPEG_MAX_BPM];
    int32_t mcus_x, bpm, total_blocks;
};
__device__ __forceinline__ void jh_load_tabs(JhTabs& T, const DfdJpegHeader* __restrict__ h, int tid, int nthreads) {
    for (int i = tid; i < 4 * 128; i += nthreads) {            // 2 u16 per thread and step
        const int t = i >> 7, j = i & 127;
        const DfdHuffTab* src = t < 2 ? &h->dc[t] : &h->ac[t - 2];
        ((uint32_t*)T.look[t])[j] = ((const uint32_t*)src->look)[j];
    }
    if (tid < 16) { const int comp = tid < h->bpm ? h->blk_comp[tid] : 0; T.sel[tid] = (uint8_t)(h->dc_tab[comp] | (h->ac_tab[comp] << 1)); }
    if (tid < 64) T.zz[tid] = DFD_ZIGZAG_DEV[tid];
    if (tid < h->bpm) {
        const int c = h->blk_comp[tid], jj = tid - h->blk_first[c];
        int dc_off = 0;
        for (int q = 0; q < c; q++) dc_off += h->comp_bw[q] * h->comp_bh[q];
        T.base[tid] = h->comp_blk0[c] + (jj / h->hs[c]) * h->comp_bw[c] + jj % h->hs[c];
        T.ystep[tid] = h->vs[c] * h->comp_bw[c];
        T.xstep[tid] = h->hs[c];
        T.nbk[tid] = h->hs[c] * h->vs[c];
        T.dcb[tid] = dc_off + jj;
    }
    if (tid == 0) { T.mcus_x = h->mcus_x; T.bpm = h->bpm; T.total_blocks = h->total_blocks; }
    __syncthreads();
}

// dfd_jpeg_decode_sub, restated on the register bit window (same state transitions, same outputs).  `rem` counts the bits
// left before the subsequence's limit: a symbol belongs to the subsequence in which it STARTS.
template <bool WRITE>
__device__ __forceinline__ unsigned long long jh_decode(const DfdJpegHeader* __restrict__ h, const uint32_t* __restrict__ words,
                                                        unsigned long long start, uint32_t limit, int* nblk, int blk0,
                                                        int16_t* __restrict__ coef, int32_t* __restrict__ dcd, const int32_t* dc_off,
                                                        const JhTabs& T, uint32_t* s_blk = nullptr) {
    const uint32_t p0 = (uint32_t)(start >> 16);
    int c = (int)((start >> 8) & 255u), z = (int)(start & 255u);
    int rem = (int)(limit - p0);
    int blocks = 0;
    const int bpm = T.bpm;
    int sel = T.sel[c];
    const uint16_t* ldc = T.look[sel & 1];
    const uint16_t* lac = T.look[2 + (sel >> 1)];
    // write position: block number blk0 = MCU (mx, my), block-in-MCU c; advanced incrementally
    int blk = blk0, index = 0, dslot = 0, mx = 0, my = 0;
    if (WRITE) {
        const int mcu = blk / bpm;
        my = mcu / T.mcus_x; mx = mcu - my * T.mcus_x;
        index = T.base[c] + my * T.ystep[c] + mx * T.xstep[c];
        dslot = T.dcb[c] + mcu * T.nbk[c];
    }
    bool wr = WRITE && blk < T.total_blocks;
    // Write pass: a block that STARTS in this subsequence is assembled in the thread's private 128-byte slice of shared memory
    // (32 words, interleaved over the CTA's threads: conflict-free) and leaves as eight 16-byte stores when it completes;
    // the block in progress at the subsequence's start (shared with the previous thread) and the one left unfinished at its end
    // are written coefficient by coefficient into the zero-initialised array.  (Scattered 2-byte stores for everything cost
    // 1.1 of the write pass's 1.9 ms per 256 natural 720p frames.)
    bool own = WRITE && z == 0;
    unsigned long long nzmask = 0ull;                           // natural positions of the owned block that hold a coefficient
    int16_t* my16 = (int16_t*)s_blk;
    const int tid2 = threadIdx.x * 2;
    JhBits r;
    jh_init(r, words, p0);
    while (rem > 0) {
        jh_refill(r, words);
        const bool is_dc = z == 0;
        const uint32_t e = (is_dc ? ldc : lac)[(uint32_t)(r.buf >> 56)];
        int len, sym;
        if (e) { len = (int)(e >> 8); sym = (int)(e & 255u); }
        else {                                                  // code longer than 8 bits: canonical search (jdhuff.c)
            const DfdHuffTab* t = is_dc ? &h->dc[sel & 1] : &h->ac[sel >> 1];
            const uint32_t b16 = (uint32_t)(r.buf >> 48);
            int l = 9;
            int32_t code = (int32_t)(b16 >> 7);
            while (l <= 16 && code > __ldg(&t->maxcode[l])) { l++; code = (int32_t)(b16 >> (16 - l)); }
            if (l > 16) { len = 1; sym = 0; }                   // invalid code: advance one bit (a blind decoder must not stall)
            else { len = l; sym = __ldg(&t->huffval[(code + __ldg(&t->valoff[l])) & 255]); }
        }
        r.buf <<= len;
        const int size = sym & 15;
        int val = 0;
        if (size) {
            const int v = (int)(r.buf >> (64 - size));
            r.buf <<= size;
            val = v - (((v >> (size - 1)) ^ 1) * ((1 << size) - 1));     // EXTEND: v < 2^(size-1) ? v - (2^size - 1) : v
        }
        const int used = len + size;
        r.nb -= used; rem -= used;
        int k = -1;
        if (is_dc) { k = 0; z = 1; }
        else if (size) { z += sym >> 4; if (z < 64) k = z; z++; }
        else z = (sym >> 4) == 15 ? z + 16 : 64;
        if (WRITE && k >= 0 && wr) {
            if (k == 0) dcd[dslot] = val;
            else {
                const int nat = T.zz[k];
                if (own) { my16[(nat >> 1) * (2 * JH_THREADS) + tid2 + (nat & 1)] = (int16_t)val; nzmask |= 1ull << nat; }
                else coef[(size_t)index * 64 + nat] = (int16_t)val;
            }
        }
        if (z >= 64) {
            z = 0;
            c = c + 1 == bpm ? 0 : c + 1;
            blocks++;
            sel = T.sel[c];
            ldc = T.look[sel & 1];
            lac = T.look[2 + (sel >> 1)];
            if (WRITE) {
                if (own && __popcll(nzmask) <= 6) {              // sparse block: the few coefficients one by one (array is zero-initialised)
                    while (nzmask) {
                        const int nat = __ffsll((long long)nzmask) - 1;
                        nzmask &= nzmask - 1;
                        const int slot = (nat >> 1) * (2 * JH_THREADS) + tid2 + (nat & 1);
                        if (wr) coef[(size_t)index * 64 + nat] = my16[slot];
                        my16[slot] = 0;
                    }
                } else if (own) {                               // dense block: flush it whole (zeros included) and clear the slice
                    nzmask = 0ull;
                    uint4* dst = (uint4*)(coef + (size_t)index * 64);
#pragma unroll
                    for (int w = 0; w < 8; w++) {
                        uint4 v;
                        v.x = s_blk[(4 * w) * JH_THREADS + threadIdx.x]; v.y = s_blk[(4 * w + 1) * JH_THREADS + threadIdx.x];
                        v.z = s_blk[(4 * w + 2) * JH_THREADS + threadIdx.x]; v.w = s_blk[(4 * w + 3) * JH_THREADS + threadIdx.x];
                        if (wr) dst[w] = v;
                        s_blk[(4 * w) * JH_THREADS + threadIdx.x] = 0u; s_blk[(4 * w + 1) * JH_THREADS + threadIdx.x] = 0u;
                        s_blk[(4 * w + 2) * JH_THREADS + threadIdx.x] = 0u; s_blk[(4 * w + 3) * JH_THREADS + threadIdx.x] = 0u;
                    }
                }
                own = true;
                blk++;
                wr = blk < T.total_blocks;
                if (c == 0) { if (++mx == T.mcus_x) { mx = 0; my++; } }
                index = T.base[c] + my * T.ystep[c] + mx * T.xstep[c];
                dslot = T.dcb[c] + (my * T.mcus_x + mx) * T.nbk[c];
            }
        }
    }
    if (WRITE && own && wr && z > 0) {                         // unfinished block: its coefficients so far, one by one
#pragma unroll 4
        for (int w = 0; w < 32; w++) {
            const uint32_t v = s_blk[w * JH_THREADS + threadIdx.x];
            if (v & 0xffffu) coef[(size_t)index * 64 + 2 * w] = (int16_t)(v & 0xffffu);
            if (v >> 16) coef[(size_t)index * 64 + 2 * w + 1] = (int16_t)(v >> 16);
        }
    }
    *nblk = blocks;
    const uint32_t p = limit - (uint32_t)rem;                   // rem <= 0: the last symbol may end beyond the limit
    return ((unsigned long long)p << 16) | ((unsigned long long)(uint32_t)c << 8) | (unsigned long long)(uint32_t)z;
}

struct JhArgs {
    const JpgMeta* meta; const DfdJpegHeader* hdr; const uint32_t* words; const uint32_t* nbits;
    unsigned long long* E; unsigned long long* used; int* cnt; int* blk0; int* changed;
    int16_t* coef; int32_t* dc; long long blocks_stride;
    int last_round;                   // index of the last flat synchronisation round that was launched
};

// grid (subsequence groups, frames); mode 0: blind pass, 1: synchronisation round `round`, 2: write pass
template <int MODE>
__global__ void __launch_bounds__(JH_THREADS) k_jh_pass(const JhArgs a, int round) {
    __shared__ JhTabs T;
    __shared__ uint32_t s_blk[MODE == 2 ? 32 * JH_THREADS : 1];
    const int f = blockIdx.y;
    if (MODE == 2) {
#pragma unroll
        for (int w = 0; w < 32; w++) s_blk[w * JH_THREADS + threadIdx.x] = 0u;       // (each thread only ever touches its own slice)
    }
    if (MODE == 1 && round > 0 && a.changed[f * JH_ROUNDS + round - 1] == 0) return;        // this frame has converged
    const uint32_t nbits = a.nbits[f];
    const int nsub = (int)((nbits + JPG_SUB_BITS - 1) / JPG_SUB_BITS);
    if (blockIdx.x * JH_THREADS >= nsub) return;                // whole CTA (barriers below)
    const DfdJpegHeader* h = a.hdr + f;
    const int i = blockIdx.x * JH_THREADS + threadIdx.x;
    const JpgMeta M = a.meta[f];
    const uint32_t* words = a.words + M.words_off;
    unsigned long long* E = a.E + M.sub_off;
    unsigned long long* used = a.used + M.sub_off;
    unsigned long long st1 = 0;
    if (MODE == 1) {                                            // a CTA none of whose subsequences has a new start state is done
        int need = 0;
        if (i > 0 && i < nsub) { st1 = __ldcg(E + i - 1); need = st1 != used[i]; }   // one 64-bit word: seen whole, old or new
        if (!__syncthreads_or(need)) return;
        jh_load_tabs(T, h, threadIdx.x, JH_THREADS);
        if (!need) return;
    } else {
        jh_load_tabs(T, h, threadIdx.x, JH_THREADS);
        if (i >= nsub) return;
    }
    const uint32_t lim = min((uint32_t)(i + 1) * JPG_SUB_BITS, nbits);
    int nb;
    if (MODE == 0) {
        const unsigned long long st = (unsigned long long)((uint32_t)i * JPG_SUB_BITS) << 16;
        const unsigned long long e = jh_decode<false>(h, words, st, lim, &nb, 0, nullptr, nullptr, nullptr, T);
        used[i] = st; E[i] = e; a.cnt[M.sub_off + i] = nb;
    } else if (MODE == 1) {
        const unsigned long long st = st1;
        const unsigned long long e = jh_decode<false>(h, words, st, lim, &nb, 0, nullptr, nullptr, nullptr, T);
        if (e != E[i]) a.changed[f * JH_ROUNDS + round] = 1;
        used[i] = st; __stcg(E + i, e); a.cnt[M.sub_off + i] = nb;
    } else {
        // Write pass: coefficients leave as scattered 2-byte stores into the zero-initialised array.  (Staging a CTA's block
        // range in 46 KB of shared memory and writing whole 128-byte blocks was measured and is SLOWER, 1.9 -> 2.9 ms per
        // 256 natural frames: the decode loop is latency-bound and lives on occupancy, which the buffer cut to 16 warps / SM.)
        int32_t dc_off[3] = {0, 0, 0};
        { int q = 0; for (int c = 0; c < h->ncomp; c++) { dc_off[c] = q; q += h->comp_bw[c] * h->comp_bh[c]; } }
        jh_decode<true>(h, words, used[i], lim, &nb, a.blk0[M.sub_off + i], a.coef + (size_t)f * a.blocks_stride * 64,
                        a.dc + (size_t)f * a.blocks_stride, dc_off, T, s_blk);
    }
}

// CTA per frame, only frames that were still changing in the last flat round: rounds until a whole round changes nothing
__global__ void __launch_bounds__(JPG_THREADS) k_jh_finish(const JhArgs a) {
    __shared__ JhTabs T;
    const int f = blockIdx.x, tid = threadIdx.x;
    if (a.changed[f * JH_ROUNDS + a.last_round] == 0) return;
    const JpgMeta M = a.meta[f];
    const DfdJpegHeader* h = a.hdr + f;
    jh_load_tabs(T, h, tid, JPG_THREADS);
    const uint32_t* words = a.words + M.words_off;
    const uint32_t nbits = a.nbits[f];
    const int nsub = (int)((nbits + JPG_SUB_BITS - 1) / JPG_SUB_BITS);
    volatile unsigned long long* E = a.E + M.sub_off;
    volatile unsigned long long* used = a.used + M.sub_off;
    for (int round = 0; round <= nsub + 1; round++) {
        __syncthreads();
        int changed = 0;
        for (int i = tid; i < nsub; i += JPG_THREADS) {
            if (i == 0) continue;
            const unsigned long long st = E[i - 1];
            if (st == used[i]) continue;
            int nb;
            const unsigned long long e = jh_decode<false>(h, words, st, min((uint32_t)(i + 1) * JPG_SUB_BITS, nbits), &nb, 0, nullptr, nullptr, nullptr, T);
            if (e != E[i]) changed = 1;
            used[i] = st; E[i] = e; a.cnt[M.sub_off + i] = nb;
        }
        if (!__syncthreads_or(changed)) break;
    }
}

// per frame: first block number of every subsequence (exclusive scan of the block counts), completeness check
__global__ void __launch_bounds__(JPG_THREADS) k_jh_scan(const JhArgs a, int32_t* __restrict__ status) {
    __shared__ int s_warp[JPG_THREADS / 32 + 1];
    __shared__ int s_carry;
    const int f = blockIdx.x, tid = threadIdx.x;
    const JpgMeta M = a.meta[f];
    const uint32_t nbits = a.nbits[f];
    const int nsub = (int)((nbits + JPG_SUB_BITS - 1) / JPG_SUB_BITS);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < nsub; c0 += JPG_THREADS) {
        const int i = c0 + tid;
        const int v = i < nsub ? a.cnt[M.sub_off + i] : 0;
        int total;
        const int ex = cta_exscan(v, &total, s_warp);
        if (i < nsub) a.blk0[M.sub_off + i] = s_carry + ex;
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (tid == 0) status[f] = s_carry >= a.hdr[f].total_blocks ? DFD_JPEG_OK : DFD_JPEG_ERR_DATA;
}

// per frame: DC prediction chains = inclusive prefix sums per component, in chain order
__global__ void __launch_bounds__(JPG_THREADS) k_jh_dc(const JhArgs a) {
    __shared__ int s_warp[JPG_THREADS / 32 + 1];
    __shared__ int s_carry;
    const int f = blockIdx.x, tid = threadIdx.x;
    const DfdJpegHeader* h = a.hdr + f;
    int32_t* dcd = a.dc + (size_t)f * a.blocks_stride;
    int off = 0;
    for (int c = 0; c < h->ncomp; c++) {
        const int n = h->comp_bw[c] * h->comp_bh[c];
        int32_t* d = dcd + off;
        off += n;
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (int c0 = 0; c0 < n; c0 += JPG_THREADS) {
            const int i = c0 + tid;
            const int v = i < n ? d[i] : 0;
            int total;
            const int ex = cta_exscan(v, &total, s_warp);
            if (i < n) d[i] = s_carry + ex + v;
            __syncthreads();
            if (tid == 0) s_carry += total;
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// thread = one 8x8 block.  Planes of frame f: component c at plane_base + plane_off[c], pitch comp_bw[c] * 8.
__global__ void __launch_bounds__(128)
k_jpeg_idct(const DfdJpegHeader* __restrict__ hdr, const int16_t* __restrict__ coef_all, const int32_t* __restrict__ dc_all,
            long long blocks_stride, uint8_t* __restrict__ planes_all, long long plane_stride) {
    const int f = blockIdx.y;
    const DfdJpegHeader* h = hdr + f;
    const int blk = blockIdx.x * 128 + threadIdx.x;
    if (blk >= h->total_blocks) return;
    int c = 0;
    if (h->ncomp == 3) c = blk >= h->comp_blk0[2] ? 2 : (blk >= h->comp_blk0[1] ? 1 : 0);
    const int j = blk - h->comp_blk0[c];
    const int by = j / h->comp_bw[c], bx = j - by * h->comp_bw[c];
    int dc_off = 0;
    for (int q = 0; q < c; q++) dc_off += h->comp_bw[q] * h->comp_bh[q];
    const int16_t* cf = coef_all + ((size_t)f * blocks_stride + blk) * 64;
    const int dc = dc_all[(size_t)f * blocks_stride + dc_off + dfd_jpeg_dc_seq(h, c, bx, by)];
    int b[64];
    const uint4* cv = (const uint4*)cf;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint4 v = cv[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            b[i * 8 + 2 * k] = (int)(short)(w[k] & 0xffffu) * (int)h->qt[c][i * 8 + 2 * k];
            b[i * 8 + 2 * k + 1] = (int)(short)(w[k] >> 16) * (int)h->qt[c][i * 8 + 2 * k + 1];
        }
    }
    b[0] = dc * (int)h->qt[c][0];
#pragma unroll
    for (int q = 0; q < 8; q++) dfd_idct8(b + q, 8, 1);
#pragma unroll
    for (int r = 0; r < 8; r++) dfd_idct8(b + 8 * r, 1, 0);
    size_t poff = 0;
    for (int q = 0; q < c; q++) poff += (size_t)h->comp_bw[q] * h->comp_bh[q] * 64;
    const int pitch = h->comp_bw[c] * 8;
    uint8_t* out = planes_all + (size_t)f * plane_stride + poff + (size_t)(by * 8) * pitch + bx * 8;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            lo |= (uint32_t)dfd_sat_u8(b[r * 8 + k] + 128) << (8 * k);
            hi |= (uint32_t)dfd_sat_u8(b[r * 8 + 4 + k] + 128) << (8 * k);
        }
        *(uint2*)(out + (size_t)r * pitch) = make_uint2(lo, hi);
    }
}

// 4:2:0 (the wire format): thread = 8 horizontally adjacent pixels of TWO rows (2r, 2r + 1: they share chroma row r).
// Per plane the thread needs chroma columns cx0 - 1 .. cx0 + 4 of rows r - 1, r, r + 1: one aligned 32-bit load + two edge
// bytes per row (18 loads for 16 pixels; the 4-pixel kernel below issues 68), luma as two 8-byte loads, and the 2 x 24 output
// bytes leave as 8-byte stores (consecutive lanes write consecutive 24-byte runs).  Same arithmetic as jdsample.c
// h2v2_fancy_upsample: vertical 3:1 blend first, then horizontal 3:1 with the +8 / +7 rounding pair; clamping the neighbour
// index at the plane's edges reproduces libjpeg's edge special cases exactly ((3t + t + 8) >> 4 == (4t + 8) >> 4).
__global__ void __launch_bounds__(256)
k_jpeg_color420(const DfdJpegHeader* __restrict__ hdr, const uint8_t* __restrict__ planes_all, long long plane_stride,
                uint8_t* __restrict__ frames, size_t frame_stride, int row_pitch, int H, int W) {
    const int f = blockIdx.z;
    const DfdJpegHeader* h = hdr + f;
    const int xg = blockIdx.x * 64 + (threadIdx.x & 63), x0 = xg * 8;
    const int r = blockIdx.y * 4 + (threadIdx.x >> 6), y0 = r * 2;
    if (y0 >= H || x0 >= W) return;
    const uint8_t* P = planes_all + (size_t)f * plane_stride;
    const int pitch0 = h->comp_bw[0] * 8, pitch1 = h->comp_bw[1] * 8;
    const uint8_t* p1 = P + (size_t)h->comp_bw[0] * h->comp_bh[0] * 64;
    const uint8_t* p2 = p1 + (size_t)h->comp_bw[1] * h->comp_bh[1] * 64;
    const int cw = (W + 1) >> 1, chh = (H + 1) >> 1;
    const int cx0 = xg * 4;
    const int rm = max(r - 1, 0), rp = min(r + 1, chh - 1);
    const bool interior = cx0 >= 1 && cx0 + 4 <= cw - 1;
    int te[2][6], to[2][6];                                 // vertical blends for the even / odd output row, per plane
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
        const uint8_t* base = pl ? p2 : p1;
        int v[3][6];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const uint8_t* row = base + (size_t)(k == 0 ? rm : (k == 1 ? r : rp)) * pitch1;
            if (interior) {
                const uint32_t w = *(const uint32_t*)(row + cx0);          // cx0 % 4 == 0, pitch1 % 8 == 0: aligned
                v[k][0] = row[cx0 - 1]; v[k][5] = row[cx0 + 4];
                v[k][1] = w & 255u; v[k][2] = (w >> 8) & 255u; v[k][3] = (w >> 16) & 255u; v[k][4] = w >> 24;
            } else {
#pragma unroll
                for (int q = 0; q < 6; q++) v[k][q] = row[min(max(cx0 - 1 + q, 0), cw - 1)];
            }
        }
#pragma unroll
        for (int q = 0; q < 6; q++) { te[pl][q] = 3 * v[1][q] + v[0][q]; to[pl][q] = 3 * v[1][q] + v[2][q]; }
    }
    const bool al8 = (((uintptr_t)frames | frame_stride | (size_t)row_pitch) & 7) == 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int y = y0 + half;
        if (y >= H) break;
        const uint2 yw = *(const uint2*)(P + (size_t)y * pitch0 + x0);
        uint8_t px[24];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int q = 1 + (i >> 1);
            int cbv, crv;
            if (half == 0) {
                cbv = (i & 1) ? (te[0][q] * 3 + te[0][q + 1] + 7) >> 4 : (te[0][q] * 3 + te[0][q - 1] + 8) >> 4;
                crv = (i & 1) ? (te[1][q] * 3 + te[1][q + 1] + 7) >> 4 : (te[1][q] * 3 + te[1][q - 1] + 8) >> 4;
            } else {
                cbv = (i & 1) ? (to[0][q] * 3 + to[0][q + 1] + 7) >> 4 : (to[0][q] * 3 + to[0][q - 1] + 8) >> 4;
                crv = (i & 1) ? (to[1][q] * 3 + to[1][q + 1] + 7) >> 4 : (to[1][q] * 3 + to[1][q - 1] + 8) >> 4;
            }
            const int Y = (int)(((i < 4 ? yw.x : yw.y) >> (8 * (i & 3))) & 255u);
            int rr, gg, bb;
            dfd_jpeg_ycc2rgb(Y, cbv, crv, &rr, &gg, &bb);
            px[3 * i] = (uint8_t)bb; px[3 * i + 1] = (uint8_t)gg; px[3 * i + 2] = (uint8_t)rr;
        }
        uint8_t* o = frames + (size_t)f * frame_stride + (size_t)y * row_pitch + (size_t)x0 * 3;
        if (x0 + 8 <= W && al8) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                uint2 w;
                w.x = px[8 * k] | (px[8 * k + 1] << 8) | (px[8 * k + 2] << 16) | ((uint32_t)px[8 * k + 3] << 24);
                w.y = px[8 * k + 4] | (px[8 * k + 5] << 8) | (px[8 * k + 6] << 16) | ((uint32_t)px[8 * k + 7] << 24);
                ((uint2*)o)[k] = w;
            }
        } else {
            for (int i = 0; i < 8 && x0 + i < W; i++) { o[3 * i] = px[3 * i]; o[3 * i + 1] = px[3 * i + 1]; o[3 * i + 2] = px[3 * i + 2]; }
        }
    }
}

// thread = 4 horizontally adjacent output pixels
__global__ void __launch_bounds__(256)
k_jpeg_color(const DfdJpegHeader* __restrict__ hdr, const uint8_t* __restrict__ planes_all, long long plane_stride,
             uint8_t* __restrict__ frames, size_t frame_stride, int row_pitch, int H, int W) {
    const int f = blockIdx.z;
    const DfdJpegHeader* h = hdr + f;
    const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (y >= H || x0 >= W) return;
    const uint8_t* P = planes_all + (size_t)f * plane_stride;
    const int pitch0 = h->comp_bw[0] * 8;
    uint8_t px[12];
    const uint8_t* p1 = nullptr;
    const uint8_t* p2 = nullptr;
    int pitch1 = 0, cw = 0, chh = 0;
    if (h->ncomp == 3) {
        p1 = P + (size_t)h->comp_bw[0] * h->comp_bh[0] * 64;
        p2 = p1 + (size_t)h->comp_bw[1] * h->comp_bh[1] * 64;
        pitch1 = h->comp_bw[1] * 8;
        cw = (W * h->hs[1] + h->hmax - 1) / h->hmax; chh = (H * h->vs[1] + h->vmax - 1) / h->vmax;
    }
    const bool h2v2 = h->ncomp == 3 && h->hs[1] * 2 == h->hmax && h->vs[1] * 2 == h->vmax && cw > 2;
    if (h2v2) {
        // 4:2:0 fast path (the wire format): the thread's 4 pixels use chroma columns cx, cx + 1 and their neighbours; the
        // vertical part of the triangle filter (3 * near row + far row) is computed once per column and plane -- the same
        // arithmetic as dfd_jpeg_chroma_at (jdsample.c h2v2_fancy_upsample), 16 byte loads instead of 48.
        const int cy = y >> 1, cx = x0 >> 1;
        const int ny = min(max((y & 1) ? cy + 1 : cy - 1, 0), chh - 1);
        const uint32_t yw = *(const uint32_t*)(P + (size_t)y * pitch0 + x0);          // x0 % 4 == 0, pitch % 8 == 0: aligned
        int up[2][4];                                                                 // up-sampled chroma of the 4 pixels
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            const uint8_t* r0 = (pl ? p2 : p1) + (size_t)cy * pitch1;
            const uint8_t* r1 = (pl ? p2 : p1) + (size_t)ny * pitch1;
            int t[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int c = min(max(cx - 1 + q, 0), cw - 1);
                t[q] = 3 * r0[c] + r1[c];
            }
            // pixel x0 (even, column cx), x0+1 (odd, cx), x0+2 (even, cx+1), x0+3 (odd, cx+1)
            up[pl][0] = cx == 0 ? (t[1] * 4 + 8) >> 4 : (t[1] * 3 + t[0] + 8) >> 4;
            up[pl][1] = cx == cw - 1 ? (t[1] * 4 + 7) >> 4 : (t[1] * 3 + t[2] + 7) >> 4;
            up[pl][2] = (t[2] * 3 + t[1] + 8) >> 4;                                   // cx + 1 >= 1: never the left edge
            up[pl][3] = cx + 1 == cw - 1 ? (t[2] * 4 + 7) >> 4 : (t[2] * 3 + t[3] + 7) >> 4;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int r, g, b;
            dfd_jpeg_ycc2rgb((int)((yw >> (8 * i)) & 255u), up[0][i], up[1][i], &r, &g, &b);
            px[3 * i] = (uint8_t)b; px[3 * i + 1] = (uint8_t)g; px[3 * i + 2] = (uint8_t)r;
        }
    } else {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = x0 + i;
        int r = 0, g = 0, b = 0;
        if (x < W) {
            const int Y = P[(size_t)y * pitch0 + x];
            r = g = b = Y;
            if (h->ncomp == 3) {
                const int cb = dfd_jpeg_chroma_at(p1, pitch1, cw, chh, h->hs[1], h->vs[1], h->hmax, h->vmax, x, y);
                const int cr = dfd_jpeg_chroma_at(p2, pitch1, cw, chh, h->hs[2], h->vs[2], h->hmax, h->vmax, x, y);
                dfd_jpeg_ycc2rgb(Y, cb, cr, &r, &g, &b);
            }
        }
        px[3 * i] = (uint8_t)b; px[3 * i + 1] = (uint8_t)g; px[3 * i + 2] = (uint8_t)r;
    }
    }
    uint8_t* o = frames + (size_t)f * frame_stride + (size_t)y * row_pitch + (size_t)x0 * 3;
    if (x0 + 4 <= W && ((uintptr_t)o & 3) == 0) {
        uint32_t w[3];
#pragma unroll
        for (int k = 0; k < 3; k++) w[k] = px[4 * k] | (px[4 * k + 1] << 8) | (px[4 * k + 2] << 16) | ((uint32_t)px[4 * k + 3] << 24);
        ((uint32_t*)o)[0] = w[0]; ((uint32_t*)o)[1] = w[1]; ((uint32_t*)o)[2] = w[2];
    } else {
        for (int i = 0; i < 4 && x0 + i < W; i++) { o[3 * i] = px[3 * i]; o[3 * i + 1] = px[3 * i + 1]; o[3 * i + 2] = px[3 * i + 2]; }
    }
}

// ---------------------------------------------------------------------------------------------
struct JpgHost {
    std::vector<DfdJpegHeader> hdr;
    std::vector<JpgMeta> meta;
    DfdJpegHeader* h_hdr_pinned = nullptr;
    JpgMeta* h_meta_pinned = nullptr;
    size_t pinned_n = 0;
};

int dfd_jpeg_decode_launch(dfd_ctx* ctx, const uint8_t* bytes_host, const int64_t* offsets_host, int n, int H, int W,
                           uint8_t* frames_out, size_t frame_stride, int row_pitch, int32_t* status_dev, cudaStream_t st) {
    DFD_REQUIRE(n > 0 && n <= ctx->cfg.max_batch, DFD_ERR_CAPACITY, "decode_jpeg: batch exceeds max_batch");
    DFD_REQUIRE(H >= 1 && W >= 1 && H <= 16384 && W <= 16384 && row_pitch >= 3 * W, DFD_ERR_INVALID, "decode_jpeg: bad frame geometry");
    if (!ctx->jpg_host) ctx->jpg_host = new JpgHost;
    JpgHost* J = (JpgHost*)ctx->jpg_host;
    if (J->pinned_n < (size_t)n) {
        if (J->h_hdr_pinned) { cudaFreeHost(J->h_hdr_pinned); cudaFreeHost(J->h_meta_pinned); }
        DFD_CUDA(cudaHostAlloc((void**)&J->h_hdr_pinned, sizeof(DfdJpegHeader) * ctx->cfg.max_batch, cudaHostAllocDefault));
        DFD_CUDA(cudaHostAlloc((void**)&J->h_meta_pinned, sizeof(JpgMeta) * ctx->cfg.max_batch, cudaHostAllocDefault));
        J->pinned_n = ctx->cfg.max_batch;
    }
    // the staging arrays are reused by the next call: the previous call's copies must have left them
    if (ctx->jpg_ev) DFD_CUDA(cudaEventSynchronize(ctx->jpg_ev));
    else DFD_CUDA(cudaEventCreateWithFlags(&ctx->jpg_ev, cudaEventDisableTiming));
    // ---- host: headers only ----
    long long words = 0, subs = 0;
    int chunks = 0, max_ecs = 0, max_blocks = 0;
    const long long total_bytes = offsets_host[n] - offsets_host[0];
    for (int i = 0; i < n; i++) {
        const long long b0 = offsets_host[i], b1 = offsets_host[i + 1];
        DFD_REQUIRE(b1 > b0, DFD_ERR_INVALID, "decode_jpeg: empty stream");
        DfdJpegHeader* h = &J->h_hdr_pinned[i];
        const int rc = dfd_jpeg_parse(bytes_host + b0, (size_t)(b1 - b0), h);
        if (rc == DFD_JPEG_ERR_UNSUPPORTED) {
            ctx->err = "decode_jpeg: frame " + std::to_string(i) + " is not a baseline 8-bit 4:4:4 / 4:2:2 / 4:2:0 / gray JPEG without restart markers";
            return DFD_ERR_UNSUPPORTED;
        }
        if (rc != DFD_JPEG_OK) { ctx->err = "decode_jpeg: frame " + std::to_string(i) + " is not a valid JPEG stream"; return DFD_ERR_INVALID; }
        if (h->height != H || h->width != W) {
            ctx->err = "decode_jpeg: frame " + std::to_string(i) + " is " + std::to_string(h->width) + "x" + std::to_string(h->height) + ", the batch is " +
                       std::to_string(W) + "x" + std::to_string(H);
            return DFD_ERR_INVALID;
        }
        JpgMeta& m = J->h_meta_pinned[i];
        m.raw_off = b0 - offsets_host[0];
        m.ecs_bytes = h->ecs_end - h->ecs_begin;
        m.words_off = words;
        m.sub_off = subs;
        m.chunk_off = chunks;
        words += (m.ecs_bytes + 3) / 4 + 6;
        subs += ((long long)m.ecs_bytes * 8 + JPG_SUB_BITS - 1) / JPG_SUB_BITS + 1;
        chunks += (m.ecs_bytes + JU_CHUNK - 1) / JU_CHUNK + 1;
        if (m.ecs_bytes > max_ecs) max_ecs = m.ecs_bytes;
        if (h->total_blocks > max_blocks) max_blocks = h->total_blocks;
    }
    // workspaces: per frame, the largest block count of the batch (4:2:0 needs half of 4:4:4; the coefficient array is zeroed
    // on every call, so its size is time)
    const long long blocks_stride = (max_blocks + 1) & ~1;
    const long long plane_stride = blocks_stride * 64;
    int rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_raw, (size_t)total_bytes + 16))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_words, (size_t)words * 4 + 64))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_sub, (size_t)subs * (8 + 8 + 4 + 4) + (size_t)chunks * 4 + (size_t)n * JH_ROUNDS * 4 + 64))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_coef, (size_t)n * blocks_stride * 64 * sizeof(int16_t)))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_dc, (size_t)n * blocks_stride * sizeof(int32_t)))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_planes, (size_t)n * plane_stride))) return rc;
    if ((rc = dfd_ensure(ctx, ctx->jpg_hdr, (size_t)ctx->cfg.max_batch * (sizeof(DfdJpegHeader) + sizeof(JpgMeta) + 8)))) return rc;
    DfdJpegHeader* d_hdr = (DfdJpegHeader*)ctx->jpg_hdr.p;
    JpgMeta* d_meta = (JpgMeta*)((uint8_t*)ctx->jpg_hdr.p + (size_t)ctx->cfg.max_batch * sizeof(DfdJpegHeader));
    uint32_t* d_nbits = (uint32_t*)((uint8_t*)d_meta + (size_t)ctx->cfg.max_batch * sizeof(JpgMeta));
    unsigned long long* d_E = (unsigned long long*)ctx->jpg_sub.p;
    unsigned long long* d_used = d_E + subs;
    int* d_cnt = (int*)(d_used + subs);
    int* d_blk0 = d_cnt + subs;
    int* d_chunk = d_blk0 + subs;
    int* d_changed = d_chunk + chunks;
    // ---- device ----
    DFD_CUDA(cudaMemcpyAsync(ctx->jpg_raw.p, bytes_host + offsets_host[0], (size_t)total_bytes, cudaMemcpyHostToDevice, st));
    DFD_CUDA(cudaMemcpyAsync(d_hdr, J->h_hdr_pinned, sizeof(DfdJpegHeader) * n, cudaMemcpyHostToDevice, st));
    DFD_CUDA(cudaMemcpyAsync(d_meta, J->h_meta_pinned, sizeof(JpgMeta) * n, cudaMemcpyHostToDevice, st));
    DFD_CUDA(cudaEventRecord(ctx->jpg_ev, st));
    DFD_CUDA(cudaMemsetAsync(ctx->jpg_coef.p, 0, (size_t)n * blocks_stride * 64 * sizeof(int16_t), st));
    DFD_CUDA(cudaMemsetAsync(ctx->jpg_dc.p, 0, (size_t)n * blocks_stride * sizeof(int32_t), st));
    DFD_CUDA(cudaMemsetAsync(d_changed, 0, (size_t)n * JH_ROUNDS * sizeof(int), st));
    const unsigned gc = (unsigned)((max_ecs + JU_CHUNK - 1) / JU_CHUNK);
    k_ju_count<<<dim3(gc, n), 128, 0, st>>>((const uint8_t*)ctx->jpg_raw.p, d_meta, d_hdr, d_chunk);
    DFD_LAUNCH_CHECK("k_ju_count", st);
    k_ju_scan<<<n, JPG_THREADS, 0, st>>>(d_meta, d_chunk, (uint32_t*)ctx->jpg_words.p, d_nbits);
    DFD_LAUNCH_CHECK("k_ju_scan", st);
    k_ju_scatter<<<dim3(gc, n), 128, 0, st>>>((const uint8_t*)ctx->jpg_raw.p, d_meta, d_hdr, d_chunk, (uint32_t*)ctx->jpg_words.p);
    DFD_LAUNCH_CHECK("k_ju_scatter", st);
    static const int rounds_env = getenv("DFD_JH_ROUNDS") ? atoi(getenv("DFD_JH_ROUNDS")) : JH_ROUNDS_DEFAULT;
    const int n_rounds = rounds_env < 1 ? 1 : (rounds_env > JH_ROUNDS ? JH_ROUNDS : rounds_env);
    JhArgs ja;
    ja.last_round = n_rounds - 1;
    ja.meta = d_meta; ja.hdr = d_hdr; ja.words = (const uint32_t*)ctx->jpg_words.p; ja.nbits = d_nbits;
    ja.E = d_E; ja.used = d_used; ja.cnt = d_cnt; ja.blk0 = d_blk0; ja.changed = d_changed;
    ja.coef = (int16_t*)ctx->jpg_coef.p; ja.dc = (int32_t*)ctx->jpg_dc.p; ja.blocks_stride = blocks_stride;
    const unsigned gs = (unsigned)(((long long)max_ecs * 8 + JPG_SUB_BITS - 1) / JPG_SUB_BITS + JH_THREADS - 1) / JH_THREADS;
    k_jh_pass<0><<<dim3(gs, n), JH_THREADS, 0, st>>>(ja, 0);
    DFD_LAUNCH_CHECK("k_jh_blind", st);
    for (int r = 0; r < n_rounds; r++) {
        k_jh_pass<1><<<dim3(gs, n), JH_THREADS, 0, st>>>(ja, r);
        DFD_LAUNCH_CHECK("k_jh_round", st);
    }
    k_jh_finish<<<n, JPG_THREADS, 0, st>>>(ja);
    DFD_LAUNCH_CHECK("k_jh_finish", st);
    k_jh_scan<<<n, JPG_THREADS, 0, st>>>(ja, status_dev);
    DFD_LAUNCH_CHECK("k_jh_scan", st);
    k_jh_pass<2><<<dim3(gs, n), JH_THREADS, 0, st>>>(ja, 0);
    DFD_LAUNCH_CHECK("k_jh_write", st);
    k_jh_dc<<<n, JPG_THREADS, 0, st>>>(ja);
    DFD_LAUNCH_CHECK("k_jh_dc", st);
    k_jpeg_idct<<<dim3((unsigned)((blocks_stride + 127) / 128), n), 128, 0, st>>>(d_hdr, (const int16_t*)ctx->jpg_coef.p, (const int32_t*)ctx->jpg_dc.p,
                                                                                  blocks_stride, (uint8_t*)ctx->jpg_planes.p, plane_stride);
    DFD_LAUNCH_CHECK("k_jpeg_idct", st);
    // every frame of the batch 4:2:0 with more than two chroma columns (what browsers and cv2.imencode emit): the 16-pixel kernel
    bool all420 = true;
    for (int i = 0; i < n; i++) {
        const DfdJpegHeader& h = J->h_hdr_pinned[i];
        all420 = all420 && h.ncomp == 3 && h.hs[0] == 2 && h.vs[0] == 2 && h.hs[1] == 1 && h.vs[1] == 1 && h.hs[2] == 1 && h.vs[2] == 1 && W > 4;
    }
    static const bool no420 = getenv("DFD_JPEG_NO420") != nullptr;
    if (all420 && !no420)
        k_jpeg_color420<<<dim3((W + 511) / 512, (H + 7) / 8, n), 256, 0, st>>>(d_hdr, (const uint8_t*)ctx->jpg_planes.p, plane_stride, frames_out,
                                                                               frame_stride, row_pitch, H, W);
    else
        k_jpeg_color<<<dim3((W + 255) / 256, (H + 3) / 4, n), 256, 0, st>>>(d_hdr, (const uint8_t*)ctx->jpg_planes.p, plane_stride, frames_out,
                                                                            frame_stride, row_pitch, H, W);
    DFD_LAUNCH_CHECK("k_jpeg_color", st);
    return DFD_OK;
}

void dfd_jpeg_free(dfd_ctx* ctx) {
    JpgHost* J = (JpgHost*)ctx->jpg_host;
    if (J) {
        if (J->h_hdr_pinned) { cudaFreeHost(J->h_hdr_pinned); cudaFreeHost(J->h_meta_pinned); }
        delete J;
        ctx->jpg_host = nullptr;
    }
    if (ctx->jpg_ev) { cudaEventDestroy(ctx->jpg_ev); ctx->jpg_ev = nullptr; }
}

// host-only header peek (no context, no device): dims and sampling of a stream, or a negative DFD_JPEG_* code
extern "C" int dfd_jpeg_info(const uint8_t* bytes_host, size_t n, int32_t* info /* H, W, components, luma h, luma v */) {
    if (!bytes_host || !info) return DFD_ERR_INVALID;
    DfdJpegHeader* h = new DfdJpegHeader;
    const int rc = dfd_jpeg_parse(bytes_host, n, h);
    info[0] = h->height; info[1] = h->width; info[2] = h->ncomp; info[3] = h->hs[0]; info[4] = h->vs[0];
    delete h;
    return rc;
}
