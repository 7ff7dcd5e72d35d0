// Baseline JPEG decoding, bit-exact with cv2.imdecode(IMREAD_COLOR) -- the reference's frame ingest
// (backend_server.py:140-142: the browser extension posts canvas.toDataURL('image/jpeg', 0.85) frames,
// extension/content.js:86-109).  OpenCV decodes with libjpeg-turbo defaults: islow integer IDCT, "fancy" (triangle)
// chroma up-sampling, 16-bit fixed-point YCbCr -> RGB.  Those pixel stages are the functions of px_jpeg.h (already
// checked against cv2 for the ELA signal); this header adds what a DECODER needs on top of them:
//
//   * header parsing (host): SOF0 / DQT / DHT / SOS / DRI -> DfdJpegHeader (one per frame)
//   * Huffman decoding as a pure state-transition function on a bit position: dfd_jpeg_step() consumes exactly one
//     symbol (+ its value bits) of the entropy-coded segment and returns the coefficient it produced.  The state is
//     (bit position, block-in-MCU index, zig-zag index); because the function has no other memory, a decoder can be
//     started ANYWHERE in the stream, which is what the GPU kernel's self-synchronising parallel decode relies on.
//   * the geometry helpers shared by the device kernels and the host-side checker (tests/hostcheck).
//
// Everything is __host__ __device__ so that tests/hostcheck can run the very same functions on the CPU -- sequentially
// and as a faithful simulation of the GPU kernel's parallel schedule -- against cv2.imdecode, without a GPU.
#pragma once
#include "px_common.h"
#include "px_jpeg.h"

#define DFD_JPEG_OK 0
#define DFD_JPEG_ERR_FORMAT (-1)          /* not a JPEG / truncated / malformed marker segment */
#define DFD_JPEG_ERR_UNSUPPORTED (-2)     /* progressive, arithmetic, 12-bit, CMYK, restart intervals, exotic sampling */
#define DFD_JPEG_ERR_SIZE (-3)            /* dimensions differ from the batch's H x W */
#define DFD_JPEG_ERR_DATA (-4)            /* entropy-coded data does not decode to the expected number of blocks */

#define DFD_JPEG_MAX_BPM 10               /* blocks per MCU (spec limit) */

// Huffman table in lookup form (jdhuff.h, restated): 8-bit look-ahead for the common short codes, canonical
// max-code / value-offset arrays for the rest.
struct DfdHuffTab {
    uint16_t look[256];                   // (length << 8) | symbol for codes of <= 8 bits, 0 = longer code
    int32_t maxcode[18];                  // largest code of length l (-1 if none); maxcode[17] = sentinel
    int32_t valoff[17];                   // huffval index of the first code of length l, minus that code
    uint8_t huffval[256];
};

struct DfdJpegHeader {
    int32_t width, height, ncomp;
    int32_t hs[3], vs[3];                 // sampling factors
    int32_t hmax, vmax;
    int32_t mcus_x, mcus_y, bpm;          // MCU grid, blocks per MCU
    int32_t blk_comp[DFD_JPEG_MAX_BPM];   // component of the j-th block of an MCU
    int32_t blk_first[3];                 // index of the component's first block inside the MCU
    int32_t comp_bw[3], comp_bh[3];       // component size in blocks (MCU-padded)
    int32_t comp_blk0[3];                 // first block of the component in the frame's coefficient array
    int32_t total_blocks;
    int32_t ecs_begin, ecs_end;           // entropy-coded segment [begin, end) inside the stream (bytes)
    int32_t restart_interval;
    int32_t dc_tab[3], ac_tab[3];         // table selectors per component
    int32_t status;
    uint16_t qt[3][64];                   // per COMPONENT, natural (row-major) order
    DfdHuffTab dc[2], ac[2];
};

static const unsigned char DFD_ZIGZAG[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                             41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                             15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
                                             62, 63};
#if defined(__CUDACC__)
__constant__ unsigned char DFD_ZIGZAG_DEV[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                                 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                                 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
                                                 62, 63};
#endif
DFD_HD int dfd_zigzag(int k) {
#if defined(__CUDA_ARCH__)
    return DFD_ZIGZAG_DEV[k];
#else
    return DFD_ZIGZAG[k];
#endif
}

// ---- bit access ------------------------------------------------------------------------------------------------------
// The entropy-coded segment with its byte stuffing removed (FF 00 -> FF), stored as 32-bit words holding the bytes
// MSB-first (word w = b[4w] << 24 | b[4w+1] << 16 | ...), so that bit position p is bit (31 - p % 32) of word p / 32.
// Returns the 16 bits starting at p (zeros beyond the end: the caller bounds p).
DFD_HD uint32_t dfd_peek16(const uint32_t* words, uint32_t nwords, uint32_t p) {
    const uint32_t w = p >> 5, s = p & 31u;
    const uint32_t w0 = w < nwords ? words[w] : 0u, w1 = w + 1 < nwords ? words[w + 1] : 0u;
    const uint64_t both = ((uint64_t)w0 << 32) | w1;
    return (uint32_t)(both >> (48 - s)) & 0xffffu;
}

// One Huffman symbol at bit position p: returns the symbol, *len = code length (0 = invalid code).
DFD_HD int dfd_huff_decode(const DfdHuffTab* t, uint32_t bits16, int* len) {
    const uint32_t e = t->look[bits16 >> 8];
    if (e) { *len = (int)(e >> 8); return (int)(e & 255u); }
    int l = 9;
    int32_t code = (int32_t)(bits16 >> 7);
    while (l <= 16 && code > t->maxcode[l]) { l++; code = (int32_t)(bits16 >> (16 - l)); }
    if (l > 16) { *len = 0; return 0; }
    *len = l;
    return t->huffval[(code + t->valoff[l]) & 255];
}

// Decoder state between two symbols.
struct DfdJpegState {
    uint32_t p;          // bit position of the next symbol
    int32_t c;           // block-in-MCU index of the block being decoded
    int32_t z;           // zig-zag index of the next coefficient (0 = the block's DC symbol comes next)
};

// Consumes one symbol (+ value bits).  Outputs: *k = zig-zag index of the coefficient produced (-1 = none: EOB / ZRL),
// *val = its value (the DC value is the DIFFERENCE to the previous block of the component), *done = 1 when the symbol
// completed a block.  Returns 0, or -1 on an invalid code (the state still advances so a blind decoder cannot stall).
DFD_HD int dfd_jpeg_step(const DfdJpegHeader* h, const uint32_t* words, uint32_t nwords, DfdJpegState* s, int* k, int* val,
                         int* done) {
    const int comp = h->blk_comp[s->c];
    int len, rc = 0;
    *k = -1; *val = 0; *done = 0;
    uint32_t b = dfd_peek16(words, nwords, s->p);
    if (s->z == 0) {
        const int sym = dfd_huff_decode(&h->dc[h->dc_tab[comp]], b, &len);
        if (len == 0) { len = 1; rc = -1; }
        const int size = sym & 15;
        s->p += (uint32_t)len;
        if (size) {
            const uint32_t v = dfd_peek16(words, nwords, s->p) >> (16 - size);
            s->p += (uint32_t)size;
            *val = (int)v < (1 << (size - 1)) ? (int)v - (1 << size) + 1 : (int)v;      // EXTEND
        }
        *k = 0;
        s->z = 1;
    } else {
        const int sym = dfd_huff_decode(&h->ac[h->ac_tab[comp]], b, &len);
        if (len == 0) { len = 1; rc = -1; }
        const int run = sym >> 4, size = sym & 15;
        s->p += (uint32_t)len;
        if (size == 0) {
            if (run == 15) s->z += 16;           // ZRL
            else s->z = 64;                      // EOB
        } else {
            s->z += run;
            const uint32_t v = dfd_peek16(words, nwords, s->p) >> (16 - size);
            s->p += (uint32_t)size;
            if (s->z < 64) {
                *k = s->z;
                *val = (int)v < (1 << (size - 1)) ? (int)v - (1 << size) + 1 : (int)v;
            }
            s->z += 1;
        }
    }
    if (s->z >= 64) {
        s->z = 0;
        s->c = s->c + 1 == h->bpm ? 0 : s->c + 1;
        *done = 1;
    }
    return rc;
}

// Position of block number `blk` (decode order) of the frame: component, block coordinates, index into the coefficient
// array (blocks of a component are stored row-major over the component's MCU-padded block grid) and the block's rank
// in its component's DC-prediction chain (= decode order within the component).
DFD_HD void dfd_jpeg_block_pos(const DfdJpegHeader* h, int blk, int* comp, int* index, int* dc_seq) {
    const int mcu = blk / h->bpm, j = blk - mcu * h->bpm;
    const int c = h->blk_comp[j], jj = j - h->blk_first[c];
    const int my = mcu / h->mcus_x, mx = mcu - my * h->mcus_x;
    const int bx = mx * h->hs[c] + jj % h->hs[c], by = my * h->vs[c] + jj / h->hs[c];
    *comp = c;
    *index = h->comp_blk0[c] + by * h->comp_bw[c] + bx;
    *dc_seq = mcu * (h->hs[c] * h->vs[c]) + jj;
}
// inverse direction for the IDCT stage: DC chain rank of the component's block (bx, by)
DFD_HD int dfd_jpeg_dc_seq(const DfdJpegHeader* h, int c, int bx, int by) {
    const int mx = bx / h->hs[c], my = by / h->vs[c];
    return (my * h->mcus_x + mx) * (h->hs[c] * h->vs[c]) + (by - my * h->vs[c]) * h->hs[c] + (bx - mx * h->hs[c]);
}

// Dequantise + islow IDCT (jidctint.c: columns, then rows, +128, range-limit) of one block; coef in natural order.
DFD_HD void dfd_jpeg_idct_block(const int16_t* coef, int dc, const uint16_t* qt, uint8_t* out /* 64, row-major */) {
    int blk[64];
    for (int i = 0; i < 64; i++) blk[i] = (int)coef[i] * (int)qt[i];
    blk[0] = dc * (int)qt[0];
    for (int c = 0; c < 8; c++) dfd_idct8(blk + c, 8, 1);
    for (int r = 0; r < 8; r++) dfd_idct8(blk + 8 * r, 1, 0);
    for (int i = 0; i < 64; i++) out[i] = (uint8_t)dfd_sat_u8(blk[i] + 128);
}

// Chroma sample at full-resolution (X, Y) for sampling (hs, vs) of the chroma component relative to luma (hmax, vmax):
// h2v2 and h2v1 "fancy" triangle filters, h1v1 copy, anything else replication -- jdsample.c with
// do_fancy_upsampling = TRUE.  plane: stride `pitch`; cw x ch = the component's DOWNSAMPLED size (ceil(W*hs/hmax), ...):
// the edge cases apply at the last REAL column / row, not at the MCU padding.
DFD_HD int dfd_jpeg_chroma_at(const uint8_t* plane, int pitch, int cw, int ch, int hs, int vs, int hmax, int vmax, int X, int Y) {
    if (hs == hmax && vs == vmax) return plane[Y * pitch + X];
    // jdsample.c (jinit_upsampler): the triangle filters are used only when the component is more than 2 samples wide;
    // narrower components (image width <= 4) are up-sampled by plain replication
    if (cw <= 2) return plane[(Y * vs / vmax) * pitch + (X * hs / hmax)];
    if (hs * 2 == hmax && vs * 2 == vmax) {            // h2v2
        const int cy = Y >> 1, cx = X >> 1;
        int ny = (Y & 1) ? cy + 1 : cy - 1;
        ny = dfd_clampi(ny, 0, ch - 1);
        const uint8_t* r0 = plane + cy * pitch;
        const uint8_t* r1 = plane + ny * pitch;
        const int cur = 3 * r0[cx] + r1[cx];
        if (X & 1) {
            if (cx == cw - 1) return (cur * 4 + 7) >> 4;
            return (cur * 3 + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
        }
        if (cx == 0) return (cur * 4 + 8) >> 4;
        return (cur * 3 + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
    }
    if (hs * 2 == hmax && vs == vmax) {                // h2v1
        const int cx = X >> 1;
        const uint8_t* r0 = plane + Y * pitch;
        const int cur = r0[cx];
        if (X & 1) {
            if (cx == cw - 1) return cur;
            return (cur * 3 + r0[cx + 1] + 2) >> 2;
        }
        if (cx == 0) return cur;
        return (cur * 3 + r0[cx - 1] + 1) >> 2;
    }
    // integral replication (h1v2 is handled by the caller's support check)
    return plane[(Y * vs / vmax) * pitch + (X * hs / hmax)];
}

// ---- host: header parsing ---------------------------------------------------------------------------------------------
static inline void dfd_jpeg_build_huff(const uint8_t* bits /* [17], bits[0] unused */, const uint8_t* vals, int nvals, DfdHuffTab* t) {
    for (int i = 0; i < 256; i++) { t->look[i] = 0; t->huffval[i] = i < nvals ? vals[i] : 0; }
    int code = 0, p = 0;
    for (int l = 1; l <= 16; l++) {
        if (bits[l]) {
            t->valoff[l] = p - code;
            for (int i = 0; i < bits[l]; i++, p++, code++) {
                if (l <= 8) {
                    const int lo = code << (8 - l), n = 1 << (8 - l);
                    for (int q = 0; q < n && lo + q < 256; q++) t->look[lo + q] = (uint16_t)((l << 8) | (p < nvals ? vals[p] : 0));
                }
            }
            t->maxcode[l] = code - 1;
        } else { t->maxcode[l] = -1; t->valoff[l] = 0; }
        code <<= 1;
    }
    t->maxcode[0] = -1;
    t->maxcode[17] = 0x7fffffff;
}

// Parses the markers of one JPEG stream.  Supported: baseline sequential DCT (SOF0; SOF1 with 8-bit samples), 8-bit,
// 1 or 3 components, luma sampling 1x1 / 2x1 / 2x2 with 1x1 chroma (4:4:4, 4:2:2, 4:2:0), one interleaved scan, no
// restart interval -- what browsers' canvas encoders and cv2.imencode emit.  Everything else -> DFD_JPEG_ERR_UNSUPPORTED.
static inline int dfd_jpeg_parse(const uint8_t* d, size_t n, DfdJpegHeader* h) {
    for (size_t i = 0; i < sizeof(DfdJpegHeader); i++) ((uint8_t*)h)[i] = 0;
    h->status = DFD_JPEG_ERR_FORMAT;
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return h->status;
    uint16_t qt_raw[4][64];
    bool have_qt[4] = {false, false, false, false}, have_dc[4] = {false, false, false, false}, have_ac[4] = {false, false, false, false};
    int tq[3] = {0, 0, 0}, comp_id[3] = {0, 0, 0};
    DfdHuffTab dc[4], ac[4];
    bool have_sof = false;
    size_t pos = 2;
    while (pos + 4 <= n) {
        if (d[pos] != 0xFF) return h->status;
        while (pos < n && d[pos] == 0xFF) pos++;                       // fill bytes
        if (pos >= n) return h->status;
        const int m = d[pos++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return h->status;                               // EOI before SOS
        if (pos + 2 > n) return h->status;
        const size_t len = ((size_t)d[pos] << 8) | d[pos + 1];
        if (len < 2 || pos + len > n) return h->status;
        const uint8_t* s = d + pos + 2;
        const size_t sl = len - 2;
        if (m == 0xDB) {                                               // DQT
            size_t q = 0;
            while (q < sl) {
                const int pq = s[q] >> 4, id = s[q] & 15;
                q++;
                if (id > 3) return h->status;
                if (pq != 0) return h->status = DFD_JPEG_ERR_UNSUPPORTED;        // 16-bit tables: not baseline
                if (q + 64 > sl) return h->status;
                for (int k = 0; k < 64; k++) qt_raw[id][DFD_ZIGZAG[k]] = s[q + k];
                have_qt[id] = true;
                q += 64;
            }
        } else if (m == 0xC4) {                                        // DHT
            size_t q = 0;
            while (q < sl) {
                if (q + 17 > sl) return h->status;
                const int tc = s[q] >> 4, id = s[q] & 15;
                if (tc > 1 || id > 3) return h->status;
                uint8_t bits[17];
                bits[0] = 0;
                int cnt = 0;
                for (int l = 1; l <= 16; l++) { bits[l] = s[q + l]; cnt += bits[l]; }
                q += 17;
                if (cnt > 256 || q + cnt > sl) return h->status;
                dfd_jpeg_build_huff(bits, s + q, cnt, tc ? &ac[id] : &dc[id]);
                (tc ? have_ac : have_dc)[id] = true;
                q += cnt;
            }
        } else if (m == 0xC0 || m == 0xC1) {                           // SOF0 / SOF1 (Huffman, sequential)
            if (sl < 6) return h->status;
            if (s[0] != 8) return h->status = DFD_JPEG_ERR_UNSUPPORTED;
            h->height = (s[1] << 8) | s[2]; h->width = (s[3] << 8) | s[4]; h->ncomp = s[5];
            if (h->ncomp != 1 && h->ncomp != 3) return h->status = DFD_JPEG_ERR_UNSUPPORTED;
            if (sl < 6 + 3 * (size_t)h->ncomp || h->height == 0 || h->width == 0) return h->status;
            for (int c = 0; c < h->ncomp; c++) {
                comp_id[c] = s[6 + 3 * c];
                h->hs[c] = s[7 + 3 * c] >> 4; h->vs[c] = s[7 + 3 * c] & 15;
                tq[c] = s[8 + 3 * c];
                if (tq[c] > 3 || h->hs[c] < 1 || h->vs[c] < 1) return h->status;
            }
            have_sof = true;
        } else if ((m >= 0xC2 && m <= 0xCF) && m != 0xC4 && m != 0xC8) {
            return h->status = DFD_JPEG_ERR_UNSUPPORTED;               // progressive / lossless / arithmetic
        } else if (m == 0xDD) {                                        // DRI
            if (sl < 2) return h->status;
            h->restart_interval = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {                                        // SOS
            if (!have_sof || sl < 1) return h->status;
            const int ns = s[0];
            if (ns != h->ncomp) return h->status = DFD_JPEG_ERR_UNSUPPORTED;     // non-interleaved scans
            if (sl < 1 + 2 * (size_t)ns + 3) return h->status;
            for (int c = 0; c < ns; c++) {
                if (s[1 + 2 * c] != comp_id[c]) return h->status = DFD_JPEG_ERR_UNSUPPORTED;
                h->dc_tab[c] = s[2 + 2 * c] >> 4; h->ac_tab[c] = s[2 + 2 * c] & 15;
                if (h->dc_tab[c] > 1 || h->ac_tab[c] > 1) return h->status = DFD_JPEG_ERR_UNSUPPORTED;   // baseline: tables 0, 1
                if (!have_dc[h->dc_tab[c]] || !have_ac[h->ac_tab[c]] || !have_qt[tq[c]]) return h->status;
            }
            if (h->restart_interval != 0) return h->status = DFD_JPEG_ERR_UNSUPPORTED;
            // geometry
            if (h->ncomp == 1) { h->hs[0] = h->vs[0] = 1; }            // a single-component scan is never interleaved: 1 block per MCU
            else {
                const bool chroma_ok = h->hs[1] == 1 && h->vs[1] == 1 && h->hs[2] == 1 && h->vs[2] == 1;
                const bool luma_ok = (h->hs[0] == 1 || h->hs[0] == 2) && (h->vs[0] == 1 || h->vs[0] == 2) && !(h->hs[0] == 1 && h->vs[0] == 2);
                if (!chroma_ok || !luma_ok) return h->status = DFD_JPEG_ERR_UNSUPPORTED;
            }
            h->hmax = h->hs[0]; h->vmax = h->vs[0];
            h->mcus_x = (h->width + 8 * h->hmax - 1) / (8 * h->hmax);
            h->mcus_y = (h->height + 8 * h->vmax - 1) / (8 * h->vmax);
            h->bpm = 0;
            int blk0 = 0;
            for (int c = 0; c < h->ncomp; c++) {
                h->blk_first[c] = h->bpm;
                for (int j = 0; j < h->hs[c] * h->vs[c]; j++) h->blk_comp[h->bpm++] = c;
                h->comp_bw[c] = h->mcus_x * h->hs[c]; h->comp_bh[c] = h->mcus_y * h->vs[c];
                h->comp_blk0[c] = blk0;
                blk0 += h->comp_bw[c] * h->comp_bh[c];
                for (int k = 0; k < 64; k++) h->qt[c][k] = qt_raw[tq[c]][k];
            }
            h->total_blocks = blk0;
            h->dc[0] = dc[0]; h->dc[1] = dc[1]; h->ac[0] = ac[0]; h->ac[1] = ac[1];
            h->ecs_begin = (int32_t)(pos + len);
            // the scan ends at the next marker (EOI); without restart markers that is the first FF followed by a non-zero byte
            size_t e = (size_t)h->ecs_begin;
            if (n >= 2 && d[n - 2] == 0xFF && d[n - 1] == 0xD9) {
                // common case: EOI closes the buffer; verify no other marker hides in the last bytes only (full scan not needed:
                // a stray marker inside the data shows up as a block-count mismatch on the device)
                e = n - 2;
            } else {
                while (e + 1 < n && !(d[e] == 0xFF && d[e + 1] != 0x00)) e++;
                if (e + 1 >= n) e = n;
            }
            h->ecs_end = (int32_t)e;
            h->status = DFD_JPEG_OK;
            return h->status;
        }
        pos += len;
    }
    return h->status;
}

// ---- one subsequence of the entropy-coded segment ------------------------------------------------------------------------
// Decodes every symbol that STARTS in [s.p, limit) from state s (a symbol belongs to the subsequence it starts in).
// Returns the end state packed into 64 bits (so that it can be published with one store), the number of blocks completed
// in *nblk and an error count in *nerr.  WRITE: coefficients are stored -- the first block touched has decode-order number
// blk0 -- AC values into coef[index * 64 + natural position], DC differences into dcdiff[dc_off[comp] + chain rank].
DFD_HD uint64_t dfd_jpeg_pack_state(const DfdJpegState& s) { return ((uint64_t)s.p << 16) | ((uint64_t)(uint32_t)s.c << 8) | (uint64_t)(uint32_t)s.z; }
DFD_HD DfdJpegState dfd_jpeg_unpack_state(uint64_t v) {
    DfdJpegState s;
    s.p = (uint32_t)(v >> 16); s.c = (int32_t)((v >> 8) & 255u); s.z = (int32_t)(v & 255u);
    return s;
}

template <bool WRITE>
DFD_HD uint64_t dfd_jpeg_decode_sub(const DfdJpegHeader* h, const uint32_t* words, uint32_t nwords, uint64_t start, uint32_t limit,
                                    int* nblk, int* nerr, int blk0, int16_t* coef, int32_t* dcdiff, const int32_t* dc_off) {
    DfdJpegState s = dfd_jpeg_unpack_state(start);
    int blocks = 0, errs = 0;
    int blk = blk0, comp = 0, index = 0, dc_seq = 0;
    if (WRITE && blk < h->total_blocks) dfd_jpeg_block_pos(h, blk, &comp, &index, &dc_seq);
    while (s.p < limit) {
        int k, val, done;
        if (dfd_jpeg_step(h, words, nwords, &s, &k, &val, &done)) errs++;
        if (WRITE && k >= 0 && blk < h->total_blocks) {
            if (k == 0) dcdiff[dc_off[comp] + dc_seq] = val;
            else coef[(size_t)index * 64 + dfd_zigzag(k)] = (int16_t)val;
        }
        if (done) {
            blocks++;
            if (WRITE) {
                blk++;
                if (blk < h->total_blocks) dfd_jpeg_block_pos(h, blk, &comp, &index, &dc_seq);
            }
        }
    }
    *nblk = blocks; *nerr = errs;
    return dfd_jpeg_pack_state(s);
}
