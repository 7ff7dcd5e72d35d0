// EfficientNet-B0 layer plan + packed-parameter layout shared by the weight
// loader and the launch code.  Mirrors dfd_b200/arch.py and weights.py (the
// Python packer); dfd_weights_blob_floats() lets the host cross-check.
//
// Blob (float32, every tensor offset aligned to 64 elements), BatchNorm folded
// (reference model.py:36-61):
//   stem   W[27][32]  (k = (ky*3+kx)*3+cin, cout fastest)      b[32]
//   block  [We[cexp][cin] be[cexp]]   (absent for block 0)
//          Wd[k*k][cexp] bd[cexp]
//          Wr[se][cexp] br[se]  Wx[cexp][se] bx[cexp]
//          Wp[cout][cexp] bp[cout]
//   head   Wh[1280][320] bh[1280]
//   fc     W1[512][1280] b1[512]  W2[256][512] b2[256]  W3[1][256] b3[1]
#pragma once
#include <stddef.h>

struct EffBlock { int k, s, cin, cexp, cout, se, hin, hout, pad; };

static const EffBlock EFF_BLOCKS[16] = {
    {3, 1, 32, 32, 16, 8, 112, 112, 1},    {3, 2, 16, 96, 24, 4, 112, 56, 0},     {3, 1, 24, 144, 24, 6, 56, 56, 1},
    {5, 2, 24, 144, 40, 6, 56, 28, 1},     {5, 1, 40, 240, 40, 10, 28, 28, 2},    {3, 2, 40, 240, 80, 10, 28, 14, 0},
    {3, 1, 80, 480, 80, 20, 14, 14, 1},    {3, 1, 80, 480, 80, 20, 14, 14, 1},    {5, 1, 80, 480, 112, 20, 14, 14, 2},
    {5, 1, 112, 672, 112, 28, 14, 14, 2},  {5, 1, 112, 672, 112, 28, 14, 14, 2},  {5, 2, 112, 672, 192, 28, 14, 7, 1},
    {5, 1, 192, 1152, 192, 48, 7, 7, 2},   {5, 1, 192, 1152, 192, 48, 7, 7, 2},   {5, 1, 192, 1152, 192, 48, 7, 7, 2},
    {3, 1, 192, 1152, 320, 48, 7, 7, 1},
};

struct EffBlockOff { size_t we, be, wd, bd, wr, br, wx, bx, wp, bp; };
struct EffOffsets {
    size_t stem_w, stem_b;
    EffBlockOff blk[16];
    size_t head_w, head_b, fc1_w, fc1_b, fc2_w, fc2_b, fc3_w, fc3_b;
    size_t total;
};

static inline size_t eff_align(size_t v) { return (v + 63) & ~(size_t)63; }

static inline EffOffsets eff_offsets() {
    EffOffsets o;
    size_t p = 0;
    auto take = [&](size_t n) { size_t at = p; p = eff_align(p + n); return at; };
    o.stem_w = take(27 * 32); o.stem_b = take(32);
    for (int i = 0; i < 16; i++) {
        const EffBlock& b = EFF_BLOCKS[i];
        EffBlockOff& f = o.blk[i];
        if (b.cexp != b.cin) { f.we = take((size_t)b.cexp * b.cin); f.be = take(b.cexp); } else { f.we = f.be = 0; }
        f.wd = take((size_t)b.k * b.k * b.cexp); f.bd = take(b.cexp);
        f.wr = take((size_t)b.se * b.cexp); f.br = take(b.se);
        f.wx = take((size_t)b.cexp * b.se); f.bx = take(b.cexp);
        f.wp = take((size_t)b.cout * b.cexp); f.bp = take(b.cout);
    }
    o.head_w = take(1280 * 320); o.head_b = take(1280);
    o.fc1_w = take(512 * 1280); o.fc1_b = take(512);
    o.fc2_w = take(256 * 512); o.fc2_b = take(256);
    o.fc3_w = take(256); o.fc3_b = take(1);
    o.total = p;
    return o;
}
