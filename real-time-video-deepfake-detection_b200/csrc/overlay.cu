// Result annotation: the drawing the reference does on every frame it returns from predict()
// (deepfake_detection.py:559-586 draw_detection_overlay, :688-726 _draw_frame_analysis_overlay), composited on the DEVICE copy
// of the frame, bit-exact with OpenCV's rasterisers:
//   DFD_DRAW_OUTLINE  cv2.rectangle(img, p0, p1, color, thickness t >= 1): the union of four thick axis-aligned lines; OpenCV
//                     draws a line of thickness t as a band of half-width r = t - 1 with diamond caps of radius r
//                     (pixel (px, py) belongs to the horizontal line x0..x1 at y iff d = |py - y| <= r and
//                     x0 - (r - d) <= px <= x1 + (r - d); likewise for vertical lines) -- checked against cv2 on the CPU
//   DFD_DRAW_FILL     cv2.rectangle(.., thickness = -1): both corners inclusive
//   DFD_DRAW_BLEND    overlay = img.copy(); fill rectangle on overlay; cv2.addWeighted(overlay, alpha, img, 1 - alpha, 0, img):
//                     saturate(round_half_even(float(color) * alpha + float(px) * beta)) in float32 inside the rectangle, identity
//                     outside
//   DFD_DRAW_MASK     cv2.putText: the string's stroke mask (rasterised by the caller with OpenCV's Hershey font -- the glyph
//                     outlines are font data, not arithmetic of this path) is stamped with the text colour
// Commands apply in order, like successive OpenCV calls.  One thread per pixel of the rows any command touches.
#include "dfd_internal.cuh"

struct DrawCmd { int32_t op, x0, y0, x1, y1, thickness; uint8_t color[4]; float alpha, beta; int32_t mask_off, mask_w, mask_h, pad; };
static_assert(sizeof(DrawCmd) == sizeof(dfd_draw_cmd), "dfd_draw_cmd layout");

__device__ __forceinline__ bool on_thick_hline(int px, int py, int xa, int xb, int y, int r) {
    const int d = py > y ? py - y : y - py;
    return d <= r && px >= xa - (r - d) && px <= xb + (r - d);
}

__global__ void k_overlay(uint8_t* __restrict__ frame, int H, int W, int pitch, const DrawCmd* __restrict__ cmds, int n_cmds,
                          const uint8_t* __restrict__ masks, int row0, int rows) {
    extern __shared__ DrawCmd s_cmd[];
    for (int i = threadIdx.x; i < n_cmds; i += blockDim.x) s_cmd[i] = cmds[i];
    __syncthreads();
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = row0 + blockIdx.y;
    if (px >= W || py >= H || blockIdx.y >= rows) return;
    uint8_t* p = frame + (size_t)py * pitch + (size_t)px * 3;
    int c0 = p[0], c1 = p[1], c2 = p[2];
    bool touched = false;
    for (int i = 0; i < n_cmds; i++) {
        const DrawCmd& c = s_cmd[i];
        bool hit = false;
        if (c.op == DFD_DRAW_MASK) {
            const int mx = px - c.x0, my = py - c.y0;
            hit = mx >= 0 && my >= 0 && mx < c.mask_w && my < c.mask_h && masks[c.mask_off + my * c.mask_w + mx] != 0;
        } else {
            const int xa = min(c.x0, c.x1), xb = max(c.x0, c.x1), ya = min(c.y0, c.y1), yb = max(c.y0, c.y1);
            if (c.op == DFD_DRAW_OUTLINE) {
                const int r = c.thickness - 1;
                hit = on_thick_hline(px, py, xa, xb, ya, r) || on_thick_hline(px, py, xa, xb, yb, r) ||
                      on_thick_hline(py, px, ya, yb, xa, r) || on_thick_hline(py, px, ya, yb, xb, r);   // vertical edges: axes swapped
            } else {
                hit = px >= xa && px <= xb && py >= ya && py <= yb;
            }
        }
        if (!hit) continue;
        touched = true;
        if (c.op == DFD_DRAW_BLEND) {
            c0 = min(max(__float2int_rn(__fadd_rn(__fmul_rn((float)c.color[0], c.alpha), __fmul_rn((float)c0, c.beta))), 0), 255);
            c1 = min(max(__float2int_rn(__fadd_rn(__fmul_rn((float)c.color[1], c.alpha), __fmul_rn((float)c1, c.beta))), 0), 255);
            c2 = min(max(__float2int_rn(__fadd_rn(__fmul_rn((float)c.color[2], c.alpha), __fmul_rn((float)c2, c.beta))), 0), 255);
        } else {
            c0 = c.color[0]; c1 = c.color[1]; c2 = c.color[2];
        }
    }
    if (touched) { p[0] = (uint8_t)c0; p[1] = (uint8_t)c1; p[2] = (uint8_t)c2; }
}

int dfd_overlay_launch(dfd_ctx* ctx, uint8_t* frame, int H, int W, int row_pitch, const dfd_draw_cmd* cmds_host, int n_cmds,
                       const uint8_t* masks_host, size_t mask_bytes, cudaStream_t st) {
    DFD_REQUIRE(H >= 1 && W >= 1 && row_pitch >= 3 * W, DFD_ERR_INVALID, "draw_overlay: bad frame geometry");
    DFD_REQUIRE(n_cmds >= 0 && n_cmds <= DFD_DRAW_MAX_CMDS, DFD_ERR_CAPACITY, "draw_overlay: more than DFD_DRAW_MAX_CMDS commands");
    if (n_cmds == 0) return DFD_OK;
    int row0 = H, row1 = -1;
    for (int i = 0; i < n_cmds; i++) {
        const dfd_draw_cmd& c = cmds_host[i];
        DFD_REQUIRE(c.op >= DFD_DRAW_OUTLINE && c.op <= DFD_DRAW_MASK, DFD_ERR_INVALID, "draw_overlay: unknown op");
        int ya, yb;
        if (c.op == DFD_DRAW_MASK) {
            DFD_REQUIRE(c.mask_w >= 0 && c.mask_h >= 0 && c.mask_off >= 0 &&
                        (size_t)c.mask_off + (size_t)c.mask_w * c.mask_h <= mask_bytes, DFD_ERR_INVALID, "draw_overlay: mask outside the mask buffer");
            ya = c.y0; yb = c.y0 + c.mask_h - 1;
        } else {
            DFD_REQUIRE(c.op != DFD_DRAW_OUTLINE || (c.thickness >= 1 && c.thickness <= 16), DFD_ERR_INVALID, "draw_overlay: thickness outside 1..16");
            const int r = c.op == DFD_DRAW_OUTLINE ? c.thickness - 1 : 0;
            ya = (c.y0 < c.y1 ? c.y0 : c.y1) - r; yb = (c.y0 < c.y1 ? c.y1 : c.y0) + r;
        }
        if (ya < row0) row0 = ya;
        if (yb > row1) row1 = yb;
    }
    if (row0 < 0) row0 = 0;
    if (row1 > H - 1) row1 = H - 1;
    if (row1 < row0) return DFD_OK;
    const size_t cmd_bytes = (size_t)n_cmds * sizeof(dfd_draw_cmd);
    int rc = dfd_ensure(ctx, ctx->draw_buf, cmd_bytes + mask_bytes + 16);
    if (rc) return rc;
    // the command list and the masks are small (a few KB): staged through the context so the caller's arrays are free on return
    if (ctx->draw_host_bytes < cmd_bytes + mask_bytes) {
        if (ctx->draw_host) cudaFreeHost(ctx->draw_host);
        ctx->draw_host = nullptr; ctx->draw_host_bytes = 0;
        DFD_CUDA(cudaMallocHost(&ctx->draw_host, (cmd_bytes + mask_bytes) * 2 + 4096));
        ctx->draw_host_bytes = (cmd_bytes + mask_bytes) * 2 + 4096;
    }
    DFD_CUDA(cudaStreamSynchronize(st));                 // the previous call's copy out of the staging buffer
    memcpy(ctx->draw_host, cmds_host, cmd_bytes);
    if (mask_bytes) memcpy((uint8_t*)ctx->draw_host + cmd_bytes, masks_host, mask_bytes);
    DFD_CUDA(cudaMemcpyAsync(ctx->draw_buf.p, ctx->draw_host, cmd_bytes + mask_bytes, cudaMemcpyHostToDevice, st));
    const int rows = row1 - row0 + 1;
    k_overlay<<<dim3((W + 255) / 256, rows), 256, cmd_bytes, st>>>(frame, H, W, row_pitch, (const DrawCmd*)ctx->draw_buf.p, n_cmds,
                                                                  (const uint8_t*)ctx->draw_buf.p + cmd_bytes, row0, rows);
    DFD_LAUNCH_CHECK("k_overlay", st);
    return DFD_OK;
}
