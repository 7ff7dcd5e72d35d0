"""Host-side engine: owns one libdfd context on one GPU and moves torch tensors
across the C-ABI.  PyTorch is used for device memory, streams and (in
sharding.py) torch.distributed only -- every kernel on the path is libdfd's.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, weights as _weights

FORENSIC_DTYPE = np.dtype([("raw", "<f8", (_lib.N_RAW,)), ("scores", "<f8", (_lib.N_SIGNALS,)),
                           ("fake_probability", "<f8"), ("frame_number", "<i4"), ("full", "<i4")])
RECORD_DTYPE = np.dtype([("stream_id", "<i4"), ("verdict", "<i4"), ("fake_count", "<i4"), ("real_count", "<i4"),
                         ("history_len", "<i4"), ("frame_count", "<i4"), ("last_vote", "<i4"), ("reserved", "<i4"),
                         ("vote_input", "<f8"),
                         ("temporal_average", "<f8"), ("stability_score", "<f8"), ("face_probability", "<f8"),
                         ("forensic_probability", "<f8")])
assert FORENSIC_DTYPE.itemsize == _lib.FORENSIC_BYTES and RECORD_DTYPE.itemsize == _lib.RECORD_BYTES

DTYPES = {"fp32": (_lib.F32, torch.float32), "bf16": (_lib.BF16, torch.bfloat16)}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    """One context per GPU.  All tensor arguments are CUDA tensors on ``device``."""

    def __init__(self, device=0, max_streams=256, max_batch=256, max_crop=1024, window_size=60, voting_window=10,
                 detection_threshold=0.5, face_weight=0.70, forensic_weight=0.30, blend_mode="reference"):
        if not torch.cuda.is_available():
            raise _lib.DfdError("no CUDA device visible: the B200 path has no CPU fallback")
        self.lib = _lib.load()
        if isinstance(device, torch.device):
            device = device.index or 0
        self.device = torch.device("cuda", int(device))
        cfg = _lib.Config()
        self.lib.dfd_default_config(C.byref(cfg))
        cfg.device, cfg.max_streams, cfg.max_batch, cfg.max_crop = int(device), max_streams, max_batch, max_crop
        cfg.window_size, cfg.voting_window = window_size, voting_window
        cfg.detection_threshold, cfg.face_weight, cfg.forensic_weight = detection_threshold, face_weight, forensic_weight
        cfg.blend_mode = {"reference": 0, "readme": 1}[blend_mode]
        self.cfg = cfg
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)          # make torch's primary context current first
            h = C.c_void_p()
            rc = self.lib.dfd_create(C.byref(cfg), C.byref(h))
            if rc != 0:
                raise _lib.DfdError(f"dfd_create failed ({rc}): {self.lib.dfd_last_error(None).decode()}")
        self.h = h
        self.has_weights = False

    # -- plumbing -----------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.dfd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise _lib.DfdError(f"{what} failed ({rc}): {self.lib.dfd_last_error(self.h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, a, dtype):
        if torch.is_tensor(a):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    @property
    def launches(self):
        return int(self.lib.dfd_launch_count(self.h))

    # -- weights --------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        sd = _weights.extract_state_dict(state_dict)
        blob = np.ascontiguousarray(_weights.pack_state_dict(sd), np.float32)
        assert blob.size == self.lib.dfd_weights_blob_floats()
        self._check(self.lib.dfd_load_weights(self.h, blob.ctypes.data_as(C.c_void_p), blob.size), "dfd_load_weights")
        self.has_weights = True
        return _weights.check_keys(sd)

    # -- forensics ------------------------------------------------------------------
    def forensics_batch(self, frames, stream_ids, full):
        """frames: (n,H,W,3) uint8 CUDA tensor (BGR).  Returns a (n,) uint8-backed CUDA tensor of
        dfd_forensic_result records; use ``forensic_to_numpy``."""
        n, H, W, _ = frames.shape
        assert frames.dtype == torch.uint8 and frames.is_cuda and frames.stride(3) == 1 and frames.stride(2) == 3
        sid = self._dev(stream_ids, torch.int32)
        fl = self._dev(full, torch.uint8)
        out = torch.empty(n * _lib.FORENSIC_BYTES, dtype=torch.uint8, device=self.device)
        rc = self.lib.dfd_forensics_batch(self.h, _ptr(frames), n, H, W, frames.stride(0), frames.stride(1), _ptr(sid),
                                          _ptr(fl), _ptr(out), self._stream())
        self._check(rc, "dfd_forensics_batch")
        return out

    @staticmethod
    def forensic_to_numpy(buf):
        return buf.cpu().numpy().view(FORENSIC_DTYPE)

    @staticmethod
    def records_to_numpy(buf):
        return buf.cpu().numpy().view(RECORD_DTYPE)

    # -- face path --------------------------------------------------------------------
    def face_prep_batch(self, frames, boxes, frame_idx, dtype="fp32"):
        n, H, W, _ = frames.shape
        code, tdt = DTYPES[dtype]
        bx = self._dev(boxes, torch.int32)
        fi = self._dev(frame_idx, torch.int32)
        m = bx.shape[0]
        out = torch.empty((m, 224, 224, 3), dtype=tdt, device=self.device)
        rc = self.lib.dfd_face_prep_batch(self.h, _ptr(frames), n, H, W, frames.stride(0), frames.stride(1), _ptr(bx),
                                          _ptr(fi), m, _ptr(out), code, self._stream())
        self._check(rc, "dfd_face_prep_batch")
        return out

    def effnet_forward(self, x_nhwc):
        """x: (m,224,224,3) float32 or bfloat16 CUDA tensor, NHWC.  Returns logits (m,) float32."""
        code = _lib.F32 if x_nhwc.dtype == torch.float32 else _lib.BF16
        assert x_nhwc.is_contiguous() and x_nhwc.shape[1:] == (224, 224, 3)
        m = x_nhwc.shape[0]
        logits = torch.empty(m, dtype=torch.float32, device=self.device)
        self._check(self.lib.dfd_effnet_forward(self.h, _ptr(x_nhwc), m, code, _ptr(logits), self._stream()),
                    "dfd_effnet_forward")
        return logits

    def face_probability(self, logits, boxes):
        bx = self._dev(boxes, torch.int32)
        m = logits.shape[0]
        prob = torch.empty(m, dtype=torch.float64, device=self.device)
        self._check(self.lib.dfd_face_probability(self.h, _ptr(logits), _ptr(bx), m, _ptr(prob), self._stream()),
                    "dfd_face_probability")
        return prob

    # -- test-time augmentation / calibration --------------------------------------------
    def face_prep_tta(self, frames, boxes, frame_idx, params, dtype="fp32"):
        """analyze_face_with_tta's inputs (deepfake_detection.py:408-438) for m boxes: ``params[i]`` = the n_pred - 1
        (flip, brightness, angle) triples of box i (``tta.draw_params``).  Returns (m * n_pred, 224, 224, 3): crop
        i * n_pred + j is prediction j of box i (j = 0 un-augmented).  ``boxes`` is a HOST array: the rotation matrices are
        built here from the box sizes after clamping to the frame (the reference's ``face_region.shape``)."""
        from . import tta as _tta
        n, H, W, _ = frames.shape
        code, tdt = DTYPES[dtype]
        bh = np.asarray(boxes, np.int64).reshape(-1, 4)
        m = bh.shape[0]
        n_pred = len(params[0]) + 1
        assert len(params) == m and all(len(p) == n_pred - 1 for p in params)
        recs = []
        for (x, y, w, h), prm in zip(bh, params):
            cw = max(min(x + w, W) - max(x, 0), 0)
            ch = max(min(y + h, H) - max(y, 0), 0)
            recs.append(_tta.pack(prm, int(cw), int(ch)))
        aug = np.concatenate(recs) if n_pred > 1 else np.zeros(0, _tta.AUG_DTYPE)
        aug_dev = torch.from_numpy(aug.view(np.uint8).copy()).to(self.device) if aug.size else None
        bx = self._dev(bh.astype(np.int32), torch.int32)
        fi = self._dev(frame_idx, torch.int32)
        out = torch.empty((m * n_pred, 224, 224, 3), dtype=tdt, device=self.device)
        rc = self.lib.dfd_face_prep_tta(self.h, _ptr(frames), n, H, W, frames.stride(0), frames.stride(1), _ptr(bx), _ptr(fi),
                                        m, n_pred, _ptr(aug_dev), _ptr(out), code, self._stream())
        self._check(rc, "dfd_face_prep_tta")
        return out

    def face_probability_tta(self, logits, boxes, n_pred):
        bx = self._dev(boxes, torch.int32)
        m = bx.shape[0]
        assert logits.shape[0] == m * n_pred
        prob = torch.empty(m, dtype=torch.float64, device=self.device)
        self._check(self.lib.dfd_face_probability_tta(self.h, _ptr(logits), _ptr(bx), m, n_pred, _ptr(prob), self._stream()),
                    "dfd_face_probability_tta")
        return prob

    def set_calibrator(self, kind="none", xs=(), ys=()):
        """apply_calibration on the device: kind "none" | "logistic" (xs = [coef], ys = [intercept]) | "piecewise_linear"."""
        code = {"none": 0, "logistic": 1, "piecewise_linear": 2}[kind]
        xa = np.ascontiguousarray(xs, np.float64)
        ya = np.ascontiguousarray(ys, np.float64)
        assert xa.shape == ya.shape
        self._check(self.lib.dfd_set_calibrator(self.h, code, int(xa.size), xa.ctypes.data_as(C.c_void_p),
                                                ya.ctypes.data_as(C.c_void_p), self._stream()), "dfd_set_calibrator")

    # -- result annotation ------------------------------------------------------------------
    def draw_overlay(self, frame_dev, command_list):
        """Composite an ``overlay.CommandList`` on an (H, W, 3) uint8 CUDA frame, in place (dfd_draw_overlay)."""
        assert frame_dev.dtype == torch.uint8 and frame_dev.is_cuda and frame_dev.dim() == 3 and frame_dev.stride(2) == 1 \
            and frame_dev.stride(1) == 3
        cmds, masks = command_list.pack()
        H, W = int(frame_dev.shape[0]), int(frame_dev.shape[1])
        rc = self.lib.dfd_draw_overlay(self.h, _ptr(frame_dev), H, W, frame_dev.stride(0), cmds.ctypes.data_as(C.c_void_p),
                                       int(cmds.size), masks.ctypes.data_as(C.c_void_p), int(masks.size), self._stream())
        self._check(rc, "dfd_draw_overlay")
        return frame_dev

    # -- vote -------------------------------------------------------------------------
    def vote_update(self, stream_ids, vote_input, np_flags=None):
        sid = self._dev(stream_ids, torch.int32)
        vi = self._dev(vote_input, torch.float64)
        fl = self._dev(np_flags, torch.uint8) if np_flags is not None else None
        n = sid.shape[0]
        rec = torch.empty(n * _lib.RECORD_BYTES, dtype=torch.uint8, device=self.device)
        self._check(self.lib.dfd_vote_update(self.h, _ptr(sid), _ptr(vi), _ptr(fl), n, _ptr(rec), self._stream()),
                    "dfd_vote_update")
        return rec

    def gemm_selftest(self, M, N, K, act=1, with_residual=0):
        err = C.c_double(-1.0)
        self._check(self.lib.dfd_gemm_selftest(self.h, M, N, K, act, with_residual, C.byref(err), self._stream()),
                    "dfd_gemm_selftest")
        return err.value

    def gemm_tf32_selftest(self, M, N, K, act=1, mode=0, iters=0):
        """-> (max relative error vs an fp64-accumulated reference, mean ms per launch or 0)."""
        err, ms = C.c_double(-1.0), C.c_double(0.0)
        self._check(self.lib.dfd_gemm_tf32_selftest(self.h, M, N, K, act, mode, iters, C.byref(err), C.byref(ms), self._stream()),
                    "dfd_gemm_tf32_selftest")
        return err.value, ms.value

    def analyze_batch(self, frames, stream_ids, full, boxes, box_frame, dtype="fp32", want_forensic=False,
                      records_out=None):
        """Whole per-frame path for one frame per stream (see dfd_analyze_batch).  dtype: "fp32" = the parity-green accuracy
        mode on the tensor cores (3xTF32, |dp| <= 1e-4: default), "bf16" = 2x faster, misses the 5e-3 gate on the synthetic
        weights."""
        n, H, W, _ = frames.shape
        code, _t = DTYPES[dtype]
        sid = self._dev(stream_ids, torch.int32)
        fl = self._dev(full, torch.uint8)
        m = 0 if boxes is None else int(boxes.shape[0])
        bx = self._dev(boxes, torch.int32) if m else None
        bf = self._dev(box_frame, torch.int32) if m else None
        rec = records_out if records_out is not None else torch.empty(n * _lib.RECORD_BYTES, dtype=torch.uint8,
                                                                      device=self.device)
        fres = torch.empty(n * _lib.FORENSIC_BYTES, dtype=torch.uint8, device=self.device) if want_forensic else None
        fprob = torch.empty(max(m, 1), dtype=torch.float64, device=self.device)
        rc = self.lib.dfd_analyze_batch(self.h, _ptr(frames), n, H, W, frames.stride(0), frames.stride(1), _ptr(sid),
                                        _ptr(fl), _ptr(bx), _ptr(bf), m, code, _ptr(fres), _ptr(fprob), _ptr(rec),
                                        self._stream())
        self._check(rc, "dfd_analyze_batch")
        return rec, fres, fprob[:m]

    # -- frame ingest ------------------------------------------------------------------
    @staticmethod
    def jpeg_info(data):
        """(H, W, components, luma_h, luma_v) of a JPEG stream (host-side header peek); raises DfdError if the stream is not
        decodable on the device (progressive, restart markers, ...)."""
        lib = _lib.load()
        buf = np.frombuffer(data, np.uint8)
        info = (C.c_int32 * 5)()
        rc = lib.dfd_jpeg_info(buf.ctypes.data_as(C.c_void_p), buf.size, info)
        if rc != 0:
            raise _lib.DfdError(f"dfd_jpeg_info: stream not decodable on the device ({rc})")
        return tuple(int(v) for v in info)

    def pack_jpegs(self, streams):
        """Concatenates JPEG byte strings into one PINNED host buffer + offsets (what dfd_decode_jpeg_batch takes)."""
        sizes = [len(s) for s in streams]
        offsets = np.zeros(len(streams) + 1, np.int64)
        np.cumsum(sizes, out=offsets[1:])
        buf = torch.empty(int(offsets[-1]), dtype=torch.uint8).pin_memory()
        view = buf.numpy()
        for s, o in zip(streams, offsets[:-1]):
            view[o:o + len(s)] = np.frombuffer(s, np.uint8)
        return buf, offsets

    def decode_jpeg_batch(self, packed, offsets, H, W, out=None):
        """cv2.imdecode on the device for n streams of equal size -> ((n, H, W, 3) uint8 BGR CUDA tensor, int32 status tensor).
        `packed`: uint8 host tensor (pinned for an asynchronous copy), `offsets`: int64 numpy array of n + 1 offsets."""
        n = len(offsets) - 1
        frames = out if out is not None else torch.empty((n, H, W, 3), dtype=torch.uint8, device=self.device)
        status = torch.empty(n, dtype=torch.int32, device=self.device)
        off = np.ascontiguousarray(offsets, np.int64)
        rc = self.lib.dfd_decode_jpeg_batch(self.h, C.c_void_p(packed.data_ptr()), off.ctypes.data_as(C.c_void_p), n, H, W,
                                            _ptr(frames), frames.stride(0), frames.stride(1), _ptr(status), self._stream())
        self._check(rc, "dfd_decode_jpeg_batch")
        return frames, status

    def configure_stream(self, stream_id, window_size=60, voting_window=10, detection_threshold=0.5):
        self._check(self.lib.dfd_configure_stream(self.h, int(stream_id), int(window_size), int(voting_window),
                                                  float(detection_threshold), self._stream()), "dfd_configure_stream")

    def capture_step(self, frames, stream_ids, full, boxes, box_frame, dtype="fp32", records_out=None, warmup=2):
        """Capture one dfd_analyze_batch call on fixed device buffers into a CUDA graph (the ~100 kernel launches
        of a step, including the forensic / classifier fork-join, replay as one submission).  Returns an object
        with .replay() and .records; the caller refreshes the contents of `frames` / `boxes` between replays."""
        sid = self._dev(stream_ids, torch.int32)
        fl = self._dev(full, torch.uint8)
        bx = self._dev(boxes, torch.int32)
        bf = self._dev(box_frame, torch.int32)
        n = frames.shape[0]
        rec = records_out if records_out is not None else torch.empty(n * _lib.RECORD_BYTES, dtype=torch.uint8, device=self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):                 # sizes every workspace and sets kernel attributes before capture
                self.analyze_batch(frames, sid, fl, bx, bf, dtype=dtype, records_out=rec)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            _, _, fprob = self.analyze_batch(frames, sid, fl, bx, bf, dtype=dtype, records_out=rec)
        torch.cuda.current_stream(self.device).wait_stream(side)

        class _Step:
            pass
        st = _Step()
        st.graph, st.records, st.face_prob, st.replay = graph, rec, fprob, graph.replay
        st._keep = (frames, sid, fl, bx, bf, side)
        return st

    def reset(self, stream_id=-1):
        self._check(self.lib.dfd_reset_stream(self.h, int(stream_id), self._stream()), "dfd_reset_stream")

    def reset_forensics(self, stream_id):
        self._check(self.lib.dfd_reset_stream_part(self.h, int(stream_id), 1, self._stream()), "dfd_reset_stream_part")

    def reset_tracker(self, stream_id):
        self._check(self.lib.dfd_reset_stream_part(self.h, int(stream_id), 2, self._stream()), "dfd_reset_stream_part")

    def gemm_bench(self, M, N, K, act=1, flags=0, iters=10):
        ms = C.c_double(0.0)
        self._check(self.lib.dfd_gemm_bench(self.h, M, N, K, act, flags, iters, C.byref(ms), self._stream()), "dfd_gemm_bench")
        return ms.value

    # -- per-kernel timing ------------------------------------------------------------
    def profile_start(self):
        self._check(self.lib.dfd_profile_start(self.h, self._stream()), "dfd_profile_start")

    def profile_stop(self):
        """-> [(kernel:label, launches, total_ms)] in order of first appearance."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self.lib.dfd_profile_stop(self.h, buf, len(buf), self._stream()), "dfd_profile_stop")
        out = []
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(",", 2)
            out.append((name, int(cnt), float(ms)))
        return out

    # -- diagnostics ------------------------------------------------------------------
    def dbg_tiles(self, n):
        tile = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device=self.device)
        gray = torch.empty((n, 256, 256), dtype=torch.uint8, device=self.device)
        self._check(self.lib.dfd_dbg_tiles(self.h, _ptr(tile), _ptr(gray), n, self._stream()), "dfd_dbg_tiles")
        return tile, gray

    def dbg_jpeg_roundtrip(self, tiles):
        out = torch.empty_like(tiles)
        self._check(self.lib.dfd_dbg_jpeg_roundtrip(self.h, _ptr(tiles), _ptr(out), tiles.shape[0], self._stream()),
                    "dfd_dbg_jpeg_roundtrip")
        return out

    def dbg_canny(self, gray):
        out = torch.empty_like(gray)
        self._check(self.lib.dfd_dbg_canny(self.h, _ptr(gray), _ptr(out), gray.shape[0], self._stream()), "dfd_dbg_canny")
        return out

    def dbg_face160(self, i):
        out = torch.empty((160, 160, 3), dtype=torch.uint8, device=self.device)
        self._check(self.lib.dfd_dbg_face160(self.h, i, _ptr(out), self._stream()), "dfd_dbg_face160")
        return out

    def dbg_face_clahe(self, frames, boxes, frame_idx, i):
        n, H, W, _ = frames.shape
        bx = self._dev(boxes, torch.int32)
        fi = self._dev(frame_idx, torch.int32)
        w, h = int(bx[i, 2]), int(bx[i, 3])
        out = torch.empty((h, w, 3), dtype=torch.uint8, device=self.device)
        rc = self.lib.dfd_dbg_face_clahe(self.h, _ptr(frames), H, W, frames.stride(0), frames.stride(1), _ptr(bx), _ptr(fi),
                                         i, _ptr(out), self._stream())
        self._check(rc, "dfd_dbg_face_clahe")
        return out

    def set_tap(self, name):
        self._check(self.lib.dfd_dbg_set_tap(self.h, (name or "").encode()), "dfd_dbg_set_tap")

    def set_option(self, name, value):
        self._check(self.lib.dfd_dbg_set_option(self.h, name.encode(), int(value)), "dfd_dbg_set_option")

    def activation(self, name):
        n = self.lib.dfd_dbg_activation(self.h, name.encode(), None, 0, self._stream())
        if n < 0:
            self._check(int(n), "dfd_dbg_activation")
        out = torch.empty(int(n), dtype=torch.float32, device=self.device)
        self.lib.dfd_dbg_activation(self.h, name.encode(), _ptr(out), n, self._stream())
        return out
