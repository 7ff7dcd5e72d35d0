"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_step.py
One JPEG decode, one dfd_analyze_batch step in both precisions (fused MBConv front kernel, tcgen05 GEMMs, 3xTF32 GEMM with
chunked accumulation, cluster SE kernel, forensic + face-prep + vote kernels) and the GEMM self-tests on ragged shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch
import dfd_b200  # noqa
from dfd_b200 import synth
from dfd_b200.engine import Engine

e = Engine(device=0, max_streams=4, max_batch=4, max_crop=256, detection_threshold=0.55)
e.load_state_dict(synth.make_state_dict())
rng = np.random.RandomState(1)
imgs = [synth.make_frame(f, 200, 264, rng) for f in ("pink", "gradient", "natural")]
streams = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 85])[1].tobytes() for im in imgs]
packed, off = e.pack_jpegs(streams)
frames, status = e.decode_jpeg_batch(packed, off, 200, 264)
torch.cuda.synchronize()
assert status.cpu().tolist() == [0, 0, 0]
for i, s in enumerate(streams):
    assert np.array_equal(frames[i].cpu().numpy(), cv2.imdecode(np.frombuffer(s, np.uint8), cv2.IMREAD_COLOR))
boxes = synth.make_boxes(3, 200, 264, rng, lo=60, hi=160)
for dtype in ("bf16", "fp32"):
    for t in range(2):
        rec, fres, fp = e.analyze_batch(frames, [0, 1, 2], [int(t == 0)] * 3, boxes, [0, 1, 2], dtype=dtype, want_forensic=True)
    torch.cuda.synchronize()
    print(dtype, "face probabilities", fp.cpu().numpy())
for (M, N, K, act, mode) in ((300, 96, 16, 1, 0), (49 * 5, 192, 672, 0, 3), (129, 24, 144, 0, 1), (500, 1280, 320, 1, 0)):
    err, _ = e.gemm_tf32_selftest(M, N, K, act, mode)
    assert err < 4e-6, (M, N, K, err)
    err = e.gemm_selftest(M, N, K, act, mode)
    assert err < 2e-2, (M, N, K, err)
print("sanitize_step ok, launches", e.launches)
e.close()
