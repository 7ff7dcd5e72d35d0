"""GPU probe: device JPEG ingest, per-kernel times (python tools/jpeg_probe.py [family ...])."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2, numpy as np, torch
import dfd_b200  # noqa
from dfd_b200 import synth
from dfd_b200.engine import Engine
fams = sys.argv[1:] or ["uniform", "pink", "blur", "flat", "gradient"]
H, W, N = 720, 1280, 256
e = Engine(device=0, max_streams=4, max_batch=N, max_crop=64)
rng = np.random.RandomState(0)
bases = [synth.make_frame(f, H, W, rng) for f in fams for _ in range(2)]
streams = []
for s in range(N):
    b = np.roll(bases[s % len(bases)], ((s * 7) % 64, (s * 13) % 64), axis=(0, 1))
    streams.append(cv2.imencode(".jpg", b, [cv2.IMWRITE_JPEG_QUALITY, 85])[1].tobytes())
packed, off = e.pack_jpegs(streams)
out = torch.empty((N, H, W, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    e.decode_jpeg_batch(packed, off, H, W, out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    e.decode_jpeg_batch(packed, off, H, W, out=out)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"families {fams}: mean stream {int(off[-1]) // N} B; host side of the call {(t1 - t0) / 5 * 1e3:.2f} ms, wall per call {(t2 - t0) / 5 * 1e3:.2f} ms "
      f"= {N / ((t2 - t0) / 5):.0f} frames/s")
e.profile_start()
e.decode_jpeg_batch(packed, off, H, W, out=out)
for name, cnt, ms in e.profile_stop():
    print(f"   {name:24s} {cnt:3d} {ms:9.3f} ms")
e.close()
