"""GPU probe / profiling driver: test-time augmentation prep (32 boxes x 3 predictions on 720p frames) and the result overlay
(python tools/tta_overlay_probe.py)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dfd_b200  # noqa
from dfd_b200 import overlay, synth, tta
from dfd_b200.engine import Engine
e = Engine(device=0, max_streams=4, max_batch=96, max_crop=512)
rng = np.random.RandomState(0)
frames = torch.from_numpy(np.stack([synth.make_frame("pink", 720, 1280, rng) for _ in range(8)])).cuda()
boxes = synth.make_boxes(32, 720, 1280, rng, lo=96, hi=400)
fidx = np.arange(32, dtype=np.int32) % 8
random.seed(1)
params = [tta.draw_params(3) for _ in range(32)]
for _ in range(2):
    out = e.face_prep_tta(frames, boxes, fidx, params, "fp32")
torch.cuda.synchronize()
e.profile_start()
out = e.face_prep_tta(frames, boxes, fidx, params, "fp32")
for name, cnt, ms in e.profile_stop():
    print(f"   {name:24s} {cnt:3d} {ms * 1e3:9.1f} us")
fr = frames[0].clone()
cl = overlay.CommandList(720, 1280)
overlay.detection_overlay(cl, 400, 200, 300, 280, 0.83, "FAKE", {"fake_count": 7, "real_count": 3, "total_frames": 10})
overlay.frame_analysis_overlay(cl, 0.4, "REAL", {"scores": {"frequency": 0.2, "noise": 0.5, "ela": 0.1, "edge": 0.6}})
for _ in range(2):
    e.draw_overlay(fr, cl)
torch.cuda.synchronize()
e.profile_start()
e.draw_overlay(fr, cl)
for name, cnt, ms in e.profile_stop():
    print(f"   {name:24s} {cnt:3d} {ms * 1e3:9.1f} us")
e.close()
