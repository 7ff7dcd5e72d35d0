"""Profiling driver: a few EfficientNet-B0 forward passes at batch 256 (run under ncu; not a benchmark).
   python tools/prof_effnet.py [batch] [reps] [bf16|fp32]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dfd_b200  # noqa: F401
from dfd_b200 import synth
from dfd_b200.engine import Engine

m = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dt = sys.argv[3] if len(sys.argv) > 3 else "bf16"
e = Engine(device=0, max_streams=8, max_batch=m, max_crop=64)
e.load_state_dict(synth.make_state_dict())
g = torch.Generator().manual_seed(1)
x = torch.randn((m, 224, 224, 3), generator=g).cuda()
if dt == "bf16":
    x = x.bfloat16()
for _ in range(reps):
    y = e.effnet_forward(x)
torch.cuda.synchronize()
# graph-replay timing of the classifier alone (CUDA events, 20 replays)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    e.effnet_forward(x)
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_, stream=side):
        y = e.effnet_forward(x)
torch.cuda.synchronize()
if os.environ.get("PROF_TIME", "1") == "1":
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        g_.replay()
    ev0.record()
    for _ in range(20):
        g_.replay()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 20
    print(f"effnet {dt} b{m}: {ms:.4f} ms/forward  {m / ms * 1e3:.0f} crops/s")
e.close()
