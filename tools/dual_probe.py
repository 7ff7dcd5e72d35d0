"""GPU probe: classifier at batch 256, one chain vs two concurrent half-batch chains (python tools/dual_probe.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dfd_b200  # noqa
from dfd_b200 import synth
from dfd_b200.engine import Engine
e = Engine(device=0, max_streams=4, max_batch=256, max_crop=64)
e.load_state_dict(synth.make_state_dict())
g = torch.Generator().manual_seed(1)
for dt in (torch.float32, torch.bfloat16):
    x = synth._calib_batch(g, 256).float().permute(0, 2, 3, 1).contiguous().cuda().to(dt)
    ref = None
    for dual in (0, 1, 0, 1):
        e.set_option("dual_chain", dual)
        z = e.effnet_forward(x)
        torch.cuda.synchronize()
        if ref is None:
            ref = z.clone()
        same = bool(torch.equal(z, ref))
        gr = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            e.effnet_forward(x)
            with torch.cuda.graph(gr, stream=s):
                e.effnet_forward(x)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            gr.replay()
        t0.record()
        for _ in range(20):
            gr.replay()
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 20
        print(f"{dt} dual={dual}: {ms:.3f} ms / 256 crops = {256 / ms * 1e3:.0f} crops/s  bit-identical={same}", flush=True)
e.close()
