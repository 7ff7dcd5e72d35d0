"""Tuning probe (not a benchmark): one 256-stream engine vs two 128-stream engines replaying their step graphs concurrently
on two CUDA streams (same GPU).  Prints frames/s of both arrangements."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dfd_b200  # noqa
from dfd_b200 import synth
from dfd_b200.engine import Engine
import bench

S = 256
dev = torch.device("cuda", 0)
sd = synth.make_state_dict()
host_frames, boxes = bench.make_inputs(S, 3, seed=1234)
dev_frames = host_frames.to(dev)

def build(n_eng):
    per = S // n_eng
    out = []
    for e in range(n_eng):
        eng = Engine(device=0, max_streams=per, max_batch=per, max_crop=512, detection_threshold=0.55)
        eng.load_state_dict(sd)
        sids = torch.arange(per, dtype=torch.int32, device=dev)
        bf = torch.arange(per, dtype=torch.int32, device=dev)
        graphs = []
        for k in range(3):
            fl = torch.full((per,), int(k == 0), dtype=torch.uint8, device=dev)
            bx = torch.from_numpy(boxes[k][e * per:(e + 1) * per]).to(dev)
            graphs.append(eng.capture_step(dev_frames[k][e * per:(e + 1) * per], sids, fl, bx, bf, dtype="bf16"))
        eng.reset(-1)
        out.append((eng, graphs, torch.cuda.Stream(dev)))
    return out

def run(engs, steps=30, warm=6):
    cur = torch.cuda.current_stream(dev)
    def step(i):
        if len(engs) == 1:
            engs[0][1][i % 3].replay()
            return
        for eng, graphs, st in engs:
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                graphs[i % 3].replay()
        for eng, graphs, st in engs:
            cur.wait_stream(st)
    for i in range(warm):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warm + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return ms, S / ms * 1e3

for n in (1, 2, 4):
    engs = build(n)
    ms, fps = run(engs)
    print(f"{n} engine(s) x {S // n} streams: {ms:.3f} ms/step  {fps:.0f} frames/s", flush=True)
    for eng, _, _ in engs:
        eng.close()
