"""GPU probe: 3xTF32 GEMM error and time per EfficientNet-B0 layer shape at batch 256 (python tools/tf32_probe.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfd_b200  # noqa
from dfd_b200.engine import Engine

e = Engine(device=0, max_streams=4, max_batch=4, max_crop=64)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
# name, M per image, N, K, act, mode
LAYERS = [("b1.expand", 12544, 96, 16, 1, 0), ("b0.project", 12544, 16, 32, 0, 2), ("b1.project", 3136, 24, 96, 0, 2),
          ("b2.expand", 3136, 144, 24, 1, 0), ("b2.project", 3136, 24, 144, 0, 3), ("b3.project", 784, 40, 144, 0, 2),
          ("b4.expand", 784, 240, 40, 1, 0), ("b4.project", 784, 40, 240, 0, 3), ("b5.project", 196, 80, 240, 0, 2),
          ("b6.expand", 196, 480, 80, 1, 0), ("b6.project", 196, 80, 480, 0, 3), ("b8.project", 196, 112, 480, 0, 2),
          ("b9.expand", 196, 672, 112, 1, 0), ("b9.project", 196, 112, 672, 0, 3), ("b11.project", 49, 192, 672, 0, 2),
          ("b12.expand", 49, 1152, 192, 1, 0), ("b12.project", 49, 192, 1152, 0, 3), ("b15.project", 49, 320, 1152, 0, 2),
          ("head", 49, 1280, 320, 1, 0)]
tot = 0.0
for name, mpi, N, K, act, mode in LAYERS:
    M = mpi * B
    err, ms = e.gemm_tf32_selftest(M, N, K, act, mode, iters=5)
    gb = (M * K + M * N * (2 if mode & 1 else 1)) * 4 / 1e9
    tot += ms
    print(f"{name:12s} M={M:8d} N={N:5d} K={K:5d} mode={mode} err={err:.2e} {ms*1e3:8.1f} us  {gb/ms*1e3/1e3:6.2f} TB/s", flush=True)
print("sum ms", tot)
e.close()

# whole classifier: error vs the fp32 oracle and time, tensor-core (3xTF32) vs CUDA-core fp32
import time
import torch
from dfd_b200 import synth
from oracle import effnet as oeff
sd = synth.make_state_dict()
e = Engine(device=0, max_streams=4, max_batch=256, max_crop=64)
e.load_state_dict(sd)
g = torch.Generator().manual_seed(256)
x = synth._calib_batch(g, 256).float()
idx = list(range(0, 256, 4))
ref = torch.sigmoid(oeff.forward(x[idx], sd).flatten())
xn = x.permute(0, 2, 3, 1).contiguous().cuda()
for simt in (0, 1):
    e.set_option("fp32_simt", simt)
    z = e.effnet_forward(xn)
    torch.cuda.synchronize()
    d = (torch.sigmoid(z.cpu()[idx]) - ref).abs()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        e.effnet_forward(xn)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    print(f"fp32 {'SIMT' if simt else '3xTF32'}: max |dp| {float(d.max()):.2e} mean {float(d.mean()):.2e}  {ms:.2f} ms / 256 crops = {256 / ms * 1e3:.0f} crops/s", flush=True)
e.set_option("fp32_simt", 0)
e.profile_start()
e.effnet_forward(xn)
for name, cnt, ms in e.profile_stop():
    print(f"   {name:40s} {cnt:3d} {ms*1e3:9.1f} us")
e.close()
