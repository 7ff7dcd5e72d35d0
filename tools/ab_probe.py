"""A/B probe: run the 3xTF32 layer table against an alternative build of libdfd (python tools/ab_probe.py <lib.so> | -)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfd_b200  # noqa
from dfd_b200 import _lib
if len(sys.argv) > 1 and sys.argv[1] != "-":
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from dfd_b200.engine import Engine
e = Engine(device=0, max_streams=4, max_batch=4, max_crop=64)
B = 256
LAYERS = [("b1.expand", 12544, 96, 16, 1, 0), ("b0.project", 12544, 16, 32, 0, 2), ("b1.project", 3136, 24, 96, 0, 2),
          ("b2.expand", 3136, 144, 24, 1, 0), ("b2.project", 3136, 24, 144, 0, 3), ("b4.project", 784, 40, 240, 0, 3),
          ("b6.expand", 196, 480, 80, 1, 0), ("b6.project", 196, 80, 480, 0, 3),
          ("b9.expand", 196, 672, 112, 1, 0), ("b9.project", 196, 112, 672, 0, 3), ("b11.project", 49, 192, 672, 0, 2),
          ("b12.expand", 49, 1152, 192, 1, 0), ("b12.project", 49, 192, 1152, 0, 3), ("b15.project", 49, 320, 1152, 0, 2),
          ("head", 49, 1280, 320, 1, 0)]
tot = 0.0
for name, mpi, N, K, act, mode in LAYERS:
    M = mpi * B
    err, ms = e.gemm_tf32_selftest(M, N, K, act, mode, iters=8)
    tot += ms
    print(f"{name:12s} M={M:8d} N={N:5d} K={K:5d} mode={mode} err={err:.2e} {ms*1e3:8.1f} us", flush=True)
print("sum ms", tot, _lib.LIB_PATH)
e.close()
