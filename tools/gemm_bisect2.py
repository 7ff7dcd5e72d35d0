"""Kernel-tuning aid for the deep-K project GEMMs: A_TMA vs A_SCALE, residual on/off, narrower N, fewer rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfd_b200  # noqa
from dfd_b200.engine import Engine
e = Engine(device=0, max_streams=4, max_batch=4, max_crop=64)
cases = [("b12 plain", 12544, 192, 1152, 0, 0), ("b12 scale", 12544, 192, 1152, 0, 4), ("b12 scale+res", 12544, 192, 1152, 0, 12),
         ("b12 N=96 scale", 12544, 96, 1152, 0, 4), ("b12 N=64 scale", 12544, 64, 1152, 0, 4),
         ("b12 2xM scale", 25088, 192, 1152, 0, 4), ("b12 K=576 scale", 12544, 192, 576, 0, 4),
         ("b9 plain", 50176, 112, 672, 0, 0), ("b9 scale", 50176, 112, 672, 0, 4), ("b9 scale+res", 50176, 112, 672, 0, 12),
         ("b15 scale", 12544, 320, 1152, 0, 4), ("head", 12544, 1280, 320, 1, 0)]
for name, M, N, K, act, fl in cases:
    row = []
    for extra in (0, 1, 2, 16):
        ms = e.gemm_bench(M, N, K, act, fl | extra, 20)
        row.append(f"{ms*1e3:7.1f}")
    print(f"{name:16s} M={M:6d} N={N:4d} K={K:4d}  full/nostore/nomath/noepi us: {' '.join(row)}", flush=True)
e.close()
