"""Kernel-tuning aid: time GEMM shapes of the classifier with parts of the kernel disabled (dfd_gemm_bench flags)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfd_b200  # noqa
from dfd_b200.engine import Engine
e = Engine(device=0, max_streams=4, max_batch=4, max_crop=64)
shapes = [("b0.project", 3211264, 16, 32, 0, 4), ("b1.project", 802816, 24, 96, 0, 4), ("b2.project", 802816, 24, 144, 0, 4 | 8),
          ("b1.expand", 3211264, 96, 16, 1, 0), ("b4.project", 200704, 40, 240, 0, 12), ("b9.project", 50176, 112, 672, 0, 12),
          ("b12.project", 12544, 192, 1152, 0, 12), ("head", 12544, 1280, 320, 1, 0)]
for name, M, N, K, act, fl in shapes:
    row = []; sys.stdout.flush()
    for extra in (0, 1, 2, 16):
        ms = e.gemm_bench(M, N, K, act, fl | extra, 10)
        row.append(f"{ms*1e3:7.1f}")
    by = (M * K + M * N * (2 if fl & 8 else 1)) * 2
    print(f"{name:12s} M={M:8d} N={N:4d} K={K:4d}  full/nostore/nomath/noepi us: {' '.join(row)}   bound {by/6.4e6:6.1f} us")
e.close()
