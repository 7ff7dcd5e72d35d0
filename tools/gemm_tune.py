import sys; sys.path.insert(0, '.')
import dfd_b200
from dfd_b200.engine import Engine
e = Engine(device=0, max_streams=4, max_batch=4, max_crop=64)
M = 256*112*112
for name,(m,n,k,act) in {"b1.expand":(M,96,16,1),"b2.expand":(256*56*56,144,24,1),"b4.expand":(256*28*28,240,40,1),"b6.expand":(256*196,480,80,1),"b9.expand":(256*196,672,112,1),"b12.expand":(256*49,1152,192,1),"head":(256*49,1280,320,1),"b0.project":(M,16,32,0)}.items():
    by = (m*k+m*n)*2
    for flags in (0,3):
        ms = e.gemm_bench(m,n,k,act,flags,10)
        print(f"{name:11s} flags={flags} {ms:.4f} ms  {by/ms/1e6:8.1f} GB/s", flush=True)
for name,(m,n,k) in {"b2.project":(256*56*56,24,144),"b5.project":(256*196,80,240),"b9.project":(256*196,112,672),"b15.project":(256*49,320,1152)}.items():
    by = (m*k+2*m*n)*2
    for flags in (4|8,):
        ms = e.gemm_bench(m,n,k,0,flags,10)
        print(f"{name:11s} flags={flags} {ms:.4f} ms  {by/ms/1e6:8.1f} GB/s", flush=True)
