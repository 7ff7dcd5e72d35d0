import sys; sys.path.insert(0, '.')
import dfd_b200
from dfd_b200.engine import Engine
e = Engine(device=0, max_streams=4, max_batch=4, max_crop=64)
for (M,N,K,act,res) in [(256,96,16,0,0),(12544,96,16,1,0),(12544,16,32,0,0),(784,240,40,1,1),(245,320,1152,0,1),(4900,1280,320,1,0),(300,24,96,0,1),(1813,192,1152,0,3),(12544,40,144,0,2),(3000,320,1152,0,2),(5000,1280,320,1,0),(5000,480,80,1,0)]:
    print((M,N,K,act,res), e.gemm_selftest(M,N,K,act,res), flush=True)
