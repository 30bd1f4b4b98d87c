"""One JC matrix at n x L through the tensor-core kernel (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api
from bench import gen_data
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
P = gen_data(n, L, 1)
ctx = api.Context(0)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx)
msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
for rep in range(int(sys.argv[3]) if len(sys.argv) > 3 else 1):
    M = msa.distMatrix(prm)
    print("dist %.2f ms" % ctx.elapsed_ms(api.T_MSA_DIST))
    M.free()
