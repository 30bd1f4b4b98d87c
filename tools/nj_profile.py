"""Per-phase cycle breakdown of the pruned NJ kernel (set DIPB_NJ_PROFILE=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api
from bench import gen_data
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
P = gen_data(n, L, 1)
ctx = api.Context(0)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx)
msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
for rep in range(int(os.environ.get("NJ_REPS", "2"))):
    nj = api.NJDeviceArrays(ctx)
    nj.getDismatrix(n, prm, msaDeviceArrays=msa)
    nj.findNeighbourJoiningTree(["T%d" % i for i in range(n)], int(os.environ.get("NJ_ALGO", "0")))
    print("n=%d dist %.2f ms  nj %.2f ms  (%.2f us/iter)" % (n, ctx.elapsed_ms(api.T_MSA_DIST), ctx.elapsed_ms(api.T_NJ), 1e3 * ctx.elapsed_ms(api.T_NJ) / n), ctx.nj_stats())
    nj.deallocateDeviceArrays()
