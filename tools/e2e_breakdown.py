"""Wall-clock breakdown of the e2e path (host buffers -> tree) for the bench workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api
from bench import gen_data
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
L = 30000
P = gen_data(n, L, 1)
lens = np.full(n, L, np.uint64)
ctx = api.Context(0)
prm = api.Param(distanceType=2, in_="m")
names = ["T%d" % i for i in range(n)]
for rep in range(3):
    t = [time.time()]
    msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, lens, n, prm); t.append(time.time())
    M = msa.distMatrix(prm); ctx.sync(); t.append(time.time())
    nj = api.NJDeviceArrays(ctx); nj.matrix, nj.d_numSequences = M, n
    nwk = nj.findNeighbourJoiningTree(names); t.append(time.time())
    nj.deallocateDeviceArrays(); t.append(time.time())
    msa.deallocateDeviceArrays(); t.append(time.time())
    d = np.diff(t) * 1e3
    print("rep %d: upload %.1f  dist(call) %.1f [kernel %.1f]  nj(call incl newick) %.1f [kernel %.1f]  free matrix %.1f  free msa %.1f  total %.1f ms"
          % (rep, d[0], d[1], ctx.elapsed_ms(api.T_MSA_DIST), d[2], ctx.elapsed_ms(api.T_NJ), d[3], d[4], d.sum()))
