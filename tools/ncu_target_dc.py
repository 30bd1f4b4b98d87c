"""ncu capture target: divide-and-conquer kernels only (6000 aligned tips x 4000 sites, backbone 300)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api, synth
ctx = api.Context(0)
n, L = 20000, 2000
codes, _ = synth.evolve(n, L, seed=3, gap_cols=0.03, gap_runs=False)
P = synth.pack4_np(codes)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
kp2 = api.KPlacementDeviceArrays(ctx); kp2.allocateDeviceArrays(n)
kp2.findTreeDC(prm, msaDeviceArrays=msa)
print("dc %.2f ms" % ctx.elapsed_ms(api.T_PLACE))
