"""ncu capture target for the kernels added late in round 1: rank-compressed Mash distances, exact placement
(data flow in one cluster), transposed D&C assignment, D&C cluster placement (small sizes: ncu replays every launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api, synth

ctx = api.Context(0)
n, L = 2000, 10000
codes, _ = synth.evolve(n, L, seed=2, gap_cols=0.01, gap_runs=False)
seqs = synth.unaligned(codes)
prm = api.Param(kmerSize=15, sketchSize=1000, in_="r")
m = api.MashDeviceArrays(ctx)
m.allocateDeviceArrays([synth.pack2_np(s) for s in seqs], np.array([len(s) for s in seqs], np.uint64), n, prm)
m.sketchConstructionOnGpu()
mat = m.distMatrix()
print("mash dist %.2f ms" % ctx.elapsed_ms(api.T_MASH_DIST))
n, L = 6000, 4000
codes, _ = synth.evolve(n, L, seed=3, gap_cols=0.03, gap_runs=False)
P = synth.pack4_np(codes)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
pl = api.PlacementDeviceArrays(ctx); pl.allocateDeviceArrays(n)
pl.findPlacementTree(prm, msaDeviceArrays=msa)
print("exact placement %.2f ms" % ctx.elapsed_ms(api.T_PLACE))
kp2 = api.KPlacementDeviceArrays(ctx); kp2.allocateDeviceArrays(n)
kp2.findTreeDC(prm, msaDeviceArrays=msa)
print("dc %.2f ms" % ctx.elapsed_ms(api.T_PLACE))
