"""ncu capture target for the secondary kernels: Mash sketching + sketch distances, k-closest placement,
divide-and-conquer (small sizes so that ncu's ~40 replays of every launch stay short)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api, synth

ctx = api.Context(0)
# Mash: 2000 unaligned sequences of ~10 kb
n, L = 2000, 10000
codes, _ = synth.evolve(n, L, seed=2, gap_cols=0.01, gap_runs=False)
seqs = synth.unaligned(codes)
packed = [synth.pack2_np(s) for s in seqs]
lens = np.array([len(s) for s in seqs], np.uint64)
prm = api.Param(kmerSize=15, sketchSize=1000, in_="r")
m = api.MashDeviceArrays(ctx); m.allocateDeviceArrays(packed, lens, n, prm); m.sketchConstructionOnGpu()
mat = m.distMatrix()
print("sketch %.2f ms, mash dist %.2f ms" % (ctx.elapsed_ms(api.T_SKETCH), ctx.elapsed_ms(api.T_MASH_DIST)))
# placement + D&C: 6000 aligned x 4000
n, L = 6000, 4000
codes, _ = synth.evolve(n, L, seed=3, gap_cols=0.03, gap_runs=False)
P = synth.pack4_np(codes)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
kp = api.KPlacementDeviceArrays(ctx); kp.allocateDeviceArrays(n)
kp.findPlacementTree(prm, msaDeviceArrays=msa)
print("placement %.2f ms" % ctx.elapsed_ms(api.T_PLACE))
kp2 = api.KPlacementDeviceArrays(ctx); kp2.allocateDeviceArrays(n)
kp2.findTreeDC(prm, msaDeviceArrays=msa)
print("dc %.2f ms" % ctx.elapsed_ms(api.T_PLACE))
