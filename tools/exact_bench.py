"""Exact placement (-p 0) timing: n aligned tips x L sites, JC; prints one JSON line (and the per-tip profile on stderr)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
codes, _ = synth.evolve(n, L, seed=3, gap_cols=0.03, gap_runs=False)
P = synth.pack4_np(codes)
ctx = api.Context(0)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
out = {"config": "exact placement -p 0, %d aligned x %d, JC" % (n, L)}
for rep in range(2):
    pl = api.PlacementDeviceArrays(ctx); pl.allocateDeviceArrays(n)
    t0 = time.time(); pl.findPlacementTree(prm, msaDeviceArrays=msa); wall = time.time() - t0
    out["place_ms"] = ctx.elapsed_ms(api.T_PLACE); out["tips_per_s"] = n / wall
    pl.deallocateDeviceArrays()
kp = api.KPlacementDeviceArrays(ctx); kp.allocateDeviceArrays(n)
kp.findPlacementTree(prm, msaDeviceArrays=msa)
out["kclosest_ms_same_input"] = ctx.elapsed_ms(api.T_PLACE)
print(json.dumps(out))
json.dump(out, open(os.path.join("gpurun_out", "r1_exact_placement.json"), "w"), indent=1)
