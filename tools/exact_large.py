"""Exact placement beyond one cluster's capacity (global-memory data-flow kernel): n aligned tips x L sites, JC."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
codes, _ = synth.evolve(n, L, seed=3, gap_cols=0.0, gap_runs=False)
P = synth.pack4_np(codes)
ctx = api.Context(0)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
pl = api.PlacementDeviceArrays(ctx); pl.allocateDeviceArrays(n)
t0 = time.time(); pl.findPlacementTree(prm, msaDeviceArrays=msa); wall = time.time() - t0
out = {"config": "exact placement -p 0 beyond the cluster capacity (%d > %d tips), %d sites, JC" % (n, api.PlacementDeviceArrays.maxTips(), L),
       "place_ms": ctx.elapsed_ms(api.T_PLACE), "tips_per_s": n / wall, "kernel": "place_exact_global_kernel"}
print(json.dumps(out))
json.dump(out, open(os.path.join("gpurun_out", "r1_exact_placement_large.json"), "w"), indent=1)
