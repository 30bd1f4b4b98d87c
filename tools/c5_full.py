"""BASELINE configs[4] on ONE GPU: divide-and-conquer (-m 3) of n aligned tips x L sites, backbone n / 20.
Synthetic, seeded.  Prints one JSON line and writes gpurun_out/r1_c5_full_1gpu.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
t0 = time.time()
codes, _ = synth.evolve(n, L, seed=4, gap_cols=0.0, gap_runs=False)
P = synth.pack4_np(codes)
del codes
t_gen = time.time() - t0
print("generated in %.1f s" % t_gen, flush=True)
ctx = api.Context(0)
prm = api.Param(distanceType=2, in_="m")
t0 = time.time()
msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
t_up = time.time() - t0
kp = api.KPlacementDeviceArrays(ctx); kp.allocateDeviceArrays(n)
t0 = time.time(); kp.findTreeDC(prm, msaDeviceArrays=msa); wall = time.time() - t0
out = {"config": "C5 on one GPU: divide-and-conquer -m 3, %d aligned x %d, backbone %d" % (n, L, n // 20), "upload_s": t_up,
       "dc_ms": ctx.elapsed_ms(api.T_PLACE), "dc_wall_s": wall, "tips_per_s": n / wall,
       "clusters": int(len(set(kp.clusterID[kp.clusterID >= 0]))), "host_gen_s": t_gen}
print(json.dumps(out))
json.dump(out, open(os.path.join("gpurun_out", "r1_c5_full_1gpu.json"), "w"), indent=1)
