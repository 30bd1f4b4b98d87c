"""BASELINE configs[3] at full size: k-closest placement (-m 1) of n unaligned sequences of ~L bases through Mash
(sketch -> sketch distances in 512-row blocks -> placement).  Synthetic, seeded; sequences carry no gaps so that the
2-bit packing can be vectorised.  Prints one JSON line and writes gpurun_out/r1_c4_full.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
print(open("/proc/meminfo").read().split("\n")[1], flush=True)
t0 = time.time()
codes, _ = synth.evolve(n, L, seed=2, gap_cols=0.0, gap_runs=False)
W = (L + 31) // 32
P = np.zeros((n, W), np.uint64)
sh = np.arange(32, dtype=np.uint64) * np.uint64(2)
for r0 in range(0, n, 20000):
    blk = codes[r0:r0 + 20000]
    pad = np.zeros((blk.shape[0], W * 32), np.uint64)
    pad[:, :L] = blk
    P[r0:r0 + 20000] = (pad.reshape(blk.shape[0], W, 32) << sh).sum(axis=2, dtype=np.uint64)
del codes
t_gen = time.time() - t0
print("generated in %.1f s" % t_gen, flush=True)
ctx = api.Context(0)
prm = api.Param(kmerSize=15, sketchSize=1000, in_="r")
lens = np.full(n, L, np.uint64)
t0 = time.time()
m = api.MashDeviceArrays(ctx); m.allocateDeviceArrays(P, lens, n, prm)
t_up = time.time() - t0
m.sketchConstructionOnGpu()
sketch_ms = ctx.elapsed_ms(api.T_SKETCH)
kp = api.KPlacementDeviceArrays(ctx); kp.allocateDeviceArrays(n)
t0 = time.time(); kp.findPlacementTree(prm, mashDeviceArrays=m); wall = time.time() - t0
nwk = kp.printTree(synth.names(n))
out = {"config": "C4: placement -m 1, %d unaligned x %d (Mash k=15 s=1000)" % (n, L), "upload_s": t_up, "sketch_ms": sketch_ms,
       "kmers_per_s": n * (L - 14) / sketch_ms * 1e3, "place_s": wall, "tips_per_s": n / wall,
       "mash_pairs": n * (n - 1) // 2, "newick_bytes": len(nwk), "host_gen_s": t_gen}
print(json.dumps(out))
json.dump(out, open(os.path.join("gpurun_out", "r1_c4_full.json"), "w"), indent=1)
