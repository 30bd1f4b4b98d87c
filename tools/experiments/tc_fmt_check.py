"""Is the f32 accumulation of tcgen05.mma kind::f8f6f4 exact for sums of +-1 products up to 90 000?  Near-identical sequences
make every match count ~ L; the popcount kernel is the exact reference, the int8 tensor-core kernel (DIPB_TC_FMT=0) and the
e2m1 one (default) must both equal it bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from dipper_b200 import api, synth
n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 1536, int(sys.argv[2]) if len(sys.argv) > 2 else 30000
rng = np.random.default_rng(5)
base = rng.integers(0, 4, L).astype(np.uint8)
codes = np.tile(base, (n, 1))
mut = rng.random((n, L)) < 0.01
codes[mut] = rng.integers(0, 4, int(mut.sum())).astype(np.uint8)
codes[rng.random((n, L)) < 0.002] = 4          # a few gaps / N
P = synth.pack4_np(codes)
lens = np.full(n, L, np.uint64)
ctx = api.Context(0)
prm = api.Param(distanceType=2, in_="m")
os.environ["DIPB_MSA_TC2"] = "1"
res = {}
for tag, env in (("popc", {"DIPB_MSA_TC": "0"}), ("i8", {"DIPB_MSA_TC": "2", "DIPB_TC_FMT": "0"}), ("fmt", {"DIPB_MSA_TC": "2", "DIPB_TC_FMT": "2"})):
    os.environ.update(env)
    msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, lens, n, prm)
    M = msa.distMatrix(prm); res[tag] = M.to_host(); ms = ctx.elapsed_ms(api.T_MSA_DIST)
    M.free(); msa.deallocateDeviceArrays()
    print(tag, "%.2f ms" % ms, flush=True)
print("i8 == popc:", np.array_equal(res["i8"], res["popc"], equal_nan=True))
eq = np.array_equal(res["fmt"], res["popc"], equal_nan=True)
print("fmt == popc:", eq)
if not eq:
    d = np.abs(res["fmt"] - res["popc"]); print("max abs diff", np.nanmax(d), "entries differing", int((d > 0).sum()), "of", d.size)
