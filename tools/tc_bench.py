"""Times the tensor-core distance path against the popcount path on the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api
from bench import gen_data
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
P = gen_data(n, L, 1)
ctx = api.Context(0)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
res = {}
for mode in ("0", "1", "1", "0"):
    os.environ["DIPB_MSA_TC"] = mode
    M = msa.distMatrix(prm); ctx.sync()
    res.setdefault(mode, []).append(ctx.elapsed_ms(api.T_MSA_DIST))
    if mode == "1" and "tc" not in res: res["tc"] = M.to_host()
    if mode == "0" and "pc" not in res: res["pc"] = M.to_host()
    M.free()
print("n=%d L=%d popcount ms %s  tensor-core ms %s  identical %s" % (n, L, res["0"], res["1"], np.array_equal(res["tc"], res["pc"], equal_nan=True)))
