import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from dipper_b200 import api, synth
from oracle import oracle
from conftest import make_msa
ctx = api.Context(0)
def run(D, algo):
    nj = api.NJDeviceArrays(ctx); nj.setMatrix(D)
    nj.findNeighbourJoiningTree(synth.names(D.shape[0]), algo)
    r = nj.result; nj.deallocateDeviceArrays(); return r
for n in [int(a) for a in sys.argv[1:]]:
    codes, P, _ = make_msa(n, 900, seed=100 + n)
    D = oracle.msa_dist_matrix(P, 900, 2)
    o = oracle.nj(D)
    for rep in range(4):
        r = run(D, 3)
        bad = [k for k in range(4) if not np.array_equal(r[k], o[k])]
        first = int(np.nonzero((r[0] != o[0]) | (r[1] != o[1]) | (r[2] != o[2]) | (r[3] != o[3]))[0][0]) if bad else -1
        print(n, 'bad arrays', bad, 'first diff merge', first, (r[0][first], r[1][first], r[2][first], r[3][first], o[0][first], o[1][first], o[2][first], o[3][first]) if first >= 0 else '')
