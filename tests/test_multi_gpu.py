"""Several GPUs in one process (dipb_multi_*, csrc/multi.cu): same matrix and same D&C tree as one GPU.  Runs on
however many devices the box has (1 still exercises the entry points; `gpurun --gpus 2` the peer copies)."""
import numpy as np
import pytest

from dipper_b200 import api, synth
from conftest import make_msa

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return list(range(min(torch.cuda.device_count(), 4)))


@pytest.mark.parametrize("dist_type", [1, 2])
def test_multi_device_matrix_equals_single(ctx, dist_type):
    n, L = 1500, 2000                                  # several 128-row blocks per device
    codes, P, _ = make_msa(n, L, seed=91, gap_cols=0.03)
    prm = api.Param(distanceType=dist_type, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    D1 = msa.distMatrix(prm).to_host()
    md = api.MultiDevice(_devices())
    md.allocateDeviceArrays(P, L)
    M = md.distMatrix(prm)
    D = M.to_host()
    assert np.array_equal(D, D1)                       # bit-identical, mirrored, zero diagonal
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    M.free()
    md.close()


def test_multi_device_dc_equals_single(ctx, oracle):
    n, L, B = 1200, 1500, 60
    codes, P, _ = make_msa(n, L, seed=92)
    prm = api.Param(distanceType=2, in_="m")
    md = api.MultiDevice(_devices())
    md.allocateDeviceArrays(P, L)
    kp = md.findTreeDC(prm, backboneSize=B)
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    D = msa.distMatrix(prm).to_host()
    ot, ocl = oracle.dc(D, B)
    assert np.array_equal(kp.clusterID, ocl)
    assert kp.printTree(synth.names(n)) == ot.newick(synth.names(n))
    kp.deallocateDeviceArrays()
    md.close()
