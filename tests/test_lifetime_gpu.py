"""Handle life times across the C ABI: children keep the context alive (round-1 smoke() crashed with a
use-after-free when the context was destroyed first)."""
import ctypes as C

import numpy as np
import pytest

from dipper_b200 import api, synth
from dipper_b200._lib import lib

pytestmark = pytest.mark.gpu


def test_children_may_outlive_dipb_destroy():
    n, L = 40, 600
    codes, _ = synth.evolve(n, L, seed=3)
    P = synth.pack4_np(codes)
    ctx = api.Context(0)
    raw = ctx.h
    assert lib().dipb_ctx_refs(raw) == 1
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    assert lib().dipb_ctx_refs(raw) == 2
    nj = api.NJDeviceArrays(ctx)
    nj.getDismatrix(n, prm, msaDeviceArrays=msa)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findPlacementTree(prm, msaDeviceArrays=msa)
    refs = lib().dipb_ctx_refs(raw)
    assert refs >= 4                      # creator + msa + matrix + tree
    ctx.close()                           # dipb_destroy FIRST: only drops the creator's reference
    assert lib().dipb_ctx_refs(raw) == refs - 1
    D = nj.matrix.to_host()               # the children still work: stream and events are alive
    assert D.shape == (n, n) and np.all(np.diag(D) == 0)
    row = msa.distConstructionOnGpu(prm, n - 1)
    assert np.allclose(row[: n - 1], D[n - 1, : n - 1], rtol=1e-12)
    msa.deallocateDeviceArrays()
    nj.matrix.free()
    assert lib().dipb_ctx_refs(raw) == refs - 3
    kp.deallocateDeviceArrays()           # last reference: the context is torn down here


def test_failed_upload_leaks_no_reference():
    ctx = api.Context(0)
    raw = ctx.h
    msa = api.MSADeviceArrays(ctx)
    with pytest.raises(api.DipperError):
        msa.allocateDeviceArrays([np.zeros(4, np.uint64), np.zeros(2, np.uint64)], np.array([64, 32], np.uint64), 2)
    assert lib().dipb_ctx_refs(raw) == 1
    ctx.close()
