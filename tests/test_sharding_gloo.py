"""CPU tests of the N>1 path: shard ranges + merge by reduce, world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dipper_b200 import sharding, synth


def test_row_block_shards_cover_and_balance():
    for n in (130, 1000, 30000):
        for world in (1, 2, 4, 8):
            sh = sharding.row_block_shards(n, world)
            assert sh[0][0] == 0 and sh[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            assert all(r0 % 128 == 0 for r0, _ in sh)
            if n == 30000:
                areas = [sharding.triangle_area(*s) for s in sh]
                assert max(areas) / (sum(areas) / world) < 1.05
    assert sharding.split_units(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]


def _worker(rank, world, port, n, L, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O          # checker stands in for the GPU kernel in this CPU test
    codes, _ = synth.evolve(n, L, seed=5)
    P = synth.pack4_np(codes)
    r0, r1 = sharding.row_block_shards(n, world)[rank]
    part = np.zeros((n, n))
    full = O.msa_dist_matrix(P, L, 2)
    # what dipb_msa_dist_matrix_rows produces for this rank: lower-triangle rows [r0, r1) + mirror
    for i in range(r0, r1):
        part[i, :i] = full[i, :i]
        part[:i, i] = full[i, :i]
    t = torch.from_numpy(part)
    dist.reduce(t, dst=0)
    if rank == 0:
        out.put(bool(np.array_equal(t.numpy(), full)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_matrix_merges_by_reduce_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 300, 400, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok
