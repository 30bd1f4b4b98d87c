"""Host-side partitioning for the phases that shard over GPUs."""
import numpy as np


def row_block_shards(n, world, tile=128):
    """Row ranges [r0, r1) of the lower triangle per rank, balanced by triangle AREA and aligned
    to the 128-row tile of the distance kernel (a tile is computed by exactly one rank, so the
    per-rank matrices sum to the full matrix)."""
    cuts = [int(round(n * np.sqrt(k / world) / tile)) * tile for k in range(world + 1)]
    cuts[0], cuts[-1] = 0, n
    for k in range(1, world + 1):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def triangle_area(r0, r1):
    return (r1 * (r1 - 1) - r0 * (r0 - 1)) // 2


def split_units(num_units, world):
    """Contiguous, near-equal split of independent units (queries, clusters) over ranks."""
    base, rem = divmod(num_units, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def balance_clusters(sizes, world):
    """Contiguous cluster ranges [c0, c1) per rank balanced by the in-cluster work
    ~ 10*s + s^2/2 pair distances for a cluster of s tips."""
    sizes = np.asarray(sizes, np.float64)
    cost = 10.0 * sizes + 0.5 * sizes * sizes + 8.0
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world)))
    cuts.append(len(sizes))
    for k in range(1, len(cuts)):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]
