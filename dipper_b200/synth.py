"""Seeded synthetic "alisim-shaped" inputs for tests and bench.py.

iqtree2/alisim is not installed, so the generator lives here: Yule-Harding tree, Exp
branch lengths clipped to [min,max], JC69 substitutions (uniformisation), optional
Gamma(0.5)x4 + invariant-site rate heterogeneity, gap runs + gap columns.  Codes are
the reference's 4-bit alphabet: A0 C1 G2 T3, 4 = anything else
(src/fourBitCompressor.cpp:17-36).
"""
import os

import numpy as np

REGIMES = {
    # name: (min, mean, max) branch length in substitutions/site
    "alisim": (2e-6, 2e-5, 2e-4),     # scripts/alisim.sh-like: tie-heavy, many identical tips
    "tiefree": (1e-3, 1e-2, 1e-1),    # used for the strict RF = 0 gates and the bench
}


def yule_tree(n, rng, regime="tiefree"):
    """Returns (parent, bl, children, leaf_nodes): nodes 0..2n-2, root = 0."""
    lo, mean, hi = REGIMES[regime]
    parent = [-1]
    children = [[]]
    leaves = [0]
    while len(leaves) < n:
        k = int(rng.integers(0, len(leaves)))
        v = leaves[k]
        a, b = len(parent), len(parent) + 1
        parent += [v, v]
        children += [[], []]
        children[v] = [a, b]
        leaves[k] = a
        leaves.append(b)
    m = len(parent)
    bl = np.clip(rng.exponential(mean, m), lo, hi)
    bl[0] = 0.0
    return np.array(parent), bl, children, leaves


def evolve(n, L, seed=1, regime="tiefree", rate_het=False, gap_cols=0.03, gap_runs=True, tree=None):
    """uint8 codes [n, L] for tips in tree-leaf order shuffled by the seed, plus the tree.

    Returns (codes, info) with info = dict(parent, bl, children, leaves, order) where
    order[i] = tree node of output row i.
    """
    rng = np.random.default_rng(seed)
    parent, bl, children, leaves = tree if tree is not None else yule_tree(n, rng, regime)
    if rate_het:
        cat = np.array([0.0334, 0.2519, 0.8203, 2.8944])  # Gamma(alpha=0.5) 4-category means
        rates = cat[rng.integers(0, 4, L)]
        rates[rng.random(L) < 0.2] = 0.0
        rates = rates / max(rates.mean(), 1e-12)
        cdf = np.cumsum(rates) / rates.sum()
    order = np.array(leaves)
    rng.shuffle(order)
    row_of = {int(v): i for i, v in enumerate(order)}
    codes = np.empty((n, L), np.uint8)
    root_seq = rng.integers(0, 4, L).astype(np.uint8)
    stack = [(0, root_seq)]
    while stack:
        v, seq = stack.pop()
        if not children[v]:
            codes[row_of[v]] = seq
            continue
        for c in children[v]:
            s = seq.copy()
            ev = rng.poisson(L * bl[c] * 4.0 / 3.0)
            if ev:
                if rate_het:
                    pos = np.searchsorted(cdf, rng.random(ev))
                else:
                    pos = rng.integers(0, L, ev)
                s[pos] = rng.integers(0, 4, ev).astype(np.uint8)
            stack.append((c, s))
    if gap_cols:
        ncol = int(L * gap_cols)
        if ncol:
            cols = rng.choice(L, ncol, replace=False)
            for c in cols:
                codes[rng.random(n) < 0.5, c] = 4
    if gap_runs:
        for i in range(n):
            for _ in range(int(rng.integers(0, 3))):
                st = int(rng.integers(0, L))
                ln = int(rng.integers(1, 51))
                codes[i, st:st + ln] = 4
    return codes, dict(parent=parent, bl=bl, children=children, leaves=leaves, order=order)


_LET = np.frombuffer(b"ACGT-", np.uint8)


def codes_to_strings(codes):
    return [_LET[row].tobytes().decode() for row in codes]


def names(n):
    return ["T%d" % (i + 1) for i in range(n)]


def pack4_np(codes):
    """[n, L] codes -> uint64 [n, ceil(L/16)], bit-identical to fourBitCompressor."""
    n, L = codes.shape
    W = (L + 15) // 16
    out = np.zeros((n, W), np.uint64)
    sh = (np.arange(16, dtype=np.uint64) * np.uint64(4))
    step = max(1, (1 << 24) // max(W * 16, 1))
    for r0 in range(0, n, step):
        blk = codes[r0:r0 + step]
        pad = np.zeros((blk.shape[0], W * 16), np.uint64)
        pad[:, :L] = blk
        out[r0:r0 + step] = (pad.reshape(blk.shape[0], W, 16) << sh).sum(axis=2, dtype=np.uint64)
    return out


def pack2_np(seq_codes):
    """1-D codes (0..3; anything else -> 0) -> uint64 [ceil(len/32)], as twoBitCompressor."""
    L = len(seq_codes)
    W = (L + 31) // 32
    pad = np.zeros(W * 32, np.uint64)
    c = np.asarray(seq_codes, np.uint64)
    pad[:L] = np.where(c < 4, c, 0)
    sh = (np.arange(32, dtype=np.uint64) * np.uint64(2))
    return (pad.reshape(W, 32) << sh).sum(axis=1, dtype=np.uint64)


def unaligned(codes):
    """Strip gap codes: list of 1-D code arrays (ragged)."""
    return [row[row < 4] for row in codes]


def flatten2(seqs2):
    """Ragged 2-bit packed sequences -> (flat words, word offsets, lengths in bases)."""
    packed = [pack2_np(s) for s in seqs2]
    lens = np.array([len(s) for s in seqs2], np.uint64)
    offs = np.zeros(len(seqs2), np.uint64)
    if len(packed) > 1:
        offs[1:] = np.cumsum([len(p) for p in packed[:-1]])
    flat = np.concatenate(packed) if packed else np.zeros(0, np.uint64)
    return flat, offs, lens


def write_fasta(path, names_, seqs):
    with open(path, "w") as f:
        for nm, s in zip(names_, seqs):
            f.write(">%s\n%s\n" % (nm, s))


def write_phylip(path, names_, D, lower=True):
    """`%.6f` PHYLIP text; lower-triangular rows carry j < i entries (src/matrix_reader.cu:23-42)."""
    n = len(names_)
    with open(path, "w") as f:
        f.write("%d\n" % n)
        for i in range(n):
            vals = D[i, :i] if lower else D[i]
            f.write(names_[i] + (" " if len(vals) else "") + " ".join("%.6f" % v for v in vals) + "\n")


def tree_with_queries(newick_text, n_queries, seed=1, scale=1.0, query_bl=None):
    """Yule-style extension of a given rooted tree (the reference's only fixture, dataset/t2.backbone.nwk, in
    BASELINE config 4b): every query tip is attached in the middle of a random existing edge.  Returns the
    `tree` tuple evolve() takes, plus (backbone_leaf_nodes, backbone_leaf_names, query_nodes); branch lengths
    are multiplied by `scale` (the fixture's alisim-regime lengths give almost identical sequences at a few
    thousand sites)."""
    from . import newick as _nw
    children, length, name = _nw.parse(newick_text)
    children = [list(c) for c in children]
    parent = [-1] * len(children)
    for v, ch in enumerate(children):
        for c in ch:
            parent[c] = v
    bl = [float(x) * scale for x in length]
    bb_leaves = [v for v in range(len(children)) if not children[v]]
    bb_names = [name[v] for v in bb_leaves]
    rng = np.random.default_rng(seed)
    mean = float(np.mean([bl[v] for v in range(1, len(bl))])) if query_bl is None else query_bl
    qnodes = []
    for _ in range(n_queries):
        v = int(rng.integers(1, len(children)))          # edge above v (never the root)
        p = parent[v]
        m, q = len(children), len(children) + 1         # new internal node, new tip
        children += [[v, q], []]
        parent += [p, m]
        bl += [bl[v] * 0.5, max(float(rng.exponential(mean)), 1e-6)]
        bl[v] *= 0.5
        children[p][children[p].index(v)] = m
        parent[v] = m
        qnodes.append(q)
    leaves = bb_leaves + qnodes
    return (np.array(parent), np.array(bl), children, leaves), bb_leaves, bb_names, qnodes


_POOL_TREE = None   # (bl, children, L) of evolve_parallel_packed, inherited by the forked workers (not pickled per job)


def _evolve_subtree(args):
    """Worker of evolve_parallel: evolves one subtree from its root sequence; returns (leaf node ids, codes)."""
    root, seq, seed = args
    bl, children, L = _POOL_TREE
    rng = np.random.default_rng(seed)
    leaves, rows = [], []
    stack = [(root, seq)]
    while stack:
        v, s = stack.pop()
        if not children[v]:
            leaves.append(v)
            rows.append(s)
            continue
        for c in children[v]:
            t = s.copy()
            ev = rng.poisson(L * bl[c] * 4.0 / 3.0)
            if ev:
                t[rng.integers(0, L, ev)] = rng.integers(0, 4, ev).astype(np.uint8)
            stack.append((c, t))
    return leaves, pack4_np(np.stack(rows)) if rows else np.zeros((0, (L + 15) // 16), np.uint64)


def evolve_parallel_packed(n, L, seed=1, regime="tiefree", workers=0, frontier=4096):
    """Same model as evolve() (Yule tree, JC substitutions, no gaps) for the multi-million-tip configurations, with the
    subtrees below a `frontier`-node cut evolved and 4-bit packed by a process pool (the single-threaded generator needs
    3 minutes for 2 000 000 tips x 10 000 sites).  Returns the packed rows [n, ceil(L/16)] in a seeded random order."""
    import multiprocessing as mp
    rng = np.random.default_rng(seed)
    parent, bl, children, leaves = yule_tree(n, rng, regime)
    # breadth-first down to `frontier` open nodes, sequences evolved serially on the way
    seqs = {0: rng.integers(0, 4, L).astype(np.uint8)}
    open_nodes, done_leaves = [0], []
    while open_nodes and len(open_nodes) < frontier:
        v = open_nodes.pop(0)
        if not children[v]:
            done_leaves.append(v)
            continue
        for c in children[v]:
            t = seqs[v].copy()
            ev = rng.poisson(L * bl[c] * 4.0 / 3.0)
            if ev:
                t[rng.integers(0, L, ev)] = rng.integers(0, 4, ev).astype(np.uint8)
            seqs[c] = t
            open_nodes.append(c)
        del seqs[v]
    # largest subtrees first (Yule subtrees are very unequal)
    size = np.zeros(len(parent), np.int64)
    size[np.array(leaves)] = 1
    for v in range(len(parent) - 1, 0, -1):
        size[parent[v]] += size[v]
    open_nodes.sort(key=lambda v: -int(size[v]))
    global _POOL_TREE
    _POOL_TREE = (bl, children, L)
    jobs = [(v, seqs[v], seed * 1000003 + k) for k, v in enumerate(open_nodes)]
    workers = workers or min(len(jobs), os.cpu_count() or 1)
    W = (L + 15) // 16
    out = np.zeros((n, W), np.uint64)
    order = np.array(leaves)
    rng.shuffle(order)
    row_of = np.empty(len(parent), np.int64)
    row_of[order] = np.arange(n)
    with mp.get_context("fork").Pool(workers) as pool:
        for lv, packed in pool.imap_unordered(_evolve_subtree, jobs, chunksize=1):
            if len(lv):
                out[row_of[np.array(lv)]] = packed
    for v in done_leaves:
        out[row_of[v]] = pack4_np(seqs[v][None, :])[0]
    return out
