"""Writes the small committed fixtures under tests/golden/.

The reference is CUDA-only and this container has no GPU, so fixtures come from two
sources: (a) this script, which records the CPU oracle's outputs on seeded inputs
(regression pins: files gold_*.npz), and (b) tools/make_ref_golden.py, which runs the
reference's own CUDA objects (oracle/_ref/dipper_ref) on the GPU box and stores their
outputs (files ref_*.npz).  The oracle is checked against both in tests/test_oracle.py.
Run from the repo root: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dipper_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def msa_case(name, n, L, seed, regime, types=(1, 2, 3, 4, 5, 6), nj=True):
    codes, _ = synth.evolve(n, L, seed=seed, regime=regime, gap_cols=0.05)
    P = synth.pack4_np(codes)
    m, u = O.msa_counts(P, L, 0, n, 0, n)
    out = dict(kind="msa", packed=P, seq_len=L, match=m, useful=u, dist_types=np.array(types))
    for t in types:
        out["dist_%d" % t] = O.msa_dist_matrix(P, L, t)
    if nj:
        c0, c1, l0, l1 = O.nj(out["dist_2"])
        out.update(nj_child0=c0, nj_child1=c1, nj_len0=l0, nj_len1=l1)
    np.savez_compressed(os.path.join(HERE, name), **out)


def mash_case(name, n, L, seed, k=15, s=1000):
    codes, _ = synth.evolve(n, L, seed=seed, regime="tiefree", gap_cols=0.0)
    flat, offs, lens = synth.flatten2(synth.unaligned(codes))
    sk = O.sketch_all(flat, offs, lens, k, s)
    D = O.mash_dist_matrix(sk, k)
    np.savez_compressed(os.path.join(HERE, name), kind="mash", flat=flat, offsets=offs, lens=lens, k=k, s=s,
                        sketches=sk, dist=D)


if __name__ == "__main__":
    msa_case("gold_msa_tiefree_24x700.npz", 24, 700, 1, "tiefree")
    msa_case("gold_msa_alisim_20x2000.npz", 20, 2000, 2, "alisim", types=(1, 2), nj=True)
    mash_case("gold_mash_6x2500.npz", 6, 2500, 3)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
