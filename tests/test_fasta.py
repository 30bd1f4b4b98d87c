"""CPU tests of the parallel FASTA ingest + packing (dipb_fasta_open) against the per-sequence packers
(dipb_pack4 / dipb_pack2, themselves checked against the oracle in test_abi.py)."""
import gzip

import numpy as np
import pytest

from dipper_b200 import api


def _write(path, recs, width=60, crlf=False, junk_before=True):
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        if junk_before:
            f.write("; a comment line before the first record" + nl)
        for k, (name, seq) in enumerate(recs):
            f.write(">" + name + (" some description\tmore" if k % 2 else "") + nl)
            w = width + (k % 7)
            for i in range(0, len(seq), w):
                f.write(seq[i:i + w] + ("  " if k % 3 == 0 else "") + nl)
            if k % 5 == 0:
                f.write(nl)          # blank line inside the file


def _recs(n, seed, aligned):
    rng = np.random.default_rng(seed)
    alphabet = np.array(list("ACGTUNacgt-RYKM"))
    p = np.array([.2, .2, .2, .2, .02, .04, .02, .02, .02, .02, .03, .01, .01, .005, .005])
    p = p / p.sum()
    recs = []
    L = 1000
    for i in range(n):
        ln = L if aligned else int(rng.integers(1, 3000))
        recs.append(("T%d" % (i + 1), "".join(rng.choice(alphabet, ln, p=p))))
    return recs


@pytest.mark.parametrize("bits,aligned", [(4, True), (2, False)])
@pytest.mark.parametrize("crlf", [False, True])
@pytest.mark.parametrize("threads", [1, 0])
def test_fasta_matches_per_sequence_packers(tmp_path, bits, aligned, crlf, threads):
    recs = _recs(257, 11 + bits, aligned)
    path = str(tmp_path / "x.fa")
    _write(path, recs, crlf=crlf)
    names, lens, off, words = api.read_fasta_packed(path, bits, threads)
    assert names == [r[0] for r in recs]
    assert lens.tolist() == [len(r[1]) for r in recs]
    pack = api.pack4 if bits == 4 else api.pack2
    for i, (_, seq) in enumerate(recs):
        want = pack(seq)
        assert np.array_equal(words[int(off[i]):int(off[i + 1])], want), i


def test_fasta_large_file_uses_all_threads(tmp_path):
    # > 1 MiB so that the chunked record search runs with several threads; records straddle chunk borders
    recs = _recs(3000, 5, True)
    path = str(tmp_path / "big.fa")
    _write(path, recs, junk_before=False)
    names, lens, off, words = api.read_fasta_packed(path, 4, 0)
    assert names == [r[0] for r in recs] and int(off[-1]) == 3000 * 63
    for i in (0, 1, 1499, 2999):
        assert np.array_equal(words[int(off[i]):int(off[i + 1])], api.pack4(recs[i][1]))


def test_fasta_empty_and_gzip(tmp_path):
    p = str(tmp_path / "empty.fa")
    open(p, "w").close()
    names, lens, off, words = api.read_fasta_packed(p, 4)
    assert names == [] and off.tolist() == [0]
    g = str(tmp_path / "x.fa.gz")
    with gzip.open(g, "wt") as f:
        f.write(">a\nACGT\n")
    with pytest.raises(api.DipperError):
        api.read_fasta_packed(g, 4)
