"""Throughput of the parallel FASTA ingest + packing (dipb_fasta_open) on the host cores."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api

n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 30000, int(sys.argv[2]) if len(sys.argv) > 2 else 10000
path = "/tmp/dipb_fasta_bench_%d_%d.fa" % (n, L)
if not os.path.exists(path):
    rng = np.random.default_rng(1)
    with open(path, "wb") as f:
        lut = np.frombuffer(b"ACGT", np.uint8)
        for i in range(n):
            seq = lut[rng.integers(0, 4, L)]
            rows = np.concatenate([seq.reshape(-1, 100), np.full((L // 100, 1), 10, np.uint8)], axis=1)   # 100 columns per line
            f.write(b">T%d\n" % (i + 1))
            f.write(rows.tobytes())
size = os.path.getsize(path)
out = {"file_bytes": size, "records": n, "sites": L, "host_threads": os.cpu_count()}
for bits in (4, 2):
    for thr in (1, 0):
        best = 1e9
        for rep in range(3):
            t0 = time.time()
            names, lens, off, words = api.read_fasta_packed(path, bits, thr)
            best = min(best, time.time() - t0)
        out["bits%d_threads%s_GBps" % (bits, "all" if thr == 0 else "1")] = size / best / 1e9
print(json.dumps(out))
