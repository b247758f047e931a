"""SHA-256 fixture of the UNMODIFIED reference on the >= 1 M-read synthetic set BASELINE.md 3.6 asks for.

    make -C oracle ref && python tests/golden/make_golden_1m.py        (build container only: needs oracle/_ref/ref_probe)

Reads: spades_b200.packing.synthetic_reads(1_000_000, 150, genome_len=1_000_000, err=0.01, seed=42) -- the SURVEY 8(d) generator
(numpy default_rng(42): uniform genome, uniform start, random strand, 1 % substitutions), i.e. BASELINE config 3 at 1/100 scale
(150x coverage). `ref_probe graph` = reads -> (k+1)-mers -> k-mers -> KMerIndex -> masks -> unitigs -> coverage -> GFA with the
reference's own classes; every artefact it dumps is hashed. tests/test_gpu_parity.py::test_million_reads_sha256 recomputes the same
bytes through the C ABI on the B200 and compares the digests (the artefacts themselves are 0.1-1.5 GB and stay out of the repo).
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from spades_b200.packing import synthetic_reads  # noqa: E402

PROBE = os.path.join(ROOT, "oracle", "_ref", "ref_probe")
N_READS, READ_LEN, GENOME, ERR, SEED = 1_000_000, 150, 1_000_000, 0.01, 42
CASES = [(21, 16), (55, 80)]          # (k, num_buckets): spades-kmercount's 16 and the graph path's 10 x 8 threads
FILES = ["kpomers", "kpomer_bucket_sizes.txt", "kmers", "kmer_index.bin", "kpomer_index.bin", "masks.bin", "coverage.bin", "histogram.txt",
         "unitigs.txt", "graph.gfa"]


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    codes = synthetic_reads(N_READS, READ_LEN, GENOME, ERR, seed=SEED, as_codes=True)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    txt = np.empty((N_READS, READ_LEN + 1), np.uint8)
    txt[:, :READ_LEN] = lut[codes]
    txt[:, READ_LEN] = 10
    out = {"reads": {"n": N_READS, "len": READ_LEN, "genome_len": GENOME, "err": ERR, "seed": SEED,
                     "sha256_text": hashlib.sha256(txt.tobytes()).hexdigest()}, "cases": {}}
    with tempfile.TemporaryDirectory() as d:
        rf = os.path.join(d, "reads.txt")
        txt.tofile(rf)
        for k, B in CASES:
            od = os.path.join(d, "out_k%d" % k)
            t0 = time.time()
            subprocess.check_call([PROBE, "graph", rf, str(k), str(B), "8", od], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            c = {"k": k, "B": B, "reference_wall_s": round(time.time() - t0, 1), "sha256": {}, "bytes": {}}
            for f in FILES:
                p = os.path.join(od, f)
                c["sha256"][f] = sha(p); c["bytes"][f] = os.path.getsize(p)
            out["cases"]["k%d" % k] = c
            print(k, B, c["reference_wall_s"], "s", c["bytes"], flush=True)
    json.dump(out, open(os.path.join(HERE, "syn1M_sha256.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
