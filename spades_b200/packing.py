"""2-bit packing of reads in the reference's layout.

Nucleotide i of a sequence sits at bits 2(i%32)..2(i%32)+1 of word i//32, A=0 C=1 G=2 T=3
(reference: src/common/sequence/rtseq.hpp:379-382, src/common/sequence/nucl.hpp:132-142,
`Sequence` uses the same packing, src/common/sequence/sequence.hpp:147-200). Every read starts on
a u64 word boundary, like the records of the reference's binary read files
(src/common/io/reads/single_read.hpp:317-323: u64 size + ceil(size/32) words).
"""
import re
import numpy as np

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[_c + 32] = _i  # lower case


def longest_valid(seq: str) -> str:
    """First longest ACGT run of a read (io/reads/longest_valid_wrapper.hpp:16-53)."""
    runs = re.findall("[ACGTacgt]+", seq)
    if not runs:
        return ""
    return max(runs, key=len)  # max() keeps the first maximal element


def pack_reads(reads):
    """list[str|bytes] (ACGT only) -> (words u64[], offs u64[n] (word offsets), lens u32[n])."""
    n = len(reads)
    lens = np.fromiter((len(r) for r in reads), dtype=np.uint32, count=n)
    nw = (lens.astype(np.int64) + 31) // 32
    offs = np.zeros(n, dtype=np.uint64)
    if n:
        offs[1:] = np.cumsum(nw)[:-1].astype(np.uint64)
    total = int(nw.sum())
    words = np.zeros(total, dtype=np.uint64)
    if n == 0:
        return words, offs, lens
    # flat base array, padded per read to a multiple of 32
    padded = np.zeros(total * 32, dtype=np.uint8)
    flat = np.frombuffer(("".join(r if isinstance(r, str) else r.decode() for r in reads)).encode(), dtype=np.uint8)
    codes = _CODE[flat]
    if codes.size and codes.max() > 3:
        raise ValueError("pack_reads: non-ACGT symbol (apply longest_valid first)")
    starts = offs.astype(np.int64) * 32
    pos = np.repeat(starts - np.concatenate(([0], np.cumsum(lens.astype(np.int64))[:-1])), lens.astype(np.int64)) + np.arange(codes.size)
    padded[pos] = codes
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))
    words = (padded.reshape(-1, 32).astype(np.uint64) << shifts).sum(axis=1, dtype=np.uint64)
    return words, offs, lens


def pack_fixed(codes2d: np.ndarray):
    """uint8 [n, L] base codes -> (words, offs, lens) with a fixed stride of ceil(L/32) words."""
    n, L = codes2d.shape
    nw = (L + 31) // 32
    padded = np.zeros((n, nw * 32), dtype=np.uint8)
    padded[:, :L] = codes2d
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))
    words = (padded.reshape(n * nw, 32).astype(np.uint64) << shifts).sum(axis=1, dtype=np.uint64)
    offs = (np.arange(n, dtype=np.uint64) * np.uint64(nw))
    lens = np.full(n, L, dtype=np.uint32)
    return words, offs, lens


def unpack_reads(words, offs, lens):
    """inverse of pack_reads: the reads of a packed read set as ACGT strings"""
    out = []
    for o, L in zip(offs, lens):
        L = int(L)
        w = np.asarray(words[int(o): int(o) + (L + 31) // 32], dtype=np.uint64)
        codes = ((w[:, None] >> (2 * np.arange(32, dtype=np.uint64))[None, :]) & np.uint64(3)).reshape(-1)[:L]
        out.append("".join("ACGT"[int(c)] for c in codes))
    return out


def unpack_kmers(keys: np.ndarray, K: int):
    """u64 [n, nw] -> list[str]."""
    out = []
    for row in keys.reshape(-1, (K + 31) // 32):
        s = []
        for i in range(K):
            s.append("ACGT"[(int(row[i // 32]) >> (2 * (i % 32))) & 3])
        out.append("".join(s))
    return out


def revcomp(s: str) -> str:
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def synthetic_reads(n_reads, read_len=150, genome_len=None, err=0.01, seed=42, as_codes=False):
    """SURVEY 8(d) generator: uniform genome, uniform start, random strand, 1% substitutions."""
    rng = np.random.default_rng(seed)
    if genome_len is None:
        genome_len = max(read_len + 1, n_reads)  # ~150x like config 3 (100 M reads / 100 Mbp)
    genome = rng.integers(0, 4, size=genome_len, dtype=np.uint8)
    starts = rng.integers(0, genome_len - read_len + 1, size=n_reads)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    reads = genome[idx]
    strand = rng.random(n_reads) < 0.5
    reads[strand] = (3 - reads[strand])[:, ::-1]
    errs = rng.random(reads.shape) < err
    reads[errs] = (reads[errs] + rng.integers(1, 4, size=int(errs.sum()), dtype=np.uint8)) & 3
    if as_codes:
        return reads
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    return [lut[r].tobytes().decode() for r in reads]
