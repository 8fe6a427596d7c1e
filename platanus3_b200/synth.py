"""Synthetic genomes and reads for tests and bench (numpy, host side).

BASELINE.json's configs are all "synthetic random genome + simulated reads"; this is the one
generator every test and bench leg shares so that the reference arm, the oracle and the CUDA
path see byte-identical input.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_genome(n_bases, seed=0):
    """uint8 codes 0..3 (A,C,G,T)"""
    return np.random.default_rng(seed).integers(0, 4, size=n_bases, dtype=np.uint8)


def simulate_reads(genome, coverage, read_len, error_rate=0.0, seed=1, both_strands=True):
    """Uniform shotgun reads -> (codes uint8 [n_reads, read_len]).

    Substitution errors replace a base by one of the three others (never a no-op), reads come
    from either strand when both_strands is set.
    """
    rng = np.random.default_rng(seed)
    g = len(genome)
    n_reads = max(1, int(round(g * coverage / read_len)))
    starts = rng.integers(0, g - read_len + 1, size=n_reads)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    codes = genome[idx]
    if both_strands:
        flip = rng.random(n_reads) < 0.5
        codes[flip] = (3 - codes[flip])[:, ::-1]
    if error_rate > 0:
        err = rng.random(codes.shape) < error_rate
        shift = rng.integers(1, 4, size=codes.shape, dtype=np.uint8)
        codes = np.where(err, (codes + shift) & 3, codes).astype(np.uint8)
    return np.ascontiguousarray(codes)


def codes_to_ascii(codes):
    return _ACGT[codes]


def reads_as_bytes(codes):
    """[n_reads, read_len] codes -> list[bytes] of ACGT strings"""
    asc = codes_to_ascii(codes)
    return [row.tobytes() for row in asc]


def write_fasta(path, reads, width=0, names=None):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write((names[i] if names else ">read_%d" % i).encode() + b"\n")
            if width and width > 0:
                for o in range(0, len(r), width):
                    f.write(r[o:o + width] + b"\n")
            else:
                f.write(r + b"\n")


def write_fastq(path, reads, names=None):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write((names[i] if names else "@read_%d" % i).encode() + b"\n" + r + b"\n+\n" + b"I" * len(r) + b"\n")
