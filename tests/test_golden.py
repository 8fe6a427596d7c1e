"""Golden fixtures (tests/golden/, generated from the compiled reference by make_golden.py):
the oracle must reproduce them on CPU, the CUDA path on the GPU. /root/reference is not read."""
import glob
import os

import numpy as np
import pytest

from _checkers import kmer_str_to_words, reads_to_arrays
from platanus3_b200 import _lib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))


def _load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    return g, os.path.join(GOLD, str(g["read_file"]))


def test_fixtures_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(oracle, name):
    g, path = _load(name)
    k, m = int(g["k"]), int(g["m"])
    seq, off, all_bases = oracle.load_reads(path, k)
    assert all_bases == int(g["all_bases"]) and len(off) - 1 == int(g["n_reads"])
    fs, nh = (m, 10) if m else oracle.estimate_bloomfilter(all_bases, k)
    assert (fs, nh) == (int(g["filter_size"]), int(g["num_hashes"]))
    keys, counts = oracle.count_short_kmers(seq, off)
    assert np.array_equal(keys, g["keys"]) and np.array_equal(counts, g["counts"])
    bloom, seed_pos, _, _ = oracle.make_bf(seq, off, k, keys, counts, fs, nh)
    assert np.array_equal(bloom, g["bloom"])
    tr = bytes.maketrans(bytes(set(range(256)) - set(b"ACGT")), b"A" * 252)
    seeds = sorted({seq[int(off[r]) + p:int(off[r]) + p + k].tobytes().translate(tr).decode()
                    for r, p in enumerate(seed_pos) if p >= 0})
    assert seeds == [str(s) for s in g["seeds"]]
    for s, want in zip(g["probe_kmers"], g["probe_masks"]):
        assert oracle.check_directions(bloom, fs, nh, kmer_str_to_words(str(s), k), k) == int(want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_golden(oracle, name):
    g, path = _load(name)
    k, m = int(g["k"]), int(g["m"])
    seq, off, all_bases = oracle.load_reads(path, k)  # loader parity is covered on CPU; here it only feeds bytes
    fs, nh = int(g["filter_size"]), int(g["num_hashes"])
    if not m:
        assert _lib.estimate_bloomfilter(all_bases, k) == (fs, nh)
    with _lib.Context(0) as ctx:
        ctx.load_ascii(seq, off)
        ctx.count_short_kmers()
        keys, counts = ctx.short_kmer_export()
        assert np.array_equal(keys, g["keys"]) and np.array_equal(counts, g["counts"])
        ctx.make_bf(k, fs, nh)
        assert np.array_equal(ctx.bf_export(), g["bloom"])
        seed_pos = ctx.seed_export()
        tr = bytes.maketrans(bytes(set(range(256)) - set(b"ACGT")), b"A" * 252)
        seeds = sorted({seq[int(off[r]) + p:int(off[r]) + p + k].tobytes().translate(tr).decode()
                        for r, p in enumerate(seed_pos) if p >= 0})
        assert seeds == [str(s) for s in g["seeds"]]
        probes = np.stack([kmer_str_to_words(str(s), k) for s in g["probe_kmers"]])
        if probes.shape[1] == 1:
            probes = probes[:, 0]
        assert np.array_equal(ctx.check_directions(probes), g["probe_masks"])
