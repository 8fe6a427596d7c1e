"""Pins the oracle (oracle/p3_oracle.c) against the UNMODIFIED reference compiled into
oracle/_ref/libp3ref.so. CPU only. Skipped where the reference build is absent (it travels to
the GPU box as a prebuilt file, so these also run there)."""
import os

import numpy as np
import pytest

from _checkers import Oracle, Ref, have_ref, kmer_str_to_words, nwords, reads_to_arrays, words_to_kmer_str
from platanus3_b200 import synth

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libp3ref.so not built")

KS = [21, 22, 25, 27, 31, 32, 33, 47, 63, 64, 65, 101]


def _rand_kmer(rng, k):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, size=k))


@pytest.mark.parametrize("k", KS + [501, 1001, 3001])
def test_std_hash_and_double_hash(oracle, k):
    """std::hash<bitset<2k>> (libstdc++ _Hash_bytes) + MyHash.cpp:22 GetDoubleHash_64bit"""
    rng = np.random.default_rng(k)
    ref = Ref(k)
    kmers = [_rand_kmer(rng, k) for _ in range(50)] + ["A" * k, "T" * k, "C" * k, "ACGT" * (k // 4) + "A" * (k % 4)]
    for s in kmers:
        w = kmer_str_to_words(s, k)
        h0 = oracle.std_hash_kmer(w, k)
        assert h0 == ref.std_hash(s)
        assert oracle.double_hash(h0) == ref.double_hash(s)


@pytest.mark.parametrize("k", KS)
def test_canonical(oracle, k):
    """BitCalc.cpp GetFirstKmerForward/Backward + CompareBit, incl. non-ACGT characters"""
    rng = np.random.default_rng(100 + k)
    ref = Ref(k)
    for i in range(60):
        s = _rand_kmer(rng, k)
        if i % 3 == 0:  # reference quirk: non-ACGT reads as code 0 on BOTH strands
            p = int(rng.integers(0, k))
            s = s[:p] + "Nnacgt"[i % 6] + s[p + 1:]
        assert words_to_kmer_str(oracle.canonical_words(s, k), k) == ref.canonical(s)


def test_estimate_bloomfilter(oracle):
    """Options.cpp:50-60"""
    ref = Ref(21)
    rng = np.random.default_rng(7)
    cases = [10 ** e for e in range(3, 13)] + [int(x) for x in rng.integers(2000, 2 ** 40, size=200)]
    for ab in cases:
        for k in (21, 25, 32, 63, 101, 3001):
            assert oracle.estimate_bloomfilter(ab, k) == ref.estimate_only(ab, k)


def _dataset(seed, genome=3000, cov=12, rl=80, err=0.01):
    g = synth.random_genome(genome, seed)
    return synth.reads_as_bytes(synth.simulate_reads(g, cov, rl, err, seed + 1))


def _run_ref(reads, k, m=0):
    ref = Ref(k, m=m)
    ref.add_reads(reads)
    if m == 0:
        ref.estimate()
    keys, counts = ref.count_short()
    bits, seeds = ref.make_bf()
    return ref, keys, counts, bits, seeds


@pytest.mark.parametrize("k,rl", [(21, 50), (25, 60), (32, 80), (31, 80), (33, 80), (63, 150), (64, 150), (101, 300)])
def test_count_and_makebf(oracle, k, rl):
    """Load.cpp:105 CountShortKmer + MakeBloomFilter.cpp:25 MakeBF (bits, seeds)"""
    reads = _dataset(k, genome=2500, cov=10, rl=rl, err=0.01)
    ref, rkeys, rcounts, rbits, rseeds = _run_ref(reads, k)
    seq, off = reads_to_arrays(reads)
    keys, counts = oracle.count_short_kmers(seq, off)
    assert np.array_equal(keys, rkeys) and np.array_equal(counts, rcounts)
    fs, nh = oracle.estimate_bloomfilter(int(off[-1]), k)
    assert (fs, nh) == (ref.filter_size, ref.num_hashes)
    bloom, seed_pos, solid, adds = oracle.make_bf(seq, off, k, keys, counts, fs, nh, want_solid=True)
    assert np.array_equal(bloom, rbits)
    seeds = sorted({reads[r][p:p + k].decode() for r, p in enumerate(seed_pos) if p >= 0})
    assert seeds == rseeds
    assert adds == int(solid.sum()) > 0


def test_makebf_with_m_option(oracle):
    """-m given: filter_size = m, num_hashes stays 10 (Options.cpp:10-11,51)"""
    k = 25
    reads = _dataset(5, genome=2000, cov=8, rl=70, err=0.02)
    ref, rkeys, rcounts, rbits, rseeds = _run_ref(reads, k, m=100003)
    assert ref.num_hashes == 10 and ref.filter_size == 100003
    seq, off = reads_to_arrays(reads)
    keys, counts = oracle.count_short_kmers(seq, off)
    bloom, _, _, _ = oracle.make_bf(seq, off, k, keys, counts, 100003, 10)
    assert np.array_equal(bloom, rbits)


def test_non_acgt_and_ragged_reads(oracle):
    """reads of different lengths, one exactly k long, N / lower-case characters"""
    k = 25
    rng = np.random.default_rng(3)
    g = synth.random_genome(1500, 11)
    reads = []
    for i in range(300):
        L = int(rng.integers(k, 120)) if i else k
        s = int(rng.integers(0, len(g) - L))
        r = bytearray(synth.codes_to_ascii(g[s:s + L]).tobytes())
        if i % 7 == 0:
            r[int(rng.integers(0, L))] = ord("N")
        if i % 11 == 0:
            r[int(rng.integers(0, L))] = ord("a")
        reads.append(bytes(r))
    ref, rkeys, rcounts, rbits, rseeds = _run_ref(reads, k, m=50021)
    seq, off = reads_to_arrays(reads)
    keys, counts = oracle.count_short_kmers(seq, off)
    assert np.array_equal(keys, rkeys) and np.array_equal(counts, rcounts)
    bloom, seed_pos, _, _ = oracle.make_bf(seq, off, k, keys, counts, 50021, 10)
    assert np.array_equal(bloom, rbits)
    seeds = sorted({reads[r][p:p + k].decode().translate(str.maketrans("Na", "AA")) for r, p in enumerate(seed_pos) if p >= 0})
    assert seeds == rseeds


@pytest.mark.parametrize("k", [21, 25, 32, 33, 63, 101])
def test_check_directions(oracle, k):
    """DeBruijnGraph.cpp:326 CheckDirections / :318 IsRecorded on oriented k-mers"""
    reads = _dataset(40 + k, genome=2000, cov=10, rl=max(60, k + 40), err=0.005)
    ref, rkeys, rcounts, rbits, rseeds = _run_ref(reads, k)
    fs, nh = ref.filter_size, ref.num_hashes
    rng = np.random.default_rng(k)
    probes = []
    for r in reads[:40]:
        for p in range(0, len(r) - k + 1, 7):
            probes.append(r[p:p + k].decode())
    probes += [_rand_kmer(rng, k) for _ in range(50)]
    n_edges = 0
    for s in probes:
        w = kmer_str_to_words(s, k)
        ign = int(rng.integers(-1, 8))
        got = oracle.check_directions(rbits, fs, nh, w, k, ign)
        assert got == ref.check_directions(s, ign)
        assert oracle.is_recorded(rbits, fs, nh, w, k) == ref.is_recorded(s)
        n_edges += bin(got).count("1")
    assert n_edges > 0


def test_load_fasta_fastq(oracle, tmp_path):
    """Load.cpp:32-103: multi-line FASTA, single-line FASTQ, short reads dropped, duplicate
    names collapse to the last record while all_bases counts both"""
    k = 25
    reads = _dataset(9, genome=1200, cov=6, rl=60, err=0.0)[:40]
    reads[3] = reads[3][:10]  # shorter than k -> dropped
    names = [">read_%d some comment" % i for i in range(len(reads))]
    names[7] = names[2]  # duplicate name line
    fa = str(tmp_path / "in.fasta")
    synth.write_fasta(fa, reads, width=17, names=names)
    fq = str(tmp_path / "in.fastq")
    synth.write_fastq(fq, reads, names=["@" + n[1:] for n in names])
    for path in (fa, fq):
        ref = Ref(k, readfile=path)
        ref.load_file()
        seq, off, all_bases = oracle.load_reads(path, k)
        assert all_bases == ref.all_bases == sum(len(r) for r in reads if len(r) >= k)
        mine = sorted(seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1))
        assert mine == sorted(ref.reads())
        assert len(mine) == len(reads) - 2


def test_rmq_matches_window_min(oracle):
    """MakeBloomFilter.cpp:8-22 is a sliding-window minimum for x >= 1"""
    rng = np.random.default_rng(5)
    for x in (1, 2, 5, 12, 43):
        v = rng.integers(0, 9, size=200).astype(np.uint64)
        got = oracle.rmq(v, x)
        want = np.array([v[j:j + x].min() for j in range(len(v) - x + 1)], np.uint64)
        assert np.array_equal(got, want)
