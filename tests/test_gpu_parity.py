"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle
(bit-exact: integer/byte work) and against the golden fixtures from the reference build."""
import numpy as np
import pytest

from _checkers import Oracle, kmer_str_to_words, reads_to_arrays
from platanus3_b200 import _lib, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def _dataset(seed, genome, cov, rl, err):
    g = synth.random_genome(genome, seed)
    return synth.reads_as_bytes(synth.simulate_reads(g, cov, rl, err, seed + 1))


def _full_check(ctx, oracle, reads, k, m=0, n_adj=3000, table_slots=0, solid_slots=0):
    seq, off = reads_to_arrays(reads)
    if m:
        fs, nh = m, 10
    else:
        fs, nh = _lib.estimate_bloomfilter(int(off[-1]), k)
        assert (fs, nh) == oracle.estimate_bloomfilter(int(off[-1]), k)
    ctx.load_ascii(seq, off)
    n_pos, n_distinct = ctx.count_short_kmers(table_slots)
    keys, counts = ctx.short_kmer_export()
    okeys, ocounts = oracle.count_short_kmers(seq, off)
    assert n_pos == int(ocounts.sum()) and n_distinct == len(okeys)
    assert np.array_equal(keys, okeys)
    assert np.array_equal(counts, ocounts)

    n_adds, n_solid = ctx.make_bf(k, fs, nh, solid_slots=solid_slots)
    obits, oseeds, osolid_flags, oadds = oracle.make_bf(seq, off, k, okeys, ocounts, fs, nh, want_solid=True)
    assert n_adds == oadds
    assert np.array_equal(ctx.bf_export(), obits)
    assert np.array_equal(ctx.seed_export(), oseeds)
    assert np.array_equal(ctx.solid_flags_export(), osolid_flags)

    n_kmers, n_edges = ctx.dbg_adjacency()
    kmers, adj = ctx.dbg_export()
    osolid = oracle.solid_kmers(seq, off, k, okeys, ocounts)[:, 0]
    assert n_kmers == n_solid == len(osolid)
    assert np.array_equal(kmers, osolid)
    assert n_edges == int(np.unpackbits(adj).sum())
    step = max(1, len(osolid) // n_adj)
    for i in range(0, len(osolid), step):
        assert adj[i] == oracle.check_directions(obits, fs, nh, osolid[i:i + 1], k), (i, hex(int(osolid[i])))
    return dict(keys=okeys, bits=obits, fs=fs, nh=nh, solid=osolid)


@pytest.mark.parametrize("k", [21, 22, 25, 27, 28, 29, 31, 32])
def test_pipeline_small(ctx, oracle, k):
    reads = _dataset(k, genome=5000, cov=15, rl=90, err=0.01)
    _full_check(ctx, oracle, reads, k)


@pytest.mark.parametrize("env", [
    {"P3_COUNT_MODE": "direct"},
    {"P3_PARTS": "1"},
    {"P3_PARTS": "7"},
    {"P3_PARTS": "1024", "P3_BIN_BUDGET_BYTES": "200000"},   # many partitions, many chunks
    {"P3_PARTS": "3", "P3_BIN_BUDGET_BYTES": "1"},            # one tile per chunk
    {"P3_BINNED_CLEARS": "1", "P3_BLOOM_SEG_BITS": "4096"},   # coverage-bit clears binned by plane segment
    {"P3_BINNED_CLEARS": "1", "P3_BLOOM_SEG_BITS": "1024", "P3_PARTS": "5"},
    {"P3_DEDUPE_BINNED": "1", "P3_SET_PARTS": "7"},           # k-mer de-duplication binned by set partition
    {"P3_DEDUPE_BINNED": "1", "P3_SET_PARTS": "1"},
    {"P3_DEDUPE_BINNED": "0", "P3_SET_PARTS": "13"},          # direct de-duplication into a partitioned set
    {"P3_EXACT_BINS": "1"},                                    # histogram-sized bins instead of fixed-capacity ones
    {"P3_EXACT_BINS": "1", "P3_PARTS": "9", "P3_BIN_BUDGET_BYTES": "300000"},
])
def test_count_modes(ctx, oracle, env, monkeypatch):
    """direct table vs binned/partitioned count (any partition count, any chunking) are all exact"""
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    reads = _dataset(61, genome=6000, cov=14, rl=110, err=0.01)
    _full_check(ctx, oracle, reads, 32)
    _full_check(ctx, oracle, [b"A" * 300, b"T" * 250, b"ACGT" * 50] * 20 + reads[:200], 25, m=30011)


@pytest.mark.parametrize("seg_bits", [1024, 4096, 1 << 16])
def test_binned_bloom_adds(ctx, oracle, seg_bits, monkeypatch):
    """BF.add binned by filter segment (p3_bloom.inc.cu) sets exactly the reference's bits, for the
    auto-sized filter (19 hashes), -m filters (10 hashes; sizes that are not a multiple of the segment,
    and tiny ones where (h1 + n*h2) wraps past 2^64 relative to the modulus) and multi-word k"""
    monkeypatch.setenv("P3_BLOOM_BINNED", "1")
    monkeypatch.setenv("P3_BLOOM_SEG_BITS", str(seg_bits))
    reads = _dataset(77, genome=6000, cov=16, rl=100, err=0.01)
    _full_check(ctx, oracle, reads, 32)
    _full_check(ctx, oracle, reads, 25, m=30011)
    _full_check(ctx, oracle, reads, 27, m=3 * seg_bits)
    _full_check(ctx, oracle, reads[:300], 31, m=1000003)
    _long_check(ctx, oracle, _dataset(78, genome=3000, cov=30, rl=150, err=0.003), 63)


def test_binned_dedupe_overflow_falls_back(ctx, oracle, monkeypatch):
    """a solid k-mer with very many occurrences overflows its set partition's bin: the direct path takes over"""
    monkeypatch.setenv("P3_DEDUPE_BINNED", "1")
    monkeypatch.setenv("P3_SET_PARTS", "64")
    reads = [b"ACGTTGCA" * 40] * 4000 + _dataset(93, genome=5000, cov=12, rl=100, err=0.01)
    _full_check(ctx, oracle, reads, 32, n_adj=300)
    _full_check(ctx, oracle, reads, 25, m=30011, n_adj=300)


def test_fixed_capacity_bins_overflow_falls_back(ctx, oracle, monkeypatch):
    """The binned count gives every table partition room for its expected share of a chunk; a heavy-hitter
    key (here a homopolymer: 1.1 M occurrences of ONE 21-mer) overflows its partition, which must be
    detected and redone with exact bins — counts stay bit-exact."""
    monkeypatch.setenv("P3_PARTS", "64")
    reads = [b"A" * 400] * 3000 + _dataset(91, genome=5000, cov=12, rl=100, err=0.01)
    _full_check(ctx, oracle, reads, 32, n_adj=300)


def test_threshold_other_than_two(ctx, oracle):
    """cov_threshold != 2 takes the full-lookup path; 1 makes every k-mer solid, 3 fewer"""
    reads = _dataset(62, genome=3000, cov=12, rl=90, err=0.01)
    seq, off = reads_to_arrays(reads)
    ctx.load_ascii(seq, off)
    ctx.count_short_kmers()
    a1, _ = ctx.make_bf(25, 50021, 10, cov_threshold=1)
    assert a1 == sum(len(r) - 24 for r in reads)
    a2, _ = ctx.make_bf(25, 50021, 10, cov_threshold=2)
    a3, _ = ctx.make_bf(25, 50021, 10, cov_threshold=3)
    assert a1 > a2 > a3 > 0
    okeys, ocounts = oracle.count_short_kmers(seq, off)
    _, _, _, oadds = oracle.make_bf(seq, off, 25, okeys, ocounts, 50021, 10)
    assert a2 == oadds
    # threshold 3 == oracle with every count-2 key demoted to 1
    dem = ocounts.copy(); dem[dem == 2] = 1
    _, _, _, oadds3 = oracle.make_bf(seq, off, 25, okeys, dem, 50021, 10)
    assert a3 == oadds3


def test_pipeline_error_free(ctx, oracle):
    """configs[0] shape in miniature: error-free reads, every k-mer solid"""
    reads = _dataset(77, genome=20000, cov=30, rl=150, err=0.0)
    _full_check(ctx, oracle, reads, 32)


def test_pipeline_medium_with_errors(ctx, oracle):
    """configs[1] shape in miniature: 1% substitutions, Bloom screening of error k-mers"""
    reads = _dataset(78, genome=200000, cov=50, rl=150, err=0.01)
    _full_check(ctx, oracle, reads, 32, n_adj=5000)


def test_pipeline_m_option(ctx, oracle):
    reads = _dataset(5, genome=4000, cov=12, rl=80, err=0.02)
    _full_check(ctx, oracle, reads, 25, m=100003)


def test_ragged_reads_and_non_acgt(ctx, oracle):
    """reads of many lengths (one exactly k), N / lower-case bases (reference quirk: code 0 on
    both strands), read ends at every word phase"""
    k = 25
    rng = np.random.default_rng(3)
    g = synth.random_genome(3000, 11)
    reads = []
    for i in range(700):
        L = int(rng.integers(k, 140)) if i else k
        s = int(rng.integers(0, len(g) - L))
        r = bytearray(synth.codes_to_ascii(g[s:s + L]).tobytes())
        if i % 7 == 0:
            r[int(rng.integers(0, L))] = ord("N")
        if i % 11 == 0:
            r[int(rng.integers(0, L))] = ord("a")
        reads.append(bytes(r))
    _full_check(ctx, oracle, reads, k, m=70001)


def test_long_reads(ctx, oracle):
    reads = _dataset(21, genome=30000, cov=12, rl=5000, err=0.01)
    _full_check(ctx, oracle, reads, 32)


def test_tight_tables_grow_or_fail_loudly(ctx, oracle):
    reads = _dataset(9, genome=5000, cov=15, rl=90, err=0.01)
    seq, off = reads_to_arrays(reads)
    ctx.load_ascii(seq, off)
    with pytest.raises(_lib.P3Error) as e:
        ctx.count_short_kmers(table_slots=64)
    assert e.value.code == -3
    # high load factor still exact; tiny solid set grows by itself
    okeys, _ = oracle.count_short_kmers(seq, off)
    _full_check(ctx, oracle, reads, 32, table_slots=int(len(okeys) * 1.05), solid_slots=16)


def test_repeats_and_homopolymers(ctx, oracle):
    """hot keys: poly-A / poly-T reads (canonical 0), tandem repeats, palindromic k-mers"""
    k = 22
    reads = [b"A" * 200, b"T" * 180, b"AC" * 100, b"ACGT" * 60, b"GAATTC" * 40] * 30
    reads += _dataset(4, genome=2000, cov=10, rl=100, err=0.0)
    _full_check(ctx, oracle, reads, k, m=40009)


def test_count_overflow_side_table(ctx, oracle):
    """counts beyond the 22-bit slot field stay exact (overflow side table)"""
    n = 5_000_000
    seq = np.full(n, ord("A"), np.uint8)
    off = np.array([0, n], np.uint64)
    ctx.load_ascii(seq, off)
    n_pos, n_distinct = ctx.count_short_kmers(4096)
    keys, counts = ctx.short_kmer_export()
    assert n_distinct == 1 and keys[0] == 0
    assert counts[0] == n - 20 == n_pos and counts[0] > (1 << 22)
    assert ctx.short_kmer_lookup(np.array([0, 5], np.uint64)).tolist() == [n - 20, 0]
    n_adds, n_solid = ctx.make_bf(32, 10007, 10)
    assert n_adds == n - 31 and n_solid == 1


def test_bf_primitives(ctx, oracle):
    """BF::add / possiblyContains / GetDoubleHash_64bit batched over canonical k-mers"""
    rng = np.random.default_rng(8)
    for k in (21, 25, 28, 29, 32):
        fs, nh = 50021, 7
        kmers = np.array([oracle.canonical_words("".join("ACGT"[i] for i in rng.integers(0, 4, k)), k)[0]
                          for _ in range(400)], np.uint64)
        dh = ctx.double_hash(k, kmers)
        for i in range(0, 400, 13):
            assert tuple(int(x) for x in dh[i]) == oracle.double_hash(oracle.std_hash_kmer(kmers[i:i + 1], k))
        ctx.bf_import(k, fs, nh, None)
        ctx.bf_add(kmers[:200])
        obits = np.zeros((fs + 7) // 8, np.uint8)
        for i in range(200):
            oracle.bf_add(obits, fs, nh, kmers[i:i + 1], k)
        assert np.array_equal(ctx.bf_export(), obits)
        got = ctx.bf_possibly_contains(kmers)
        want = np.array([oracle.L.p3o_bf_possibly_contains(obits, fs, nh, kmers[i:i + 1], k) for i in range(400)], np.uint8)
        assert np.array_equal(got, want) and got[:200].all()


def test_check_directions_oriented(ctx, oracle):
    """CheckDirections on oriented (non-canonical) k-mers, incl. k-mers that are not in the set"""
    k = 31
    reads = _dataset(31, genome=4000, cov=12, rl=100, err=0.005)
    info = _full_check(ctx, oracle, reads, k)
    rng = np.random.default_rng(2)
    probes = [kmer_str_to_words(r[p:p + k].decode(), k)[0] for r in reads[:50] for p in range(0, 60, 9)]
    probes += [int(x) for x in rng.integers(0, 1 << 62, 200)]
    probes = np.array(probes, np.uint64)
    got = ctx.check_directions(probes)
    want = np.array([oracle.check_directions(info["bits"], info["fs"], info["nh"], probes[i:i + 1], k) for i in range(len(probes))], np.uint8)
    assert np.array_equal(got, want)


def test_whole_path_entry_point(ctx, oracle):
    """p3_assemble_hot_path on host staging buffers == staged calls"""
    k = 32
    reads = _dataset(12, genome=8000, cov=20, rl=120, err=0.01)
    seq, off = reads_to_arrays(reads)
    packed, nmask = _lib.pack_reads(seq, off)
    ctx.assemble_hot_path(packed, off, k, nmask)
    fs, nh = oracle.estimate_bloomfilter(int(off[-1]), k)
    assert (ctx.filter_size, ctx.num_hashes) == (fs, nh)
    okeys, ocounts = oracle.count_short_kmers(seq, off)
    obits, oseeds, _, _ = oracle.make_bf(seq, off, k, okeys, ocounts, fs, nh)
    assert np.array_equal(ctx.bf_export(), obits)
    assert np.array_equal(ctx.seed_export(), oseeds)
    kmers, adj = ctx.dbg_export()
    assert np.array_equal(kmers, oracle.solid_kmers(seq, off, k, okeys, ocounts)[:, 0])


def test_properties_at_scale(ctx):
    """size-independent properties on a workload too big for the oracle: counts sum to the
    number of positions, every solid k-mer answers possiblyContains, adjacency is symmetric
    (a right edge K->N implies N, re-oriented, has the left edge back to K)"""
    k = 32
    g = synth.random_genome(2_000_000, 5)
    codes = synth.simulate_reads(g, 30, 150, 0.01, 6)
    seq = synth.codes_to_ascii(codes).reshape(-1)
    off = (np.arange(codes.shape[0] + 1, dtype=np.uint64) * np.uint64(150))
    ctx.load_ascii(seq, off)
    n_pos, n_distinct = ctx.count_short_kmers()
    assert n_pos == codes.shape[0] * 130
    keys, counts = ctx.short_kmer_export()
    assert int(counts.sum()) == n_pos and len(np.unique(keys)) == len(keys) == n_distinct
    fs, nh = _lib.estimate_bloomfilter(int(off[-1]), k)
    n_adds, n_solid = ctx.make_bf(k, fs, nh)
    ctx.dbg_adjacency()
    kmers, adj = ctx.dbg_export()
    assert len(np.unique(kmers)) == len(kmers) == n_solid
    assert ctx.bf_possibly_contains(kmers).all()
    # genome k-mers are (almost all) solid at 30x
    gk = 0
    sel = np.arange(0, len(g) - k, 997)
    win = g[sel[:, None] + np.arange(k)[None, :]].astype(np.uint64)
    f = np.zeros(len(sel), np.uint64)
    r = np.zeros(len(sel), np.uint64)
    for j in range(k):
        f = (f << np.uint64(2)) | win[:, j]
        r = r | ((np.uint64(3) - win[:, j]) << np.uint64(2 * j))
    canon = np.minimum(f, r)
    assert np.isin(canon, kmers).mean() > 0.99
    # edge symmetry through the oriented query entry point
    sub = kmers[:: max(1, len(kmers) // 20000)]
    a = ctx.check_directions(sub)
    mask = np.uint64(0xFFFFFFFFFFFFFFFF)
    for d in range(4, 8):
        has = (a >> d) & 1 == 1
        nbr = ((sub[has] << np.uint64(2)) | np.uint64(d - 4)) & mask
        back = ctx.check_directions(nbr)
        left_base = (sub[has] >> np.uint64(2 * k - 2)).astype(np.int64)
        assert np.all((back >> left_base.astype(np.uint8)) & 1 == 1)


# ---------------------------------------------------------------- multi-word k (33 <= k <= 3001)
def _long_check(ctx, oracle, reads, k, m=0, n_adj=400):
    from _checkers import nwords
    W = nwords(k)
    seq, off = reads_to_arrays(reads)
    fs, nh = (m, 10) if m else _lib.estimate_bloomfilter(int(off[-1]), k)
    ctx.load_ascii(seq, off)
    ctx.count_short_kmers()
    okeys, ocounts = oracle.count_short_kmers(seq, off)
    n_adds, n_solid = ctx.make_bf(k, fs, nh)
    obits, oseeds, osolid_flags, oadds = oracle.make_bf(seq, off, k, okeys, ocounts, fs, nh, want_solid=True)
    assert n_adds == oadds
    assert np.array_equal(ctx.solid_flags_export(), osolid_flags)
    assert np.array_equal(ctx.seed_export(), oseeds)
    assert np.array_equal(ctx.bf_export(), obits)
    osolid = oracle.solid_kmers(seq, off, k, okeys, ocounts)
    n_kmers, n_edges = ctx.dbg_adjacency()
    kmers, adj = ctx.dbg_export()
    assert n_kmers == n_solid == len(osolid)
    assert kmers.shape == (len(osolid), W) and np.array_equal(kmers, osolid)
    assert n_edges == int(np.unpackbits(adj).sum())
    for i in range(0, len(osolid), max(1, len(osolid) // n_adj)):
        assert adj[i] == oracle.check_directions(obits, fs, nh, osolid[i], k), (k, i)
    return dict(bits=obits, fs=fs, nh=nh, solid=osolid)


@pytest.mark.parametrize("k,rl", [(33, 100), (47, 120), (63, 150), (64, 150), (65, 150), (96, 200), (101, 250)])
def test_long_k_pipeline(ctx, oracle, k, rl):
    reads = _dataset(300 + k, genome=4000, cov=40, rl=rl, err=0.003)
    _long_check(ctx, oracle, reads, k)


def test_long_k_quirks_and_m(ctx, oracle):
    """k = 63 with ragged reads (one exactly k long), non-ACGT bases, homopolymers, explicit -m"""
    k = 63
    rng = np.random.default_rng(4)
    g = synth.random_genome(3000, 12)
    reads = [b"A" * 200, b"T" * 150, b"AC" * 90]
    for i in range(500):
        L = int(rng.integers(k, 260)) if i else k
        s0 = int(rng.integers(0, len(g) - L))
        r = bytearray(synth.codes_to_ascii(g[s0:s0 + L]).tobytes())
        if i % 9 == 0:
            r[int(rng.integers(0, L))] = ord("N")
        reads.append(bytes(r))
    _long_check(ctx, oracle, reads, k, m=90001)


def test_very_long_k(ctx, oracle):
    """k = 501 and the reference's maximum k = 3001 on long error-free reads"""
    g = synth.random_genome(9000, 21)
    for k, rl in ((501, 2000), (3001, 6000)):
        reads = synth.reads_as_bytes(synth.simulate_reads(g, 12, rl, 0.0, 22 + k))
        _long_check(ctx, oracle, reads, k, m=200003, n_adj=60)


def test_long_k_batch_primitives(ctx, oracle):
    """BF.add / possiblyContains / GetDoubleHash_64bit / CheckDirections on explicit W-word k-mers"""
    from _checkers import nwords
    rng = np.random.default_rng(6)
    for k in (33, 63, 64, 101, 1001):
        W = nwords(k)
        fs, nh = 70001, 9
        kmers = np.stack([oracle.canonical_words("".join("ACGT"[i] for i in rng.integers(0, 4, k)), k) for _ in range(120)])
        dh = ctx.double_hash(k, kmers)
        for i in range(0, 120, 7):
            assert tuple(int(x) for x in dh[i]) == oracle.double_hash(oracle.std_hash_kmer(kmers[i], k))
        ctx.bf_import(k, fs, nh, None)
        ctx.bf_add(kmers[:60])
        obits = np.zeros((fs + 7) // 8, np.uint8)
        for i in range(60):
            oracle.bf_add(obits, fs, nh, kmers[i], k)
        assert np.array_equal(ctx.bf_export(), obits)
        got = ctx.bf_possibly_contains(kmers)
        want = np.array([oracle.L.p3o_bf_possibly_contains(obits, fs, nh, np.ascontiguousarray(kmers[i]), k) for i in range(120)], np.uint8)
        assert np.array_equal(got, want) and got[:60].all()
        masks = ctx.check_directions(kmers)
        for i in range(0, 120, 5):
            assert masks[i] == oracle.check_directions(obits, fs, nh, kmers[i], k)
