"""Host side of the drop-in: Load parity on CPU; the whole run (GPU hot path + host walk +
coverage + GFA) against the golden fixtures made from the reference with -t 1."""
import glob
import os
import subprocess

import numpy as np
import pytest

from platanus3_b200 import _lib, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    return g, os.path.join(GOLD, str(g["read_file"]))


@pytest.mark.parametrize("name", CASES)
def test_load_file_matches_oracle(oracle, name):
    """ReadFile::LoadFile: same reads, same all_bases, 2-bit staging == p3_pack_reads of them"""
    g, path = _load(name)
    k = int(g["k"])
    mine = _lib.load_file(path, k)
    seq, off, all_bases = oracle.load_reads(path, k)
    assert mine["all_bases"] == all_bases == int(g["all_bases"])
    assert len(mine["off"]) - 1 == int(g["n_reads"])
    a = sorted(mine["seq"][int(mine["off"][i]):int(mine["off"][i + 1])].tobytes() for i in range(len(mine["off"]) - 1))
    b = sorted(seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1))
    assert a == b
    packed, nmask = _lib.pack_reads(mine["seq"], mine["off"])
    assert np.array_equal(packed, mine["packed"])
    assert (nmask is None) == (mine["nmask"] is None)
    if nmask is not None:
        assert np.array_equal(nmask, mine["nmask"])


def _numpy_pack(seq):
    """straightforward restatement of the staging layout: 32 bases per uint64, first base on top; a
    byte outside ACGT packs as 0 and sets its bit (MSB first) in the non-ACGT plane"""
    n = len(seq)
    words = (n + 31) // 32 + 1
    code = np.full(256, 4, np.uint8)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
    c = code[seq]
    bad = (c >> 2).astype(np.uint64)
    c = (c & 3).astype(np.uint64)
    pad = (-n) % 32
    c = np.concatenate([c, np.zeros(pad, np.uint64)]).reshape(-1, 32)
    bad = np.concatenate([bad, np.zeros(pad, np.uint64)]).reshape(-1, 32)
    pk = np.zeros(words, np.uint64)
    nm = np.zeros(words, np.uint32)
    pk[:len(c)] = (c << np.arange(62, -2, -2, dtype=np.uint64)).sum(axis=1, dtype=np.uint64)
    nm[:len(c)] = (bad << np.arange(31, -1, -1, dtype=np.uint64)).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    return pk, nm


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 257, 100_003, (1 << 25) + 77])
def test_pack_reads_matches_the_layout_definition(n):
    """p3_pack_reads (8 bases per step with BMI2 bit gathers, threads above 2^20 words) against the plain
    definition, with every byte value present and ~2 % arbitrary bytes"""
    rng = np.random.default_rng(n)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    if n > 40:
        idx = rng.integers(0, n, max(1, n // 50))
        seq[idx] = rng.integers(0, 256, len(idx)).astype(np.uint8)
    if n >= 257:
        seq[:256] = np.arange(256, dtype=np.uint8)
    pk, nm = _lib.pack_reads(seq, np.array([0, n], np.uint64))
    rpk, rnm = _numpy_pack(seq)
    assert np.array_equal(pk, rpk)
    assert (nm is None and not rnm.any()) or (nm is not None and np.array_equal(nm, rnm))


def _messy_file(rng, kind, k):
    """a read file full of the things real files do: multi-line FASTA with ragged widths, blank lines,
    repeated and empty name lines, reads shorter than / exactly k, non-ACGT and lower-case bytes,
    CRLF line ends, a missing final newline, '@' and '>' inside FASTQ quality lines"""
    out = []
    n = int(rng.integers(1, 40))
    names = ["%s r%d x" % (">" if kind == "fasta" else "@", i) for i in range(n)]
    for i in range(n):
        r = int(rng.integers(0, 12))
        if r == 0:
            names[i] = names[int(rng.integers(0, n))]              # repeated name line
        elif r == 1:
            names[i] = ">" if kind == "fasta" else "@"              # nothing but the marker
        L = int(rng.choice([k - 1, k, k + 1, int(rng.integers(1, 4 * k))]))
        seq = bytearray(b"ACGT"[j] for j in rng.integers(0, 4, L))
        for _ in range(int(rng.integers(0, 3))):
            if L:
                seq[int(rng.integers(0, L))] = int(rng.choice(list(b"NnacgtRY-")))
        eol = b"\r\n" if rng.integers(0, 10) == 0 else b"\n"
        if kind == "fasta":
            out.append(names[i].encode() + eol)
            w = int(rng.integers(1, 2 * k))
            for j in range(0, max(L, 1), w):
                out.append(bytes(seq[j:j + w]) + eol)
                if rng.integers(0, 25) == 0:
                    out.append(eol)                                   # blank line inside a record
        else:
            qual = bytes(rng.choice(list(b"@>+I!#5"), L).tolist())
            out.append(names[i].encode() + eol + bytes(seq) + eol + b"+" + eol + qual + eol)
            if rng.integers(0, 30) == 0:
                out.append(eol)             # a stray blank line: shifts the 4-line rhythm, empty "name" lines follow
    data = b"".join(out)
    if rng.integers(0, 3) == 0 and data.endswith(b"\n"):
        data = data[:-1]                                              # no newline at the end of the file
    return data


@pytest.mark.parametrize("chunk", [None, 1, 7, 64, 1000])
@pytest.mark.parametrize("kind", ["fasta", "fastq"])
def test_load_file_differential(oracle, tmp_path, kind, chunk, monkeypatch):
    """p3_load_file (parallel over byte chunks of the mapped file) against the reference's own LoadFile
    (when the reference is compiled here) and the oracle's restatement, on 60 messy files per format;
    chunk sizes down to ONE byte put a chunk boundary at every possible place of a record"""
    from _checkers import Ref, have_ref
    if chunk is not None:
        monkeypatch.setenv("P3_LOAD_CHUNK_BYTES", str(chunk))
    k = 21
    for trial in range(60):
        rng = np.random.default_rng(1000 * (kind == "fastq") + trial)
        path = str(tmp_path / ("t%d_x.%s" % (trial, kind)))
        with open(path, "wb") as f:
            f.write(_messy_file(rng, kind, k))
        mine = _lib.load_file(path, k)
        a = sorted(mine["seq"][int(mine["off"][i]):int(mine["off"][i + 1])].tobytes() for i in range(len(mine["off"]) - 1))
        seq, off, all_bases = oracle.load_reads(path, k)
        b = sorted(seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1))
        assert a == b and mine["all_bases"] == all_bases, trial
        if have_ref():
            ref = Ref(k, readfile=path)
            ref.load_file()
            assert a == sorted(ref.reads()) and mine["all_bases"] == ref.all_bases, trial
            ref.close()


def test_load_file_edge_cases(tmp_path):
    p = tmp_path / "empty.fasta"
    p.write_bytes(b"")
    assert len(_lib.load_file(str(p), 21)["off"]) == 1
    p = tmp_path / "noheader.fasta"
    p.write_bytes(b"ACGTACGTACGTACGTACGTACGTACGT\n")
    assert len(_lib.load_file(str(p), 21)["off"]) == 1          # first byte is neither '>' nor '@'
    p = tmp_path / "crlf_.fasta"
    p.write_bytes(b">a\r\nACGTACGTACGTACGTACGTACGT\r\n")
    r = _lib.load_file(str(p), 21)                                # '\r' stays in the read, like std::getline
    assert r["seq"].tobytes() == b"ACGTACGTACGTACGTACGTACGT\r" and r["nmask"] is not None
    with pytest.raises(_lib.P3Error):
        _lib.load_file("a.fa", 21)                                 # name shorter than 5 (Load.cpp:26)
    with pytest.raises(_lib.P3Error):
        _lib.load_file(str(tmp_path / "missing.fasta"), 21)


def _closed_table_from_oracle(oracle, g, path):
    """solid k-mers + adjacency + Bloom-false-positive closure + seeds, all from the oracle (CPU)"""
    from _checkers import kmer_str_to_words, nwords
    k, fs, nh = int(g["k"]), int(g["filter_size"]), int(g["num_hashes"])
    W = nwords(k)
    seq, off, _ = oracle.load_reads(path, k)
    keys, counts = np.ascontiguousarray(g["keys"], np.uint64), np.ascontiguousarray(g["counts"], np.uint64)
    bloom = np.ascontiguousarray(g["bloom"], np.uint8)
    solid = oracle.solid_kmers(seq, off, k, keys, counts)
    mask = (1 << (2 * k)) - 1

    def to_int(w):
        return sum(int(x) << (64 * i) for i, x in enumerate(w))

    def to_words(v):
        return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(W)], np.uint64)

    def canon(v):
        r, x = 0, v
        for _ in range(k):
            r = (r << 2) | (3 - (x & 3))
            x >>= 2
        return min(v, r)

    table = {}
    frontier = []
    for w in solid:
        v = to_int(w)
        table[v] = oracle.check_directions(bloom, fs, nh, w, k)
        frontier.append(v)
    seeds = [to_int(kmer_str_to_words(str(sd), k)) for sd in g["seeds"]]
    for sd in seeds:                       # walk roots that are not solid get an entry too
        c = canon(sd)
        if c not in table:
            table[c] = oracle.check_directions(bloom, fs, nh, to_words(c), k)
            frontier.append(c)
    while frontier:                        # closure under reported neighbours (Bloom false positives)
        nxt = []
        for v in frontier:
            a = table[v]
            for d in range(8):
                if (a >> d) & 1:
                    nb = (v >> 2) | (d << (2 * k - 2)) if d < 4 else ((v << 2) | (d - 4)) & mask
                    c = canon(nb)
                    if c not in table:
                        table[c] = oracle.check_directions(bloom, fs, nh, to_words(c), k)
                        nxt.append(c)
        frontier = nxt
    kk = np.array([to_words(v) for v in table], np.uint64).reshape(-1, W)
    aa = np.array(list(table.values()), np.uint8)
    ss = np.array([to_words(v) for v in seeds], np.uint64).reshape(-1, W)
    return kk, aa, ss, len(solid)


@pytest.mark.parametrize("name", CASES)
def test_host_walk_over_oracle_table_matches_reference_gfa(oracle, name, tmp_path):
    """The host half of the drop-in (Load, MakeDBG in -t 1 order, CountNodeCoverage on the host, PrintGraph)
    needs no GPU when it is handed a closed CheckDirections table: here the table comes from the oracle, the
    GFA must equal the reference's — for single-word k (uint64 walk) and for k = 63 / 101 (string walk)."""
    g, path = _load(name)
    k = int(g["k"])
    kk, aa, ss, n_solid = _closed_table_from_oracle(oracle, g, path)
    assert len(aa) >= n_solid
    gfa = str(tmp_path / "walk.gfa")
    st = _lib.walk_table(path, k, kk, aa, ss, gfa_path=gfa)
    assert (st["junctions"], st["joints"], st["straights"]) == (int(g["n_junctions"]), int(g["n_joints"]), int(g["n_straights"]))
    assert sorted(open(gfa).read().splitlines()) == [str(x) for x in g["gfa"]]


def test_host_walk_refuses_an_open_table(oracle, tmp_path):
    g, path = _load("k25_err")
    kk, aa, ss, n_solid = _closed_table_from_oracle(oracle, g, path)
    if len(aa) == n_solid:
        pytest.skip("no false-positive k-mers in this fixture")
    with pytest.raises(_lib.P3Error):      # drop the closure's additions: the walk must notice, not mis-assemble
        _lib.walk_table(path, 25, kk[:n_solid], aa[:n_solid], ss, gfa_path=str(tmp_path / "x.gfa"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_assemble_file_matches_reference_gfa(name, tmp_path):
    """GFA (as a set of lines), node counts and seed count equal the reference's -t 1 run"""
    g, path = _load(name)
    gfa, log = str(tmp_path / "out.gfa"), str(tmp_path / "out.log")
    st = _lib.assemble_file(path, int(g["k"]), m=int(g["m"]), threads=1, gfa_path=gfa, log_path=log)
    assert (st["junctions"], st["joints"], st["straights"]) == (int(g["n_junctions"]), int(g["n_joints"]), int(g["n_straights"]))
    assert st["reads"] == int(g["n_reads"]) and st["all_bases"] == int(g["all_bases"])
    assert st["distinct_21mers"] == len(g["keys"])
    assert sorted(open(gfa).read().splitlines()) == [str(x) for x in g["gfa"]]
    lines = open(log).read().splitlines()
    assert "seed kmer num= %d" % len(g["seeds"]) in lines
    assert "filter_size : %d" % int(g["filter_size"]) in lines and "num_hashes : %d" % int(g["num_hashes"]) in lines
    assert lines[-1] == "finish"


@pytest.mark.gpu
def test_cli_drop_in(tmp_path):
    """platanus3-compatible command line: -i -k -t [-m], ./platanus3.log and ./de_bruijn_graph.gfa in cwd"""
    g, path = _load("k25_err")
    exe = os.path.join(ROOT, "platanus3_b200", "platanus3_b200")
    r = subprocess.run([exe, "-i", path, "-k", "25", "-t", "4"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert sorted((tmp_path / "de_bruijn_graph.gfa").read_text().splitlines()) == [str(x) for x in g["gfa"]]
    assert (tmp_path / "platanus3.log").read_text().splitlines()[-1] == "finish"
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.startswith("Usage: platanus3 -i")      # main.cpp:16-19
    r = subprocess.run([exe, "-i", path, "-k", "3002"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "not supported" in r.stderr


@pytest.mark.gpu
def test_closure_covers_every_reported_neighbour(oracle):
    """after p3_dbg_close every neighbour the table reports is itself in the table"""
    from _checkers import reads_to_arrays
    k = 27
    gnm = synth.random_genome(8000, 9)
    reads = synth.reads_as_bytes(synth.simulate_reads(gnm, 30, 100, 0.01, 10))
    seq, off = reads_to_arrays(reads)
    fs, nh = _lib.estimate_bloomfilter(int(off[-1]), k)
    with _lib.Context(0) as ctx:
        ctx.load_ascii(seq, off)
        ctx.count_short_kmers()
        ctx.make_bf(k, fs, nh)
        n_solid, _ = ctx.dbg_adjacency()
        n_total = ctx.dbg_close()
        kmers, adj = ctx.dbg_export()
        assert len(kmers) == n_total >= n_solid
        assert np.array_equal(adj, ctx.check_directions(kmers))
        table = set(int(x) for x in kmers)
        mask = (1 << (2 * k)) - 1

        def canon(v):
            r = 0
            x = v
            for _ in range(k):
                r = (r << 2) | (3 - (x & 3))
                x >>= 2
            return min(v, r)
        for km, a in list(zip(kmers.tolist(), adj.tolist()))[:: max(1, len(kmers) // 3000)]:
            for d in range(8):
                if (a >> d) & 1:
                    nb = (km >> 2) | (d << (2 * k - 2)) if d < 4 else ((km << 2) | (d - 4)) & mask
                    assert canon(nb) in table
