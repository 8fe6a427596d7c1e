"""ctypes bindings for the two parity checkers (TEST INFRASTRUCTURE):

  Oracle  -> oracle/libp3oracle.so   (our C restatement, oracle/p3_oracle.c)
  Ref     -> oracle/_ref/libp3ref.so (the unmodified reference compiled by oracle/Makefile)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libp3oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libp3ref.so")

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build_checkers():
    """Compile the oracle (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def nwords(k):
    return (2 * k + 63) // 64


def reads_to_arrays(reads):
    """list[str|bytes] -> (uint8 concatenated ASCII, uint64 offsets[n+1])"""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    seq = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return seq, off


def kmer_str_to_words(s, k):
    """ASCII k-mer -> little-endian uint64 words of the reference's bitset<2k> value."""
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    v = 0
    for ch in (s if isinstance(s, str) else s.decode()):
        v = (v << 2) | code.get(ch, 0)
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(nwords(k))], dtype=np.uint64)


def words_to_kmer_str(w, k):
    v = 0
    for i, x in enumerate(w):
        v |= int(x) << (64 * i)
    return "".join("ACGT"[(v >> (2 * (k - 1 - j))) & 3] for j in range(k))


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_checkers()
        L = self.L = C.CDLL(ORACLE_SO)
        L.p3o_hash_bytes.restype = C.c_uint64
        L.p3o_hash_bytes.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64]
        L.p3o_std_hash_kmer.restype = C.c_uint64
        L.p3o_std_hash_kmer.argtypes = [u64p, C.c_int]
        L.p3o_double_hash.argtypes = [C.c_uint64, u64p]
        L.p3o_estimate_bloomfilter.argtypes = [C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.p3o_first_kmer_forward.argtypes = [C.c_char_p, C.c_int, u64p]
        L.p3o_first_kmer_backward.argtypes = [C.c_char_p, C.c_int, u64p]
        L.p3o_complement_kmer.argtypes = [u64p, C.c_int, u64p]
        L.p3o_compare_bit.argtypes = [u64p, u64p, C.c_int]
        L.p3o_load_reads.restype = C.c_int64
        L.p3o_load_reads.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.p3o_count_short_kmers.restype = C.c_uint64
        L.p3o_count_short_kmers.argtypes = [u8p, u64p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.p3o_rmq.restype = C.c_uint64
        L.p3o_rmq.argtypes = [u64p, C.c_uint64, C.c_int, u64p]
        L.p3o_make_bf.restype = C.c_uint64
        L.p3o_make_bf.argtypes = [u8p, u64p, C.c_uint64, C.c_int, u64p, u64p, C.c_uint64,
                                  C.c_uint64, C.c_int, u8p, i64p, C.c_void_p]
        L.p3o_bf_add.argtypes = [u8p, C.c_uint64, C.c_int, u64p, C.c_int]
        L.p3o_bf_possibly_contains.argtypes = [u8p, C.c_uint64, C.c_int, u64p, C.c_int]
        L.p3o_is_recorded.argtypes = [u8p, C.c_uint64, C.c_int, u64p, C.c_int]
        L.p3o_check_directions.argtypes = [u8p, C.c_uint64, C.c_int, u64p, C.c_int, C.c_int]
        L.p3o_solid_kmers.restype = C.c_uint64
        L.p3o_solid_kmers.argtypes = [u8p, u64p, C.c_uint64, C.c_int, u64p, u64p, C.c_uint64, C.c_void_p]

    def hash_bytes(self, b, seed=0xC70F6907):
        return self.L.p3o_hash_bytes(b, len(b), seed)

    def std_hash_kmer(self, words, k):
        return self.L.p3o_std_hash_kmer(np.ascontiguousarray(words, np.uint64), k)

    def double_hash(self, h0):
        out = np.zeros(2, np.uint64)
        self.L.p3o_double_hash(h0, out)
        return int(out[0]), int(out[1])

    def estimate_bloomfilter(self, all_bases, k):
        fs, nh = C.c_uint64(), C.c_int()
        self.L.p3o_estimate_bloomfilter(all_bases, k, C.byref(fs), C.byref(nh))
        return fs.value, nh.value

    def canonical_words(self, s, k):
        b = s.encode() if isinstance(s, str) else s
        f = np.zeros(nwords(k), np.uint64)
        r = np.zeros(nwords(k), np.uint64)
        self.L.p3o_first_kmer_forward(b, k, f)
        self.L.p3o_first_kmer_backward(b, k, r)
        return f if self.L.p3o_compare_bit(f, r, k) == 0 else r

    def load_reads(self, path, k):
        tot, ab = C.c_uint64(), C.c_uint64()
        n = self.L.p3o_load_reads(path.encode(), k, None, None, C.byref(tot), C.byref(ab))
        if n < 0:
            raise IOError(path)
        seq = np.zeros(max(tot.value, 1), np.uint8)
        off = np.zeros(n + 1, np.uint64)
        self.L.p3o_load_reads(path.encode(), k, seq.ctypes.data, off.ctypes.data, C.byref(tot), C.byref(ab))
        return seq[: tot.value], off, ab.value

    def count_short_kmers(self, seq, off):
        n_reads = len(off) - 1
        n = self.L.p3o_count_short_kmers(seq, off, n_reads, None, None)
        keys = np.zeros(n, np.uint64)
        counts = np.zeros(n, np.uint64)
        if n:
            self.L.p3o_count_short_kmers(seq, off, n_reads, keys.ctypes.data, counts.ctypes.data)
        return keys, counts

    def rmq(self, v, x):
        v = np.ascontiguousarray(v, np.uint64)
        out = np.zeros(len(v) + 1, np.uint64)
        m = self.L.p3o_rmq(v, len(v), x, out)
        return out[:m]

    def make_bf(self, seq, off, k, keys, counts, filter_size, num_hashes, want_solid=False):
        n_reads = len(off) - 1
        bloom = np.zeros((filter_size + 7) // 8, np.uint8)
        seeds = np.zeros(max(n_reads, 1), np.int64)
        solid = np.zeros(max(len(seq), 1), np.uint8) if want_solid else None
        adds = self.L.p3o_make_bf(seq, off, n_reads, k, keys, counts, len(keys), filter_size,
                                  num_hashes, bloom, seeds, solid.ctypes.data if want_solid else None)
        return bloom, seeds[:n_reads], (solid[: len(seq)] if want_solid else None), adds

    def solid_kmers(self, seq, off, k, keys, counts):
        n_reads = len(off) - 1
        n = self.L.p3o_solid_kmers(seq, off, n_reads, k, keys, counts, len(keys), None)
        out = np.zeros((n, nwords(k)), np.uint64)
        if n:
            self.L.p3o_solid_kmers(seq, off, n_reads, k, keys, counts, len(keys), out.ctypes.data)
        return out

    def check_directions(self, bloom, filter_size, num_hashes, words, k, ignored=-1):
        return self.L.p3o_check_directions(bloom, filter_size, num_hashes,
                                           np.ascontiguousarray(words, np.uint64), k, ignored)

    def bf_add(self, bloom, filter_size, num_hashes, words, k):
        self.L.p3o_bf_add(bloom, filter_size, num_hashes, np.ascontiguousarray(words, np.uint64), k)

    def is_recorded(self, bloom, filter_size, num_hashes, words, k):
        return self.L.p3o_is_recorded(bloom, filter_size, num_hashes, np.ascontiguousarray(words, np.uint64), k)


def have_ref():
    return os.path.exists(REF_SO)


class Ref:
    """One run of the unmodified reference (oracle/_ref/libp3ref.so)."""

    def __init__(self, k, readfile=None, m=0, threads=1, log_path=None):
        L = self.L = C.CDLL(REF_SO)
        L.p3ref_new.restype = C.c_void_p
        L.p3ref_new.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.c_int, C.c_char_p]
        for name in ("p3ref_free", "p3ref_load_file", "p3ref_estimate", "p3ref_count_short",
                     "p3ref_make_bf", "p3ref_make_dbg", "p3ref_count_node_coverage", "p3ref_print_graph"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        for name in ("p3ref_all_bases", "p3ref_n_reads", "p3ref_filter_size", "p3ref_short_size",
                     "p3ref_bf_size", "p3ref_seed_count", "p3ref_n_junctions", "p3ref_n_joints",
                     "p3ref_n_straights"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_uint64
        L.p3ref_num_hashes.argtypes = [C.c_void_p]
        L.p3ref_num_hashes.restype = C.c_int
        L.p3ref_add_read.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint64]
        L.p3ref_add_reads.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64]
        L.p3ref_check_directions_batch.argtypes = [C.c_void_p, u8p, C.c_uint64, u8p]
        L.p3ref_reads_export.restype = C.c_uint64
        L.p3ref_reads_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.p3ref_estimate_only.argtypes = [C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.p3ref_set_filter.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.p3ref_short_export.argtypes = [C.c_void_p, u64p, u64p]
        L.p3ref_bf_bits.argtypes = [C.c_void_p, u8p]
        L.p3ref_seed_export.argtypes = [C.c_void_p, C.c_char_p]
        L.p3ref_bf_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.p3ref_bf_add.argtypes = [C.c_void_p, C.c_char_p]
        L.p3ref_check_directions.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.p3ref_is_recorded.argtypes = [C.c_void_p, C.c_char_p]
        L.p3ref_std_hash.restype = C.c_uint64
        L.p3ref_std_hash.argtypes = [C.c_void_p, C.c_char_p]
        L.p3ref_double_hash.argtypes = [C.c_void_p, C.c_char_p, u64p]
        L.p3ref_canonical.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        self.k = k
        self.h = L.p3ref_new(readfile.encode() if readfile else None, k, m, threads,
                             log_path.encode() if log_path else None)
        if not self.h:
            raise ValueError("reference build has no instantiation for k=%d" % k)

    def close(self):
        if self.h:
            self.L.p3ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # stages, in main.cpp / Assemble.cpp order
    def load_file(self):
        self.L.p3ref_load_file(self.h)

    def add_reads(self, reads):
        for i, r in enumerate(reads):
            b = r.encode() if isinstance(r, str) else bytes(r)
            self.L.p3ref_add_read(self.h, (">r%d" % i).encode(), b, len(b))

    def add_reads_arrays(self, seq, off):
        self.L.p3ref_add_reads(self.h, np.ascontiguousarray(seq, np.uint8), np.ascontiguousarray(off, np.uint64), len(off) - 1)

    def check_directions_batch(self, kmers_ascii, n):
        out = np.zeros(max(n, 1), np.uint8)
        self.L.p3ref_check_directions_batch(self.h, np.ascontiguousarray(kmers_ascii, np.uint8), n, out)
        return out[:n]

    def reads(self):
        n = self.n_reads
        lens = np.zeros(max(n, 1), np.uint64)
        tot = self.L.p3ref_reads_export(self.h, lens.ctypes.data, None)
        seq = np.zeros(max(tot, 1), np.uint8)
        self.L.p3ref_reads_export(self.h, lens.ctypes.data, seq.ctypes.data)
        out, o = [], 0
        for i in range(n):
            out.append(seq[o:o + int(lens[i])].tobytes())
            o += int(lens[i])
        return out

    def estimate(self):
        self.L.p3ref_estimate(self.h)

    def estimate_only(self, all_bases, k):
        fs, nh = C.c_uint64(), C.c_int()
        self.L.p3ref_estimate_only(all_bases, k, C.byref(fs), C.byref(nh))
        return fs.value, nh.value

    def set_filter(self, m, nh):
        self.L.p3ref_set_filter(self.h, m, nh)

    all_bases = property(lambda s: s.L.p3ref_all_bases(s.h))
    n_reads = property(lambda s: s.L.p3ref_n_reads(s.h))
    filter_size = property(lambda s: s.L.p3ref_filter_size(s.h))
    num_hashes = property(lambda s: s.L.p3ref_num_hashes(s.h))

    def count_short(self):
        self.L.p3ref_count_short(self.h)
        n = self.L.p3ref_short_size(self.h)
        keys = np.zeros(n, np.uint64)
        counts = np.zeros(n, np.uint64)
        if n:
            self.L.p3ref_short_export(self.h, keys, counts)
        return keys, counts

    def make_bf(self):
        self.L.p3ref_make_bf(self.h)
        n = self.L.p3ref_bf_size(self.h)
        bits = np.zeros((n + 7) // 8, np.uint8)
        if n:
            self.L.p3ref_bf_bits(self.h, bits)
        ns = self.L.p3ref_seed_count(self.h)
        buf = C.create_string_buffer(ns * self.k + 1)
        self.L.p3ref_seed_export(self.h, buf)
        raw = buf.raw[: ns * self.k]
        seeds = [raw[i * self.k:(i + 1) * self.k].decode() for i in range(ns)]
        return bits, seeds

    def bf_reset(self, size, nh):
        self.L.p3ref_bf_reset(self.h, size, nh)

    def bf_add(self, kmer):
        self.L.p3ref_bf_add(self.h, kmer.encode())

    def bf_bits(self):
        n = self.L.p3ref_bf_size(self.h)
        bits = np.zeros((n + 7) // 8, np.uint8)
        self.L.p3ref_bf_bits(self.h, bits)
        return bits

    def check_directions(self, kmer, ignored=-1):
        return self.L.p3ref_check_directions(self.h, kmer.encode(), ignored)

    def is_recorded(self, kmer):
        return self.L.p3ref_is_recorded(self.h, kmer.encode())

    def std_hash(self, kmer):
        return self.L.p3ref_std_hash(self.h, kmer.encode())

    def double_hash(self, kmer):
        out = np.zeros(2, np.uint64)
        self.L.p3ref_double_hash(self.h, kmer.encode(), out)
        return int(out[0]), int(out[1])

    def canonical(self, kmer):
        buf = C.create_string_buffer(self.k + 1)
        self.L.p3ref_canonical(self.h, kmer.encode(), buf)
        return buf.raw[: self.k].decode()

    def make_dbg(self):
        self.L.p3ref_make_dbg(self.h)

    def count_node_coverage(self):
        self.L.p3ref_count_node_coverage(self.h)

    def print_graph(self, workdir):
        """PrintGraph writes ./de_bruijn_graph.gfa (DeBruijnGraph.cpp:454) -> run in workdir."""
        cwd = os.getcwd()
        os.chdir(workdir)
        try:
            self.L.p3ref_print_graph(self.h)
        finally:
            os.chdir(cwd)
        with open(os.path.join(workdir, "de_bruijn_graph.gfa")) as f:
            return f.read().splitlines()

    def counts(self):
        return (self.L.p3ref_n_junctions(self.h), self.L.p3ref_n_joints(self.h),
                self.L.p3ref_n_straights(self.h))
