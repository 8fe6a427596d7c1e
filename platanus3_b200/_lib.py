"""ctypes binding of include/platanus3_b200.h (libplatanus3_b200.so, built in-tree by
__graft_entry__.build()). No fallback: a missing library or a missing GPU raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplatanus3_b200.so")

P3_OK = 0
ERR_NAMES = {-1: "P3_ERR_CUDA", -2: "P3_ERR_ARG", -3: "P3_ERR_TABLE_FULL", -4: "P3_ERR_STATE",
             -5: "P3_ERR_NOMEM", -6: "P3_ERR_IO"}

# every symbol include/platanus3_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "p3_last_error", "p3_version", "p3_device_count", "p3_estimate_bloomfilter", "p3_packed_words",
    "p3_pack_reads", "p3_host_alloc", "p3_host_free", "p3_create", "p3_destroy", "p3_synchronize",
    "p3_reads_upload", "p3_reads_attach", "p3_count_short_kmers", "p3_short_kmer_stats",
    "p3_short_kmer_export", "p3_short_kmer_lookup", "p3_make_bf", "p3_make_bf_stats", "p3_bf_export",
    "p3_bf_import", "p3_seed_export", "p3_solid_flags_export", "p3_bf_add", "p3_bf_possibly_contains",
    "p3_double_hash", "p3_dbg_adjacency", "p3_dbg_stats", "p3_dbg_close", "p3_dbg_export", "p3_check_directions",
    "p3_owner_of_key", "p3_ipc_export", "p3_ipc_open", "p3_ipc_close", "p3_mg_arena", "p3_mg_connect", "p3_mg_sync",
    "p3_mg_staged_buffers", "p3_mg_count_begin", "p3_mg_count_send", "p3_mg_count_recv", "p3_mg_count_finish",
    "p3_mg_count_end", "p3_mg_cover_begin", "p3_mg_cover_send", "p3_mg_cover_recv", "p3_mg_solid_begin",
    "p3_mg_solid_send", "p3_mg_solid_recv", "p3_mg_solid_finish", "p3_mg_solid_end", "p3_bloom_seg_bits",
    "p3_mg_bloom_buffer", "p3_mg_bloom_bin", "p3_mg_bloom_apply", "p3_mg_bloom_direct", "p3_mg_filter",
    "p3_mg_makebf_done", "p3_device_mem_used",
    "p3_mg_count_next_round", "p3_mg_cover_rebin_begin", "p3_mg_cover_rebin_end", "p3_mg_solid_next_round", "p3_mg_bloom_bin_range",
    "p3_mg_count_begin_keyed", "p3_mg_key_round_begin", "p3_mg_cover_begin_keyed", "p3_mg_cover_key_round",
    "p3_mg_long_solid", "p3_mg_long_begin", "p3_mg_long_send", "p3_mg_long_recv", "p3_mg_long_finish",
    "p3_load_file", "p3_reads_free", "p3_reads_count", "p3_reads_all_bases", "p3_reads_total_bases",
    "p3_reads_offsets", "p3_reads_packed", "p3_reads_nmask", "p3_reads_ascii", "p3_assemble_file", "p3_walk_table", "p3_node_coverage",
    "p3_assemble_hot_path", "p3_assemble_hot_path_to_host", "p3_stage_ms", "p3_count_substage_ms", "p3_launch_count", "p3_bf_params", "p3_probe_stats", "p3_table_capacity", "p3_table_partitions",
]


class P3Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.p3_last_error.restype = C.c_char_p
        L.p3_packed_words.restype = u64
        L.p3_packed_words.argtypes = [u64]
        L.p3_estimate_bloomfilter.argtypes = [u64, u32, C.POINTER(u64), C.POINTER(u32)]
        L.p3_pack_reads.argtypes = [vp, vp, u64, vp, vp, C.POINTER(i32)]
        L.p3_host_alloc.restype = vp
        L.p3_host_alloc.argtypes = [C.c_size_t]
        L.p3_host_free.argtypes = [vp]
        L.p3_create.restype = vp
        L.p3_create.argtypes = [i32, vp]
        L.p3_destroy.argtypes = [vp]
        L.p3_synchronize.argtypes = [vp]
        L.p3_reads_upload.argtypes = [vp, vp, u64, vp, u64, vp]
        L.p3_reads_attach.argtypes = [vp, vp, u64, vp, u64, vp]
        L.p3_count_short_kmers.argtypes = [vp, u64]
        L.p3_short_kmer_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
        L.p3_short_kmer_export.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
        L.p3_short_kmer_lookup.argtypes = [vp, vp, u64, vp]
        L.p3_make_bf.argtypes = [vp, u32, u64, u32, u32, u64]
        L.p3_make_bf_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
        L.p3_bf_export.argtypes = [vp, vp]
        L.p3_bf_import.argtypes = [vp, u32, u64, u32, vp]
        L.p3_seed_export.argtypes = [vp, vp]
        L.p3_solid_flags_export.argtypes = [vp, vp]
        L.p3_bf_add.argtypes = [vp, vp, u64]
        L.p3_bf_possibly_contains.argtypes = [vp, vp, u64, vp]
        L.p3_double_hash.argtypes = [vp, u32, vp, u64, vp]
        L.p3_dbg_adjacency.argtypes = [vp]
        L.p3_dbg_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
        L.p3_dbg_close.argtypes = [vp, vp, u64, C.POINTER(u64)]
        L.p3_dbg_export.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
        L.p3_check_directions.argtypes = [vp, vp, u64, vp]
        L.p3_assemble_hot_path.argtypes = [vp, vp, u64, vp, u64, vp, u64, u32, u64, u32, u64, u64]
        L.p3_assemble_hot_path_to_host.argtypes = [vp, vp, u64, vp, u64, vp, u64, u32, u64, u32, u64, u64, vp, vp, vp, vp, u64, C.POINTER(u64)]
        L.p3_owner_of_key.restype = u32
        L.p3_owner_of_key.argtypes = [u64, u32]
        L.p3_ipc_export.argtypes = [vp, vp]
        L.p3_ipc_open.argtypes = [C.c_int, vp, C.POINTER(vp)]
        L.p3_ipc_close.argtypes = [C.c_int, vp]
        L.p3_mg_arena.argtypes = [vp, u32, u32, u64, i32, C.POINTER(vp)]
        L.p3_mg_connect.argtypes = [vp, vp, i32]
        L.p3_mg_sync.argtypes = [vp]
        L.p3_mg_staged_buffers.argtypes = [vp, i32, i32, vp]
        L.p3_mg_count_begin.argtypes = [vp, u64, u64, u64, u64]
        L.p3_mg_count_send.argtypes = [vp, u64]
        L.p3_mg_count_recv.argtypes = [vp, u64]
        L.p3_mg_count_finish.argtypes = [vp]
        L.p3_mg_count_end.argtypes = [vp]
        L.p3_mg_cover_begin.argtypes = [vp, u32, u64, C.POINTER(u32)]
        L.p3_mg_cover_send.argtypes = [vp, u32, u32]
        L.p3_mg_cover_recv.argtypes = [vp, u32]
        L.p3_mg_solid_begin.argtypes = [vp, u32, u64]
        L.p3_mg_solid_send.argtypes = [vp, u64]
        L.p3_mg_solid_recv.argtypes = [vp, u64]
        L.p3_mg_solid_finish.argtypes = [vp]
        L.p3_mg_solid_end.argtypes = [vp, u64, u32, u64, C.POINTER(u64), C.POINTER(u64)]
        L.p3_mg_filter.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
        L.p3_bloom_seg_bits.restype = u64
        L.p3_bloom_seg_bits.argtypes = []
        L.p3_mg_bloom_buffer.argtypes = [vp, u64, C.POINTER(vp)]
        L.p3_mg_bloom_bin.argtypes = [vp, u32, vp, u64, vp]
        L.p3_mg_bloom_bin_range.argtypes = [vp, u32, vp, u64, vp, u64, u64]
        L.p3_table_partitions.restype = u32
        L.p3_table_partitions.argtypes = [u64]
        L.p3_mg_count_begin_keyed.argtypes = [vp, u64, u64, u64, u64, u32]
        L.p3_mg_key_round_begin.argtypes = [vp, u32]
        L.p3_mg_cover_begin_keyed.argtypes = [vp, u32, u64, C.POINTER(u32)]
        L.p3_mg_cover_key_round.argtypes = [vp, u32]
        L.p3_mg_long_solid.argtypes = [vp, u32, C.POINTER(u64)]
        L.p3_mg_long_begin.argtypes = [vp, u64, u64, C.c_double, u64, C.POINTER(u64), C.POINTER(u64)]
        L.p3_mg_long_send.argtypes = [vp, u64]
        L.p3_mg_long_recv.argtypes = [vp, u64]
        L.p3_mg_long_finish.argtypes = [vp]
        for fn in (L.p3_mg_count_next_round, L.p3_mg_cover_rebin_begin, L.p3_mg_cover_rebin_end, L.p3_mg_solid_next_round):
            fn.argtypes = [vp]
        L.p3_mg_bloom_apply.argtypes = [vp, u64, u32, u32, vp, vp]
        L.p3_mg_bloom_direct.argtypes = [vp]
        L.p3_mg_makebf_done.argtypes = [vp]
        L.p3_device_mem_used.restype = u64
        L.p3_device_mem_used.argtypes = [vp]
        L.p3_load_file.argtypes = [C.c_char_p, u32, C.POINTER(vp)]
        L.p3_reads_free.argtypes = [vp]
        for nm in ("p3_reads_count", "p3_reads_all_bases", "p3_reads_total_bases"):
            getattr(L, nm).restype = u64
            getattr(L, nm).argtypes = [vp]
        for nm in ("p3_reads_offsets", "p3_reads_packed", "p3_reads_nmask", "p3_reads_ascii"):
            getattr(L, nm).restype = vp
            getattr(L, nm).argtypes = [vp]
        L.p3_node_coverage.argtypes = [vp, u32, vp, u64, vp, u64, vp, vp]
        L.p3_assemble_file.argtypes = [C.c_char_p, u32, u64, i32, i32, C.c_char_p, C.c_char_p, C.POINTER(u64)]
        L.p3_walk_table.argtypes = [C.c_char_p, u32, vp, vp, u64, vp, u64, C.c_char_p, C.POINTER(u64)]
        L.p3_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.p3_count_substage_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(u32), C.POINTER(u64)]
        L.p3_probe_stats.argtypes = [vp, vp]
        L.p3_table_capacity.argtypes = [vp, vp]
        L.p3_launch_count.restype = u64
        L.p3_launch_count.argtypes = [vp]
        L.p3_bf_params.argtypes = [vp, C.POINTER(u64), C.POINTER(u32), C.POINTER(u32)]
        _lib = L
    return _lib


def check(rc):
    if rc != P3_OK:
        raise P3Error(rc, lib().p3_last_error().decode())


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


def estimate_bloomfilter(all_bases, k):
    """Options::EstimateBloomfilter (reference src/Options.cpp:50)"""
    fs, nh = C.c_uint64(), C.c_uint32()
    check(lib().p3_estimate_bloomfilter(all_bases, k, C.byref(fs), C.byref(nh)))
    return fs.value, nh.value


def pack_reads(seq, off, want_mask=True):
    """ASCII reads (uint8 array + uint64 offsets) -> (packed uint64, nmask uint32 or None)"""
    seq = np.ascontiguousarray(seq, np.uint8)
    off = np.ascontiguousarray(off, np.uint64)
    n_reads = len(off) - 1
    total = int(off[-1]) if n_reads > 0 else 0
    words = lib().p3_packed_words(total)
    packed = np.zeros(words, np.uint64)
    nmask = np.zeros(words, np.uint32)
    bad = C.c_int(0)
    check(lib().p3_pack_reads(_ptr(seq) if total else None, _ptr(off), n_reads, _ptr(packed), _ptr(nmask), C.byref(bad)))
    return packed, (nmask if (bad.value and want_mask) else None)


def load_file(path, k):
    """ReadFile::LoadFile (reference src/Load.cpp:32) -> dict(seq uint8, off uint64, all_bases, packed, nmask)"""
    h = C.c_void_p()
    check(lib().p3_load_file(path.encode(), k, C.byref(h)))
    L = lib()
    try:
        n = L.p3_reads_count(h)
        total = L.p3_reads_total_bases(h)
        off = np.ctypeslib.as_array(C.cast(L.p3_reads_offsets(h), C.POINTER(C.c_uint64)), (n + 1,)).copy()
        seq = np.ctypeslib.as_array(C.cast(L.p3_reads_ascii(h), C.POINTER(C.c_uint8)), (max(total, 1),))[:total].copy()
        words = L.p3_packed_words(total)
        packed = np.ctypeslib.as_array(C.cast(L.p3_reads_packed(h), C.POINTER(C.c_uint64)), (words,)).copy()
        nm = L.p3_reads_nmask(h)
        nmask = np.ctypeslib.as_array(C.cast(nm, C.POINTER(C.c_uint32)), (words,)).copy() if nm else None
        return dict(seq=seq, off=off, all_bases=L.p3_reads_all_bases(h), packed=packed, nmask=nmask)
    finally:
        L.p3_reads_free(h)


def walk_table(path, k, kmers, adj, seeds, gfa_path=None):
    """host half of the drop-in over a closed CheckDirections table (no GPU needed)"""
    kmers = np.ascontiguousarray(kmers, np.uint64)
    adj = np.ascontiguousarray(adj, np.uint8)
    seeds = np.ascontiguousarray(seeds, np.uint64)
    W = (2 * k + 63) // 64
    st = (C.c_uint64 * 3)()
    check(lib().p3_walk_table(path.encode(), k, _ptr(kmers), _ptr(adj), len(adj), _ptr(seeds), seeds.size // W,
                              gfa_path.encode() if gfa_path else None, st))
    return dict(junctions=int(st[0]), joints=int(st[1]), straights=int(st[2]))


def assemble_file(path, k, m=0, threads=1, device=0, gfa_path=None, log_path=None):
    """main() + Assemble<> for one read file; returns the 8 run statistics"""
    st = (C.c_uint64 * 8)()
    check(lib().p3_assemble_file(path.encode(), k, m, threads, device,
                                 gfa_path.encode() if gfa_path else None,
                                 log_path.encode() if log_path else None, st))
    names = ("reads", "all_bases", "distinct_21mers", "solid_kmers", "table_kmers", "junctions", "joints", "straights")
    return dict(zip(names, [int(x) for x in st]))


class Context:
    """One GPU context (p3_ctx). Mirrors the order of reference src/Assemble.cpp:7-21."""

    def __init__(self, device=0, stream=None):
        self.L = lib()
        self.h = self.L.p3_create(device, stream)
        if not self.h:
            raise P3Error(-1, self.L.p3_last_error().decode())
        self.n_reads = 0
        self.total_bases = 0
        self._keep = None

    def close(self):
        if self.h:
            self.L.p3_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # reads -------------------------------------------------------------------
    def upload(self, packed, off, nmask=None):
        off = np.ascontiguousarray(off, np.uint64)
        self.n_reads = len(off) - 1
        self.total_bases = int(off[-1]) if self.n_reads > 0 else 0
        check(self.L.p3_reads_upload(self.h, _ptr(packed), self.total_bases, _ptr(off), self.n_reads, _ptr(nmask)))
        self.L.p3_synchronize(self.h)

    def attach(self, d_packed_ptr, total_bases, d_off_ptr, n_reads, d_nmask_ptr=None, keep=None):
        self.n_reads, self.total_bases, self._keep = n_reads, total_bases, keep
        check(self.L.p3_reads_attach(self.h, d_packed_ptr, total_bases, d_off_ptr, n_reads, d_nmask_ptr))

    def load_ascii(self, seq, off):
        packed, nmask = pack_reads(seq, off)
        self.upload(packed, off, nmask)

    # stage A: ReadFile::CountShortKmer ------------------------------------------
    def count_short_kmers(self, table_slots=0):
        check(self.L.p3_count_short_kmers(self.h, table_slots))
        a, b = C.c_uint64(), C.c_uint64()
        check(self.L.p3_short_kmer_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def short_kmer_export(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(self.L.p3_short_kmer_stats(self.h, C.byref(a), C.byref(b)))
        keys = np.zeros(max(b.value, 1), np.uint64)
        counts = np.zeros(max(b.value, 1), np.uint64)
        n = C.c_uint64()
        check(self.L.p3_short_kmer_export(self.h, _ptr(keys), _ptr(counts), len(keys), C.byref(n)))
        keys, counts = keys[: n.value], counts[: n.value]
        o = np.argsort(keys, kind="stable")
        return keys[o], counts[o]

    def short_kmer_lookup(self, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        out = np.zeros(len(keys), np.uint64)
        check(self.L.p3_short_kmer_lookup(self.h, _ptr(keys), len(keys), _ptr(out)))
        return out

    # stage B: MakeBF ---------------------------------------------------------------
    def make_bf(self, k, filter_size, num_hashes, cov_threshold=2, solid_slots=0):
        check(self.L.p3_make_bf(self.h, k, filter_size, num_hashes, cov_threshold, solid_slots))
        a, b = C.c_uint64(), C.c_uint64()
        check(self.L.p3_make_bf_stats(self.h, C.byref(a), C.byref(b)))
        self.k, self.filter_size, self.num_hashes = k, filter_size, num_hashes
        return a.value, b.value

    def bf_export(self):
        bits = np.zeros((self.filter_size + 7) // 8, np.uint8)
        check(self.L.p3_bf_export(self.h, _ptr(bits)))
        return bits

    def bf_import(self, k, filter_size, num_hashes, bits=None):
        check(self.L.p3_bf_import(self.h, k, filter_size, num_hashes, _ptr(bits)))
        self.k, self.filter_size, self.num_hashes = k, filter_size, num_hashes

    def seed_export(self):
        s = np.zeros(max(self.n_reads, 1), np.int64)
        check(self.L.p3_seed_export(self.h, _ptr(s)))
        return s[: self.n_reads]

    def solid_flags_export(self):
        words = (self.total_bases + 31) // 32
        bm = np.zeros(max(words, 1), np.uint32)
        check(self.L.p3_solid_flags_export(self.h, _ptr(bm)))
        # -> one byte per stream position
        bits = np.unpackbits(bm[:words].astype(">u4").view(np.uint8))
        return bits[: self.total_bases]

    def bf_add(self, kmers):
        """kmers: uint64[n] (k <= 32) or uint64[n, W]"""
        kmers = np.ascontiguousarray(kmers, np.uint64)
        check(self.L.p3_bf_add(self.h, _ptr(kmers), len(kmers)))

    def bf_possibly_contains(self, kmers):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        out = np.zeros(len(kmers), np.uint8)
        check(self.L.p3_bf_possibly_contains(self.h, _ptr(kmers), len(kmers), _ptr(out)))
        return out

    def double_hash(self, k, kmers):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        out = np.zeros((len(kmers), 2), np.uint64)
        check(self.L.p3_double_hash(self.h, k, _ptr(kmers), len(kmers), _ptr(out)))
        return out

    # stage C: DeBruijnGraph::CheckDirections -----------------------------------------
    def dbg_adjacency(self):
        check(self.L.p3_dbg_adjacency(self.h))
        a, b = C.c_uint64(), C.c_uint64()
        check(self.L.p3_dbg_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def dbg_close(self, roots=None):
        n = C.c_uint64()
        roots = None if roots is None else np.ascontiguousarray(roots, np.uint64)
        check(self.L.p3_dbg_close(self.h, _ptr(roots), 0 if roots is None else len(roots), C.byref(n)))
        return n.value

    def dbg_export(self, sort=True):
        """distinct k-mers and adjacency bytes; k <= 32: kmers is uint64[n], else uint64[n, W] (little-endian words)"""
        n = C.c_uint64()
        self.L.p3_dbg_export(self.h, None, None, 0, C.byref(n))   # size query (capacity error ignored)
        W = (2 * self.k + 63) // 64
        if W > 1:
            kmers = np.zeros((max(n.value, 1), W), np.uint64)
            adj = np.zeros(max(n.value, 1), np.uint8)
            check(self.L.p3_dbg_export(self.h, _ptr(kmers), _ptr(adj), len(adj), C.byref(n)))
            kmers, adj = kmers[: n.value], adj[: n.value]
            if sort and len(kmers):
                o = np.lexsort(tuple(kmers[:, j] for j in range(W)))   # last key (top word) is primary
                kmers, adj = kmers[o], adj[o]
            return kmers, adj
        kmers = np.zeros(max(n.value, 1), np.uint64)
        adj = np.zeros(max(n.value, 1), np.uint8)
        check(self.L.p3_dbg_export(self.h, _ptr(kmers), _ptr(adj), len(kmers), C.byref(n)))
        kmers, adj = kmers[: n.value], adj[: n.value]
        if sort:
            o = np.argsort(kmers, kind="stable")
            kmers, adj = kmers[o], adj[o]
        return kmers, adj

    def check_directions(self, kmers):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        out = np.zeros(len(kmers), np.uint8)
        check(self.L.p3_check_directions(self.h, _ptr(kmers), len(kmers), _ptr(out)))
        return out

    # whole path ----------------------------------------------------------------------
    def assemble_hot_path(self, packed, off, k, nmask=None, all_bases=None, filter_size=0, num_hashes=10,
                          table_slots=0, solid_slots=0):
        off = np.ascontiguousarray(off, np.uint64) if not isinstance(off, int) else off
        if not isinstance(off, int):
            self.n_reads = len(off) - 1
            self.total_bases = int(off[-1]) if self.n_reads > 0 else 0
        if all_bases is None:
            all_bases = self.total_bases
        check(self.L.p3_assemble_hot_path(self.h, _ptr(packed), self.total_bases, _ptr(off), self.n_reads,
                                          _ptr(nmask), all_bases, k, filter_size, num_hashes, table_slots, solid_slots))
        fs, nh, kk = C.c_uint64(), C.c_uint32(), C.c_uint32()
        check(self.L.p3_bf_params(self.h, C.byref(fs), C.byref(nh), C.byref(kk)))
        self.k, self.filter_size, self.num_hashes = kk.value, fs.value, nh.value

    def stage_ms(self):
        ms = (C.c_float * 5)()
        check(self.L.p3_stage_ms(self.h, ms))
        return dict(zip(("count21", "flags21", "makebf", "seeds", "adjacency"), [float(x) for x in ms]))

    def count_substage_ms(self):
        ms = (C.c_float * 4)()
        parts, chunks = C.c_uint32(), C.c_uint64()
        check(self.L.p3_count_substage_ms(self.h, ms, C.byref(parts), C.byref(chunks)))
        return {"hist": float(ms[0]), "scatter": float(ms[1]), "insert": float(ms[2]), "bloom_add": float(ms[3]),
                "parts": parts.value, "chunks": chunks.value}

    def launch_count(self):
        return int(self.L.p3_launch_count(self.h))

    def stats(self):
        a, b, c, d, e, f = (C.c_uint64() for _ in range(6))
        out = {}
        if self.L.p3_short_kmer_stats(self.h, C.byref(a), C.byref(b)) == 0:
            out.update(n_positions=a.value, n_distinct21=b.value)
        if self.L.p3_make_bf_stats(self.h, C.byref(c), C.byref(d)) == 0:
            out.update(n_adds=c.value, n_distinct_solid=d.value)
        if self.L.p3_dbg_stats(self.h, C.byref(e), C.byref(f)) == 0:
            out.update(n_kmers=e.value, n_edges=f.value)
        return out
