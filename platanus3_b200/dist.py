"""Multi-GPU hot path: canonical k-mers hash-partitioned by owner rank, exchanged with all-to-all.

One process per GPU (torch.distributed / NCCL). All heavy work is in the CUDA library
(csrc/p3_multi.inc.cu); this module only sequences the stages and moves the device buffers:

  A   every rank bins its 21-mers by owner -> all-to-all (12 B records) -> owner counts them
  B1  owner lists its count-1 keys' (rank, position) -> all-to-all -> ranks clear coverage bits
  B2  ranks build solid planes/seeds and their locally distinct solid k-mers -> all-to-all by
      owner -> owner de-duplicates and BF.adds into its filter copy -> OR-reduce of the copies
  C   owner runs CheckDirections for its k-mers against the (now complete, local) filter

The same driver runs over an emulated communicator (several contexts of one process on one GPU),
which is how the parity tests exercise the distributed algorithm on a single-GPU box.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class _DevView:
    """torch view of a raw device pointer via __cuda_array_interface__"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def dev_tensor(ptr, n, dtype, device):
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device=device)
    typestr = {torch.int64: "<i8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_DevView(ptr, n, typestr), device=device)


# ---------------------------------------------------------------------------- communicators
class TorchDistComm:
    """all-to-all with variable splits and bitwise-OR reduction over torch.distributed"""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.local_ranks = [self.rank]

    def exchange(self, sends):
        """sends: [(tensors, counts)] for the one local rank -> [(recv_tensors, recv_counts)]"""
        (tensors, counts), = sends
        dev = tensors[0].device
        sc = torch.tensor(counts, dtype=torch.int64, device=dev)
        rc = torch.empty_like(sc)
        self.dist.all_to_all_single(rc, sc, group=self.group)
        rcounts = [int(x) for x in rc.tolist()]
        outs = []
        for t in tensors:
            out = torch.empty(sum(rcounts), dtype=t.dtype, device=dev)
            self.dist.all_to_all_single(out, t, output_split_sizes=rcounts, input_split_sizes=[int(c) for c in counts], group=self.group)
            outs.append(out)
        return [(outs, rcounts)]

    def or_reduce(self, filters):
        """bitwise OR of the int32 filter copies of all ranks, in place: all-to-all of shards, local OR,
        all-gather (NCCL has no bitwise-or reduction op)"""
        f, = filters
        n, w = f.numel(), self.world
        if w == 1:
            return
        shard = (n + w - 1) // w
        buf = torch.zeros(shard * w, dtype=f.dtype, device=f.device)
        buf[:n] = f
        recv = torch.empty_like(buf)
        self.dist.all_to_all_single(recv, buf, group=self.group)
        acc = recv[:shard].clone()
        for i in range(1, w):
            torch.bitwise_or(acc, recv[i * shard:(i + 1) * shard], out=acc)
        self.dist.all_gather_into_tensor(buf, acc, group=self.group)
        f.copy_(buf[:n])

    def _dev(self):
        return "cuda" if torch.cuda.is_available() and self.dist.get_backend(self.group) == "nccl" else "cpu"

    def all_gather_shards(self, filters, shard):
        """filters: the one local filter tensor of world*shard words whose shard `rank` is final ->
        every shard final on every rank (in-place all-gather)"""
        f, = filters
        if self.world > 1:
            self.dist.all_gather_into_tensor(f[: self.world * shard], f[self.rank * shard:(self.rank + 1) * shard], group=self.group)

    def all_sum(self, values):
        t = torch.tensor(values, dtype=torch.int64, device=self._dev())
        self.dist.all_reduce(t, group=self.group)
        return [int(x) for x in t.tolist()]

    def all_gather(self, rows):
        """rows: [list of ints] of the one local rank (same length on every rank) -> [world][len]"""
        (row,) = rows
        t = torch.tensor(row, dtype=torch.int64, device=self._dev())
        out = torch.empty(self.world * max(t.numel(), 1), dtype=torch.int64, device=t.device)[: self.world * t.numel()]
        if t.numel():
            self.dist.all_gather_into_tensor(out, t, group=self.group)
        return out.view(self.world, -1).tolist()

    def barrier(self):
        torch.cuda.synchronize()
        self.dist.barrier(group=self.group)

    def share(self, ptr_rows, device_index):
        """ptr_rows: [[device pointers of cudaMalloc'ed buffers]] of the one local rank -> the same
        buffers of EVERY rank as pointers valid in this process ([world][n]): CUDA IPC handles are
        all-gathered and opened once (peer access over NVLink); the mappings are cached by handle."""
        L = _lib.lib()
        (ptrs,) = ptr_rows
        handles = []
        for p in ptrs:
            h = (C.c_uint8 * 64)()
            _check(L.p3_ipc_export(C.c_void_p(p), h))
            handles.append(bytes(h))
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, handles, group=self.group)
        if not hasattr(self, "_mapped"):
            self._mapped = {}
        table = []
        for r in range(self.world):
            if r == self.rank:
                table.append(list(ptrs))
                continue
            row = []
            for h in gathered[r]:
                if h not in self._mapped:
                    out = C.c_void_p()
                    _check(L.p3_ipc_open(device_index, (C.c_uint8 * 64).from_buffer_copy(h), C.byref(out)))
                    self._mapped[h] = (out.value, device_index)
                row.append(self._mapped[h][0])
            table.append(row)
        return table

    def close_shared(self):
        """unmap every peer buffer (call on all ranks BEFORE the owners free their buffers)"""
        L = _lib.lib()
        for ptr, dev in getattr(self, "_mapped", {}).values():
            L.p3_ipc_close(dev, C.c_void_p(ptr))
        self._mapped = {}


class EmulatedComm:
    """all ranks live in this process (one context each); exchanges are slicing and concatenation"""

    def __init__(self, world):
        self.world = world
        self.local_ranks = list(range(world))

    def exchange(self, sends):
        w = self.world
        offs = [np.concatenate([[0], np.cumsum(counts)]).astype(np.int64) for _, counts in sends]
        out = []
        for j in range(w):
            rcounts = [int(sends[i][1][j]) for i in range(w)]
            outs = []
            for ti in range(len(sends[0][0])):
                parts = [sends[i][0][ti][int(offs[i][j]):int(offs[i][j + 1])] for i in range(w)]
                outs.append(torch.cat(parts) if parts else sends[0][0][ti][:0])
            out.append((outs, rcounts))
        return out

    def or_reduce(self, filters):
        acc = filters[0].clone()
        for f in filters[1:]:
            torch.bitwise_or(acc, f, out=acc)
        for f in filters:
            f.copy_(acc)

    def all_gather_shards(self, filters, shard):
        for r, f in enumerate(filters):
            for o, g in enumerate(filters):
                if o != r:
                    g[r * shard:(r + 1) * shard] = f[r * shard:(r + 1) * shard]

    def all_sum(self, values_per_rank):
        return [int(sum(v)) for v in zip(*values_per_rank)]

    def all_gather(self, rows):
        return [list(r) for r in rows]

    def barrier(self):
        torch.cuda.synchronize()

    def share(self, ptr_rows, device_index):
        return [list(r) for r in ptr_rows]   # one process: every rank's pointers are valid as they are

    def close_shared(self):
        pass


# ---------------------------------------------------------------------------- driver
def _check(rc):
    _lib.check(rc)


def _exchange(comm, sends):
    """all-to-all, then wait for it: the library launches on its context's stream, which need not be
    the torch stream NCCL synchronises with (every p3_mg_* call itself returns synchronised)"""
    out = comm.exchange(sends)
    torch.cuda.synchronize()
    return out


def run_hot_path(ctxs, comm, k, filter_size, num_hashes, table_slots, solid_slots=0, owned_slots=0,
                 chunk_words=None, device=None):
    """ctxs: the Context of every LOCAL rank (reads already attached/uploaded), in comm.local_ranks
    order. table_slots: per-rank count-table capacity. Returns one stats dict per local rank; the
    results stay in the contexts (owned 21-mer counts, owned k-mers + adjacency, local seeds, the
    complete filter)."""
    L = _lib.lib()
    w = comm.world
    device = device or torch.device("cuda", torch.cuda.current_device())
    stats = [dict(rank=r) for r in comm.local_ranks]
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))
    mark("start")
    n_words = [(c.total_bases + 31) // 32 for c in ctxs]
    cw = chunk_words or max(max(n_words), 1)
    n_chunks = max((nw + cw - 1) // cw for nw in n_words) if n_words else 0
    if hasattr(comm, "dist"):   # ranks may hold different amounts of reads: agree on the chunk count
        t = torch.tensor([n_chunks], dtype=torch.int64, device=device)
        comm.dist.all_reduce(t, op=comm.dist.ReduceOp.MAX, group=comm.group)
        n_chunks = int(t.item())

    import os
    import time
    peer = os.environ.get("P3_MG_EXCHANGE", "peer") != "nccl" and w <= 16
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    u64a = C.c_uint64 * w
    sub = dict(bin=0.0, exchange=0.0, insert=0.0)
    # ---- A: count ------------------------------------------------------------------------------
    for c in ctxs:
        _check(L.p3_mg_count_begin(c.h, table_slots, 1))

    def chunk_range(ch, nw):
        return min(ch * cw, nw), min((ch + 1) * cw, nw)

    if peer:
        # Fused bin + exchange (csrc: scatter21_kernel<.., PEER>): every rank stores owner j's records
        # straight into rank j's receive buffer over NVLink, tile by tile, while it bins. Needs the
        # per-chunk owner histograms of all ranks first (where each source's region starts).
        t0 = time.perf_counter()
        rows = []
        for c, nw in zip(ctxs, n_words):
            row = []
            for ch in range(n_chunks):
                w0, w1 = chunk_range(ch, nw)
                counts = u64a()
                _check(L.p3_mg_owner_hist(c.h, w, w0, w1, counts))
                row += [int(x) for x in counts]
            rows.append(row)
        hist = np.array(comm.all_gather(rows), dtype=np.int64).reshape(w, n_chunks, w)   # [source][chunk][owner]
        cap = hist.sum(axis=0).max(axis=0) if n_chunks else np.zeros(w, np.int64)        # per owner
        n_buf = 2 if n_chunks > 1 else 1
        ptr_rows = []
        for c, r in zip(ctxs, comm.local_ranks):
            row = []
            for b in range(n_buf):
                pk, pw = C.c_void_p(), C.c_void_p()
                _check(L.p3_mg_recv_buffers(c.h, int(cap[r]), b, C.byref(pk), C.byref(pw)))
                row += [pk.value, pw.value]
            ptr_rows.append(row)
        table = comm.share(ptr_rows, dev_index)     # [rank][2*buffer + (0 = keys, 1 = words)]
        comm.barrier()
        sub["bin"] += 1e3 * (time.perf_counter() - t0)

        def scatter(ch, asynchronous):
            before = np.cumsum(hist[:, ch, :], axis=0) - hist[:, ch, :]   # [source][owner]: records of lower sources
            b = 2 * (ch % n_buf)
            for c, r, nw in zip(ctxs, comm.local_ranks, n_words):
                w0, w1 = chunk_range(ch, nw)
                kb = u64a(*[table[j][b] + 8 * int(before[r, j]) for j in range(w)])
                wb = u64a(*[table[j][b + 1] + 4 * int(before[r, j]) for j in range(w)])
                _check(L.p3_mg_owner_scatter_peer(c.h, w, r, w0, w1, kb, wb, 1 if asynchronous else 0))

        # software pipeline over the chunks: while the owners insert chunk ch (L2-latency bound), the
        # binning kernel of chunk ch+1 (ALU / NVLink bound) already stores into the other buffer set
        overlap = os.environ.get("P3_MG_OVERLAP", "0") != "0"
        t0 = time.perf_counter()
        if n_chunks:
            scatter(0, False)
        comm.barrier()              # every source's stores of chunk 0 have landed
        sub["bin"] += 1e3 * (time.perf_counter() - t0)
        for ch in range(n_chunks):
            t2 = time.perf_counter()
            if ch + 1 < n_chunks and overlap:
                scatter(ch + 1, True)
            b = 2 * (ch % n_buf)
            for c, r, row in zip(ctxs, comm.local_ranks, ptr_rows):
                n = int(hist[:, ch, r].sum())
                if n:
                    _check(L.p3_mg_count_records(c.h, row[b], row[b + 1], n))
            if ch + 1 < n_chunks and not overlap:
                scatter(ch + 1, False)
            for c in ctxs:
                _check(L.p3_mg_scatter_wait(c.h))
            comm.barrier()          # chunk ch+1 has landed everywhere; buffer set ch % 2 is free again
            sub["insert"] += 1e3 * (time.perf_counter() - t2)
    for ch in range(n_chunks if not peer else 0):
        t0 = time.perf_counter()
        sends = []
        for c, r, nw in zip(ctxs, comm.local_ranks, n_words):
            w0, w1 = chunk_range(ch, nw)
            counts = (C.c_uint64 * w)()
            _check(L.p3_mg_owner_hist(c.h, w, w0, w1, counts))
            counts = [int(x) for x in counts]
            tot = sum(counts)
            keys = torch.empty(max(tot, 1), dtype=torch.int64, device=device)
            words = torch.empty(max(tot, 1), dtype=torch.int32, device=device)
            _check(L.p3_mg_owner_scatter(c.h, w, r, w0, w1, keys.data_ptr(), words.data_ptr()))
            sends.append(([keys[:tot], words[:tot]], counts))
        t1 = time.perf_counter()
        recvs = _exchange(comm, sends)
        del sends
        t2 = time.perf_counter()
        for c, (tensors, rcounts) in zip(ctxs, recvs):
            n = sum(rcounts)
            if n:
                _check(L.p3_mg_count_records(c.h, tensors[0].data_ptr(), tensors[1].data_ptr(), n))
        del recvs
        t3 = time.perf_counter()
        sub["bin"] += 1e3 * (t1 - t0); sub["exchange"] += 1e3 * (t2 - t1); sub["insert"] += 1e3 * (t3 - t2)
    for c, st in zip(ctxs, stats):
        _check(L.p3_mg_count_end(c.h))
        a, b = C.c_uint64(), C.c_uint64()
        _check(L.p3_short_kmer_stats(c.h, C.byref(a), C.byref(b)))
        st.update(owned_positions=a.value, owned_distinct21=b.value, exchange="peer" if peer else "nccl")
        st["owner_count_ms"] = {kk: v for kk, v in c.count_substage_ms().items() if kk in ("hist", "scatter", "insert")}

    mark("count")
    laps = {}
    lap_t = [time.perf_counter()]

    def lap(name):      # wall-clock laps between library calls (each of which returns synchronised)
        torch.cuda.synchronize()
        now = time.perf_counter()
        laps[name] = laps.get(name, 0.0) + 1e3 * (now - lap_t[0])
        lap_t[0] = now
    # ---- B1: singleton verdicts back to the reads ----------------------------------------------------
    if peer and os.environ.get("P3_MG_COVER", "nccl") == "peer":
        # owners clear the bits of their count-1 keys directly in the source ranks' planes (NVLink RED.AND).
        # NOT the default: remote atomics are slow — equal to the all-to-all route at 2 GPUs (57 vs 60 ms)
        # but 2168 ms instead of 74 ms at 8 GPUs (profiles/r01_summary.md)
        ptr_rows = []
        for c in ctxs:
            _check(L.p3_mg_cover_begin(c.h))
            pp = C.c_void_p()
            _check(L.p3_mg_cover_plane(c.h, C.byref(pp)))
            ptr_rows.append([pp.value])
        planes = comm.share(ptr_rows, dev_index)
        comm.barrier()
        pl = u64a(*[planes[j][0] for j in range(w)])
        for c in ctxs:
            _check(L.p3_mg_cover_peer(c.h, w, pl))
        comm.barrier()
    else:
        sends = []
        for c in ctxs:
            counts = (C.c_uint64 * w)()
            ptr = C.c_void_p()
            _check(L.p3_mg_singletons(c.h, w, counts, C.byref(ptr)))
            counts = [int(x) for x in counts]
            sends.append(([dev_tensor(ptr.value, sum(counts), torch.int64, device)], counts))
        recvs = _exchange(comm, sends)
        for c, (tensors, rcounts) in zip(ctxs, recvs):
            _check(L.p3_mg_cover_begin(c.h))
            n = sum(rcounts)
            if n:
                t = tensors[0].contiguous()
                _check(L.p3_mg_cover_clear(c.h, t.data_ptr(), n))
        del sends, recvs

    lap("coverage")
    mark("coverage")
    # ---- B2: solid k-mers to their owners ---------------------------------------------------------------
    sends = []
    for c, st in zip(ctxs, stats):
        a, b = C.c_uint64(), C.c_uint64()
        _check(L.p3_mg_solid_local(c.h, k, solid_slots, C.byref(a), C.byref(b)))
        st.update(n_adds=a.value, local_distinct_solid=b.value)
        lap("solid_local")
        counts = (C.c_uint64 * w)()
        _check(L.p3_mg_kmer_owner_hist(c.h, w, counts))
        counts = [int(x) for x in counts]
        buf = torch.empty(max(sum(counts), 1), dtype=torch.int64, device=device)
        _check(L.p3_mg_kmer_owner_scatter(c.h, w, buf.data_ptr()))
        sends.append(([buf[:sum(counts)]], counts))
        lap("kmer_bin")
    recvs = _exchange(comm, sends)
    del sends
    lap("kmer_exchange")
    # the filter: sharded, binned adds (each rank owns a contiguous run of 16 MB segments and receives
    # the bit indices that fall into them) or, as fallback, adds into replicated copies + OR-reduce
    seg_bits = int(L.p3_bloom_seg_bits())
    nseg = (filter_size + seg_bits - 1) // seg_bits
    spr = (nseg + w - 1) // w                  # segments per rank
    seg_words = seg_bits // 32
    sharded = peer and nseg <= 1024 and 0 < num_hashes <= 32 and os.environ.get("P3_MG_FILTER", "sharded") != "replicated"
    for c, st, (tensors, rcounts) in zip(ctxs, stats, recvs):
        n = sum(rcounts)
        _check(L.p3_mg_owned_begin(c.h, owned_slots or max(2 * n, 1024)))
        if n:
            t = tensors[0].contiguous()
            _check(L.p3_mg_owned_insert(c.h, t.data_ptr(), n))
        no = C.c_uint64()
        if sharded:
            _check(L.p3_mg_owned_list(c.h, k, filter_size, num_hashes, w * spr * seg_words, C.byref(no)))
        else:
            _check(L.p3_mg_owned_end(c.h, k, filter_size, num_hashes, C.byref(no)))
        st.update(owned_solid=no.value)
        c.k, c.filter_size, c.num_hashes = k, filter_size, num_hashes
    del recvs
    lap("owned_dedupe")

    def filter_tensors():
        out = []
        for c in ctxs:
            ptr, nwords = C.c_void_p(), C.c_uint64()
            _check(L.p3_mg_filter(c.h, C.byref(ptr), C.byref(nwords)))
            out.append(dev_tensor(ptr.value, nwords.value, torch.int32, device))
        return out

    binned = False
    if sharded:
        n_all = [row[0] for row in comm.all_gather([[st["owned_solid"]] for st in stats])]
        force = os.environ.get("P3_BLOOM_BINNED")
        binned = (force != "0") if force is not None else (nseg >= 2 and sum(n_all) * num_hashes >= (1 << 22))
    if binned:
        # hashed indices are uniform: a full segment gets seg_bits / filter_size of a source's indices
        share = min(1.0, seg_bits / filter_size)
        cap_src = [int(n * num_hashes * share * 1.05) + 65536 for n in n_all]
        prefix = [sum(cap_src[:r]) for r in range(w)]
        tot_cap = sum(cap_src)
        ptr_rows = []
        for c in ctxs:
            pb = C.c_void_p()
            _check(L.p3_mg_bloom_buffer(c.h, spr * tot_cap, C.byref(pb)))
            ptr_rows.append([pb.value])
        table = comm.share(ptr_rows, dev_index)
        comm.barrier()
        count_rows = []
        for c, r in zip(ctxs, comm.local_ranks):
            base = (C.c_uint64 * nseg)(*[table[s // spr][0] + 4 * ((s % spr) * tot_cap + prefix[r]) for s in range(nseg)])
            counts = (C.c_uint64 * nseg)()
            _check(L.p3_mg_bloom_bin(c.h, nseg, base, cap_src[r], counts))
            count_rows.append([int(x) for x in counts])
        cnt = np.array(comm.all_gather(count_rows), dtype=np.int64).reshape(w, nseg)
        binned = bool((cnt <= np.array(cap_src)[:, None]).all())     # the same verdict on every rank
        comm.barrier()
        lap("bloom_bin")
    if binned:
        for c, r, row in zip(ctxs, comm.local_ranks, ptr_rows):
            first = r * spr
            nloc = max(0, min(spr, nseg - first))
            if nloc:
                hp = (C.c_uint64 * (nloc * w))(*[row[0] + 4 * (sl * tot_cap + prefix[src]) for sl in range(nloc) for src in range(w)])
                hn = (C.c_uint64 * (nloc * w))(*[int(cnt[src, first + sl]) for sl in range(nloc) for src in range(w)])
                _check(L.p3_mg_bloom_apply(c.h, first, nloc, w, hp, hn))
        lap("bloom_apply")
        comm.all_gather_shards(filter_tensors(), spr * seg_words)
        lap("filter_gather")
    else:
        if sharded:
            for c in ctxs:
                _check(L.p3_mg_bloom_direct(c.h))
        comm.or_reduce([f[: (filter_size + 31) // 32] for f in filter_tensors()])
    for st in stats:
        st["filter"] = "sharded" if binned else "replicated"
    torch.cuda.synchronize()

    mark("makebf")
    # ---- C: adjacency of the owned k-mers ---------------------------------------------------------------------
    for c, st in zip(ctxs, stats):
        nk, ne = c.dbg_adjacency()
        st.update(owned_kmers=nk, owned_edges=ne)
    mark("adjacency")
    torch.cuda.synchronize()
    ms = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
    for st in stats:
        st["stage_ms"] = ms
        st["count_sub_ms"] = sub
        st["lap_ms"] = dict(laps, **{"owner_" + kk: v for kk, v in st.get("owner_count_ms", {}).items()})
    return stats
