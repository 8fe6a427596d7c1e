"""Multi-GPU hot path: canonical k-mers hash-partitioned by owner rank (csrc/p3_multi.inc.cu).

One process per GPU (torch.distributed for the plumbing). All heavy work is in the CUDA library; this
module only sequences the stages. With the default transport the binning kernels store every
destination rank's records straight into that rank's receive region over NVLink peer memory and the
ranks meet at device-side barriers (p3_mg_sync), so a step is one long stream of kernel launches with
three host waits (after the count, after the k-mer de-duplication, after the adjacency):

  A   every rank bins its 21-mers by owner -> owner sorts them into its partition bins -> one insert sweep
  B1  owner sweeps its bins again: positions of count-1 keys -> back to their source rank -> bits cleared
  B2  solid plane and seeds are local; every solid occurrence (k-mer + adjacency hint) -> owner ->
      sorted by set partition -> one de-duplication sweep; sharded Bloom adds; all-gather of the shards
  C   owner runs CheckDirections for its k-mers against the (now complete, local) filter

P3_MG_EXCHANGE=nccl selects the staged transport: the same regions are filled in a local buffer and moved
with all_to_all_single (the baseline the fused path is measured against). The same driver runs over an
emulated communicator (several contexts of one process on one GPU and one stream), which is how the
parity tests exercise the distributed algorithm on a single-GPU box.

Inputs whose records an owner cannot hold at once (BASELINE.json configs[3]) run A and B1 in rounds — over key
ranges (every round scans all reads and moves the keys of a group of table partitions, whose verdicts follow at once)
or, with P3_MG_ROUNDS=chunks, over read chunks (the records travel a second time for the verdicts) — B2 in rounds
of chunks and the sharded Bloom adds in passes (plan_rounds, bin_budget_bytes, bloom_budget_bytes).
Multi-word k-mers (k > 32) take B2 as W-word records to the owner of their std::hash (p3_mg_long_*).
"""
import ctypes as C
import os

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


def _check(rc):
    _lib.check(rc)


def _staged_info(ctx, stage, rset):
    out = (C.c_uint64 * 8)()
    _check(_lib.lib().p3_mg_staged_buffers(ctx.h, stage, rset, out))
    return [int(x) for x in out]


# ---------------------------------------------------------------------------- communicators
class TorchDistComm:
    """one rank per process over torch.distributed (NCCL on the box, gloo in the CPU tests)"""
    same_stream = 0

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.local_ranks = [self.rank]
        self._mapped = {}

    def _dev(self):
        return "cuda" if torch.cuda.is_available() and self.dist.get_backend(self.group) == "nccl" else "cpu"

    def all_sum(self, values):
        """values: [ints] or, like the emulated communicator, [[ints]] of the one local rank"""
        if values and isinstance(values[0], (list, tuple)):
            (values,) = values
        t = torch.tensor(values, dtype=torch.int64, device=self._dev())
        self.dist.all_reduce(t, group=self.group)
        return [int(x) for x in t.tolist()]

    def all_max(self, values):
        (row,) = values
        t = torch.tensor(row, dtype=torch.int64, device=self._dev())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
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
        for ptr, dev in self._mapped.values():
            L.p3_ipc_close(dev, C.c_void_p(ptr))
        self._mapped = {}
        self.__dict__.pop("_p3_setup", None)
        self.__dict__.pop("_p3_bloom", None)

    def staged_exchange(self, ctxs, stage, rset, device):
        """staged transport: all-to-all of the fixed-size regions (8-byte records, auxiliary block, counts)"""
        (c,) = ctxs
        w = self.world
        info = _staged_info(c, stage, rset)
        for s, r, per in ((info[0], info[1], info[2]), (info[3], info[4], info[5])):
            if per:
                self.dist.all_to_all_single(dev_tensor(r, w * per, torch.uint8, device), dev_tensor(s, w * per, torch.uint8, device), group=self.group)
        self.dist.all_to_all_single(dev_tensor(info[7], w, torch.int64, device), dev_tensor(info[6], w, torch.int64, device), group=self.group)

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

    def all_gather_shards(self, filters, shard):
        """filters: the one local filter tensor of world*shard words whose shard `rank` is final ->
        every shard final on every rank (in-place all-gather)"""
        f, = filters
        if self.world > 1:
            self.dist.all_gather_into_tensor(f[: self.world * shard], f[self.rank * shard:(self.rank + 1) * shard], group=self.group)


class EmulatedComm:
    """all ranks live in this process (one context each, all on ONE stream); exchanges are copies"""
    same_stream = 1

    def __init__(self, world):
        self.world = world
        self.local_ranks = list(range(world))

    def all_sum(self, values_per_rank):
        return [int(sum(v)) for v in zip(*values_per_rank)]

    def all_max(self, values_per_rank):
        return [int(max(v)) for v in zip(*values_per_rank)]

    def all_gather(self, rows):
        return [list(r) for r in rows]

    def barrier(self):
        torch.cuda.synchronize()

    def share(self, ptr_rows, device_index):
        return [list(r) for r in ptr_rows]   # one process: every rank's pointers are valid as they are

    def close_shared(self):
        self.__dict__.pop("_p3_setup", None)
        self.__dict__.pop("_p3_bloom", None)

    def staged_exchange(self, ctxs, stage, rset, device):
        w = self.world
        infos = [_staged_info(c, stage, rset) for c in ctxs]
        for a, b, p in ((0, 1, 2), (3, 4, 5)):
            per = infos[0][p]
            if not per:
                continue
            send = [dev_tensor(i[a], w * per, torch.uint8, device) for i in infos]
            recv = [dev_tensor(i[b], w * per, torch.uint8, device) for i in infos]
            for i in range(w):
                for j in range(w):
                    recv[j][i * per:(i + 1) * per] = send[i][j * per:(j + 1) * per]
        sc = [dev_tensor(i[6], w, torch.int64, device) for i in infos]
        rc = [dev_tensor(i[7], w, torch.int64, device) for i in infos]
        for i in range(w):
            for j in range(w):
                rc[j][i] = sc[i][j]

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


# ---------------------------------------------------------------------------- driver
def default_set_bytes(chunk_words, world):
    """receive set for one chunk of 21-mer records (12 B each): expected share + 3 % + slack per region"""
    per_region = int(chunk_words * 32 / world * 1.03) + 8192 + 2 * 4096    # the library rounds a region DOWN to whole sort tiles (4096 records)
    return per_region * 12 * world


def _setup(ctxs, comm, set_bytes, transport, dev_index):
    """arena of every local rank; handles exchanged and peers mapped ONCE per (contexts, size, transport)"""
    L = _lib.lib()
    key = (tuple(c.h for c in ctxs), set_bytes, transport)
    if getattr(comm, "_p3_setup", None) == key:
        return
    ptr_rows = []
    for c, r in zip(ctxs, comm.local_ranks):
        p = C.c_void_p()
        _check(L.p3_mg_arena(c.h, comm.world, r, set_bytes, transport, C.byref(p)))
        ptr_rows.append([p.value])
    table = comm.share(ptr_rows, dev_index)     # [rank][0]
    arr = (C.c_uint64 * comm.world)(*[table[j][0] for j in range(comm.world)])
    for c in ctxs:
        _check(L.p3_mg_connect(c.h, arr, comm.same_stream))
    comm.barrier()      # every arena is zeroed and mapped before anybody stores into it
    comm._p3_setup = key


def plan_rounds(n_chunks, owner_total, bin_budget_bytes):
    """how many insert rounds keep the owner's bins (16 B per received record + slack) inside the budget -> (n_rounds, chunks per round)"""
    want = max(1, -(-int(owner_total * 16.5) // max(int(bin_budget_bytes), 1)))
    n_rounds = max(1, min(n_chunks, want))
    cpr = -(-n_chunks // n_rounds)
    return -(-n_chunks // cpr), cpr


def run_hot_path(ctxs, comm, k, filter_size, num_hashes, table_slots, solid_slots=0, owned_slots=0,
                 chunk_words=None, device=None, set_bytes=None, cov_threshold=2, bin_budget_bytes=None, bloom_budget_bytes=None):
    """ctxs: the Context of every LOCAL rank (reads already attached/uploaded, all contexts of one process on
    the current torch stream), in comm.local_ranks order. table_slots / owned_slots: per-rank capacities of
    the count table and of the owned solid k-mer set. Returns one stats dict per local rank; the results
    stay in the contexts (owned 21-mer counts, owned k-mers + adjacency, local seeds, the complete filter).
    solid_slots is accepted for compatibility and unused (there is no local k-mer set any more).
    bin_budget_bytes (default 80 GB, env P3_MG_BIN_BUDGET): when an owner's partition bins for ALL its records would not fit,
    the count, the verdicts and the de-duplication run in rounds of chunks over bins sized for one round (human-scale
    inputs, BASELINE.json configs[3]); bloom_budget_bytes (default 16 GB, env P3_MG_BLOOM_BUDGET) bounds the shard owners'
    bit-index buffers the same way (several passes over the owned k-mer lists). Results do not depend on either."""
    L = _lib.lib()
    w = comm.world
    device = device or torch.device("cuda", torch.cuda.current_device())
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    stats = [dict(rank=r) for r in comm.local_ranks]
    marks = []
    mem_peak = [0]

    sample_mem = os.environ.get("P3_MG_SAMPLE_MEM", "0") != "0"    # cudaMemGetInfo is not free: only on request (bench warm-up)

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))
        if sample_mem:
            mem_peak[0] = max([mem_peak[0]] + [int(L.p3_device_mem_used(c.h)) for c in ctxs])

    peer = os.environ.get("P3_MG_EXCHANGE", "peer") != "nccl"
    transport = 0 if peer else 1
    n_words = [(c.total_bases + 31) // 32 for c in ctxs]
    pos_upper = [max(c.total_bases - 20 * c.n_reads, 0) for c in ctxs]
    total_pos, = comm.all_sum([[p] for p in pos_upper])
    max_words, ppw_e6 = comm.all_max([[nw, min(32_000_000, -(-p * 1_000_000 // max(nw, 1)))] for nw, p in zip(n_words, pos_upper)])
    owner_total = int(total_pos / w * 1.02) + 65536
    bin_budget = int(bin_budget_bytes or float(os.environ.get("P3_MG_BIN_BUDGET", 80e9)))
    cw = chunk_words
    if not cw:     # one chunk per round
        cw = -(-max(max_words, 128) // plan_rounds(1 << 30, owner_total, bin_budget)[0])
    cw = (cw + 127) // 128 * 128
    n_chunks, = comm.all_max([[max((nw + cw - 1) // cw, 1)] for nw in n_words])
    n_rounds, cpr = plan_rounds(n_chunks, owner_total, bin_budget)
    rounds = [range(r * cpr, min((r + 1) * cpr, n_chunks)) for r in range(n_rounds)]
    # what one round can bring an owner: every rank's chunks of the round, spread evenly over the owners (a hash)
    owner_positions = owner_total if n_rounds == 1 else min(owner_total, int(cpr * cw * (ppw_e6 / 1e6) * 1.02) + 65536)
    set_bytes = set_bytes or default_set_bytes(min(cw, max_words + 128), w)
    if k > 32:      # a region must hold a useful number of W-word k-mer records
        set_bytes = max(set_bytes, w * 8 * ((2 * k + 63) // 64) * 8192)
    _setup(ctxs, comm, set_bytes, transport, dev_index)

    def sync(stage, rset):
        if peer:
            for c in ctxs:
                _check(L.p3_mg_sync(c.h))
        else:
            comm.staged_exchange(ctxs, stage, rset, device)

    def stage_start():
        for c in ctxs:
            _check(L.p3_mg_sync(c.h))

    mark("start")
    # ---- A: count ------------------------------------------------------------------------------
    # rounds over key ranges need at least one table partition per round; otherwise (tiny tables) rounds of read chunks
    keyed = n_rounds > 1 and os.environ.get("P3_MG_ROUNDS", "keys") != "chunks" and int(L.p3_table_partitions(table_slots)) >= n_rounds
    for c in ctxs:
        if keyed:
            _check(L.p3_mg_count_begin_keyed(c.h, table_slots, owner_total, cw, n_chunks, n_rounds))
        else:
            _check(L.p3_mg_count_begin(c.h, table_slots, owner_positions, cw, n_chunks))
    stage_start()

    def count_chunks(chs):
        for ch in chs:
            for c in ctxs:
                _check(L.p3_mg_count_send(c.h, ch))
            sync(0, ch & 1)
            for c in ctxs:
                _check(L.p3_mg_count_recv(c.h, ch))

    n_slices = 1
    cover_ms_events = []

    def cover_slices():
        stage_start()
        for sl in range(n_slices):
            for c in ctxs:
                _check(L.p3_mg_cover_send(c.h, cov_threshold, sl))
            sync(1, sl & 1)
            for c in ctxs:
                _check(L.p3_mg_cover_recv(c.h, sl))

    if keyed:
        # rounds over key ranges: every round scans all reads, only the keys of the round's table partitions travel; their
        # counts are final when the round's insert ends, so the round's verdicts follow at once (nothing is sent twice)
        for r in range(n_rounds):
            if r:
                for c in ctxs:
                    _check(L.p3_mg_count_next_round(c.h))
                stage_start()
            for c in ctxs:
                _check(L.p3_mg_key_round_begin(c.h, r))
            count_chunks(range(n_chunks))
            for c in ctxs:
                _check(L.p3_mg_count_finish(c.h))
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
            if r == 0:
                for c in ctxs:
                    ns = C.c_uint32()
                    _check(L.p3_mg_cover_begin_keyed(c.h, cov_threshold, int(table_slots * 0.7 / n_rounds) + 4096, C.byref(ns)))
                    n_slices = ns.value
            for c in ctxs:
                _check(L.p3_mg_cover_key_round(c.h, cov_threshold))
            cover_slices()
            e1 = torch.cuda.Event(enable_timing=True); e1.record()
            cover_ms_events.append((e0, e1))
    else:
        for r, chs in enumerate(rounds):
            if r:
                for c in ctxs:
                    _check(L.p3_mg_count_next_round(c.h))
            count_chunks(chs)
            for c in ctxs:
                _check(L.p3_mg_count_finish(c.h))
    mark("count")
    for c, st in zip(ctxs, stats):
        _check(L.p3_mg_count_end(c.h))
        a, b = C.c_uint64(), C.c_uint64()
        _check(L.p3_short_kmer_stats(c.h, C.byref(a), C.byref(b)))
        st.update(owned_positions=a.value, owned_distinct21=b.value, exchange="peer" if peer else "nccl")
        st["owner_count_ms"] = {kk: v for kk, v in c.count_substage_ms().items() if kk in ("scatter", "insert")}
    # ---- B1: verdicts back to the reads ---------------------------------------------------------------
    if not keyed:
        owner_distinct, = comm.all_max([[st["owned_distinct21"]] for st in stats])
        for c in ctxs:
            ns = C.c_uint32()
            _check(L.p3_mg_cover_begin(c.h, cov_threshold, owner_distinct, C.byref(ns)))
            n_slices = ns.value
        if n_rounds == 1:
            cover_slices()
        else:       # rounds of read chunks (P3_MG_ROUNDS=chunks): send and sort every round's records again, then that round's verdicts
            for chs in rounds:
                for c in ctxs:
                    _check(L.p3_mg_cover_rebin_begin(c.h))
                stage_start()
                count_chunks(chs)
                for c in ctxs:
                    _check(L.p3_mg_cover_rebin_end(c.h))
                cover_slices()
    mark("coverage")
    # ---- B2: solid occurrences to their owners ---------------------------------------------------------
    n_long_chunks = 0
    if k > 32:
        # multi-word k-mers: every solid occurrence travels as its W canonical words; the owner stores what arrives and
        # de-duplicates the store after the last chunk (csrc/p3_multi.inc.cu, p3_mg_long_*)
        local_adds = []
        for c in ctxs:
            na = C.c_uint64()
            _check(L.p3_mg_long_solid(c.h, k, C.byref(na)))
            local_adds.append(na.value)
        total_adds, = comm.all_sum([[a] for a in local_adds])
        opw_e6, = comm.all_max([[-(-a * 1_000_000 // max(nw, 1))] for a, nw in zip(local_adds, n_words)])
        owner_occ = int(total_adds / w * 1.05) + 65536
        cwl = nchl = 0
        for c in ctxs:
            a, b = C.c_uint64(), C.c_uint64()
            _check(L.p3_mg_long_begin(c.h, owned_slots or max(2 * owner_occ, 4096), owner_occ, min(32.0, opw_e6 / 1e6 * 1.4 + 0.5), max_words,
                                      C.byref(a), C.byref(b)))
            cwl, nchl = a.value, b.value
        n_long_chunks = nchl if total_adds else 0
        stage_start()
        for ch in range(n_long_chunks):
            for c in ctxs:
                _check(L.p3_mg_long_send(c.h, ch))
            sync(3, ch & 1)
            for c in ctxs:
                _check(L.p3_mg_long_recv(c.h, ch))
        for c in ctxs:
            _check(L.p3_mg_long_finish(c.h))
    else:
        for c in ctxs:     # default capacity: a rank owns about 1/w of all k-mers, whatever share of the reads it parsed
            _check(L.p3_mg_solid_begin(c.h, k, owned_slots or max(2 * total_pos // w + 4096, 4096)))
        stage_start()
        for r, chs in enumerate(rounds):
            if r:
                for c in ctxs:
                    _check(L.p3_mg_solid_next_round(c.h))
            for ch in chs:
                for c in ctxs:
                    _check(L.p3_mg_solid_send(c.h, ch))
                sync(2, ch & 1)
                for c in ctxs:
                    _check(L.p3_mg_solid_recv(c.h, ch))
        for c in ctxs:
            _check(L.p3_mg_solid_finish(c.h))
    mark("dedupe")
    # the filter: sharded, binned adds (each rank owns a contiguous run of 16 MB segments and receives
    # the bit indices that fall into them) or, as fallback, adds into replicated copies + OR-reduce
    seg_bits = int(L.p3_bloom_seg_bits())
    nseg = (filter_size + seg_bits - 1) // seg_bits
    spr = (nseg + w - 1) // w                  # segments per rank
    seg_words = seg_bits // 32
    sharded = peer and nseg <= 1024 and 0 < num_hashes <= 32 and os.environ.get("P3_MG_FILTER", "sharded") != "replicated"
    for c, st in zip(ctxs, stats):
        na, no = C.c_uint64(), C.c_uint64()
        _check(L.p3_mg_solid_end(c.h, filter_size, num_hashes, w * spr * seg_words if sharded else 0, C.byref(na), C.byref(no)))
        st.update(n_adds=na.value, owned_solid=no.value)
        c.k, c.filter_size, c.num_hashes = k, filter_size, num_hashes

    def filter_tensors():
        out = []
        for c in ctxs:
            ptr, nwords = C.c_void_p(), C.c_uint64()
            _check(L.p3_mg_filter(c.h, C.byref(ptr), C.byref(nwords)))
            out.append(dev_tensor(ptr.value, nwords.value, torch.int32, device))
        return out

    binned = False
    n_pass = 1
    n_all = [row[0] for row in comm.all_gather([[st["owned_solid"]] for st in stats])]
    if sharded:
        force = os.environ.get("P3_BLOOM_BINNED")
        binned = (force != "0") if force is not None else (nseg >= 2 and sum(n_all) * num_hashes >= (1 << 22))
    if binned:
        # hashed indices are uniform: a full segment gets seg_bits / filter_size of a source's indices. The shard owners'
        # buffers hold one PASS over the sources' k-mer lists; several passes when all indices at once exceed the budget
        share = min(1.0, seg_bits / filter_size)
        bloom_budget = int(bloom_budget_bytes or float(os.environ.get("P3_MG_BLOOM_BUDGET", 16e9)))
        full = 4 * spr * sum(int(n * num_hashes * share * 1.05) + 65536 for n in n_all)
        n_pass = max(1, min(-(-full // max(bloom_budget, 1)), max(max(n_all), 1)))
        per_pass = [-(-n // n_pass) for n in n_all]
        cap_src = [int(n * num_hashes * share * 1.05) + 65536 for n in per_pass]
        prefix = [sum(cap_src[:r]) for r in range(w)]
        tot_cap = sum(cap_src)
        # the buffers only ever grow, and their size is the same function of the same numbers on every rank:
        # when it does not exceed what was shared before, every rank still has the same buffer and mapping
        cached = getattr(comm, "_p3_bloom", None)
        if cached and cached[0] == tuple(c.h for c in ctxs) and spr * tot_cap <= cached[1]:
            ptr_rows, table = cached[2], cached[3]
        else:
            ptr_rows = []
            for c in ctxs:
                pb = C.c_void_p()
                _check(L.p3_mg_bloom_buffer(c.h, spr * tot_cap, C.byref(pb)))
                ptr_rows.append([pb.value])
            table = comm.share(ptr_rows, dev_index)
            comm.barrier()
            comm._p3_bloom = (tuple(c.h for c in ctxs), spr * tot_cap, ptr_rows, table)
        for ps in range(n_pass):
            count_rows = []
            for c, r in zip(ctxs, comm.local_ranks):
                base = (C.c_uint64 * nseg)(*[table[s // spr][0] + 4 * ((s % spr) * tot_cap + prefix[r]) for s in range(nseg)])
                counts = (C.c_uint64 * nseg)()
                _check(L.p3_mg_bloom_bin_range(c.h, nseg, base, cap_src[r], counts, ps * per_pass[r], per_pass[r]))
                count_rows.append([int(x) for x in counts])
            cnt = np.array(comm.all_gather(count_rows), dtype=np.int64).reshape(w, nseg)
            binned = bool((cnt <= np.array(cap_src)[:, None]).all())     # the same verdict on every rank
            stage_start()       # every source's stores have landed
            if not binned:      # a segment outgrew its share: every rank adds directly (from scratch) and the copies are OR-reduced
                break
            for c, r, row in zip(ctxs, comm.local_ranks, ptr_rows):
                first = r * spr
                nloc = max(0, min(spr, nseg - first))
                if nloc:
                    hp = (C.c_uint64 * (nloc * w))(*[row[0] + 4 * (sl * tot_cap + prefix[src]) for sl in range(nloc) for src in range(w)])
                    hn = (C.c_uint64 * (nloc * w))(*[int(cnt[src, first + sl]) for sl in range(nloc) for src in range(w)])
                    _check(L.p3_mg_bloom_apply(c.h, first, nloc, w, hp, hn))
            if ps + 1 < n_pass:
                stage_start()   # nobody overwrites a buffer of the next pass before it has been applied
    if binned:
        comm.all_gather_shards(filter_tensors(), spr * seg_words)
        stage_start()       # nobody overwrites a bloom buffer of the next step before it has been applied
    elif sum(n_all):
        for c in ctxs:
            _check(L.p3_mg_bloom_direct(c.h))
        comm.or_reduce([f[: (filter_size + 31) // 32] for f in filter_tensors()])
    # (no solid k-mer anywhere: every copy of the filter is empty already)
    for c, st in zip(ctxs, stats):
        st["filter"] = "sharded" if binned else "replicated"
        _check(L.p3_mg_makebf_done(c.h))

    mark("bloom")
    # ---- C: adjacency of the owned k-mers ---------------------------------------------------------------------
    for c, st in zip(ctxs, stats):
        nk, ne = c.dbg_adjacency()
        st.update(owned_kmers=nk, owned_edges=ne)
    mark("adjacency")
    torch.cuda.synchronize()
    ms = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
    if cover_ms_events:      # key-range rounds interleave the verdicts with the count: book them where they belong
        cov = sum(a.elapsed_time(b) for a, b in cover_ms_events)
        ms["count"] -= cov
        ms["coverage"] += cov
    ms["makebf"] = ms["dedupe"] + ms["bloom"]
    for st in stats:
        st["stage_ms"] = ms
        st["lap_ms"] = {"owner_" + kk: v for kk, v in st.get("owner_count_ms", {}).items()}
        st["hbm_used_peak_bytes"] = mem_peak[0]
        st["n_chunks"], st["cover_slices"], st["set_bytes"] = n_chunks, n_slices, set_bytes
        st["insert_rounds"], st["bloom_passes"], st["long_chunks"] = n_rounds, n_pass, n_long_chunks
        st["round_mode"] = "one round" if n_rounds == 1 else ("key ranges" if keyed else "read chunks")
    return stats
