#!/usr/bin/env python
"""bench.py — headline benchmark of the k-mer-to-graph hot path (BASELINE.json metric).

A "step" is one pass of CountShortKmer -> MakeBF -> CheckDirections over one synthetic read
set. At N=1 the workload is BASELINE.json configs[1]: 100 Mbp random genome, 50x reads of
150 bp with 1% substitution errors, k=32. At N>1 (torchrun, one rank per GPU) the genome grows
with N (weak scaling: N x 100 Mbp at 50x; every rank parses 1/N of the reads = 5 Gbp).

  value     k-mers/s with the 2-bit read staging already resident in HBM (p3_reads_attach)
  e2e       the same pass through p3_assemble_hot_path_to_host on pinned HOST staging buffers, with
            the H2D of the reads and the D2H of filter/seeds/k-mers/adjacency inside the timing
  roofline  the count stage (scatter21 + insert_bins, the kernels that count the metric's k-mers):
            algorithmic bytes / CUDA-event time vs the measured HBM copy peak
  verified  before timing, a small instance of the same generator goes through the same code path (same
            number of ranks, real NVLink) and is compared with the oracle; the timed run's counts are
            compared with tests/golden/expected_counts.json (oracle/p3_scalecheck, same seed)
  cpu_baseline / --impl reference
            the unmodified reference (oracle/_ref/libp3ref.so; oracle port when absent) on a
            bounded, scaled-down sample of the same workload on the host cores
  --config 0 | 2 | 4   the other BASELINE.json configs as their own JSON lines (bench_configs.py)
  --config 3           human scale (3.1 Gbp, 30x) on 8 GPUs, or its per-rank share (387.5 Mbp, 11.6 Gbp of reads per rank)
                       on fewer: the multi-GPU path with the owner's insert in rounds (platanus3_b200/dist.py)
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "canonical k-mers/s counted + DBG edges/s"
UNIT = "k-mers/s"
K = 32
READ_LEN = 150
GENOME = 100_000_000
COVERAGE = 50
ERR = 0.01
SEED = 1234
# count stage's algorithmic bytes per 21-mer occurrence (DESIGN.md "Roofline"): 2 bits of read
# staging in, one 8-byte table slot read and written back
ALGO_BYTES_PER_KMER = 0.25 + 8 + 8
SAMPLE_GENOME = 200_000   # cpu_baseline / reference arm: same generator, 500x smaller genome
VERIFY_GENOME = 400_000   # the instance checked against the oracle before timing
CONFIG = 1                # BASELINE.json configs index of the N-GPU line (1 = headline; 3 = human scale, see use_config3)
EXPECTED = os.path.join(ROOT, "tests", "golden", "expected_counts.json")
TRAFFIC = os.path.join(ROOT, "profiles", "count_stage_traffic.json")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def count_stage_traffic():
    """DRAM bytes of the count stage's kernels for ONE step of configs[1] on one B200 (ncu dram__bytes_read.sum +
    dram__bytes_write.sum, profiles/); None until this round's capture exists"""
    if os.path.exists(TRAFFIC):
        d = json.load(open(TRAFFIC))
        return d.get("bytes_per_step"), d.get("source")
    return None, None


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(sm)}


def use_config3():
    """BASELINE.json configs[3]: human-scale 3.1 Gbp genome at 30x on 8 GPUs = 387.5 Mbp of genome and 11.6 Gbp of reads per
    rank. With fewer ranks the per-rank share stays the same (weak scaling towards the 8-GPU point): every rank runs the
    multi-GPU path at human-scale load (count table ~41 GB, owner bins re-used over several insert rounds)."""
    global CONFIG, GENOME, COVERAGE
    CONFIG, GENOME, COVERAGE = 3, 387_500_000, 30


def workload_config(n_gpus, genome=None):
    genome = genome or GENOME * n_gpus
    if CONFIG == 3:
        what = ("configs[3]: human-scale synthetic %.2f Gbp genome, %dx reads of %d bp (%.1f Gbp of reads), %.0f%% substitution errors, k=%d, %d GPUs"
                if n_gpus == 8 else
                "configs[3] per-rank share on fewer GPUs: synthetic %.2f Gbp genome, %dx reads of %d bp (%.1f Gbp of reads), %.0f%% substitution errors, k=%d, "
                "%d rank(s) x 11.6 Gbp of reads (the 8-GPU run is 3.1 Gbp)") % (genome / 1e9, COVERAGE, READ_LEN, genome * COVERAGE / 1e9, ERR * 100, K, n_gpus)
        return {"workload": what, "k": K, "short_k": 21, "cov_threshold": 2, "read_len": READ_LEN, "genome_bp": genome, "coverage": COVERAGE,
                "error_rate": ERR, "seed": SEED, "generator": "platanus3_b200/workload.py (hash-defined, identical on GPU / numpy / C)",
                "l2": "inputs larger than L2 (2.9 GB read staging, ~41 GB count table per GPU)", "parallelism": "%d GPUs" % n_gpus}
    return {"workload": ("configs[1]: synthetic %d Mbp genome, %dx reads of %d bp, %.0f%% substitution errors, k=%d"
                         % (genome // 10 ** 6, COVERAGE, READ_LEN, ERR * 100, K)) if n_gpus == 1 else
                        ("configs[1] scaled weakly: synthetic %d Mbp genome, %dx reads of %d bp, %.0f%% substitution errors, k=%d, %d ranks x 5 Gbp of reads"
                         % (genome // 10 ** 6, COVERAGE, READ_LEN, ERR * 100, K, n_gpus)),
            "k": K, "short_k": 21, "cov_threshold": 2, "read_len": READ_LEN, "genome_bp": genome, "coverage": COVERAGE,
            "error_rate": ERR, "seed": SEED, "generator": "platanus3_b200/workload.py (hash-defined, identical on GPU / numpy / C)",
            "l2": "inputs larger than L2 (1.25 GB read staging, >10 GB count table per GPU)",
            "parallelism": "1 GPU" if n_gpus == 1 else "%d GPUs" % n_gpus}


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """--impl reference: the reference's own CPU path on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _checkers import Oracle, Ref, have_ref, words_to_kmer_str
    from platanus3_b200 import workload
    seq, off = workload.make_reads_numpy(SAMPLE_GENOME, COVERAGE, READ_LEN, ERR, SEED)
    n_reads = len(off) - 1
    n_kmers = n_reads * (READ_LEN - 20)
    orc = Oracle()
    kind = "reference" if have_ref() else "port"
    fs, nh = orc.estimate_bloomfilter(int(off[-1]), K)

    def one_step():
        t0 = time.perf_counter()
        if kind == "reference":
            ref = Ref(K, threads=1)
            ref.add_reads_arrays(seq, off)
            ref.estimate()
            t0 = time.perf_counter()  # reads are "already loaded" like the GPU arm's resident staging
            keys, counts = ref.count_short()
            ref.make_bf()
            t1 = time.perf_counter()
            # the neighbour queries of every distinct solid k-mer (set taken from the oracle, untimed)
            solid = orc.solid_kmers(seq, off, K, keys, counts)[:, 0]
            km = np.frombuffer("".join(words_to_kmer_str([x], K) for x in solid).encode(), np.uint8)
            t2 = time.perf_counter()
            adj = ref.check_directions_batch(km, len(solid))
            t3 = time.perf_counter()
            ref.close()
            return (t1 - t0) + (t3 - t2), int(np.unpackbits(adj).sum())
        keys, counts = orc.count_short_kmers(seq, off)
        bloom, _, _, _ = orc.make_bf(seq, off, K, keys, counts, fs, nh)
        solid = orc.solid_kmers(seq, off, K, keys, counts)
        e = sum(bin(orc.check_directions(bloom, fs, nh, solid[i], K)).count("1") for i in range(len(solid)))
        return time.perf_counter() - t0, e

    for _ in range(args.warmup):
        one_step()
    times, edges = [], 0
    for _ in range(args.steps):
        t, edges = one_step()
        times.append(t)
    tot = sum(times)
    value = n_kmers * args.steps / tot
    sample = "a %d bp sample of the configured genome (same generator, %dx, %d bp reads, %.0f%% subs, k=%d): %d reads / %d 21-mer positions per step" % (
        SAMPLE_GENOME, COVERAGE, READ_LEN, ERR * 100, K, n_reads, n_kmers)
    cfg = workload_config(args.gpus)
    cfg["reference_sample"] = sample      # what this arm actually ran: a bounded sample, as the bench contract prescribes
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": cfg,
        "dbg_edges_per_s": edges * args.steps / tot,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "threads_note": "the reference's CountShortKmer (src/Load.cpp:105) and MakeBF (src/MakeBloomFilter.cpp:8) "
                                         "are single-threaded whatever -t says; -t only feeds MakeDBG's walk, which is outside this path"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline():
    """reference CPU path on the bounded sample, in a subprocess (rank 0, N=1 only)"""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                             capture_output=True, text=True, timeout=900, env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:  # report, never hide
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": "failed: %r" % (e,)}


# ------------------------------------------------------------------------------------------ verification
def filter_checksum_device(ctx, fs, device):
    """popcount and position-sensitive xor fold of the device filter (as oracle/p3_scalecheck.c computes them)"""
    import torch
    from platanus3_b200 import _lib, dist as pdist
    ptr, nwords = ctypes.c_void_p(), ctypes.c_uint64()
    _lib.check(_lib.lib().p3_mg_filter(ctx.h, ctypes.byref(ptr), ctypes.byref(nwords)))
    n32 = (fs + 31) // 32
    f = pdist.dev_tensor(ptr.value, n32, torch.int32, device)
    if n32 % 2:
        f = torch.cat([f, torch.zeros(1, dtype=torch.int32, device=device)])
    w64 = f.view(torch.int64)
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=device)
    pop = 0
    fx = torch.zeros((), dtype=torch.int64, device=device)
    step = 1 << 24
    for a in range(0, w64.numel(), step):
        part = w64[a:a + step]
        pop += int(lut[part.view(torch.uint8).to(torch.int64)].sum().item())
        idx = torch.arange(a, a + part.numel(), dtype=torch.int64, device=device)
        prod = part * (2 * idx + 1)
        # xor-reduce: fold halves
        while prod.numel() > 1:
            if prod.numel() % 2:
                prod = torch.cat([prod, torch.zeros(1, dtype=torch.int64, device=device)])
            h = prod.numel() // 2
            prod = prod[:h] ^ prod[h:]
        fx = fx ^ prod[0]
    return pop, int(fx.item()) & ((1 << 64) - 1)


def canonical_kmers_numpy(seq, positions, k):
    """canonical k-mers (uint64, k <= 32) starting at the given stream positions of an ACGT byte array"""
    code = np.zeros(256, np.uint64)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
    fw = np.zeros(len(positions), np.uint64)
    bw = np.zeros(len(positions), np.uint64)
    for j in range(k):
        c = code[seq[positions + j]]
        fw = (fw << np.uint64(2)) | c
        bw |= (np.uint64(3) - c) << np.uint64(2 * j)
    return np.minimum(fw, bw)


def verify_small_instance(world, rank, local, dev, stream, comm=None):
    """A small instance of the same generator through the same code path (same ranks, same transport), compared
    with the oracle: every owned 21-mer count, the filter bits, this rank's seeds, the owned solid k-mers and a
    sample of their adjacency bytes. Returns a dict for the JSON line; raises on any difference."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _checkers import Oracle
    from platanus3_b200 import _lib, workload, dist as pdist
    t0 = time.perf_counter()
    L = _lib.lib()
    n_total = workload.n_reads_for(VERIFY_GENOME, COVERAGE, READ_LEN)
    per = n_total // world // 16 * 16
    first = rank * per
    mine = per if rank < world - 1 else n_total - first
    seq, off = workload.make_reads_numpy(VERIFY_GENOME, COVERAGE, READ_LEN, ERR, SEED + 1)
    orc = Oracle()
    fs, nh = orc.estimate_bloomfilter(int(off[-1]), K)
    okeys, ocounts = orc.count_short_kmers(seq, off)
    obits, oseeds, osolid_flags, oadds = orc.make_bf(seq, off, K, okeys, ocounts, fs, nh, want_solid=True)
    osolid = np.unique(canonical_kmers_numpy(seq, np.flatnonzero(osolid_flags), K))
    wl = workload.make_reads(VERIFY_GENOME, COVERAGE, READ_LEN, ERR, SEED + 1, dev, first_read=first, n_reads=mine)
    ctx = _lib.Context(local, ctypes.c_void_p(stream.cuda_stream))
    ctx.attach(wl["packed"].data_ptr(), wl["total_bases"], wl["off"].data_ptr(), mine, None, keep=wl)
    # the small instance must take the code paths of the full-size run, not the small-input shortcuts
    forced = {"P3_DEDUPE_BINNED": "1", "P3_SET_PARTS": "5", "P3_BINNED_CLEARS": "1", "P3_BLOOM_BINNED": "1", "P3_BLOOM_SEG_BITS": str(1 << 20), "P3_PARTS": "24"}
    if CONFIG == 3:     # human scale runs the insert in rounds and the Bloom adds in passes: so does its small instance
        forced.update({"P3_MG_BIN_BUDGET": str(24_000_000 // world + 2_000_000), "P3_MG_BLOOM_BUDGET": "2000000"})
    saved = {kk: os.environ.get(kk) for kk in forced}
    os.environ.update({kk: v for kk, v in forced.items() if saved[kk] is None})
    if world == 1 and comm is None:
        ctx.count_short_kmers(int(len(okeys) / 0.5))
        n_adds, _ = ctx.make_bf(K, fs, nh, 2, 0)
        ctx.dbg_adjacency()
    else:
        st = pdist.run_hot_path([ctx], comm, K, fs, nh, int(len(okeys) / world / 0.5), owned_slots=int(len(osolid) / world / 0.4),
                                chunk_words=(1 << 16) if CONFIG != 3 else max(1024, (1 << 16) // world), device=dev)[0]
        n_adds = st["n_adds"]
        if CONFIG == 3:
            assert st["insert_rounds"] > 1, "verify: the small instance did not run the insert in rounds"
    for kk, v in saved.items():
        if v is None:
            os.environ.pop(kk, None)
    keys, counts = ctx.short_kmer_export()
    idx = np.searchsorted(okeys, keys)
    assert len(keys) and np.all(idx < len(okeys)) and np.array_equal(okeys[np.minimum(idx, len(okeys) - 1)], keys), "verify: a counted key is not in the oracle's table"
    assert np.array_equal(ocounts[idx], counts), "verify: 21-mer counts differ from the oracle"
    assert np.array_equal(ctx.bf_export(), obits), "verify: Bloom filter bits differ from the oracle"
    assert np.array_equal(ctx.seed_export(), oseeds[first:first + mine]), "verify: seed positions differ from the oracle"
    kmers, adj = ctx.dbg_export(sort=False)
    assert np.all(np.isin(kmers, osolid)) and len(np.unique(kmers)) == len(kmers), "verify: solid k-mer set differs from the oracle"
    if world > 1:
        assert all(L.p3_owner_of_key(int(x), world) == rank for x in kmers[:: max(1, len(kmers) // 500)]), "verify: k-mer on the wrong owner"
        assert all(L.p3_owner_of_key(int(x), world) == rank for x in keys[:: max(1, len(keys) // 500)]), "verify: 21-mer on the wrong owner"
    sample = range(0, len(kmers), max(1, len(kmers) // 4000))
    for i in sample:
        assert adj[i] == orc.check_directions(obits, fs, nh, kmers[i:i + 1], K), "verify: adjacency differs from the oracle"
    pop, fx = filter_checksum_device(ctx, fs, dev)
    b = np.concatenate([obits, np.zeros((-len(obits)) % 8, np.uint8)]).view("<u8")
    with np.errstate(over="ignore"):
        ofx = int(np.bitwise_xor.reduce(b * (2 * np.arange(len(b), dtype=np.uint64) + np.uint64(1))))
    assert (pop, fx) == (int(np.unpackbits(obits).sum()), ofx), "verify: device filter checksum differs from the host one"
    totals = [len(keys), int(counts.sum()), n_adds, len(kmers)]
    if world > 1:
        totals = comm.all_sum(totals)
    assert totals == [len(okeys), int(ocounts.sum()), oadds, len(osolid)], ("verify: totals differ from the oracle", totals)
    if comm is not None:
        comm.barrier()
        comm.close_shared()
        comm.barrier()
    ctx.close()
    return {"instance": "synthetic %d bp genome, %dx, %d bp reads, %.0f%% subs, k=%d, seed %d: %d reads" % (VERIFY_GENOME, COVERAGE, READ_LEN, ERR * 100, K, SEED + 1, n_total),
            "checked": "all owned 21-mer counts, filter bits, seeds of this rank's reads, owned solid k-mers, %d sampled adjacency bytes, totals over ranks" % len(sample),
            "against": "oracle/p3_oracle.c (pinned to the compiled reference)", "seconds": round(time.perf_counter() - t0, 1)}


def check_expected(n_gpus, counts, genome):
    """the timed run's counts against oracle/p3_scalecheck's for the same generator and seed"""
    if not os.path.exists(EXPECTED):
        return None
    for e in json.load(open(EXPECTED)):
        if e["genome_bp"] == genome and e["seed"] == SEED and e["k"] == K and e["coverage"] == COVERAGE and e["read_len"] == READ_LEN:
            keys = ["kmer_positions", "distinct_21mers", "bf_adds", "solid_kmers", "dbg_edges", "filter_size_bits", "num_hashes", "filter_popcount", "filter_xor"]
            bad = {kk: (counts.get(kk), e[kk]) for kk in keys if kk in counts and counts[kk] != e[kk]}
            assert not bad, "timed run differs from oracle/p3_scalecheck: %r" % (bad,)
            return {"source": "tests/golden/expected_counts.json (oracle/p3_scalecheck)", "matched": [kk for kk in keys if kk in counts]}
    return None


# ------------------------------------------------------------------------------------------ N > 1
def main_multi(args, rank, world, local, dev):
    """N > 1: one rank per GPU; k-mers hash-partitioned by owner (platanus3_b200/dist.py)."""
    import torch
    import torch.distributed as dist
    from platanus3_b200 import _lib, workload, dist as pdist
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ.pop("NCCL_DEBUG")   # NCCL prints its banner on stdout; stdout carries the one JSON line
    if "MASTER_ADDR" not in os.environ:      # plain `python bench.py --config 3`: a world of one rank
        import socket
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1", LOCAL_RANK=str(local))
    dist.init_process_group("nccl", device_id=dev)
    comm = pdist.TorchDistComm()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    verified = None
    if not args.no_verify:
        verified = verify_small_instance(world, rank, local, dev, stream, comm)
    genome = args.genome * world
    n_total = workload.n_reads_for(genome, COVERAGE, READ_LEN)
    per = n_total // world // 16 * 16
    first = rank * per
    n_reads = per if rank < world - 1 else n_total - first
    wl = workload.make_reads(genome, COVERAGE, READ_LEN, ERR, SEED, dev, first_read=first, n_reads=n_reads)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()        # the generator's temporaries (several GB) go back to the device
    total = wl["total_bases"]
    n_pos_local = n_reads * (READ_LEN - 20)
    all_bases, n_pos = comm.all_sum([total, n_pos_local])
    fs, nh = _lib.estimate_bloomfilter(all_bases, K)
    distinct21 = genome + int(all_bases * ERR * 21 * 1.05)
    table_slots = int(distinct21 / world / 0.55)
    owned_slots = int(genome * 1.2 / world / 0.5)
    chunk_words = 1 << 23
    ctx = _lib.Context(local, ctypes.c_void_p(stream.cuda_stream))
    ctx.attach(wl["packed"].data_ptr(), total, wl["off"].data_ptr(), n_reads, None, keep=wl)

    # human scale: owner bins for a quarter of the records at a time, shard buffers for an eighth of the bit indices
    budgets = dict(bin_budget_bytes=45e9, bloom_budget_bytes=8e9) if CONFIG == 3 else {}

    def step():
        return pdist.run_hot_path([ctx], comm, K, fs, nh, table_slots, 0, owned_slots, chunk_words, dev, **budgets)[0]

    def timed(fn, steps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    os.environ["P3_MG_SAMPLE_MEM"] = "1"       # device memory in use is sampled after every stage of the warm-up steps only
    hbm_local = 0
    for _ in range(max(args.warmup, 1)):
        hbm_local = max(hbm_local, step()["hbm_used_peak_bytes"])
    os.environ["P3_MG_SAMPLE_MEM"] = "0"
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    total_ms, st = timed(step, args.steps)
    launches = ctx.launch_count() - l0
    clocks = sampler.summary()
    pop, fx = filter_checksum_device(ctx, fs, dev)
    sums = comm.all_sum([st["owned_distinct21"], st["n_adds"], st["owned_solid"], st["owned_edges"], st["owned_positions"], launches])
    hbm_peak, = comm.all_max([[hbm_local]])
    assert sums[4] == n_pos, (sums, n_pos)
    counts = {"kmer_positions": n_pos, "distinct_21mers": sums[0], "bf_adds": sums[1], "solid_kmers": sums[2],
              "dbg_edges": sums[3], "filter_size_bits": fs, "num_hashes": nh, "filter_popcount": pop, "filter_xor": fx}
    expected = check_expected(world, counts, genome) if rank == 0 else None

    # e2e: pinned host staging -> upload -> distributed pass -> results back to the host
    e2e = None
    if not args.no_e2e:
        h_packed, h_off = wl["packed"].cpu().pin_memory(), wl["off"].cpu().pin_memory()
        # every rank holds the complete filter after the step; together the ranks return it ONCE (rank r its r-th part),
        # plus each rank's own seeds, owned k-mers and adjacency bytes
        nw32 = (fs + 31) // 32
        part = (nw32 + world - 1) // world
        fa, fb = min(rank * part, nw32), min((rank + 1) * part, nw32)
        out_bits = torch.empty(max(fb - fa, 1), dtype=torch.int32).pin_memory()
        out_seeds = torch.empty(n_reads, dtype=torch.int64).pin_memory()
        out_kmers = torch.empty(owned_slots, dtype=torch.int64).pin_memory()
        out_adj = torch.empty(owned_slots, dtype=torch.uint8).pin_memory()
        L = _lib.lib()
        got = [0]

        def step_e2e():
            _lib.check(L.p3_reads_upload(ctx.h, h_packed.data_ptr(), total, h_off.data_ptr(), n_reads, None))
            pdist.run_hot_path([ctx], comm, K, fs, nh, table_slots, 0, owned_slots, chunk_words, dev, **budgets)
            ptr, nwords = ctypes.c_void_p(), ctypes.c_uint64()
            _lib.check(L.p3_mg_filter(ctx.h, ctypes.byref(ptr), ctypes.byref(nwords)))
            if fb > fa:
                out_bits[: fb - fa].copy_(pdist.dev_tensor(ptr.value, nw32, torch.int32, dev)[fa:fb], non_blocking=True)
            _lib.check(L.p3_seed_export(ctx.h, out_seeds.data_ptr()))
            n = ctypes.c_uint64()
            _lib.check(L.p3_dbg_export(ctx.h, out_kmers.data_ptr(), out_adj.data_ptr(), owned_slots, ctypes.byref(n)))
            got[0] = n.value

        step_e2e()
        e2e_ms, _ = timed(step_e2e, args.steps)
        io = comm.all_sum([h_packed.numel() * 8 + h_off.numel() * 8, (fb - fa) * 4 + out_seeds.numel() * 8 + got[0] * 9])
        e2e = {"value": n_pos / (e2e_ms / args.steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
               "h2d_bytes_per_step": io[0], "d2h_bytes_per_step": io[1]}

    ms_per_step = total_ms / args.steps
    peak, peak_src = measured_peak()
    count_ms = st["stage_ms"]["count"]
    achieved = ALGO_BYTES_PER_KMER * (n_pos / world) / (count_ms * 1e-3) / 1e9
    if rank == 0:
        cfg = workload_config(world, genome)
        cfg["parallelism"] = ("%d GPUs, one process each: k-mers hash-partitioned by owner; 21-mer records, coverage verdicts and solid "
                              "k-mer occurrences are stored into the owners' receive regions over NVLink peer memory inside the binning "
                              "kernels (device-side barriers, no all-to-all), sharded Bloom filter + NCCL all-gather of the shards" % world
                              ) if st.get("exchange") == "peer" else (
                              "%d GPUs: k-mers hash-partitioned by owner; the same regions staged locally and moved by NCCL all_to_all_single" % world)
        cfg["chunk_words"], cfg["receive_set_bytes"] = chunk_words, st["set_bytes"]
        cfg["insert_rounds"], cfg["bloom_passes"] = st.get("insert_rounds"), st.get("bloom_passes")
        print(json.dumps({
            "metric": METRIC, "value": n_pos / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": cfg,
            "dbg_edges_per_s": sums[3] / (ms_per_step * 1e-3),
            "counts": counts, "verified": verified is not None, "verification": verified, "expected_counts": expected,
            "stage_ms": st["stage_ms"], "lap_ms": st.get("lap_ms"),
            "hbm_peak_bytes": hbm_peak,
            "roofline": {"kernel": "count stage, rank 0 (scatter21_kernel<PEER> + scatter_rec_kernel + insert_bins_kernel)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_kmer": ALGO_BYTES_PER_KMER, "kernel_ms": count_ms},
            "e2e": e2e, "gpu_launches": sums[5], "clocks": clocks,
        }))
    comm.barrier()
    comm.close_shared()
    comm.barrier()
    ctx.close()
    dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ N = 1
def main_single(args, local, dev):
    import torch
    from platanus3_b200 import _lib, workload
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    verified = None
    if not args.no_verify:
        verified = verify_small_instance(1, 0, local, dev, stream)
    genome = args.genome
    wl = workload.make_reads(genome, COVERAGE, READ_LEN, ERR, SEED, dev)
    torch.cuda.synchronize()
    n_reads, total = wl["n_reads"], wl["total_bases"]
    n_pos = n_reads * (READ_LEN - 20)
    fs, nh = _lib.estimate_bloomfilter(total, K)
    # capacity hints (a user gives these from the expected genome size / error rate)
    distinct21 = genome + int(total * ERR * 21 * 1.05)
    table_slots = int(distinct21 / 0.55)
    solid_slots = int(genome * 1.2 / 0.5)
    L = _lib.lib()
    ctx = _lib.Context(local, ctypes.c_void_p(stream.cuda_stream))
    ctx.attach(wl["packed"].data_ptr(), total, wl["off"].data_ptr(), n_reads, None, keep=wl)
    mem_peak = [0]

    def step_resident(sample_mem=False):
        ctx.count_short_kmers(table_slots)
        if sample_mem:      # cudaMemGetInfo is not free: sampled during warm-up only (the buffers are grow-only)
            mem_peak[0] = max(mem_peak[0], int(L.p3_device_mem_used(ctx.h)))
        ctx.make_bf(K, fs, nh, 2, solid_slots)
        if sample_mem:
            mem_peak[0] = max(mem_peak[0], int(L.p3_device_mem_used(ctx.h)))
        ctx.dbg_adjacency()

    for _ in range(max(args.warmup, 1)):
        step_resident(True)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    count_ms, stage_acc, sub_acc = [], {}, {}
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        ms = ctx.stage_ms()
        count_ms.append(ms["count21"])
        for kk, v in ms.items():
            stage_acc[kk] = stage_acc.get(kk, 0.0) + v
        for kk, v in ctx.count_substage_ms().items():
            sub_acc[kk] = sub_acc.get(kk, 0.0) + v
    e1.record(stream)
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    clocks = sampler.summary()
    st = ctx.stats()
    assert st["n_positions"] == n_pos, (st, n_pos)
    pop, fx = filter_checksum_device(ctx, fs, dev)
    counts = {"reads": n_reads, "kmer_positions": n_pos, "distinct_21mers": st["n_distinct21"],
              "bf_adds": st["n_adds"], "solid_kmers": st["n_distinct_solid"], "dbg_edges": st["n_edges"],
              "filter_size_bits": fs, "num_hashes": nh, "filter_popcount": pop, "filter_xor": fx}
    expected = check_expected(1, counts, genome)

    # e2e leg: pinned host staging in, filter / seeds / k-mers / adjacency out, all inside the timing
    e2e = None
    if not args.no_e2e:
        h_packed = wl["packed"].cpu().pin_memory()
        h_off = wl["off"].cpu().pin_memory()
        ctx.close()
        del wl
        torch.cuda.empty_cache()
        ctx = _lib.Context(local, ctypes.c_void_p(stream.cuda_stream))
        out_bits = torch.empty((fs + 7) // 8, dtype=torch.uint8).pin_memory()
        out_seeds = torch.empty(n_reads, dtype=torch.int64).pin_memory()
        out_kmers = torch.empty(solid_slots, dtype=torch.int64).pin_memory()
        out_adj = torch.empty(solid_slots, dtype=torch.uint8).pin_memory()
        n = ctypes.c_uint64()

        def step_e2e():
            _lib.check(L.p3_assemble_hot_path_to_host(ctx.h, h_packed.data_ptr(), total, h_off.data_ptr(), n_reads, None,
                                                      total, K, fs, nh, table_slots, solid_slots, out_bits.data_ptr(), out_seeds.data_ptr(),
                                                      out_kmers.data_ptr(), out_adj.data_ptr(), solid_slots, ctypes.byref(n)))

        step_e2e()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            step_e2e()
        e1.record(stream)
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1)
        assert n.value == st["n_distinct_solid"]
        assert int(np.unpackbits(out_adj[:n.value].numpy()).sum()) == st["n_edges"], "e2e: adjacency differs from the resident run"
        e2e = {"value": n_pos / (e2e_ms / args.steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
               "h2d_bytes_per_step": h_packed.numel() * 8 + h_off.numel() * 8,
               "d2h_bytes_per_step": out_bits.numel() + out_seeds.numel() * 8 + n.value * 9}

    ms_per_step = total_ms / args.steps
    value = n_pos / (ms_per_step * 1e-3)
    peak, peak_src = measured_peak()
    k_ms = sum(count_ms) / len(count_ms)
    achieved = ALGO_BYTES_PER_KMER * n_pos / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = count_stage_traffic()
    direct = os.environ.get("P3_COUNT_MODE", "binned") == "direct"
    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(1, genome),
        "dbg_edges_per_s": st["n_edges"] / (ms_per_step * 1e-3),
        "counts": counts, "verified": verified is not None, "verification": verified, "expected_counts": expected,
        "stage_ms": {kk: v / args.steps for kk, v in stage_acc.items()},
        "count_substage": {kk: v / args.steps for kk, v in sub_acc.items()},
        "hbm_peak_bytes": mem_peak[0],
        "roofline": {"kernel": "count21_kernel" if direct else "count stage = scatter21_kernel + insert_bins_kernel (dominant: insert_bins_kernel)",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": traffic if (genome == GENOME and not direct) else None, "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_kmer": ALGO_BYTES_PER_KMER, "kernel_ms": k_ms},
        "count_mode": os.environ.get("P3_COUNT_MODE", "binned"),
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    if genome != GENOME:
        result["config"]["workload"] += " [DEBUG OVERRIDE genome=%d: not the headline config]" % genome
    ctx.close()
    if not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline()
    print(json.dumps(result))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=1, help="BASELINE.json configs index: 1 (default, the headline), 0, 2, 3 or 4")
    ap.add_argument("--genome", type=int, default=GENOME, help="override genome size (debug only; invalidates the metric)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling only: skip the host-buffer leg (e2e is then null)")
    ap.add_argument("--with-reference", action="store_true", help="--config 0: also rerun the unmodified reference on the same file (minutes)")
    ap.add_argument("--no-verify", action="store_true", help="profiling only: skip the small-instance check against the oracle")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    try:
        import torch
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        if args.config == 3:    # human scale: the multi-GPU path at 11.6 Gbp of reads per rank, whatever the number of ranks
            use_config3()
            if args.genome == 100_000_000:
                args.genome = GENOME
            args.no_e2e = args.no_e2e or os.environ.get("P3_BENCH_E2E") != "1"   # 8 GB of pinned result buffers per rank: on request only
            return main_multi(args, rank, world, local, dev)
        if args.config != 1:
            import bench_configs
            return bench_configs.run(args, rank, world, local, dev)
        if world > 1:
            return main_multi(args, rank, world, local, dev)
        return main_single(args, local, dev)
    except BaseException as e:     # the exception text and the library's last error are the last stderr lines; exit non-zero
        if isinstance(e, SystemExit) and not e.code:
            raise
        traceback.print_exc()
        try:
            from platanus3_b200 import _lib
            err = _lib.lib().p3_last_error().decode()
        except Exception:
            err = "(library not loaded)"
        sys.stderr.write("bench.py rank %d/%d FAILED: %r | p3_last_error: %s\n" % (rank, world, e, err))
        sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
