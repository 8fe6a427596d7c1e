#!/usr/bin/env python
"""bench.py — headline benchmark of the k-mer-to-graph hot path (BASELINE.json metric).

A "step" is one pass of CountShortKmer -> MakeBF -> CheckDirections over one synthetic read
set. At N=1 the workload is BASELINE.json configs[1]: 100 Mbp random genome, 50x reads of
150 bp with 1% substitution errors, k=32.

  value     k-mers/s with the 2-bit read staging already resident in HBM (p3_reads_attach)
  e2e       the same pass through p3_assemble_hot_path on pinned HOST staging buffers, with
            the H2D of the reads and the D2H of filter/seeds/k-mers/adjacency inside the timing
  roofline  the count stage (scatter21 + insert_bins, the kernels that count the metric's k-mers):
            algorithmic bytes / CUDA-event time vs the measured HBM copy peak
  cpu_baseline / --impl reference
            the unmodified reference (oracle/_ref/libp3ref.so; oracle port when absent) on a
            bounded, scaled-down sample of the same workload on the host cores
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

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
# count stage's algorithmic bytes per 21-mer occurrence (DESIGN.md "Roofline"): 2 bits of read
# staging in, one 8-byte table slot read and written back
ALGO_BYTES_PER_KMER = 0.25 + 8 + 8
# DRAM bytes of the count stage's kernels for ONE step of configs[1] on one B200, from
# `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` (profiles/r01p3_traffic_full_workload.csv):
# scatter21 4.90 + 55.02 GB, insert_bins 69.50 + 32.51 GB (the histogram pass is gone: fixed-capacity bins)
COUNT_STAGE_TRAFFIC_BYTES = 161.93e9
SAMPLE_GENOME = 200_000  # cpu_baseline / reference arm: same generator, 500x smaller genome


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def run_reference(args):
    """--impl reference: the reference's own CPU path on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _checkers import Oracle, Ref, have_ref, words_to_kmer_str
    from platanus3_b200 import workload
    wl = workload.make_reads(SAMPLE_GENOME, COVERAGE, READ_LEN, ERR, 1234, "cpu", return_codes=True)
    seq = np.frombuffer(b"ACGT", np.uint8)[wl["codes"].numpy()]
    off = wl["off"].numpy().astype(np.uint64)
    n_kmers = wl["n_reads"] * (READ_LEN - 20)
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
    sample = "synthetic %d bp genome, %dx, %d bp reads, %.0f%% subs, k=%d: %d reads / %d 21-mer positions per step" % (
        SAMPLE_GENOME, COVERAGE, READ_LEN, ERR * 100, K, wl["n_reads"], n_kmers)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "dbg_edges_per_s": edges * args.steps / tot,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "threads_note": "the reference's CountShortKmer (src/Load.cpp:105) and MakeBF (src/MakeBloomFilter.cpp:8) "
                                         "are single-threaded whatever -t says; -t only feeds MakeDBG's walk, which is outside this path"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n_gpus):
    return {"workload": "configs[1]: synthetic %d Mbp genome, %dx reads of %d bp, %.0f%% substitution errors, k=%d"
                        % (GENOME // 10 ** 6, COVERAGE, READ_LEN, ERR * 100, K),
            "k": K, "short_k": 21, "cov_threshold": 2, "read_len": READ_LEN, "genome_bp": GENOME, "coverage": COVERAGE,
            "error_rate": ERR, "l2": "inputs larger than L2 (1.25 GB read staging, >10 GB count table)",
            "parallelism": "1 GPU" if n_gpus == 1 else "%d GPUs" % n_gpus}


def cpu_baseline():
    """reference CPU path on the bounded sample, in a subprocess (rank 0, N=1 only)"""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                             capture_output=True, text=True, timeout=900, env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:  # report, never hide
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": "failed: %r" % (e,)}


def main_multi(args, rank, world, local, dev):
    """N > 1: one rank per GPU; k-mers hash-partitioned by owner, NCCL all-to-all (platanus3_b200/dist.py).
    Weak scaling: the genome grows with N (N x 100 Mbp at 50x), every rank parses the same number of
    reads as the single-GPU run."""
    import torch
    import torch.distributed as dist
    from platanus3_b200 import _lib, workload, dist as pdist
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ.pop("NCCL_DEBUG")   # NCCL prints its banner on stdout; stdout carries the one JSON line
    dist.init_process_group("nccl", device_id=dev)
    comm = pdist.TorchDistComm()
    genome = args.genome * world
    wl = workload.make_reads(genome, COVERAGE / world, READ_LEN, ERR, 1234, dev, read_seed=5678 + rank)
    torch.cuda.synchronize()
    n_reads, total = wl["n_reads"], wl["total_bases"]
    n_pos_local = n_reads * (READ_LEN - 20)
    tot = comm.all_sum([total, n_pos_local])
    all_bases, n_pos = tot
    fs, nh = _lib.estimate_bloomfilter(all_bases, K)
    distinct21 = genome + int(all_bases * ERR * 21 * 1.05)
    table_slots = int(distinct21 / world / 0.55)
    owned_slots = int(genome * 1.2 / world / 0.5)
    solid_slots = int(min(genome, total) * 1.2 / 0.5)
    chunk_words = 1 << 25
    stream = torch.cuda.current_stream()
    ctx = _lib.Context(local, ctypes.c_void_p(stream.cuda_stream))
    ctx.attach(wl["packed"].data_ptr(), total, wl["off"].data_ptr(), n_reads, None, keep=wl)

    def step():
        return pdist.run_hot_path([ctx], comm, K, fs, nh, table_slots, solid_slots, owned_slots, chunk_words, dev)[0]

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

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    total_ms, st = timed(step, args.steps)
    launches = ctx.launch_count() - l0
    clocks = sampler.summary()
    sums = comm.all_sum([st["owned_distinct21"], st["n_adds"], st["owned_solid"], st["owned_edges"], st["owned_positions"], launches])
    assert sums[4] == n_pos, (sums, n_pos)

    # e2e: pinned host staging -> upload -> distributed pass -> results back to the host
    h_packed, h_off = wl["packed"].cpu().pin_memory(), wl["off"].cpu().pin_memory()
    comm.barrier()
    comm.close_shared()     # peers unmap this rank's receive buffers before it frees them
    comm.barrier()
    ctx.close()
    del wl
    ctx2 = _lib.Context(local, ctypes.c_void_p(stream.cuda_stream))
    out_bits = torch.empty((fs + 7) // 8, dtype=torch.uint8).pin_memory()
    out_seeds = torch.empty(n_reads, dtype=torch.int64).pin_memory()
    out_kmers = torch.empty(owned_slots, dtype=torch.int64).pin_memory()
    out_adj = torch.empty(owned_slots, dtype=torch.uint8).pin_memory()
    L = _lib.lib()
    got = [0]

    def step_e2e():
        _lib.check(L.p3_reads_upload(ctx2.h, h_packed.data_ptr(), total, h_off.data_ptr(), n_reads, None))
        ctx2.n_reads, ctx2.total_bases = n_reads, total
        pdist.run_hot_path([ctx2], comm, K, fs, nh, table_slots, solid_slots, owned_slots, chunk_words, dev)
        _lib.check(L.p3_bf_export(ctx2.h, out_bits.data_ptr()))
        _lib.check(L.p3_seed_export(ctx2.h, out_seeds.data_ptr()))
        n = ctypes.c_uint64()
        _lib.check(L.p3_dbg_export(ctx2.h, out_kmers.data_ptr(), out_adj.data_ptr(), owned_slots, ctypes.byref(n)))
        got[0] = n.value

    step_e2e()
    e2e_ms, _ = timed(step_e2e, args.steps)
    h2d = h_packed.numel() * 8 + h_off.numel() * 8
    d2h = out_bits.numel() + out_seeds.numel() * 8 + got[0] * 9
    io = comm.all_sum([h2d, d2h])

    ms_per_step = total_ms / args.steps
    peak, peak_src = measured_peak()
    count_ms = st["stage_ms"]["count"]
    achieved = ALGO_BYTES_PER_KMER * (n_pos / world) / (count_ms * 1e-3) / 1e9
    if rank == 0:
        cfg = workload_config(world)
        cfg["workload"] = "configs[1] scaled weakly: synthetic %d Mbp genome, %dx reads of %d bp, %.0f%% substitution errors, k=%d, %d ranks" % (
            genome // 10 ** 6, COVERAGE, READ_LEN, ERR * 100, K, world)
        cfg["genome_bp"] = genome
        cfg["parallelism"] = ("%d GPUs: k-mers hash-partitioned by owner; 21-mer records stored into the owners' buffers over NVLink "
                              "peer memory inside the binning kernel, coverage verdicts by remote RED.AND, NCCL all-to-all of "
                              "solid k-mers, filter OR-reduce" % world) if st.get("exchange") == "peer" else (
                              "%d GPUs: k-mers hash-partitioned by owner, NCCL all-to-all of binned 21-mers / k-mers, filter OR-reduce" % world)
        print(json.dumps({
            "metric": METRIC, "value": n_pos / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": cfg,
            "dbg_edges_per_s": sums[3] / (ms_per_step * 1e-3),
            "counts": {"kmer_positions": n_pos, "distinct_21mers": sums[0], "bf_adds": sums[1], "solid_kmers": sums[2],
                       "dbg_edges": sums[3], "filter_size_bits": fs, "num_hashes": nh},
            "stage_ms": st["stage_ms"], "count_substage": st["count_sub_ms"], "lap_ms": st.get("lap_ms"),
            "roofline": {"kernel": "count stage (owner binning fused with the exchange + L2-resident insert), rank 0", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_kmer": ALGO_BYTES_PER_KMER, "kernel_ms": count_ms},
            "e2e": {"value": n_pos / (e2e_ms / args.steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": io[0], "d2h_bytes_per_step": io[1]},
            "gpu_launches": sums[5], "clocks": clocks,
        }))
    comm.barrier()
    comm.close_shared()
    comm.barrier()
    ctx2.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--genome", type=int, default=GENOME, help="override genome size (debug only; invalidates the metric)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling only: skip the host-buffer leg (e2e is then null)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from platanus3_b200 import _lib, workload

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        return main_multi(args, rank, world, local, dev)

    genome = args.genome
    wl = workload.make_reads(genome, COVERAGE, READ_LEN, ERR, 1234 + rank, dev)
    torch.cuda.synchronize()
    n_reads, total = wl["n_reads"], wl["total_bases"]
    n_pos = n_reads * (READ_LEN - 20)
    fs, nh = _lib.estimate_bloomfilter(total, K)
    if os.environ.get("P3_BENCH_FS_SCALE"):      # experiment knob (invalidates the metric): a larger filter on the same reads
        fs = int(fs * float(os.environ["P3_BENCH_FS_SCALE"]))
        args.genome = -abs(genome)
    # capacity hints (a user gives these from the expected genome size / error rate)
    distinct21 = genome + int(total * ERR * 21 * 1.05)
    table_slots = int(distinct21 / 0.55)
    solid_slots = int(genome * 1.2 / 0.5)
    if os.environ.get("P3_BENCH_SET_SCALE"):     # experiment knob (invalidates the metric): a sparser / larger solid set
        solid_slots = int(solid_slots * float(os.environ["P3_BENCH_SET_SCALE"]))
        args.genome = -abs(genome)

    stream = torch.cuda.current_stream()
    ctx = _lib.Context(local, ctypes.c_void_p(stream.cuda_stream))
    ctx.attach(wl["packed"].data_ptr(), total, wl["off"].data_ptr(), n_reads, None, keep=wl)

    def step_resident():
        ctx.count_short_kmers(table_slots)
        ctx.make_bf(K, fs, nh, 2, solid_slots)
        ctx.dbg_adjacency()

    # pinned host staging for the e2e leg
    h_packed = wl["packed"].cpu().pin_memory()
    h_off = wl["off"].cpu().pin_memory()
    out_bits = torch.empty((fs + 7) // 8, dtype=torch.uint8).pin_memory()
    out_seeds = torch.empty(n_reads, dtype=torch.int64).pin_memory()
    out_kmers = torch.empty(solid_slots, dtype=torch.int64).pin_memory()
    out_adj = torch.empty(solid_slots, dtype=torch.uint8).pin_memory()
    L = _lib.lib()

    def step_e2e():
        _lib.check(L.p3_assemble_hot_path(ctx2.h, h_packed.data_ptr(), total, h_off.data_ptr(), n_reads, None,
                                          total, K, fs, nh, table_slots, solid_slots))
        _lib.check(L.p3_bf_export(ctx2.h, out_bits.data_ptr()))
        _lib.check(L.p3_seed_export(ctx2.h, out_seeds.data_ptr()))
        n = ctypes.c_uint64()
        _lib.check(L.p3_dbg_export(ctx2.h, out_kmers.data_ptr(), out_adj.data_ptr(), solid_slots, ctypes.byref(n)))
        return n.value

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(args.warmup):
        step_resident()
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

    # e2e leg (its own context; the resident one is released first so both fit in HBM)
    ctx.close()
    ctx2 = _lib.Context(local, ctypes.c_void_p(stream.cuda_stream))
    n_solid = 0
    e2e_ms = float("nan")
    if not args.no_e2e:
        step_e2e()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            n_solid = step_e2e()
        e1.record(stream)
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1)
    h2d = h_packed.numel() * 8 + h_off.numel() * 8
    d2h = out_bits.numel() + out_seeds.numel() * 8 + n_solid * 9

    ms_per_step = total_ms / args.steps
    value = n_pos / (ms_per_step * 1e-3)
    peak, peak_src = measured_peak()
    k_ms = sum(count_ms) / len(count_ms)
    achieved = ALGO_BYTES_PER_KMER * n_pos / (k_ms * 1e-3) / 1e9
    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(world),
        "dbg_edges_per_s": st["n_edges"] / (ms_per_step * 1e-3),
        "counts": {"reads": n_reads, "kmer_positions": n_pos, "distinct_21mers": st["n_distinct21"],
                   "bf_adds": st["n_adds"], "solid_kmers": st["n_distinct_solid"], "dbg_edges": st["n_edges"],
                   "filter_size_bits": fs, "num_hashes": nh},
        "stage_ms": {kk: v / args.steps for kk, v in stage_acc.items()},
        "count_substage": {kk: v / args.steps for kk, v in sub_acc.items()},
        "roofline": {"kernel": "count stage = scatter21_kernel + insert_bins_kernel (dominant: insert_bins_kernel)"
                               if os.environ.get("P3_COUNT_MODE", "binned") != "direct" else "count21_kernel",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": COUNT_STAGE_TRAFFIC_BYTES if (genome == GENOME and os.environ.get("P3_COUNT_MODE", "binned") != "direct") else None,
                     "traffic_source": "profiles/r01p3_traffic_full_workload.csv (ncu dram__bytes_read+write of the two kernels, one step)",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_kmer": ALGO_BYTES_PER_KMER, "kernel_ms": k_ms},
        "count_mode": os.environ.get("P3_COUNT_MODE", "binned"),
        "e2e": {"value": n_pos / (e2e_ms / args.steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks,
    }
    if genome != GENOME or args.genome != genome:
        result["config"]["workload"] += " [DEBUG OVERRIDE genome=%d: not the headline config]" % genome
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ctx2.close()
        result["cpu_baseline"] = cpu_baseline()
    if rank == 0:
        print(json.dumps(result))


if __name__ == "__main__":
    main()
