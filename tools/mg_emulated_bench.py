#!/usr/bin/env python
"""Times the DISTRIBUTED algorithm (platanus3_b200/dist.py) with R ranks emulated as R contexts of
one process on one GPU — the configuration ncu can profile (never a multi-rank command).
  python tools/mg_emulated_bench.py --world 1 --genome 100000000 [--steps 2]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--genome", type=int, default=100_000_000, help="genome size PER RANK (weak scaling, as bench.py)")
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    import torch
    from platanus3_b200 import _lib, workload, dist as pdist
    K, COV, RL, ERR = 32, 50, 150, 0.01
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    w = args.world
    genome = args.genome * w
    stream = torch.cuda.current_stream()
    ctxs, keep, total_all = [], [], 0
    for r in range(w):
        wl = workload.make_reads(genome, COV / w, RL, ERR, 1234, dev, read_seed=5678 + r)
        torch.cuda.synchronize()    # the reads must be complete before the library's own stream looks at them
        c = _lib.Context(0, ctypes.c_void_p(stream.cuda_stream))
        c.attach(wl["packed"].data_ptr(), wl["total_bases"], wl["off"].data_ptr(), wl["n_reads"], None, keep=wl)
        ctxs.append(c); keep.append(wl); total_all += wl["total_bases"]
    per_rank = keep[0]["total_bases"]
    fs, nh = _lib.estimate_bloomfilter(total_all, K)
    distinct21 = genome + int(total_all * ERR * 21 * 1.05)
    table_slots = int(distinct21 / w / 0.55)
    owned_slots = int(genome * 1.2 / w / 0.5)
    solid_slots = int(min(genome, per_rank) * 1.2 / 0.5)
    comm = pdist.EmulatedComm(w)
    out = None
    for _ in range(args.steps):
        out = pdist.run_hot_path(ctxs, comm, K, fs, nh, table_slots, solid_slots, owned_slots, 1 << 25, dev)
    print(json.dumps({"world": w, "genome_bp": genome, "stage_ms": out[0]["stage_ms"], "count_sub_ms": out[0]["count_sub_ms"], "lap_ms": out[0]["lap_ms"],
                      "owned_solid": [s["owned_solid"] for s in out], "filter": out[0]["filter"], "exchange": out[0]["exchange"]}))
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    main()
