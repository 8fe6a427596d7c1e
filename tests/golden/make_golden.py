"""Generates tests/golden/*.npz + *.fasta|fastq from the UNMODIFIED reference
(oracle/_ref/libp3ref.so, built by oracle/Makefile from /root/reference/src).

Run in the dev container:  python tests/golden/make_golden.py
Each fixture holds the input read file and what the reference computed from it:
  all_bases, filter_size, num_hashes           (main.cpp:22-23)
  keys, counts                                 shortk_database, sorted by key (Load.cpp:105)
  bloom                                        BF::m_bits packed LSB-first (MakeBloomFilter.cpp:25)
  seeds                                        sorted seed k-mers (MakeBloomFilter.cpp:79-83)
  probe_kmers, probe_masks                     CheckDirections answers (DeBruijnGraph.cpp:326)
  gfa                                          sorted lines of de_bruijn_graph.gfa with -t 1
  n_junctions, n_joints, n_straights           graph sizes after MakeDBG with -t 1
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _checkers import Ref  # noqa: E402
from platanus3_b200 import synth  # noqa: E402

CASES = [
    # name, k, m, genome, cov, read_len, err, kind
    # coverage 50 where the filter is auto-sized: the reference sizes it for all_bases*0.0005*k
    # items (Options.cpp:53), so at low coverage it saturates and MakeDBG never terminates
    ("k21_clean", 21, 0, 1200, 50, 60, 0.0, "fasta"),
    ("k25_err", 25, 0, 1500, 50, 80, 0.01, "fasta"),
    ("k32_err", 32, 0, 1500, 50, 100, 0.01, "fastq"),
    ("k32_m", 32, 200003, 2000, 10, 90, 0.02, "fasta"),
    ("k25_quirks", 25, 60013, 1500, 10, 70, 0.005, "fasta"),
    # k = 32 with runs of 64 T's and 64 A's inside the genome: the all-T 32-mer is 2^64 - 1 as an integer, the value
    # hash tables like to use as "empty"; it becomes a junction node here, counted by CountNodeCoverage in both orientations
    ("k32_homopoly", 32, 400009, 1800, 40, 120, 0.004, "fasta"),
    # multi-word k-mers, two of the values the reference's own Assemble_k offers (Assemble.cpp:38-41)
    ("k63_err", 63, 0, 1500, 50, 150, 0.005, "fasta"),
    ("k101_m", 101, 300007, 1500, 30, 250, 0.003, "fastq"),
    # 16 words per k-mer; long reads (Assemble.cpp:42)
    ("k501_m", 501, 400009, 3000, 30, 800, 0.002, "fasta"),
    # the largest k the reference offers: 94 words per k-mer (Assemble.cpp:45)
    ("k3001_m", 3001, 900001, 9000, 25, 4500, 0.001, "fastq"),
]


def build_reads(name, genome, cov, rl, err, seed):
    g = synth.random_genome(genome, seed)
    if name.endswith("homopoly"):
        g = g.copy()
        g[300:364] = 3          # 64 T's
        g[900:964] = 0          # 64 A's
        g[1400:1440] = 3        # a shorter T run: all-T 32-mers from a second context
    reads = synth.reads_as_bytes(synth.simulate_reads(g, cov, rl, err, seed + 1))
    names = [">read_%d" % i for i in range(len(reads))]
    if name.endswith("quirks"):
        rng = np.random.default_rng(seed)
        reads = [bytearray(r) for r in reads]
        for i in range(0, len(reads), 9):
            reads[i][int(rng.integers(0, len(reads[i])))] = ord("N")
        for i in range(4, len(reads), 13):
            reads[i][int(rng.integers(0, len(reads[i])))] = ord("g")
        reads[5] = reads[5][:20]          # shorter than k: dropped
        reads[6] = reads[6][:25]          # exactly k
        names[11] = names[3]              # duplicate name: last record wins, all_bases counts both
        reads = [bytes(r) for r in reads]
    return reads, names


def main():
    import multiprocessing as mp
    only = set(sys.argv[1:])       # optional: names of the cases to (re)generate
    for ci, case in enumerate(CASES):
        if only and case[0] not in only:
            continue
        p = mp.Process(target=one_case, args=(ci,) + case)
        p.start()
        p.join(600)
        if p.is_alive():
            p.kill()
            raise SystemExit("reference did not terminate on %s" % case[0])


def one_case(ci, name, k, m, genome, cov, rl, err, kind):
    if True:
        reads, names = build_reads(name, genome, cov, rl, err, 1000 + ci)
        path = os.path.join(HERE, "%s.%s" % (name, kind))
        if kind == "fasta":
            synth.write_fasta(path, reads, width=0 if ci % 2 else 37, names=names)
        else:
            synth.write_fastq(path, reads, names=["@" + n[1:] for n in names])
        ref = Ref(k, readfile=path, m=m, threads=1)
        ref.load_file()
        ref.estimate()
        keys, counts = ref.count_short()
        bloom, seeds = ref.make_bf()
        loaded = ref.reads()
        rng = np.random.default_rng(ci)
        probes = []
        for r in loaded[:: max(1, len(loaded) // 40)]:
            for p in range(0, len(r) - k + 1, 11 if k <= 1000 else 401):    # long k-mers: fewer, the fixture stays small
                probes.append(r[p:p + k].decode())
        probes += ["".join("ACGT"[i] for i in rng.integers(0, 4, k)) for _ in range(40 if k <= 1000 else 8)]
        probes = [p for p in probes if set(p) <= set("ACGT")]
        masks = np.array([ref.check_directions(p) for p in probes], np.uint8)
        ref.make_dbg()
        ref.count_node_coverage()
        with tempfile.TemporaryDirectory() as td:
            gfa = sorted(ref.print_graph(td))
        nj, njo, ns = ref.counts()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), k=k, m=m, all_bases=ref.all_bases, n_reads=ref.n_reads,
            filter_size=ref.filter_size, num_hashes=ref.num_hashes, keys=keys, counts=counts, bloom=bloom,
            seeds=np.array(seeds), probe_kmers=np.array(probes), probe_masks=masks, gfa=np.array(gfa),
            n_junctions=nj, n_joints=njo, n_straights=ns, read_file=os.path.basename(path))
        print("%-12s k=%d reads=%d distinct21=%d seeds=%d bloom_bits=%d junctions=%d joints=%d straights=%d gfa_lines=%d"
              % (name, k, ref.n_reads, len(keys), len(seeds), int(np.unpackbits(bloom).sum()), nj, njo, ns, len(gfa)))
        ref.close()


if __name__ == "__main__":
    main()
