#!/usr/bin/env python
"""BASELINE.json configs[0] end to end: synthetic 4.6 Mbp random genome, 30x error-free reads,
k=32. Runs (1) the platanus3_b200 command line and (2) the unmodified reference's own pipeline
(oracle/_ref/libp3ref.so: LoadFile, EstimateBloomfilter, CountShortKmer, MakeBF, MakeDBG -t 1,
CountNodeCoverage, PrintGraph) on the same FASTA, compares the complete GFA as a set of lines and
prints one JSON line with the wall times. Test infrastructure (uses oracle/_ref)."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from platanus3_b200 import synth  # noqa: E402


def main():
    genome = int(sys.argv[1]) if len(sys.argv) > 1 else 4_600_000
    cov, rl, k = 30, 150, 32
    work = tempfile.mkdtemp(prefix="p3cfg0_")
    fa = os.path.join(work, "reads.fasta")
    g = synth.random_genome(genome, 7)
    reads = synth.reads_as_bytes(synth.simulate_reads(g, cov, rl, 0.0, 8))
    synth.write_fasta(fa, reads)
    out = {"config": "configs[0]: synthetic %d bp genome, %dx error-free %d bp reads, k=%d" % (genome, cov, rl, k),
           "reads": len(reads)}

    d1 = os.path.join(work, "b200"); os.makedirs(d1)
    exe = os.path.join(ROOT, "platanus3_b200", "platanus3_b200")
    t = time.perf_counter()
    r = subprocess.run([exe, "-i", fa, "-k", str(k), "-t", "1"], cwd=d1, capture_output=True, text=True)
    out["b200_cli_s"] = time.perf_counter() - t
    out["b200_cli_rc"] = r.returncode
    out["b200_cli_msg"] = r.stderr.strip()[-300:]
    mine = sorted(open(os.path.join(d1, "de_bruijn_graph.gfa")).read().splitlines()) if r.returncode == 0 else []

    if "--no-ref" not in sys.argv:
        from _checkers import Ref
        d2 = os.path.join(work, "ref"); os.makedirs(d2)
        t0 = time.perf_counter()
        ref = Ref(k, readfile=fa, threads=1)
        ref.load_file(); ref.estimate()
        t1 = time.perf_counter()
        ref.count_short()
        t2 = time.perf_counter()
        ref.make_bf()
        t3 = time.perf_counter()
        ref.make_dbg()
        t4 = time.perf_counter()
        ref.count_node_coverage()
        t5 = time.perf_counter()
        theirs = sorted(ref.print_graph(d2))
        t6 = time.perf_counter()
        out.update(ref_load_s=t1 - t0, ref_count_s=t2 - t1, ref_makebf_s=t3 - t2, ref_walk_s=t4 - t3,
                   ref_coverage_s=t5 - t4, ref_print_s=t6 - t5, ref_total_s=t6 - t0,
                   ref_nodes=dict(zip(("junctions", "joints", "straights"), ref.counts())),
                   gfa_lines=len(theirs), gfa_identical=(mine == theirs))
        if mine != theirs:
            out["first_diff"] = next(((a, b) for a, b in zip(mine, theirs) if a != b), None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
