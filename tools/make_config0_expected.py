#!/usr/bin/env python
"""Runs the UNMODIFIED reference (oracle/_ref/libp3ref.so: LoadFile, EstimateBloomfilter, CountShortKmer, MakeBF,
MakeDBG -t 1, CountNodeCoverage, PrintGraph) on the BASELINE.json configs[0] FASTA that bench.py --config 0 generates
(hash-defined generator, so the file is the same everywhere) and commits what it produced:
tests/golden/config0_expected.json = sha256 of the sorted GFA lines, line and node counts, the reference's stage times.
Dev container only (needs /root/reference compiled by oracle/Makefile). Test infrastructure."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench_configs as bc  # noqa: E402
from _checkers import Ref  # noqa: E402


def main():
    work = tempfile.mkdtemp(prefix="p3cfg0ref_")
    fa = os.path.join(work, "reads.fasta")
    n = bc.write_config0_fasta(fa)
    t0 = time.perf_counter()
    ref = Ref(bc.C0["k"], readfile=fa, threads=1)
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
    ref.print_graph(work)
    t6 = time.perf_counter()
    sha, lines = bc.gfa_digest(os.path.join(work, "de_bruijn_graph.gfa"))
    j, jo, s = ref.counts()
    out = dict(config=bc.C0, reads=n, gfa_sha256=sha, gfa_lines=lines, junctions=j, joints=jo, straights=s,
               ref_load_s=t1 - t0, ref_count_s=t2 - t1, ref_makebf_s=t3 - t2, ref_walk_s=t4 - t3, ref_coverage_s=t5 - t4,
               ref_print_s=t6 - t5, ref_total_s=t6 - t0, ref_threads=1, ref_host="dev container, %d cores" % (os.cpu_count() or 0))
    json.dump(out, open(bc.C0_EXPECTED, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
