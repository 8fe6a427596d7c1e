"""Differential test of the host half of the drop-in (Load, MakeDBG in -t 1 order, CountNodeCoverage,
PrintGraph; platanus3_b200/csrc/p3_assemble.cpp) against the UNMODIFIED reference compiled here
(oracle/_ref): many small random read sets built to produce awkward graphs — tandem repeats and
palindromes (cycles, k-mers equal to their own reverse complement), low coverage (tips, gaps),
substitution errors (bubbles), non-ACGT bytes, tight `-m` filters (Bloom false positives on the
walk) — for single-word and multi-word k. The CheckDirections table handed to the walk comes from
the oracle, so the test needs no GPU. Skipped where the reference is not compiled (the GPU box)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from _checkers import have_ref
from platanus3_b200 import _lib, synth
from test_host_side import _closed_table_from_oracle

pytestmark = pytest.mark.skipif(not have_ref(), reason="needs oracle/_ref (the compiled reference)")


def _reads(rng, k):
    """a small genome with planted repeats / palindromes, shotgun reads with errors and stray bytes"""
    n = int(rng.integers(3 * k, 12 * k))
    g = rng.integers(0, 4, n).astype(np.uint8)
    style = int(rng.integers(0, 4))
    if style == 1:                                  # tandem repeat longer than k
        unit = g[: int(rng.integers(2, k))]
        g = np.concatenate([g[: n // 3], np.tile(unit, 2 + 2 * k // len(unit)), g[n // 3:]])
    elif style == 2:                                # inverted repeat: a long palindromic stretch
        arm = g[: int(rng.integers(k, 2 * k))]
        g = np.concatenate([g[: n // 2], arm, (3 - arm)[::-1], g[n // 2:]])
    elif style == 3:                                # the same segment twice, far apart
        seg = g[: int(rng.integers(k + 2, 3 * k))]
        g = np.concatenate([g, rng.integers(0, 4, 2 * k).astype(np.uint8), seg])
    rl = int(rng.integers(k + 5, 3 * k + 20))
    cov = float(rng.choice([6, 12, 30]))
    err = float(rng.choice([0.0, 0.0, 0.004, 0.01]))
    reads = [bytearray(r) for r in synth.reads_as_bytes(synth.simulate_reads(g, cov, min(rl, len(g)), err, int(rng.integers(1 << 30))))]
    for r in reads[:: max(1, len(reads) // 5)]:     # stray bytes: read as A on both strands by the reference
        if rng.integers(0, 3) == 0:
            r[int(rng.integers(0, len(r)))] = int(rng.choice(list(b"Nnx")))
    return [bytes(r) for r in reads]


def _reference_run(path, k, m, out_path):
    from _checkers import Ref
    import tempfile
    ref = Ref(k, readfile=path, m=m, threads=1)
    ref.load_file()
    ref.estimate()
    keys, counts = ref.count_short()
    bloom, seeds = ref.make_bf()
    ref.make_dbg()
    ref.count_node_coverage()
    with tempfile.TemporaryDirectory() as td:
        gfa = sorted(ref.print_graph(td))
    with open(out_path, "wb") as f:
        pickle.dump(dict(k=k, filter_size=ref.filter_size, num_hashes=ref.num_hashes, keys=keys, counts=counts, bloom=bloom,
                         seeds=np.array(seeds), gfa=gfa, nodes=ref.counts()), f)


def _reference_in_fresh_process(path, k, m, out_path, timeout=30):
    """the reference loops forever on a saturated filter, so it runs in its own process under a timeout — a FRESH
    interpreter, not a fork: by the time this test runs the pytest process has OpenMP / torch threads, and a forked
    child that enters libgomp deadlocks"""
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([here, os.path.dirname(here), os.environ.get("PYTHONPATH", "")]))
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), path, str(k), str(m), out_path], env=env, timeout=timeout,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    except subprocess.TimeoutExpired:
        return None
    if r.returncode != 0:
        raise RuntimeError("reference run failed: " + r.stderr.decode(errors="replace")[-2000:])
    with open(out_path, "rb") as f:
        return pickle.load(f)


@pytest.mark.parametrize("k", [21, 25, 32, 33, 63])
def test_host_walk_matches_reference_on_awkward_graphs(oracle, tmp_path, k):
    done = 0
    for trial in range(9):
        rng = np.random.default_rng(7919 * k + trial)
        reads = _reads(rng, k)
        path = str(tmp_path / ("fz%d_%d.fasta" % (k, trial)))
        synth.write_fasta(path, reads, width=int(rng.choice([0, 37])))
        n_kmers = sum(max(0, len(r) - k + 1) for r in reads)
        # explicit filter: large enough that the reference's walk terminates, small enough for false positives
        m = int(n_kmers * float(rng.choice([1.5, 3, 8]))) + 1009
        g = _reference_in_fresh_process(path, k, m, str(tmp_path / "ref.pkl"))
        if g is None:
            continue
        kk, aa, ss, n_solid = _closed_table_from_oracle(oracle, g, path)
        gfa = str(tmp_path / "fz.gfa")
        st = _lib.walk_table(path, k, kk, aa, ss, gfa_path=gfa)
        assert (st["junctions"], st["joints"], st["straights"]) == tuple(int(x) for x in g["nodes"]), (k, trial)
        assert sorted(open(gfa).read().splitlines()) == g["gfa"], (k, trial)
        done += 1
    assert done >= 5      # the odd non-terminating reference run is skipped, not most of them


if __name__ == "__main__":
    _reference_run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
