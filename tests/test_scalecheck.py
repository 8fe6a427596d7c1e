"""oracle/p3_scalecheck (the oracle's definitions over data structures that scale to the full bench
workloads) against the oracle proper and the compiled reference, on inputs small enough for all three; and
the hash-defined workload generator: torch, numpy and C produce the same reads."""
import json
import os
import subprocess

import numpy as np
import pytest

from platanus3_b200 import workload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "p3_scalecheck")


def filter_checksum(bits):
    """popcount and position-sensitive xor fold of the filter bytes, as p3_scalecheck.c and bench.py compute them"""
    b = np.concatenate([bits, np.zeros((-len(bits)) % 8, np.uint8)]).view("<u8")
    pop = int(np.unpackbits(bits).sum())
    with np.errstate(over="ignore"):
        fx = np.bitwise_xor.reduce(b * (2 * np.arange(len(b), dtype=np.uint64) + np.uint64(1)))
    return pop, int(fx)


@pytest.mark.parametrize("genome,cov,rl,err,seed,k", [(30000, 40, 100, 0.01, 5, 32), (20000, 30, 80, 0.0, 6, 25), (25000, 50, 150, 0.02, 7, 21)])
def test_scalecheck_matches_oracle_and_reference(oracle, genome, cov, rl, err, seed, k):
    from _checkers import Ref, have_ref, words_to_kmer_str
    out = subprocess.run([EXE, str(genome), str(cov), str(rl), str(err), str(seed), str(k), "3"], capture_output=True, text=True, check=True)
    got = json.loads(out.stdout)
    seq, off = workload.make_reads_numpy(genome, cov, rl, err, seed)
    assert got["reads"] == len(off) - 1
    fs, nh = oracle.estimate_bloomfilter(int(off[-1]), k)
    keys, counts = oracle.count_short_kmers(seq, off)
    bits, _, _, adds = oracle.make_bf(seq, off, k, keys, counts, fs, nh)
    solid = oracle.solid_kmers(seq, off, k, keys, counts)
    edges = sum(bin(oracle.check_directions(bits, fs, nh, solid[i], k)).count("1") for i in range(len(solid)))
    pop, fx = filter_checksum(bits)
    want = dict(kmer_positions=int(counts.sum()), distinct_21mers=len(keys), bf_adds=adds, solid_kmers=len(solid), dbg_edges=edges,
                filter_size_bits=fs, num_hashes=nh, filter_popcount=pop, filter_xor=fx)
    assert {kk: got[kk] for kk in want} == want
    if have_ref() and k in (21, 25, 32):      # the same numbers from the unmodified reference
        ref = Ref(k, threads=1)
        ref.add_reads_arrays(seq, off)
        ref.estimate()
        rkeys, rcounts = ref.count_short()
        rbits, _ = ref.make_bf()
        assert len(rkeys) == got["distinct_21mers"] and filter_checksum(rbits) == (got["filter_popcount"], got["filter_xor"])
        km = np.frombuffer("".join(words_to_kmer_str(x, k) for x in solid).encode(), np.uint8)
        radj = ref.check_directions_batch(km, len(solid))
        assert int(np.unpackbits(radj).sum()) == got["dbg_edges"]
        ref.close()


def test_workload_generator_is_the_same_everywhere():
    """torch (any device) and numpy evaluate the same hash-defined data set; rank slices concatenate to the whole"""
    import torch
    t = workload.make_reads(50000, 20, 96, 0.01, 7, "cpu", return_codes=True, chunk_reads=64)
    seq, off = workload.make_reads_numpy(50000, 20, 96, 0.01, 7, chunk_reads=100)
    assert np.array_equal(np.frombuffer(b"ACGT", np.uint8)[t["codes"].numpy()], seq)
    assert t["n_reads"] == len(off) - 1 == workload.n_reads_for(50000, 20, 96)
    parts = [workload.make_reads(50000, 20, 96, 0.01, 7, "cpu", return_codes=True, first_read=f, n_reads=n)["codes"]
             for f, n in ((0, 160), (160, 4096), (4256, t["n_reads"] - 4256))]
    assert torch.equal(torch.cat(parts), t["codes"])
    idx = torch.arange(0, 5000, dtype=torch.int64)
    a = workload.hash_torch(1234, 3, idx).numpy().view(np.uint64)
    assert np.array_equal(a, workload.hash_numpy(1234, 3, np.arange(5000, dtype=np.uint64)))
    assert int(a[17]) == workload.mix_int((17 + workload.mix_int(4 * 1234 + 3)) & ((1 << 64) - 1))
    # substitution rate and strand balance are what they say
    g = workload.make_reads_numpy(200000, 5, 100, 0.0, 3)[0]
    e = workload.make_reads_numpy(200000, 5, 100, 0.05, 3)[0]
    assert abs((g != e).mean() - 0.05) < 0.003
