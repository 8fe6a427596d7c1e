"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol the
header declares, and its host-only entry points (packer, filter sizing) match the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as ge
from _checkers import reads_to_arrays
from platanus3_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    ge.build()


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "platanus3_b200.h")).read()
    declared = set(re.findall(r"\b(p3_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = C.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s


def test_estimate_matches_oracle(oracle):
    rng = np.random.default_rng(1)
    for ab in [10 ** 4, 10 ** 6, 138 * 10 ** 6, 5 * 10 ** 9, 93 * 10 ** 9] + [int(x) for x in rng.integers(5000, 2 ** 38, 100)]:
        for k in (21, 25, 32):
            assert _lib.estimate_bloomfilter(ab, k) == oracle.estimate_bloomfilter(ab, k)


def test_estimate_rejects_degenerate_input():
    with pytest.raises(_lib.P3Error):
        _lib.estimate_bloomfilter(10, 21)  # item_number == 0: the reference divides by zero


def _unpack(packed, total):
    b = np.unpackbits(packed.astype(">u8").view(np.uint8)).reshape(-1, 2)
    return (b[:, 0] * 2 + b[:, 1])[:total]


def test_pack_reads_layout():
    rng = np.random.default_rng(2)
    for total in (0, 1, 31, 32, 33, 64, 1000, 4097):
        codes = rng.integers(0, 4, total).astype(np.uint8)
        seq = synth.codes_to_ascii(codes)
        off = np.array([0, total], np.uint64) if total else np.zeros(1, np.uint64)
        packed, nmask = _lib.pack_reads(seq, off)
        assert len(packed) == (total + 31) // 32 + 1 and nmask is None
        assert np.array_equal(_unpack(packed, total), codes)


def test_pack_reads_non_acgt_plane():
    seq = np.frombuffer(b"ACGTNacgtRYACGTACGTACGTACGTACGTACGTACGTN", np.uint8)
    off = np.array([0, len(seq)], np.uint64)
    packed, nmask = _lib.pack_reads(seq, off)
    assert nmask is not None
    bits = np.unpackbits(nmask.astype(">u4").view(np.uint8))[: len(seq)]
    want = np.array([c not in b"ACGT" for c in seq.tobytes()], np.uint8)
    assert np.array_equal(bits, want)
    codes = _unpack(packed, len(seq))
    assert np.all(codes[want == 1] == 0)


def test_no_gpu_fails_loudly():
    if _lib.lib().p3_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.P3Error):
        _lib.Context(0)


def test_binding_example_compiles_against_the_header(tmp_path):
    """examples/reference_binding.cpp is the stub of INTEGRATION.md as a real program: it must build
    with nothing but include/platanus3_b200.h and the shared library, and on a box without a GPU it must
    stop at p3_create with the 'no CPU fallback' message (not crash, not produce a graph)"""
    import subprocess
    exe = str(tmp_path / "reference_binding")
    pkg = os.path.join(ROOT, "platanus3_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "reference_binding.cpp"), "-o", exe,
                        "-L" + pkg, "-lplatanus3_b200", "-Wl,-rpath," + pkg], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if _lib.lib().p3_device_count() > 0:
        return
    gold = os.path.join(ROOT, "tests", "golden", "k25_err.fasta")
    r = subprocess.run([exe, gold, "25", str(tmp_path / "x.gfa")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
    assert not os.path.exists(tmp_path / "x.gfa")
