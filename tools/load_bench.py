#!/usr/bin/env python
"""Loader throughput on the host (no GPU needed): p3_load_file = parse FASTA/FASTQ + 2-bit staging.
  python tools/load_bench.py reads.fastq [k]          an existing file
  python tools/load_bench.py --synthetic 2000000      writes N reads of 150 bp to /tmp first
Set P3_LOAD_TIMING=1 to see parse / allocation / pack separately."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    from platanus3_b200 import _lib
    args = sys.argv[1:]
    k = 32
    if args and args[0] == "--synthetic":
        n = int(args[1]) if len(args) > 1 else 2_000_000
        path = "/tmp/p3_load_bench_%d.fastq" % n
        if not os.path.exists(path):
            rng = np.random.default_rng(0)
            letters = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (n, 150), dtype=np.uint8)]
            qual = b"I" * 150
            with open(path, "wb") as f:
                chunk = []
                for i in range(n):
                    chunk.append(b"@r%d\n" % i + letters[i].tobytes() + b"\n+\n" + qual + b"\n")
                    if len(chunk) == 100000:
                        f.write(b"".join(chunk))
                        chunk = []
                f.write(b"".join(chunk))
    else:
        path = args[0]
        if len(args) > 1:
            k = int(args[1])
    L = _lib.lib()
    size = os.path.getsize(path)
    for it in range(3):
        h = C.c_void_p()
        t0 = time.perf_counter()
        _lib.check(L.p3_load_file(path.encode(), k, C.byref(h)))
        t = time.perf_counter() - t0
        bases, reads = L.p3_reads_total_bases(h), L.p3_reads_count(h)
        L.p3_reads_free(h)
        print("run %d: %.3f s  %.0f MB/s of file  %.0f Mbases/s  (%d reads, %d bases, cores seen: %d)"
              % (it, t, size / t / 1e6, bases / t / 1e6, reads, bases, os.cpu_count()))


if __name__ == "__main__":
    main()
