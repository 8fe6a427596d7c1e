"""Synthetic read sets for bench.py's BASELINE.json configs, generated straight into the 2-bit staging
layout of include/platanus3_b200.h (torch is device-memory plumbing only).

Every random choice is a pure function of (seed, index) — a splitmix64 hash, evaluated with wrapping
64-bit integer arithmetic — so the same data set comes out of torch on the GPU, of numpy on the host
(`make_reads_numpy`, what the oracle / reference arm is fed) and of the C scale checker
(oracle/p3_scalecheck.c), bit for bit, and a rank of a multi-GPU run can generate just its slice of the
reads (`first_read`, `n_reads`):

  genome[p]            = H(gseed, 0, p) >> 62
  read i               start = (H(rseed, 1, i) >> 2) mod (G - L + 1), reverse strand iff H(rseed, 2, i) >> 63
  base j of read i     substituted iff (e >> 40) < round(rate * 2^24), by +1 + ((e >> 8 & 0xffffffff) * 3 >> 32),
                       e = H(rseed, 3, i * L + j)
  H(seed, s, x)        = mix(x + mix(4 * seed + s)),  mix = splitmix64's output function of (x + 0x9E3779B97F4A7C15)
"""
import numpy as np
import torch

_M64 = (1 << 64) - 1
_GOLD, _C1, _C2 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB


def _s64(x):
    """unsigned 64-bit constant as the signed value torch.int64 holds"""
    x &= _M64
    return x - (1 << 64) if x >> 63 else x


def mix_int(x):
    z = (x + _GOLD) & _M64
    z = ((z ^ (z >> 30)) * _C1) & _M64
    z = ((z ^ (z >> 27)) * _C2) & _M64
    return z ^ (z >> 31)


def _lsr(t, s):
    """logical shift right of an int64 tensor"""
    return (t >> s) & ((1 << (64 - s)) - 1)


def hash_torch(seed, stream, idx):
    """H(seed, stream, idx) for an int64 tensor idx (bit pattern of the unsigned value)"""
    z = idx + _s64(mix_int((4 * seed + stream) & _M64) + _GOLD)
    z = (z ^ _lsr(z, 30)) * _s64(_C1)
    z = (z ^ _lsr(z, 27)) * _s64(_C2)
    return z ^ _lsr(z, 31)


def hash_numpy(seed, stream, idx):
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64((mix_int((4 * seed + stream) & _M64) + _GOLD) & _M64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_C1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_C2)
        return z ^ (z >> np.uint64(31))


def n_reads_for(genome_bases, coverage, read_len):
    return max(16, int(round(genome_bases * coverage / read_len)) // 16 * 16)


def _shifts(device):
    return torch.arange(62, -2, -2, device=device, dtype=torch.int64)


def pack_codes(codes_flat):
    """uint8 codes (multiple of 32 long) -> int64 words, first base in the top 2 bits"""
    v = codes_flat.view(-1, 32).to(torch.int64)
    return (v << _shifts(codes_flat.device)).sum(dim=1)


def make_reads(genome_bases, coverage, read_len, error_rate, seed, device, chunk_reads=1 << 20,
               return_codes=False, read_seed=None, first_read=0, n_reads=None):
    """Uniform shotgun reads from a random genome, both strands, substitution errors.

    Returns dict(packed=int64[n_words+1], off=int64[n_reads+1], total_bases, n_reads[, codes]).
    The data set has n_reads_for(genome_bases, coverage, read_len) reads (a multiple of 16, so that every
    chunk packs into whole words when read_len*16 is a multiple of 32); first_read / n_reads select a slice.
    """
    assert (read_len * 16) % 32 == 0
    rseed = seed + 1 if read_seed is None else read_seed
    if n_reads is None:
        n_reads = n_reads_for(genome_bases, coverage, read_len) - first_read
    assert n_reads % 16 == 0 and first_read % 16 == 0
    total = n_reads * read_len
    n_words = total // 32
    packed = torch.zeros(n_words + 1, dtype=torch.int64, device=device)
    codes_out = torch.empty(total, dtype=torch.uint8, device=device) if return_codes else None
    ar = torch.arange(read_len, device=device, dtype=torch.int64)
    thr = int(round(error_rate * (1 << 24)))
    span = genome_bases - read_len + 1
    done = 0
    while done < n_reads:
        n = min(chunk_reads, n_reads - done)
        ridx = torch.arange(first_read + done, first_read + done + n, device=device, dtype=torch.int64)
        starts = _lsr(hash_torch(rseed, 1, ridx), 2) % span
        flip = _lsr(hash_torch(rseed, 2, ridx), 63) != 0
        # base j of a reverse-strand read is the complement of genome[start + L - 1 - j]
        gpos = torch.where(flip[:, None], starts[:, None] + (read_len - 1 - ar)[None, :], starts[:, None] + ar[None, :])
        codes = _lsr(hash_torch(seed, 0, gpos), 62)
        codes = torch.where(flip[:, None], 3 - codes, codes)
        if thr > 0:
            e = hash_torch(rseed, 3, ridx[:, None] * read_len + ar[None, :])
            err = _lsr(e, 40) < thr
            shift = 1 + (((_lsr(e, 8) & 0xFFFFFFFF) * 3) >> 32)
            codes = torch.where(err, (codes + shift) & 3, codes)
            del e, err, shift
        flat = codes.to(torch.uint8).reshape(-1).contiguous()
        w0 = done * read_len // 32
        packed[w0:w0 + flat.numel() // 32] = pack_codes(flat)
        if return_codes:
            codes_out[done * read_len:(done + n) * read_len] = flat
        done += n
        del codes, flat, starts, flip, gpos, ridx
    off = torch.arange(n_reads + 1, device=device, dtype=torch.int64) * read_len
    out = dict(packed=packed, off=off, total_bases=total, n_reads=n_reads, read_len=read_len)
    if return_codes:
        out["codes"] = codes_out
    return out


def make_reads_numpy(genome_bases, coverage, read_len, error_rate, seed, read_seed=None, first_read=0, n_reads=None,
                     chunk_reads=1 << 18):
    """the same data set on the host: (uint8 ASCII [n_reads * read_len], uint64 offsets [n_reads + 1])"""
    rseed = seed + 1 if read_seed is None else read_seed
    if n_reads is None:
        n_reads = n_reads_for(genome_bases, coverage, read_len) - first_read
    thr = int(round(error_rate * (1 << 24)))
    span = np.uint64(genome_bases - read_len + 1)
    ar = np.arange(read_len, dtype=np.uint64)
    out = np.empty(n_reads * read_len, np.uint8)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    for done in range(0, n_reads, chunk_reads):
        n = min(chunk_reads, n_reads - done)
        ridx = np.arange(first_read + done, first_read + done + n, dtype=np.uint64)
        starts = (hash_numpy(rseed, 1, ridx) >> np.uint64(2)) % span
        flip = (hash_numpy(rseed, 2, ridx) >> np.uint64(63)) != 0
        gpos = np.where(flip[:, None], starts[:, None] + (np.uint64(read_len - 1) - ar)[None, :], starts[:, None] + ar[None, :])
        codes = (hash_numpy(seed, 0, gpos) >> np.uint64(62)).astype(np.uint8)
        codes = np.where(flip[:, None], 3 - codes, codes)
        if thr > 0:
            e = hash_numpy(rseed, 3, ridx[:, None] * np.uint64(read_len) + ar[None, :])
            err = (e >> np.uint64(40)) < np.uint64(thr)
            shift = (1 + ((((e >> np.uint64(8)) & np.uint64(0xFFFFFFFF)) * np.uint64(3)) >> np.uint64(32))).astype(np.uint8)
            codes = np.where(err, (codes + shift) & 3, codes)
        out[done * read_len:(done + n) * read_len] = acgt[codes.reshape(-1)]
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return out, off
