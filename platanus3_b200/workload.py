"""Synthetic read sets generated with torch (device memory plumbing only) straight into the
2-bit staging layout of include/platanus3_b200.h, for bench.py's BASELINE.json configs."""
import torch

_SHIFTS = None


def _shifts(device):
    return torch.arange(62, -2, -2, device=device, dtype=torch.int64)


def pack_codes(codes_flat):
    """uint8 codes (multiple of 32 long) -> int64 words, first base in the top 2 bits"""
    v = codes_flat.view(-1, 32).to(torch.int64)
    return (v << _shifts(codes_flat.device)).sum(dim=1)


def make_reads(genome_bases, coverage, read_len, error_rate, seed, device, chunk_reads=1 << 20,
               return_codes=False, read_seed=None):
    """Uniform shotgun reads from a random genome, both strands, substitution errors.

    Returns dict(packed=int64[n_words+1], off=int64[n_reads+1], total_bases, n_reads[, codes]).
    n_reads is rounded to a multiple of 16 so that every chunk packs into whole words when
    read_len*16 is a multiple of 32.
    """
    assert (read_len * 16) % 32 == 0
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    genome = torch.randint(0, 4, (genome_bases,), generator=gen, device=device, dtype=torch.uint8)
    if read_seed is not None:   # same genome on every rank, different reads
        gen.manual_seed(read_seed)
    n_reads = max(16, int(round(genome_bases * coverage / read_len)) // 16 * 16)
    total = n_reads * read_len
    n_words = total // 32
    packed = torch.zeros(n_words + 1, dtype=torch.int64, device=device)
    codes_out = torch.empty(total, dtype=torch.uint8, device=device) if return_codes else None
    ar = torch.arange(read_len, device=device, dtype=torch.int64)
    done = 0
    while done < n_reads:
        n = min(chunk_reads, n_reads - done)
        starts = torch.randint(0, genome_bases - read_len + 1, (n,), generator=gen, device=device, dtype=torch.int64)
        codes = genome[starts[:, None] + ar[None, :]]
        flip = torch.rand(n, generator=gen, device=device) < 0.5
        rc = (3 - codes).flip(1)
        codes = torch.where(flip[:, None], rc, codes)
        if error_rate > 0:
            err = torch.rand(codes.shape, generator=gen, device=device) < error_rate
            shift = torch.randint(1, 4, codes.shape, generator=gen, device=device, dtype=torch.uint8)
            codes = torch.where(err, (codes + shift) & 3, codes)
        flat = codes.reshape(-1).contiguous()
        w0 = done * read_len // 32
        packed[w0:w0 + flat.numel() // 32] = pack_codes(flat)
        if return_codes:
            codes_out[done * read_len:(done + n) * read_len] = flat
        done += n
        del codes, rc, flat, starts, flip
    off = torch.arange(n_reads + 1, device=device, dtype=torch.int64) * read_len
    out = dict(packed=packed, off=off, total_bases=total, n_reads=n_reads, read_len=read_len)
    if return_codes:
        out["codes"] = codes_out
    return out
