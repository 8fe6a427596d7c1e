// p3_device.cuh — device-side building blocks of the k-mer-to-graph hot path (sm_100a).
//
// Everything here is integer/byte work bound by HBM random access, so the design rules are:
// coalesced 2-bit read staging, one 256-bit load per 32-byte hash bucket, 64-bit atomics that
// stay in L2, and no per-read rolling state (each k-mer is cut straight out of two packed
// words with a funnel shift, so positions are independent and map 1:1 onto threads).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace p3 {

constexpr int kShortK = 21;                        // reference src/Options.cpp:14
constexpr uint64_t kKey42 = (1ULL << 42) - 1;      // canonical 21-mer = 42 bits
constexpr uint64_t kCntOne = 1ULL << 42;           // count lives in slot bits 42..63
constexpr uint64_t kCntFieldMax = 0x3FFFFFULL;     // 22-bit count field
constexpr uint64_t kEmpty = ~0ULL;                 // all-T is never canonical (rc = all-A = 0 is smaller)
constexpr int kOvfCap = 1024;                      // overflow side table (counts >= 2^22)

struct Stats {
    unsigned long long n_pos21;           // valid 21-mer positions counted
    unsigned long long n_distinct21;      // distinct canonical 21-mers
    unsigned long long n_adds;            // BF.add calls the reference would make
    unsigned long long n_distinct_solid;  // distinct canonical solid k-mers
    unsigned long long n_edges;           // adjacency bits set
    unsigned long long n_good21;          // distinct 21-mers with count >= threshold
    unsigned long long n_export;          // scratch counter for export kernels
    unsigned long long n_cand;            // keys created by the binned count (== distinct 21-mers)
    unsigned long long work;              // dynamic work counter of the sweep kernels
    unsigned int err_table_full;
    unsigned int err_ovf_full;
    unsigned int n_overflow;              // entries in the overflow table
    unsigned int err_bin_overflow;        // a fixed-capacity bin overflowed (deferred check: the host redoes the stage with exact bins)
    unsigned long long probes;            // bucket probes of the instrumented (PROBE_STATS) insert kernels
    unsigned int max_probe;               // longest probe chain seen by them
    unsigned int err_kbin_overflow;       // same deferred check for the k-mer bins of the binned de-duplication
    unsigned int err_peer_timeout;        // a multi-GPU device-side barrier gave up waiting for a peer
    unsigned int pad;
    unsigned long long owned_pos;         // multi-GPU owner: records inserted in the rounds before the current one
};

struct FastMod {  // x % d for a run-time invariant d (filter_size), reference bloomfilter.cpp:65
    uint64_t d, M;
};
__host__ __device__ inline FastMod make_fastmod(uint64_t d) {
    FastMod f; f.d = d; f.M = d ? (~0ULL / d) : 0; return f;
}
__device__ __forceinline__ uint64_t fastmod(uint64_t x, const FastMod &f) {
    uint64_t q = __umul64hi(x, f.M);     // q <= x/d, off by at most 2
    uint64_t r = x - q * f.d;
    while (r >= f.d) r -= f.d;
    return r;
}

// ---- hashing --------------------------------------------------------------------------------
// reference src/MyHash.cpp:12-19
__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
__device__ __forceinline__ uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }
// std::hash<std::bitset<2k>> for 2k <= 64: libstdc++ _Hash_bytes over ceil(2k/8) bytes with seed
// 0xc70f6907 (what reference src/MyHash.cpp:25 calls). One 8-byte block when nbytes == 8, else
// only the tail step (the tail bytes ARE the little-endian value of the k-mer).
__device__ __forceinline__ uint64_t std_hash_kmer1(uint64_t w, int nbytes) {
    const uint64_t mul = 0xc6a4a7935bd1e995ULL;
    uint64_t hash = 0xc70f6907ULL ^ ((uint64_t)nbytes * mul);
    if (nbytes == 8) {
        uint64_t data = shift_mix(w * mul) * mul;
        hash ^= data; hash *= mul;
    } else {
        hash ^= w; hash *= mul;
    }
    hash = shift_mix(hash) * mul;
    hash = shift_mix(hash);
    return hash;
}
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
// reference src/MyHash.cpp:22-35
__device__ __forceinline__ void double_hash(uint64_t h0, uint64_t &h1, uint64_t &h2) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    h1 = h0; h2 = h1 ^ c2;
    h2 = rotl64(h2, 31); h1 ^= c1; h1 = rotl64(h1, 33);
    h1 += h2; h2 += h1; h1 ^= c2; h2 ^= c1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
}

// ---- 2-bit k-mers ----------------------------------------------------------------------------
// reverse the order of the 32 two-bit groups of v
__device__ __forceinline__ uint64_t rev2(uint64_t v) {
    v = __brevll(v);
    return ((v >> 1) & 0x5555555555555555ULL) | ((v & 0x5555555555555555ULL) << 1);
}
// duplicate every bit of m: bit i -> bits 2i, 2i+1
__device__ __forceinline__ uint64_t spread32(uint32_t m) {
    uint64_t s = m;
    s = (s | (s << 16)) & 0x0000FFFF0000FFFFULL;
    s = (s | (s << 8)) & 0x00FF00FF00FF00FFULL;
    s = (s | (s << 4)) & 0x0F0F0F0F0F0F0F0FULL;
    s = (s | (s << 2)) & 0x3333333333333333ULL;
    s = (s | (s << 1)) & 0x5555555555555555ULL;
    return s | (s << 1);
}
// 64-bit window starting `o` bases into hi (hand-off of the neighbouring packed word)
__device__ __forceinline__ uint64_t window(uint64_t hi, uint64_t lo, int o) {
    return o ? ((hi << (2 * o)) | (lo >> (64 - 2 * o))) : hi;
}
__device__ __forceinline__ uint64_t kmask(int k) { return k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1); }
// canonical k-mer (CompareBit(Fw,Bw), reference src/BitCalc.cpp:48-54) from a left-aligned
// window x; m2 = spread non-ACGT plane of the same window (0 when the reads are clean): those
// bases read as code 0 on the reverse strand too (reference src/common.h:32-33 quirk).
__device__ __forceinline__ uint64_t canonical_from_window(uint64_t x, uint64_t m2, int k) {
    uint64_t f = x >> (64 - 2 * k);
    uint64_t r = rev2(~x & ~m2) & kmask(k);
    return f <= r ? f : r;
}
// GetComplementKmer (reference src/BitCalc.cpp:36-45) on a right-aligned k-mer
__device__ __forceinline__ uint64_t revcomp(uint64_t v, int k) { return rev2(~v) >> (64 - 2 * k); }

// ---- 32-byte bucket load ----------------------------------------------------------------------
__device__ __forceinline__ void ld_bucket(const uint64_t *p, uint64_t s[4]) {
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(s[0]), "=l"(s[1]), "=l"(s[2]), "=l"(s[3]) : "l"(p));
}
// same, but an L2 miss fills only the 64-byte half line that holds the bucket (default: all 128 B).
// For tables that stay DRAM resident (solid set, Bloom filter) the unused half is pure traffic:
// ncu showed makebf/adjacency byte-bound at 4.4-5.0 TB/s with 128-byte fills.
__device__ __forceinline__ void ld_bucket64(const uint64_t *p, uint64_t s[4]) {
    asm volatile("ld.global.cg.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(s[0]), "=l"(s[1]), "=l"(s[2]), "=l"(s[3]) : "l"(p));
}
__device__ __forceinline__ uint32_t ld_word64(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long *ull(uint64_t *p) { return reinterpret_cast<unsigned long long *>(p); }

// ---- overflow side table (counts that do not fit 22 bits) --------------------------------------
struct Ovf {
    uint64_t *keys;                  // kOvfCap, kEmpty when free
    unsigned long long *wraps;       // number of 2^22 wrap-arounds
};
__device__ inline void ovf_add(const Ovf &o, uint64_t key, Stats *st) {
    unsigned h = (unsigned)(fmix64(key) % kOvfCap);
    for (int i = 0; i < kOvfCap; i++) {
        unsigned s = (h + i) % kOvfCap;
        uint64_t cur = o.keys[s];
        if (cur == kEmpty) {
            cur = atomicCAS(ull(o.keys + s), kEmpty, key);
            if (cur == kEmpty) { atomicAdd(&st->n_overflow, 1u); cur = key; }
        }
        if (cur == key) { atomicAdd(o.wraps + s, 1ULL); return; }
    }
    atomicExch(&st->err_ovf_full, 1u);
}
__device__ inline uint64_t ovf_get(const Ovf &o, uint64_t key) {
    unsigned h = (unsigned)(fmix64(key) % kOvfCap);
    for (int i = 0; i < kOvfCap; i++) {
        unsigned s = (h + i) % kOvfCap;
        uint64_t cur = o.keys[s];
        if (cur == key) return o.wraps[s];
        if (cur == kEmpty) return 0;
    }
    return 0;
}

// ---- count table: [count:22 | key:42] slots, 4 per 32-byte bucket, linear bucket probing ------
// (the "MyHash" count table of BASELINE.json's north_star; replaces the reference's
//  std::unordered_map<std::bitset<42>,uint64_t> KmerCount, src/common.h:26)
__device__ __forceinline__ void count_bump(uint64_t *slot, uint64_t seen, uint64_t key, const Ovf &ovf, Stats *st) {
    if ((seen >> 42) < 0x200000ULL) {
        atomicAdd(ull(slot), kCntOne);  // RED: fewer than 2^21 adds can be in flight, cannot wrap
    } else {
        uint64_t old = atomicAdd(ull(slot), kCntOne);
        if ((old >> 42) == kCntFieldMax) ovf_add(ovf, key, st);  // field wrapped to 0
    }
}
// Table addressing. The table is split into P equal partitions of nbp buckets; a key lives in
// partition part(h) (top bits of h = fmix64(key)) and probes linearly INSIDE its partition
// starting at a bucket chosen from the low 32 bits of h. P = 1 is the plain table. With P > 1
// a sweep over records binned by partition touches one partition at a time, which then stays
// L2 resident (profiles/r01_randacc.md: 62-77 G inserts/s instead of 17.4 G/s from DRAM).
constexpr int kMaxParts = 1024;
// Longest probe sequence an insert walks before it reports "full": a (nearly) full partition must fail fast instead of
// crawling through every bucket for every remaining record. At the load factors in use (<= 0.6) the longest chains
// measured are ~12 buckets.
constexpr uint64_t kMaxProbe = 4096;
struct Table {
    uint64_t *slots;
    uint64_t nbp;   // buckets per partition (< 2^32)
    uint32_t P;     // partitions
    uint32_t sb;    // bits of a partition-relative slot index when the insert's index stream can also carry the record's
                    // offset-in-word (5 bits) and rank (4 bits) above it (sb + 9 <= 31), else 0 (index only)
};
__device__ __forceinline__ uint32_t part_of(uint64_t h, uint32_t P) { return (uint32_t)__umul64hi(h, (uint64_t)P); }
// owner GPU of a key (21-mer or k-mer) among n ranks: a hash independent of the table hash, so
// that the keys one rank owns still spread over all of its table partitions and buckets
__host__ __device__ inline uint64_t owner_mix(uint64_t k) {
    k ^= 0x5bd1e9955bd1e995ULL;
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
__device__ __forceinline__ uint32_t owner_of(uint64_t key, uint32_t n) { return (uint32_t)__umul64hi(owner_mix(key), (uint64_t)n); }
// record layouts
//   count record   : [rank:8 @47 | offset-in-word:5 @42 | key:42]  + uint32 word index
//   position record: [rank:8 @56 | stream position:56]
constexpr int kRecOffShift = 42, kRecRankShift = 47, kPosRankShift = 56;
// partition function of the tile-sort scatter kernels
//   0 = table partition of a count record   1 = owner of a count record
//   2 = rank field of a position record     3 = owner of a full 64-bit k-mer
//   4 = solid-set partition of a full 64-bit k-mer
template <int PMODE> __device__ __forceinline__ uint32_t pid_of(uint64_t rec, uint32_t P) {
    if (PMODE == 0) return part_of(fmix64(rec & kKey42), P);
    if (PMODE == 1) return owner_of(rec & kKey42, P);
    if (PMODE == 2) return (uint32_t)(rec >> kPosRankShift);
    if (PMODE == 4) return part_of(fmix64(rec), P);   // set partition of a full 64-bit k-mer (kset_part)
    return owner_of(rec, P);
}
__device__ __forceinline__ uint64_t sub_of(uint64_t h, uint64_t nbp) { return ((h & 0xFFFFFFFFULL) * nbp) >> 32; }

// Inserts one occurrence of key. Returns the count the key is KNOWN to have reached after this
// insert (a lower bound on the final count, because counts only grow): 1 when this call created
// the key, seen+1 otherwise; 0 on table-full. *created_slot = global slot index when this call
// created the key, else ~0.
__device__ __forceinline__ uint64_t count_insert(const Table &t, uint64_t key, const Ovf &ovf, Stats *st, uint64_t *created_slot) {
    const uint64_t h = fmix64(key);
    const uint64_t base = (uint64_t)part_of(h, t.P) * t.nbp;
    uint64_t b = sub_of(h, t.nbp);
    *created_slot = ~0ULL;
    for (uint64_t probe = 0; probe < t.nbp && probe < kMaxProbe; probe++) {
        uint64_t *bp = t.slots + 4 * (base + b);
        uint64_t s[4];
        ld_bucket(bp, s);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint64_t v = s[i];
            if ((v & kKey42) == key) { count_bump(bp + i, v, key, ovf, st); return (v >> 42) + 1; }
            if (v == kEmpty) {
                uint64_t old = atomicCAS(ull(bp + i), kEmpty, key | kCntOne);
                if (old == kEmpty) { *created_slot = 4 * (base + b) + i; return 1; }
                if ((old & kKey42) == key) { count_bump(bp + i, old, key, ovf, st); return (old >> 42) + 1; }
            }
        }
        b = (b + 1 == t.nbp) ? 0 : b + 1;
    }
    atomicExch(&st->err_table_full, 1u);
    return 0;
}
// KC[key] (0 when absent)
__device__ __forceinline__ uint64_t count_lookup(const Table &t, uint64_t key, const Ovf &ovf, unsigned n_overflow) {
    const uint64_t h = fmix64(key);
    const uint64_t base = (uint64_t)part_of(h, t.P) * t.nbp;
    uint64_t b = sub_of(h, t.nbp);
    for (uint64_t probe = 0; probe < t.nbp; probe++) {
        uint64_t s[4];
        ld_bucket(t.slots + 4 * (base + b), s);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint64_t v = s[i];
            if ((v & kKey42) == key) {
                uint64_t c = v >> 42;
                if (n_overflow) c += ovf_get(ovf, key) << 22;
                return c;
            }
            if (v == kEmpty) return 0;
        }
        b = (b + 1 == t.nbp) ? 0 : b + 1;
    }
    return 0;
}

// ---- solid k-mer set: 64-bit keys, 4 per bucket --------------------------------------------------
// Addressed like the count table: P partitions of nbp buckets, a key's partition from the top bits of
// fmix64(key), its bucket inside the partition from the low 32 bits, probing stays inside the
// partition. P = 1 is a plain table; with P > 1 the de-duplication sweep over k-mers binned by
// partition works in one ~24 MB, L2-resident window at a time.
struct KSet {
    uint64_t *slots;
    uint64_t nbp;   // buckets per partition (< 2^32)
    uint32_t P;     // partitions
};
__device__ __forceinline__ uint32_t kset_part(uint64_t key, uint32_t P) { return part_of(fmix64(key), P); }
// returns 1 = inserted now, 0 = already present, -1 = table full
__device__ __forceinline__ int set_insert(const KSet &t, uint64_t key) {
    const uint64_t h = fmix64(key);
    const uint64_t base = (uint64_t)part_of(h, t.P) * t.nbp;
    uint64_t b = sub_of(h, t.nbp);
    for (uint64_t probe = 0; probe < t.nbp && probe < kMaxProbe; probe++) {
        uint64_t *bp = t.slots + 4 * (base + b);
        uint64_t s[4];
        ld_bucket64(bp, s);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint64_t v = s[i];
            if (v == key) return 0;
            if (v == kEmpty) {
                uint64_t old = atomicCAS(ull(bp + i), kEmpty, key);
                if (old == kEmpty) return 1;
                if (old == key) return 0;
            }
        }
        b = (b + 1 == t.nbp) ? 0 : b + 1;
    }
    return -1;
}

// like set_insert, and *slot = global slot index of the key (inserted now or found)
__device__ __forceinline__ int set_insert_slot(const KSet &t, uint64_t key, uint64_t *slot) {
    const uint64_t h = fmix64(key);
    const uint64_t base = (uint64_t)part_of(h, t.P) * t.nbp;
    uint64_t b = sub_of(h, t.nbp);
    for (uint64_t probe = 0; probe < t.nbp && probe < kMaxProbe; probe++) {
        uint64_t *bp = t.slots + 4 * (base + b);
        uint64_t s[4];
        ld_bucket64(bp, s);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint64_t v = s[i];
            if (v == key) { *slot = 4 * (base + b) + i; return 0; }
            if (v == kEmpty) {
                uint64_t old = atomicCAS(ull(bp + i), kEmpty, key);
                if (old == kEmpty) { *slot = 4 * (base + b) + i; return 1; }
                if (old == key) { *slot = 4 * (base + b) + i; return 0; }
            }
        }
        b = (b + 1 == t.nbp) ? 0 : b + 1;
    }
    return -1;
}

// membership only (read-only table)
__device__ __forceinline__ bool set_contains(const KSet &t, uint64_t key) {
    const uint64_t h = fmix64(key);
    const uint64_t base = (uint64_t)part_of(h, t.P) * t.nbp;
    uint64_t b = sub_of(h, t.nbp);
    for (uint64_t probe = 0; probe < t.nbp; probe++) {
        uint64_t s[4];
        ld_bucket64(t.slots + 4 * (base + b), s);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (s[i] == key) return true;
            if (s[i] == kEmpty) return false;
        }
        b = (b + 1 == t.nbp) ? 0 : b + 1;
    }
    return false;
}

// ---- Bloom filter, reference src/bloomfilter.cpp ---------------------------------------------------
struct Bloom {
    uint32_t *bits;   // bit i = word i>>5 bit i&31 (== byte i>>3 bit i&7 of std::vector<bool> order)
    FastMod fm;       // filter_size
    int nh;           // num_hashes
    int nbytes;       // ceil(2k/8)
};
// BF::add, reference src/bloomfilter.cpp:69-74
__device__ __forceinline__ void bloom_add(const Bloom &bf, uint64_t canon) {
    uint64_t h1, h2;
    double_hash(std_hash_kmer1(canon, bf.nbytes), h1, h2);
    uint64_t x = h1;
    for (int n = 0; n < bf.nh; n++, x += h2) {
        uint64_t bit = fastmod(x, bf.fm);
        uint32_t m = 1u << (bit & 31);
        uint32_t *wp = bf.bits + (bit >> 5);
        if (!(__ldcg(wp) & m)) atomicOr(wp, m);   // bits only ever go 0->1, a stale 1 is still a 1
    }
}
// BF::possiblyContains, reference src/bloomfilter.cpp:77-86
__device__ __forceinline__ bool bloom_query(const Bloom &bf, uint64_t canon) {
    uint64_t h1, h2;
    double_hash(std_hash_kmer1(canon, bf.nbytes), h1, h2);
    uint64_t x = h1;
    for (int n = 0; n < bf.nh; n++, x += h2) {
        uint64_t bit = fastmod(x, bf.fm);
        if (!((ld_word64(bf.bits + (bit >> 5)) >> (bit & 31)) & 1u)) return false;
    }
    return true;
}
// (h1 + q*h2) mod 2^64 mod d for q = 0, 1, 2, ... advanced without a division: the remainder moves by
// h2 mod d per step and by -(2^64 mod d) whenever the reference's uint64 sum wraps
struct BloomStep {
    uint64_t x, h2, r, step, d, wrap;
    __device__ __forceinline__ void init(uint64_t h1, uint64_t h2_, const FastMod &fm, uint64_t wrap_) {
        x = h1; h2 = h2_; d = fm.d; wrap = wrap_;
        r = fastmod(h1, fm); step = fastmod(h2_, fm);
    }
    __device__ __forceinline__ void next() {
        uint64_t nx = x + h2;
        r += step;
        if (r >= d) r -= d;
        if (nx < x) r = r >= wrap ? r - wrap : r + d - wrap;
        x = nx;
    }
};
// IsRecorded, reference src/DeBruijnGraph.cpp:318-323 (oriented k-mer in, canonicalised here)
__device__ __forceinline__ bool is_recorded(const Bloom &bf, uint64_t kmer, int k) {
    uint64_t rc = revcomp(kmer, k);
    return bloom_query(bf, kmer <= rc ? kmer : rc);
}
// neighbour d of an oriented k-mer, reference src/DeBruijnGraph.cpp:327-339
__device__ __forceinline__ uint64_t neighbour(uint64_t kmer, int d, int k) {
    return d < 4 ? ((kmer >> 2) | ((uint64_t)d << (2 * k - 2))) : (((kmer << 2) | (uint64_t)(d - 4)) & kmask(k));
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned v) {
    return __reduce_add_sync(0xffffffffu, v);
}

}  // namespace p3
