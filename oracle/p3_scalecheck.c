/* oracle/p3_scalecheck.c — TEST INFRASTRUCTURE ONLY.
 *
 * The oracle's definitions (p3_oracle.c: hashes, filter sizing, BF.add / possiblyContains) over data
 * structures that scale to BASELINE.json's full sizes, so that the counts the timed GPU run reports at
 * configs[1] (and at the N-rank weak-scaling points) can be asserted, not just its throughput:
 *
 *   kmer_positions, distinct_21mers       CountShortKmer   reference src/Load.cpp:105-127
 *   bf_adds, solid_kmers                  MakeBF           reference src/MakeBloomFilter.cpp:25-89 (cov_threshold 2)
 *   filter_popcount, filter_xor           BF::m_bits       reference src/bloomfilter.cpp:69-74
 *   dbg_edges                             CheckDirections  reference src/DeBruijnGraph.cpp:326-345 over every distinct solid k-mer
 *
 * The reference itself (std::unordered_map<bitset<42>,uint64_t>, std::string reads, one thread) needs > 60 GB and hours
 * for 5 Gbp of reads; this program takes minutes. It is pinned to the oracle (and through it to the compiled
 * reference) by tests/test_scalecheck.py on inputs small enough for both. k <= 32, error-free alphabet (ACGT only).
 *
 * Data set: platanus3_b200/workload.py's hash-defined generator (genome G, coverage, read length L, substitution
 * rate, seed), so nothing has to be stored or shipped.
 *
 *   usage: p3_scalecheck G coverage L error_rate seed k [passes21]      -> one JSON line on stdout
 */
#include "p3_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define GOLD 0x9E3779B97F4A7C15ULL
static inline uint64_t mix(uint64_t x) {
    uint64_t z = x + GOLD;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t H(uint64_t seed, uint64_t stream, uint64_t x) { return mix(x + mix(4 * seed + stream)); }
static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

static uint64_t G, L, n_reads, seed, rseed;
static int K;
static uint8_t *codes;   /* n_reads * L base codes, one byte each (0..3) */

static void generate(double rate) {
    const uint64_t thr = (uint64_t)llround(rate * 16777216.0), span = G - L + 1;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < n_reads; i++) {
        const uint64_t start = (H(rseed, 1, i) >> 2) % span;
        const int flip = (int)(H(rseed, 2, i) >> 63);
        uint8_t *r = codes + i * L;
        for (uint64_t j = 0; j < L; j++) {
            uint64_t gp = flip ? start + (L - 1 - j) : start + j;
            unsigned c = (unsigned)(H(seed, 0, gp) >> 62);
            if (flip) c = 3 - c;
            if (thr) {
                uint64_t e = H(rseed, 3, i * L + j);
                if ((e >> 40) < thr) c = (c + 1 + (unsigned)((((e >> 8) & 0xFFFFFFFFULL) * 3) >> 32)) & 3;
            }
            r[j] = (uint8_t)c;
        }
    }
}

/* lock-free open addressing: slot = [count:22 | key:42], empty = all ones; counts saturate at 2 (only ">= 2" matters) */
#define KEY42 ((1ULL << 42) - 1)
static inline void table_add(uint64_t *t, uint64_t mask, uint64_t key) {
    uint64_t s = fmix64(key * 0x9E3779B97F4A7C15ULL + 1) & mask;
    for (;;) {
        uint64_t v = __atomic_load_n(t + s, __ATOMIC_RELAXED);
        if (v == ~0ULL) {
            uint64_t exp = ~0ULL;
            if (__atomic_compare_exchange_n(t + s, &exp, key | (1ULL << 42), 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return;
            v = exp;
        }
        if ((v & KEY42) == key) {
            if ((v >> 42) < 2) __atomic_fetch_add(t + s, 1ULL << 42, __ATOMIC_RELAXED);
            return;
        }
        s = (s + 1) & mask;
    }
}
static inline unsigned table_count(const uint64_t *t, uint64_t mask, uint64_t key) {
    uint64_t s = fmix64(key * 0x9E3779B97F4A7C15ULL + 1) & mask;
    for (;;) {
        uint64_t v = t[s];
        if (v == ~0ULL) return 0;
        if ((v & KEY42) == key) return (unsigned)(v >> 42);
        s = (s + 1) & mask;
    }
}
/* set of 64-bit canonical k-mers; empty = all ones (the all-T k-mer is never canonical). returns 1 when newly inserted */
static inline int set_add(uint64_t *t, uint64_t mask, uint64_t key) {
    uint64_t s = fmix64(key) & mask;
    for (;;) {
        uint64_t v = __atomic_load_n(t + s, __ATOMIC_RELAXED);
        if (v == ~0ULL) {
            uint64_t exp = ~0ULL;
            if (__atomic_compare_exchange_n(t + s, &exp, key, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return 1;
            v = exp;
        }
        if (v == key) return 0;
        s = (s + 1) & mask;
    }
}

static inline uint64_t revcomp(uint64_t v, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | (3 - (v & 3)); v >>= 2; }
    return r;
}

int main(int argc, char **argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s G coverage L error_rate seed k [passes21]\n", argv[0]); return 2; }
    G = strtoull(argv[1], 0, 10);
    const double cov = atof(argv[2]);
    L = strtoull(argv[3], 0, 10);
    const double rate = atof(argv[4]);
    seed = strtoull(argv[5], 0, 10); rseed = seed + 1;
    K = atoi(argv[6]);
    if (K < 21 || K > 32 || L < (uint64_t)K) { fprintf(stderr, "k in [21,32], L >= k\n"); return 2; }
    n_reads = (uint64_t)llround((double)G * cov / (double)L) / 16 * 16;   /* workload.n_reads_for */
    if (n_reads < 16) n_reads = 16;
    const uint64_t total = n_reads * L;
    const double t0 = omp_get_wtime();
    codes = (uint8_t *)malloc(total);
    if (!codes) { fprintf(stderr, "out of memory (reads)\n"); return 1; }
    generate(rate);
    fprintf(stderr, "[%.1fs] %llu reads generated, %d threads\n", omp_get_wtime() - t0, (unsigned long long)n_reads, omp_get_max_threads());

    /* ---- CountShortKmer: distinct canonical 21-mers and, per position, "count >= 2" -------------------------- */
    const int S = P3O_SHORTK;
    const uint64_t m21 = (1ULL << (2 * S)) - 1;
    const uint64_t pos21 = n_reads * (L - S + 1);
    const uint64_t est_distinct = G + (uint64_t)((double)total * rate * S * 1.1) + 1024;
    int passes = argc > 7 ? atoi(argv[7]) : (int)(est_distinct / 300000000ULL) + 1;     /* <= ~300 M keys per pass: 8 GB table */
    uint64_t cap = 1;
    while (cap < (est_distinct / passes) * 2 + 1024) cap <<= 1;
    uint64_t *tab = (uint64_t *)malloc(cap * 8);
    uint64_t *good = (uint64_t *)calloc((total + 63) / 64 + 1, 8);     /* bit per stream position: its 21-mer has count >= 2 */
    if (!tab || !good) { fprintf(stderr, "out of memory (table)\n"); return 1; }
    uint64_t distinct21 = 0;
    for (int q = 0; q < passes; q++) {
        memset(tab, 0xFF, cap * 8);
#pragma omp parallel for schedule(dynamic, 4096)
        for (uint64_t i = 0; i < n_reads; i++) {
            const uint8_t *r = codes + i * L;
            uint64_t fw = 0, bw = 0;
            for (uint64_t j = 0; j < L; j++) {
                fw = ((fw << 2) | r[j]) & m21;
                bw = (bw >> 2) | ((uint64_t)(3 - r[j]) << (2 * (S - 1)));
                if (j + 1 < (uint64_t)S) continue;
                const uint64_t key = fw < bw ? fw : bw;
                if ((int)(fmix64(key) % (uint64_t)passes) == q) table_add(tab, cap - 1, key);
            }
        }
        uint64_t d = 0;
#pragma omp parallel for reduction(+ : d) schedule(static)
        for (uint64_t s = 0; s < cap; s++) d += tab[s] != ~0ULL;
        distinct21 += d;
#pragma omp parallel for schedule(dynamic, 4096)
        for (uint64_t i = 0; i < n_reads; i++) {
            const uint8_t *r = codes + i * L;
            uint64_t fw = 0, bw = 0;
            for (uint64_t j = 0; j < L; j++) {
                fw = ((fw << 2) | r[j]) & m21;
                bw = (bw >> 2) | ((uint64_t)(3 - r[j]) << (2 * (S - 1)));
                if (j + 1 < (uint64_t)S) continue;
                const uint64_t key = fw < bw ? fw : bw;
                if ((int)(fmix64(key) % (uint64_t)passes) != q) continue;
                if (table_count(tab, cap - 1, key) >= P3O_COV_THRESHOLD) {
                    const uint64_t p = i * L + (j + 1 - S);
                    __atomic_fetch_or(good + (p >> 6), 1ULL << (p & 63), __ATOMIC_RELAXED);
                }
            }
        }
        fprintf(stderr, "[%.1fs] 21-mer pass %d/%d: %llu distinct so far\n", omp_get_wtime() - t0, q + 1, passes, (unsigned long long)distinct21);
    }
    free(tab);

    /* ---- MakeBF: solid positions (window minimum of the 21-mer counts >= 2), distinct solid k-mers, filter ------ */
    uint64_t fsize; int nh;
    p3o_estimate_bloomfilter(total, K, &fsize, &nh);     /* main.cpp:22-23: all_bases = every base loaded */
    uint8_t *bloom = (uint8_t *)calloc((fsize + 7) / 8 + 8, 1);
    const uint64_t est_solid = G + G / 4 + 1024;
    uint64_t scap = 1;
    while (scap < est_solid * 2) scap <<= 1;
    uint64_t *set = (uint64_t *)malloc(scap * 8);
    if (!bloom || !set) { fprintf(stderr, "out of memory (set)\n"); return 1; }
    memset(set, 0xFF, scap * 8);
    const uint64_t mk = K >= 32 ? ~0ULL : ((1ULL << (2 * K)) - 1);
    const int win = K - S + 1;
    uint64_t adds = 0, solid = 0;
#pragma omp parallel for reduction(+ : adds, solid) schedule(dynamic, 4096)
    for (uint64_t i = 0; i < n_reads; i++) {
        const uint8_t *r = codes + i * L;
        uint64_t fw = 0, bw = 0;
        int run = 0;                         /* consecutive good 21-mer positions ending at the current one */
        for (uint64_t j = 0; j < L; j++) {
            fw = ((fw << 2) | r[j]) & mk;
            bw = (bw >> 2) | ((uint64_t)(3 - r[j]) << (2 * (K - 1)));
            if (j + 1 >= (uint64_t)S) {
                const uint64_t p = i * L + (j + 1 - S);
                run = ((good[p >> 6] >> (p & 63)) & 1) ? run + 1 : 0;
            }
            if (j + 1 < (uint64_t)K) continue;
            /* the k-mer ending at j starts at j+1-K; its 21-mers start at j+1-K .. j+1-S: the last `win` positions */
            if (run < win) continue;
            adds++;
            uint64_t key = fw < bw ? fw : bw;
            if (set_add(set, scap - 1, key)) {
                solid++;
                uint64_t hh[2];
                p3o_double_hash(p3o_std_hash_kmer(&key, K), hh);
                for (int n = 0; n < nh; n++) {
                    const uint64_t bit = (hh[0] + (uint64_t)n * hh[1]) % fsize;
                    __atomic_fetch_or(bloom + (bit >> 3), (uint8_t)(1u << (bit & 7)), __ATOMIC_RELAXED);
                }
            }
        }
    }
    fprintf(stderr, "[%.1fs] MakeBF: %llu adds, %llu distinct solid k-mers\n", omp_get_wtime() - t0, (unsigned long long)adds, (unsigned long long)solid);
    free(good); free(codes);
    uint64_t pop = 0, fx = 0;
    const uint64_t fwords = (fsize + 63) / 64;     /* bytes beyond the filter are zero (calloc) */
    for (uint64_t w = 0; w < fwords; w++) {
        uint64_t v;
        memcpy(&v, bloom + 8 * w, 8);
        pop += (uint64_t)__builtin_popcountll(v);
        fx ^= v * (2 * w + 1);                      /* position-sensitive fold */
    }

    /* ---- CheckDirections of every distinct solid k-mer (canonical orientation) ------------------------------------- */
    uint64_t edges = 0;
#pragma omp parallel for reduction(+ : edges) schedule(dynamic, 65536)
    for (uint64_t s = 0; s < scap; s++) {
        const uint64_t km = set[s];
        if (km == ~0ULL) continue;
        for (int d = 0; d < 8; d++) {
            uint64_t nb = d < 4 ? ((km >> 2) | ((uint64_t)d << (2 * K - 2))) : (((km << 2) | (uint64_t)(d - 4)) & mk);
            const uint64_t rc = revcomp(nb, K);
            if (rc < nb) nb = rc;
            edges += (uint64_t)p3o_bf_possibly_contains(bloom, fsize, nh, &nb, K);
        }
    }
    fprintf(stderr, "[%.1fs] done\n", omp_get_wtime() - t0);
    printf("{\"genome_bp\": %llu, \"coverage\": %g, \"read_len\": %llu, \"error_rate\": %g, \"seed\": %llu, \"k\": %d, "
           "\"reads\": %llu, \"kmer_positions\": %llu, \"distinct_21mers\": %llu, \"bf_adds\": %llu, \"solid_kmers\": %llu, "
           "\"dbg_edges\": %llu, \"filter_size_bits\": %llu, \"num_hashes\": %d, \"filter_popcount\": %llu, \"filter_xor\": %llu}\n",
           (unsigned long long)G, cov, (unsigned long long)L, rate, (unsigned long long)seed, K, (unsigned long long)n_reads,
           (unsigned long long)pos21, (unsigned long long)distinct21, (unsigned long long)adds, (unsigned long long)solid,
           (unsigned long long)edges, (unsigned long long)fsize, nh, (unsigned long long)pop, (unsigned long long)fx);
    return 0;
}
