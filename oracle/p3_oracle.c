/* oracle/p3_oracle.c — TEST INFRASTRUCTURE ONLY (see p3_oracle.h).
 *
 * Plain-C CPU restatement of the reference's k-mer-to-graph hot path. Each function cites the
 * reference file:line it follows. Written for obviousness, not speed. Parity is pinned against
 * the unmodified reference (oracle/_ref/libp3ref.so) by tests/test_oracle_vs_ref.py and the
 * fixtures in tests/golden/.
 */
#define _GNU_SOURCE
#include "p3_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- base codes */
/* reference src/common.h:31-33 base_to_bit / trans_base are unordered_maps read with
 * operator[]: a character outside ACGT default-inserts 0, so it reads as 'A' on the forward
 * strand AND as code 0 on the reverse strand (trans_base[c] = '\0', base_to_bit['\0'] = 0).
 * GetFirstKmerForward/Backward (BitCalc.cpp:8-33) agree: no branch taken -> 0. */
static inline unsigned fwd_code(unsigned char c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 0; }
}
static inline unsigned rev_code(unsigned char c) {
    switch (c) { case 'A': return 3; case 'C': return 2; case 'G': return 1; case 'T': return 0; default: return 0; }
}

/* ---------------------------------------------------------------- multiword k-mers */
static inline int nwords(int k) { return (2 * k + 63) >> 6; }
static inline uint64_t topmask(int k) { int r = (2 * k) & 63; return r ? ((1ULL << r) - 1) : ~0ULL; }

/* (x << 2) | code on a std::bitset<2k> (bits shifted past 2k are dropped) */
static void shl2_or(uint64_t *a, int W, int k, unsigned code) {
    for (int i = W - 1; i > 0; i--) a[i] = (a[i] << 2) | (a[i - 1] >> 62);
    a[0] = (a[0] << 2) | code;
    a[W - 1] &= topmask(k);
}
/* (x >> 2) | (code << (2k-2)) */
static void shr2_or_top(uint64_t *a, int W, int k, unsigned code) {
    for (int i = 0; i < W - 1; i++) a[i] = (a[i] >> 2) | (a[i + 1] << 62);
    a[W - 1] >>= 2;
    int bit = 2 * k - 2;
    a[bit >> 6] |= (uint64_t)code << (bit & 63);
}
static int cmp_words(const uint64_t *a, const uint64_t *b, int W) {
    for (int i = W - 1; i >= 0; i--) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}

/* reference src/BitCalc.cpp:8-19 */
void p3o_first_kmer_forward(const char *s, int k, uint64_t *out) {
    int W = nwords(k);
    memset(out, 0, sizeof(uint64_t) * W);
    for (int i = 0; i < k; i++) shl2_or(out, W, k, fwd_code((unsigned char)s[i]));
}
/* reference src/BitCalc.cpp:22-33 */
void p3o_first_kmer_backward(const char *s, int k, uint64_t *out) {
    int W = nwords(k);
    memset(out, 0, sizeof(uint64_t) * W);
    for (int i = k - 1; i >= 0; i--) shl2_or(out, W, k, rev_code((unsigned char)s[i]));
}
/* reference src/BitCalc.cpp:36-45: walks bases from least significant, appends (~hi,~lo) */
void p3o_complement_kmer(const uint64_t *in, int k, uint64_t *out) {
    int W = nwords(k);
    uint64_t tmp[P3O_MAXW];
    memset(tmp, 0, sizeof(uint64_t) * W);
    for (int i = 0; i < k; i++) {
        unsigned code = (unsigned)((in[(2 * i) >> 6] >> ((2 * i) & 63)) & 3);
        shl2_or(tmp, W, k, 3u - code);
    }
    memcpy(out, tmp, sizeof(uint64_t) * W);
}
/* reference src/BitCalc.cpp:48-54: MSB-first compare, ties return Fw */
int p3o_compare_bit(const uint64_t *fw, const uint64_t *bw, int k) {
    return cmp_words(fw, bw, nwords(k)) <= 0 ? 0 : 1;
}
/* reference src/BitCalc.cpp:57-65 */
void p3o_string_kmer(const uint64_t *in, int k, char *out) {
    static const char b2c[4] = {'A', 'C', 'G', 'T'};
    for (int i = k - 1, j = 0; i >= 0; i--, j++)
        out[j] = b2c[(in[(2 * i) >> 6] >> ((2 * i) & 63)) & 3];
}

/* ---------------------------------------------------------------- hashing */
static inline uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }
/* libstdc++-v3 libsupc++/hash_bytes.cc, 64-bit size_t branch (GCC 13.3). */
uint64_t p3o_hash_bytes(const void *ptr, size_t len, uint64_t seed) {
    const uint64_t mul = (((uint64_t)0xc6a4a793UL) << 32) + (uint64_t)0x5bd1e995UL;
    const unsigned char *buf = (const unsigned char *)ptr;
    const size_t len_aligned = len & ~(size_t)0x7;
    const unsigned char *end = buf + len_aligned;
    uint64_t hash = seed ^ (len * mul);
    for (const unsigned char *p = buf; p != end; p += 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        const uint64_t data = shift_mix(w * mul) * mul;
        hash ^= data;
        hash *= mul;
    }
    if ((len & 0x7) != 0) {
        uint64_t data = 0;
        for (int n = (int)(len & 0x7) - 1; n >= 0; n--) data = (data << 8) + end[n];
        hash ^= data;
        hash *= mul;
    }
    hash = shift_mix(hash) * mul;
    hash = shift_mix(hash);
    return hash;
}
/* <bitset>:1719 — hashes ceil(2k/8) bytes of the word array, seed 0xc70f6907 */
uint64_t p3o_std_hash_kmer(const uint64_t *kmer, int k) {
    return p3o_hash_bytes(kmer, (size_t)(2 * k + 7) / 8, 0xc70f6907ULL);
}
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64(uint64_t k) { /* reference src/MyHash.cpp:12-19 */
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
/* reference src/MyHash.cpp:22-35 */
void p3o_double_hash(uint64_t h0, uint64_t out[2]) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = h0;
    uint64_t h2 = h1 ^ c2;
    h2 = rotl64(h2, 31); h1 ^= c1; h1 = rotl64(h1, 33);
    h1 += h2; h2 += h1; h1 ^= c2; h2 ^= c1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

/* reference src/Options.cpp:50-60 (filter_size==0 case; -m given keeps num_hashes=10) */
void p3o_estimate_bloomfilter(uint64_t all_bases, int k, uint64_t *filter_size, int *num_hashes) {
    const double error_rate = 0.0005; /* Options.cpp:15 */
    const double false_positive_rate = 1.0e-6;
    uint64_t item_number = (uint64_t)((double)all_bases * error_rate * (double)(uint32_t)k);
    uint64_t fs = (uint64_t)(((double)item_number * (-(log(false_positive_rate)))) / pow(log(2.0), 2));
    *filter_size = fs;
    *num_hashes = (int)(uint8_t)((log(2.0) * (double)fs) / (double)item_number);
}

/* ---------------------------------------------------------------- bloom filter */
/* reference src/bloomfilter.cpp:59-66 */
static inline uint64_t nth_hash(unsigned n, uint64_t a, uint64_t b, uint64_t size) { return (a + (uint64_t)n * b) % size; }
/* reference src/bloomfilter.cpp:69-74 */
void p3o_bf_add(uint8_t *bloom, uint64_t filter_size, int num_hashes, const uint64_t *kmer, int k) {
    uint64_t h[2];
    p3o_double_hash(p3o_std_hash_kmer(kmer, k), h);
    for (int n = 0; n < num_hashes; n++) {
        uint64_t b = nth_hash((unsigned)n, h[0], h[1], filter_size);
        bloom[b >> 3] |= (uint8_t)(1u << (b & 7));
    }
}
/* reference src/bloomfilter.cpp:77-86 */
int p3o_bf_possibly_contains(const uint8_t *bloom, uint64_t filter_size, int num_hashes,
                             const uint64_t *kmer, int k) {
    uint64_t h[2];
    p3o_double_hash(p3o_std_hash_kmer(kmer, k), h);
    for (int n = 0; n < num_hashes; n++) {
        uint64_t b = nth_hash((unsigned)n, h[0], h[1], filter_size);
        if (!((bloom[b >> 3] >> (b & 7)) & 1)) return 0;
    }
    return 1;
}
/* reference src/DeBruijnGraph.cpp:318-323 */
int p3o_is_recorded(const uint8_t *bloom, uint64_t filter_size, int num_hashes, const uint64_t *kmer, int k) {
    uint64_t bw[P3O_MAXW];
    p3o_complement_kmer(kmer, k, bw);
    const uint64_t *q = p3o_compare_bit(kmer, bw, k) == 0 ? kmer : bw;
    return p3o_bf_possibly_contains(bloom, filter_size, num_hashes, q, k);
}
/* reference src/DeBruijnGraph.cpp:326-345 */
int p3o_check_directions(const uint8_t *bloom, uint64_t filter_size, int num_hashes,
                         const uint64_t *kmer, int k, int ignored_direction) {
    int W = nwords(k), mask = 0;
    uint64_t adj[P3O_MAXW];
    for (int i = 0; i < 8; i++) {
        if (i == ignored_direction) continue;
        memcpy(adj, kmer, sizeof(uint64_t) * W);
        if (i < 4) shr2_or_top(adj, W, k, (unsigned)i);      /* back_shifted | X_left  */
        else shl2_or(adj, W, k, (unsigned)(i - 4));            /* front_shifted | X_right */
        if (p3o_is_recorded(bloom, filter_size, num_hashes, adj, k)) mask |= 1 << i;
    }
    return mask;
}

/* ---------------------------------------------------------------- Load */
typedef struct { char *name; size_t name_len; char *seq; size_t seq_len; size_t idx; } rec_t;
static int rec_cmp(const void *pa, const void *pb) {
    const rec_t *a = (const rec_t *)pa, *b = (const rec_t *)pb;
    size_t m = a->name_len < b->name_len ? a->name_len : b->name_len;
    int c = memcmp(a->name, b->name, m);
    if (c) return c;
    if (a->name_len != b->name_len) return a->name_len < b->name_len ? -1 : 1;
    return a->idx < b->idx ? -1 : (a->idx > b->idx);
}
static int rec_cmp_idx(const void *pa, const void *pb) {
    const rec_t *a = (const rec_t *)pa, *b = (const rec_t *)pb;
    return a->idx < b->idx ? -1 : (a->idx > b->idx);
}
/* reference src/Load.cpp:32-103 */
int64_t p3o_load_reads(const char *path, int k, char *seq_out, uint64_t *off, uint64_t *total_len,
                       uint64_t *all_bases_out) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)fsz + 1);
    if (fsz > 0 && fread(buf, 1, (size_t)fsz, f) != (size_t)fsz) { fclose(f); free(buf); return -1; }
    fclose(f);
    size_t n = (size_t)fsz;
    int mode = 0; /* Load.cpp:40-48: first byte of first line decides */
    if (n > 0 && buf[0] == '>') mode = 1; else if (n > 0 && buf[0] == '@') mode = 2;
    size_t cap = 1024, nrec = 0;
    rec_t *recs = (rec_t *)malloc(cap * sizeof(rec_t));
    uint64_t all_bases = 0;
    if (mode) {
        char *cur_seq = (char *)malloc(n + 1);
        size_t cur_len = 0; char *name = NULL; size_t name_len = 0; int have_name = 0;
        size_t pos = 0, line_cnt = 0;
        while (pos < n) { /* std::getline semantics */
            size_t e = pos;
            while (e < n && buf[e] != '\n') e++;
            char *line = buf + pos; size_t ll = e - pos;
            int is_header = (mode == 1) ? (ll > 0 && line[0] == '>') : (line_cnt % 4 == 0);
            if (is_header) {
                if (have_name && name_len > 0) { /* read_name != "" */
                    if (cur_len >= (size_t)k) {
                        if (nrec == cap) { cap *= 2; recs = (rec_t *)realloc(recs, cap * sizeof(rec_t)); }
                        recs[nrec].name = name; recs[nrec].name_len = name_len;
                        recs[nrec].seq = (char *)malloc(cur_len); memcpy(recs[nrec].seq, cur_seq, cur_len);
                        recs[nrec].seq_len = cur_len; recs[nrec].idx = nrec; nrec++;
                        all_bases += cur_len;
                    }
                    cur_len = 0;
                }
                name = line; name_len = ll; have_name = 1;
            } else if (mode == 1 || line_cnt % 4 == 1) {
                memcpy(cur_seq + cur_len, line, ll); cur_len += ll;
            }
            line_cnt++;
            pos = e + 1;
        }
        if (cur_len >= (size_t)k) { /* Load.cpp:71-74 / :99-102 (no read_name check) */
            if (nrec == cap) { cap *= 2; recs = (rec_t *)realloc(recs, cap * sizeof(rec_t)); }
            recs[nrec].name = name; recs[nrec].name_len = have_name ? name_len : 0;
            recs[nrec].seq = (char *)malloc(cur_len); memcpy(recs[nrec].seq, cur_seq, cur_len);
            recs[nrec].seq_len = cur_len; recs[nrec].idx = nrec; nrec++;
            all_bases += cur_len;
        }
        free(cur_seq);
    }
    /* unordered_map<string,string>: same name line -> last assignment wins */
    qsort(recs, nrec, sizeof(rec_t), rec_cmp);
    size_t kept = 0;
    for (size_t i = 0; i < nrec; i++) {
        int last = (i + 1 == nrec) || recs[i].name_len != recs[i + 1].name_len ||
                   memcmp(recs[i].name, recs[i + 1].name, recs[i].name_len) != 0;
        if (last) { rec_t t = recs[kept]; recs[kept] = recs[i]; recs[i] = t; kept++; }
    }
    for (size_t i = kept; i < nrec; i++) free(recs[i].seq);
    qsort(recs, kept, sizeof(rec_t), rec_cmp_idx);
    uint64_t tot = 0;
    for (size_t i = 0; i < kept; i++) {
        if (off) off[i] = tot;
        if (seq_out) memcpy(seq_out + tot, recs[i].seq, recs[i].seq_len);
        tot += recs[i].seq_len;
        free(recs[i].seq);
    }
    if (off) off[kept] = tot;
    if (total_len) *total_len = tot;
    if (all_bases_out) *all_bases_out = all_bases;
    free(recs); free(buf);
    return (int64_t)kept;
}

/* ---------------------------------------------------------------- CountShortKmer */
static void radix_sort_u64(uint64_t *a, uint64_t n, int bits) {
    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));
    const int RB = 11; const uint64_t NB = 1u << RB;
    uint64_t *cnt = (uint64_t *)malloc(sizeof(uint64_t) * NB);
    uint64_t *src = a, *dst = tmp;
    for (int sh = 0; sh < bits; sh += RB) {
        memset(cnt, 0, sizeof(uint64_t) * NB);
        for (uint64_t i = 0; i < n; i++) cnt[(src[i] >> sh) & (NB - 1)]++;
        uint64_t s = 0;
        for (uint64_t b = 0; b < NB; b++) { uint64_t c = cnt[b]; cnt[b] = s; s += c; }
        for (uint64_t i = 0; i < n; i++) dst[cnt[(src[i] >> sh) & (NB - 1)]++] = src[i];
        uint64_t *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, sizeof(uint64_t) * n);
    free(tmp); free(cnt);
}
/* canonical 21-mer stream of one read, reference src/Load.cpp:115-123 */
static uint64_t short_kmers_of_read(const char *r, uint64_t len, uint64_t *out) {
    const int sk = P3O_SHORTK;
    const uint64_t mask = (1ULL << (2 * sk)) - 1;
    uint64_t f = 0, b = 0, n = 0;
    for (int i = 0; i < sk; i++) f = ((f << 2) | fwd_code((unsigned char)r[i])) & mask;
    for (int i = sk - 1; i >= 0; i--) b = ((b << 2) | rev_code((unsigned char)r[i])) & mask;
    for (uint64_t i = sk - 1; i < len; i++) {
        if (i != (uint64_t)sk - 1) {
            f = ((f << 2) | fwd_code((unsigned char)r[i])) & mask;
            b = (b >> 2) | ((uint64_t)rev_code((unsigned char)r[i]) << (2 * sk - 2));
        }
        out[n++] = f <= b ? f : b; /* CompareBit(for,rev,42) */
    }
    return n;
}
uint64_t p3o_count_short_kmers(const char *seq, const uint64_t *off, uint64_t n_reads,
                               uint64_t *keys, uint64_t *counts) {
    uint64_t total = 0;
    for (uint64_t r = 0; r < n_reads; r++) total += (off[r + 1] - off[r]) - P3O_SHORTK + 1;
    uint64_t *all = (uint64_t *)malloc(sizeof(uint64_t) * (total ? total : 1));
    uint64_t n = 0;
    for (uint64_t r = 0; r < n_reads; r++) n += short_kmers_of_read(seq + off[r], off[r + 1] - off[r], all + n);
    radix_sort_u64(all, n, 2 * P3O_SHORTK);
    uint64_t d = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i;
        while (j < n && all[j] == all[i]) j++;
        if (keys) keys[d] = all[i];
        if (counts) counts[d] = j - i;
        d++; i = j;
    }
    free(all);
    return d;
}

/* ---------------------------------------------------------------- MakeBF */
/* reference src/MakeBloomFilter.cpp:8-22, monotonic deque incl. `int num` truncation */
uint64_t p3o_rmq(const uint64_t *v, uint64_t n, int x, uint64_t *out) {
    int64_t *di = (int64_t *)malloc(sizeof(int64_t) * (n ? n : 1));
    uint64_t *dv = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));
    uint64_t head = 0, tail = 0, m = 0;
    int num = 0;
    for (uint64_t i = 0; i < n; i++) {
        while (tail > head && dv[tail - 1] > v[i]) tail--;
        di[tail] = (int64_t)i; dv[tail] = v[i]; tail++;
        if (di[head] + x == (int64_t)i) head++;
        num = (int)dv[head];
        if ((int64_t)x - 1 <= (int64_t)i) out[m++] = (uint64_t)(int64_t)num;
    }
    free(di); free(dv);
    return m;
}
static uint64_t lookup_count(const uint64_t *keys, const uint64_t *counts, uint64_t n, uint64_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == key) ? counts[lo] : 0; /* KC[...] default-inserts 0 */
}
typedef void (*solid_cb)(void *ctx, uint64_t read, uint64_t pos, const uint64_t *fw, const uint64_t *canon);
/* reference src/MakeBloomFilter.cpp:43-86: per read, 21-mer coverage -> RMQ -> solid k-mers */
static void for_each_solid(const char *seq, const uint64_t *off, uint64_t n_reads, int k,
                           const uint64_t *keys, const uint64_t *counts, uint64_t n_keys,
                           solid_cb cb, void *ctx) {
    const int sk = P3O_SHORTK;
    int W = nwords(k);
    for (uint64_t r = 0; r < n_reads; r++) {
        const char *rd = seq + off[r];
        uint64_t len = off[r + 1] - off[r];
        uint64_t ncov = len - sk + 1;
        uint64_t *cov = (uint64_t *)malloc(sizeof(uint64_t) * ncov);
        uint64_t *canon21 = (uint64_t *)malloc(sizeof(uint64_t) * ncov);
        short_kmers_of_read(rd, len, canon21);
        for (uint64_t j = 0; j < ncov; j++) cov[j] = lookup_count(keys, counts, n_keys, canon21[j]);
        uint64_t *rmq = (uint64_t *)malloc(sizeof(uint64_t) * ncov);
        p3o_rmq(cov, ncov, k - sk + 1, rmq);
        uint64_t fw[P3O_MAXW], bw[P3O_MAXW];
        p3o_first_kmer_forward(rd, k, fw);
        p3o_first_kmer_backward(rd, k, bw);
        for (uint64_t i = (uint64_t)k - 1; i < len; i++) {
            if (i != (uint64_t)k - 1) {
                shl2_or(fw, W, k, fwd_code((unsigned char)rd[i]));
                shr2_or_top(bw, W, k, rev_code((unsigned char)rd[i]));
            }
            if (rmq[i - k + 1] >= P3O_COV_THRESHOLD)
                cb(ctx, r, i - k + 1, fw, p3o_compare_bit(fw, bw, k) == 0 ? fw : bw);
        }
        free(cov); free(canon21); free(rmq);
    }
}
typedef struct { uint8_t *bloom; uint64_t fs; int nh; int k; int64_t *seed; uint8_t *solid; const uint64_t *off; uint64_t adds; } mkbf_ctx;
static void mkbf_cb(void *vctx, uint64_t r, uint64_t pos, const uint64_t *fw, const uint64_t *canon) {
    mkbf_ctx *c = (mkbf_ctx *)vctx; (void)fw;
    p3o_bf_add(c->bloom, c->fs, c->nh, canon, c->k);
    c->adds++;
    if (c->seed && c->seed[r] < 0) c->seed[r] = (int64_t)pos; /* MakeBloomFilter.cpp:79-83 */
    if (c->solid) c->solid[c->off[r] + pos] = 1;
}
uint64_t p3o_make_bf(const char *seq, const uint64_t *off, uint64_t n_reads, int k,
                     const uint64_t *keys, const uint64_t *counts, uint64_t n_keys,
                     uint64_t filter_size, int num_hashes, uint8_t *bloom, int64_t *seed_pos,
                     uint8_t *solid) {
    mkbf_ctx c = {bloom, filter_size, num_hashes, k, seed_pos, solid, off, 0};
    if (seed_pos) for (uint64_t r = 0; r < n_reads; r++) seed_pos[r] = -1;
    for_each_solid(seq, off, n_reads, k, keys, counts, n_keys, mkbf_cb, &c);
    return c.adds;
}

typedef struct { uint64_t *buf; uint64_t n, cap; int W; } coll_ctx;
static void coll_cb(void *vctx, uint64_t r, uint64_t pos, const uint64_t *fw, const uint64_t *canon) {
    coll_ctx *c = (coll_ctx *)vctx; (void)r; (void)pos; (void)fw;
    if (c->n == c->cap) { c->cap = c->cap ? c->cap * 2 : 1024; c->buf = (uint64_t *)realloc(c->buf, sizeof(uint64_t) * c->cap * c->W); }
    memcpy(c->buf + c->n * c->W, canon, sizeof(uint64_t) * c->W);
    c->n++;
}
static int g_sortW;
static int kmer_cmp(const void *a, const void *b) { return cmp_words((const uint64_t *)a, (const uint64_t *)b, g_sortW); }
uint64_t p3o_solid_kmers(const char *seq, const uint64_t *off, uint64_t n_reads, int k,
                         const uint64_t *keys, const uint64_t *counts, uint64_t n_keys, uint64_t *out) {
    coll_ctx c = {NULL, 0, 0, nwords(k)};
    for_each_solid(seq, off, n_reads, k, keys, counts, n_keys, coll_cb, &c);
    g_sortW = c.W;
    qsort(c.buf, c.n, sizeof(uint64_t) * c.W, kmer_cmp);
    uint64_t d = 0;
    for (uint64_t i = 0; i < c.n; i++) {
        if (i && cmp_words(c.buf + i * c.W, c.buf + (i - 1) * c.W, c.W) == 0) continue;
        if (out) memcpy(out + d * c.W, c.buf + i * c.W, sizeof(uint64_t) * c.W);
        d++;
    }
    free(c.buf);
    return d;
}
