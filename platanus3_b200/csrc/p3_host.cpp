// p3_host.cpp — host-only entry points of the C ABI: Bloom filter sizing and the 2-bit packer
// that turns ASCII reads into the pinned staging layout the kernels consume.
#include "../../include/platanus3_b200.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

// 8 bases per step without a table (x86 BMI2): with the first base in the most significant byte,
//   code  = ((x >> 1) & 3) ^ ((x >> 2) & 1) per byte      A 0x41 -> 0, C 0x43 -> 1, G 0x47 -> 2, T 0x54 -> 3
//   valid = byte == 0x41 + 2*lo + 6*hi + 11*(lo & hi)      (lo, hi = the two code bits; no carry between bytes)
// and PEXT gathers the 8 codes (16 bits) and the 8 "not ACGT" flags (8 bits) in base order.
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
static bool pack_have_bmi2() { return __builtin_cpu_supports("bmi2"); }
__attribute__((target("bmi2")))
static int pack_range_bmi2(const unsigned char *seq, uint64_t w0, uint64_t w1, uint64_t *packed, uint32_t *nmask) {
    const uint64_t ONES = 0x0101010101010101ULL;
    int bad = 0;
    for (uint64_t w = w0; w < w1; w++) {
        const unsigned char *p = seq + 32 * w;
        uint64_t v = 0; uint32_t m = 0;
        for (int q = 0; q < 4; q++) {
            uint64_t x;
            memcpy(&x, p + 8 * q, 8);
            x = __builtin_bswap64(x);
            const uint64_t lo = ((x >> 1) ^ (x >> 2)) & ONES;     // code bit 0
            const uint64_t hi = (x >> 2) & ONES;                  // code bit 1
            const uint64_t both = lo & hi;
            const uint64_t expect = 0x41 * ONES + (lo << 1) + (hi << 2) + (hi << 1) + (both << 3) + (both << 1) + both;
            const uint64_t d = x ^ expect;
            const uint64_t nz = (d | ((d & 0x7f7f7f7f7f7f7f7fULL) + 0x7f7f7f7f7f7f7f7fULL)) & 0x8080808080808080ULL;
            const uint64_t flags = _pext_u64(nz, 0x8080808080808080ULL);                       // 8 bits, first base on top
            uint64_t codes = _pext_u64(lo | (hi << 1), 0x0303030303030303ULL);                 // 16 bits, first base on top
            // a byte that is not ACGT packs as 0 (the reference's operator[] default)
            if (flags) {
                const uint64_t spread = _pdep_u64(flags, 0x5555ULL);      // flag of base i -> bit 2i
                codes &= ~(spread | (spread << 1));
            }
            v = (v << 16) | (codes & 0xFFFF);
            m = (m << 8) | (uint32_t)flags;
        }
        packed[w] = v;
        if (m) { bad = 1; if (nmask) nmask[w] = m; }
    }
    return bad;
}
#else
static bool pack_have_bmi2() { return false; }
static int pack_range_bmi2(const unsigned char *, uint64_t, uint64_t, uint64_t *, uint32_t *) { return 0; }
#endif

extern "C" {

// Options::EstimateBloomfilter, reference src/Options.cpp:50-60. Same double arithmetic in the
// same order; the uint8_t narrowing of num_hashes is the reference's member type (Options.cpp:11).
int p3_estimate_bloomfilter(uint64_t all_bases, uint32_t k, uint64_t *filter_size, uint32_t *num_hashes) {
    if (!filter_size || !num_hashes) return P3_ERR_ARG;
    const double error_rate = 0.0005;  // Options.cpp:15
    const double false_positive_rate = 1.0e-6;
    uint64_t item_number = (uint64_t)((double)all_bases * error_rate * (double)k);
    if (item_number == 0) return P3_ERR_ARG;  // the reference computes 0/0 here and later crashes on % 0
    uint64_t fs = (uint64_t)(((double)item_number * (-(std::log(false_positive_rate)))) / std::pow(std::log(2.0), 2));
    *filter_size = fs;
    *num_hashes = (uint32_t)(uint8_t)((std::log(2.0) * (double)fs) / (double)item_number);
    return fs ? P3_OK : P3_ERR_ARG;
}

uint64_t p3_packed_words(uint64_t total_bases) { return (total_bases + 31) / 32 + 1; }

// 2-bit staging. A=0 C=1 G=2 T=3 (reference src/BitCalc.cpp:9-16); anything else packs as 0 and
// sets its bit in the non-ACGT plane, which is how the reference reads it on the forward strand
// (common.h:32 operator[] default) — the plane lets the kernels also reproduce the reverse strand.
int p3_pack_reads(const char *seq, const uint64_t *off, uint64_t n_reads, uint64_t *packed,
                  uint32_t *nmask, int *has_non_acgt) {
    if (!off || !packed || (!seq && n_reads && off[n_reads])) return P3_ERR_ARG;
    // built once, thread-safely (function-local static with an initialiser: several host threads may pack at the same time)
    static const struct Lut {
        unsigned char v[256];
        Lut() { memset(v, 4, sizeof(v)); v[(unsigned char)'A'] = 0; v[(unsigned char)'C'] = 1; v[(unsigned char)'G'] = 2; v[(unsigned char)'T'] = 3; }
    } lut_holder;
    const unsigned char *lut = lut_holder.v;
    uint64_t total = n_reads ? off[n_reads] : 0;
    uint64_t words = p3_packed_words(total);
    memset(packed, 0, sizeof(uint64_t) * words);
    if (nmask) memset(nmask, 0, sizeof(uint32_t) * words);
    int bad = 0;
    uint64_t full = total / 32;
    // whole words in parallel (independent: 32 bases -> one packed word + one mask word)
    auto pack_range_lut = [&](uint64_t w0, uint64_t w1, int *bad_out) {
        int b = 0;
        for (uint64_t w = w0; w < w1; w++) {
            const unsigned char *p = (const unsigned char *)seq + 32 * w;
            uint64_t v = 0; uint32_t m = 0;
            for (int j = 0; j < 32; j++) {
                unsigned c = lut[p[j]];
                m = (m << 1) | (c >> 2);
                v = (v << 2) | (c & 3);
            }
            packed[w] = v;
            if (m) { b = 1; if (nmask) nmask[w] = m; }
        }
        *bad_out = b;
    };
    const bool bmi2 = pack_have_bmi2();
    auto pack_range = [&](uint64_t w0, uint64_t w1, int *bad_out) {
        if (bmi2) *bad_out = pack_range_bmi2((const unsigned char *)seq, w0, w1, packed, nmask);
        else pack_range_lut(w0, w1, bad_out);
    };
    unsigned n_thr = std::thread::hardware_concurrency();
    n_thr = std::max(1u, std::min(n_thr ? n_thr : 1u, 32u));
    if (full < (1u << 20)) n_thr = 1;
    if (n_thr == 1) {
        pack_range(0, full, &bad);
    } else {
        std::vector<std::thread> th;
        std::vector<int> bads(n_thr, 0);
        for (unsigned t = 0; t < n_thr; t++)
            th.emplace_back(pack_range, full * t / n_thr, full * (t + 1) / n_thr, &bads[t]);
        for (auto &x : th) x.join();
        for (int b : bads) bad |= b;
    }
    uint64_t rem = total - 32 * full;
    if (rem) {
        const unsigned char *p = (const unsigned char *)seq + 32 * full;
        uint64_t v = 0; uint32_t m = 0;
        for (uint64_t j = 0; j < rem; j++) {
            unsigned c = lut[p[j]];
            m = (m << 1) | (c >> 2);
            v = (v << 2) | (c & 3);
        }
        packed[full] = v << (2 * (32 - rem));
        m <<= (32 - rem);
        if (m) { bad = 1; if (nmask) nmask[full] = m; }
    }
    if (has_non_acgt) *has_non_acgt = bad;
    return P3_OK;
}

}  // extern "C"
