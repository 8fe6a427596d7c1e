// p3_host.cpp — host-only entry points of the C ABI: Bloom filter sizing and the 2-bit packer
// that turns ASCII reads into the pinned staging layout the kernels consume.
#include "../../include/platanus3_b200.h"

#include <cmath>
#include <cstring>

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
    static unsigned char lut[256];
    static bool init = false;
    if (!init) {
        memset(lut, 4, sizeof(lut));
        lut[(unsigned char)'A'] = 0; lut[(unsigned char)'C'] = 1; lut[(unsigned char)'G'] = 2; lut[(unsigned char)'T'] = 3;
        init = true;
    }
    uint64_t total = n_reads ? off[n_reads] : 0;
    uint64_t words = p3_packed_words(total);
    memset(packed, 0, sizeof(uint64_t) * words);
    if (nmask) memset(nmask, 0, sizeof(uint32_t) * words);
    int bad = 0;
    uint64_t full = total / 32;
    for (uint64_t w = 0; w < full; w++) {
        const unsigned char *p = (const unsigned char *)seq + 32 * w;
        uint64_t v = 0; uint32_t m = 0;
        for (int j = 0; j < 32; j++) {
            unsigned c = lut[p[j]];
            m = (m << 1) | (c >> 2);
            v = (v << 2) | (c & 3);
        }
        packed[w] = v;
        if (m) { bad = 1; if (nmask) nmask[w] = m; }
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
