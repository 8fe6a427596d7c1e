/* oracle/p3_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's k-mer-to-graph hot path, used ONLY as the
 * parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs. Nothing under platanus3_b200/ may include, link or call this.
 *
 * Parity is PINNED: tests/test_oracle_vs_ref.py checks every function here against the
 * unmodified reference compiled into oracle/_ref/libp3ref.so, and tests/golden/ holds
 * fixtures generated from that reference build (tests/golden/make_golden.py).
 *
 * k-mer representation: W = ceil(2k/64) little-endian uint64 words holding the same 2k-bit
 * integer as the reference's std::bitset<2k> (first base = most significant 2 bits;
 * A=0 C=1 G=2 T=3, reference src/BitCalc.cpp:8-19).
 */
#ifndef P3_ORACLE_H
#define P3_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P3O_MAXK 3001
#define P3O_MAXW 94
#define P3O_SHORTK 21 /* reference src/Options.cpp:14, src/MakeBloomFilter.cpp:27 */
#define P3O_COV_THRESHOLD 2 /* reference src/MakeBloomFilter.cpp:28 */

/* libstdc++ (GCC 13.3, the toolchain the reference Makefile uses) std::_Hash_bytes:
 * 64-bit Murmur-style hash behind std::hash<std::bitset<N>> (bits/functional_hash.h:204,
 * bitset:1719). Third-party to the reference; restated from the published algorithm. */
uint64_t p3o_hash_bytes(const void *ptr, size_t len, uint64_t seed);
/* std::hash<std::bitset<2k>>()(kmer) */
uint64_t p3o_std_hash_kmer(const uint64_t *kmer, int k);
/* reference src/MyHash.cpp:22-35 GetDoubleHash_64bit, from h0 = std::hash(kmer) */
void p3o_double_hash(uint64_t h0, uint64_t out[2]);

/* reference src/Options.cpp:50-60 */
void p3o_estimate_bloomfilter(uint64_t all_bases, int k, uint64_t *filter_size, int *num_hashes);

/* k-mer helpers (reference src/BitCalc.cpp) */
void p3o_first_kmer_forward(const char *s, int k, uint64_t *out);  /* :8-19  */
void p3o_first_kmer_backward(const char *s, int k, uint64_t *out); /* :22-33 */
void p3o_complement_kmer(const uint64_t *in, int k, uint64_t *out); /* :36-45 */
/* :48-54 — returns 0 if Fw chosen, 1 if Bw chosen */
int p3o_compare_bit(const uint64_t *fw, const uint64_t *bw, int k);
void p3o_string_kmer(const uint64_t *in, int k, char *out);         /* :57-65 */

/* reference src/Load.cpp:32-103: FASTA / single-line FASTQ by first byte; reads shorter
 * than k dropped; duplicate name lines collapse (last wins) but all_bases counts every one.
 * Returns number of reads kept, or -1 on open failure. seq/off may be NULL to size first:
 * *total_len receives the concatenated length. */
int64_t p3o_load_reads(const char *path, int k, char *seq, uint64_t *off, uint64_t *total_len,
                       uint64_t *all_bases);

/* reference src/Load.cpp:105-127 CountShortKmer(21). Reads are seq[off[i]..off[i+1]).
 * Output sorted by key; returns number of distinct canonical 21-mers. keys/counts may be
 * NULL to size first. */
uint64_t p3o_count_short_kmers(const char *seq, const uint64_t *off, uint64_t n_reads,
                               uint64_t *keys, uint64_t *counts);

/* reference src/MakeBloomFilter.cpp:8-22 RMQ (sliding-window minimum incl. its int cast) */
uint64_t p3o_rmq(const uint64_t *v, uint64_t n, int x, uint64_t *out);

/* reference src/MakeBloomFilter.cpp:25-89 MakeBF.
 *   keys/counts/n_keys : the CountShortKmer table (sorted by key)
 *   bloom              : (filter_size+7)/8 bytes, bit i at byte i>>3 bit i&7; caller zeroes
 *   seed_pos[r]        : position in read r of its first solid k-mer, or -1
 *   solid              : optional; one byte per k-mer start position, laid out per read at
 *                        seq offset off[r]+j (j < len-k+1), 1 = inserted
 * Returns the number of BF.add calls. */
uint64_t p3o_make_bf(const char *seq, const uint64_t *off, uint64_t n_reads, int k,
                     const uint64_t *keys, const uint64_t *counts, uint64_t n_keys,
                     uint64_t filter_size, int num_hashes, uint8_t *bloom, int64_t *seed_pos,
                     uint8_t *solid);

/* reference src/bloomfilter.cpp:69-86 on an already-canonical k-mer */
void p3o_bf_add(uint8_t *bloom, uint64_t filter_size, int num_hashes, const uint64_t *kmer, int k);
int p3o_bf_possibly_contains(const uint8_t *bloom, uint64_t filter_size, int num_hashes,
                             const uint64_t *kmer, int k);
/* reference src/DeBruijnGraph.cpp:318-323 IsRecorded (canonicalises with GetComplementKmer) */
int p3o_is_recorded(const uint8_t *bloom, uint64_t filter_size, int num_hashes,
                    const uint64_t *kmer, int k);
/* reference src/DeBruijnGraph.cpp:326-345 CheckDirections on an ORIENTED k-mer: bit i set iff
 * direction i (0-3 = left extension by A,C,G,T; 4-7 = right extension) is recorded.
 * ignored_direction as in the reference (-1 = none). */
int p3o_check_directions(const uint8_t *bloom, uint64_t filter_size, int num_hashes,
                         const uint64_t *kmer, int k, int ignored_direction);

/* Distinct canonical solid k-mers (the set MakeBF adds), sorted ascending as 2k-bit integers.
 * out holds n*W words (W = ceil(2k/64)); may be NULL to size first. Returns n. */
uint64_t p3o_solid_kmers(const char *seq, const uint64_t *off, uint64_t n_reads, int k,
                         const uint64_t *keys, const uint64_t *counts, uint64_t n_keys,
                         uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
