/* platanus3_b200.h — C ABI of the B200-native k-mer-to-graph hot path.
 *
 * Drop-in boundary for platanus3's Load -> CountShortKmer -> MakeBF -> DeBruijnGraph
 * neighbour-query stage. The reference has no FFI of its own (one C++ translation unit,
 * reference main.cpp:1-9); every entry point below names the reference function it replaces.
 * Plain pointers and sizes only. All functions return P3_OK (0) or a negative P3_ERR_* code;
 * p3_last_error() gives the message. There is NO CPU fallback: every compute entry point
 * fails with P3_ERR_CUDA when no sm_100 device is usable.
 *
 * k-mer words: W = ceil(2k/64) little-endian uint64 words holding the same 2k-bit integer as
 * the reference's std::bitset<2k> (first base in the most significant 2 bits, A=0 C=1 G=2 T=3;
 * reference src/BitCalc.cpp:8-19). The device path and the drop-in run (p3_assemble_file, the CLI)
 * support 21 <= k <= 3001; the device-side closure p3_dbg_close, p3_node_coverage and the multi-GPU
 * entry points are limited to k <= 32 (W = 1) in this build (for k > 32 the drop-in run closes the
 * table from the host with p3_check_directions batches and counts node coverage on the host).
 *
 * 2-bit staging layout ("packed reads"): all reads back to back, 32 bases per uint64 word,
 * base j of the stream in word j/32 at bits [63-2(j%32)-1, 63-2(j%32)] (MSB first), zero
 * padded, plus read offsets off[0..n_reads] in bases. p3_packed_words() words are required
 * (one halo word included). An optional "non-ACGT" plane (uint32 per word, bit 31-(j%32))
 * reproduces the reference reading such characters as code 0 on both strands
 * (reference src/common.h:32-33 operator[] default-insert).
 */
#ifndef PLATANUS3_B200_H
#define PLATANUS3_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P3_OK 0
#define P3_ERR_CUDA (-1)       /* no device / CUDA runtime error */
#define P3_ERR_ARG (-2)        /* bad argument (unsupported k, null pointer, ...) */
#define P3_ERR_TABLE_FULL (-3) /* a hash table / bin overflowed its capacity (the solid set grows by itself, the count table does not) */
#define P3_ERR_STATE (-4)      /* stage called out of order */
#define P3_ERR_NOMEM (-5)
#define P3_ERR_IO (-6)

#define P3_SHORTK 21        /* reference src/Options.cpp:14 shortk_length */
#define P3_COV_THRESHOLD 2  /* reference src/MakeBloomFilter.cpp:28 cov_threshold */
#define P3_MIN_K 21
#define P3_MAX_K 3001      /* device path: any k in [21,3001] (the reference's largest bitset) */
#define P3_MAX_K_WALK 32   /* single-word k-mers: p3_dbg_close, p3_node_coverage, the uint64 fast path of the host walk */

typedef struct p3_ctx p3_ctx;

const char *p3_last_error(void);
int p3_version(void);
int p3_device_count(void);

/* ---- host-only helpers (no GPU needed) --------------------------------------------------- */

/* Options::EstimateBloomfilter, reference src/Options.cpp:50-60 (the filter_size==0 branch).
 * Fails with P3_ERR_ARG when the estimate degenerates (item_number == 0), where the reference
 * divides by zero. */
int p3_estimate_bloomfilter(uint64_t all_bases, uint32_t k, uint64_t *filter_size, uint32_t *num_hashes);

/* number of uint64 words of 2-bit staging for total_bases bases (incl. one halo word) */
uint64_t p3_packed_words(uint64_t total_bases);

/* GetFirstKmerForward's base coding (reference src/BitCalc.cpp:8-19) applied to whole reads:
 * ASCII reads seq[off[i]..off[i+1]) -> 2-bit staging. nmask may be NULL; *has_non_acgt tells
 * whether any character outside ACGT was seen (then the nmask plane is required for parity). */
int p3_pack_reads(const char *seq, const uint64_t *off, uint64_t n_reads, uint64_t *packed,
                  uint32_t *nmask, int *has_non_acgt);

/* pinned host staging buffers (cudaHostAlloc) */
void *p3_host_alloc(size_t bytes);
void p3_host_free(void *p);

/* ---- context ------------------------------------------------------------------------------ */

/* One context per GPU. stream: a cudaStream_t to launch on, or NULL for a private stream. */
p3_ctx *p3_create(int device, void *stream);
void p3_destroy(p3_ctx *ctx);
int p3_synchronize(p3_ctx *ctx);

/* ---- reads (ReadFile::reads, reference src/Load.cpp:8) ----------------------------------- */

/* host staging -> device (async H2D on the context stream; buffers should be pinned) */
int p3_reads_upload(p3_ctx *ctx, const uint64_t *h_packed, uint64_t total_bases,
                    const uint64_t *h_off, uint64_t n_reads, const uint32_t *h_nmask);
/* device-resident staging owned by the caller (zero copy); must stay alive while attached */
int p3_reads_attach(p3_ctx *ctx, const uint64_t *d_packed, uint64_t total_bases,
                    const uint64_t *d_off, uint64_t n_reads, const uint32_t *d_nmask);

/* ---- stage A: ReadFile::CountShortKmer, reference src/Load.cpp:105-127 -------------------- */

/* Counts every canonical 21-mer of every read into the device count table.
 * table_slots: capacity in 8-byte slots (0 = auto from the number of 21-mer positions). */
int p3_count_short_kmers(p3_ctx *ctx, uint64_t table_slots);
int p3_short_kmer_stats(p3_ctx *ctx, uint64_t *n_positions, uint64_t *n_distinct);
/* shortk_database as (key,count) pairs in table order (unsorted). cap = array capacity. */
int p3_short_kmer_export(p3_ctx *ctx, uint64_t *h_keys, uint64_t *h_counts, uint64_t cap, uint64_t *n);
/* KC[key] for a batch of canonical 21-mers (0 when absent) */
int p3_short_kmer_lookup(p3_ctx *ctx, const uint64_t *h_keys, uint64_t n, uint64_t *h_counts);

/* ---- stage B: MakeBF, reference src/MakeBloomFilter.cpp:25-89 ----------------------------- */

/* 21-mer coverage -> window minimum (RMQ) >= cov_threshold -> BF.add(canonical k-mer) and the
 * first such k-mer of each read as seed. solid_slots: capacity of the distinct-solid-k-mer
 * set (0 = auto, grows on overflow). */
int p3_make_bf(p3_ctx *ctx, uint32_t k, uint64_t filter_size, uint32_t num_hashes,
               uint32_t cov_threshold, uint64_t solid_slots);
int p3_make_bf_stats(p3_ctx *ctx, uint64_t *n_adds, uint64_t *n_distinct_solid);
/* BF::m_bits (reference src/bloomfilter.cpp:28): (filter_size+7)/8 bytes, bit i = byte i>>3 bit i&7 */
int p3_bf_export(p3_ctx *ctx, uint8_t *h_bits);
/* install a filter from host bits (for stand-alone add / query use) */
int p3_bf_import(p3_ctx *ctx, uint32_t k, uint64_t filter_size, uint32_t num_hashes, const uint8_t *h_bits);
/* seed_kmer: position in read r of its first solid k-mer, or -1 (MakeBloomFilter.cpp:79-83) */
int p3_seed_export(p3_ctx *ctx, int64_t *h_seed_pos);
/* one bit per stream position (uint32 per 32 bases, MSB first): k-mer starting there was added */
int p3_solid_flags_export(p3_ctx *ctx, uint32_t *h_bitmap);
/* BF<Key>::add / possiblyContains, reference src/bloomfilter.cpp:69-86, batched over
 * already-canonical k-mers (n*W words) */
int p3_bf_add(p3_ctx *ctx, const uint64_t *h_kmers, uint64_t n);
int p3_bf_possibly_contains(p3_ctx *ctx, const uint64_t *h_kmers, uint64_t n, uint8_t *h_out);
/* GetDoubleHash_64bit, reference src/MyHash.cpp:22-35, over canonical k-mers: out = n*2 words */
int p3_double_hash(p3_ctx *ctx, uint32_t k, const uint64_t *h_kmers, uint64_t n, uint64_t *h_out);

/* ---- stage C: DeBruijnGraph::CheckDirections, reference src/DeBruijnGraph.cpp:326-345 ----- */

/* For every distinct solid k-mer (canonical orientation) the 8 neighbour queries of
 * CheckDirections/IsRecorded (:318-323): bit i of the adjacency byte = direction i recorded
 * (0-3 left extension by A,C,G,T; 4-7 right extension). */
int p3_dbg_adjacency(p3_ctx *ctx);
int p3_dbg_stats(p3_ctx *ctx, uint64_t *n_kmers, uint64_t *n_edges);
/* Closes the k-mer table under recorded neighbours so that a host walk (DeBruijnGraph::MakeDBG,
 * reference src/DeBruijnGraph.cpp:94-297) never needs a k-mer the table lacks: neighbours that
 * answer possiblyContains but are not solid (Bloom false positives) are added with their own
 * adjacency, to a fixed point. h_roots (optional): oriented k-mers the walk starts from (the seeds),
 * which get an entry even when they are not solid (a seed with a non-ACGT base is recorded in
 * forward orientation only). p3_dbg_export then returns solid + root + added k-mers. */
int p3_dbg_close(p3_ctx *ctx, const uint64_t *h_roots, uint64_t n_roots, uint64_t *n_total);
/* distinct solid k-mers (n*W words, unsorted) and their adjacency bytes */
int p3_dbg_export(p3_ctx *ctx, uint64_t *h_kmers, uint8_t *h_adj, uint64_t cap, uint64_t *n);
/* CheckDirections on arbitrary ORIENTED k-mers (ignored_direction = -1) */
int p3_check_directions(p3_ctx *ctx, const uint64_t *h_kmers, uint64_t n, uint8_t *h_mask);

/* ---- whole path ---------------------------------------------------------------------------- */

/* Assemble<>, reference src/Assemble.cpp:7-21, up to the neighbour queries MakeDBG issues:
 * upload + stages A, B, C in one call on host staging buffers. filter_size == 0 runs
 * p3_estimate_bloomfilter(all_bases, k) first, as main.cpp:23 does. Results stay on the device
 * for the export calls above. */
int p3_assemble_hot_path(p3_ctx *ctx, const uint64_t *h_packed, uint64_t total_bases,
                         const uint64_t *h_off, uint64_t n_reads, const uint32_t *h_nmask,
                         uint64_t all_bases, uint32_t k, uint64_t filter_size, uint32_t num_hashes,
                         uint64_t table_slots, uint64_t solid_slots);

/* The same, with the results of MakeBF and CheckDirections written to host buffers (pinned for overlap; any may be
 * NULL): BF::m_bits as p3_bf_export, seed positions as p3_seed_export, the distinct solid k-mers and their adjacency
 * bytes as p3_dbg_export (cap = capacity in k-mers, *n = their number). The staging is uploaded in pieces while the
 * binning kernel already works, and filter / seeds / k-mers leave the device while CheckDirections runs. */
int p3_assemble_hot_path_to_host(p3_ctx *ctx, const uint64_t *h_packed, uint64_t total_bases,
                                 const uint64_t *h_off, uint64_t n_reads, const uint32_t *h_nmask,
                                 uint64_t all_bases, uint32_t k, uint64_t filter_size, uint32_t num_hashes,
                                 uint64_t table_slots, uint64_t solid_slots, uint8_t *h_bits, int64_t *h_seed_pos,
                                 uint64_t *h_kmers, uint8_t *h_adj, uint64_t cap, uint64_t *n);

/* ---- multi-GPU hot path (one process per GPU; csrc/p3_multi.inc.cu, driver: platanus3_b200/dist.py) ------------
 *
 * Canonical k-mers are hash-partitioned over the ranks: the OWNER of a key counts it / de-duplicates it
 * (BASELINE.json north_star; the reference itself is single-process). Every rank owns one peer-visible arena
 * (control block + two receive sets of n_ranks regions each). The binning kernels store a destination rank's
 * records straight into that rank's region over NVLink peer memory (transport 0); transport 1 stages the same
 * regions locally and leaves the movement to the caller's all-to-all (NCCL). All *_send / *_recv / *_finish
 * calls only enqueue work on the context's stream; *_end waits and checks. Order of use:
 *   p3_mg_arena, p3_ipc_export/open, p3_mg_connect                       once (again when set_bytes changes)
 *   A   p3_mg_count_begin, p3_mg_sync, per chunk: _send, sync, _recv;    p3_mg_count_finish, p3_mg_count_end
 *   B1  p3_mg_cover_begin, p3_mg_sync, per slice: _send, sync, _recv
 *   B2  p3_mg_solid_begin, p3_mg_sync, per chunk: _send, sync, _recv;    p3_mg_solid_finish, p3_mg_solid_end
 *   B3  p3_mg_bloom_bin / _apply (sharded filter, then all-gather the shards) or p3_mg_bloom_direct (then OR-reduce)
 *   C   p3_dbg_adjacency                                                  (the filter is complete and local)
 * "sync" = p3_mg_sync for transport 0, the all-to-all of p3_mg_staged_buffers for transport 1.
 * Count record = uint64 [rank:8 @47 | offset-in-word:5 @42 | canonical 21-mer:42] + uint32 word index; position record
 * = uint64 [rank:8 @56 | stream position:56]; k-mer record = uint64 canonical k-mer + uint8 adjacency hint (k <= 32). */
uint32_t p3_owner_of_key(uint64_t key, uint32_t n_ranks);
/* CUDA IPC plumbing: a 64-byte handle of a buffer of this process / a mapping of another process's buffer */
int p3_ipc_export(const void *d_ptr, uint8_t handle[64]);
int p3_ipc_open(int device, const uint8_t handle[64], void **d_ptr);
int p3_ipc_close(int device, void *d_ptr);
/* this rank's arena (set_bytes per receive set, the same on every rank; n_ranks <= 16); arena_ptrs[r] = rank r's arena as
 * mapped into this process. same_stream != 0: all ranks are contexts of one process launching on one stream (tests). */
int p3_mg_arena(p3_ctx *ctx, uint32_t n_ranks, uint32_t my_rank, uint64_t set_bytes, int transport, void **d_arena);
int p3_mg_connect(p3_ctx *ctx, const uint64_t *arena_ptrs, int same_stream);
/* device-side barrier over the peers' control blocks (a one-block kernel on the context's stream; no host wait) */
int p3_mg_sync(p3_ctx *ctx);
/* transport 1: the blocks and counts the caller's all-to-all moves for a stage (0 = A, 1 = B1, 2 = B2, 3 = B2 with multi-word k-mers), see p3_multi.inc.cu */
int p3_mg_staged_buffers(p3_ctx *ctx, int stage, int set, uint64_t out[8]);
/* A: ReadFile::CountShortKmer, reference src/Load.cpp:105-127, over all ranks' reads. table_slots: this rank's table;
 * owner_positions: upper estimate of the 21-mer positions this rank will own; chunk_words / n_chunks: the same on every rank */
int p3_mg_count_begin(p3_ctx *ctx, uint64_t table_slots, uint64_t owner_positions, uint64_t chunk_words, uint64_t n_chunks);
int p3_mg_count_send(p3_ctx *ctx, uint64_t chunk);
int p3_mg_count_recv(p3_ctx *ctx, uint64_t chunk);
int p3_mg_count_finish(p3_ctx *ctx);
int p3_mg_count_end(p3_ctx *ctx);   /* then p3_short_kmer_stats / _export / _lookup answer for the OWNED keys */
/* Human-scale inputs (BASELINE.json configs[3]): an owner cannot keep every received record (16 B each) until one insert
 * sweep. The count then runs in ROUNDS of chunks: p3_mg_count_finish after the last chunk of every round and
 * p3_mg_count_next_round between two rounds (the bins start empty again; owner_positions of p3_mg_count_begin is then the
 * estimate for ONE round). The verdict stage sends and sorts each round's chunks again between p3_mg_cover_rebin_begin /
 * _end (p3_mg_count_send / _recv), followed by that round's slices; the solid stage de-duplicates per round
 * (p3_mg_solid_next_round between two rounds). Results are identical to the one-round schedule. */
int p3_mg_count_next_round(p3_ctx *ctx);
/* The better schedule for the same inputs: rounds over KEY RANGES (groups of table partitions) instead of read chunks.
 * Every round the sources scan all their reads and send only the records of that round's partitions, so an owner sees
 * all occurrences of a key in one round: its counts are final when the round's insert ends, the round's verdicts follow
 * the insert's index stream at once, nothing is sent twice and the table passes through L2 once.
 *   p3_mg_count_begin_keyed (owner_positions = the estimate for ALL rounds), p3_mg_sync; per round r:
 *   p3_mg_key_round_begin(r); every chunk: p3_mg_count_send, sync, p3_mg_count_recv; p3_mg_count_finish;
 *   [r == 0: p3_mg_cover_begin_keyed]; p3_mg_cover_key_round; p3_mg_sync; per slice: p3_mg_cover_send, sync, _recv;
 *   p3_mg_count_next_round before the next round; p3_mg_count_end after the last. */
int p3_mg_count_begin_keyed(p3_ctx *ctx, uint64_t table_slots, uint64_t owner_positions, uint64_t chunk_words, uint64_t n_chunks, uint32_t n_rounds);
int p3_mg_key_round_begin(p3_ctx *ctx, uint32_t round);
int p3_mg_cover_begin_keyed(p3_ctx *ctx, uint32_t cov_threshold, uint64_t owner_distinct, uint32_t *n_slices);
int p3_mg_cover_key_round(p3_ctx *ctx, uint32_t cov_threshold);
int p3_mg_cover_rebin_begin(p3_ctx *ctx);
int p3_mg_cover_rebin_end(p3_ctx *ctx);
int p3_mg_solid_next_round(p3_ctx *ctx);
/* B1: MakeBF's coverage test, reference src/MakeBloomFilter.cpp:52-58: owners return the positions of keys whose count
 * stayed below cov_threshold, in *n_slices rounds (owner_distinct: the largest n_distinct of any rank) */
int p3_mg_cover_begin(p3_ctx *ctx, uint32_t cov_threshold, uint64_t owner_distinct, uint32_t *n_slices);
int p3_mg_cover_send(p3_ctx *ctx, uint32_t cov_threshold, uint32_t slice);
int p3_mg_cover_recv(p3_ctx *ctx, uint32_t slice);
/* B2: MakeBF's RMQ test, seeds and solid k-mers (reference src/MakeBloomFilter.cpp:60-83); owned_slots: capacity of this
 * rank's set of owned distinct solid k-mers. p3_mg_solid_end leaves an empty filter of >= filter_words_cap words. */
int p3_mg_solid_begin(p3_ctx *ctx, uint32_t k, uint64_t owned_slots);
int p3_mg_solid_send(p3_ctx *ctx, uint64_t chunk);
int p3_mg_solid_recv(p3_ctx *ctx, uint64_t chunk);
int p3_mg_solid_finish(p3_ctx *ctx);
int p3_mg_solid_end(p3_ctx *ctx, uint64_t filter_size, uint32_t num_hashes, uint64_t filter_words_cap, uint64_t *n_adds, uint64_t *n_owned);
/* B2 for multi-word k-mers (33 <= k <= 3001; the reference's std::bitset<2k> k-mers, src/Assemble.cpp:30-53): instead of
 * p3_mg_solid_begin .. _finish. p3_mg_long_solid builds the local solid plane and seeds and returns this rank's solid
 * occurrences; p3_mg_long_begin sizes the owner's store of received W-word records (owner_occurrences), its set
 * (owned_slots) and the chunking of the sends (occ_per_word: upper estimate of solid occurrences per packed word on any
 * rank, max_words: the largest n_words of any rank; *chunk_words / *n_chunks are the same on every rank); per chunk
 * p3_mg_long_send, sync (stage 3 for transport 1), p3_mg_long_recv; p3_mg_long_finish de-duplicates the store and leaves
 * the owned distinct k-mers as the context's n x W word list; then p3_mg_solid_end and B3 / C as for k <= 32. */
int p3_mg_long_solid(p3_ctx *ctx, uint32_t k, uint64_t *n_adds);
int p3_mg_long_begin(p3_ctx *ctx, uint64_t owned_slots, uint64_t owner_occurrences, double occ_per_word, uint64_t max_words,
                     uint64_t *chunk_words, uint64_t *n_chunks);
int p3_mg_long_send(p3_ctx *ctx, uint64_t chunk);
int p3_mg_long_recv(p3_ctx *ctx, uint64_t chunk);
int p3_mg_long_finish(p3_ctx *ctx);
/* B3: sharded BF.add (reference src/bloomfilter.cpp:69-74): the filter is cut into segments of p3_bloom_seg_bits() bits,
 * dealt out to the ranks in contiguous shards. p3_mg_bloom_bin computes every owned k-mer's num_hashes bit indices once,
 * sorts them by segment and stores segment s's 4-byte in-segment offsets to h_segbase[s] (device addresses, normally
 * inside the shard owner's p3_mg_bloom_buffer mapped over NVLink peer memory; cap records each); h_counts[s] = records
 * written. The owner ORs them in with p3_mg_bloom_apply (h_ptr/h_n: [n_local][n_src] regions), then the shards are
 * all-gathered. p3_mg_bloom_direct is the fallback (adds into the local full copy; OR-reduce afterwards). */
uint64_t p3_bloom_seg_bits(void);
int p3_mg_bloom_buffer(p3_ctx *ctx, uint64_t n_u32, uint32_t **d_buf);
int p3_mg_bloom_bin(p3_ctx *ctx, uint32_t n_seg, const uint64_t *h_segbase, uint64_t cap, uint64_t *h_counts);
/* the same for k-mers [first, first + count) of the owned list: one PASS at a time through smaller shard buffers */
int p3_mg_bloom_bin_range(p3_ctx *ctx, uint32_t n_seg, const uint64_t *h_segbase, uint64_t cap, uint64_t *h_counts, uint64_t first, uint64_t count);
int p3_mg_bloom_apply(p3_ctx *ctx, uint64_t seg_first, uint32_t n_local, uint32_t n_src, const uint64_t *h_ptr, const uint64_t *h_n);
int p3_mg_bloom_direct(p3_ctx *ctx);
int p3_mg_filter(p3_ctx *ctx, uint32_t **d_bits, uint64_t *n_words);
int p3_mg_makebf_done(p3_ctx *ctx);
/* device memory in use on the context's GPU (total - free), for the benchmark's hbm_peak_bytes */
uint64_t p3_device_mem_used(p3_ctx *ctx);

/* ---- host side of the drop-in (row f: callers / formats either side of the path) ------------- */

/* ReadFile::LoadFile / LoadFasta / LoadFastq, reference src/Load.cpp:32-103: FASTA or single-line
 * FASTQ chosen by the first byte, reads shorter than k dropped, records with the same name line
 * collapse to the last one while all_bases counts every record. The reads come back as ASCII,
 * offsets and (pinned, when a GPU is present) 2-bit staging ready for p3_reads_upload. */
typedef struct p3_reads p3_reads;
int p3_load_file(const char *path, uint32_t k, p3_reads **out);
void p3_reads_free(p3_reads *r);
uint64_t p3_reads_count(const p3_reads *r);
uint64_t p3_reads_all_bases(const p3_reads *r);
uint64_t p3_reads_total_bases(const p3_reads *r);
const uint64_t *p3_reads_offsets(const p3_reads *r);
const uint64_t *p3_reads_packed(const p3_reads *r);
const uint32_t *p3_reads_nmask(const p3_reads *r);   /* NULL when every base is ACGT */
const char *p3_reads_ascii(const p3_reads *r);

/* DeBruijnGraph::CountNodeCoverage, reference src/DeBruijnGraph.cpp:394-449, over the reads attached
 * to ctx: h_junctions / h_joints are the ORIENTED node k-mers (k <= 32) as the walk recorded them;
 * h_jcov receives 9 counters per junction [coverage, left_kmers_cov[4], right_kmers_cov[4]],
 * h_tcov one coverage counter per joint. */
int p3_node_coverage(p3_ctx *ctx, uint32_t k, const uint64_t *h_junctions, uint64_t nj,
                     const uint64_t *h_joints, uint64_t nt, int32_t *h_jcov, int32_t *h_tcov);

/* main() + Assemble<>, reference main.cpp:11-31 and src/Assemble.cpp:7-28, for one read file:
 * Load, filter sizing, the GPU hot path, then on the host the unitig walk (MakeDBG with the
 * reference's -t 1 order), CountNodeCoverage and PrintGraph. Writes the GFA to gfa_path and the
 * reference's milestone log lines to log_path (either may be NULL). stats (optional, 8 values):
 * reads, all_bases, distinct 21-mers, solid k-mers, table k-mers after closure, junctions, joints,
 * straights. */
int p3_assemble_file(const char *read_path, uint32_t k, uint64_t m, int threads, int device,
                     const char *gfa_path, const char *log_path, uint64_t *stats);

/* The host half of that run on its own: Load + MakeDBG (the reference's -t 1 order, src/DeBruijnGraph.cpp:94-297)
 * + CountNodeCoverage (:394-449, counted on the host here) + PrintGraph (:452-544) over a CLOSED
 * CheckDirections table from any source: n canonical k-mers (W words each) with one adjacency byte
 * each, closed = every neighbour a byte reports is itself in the table; h_seeds = the ORIENTED seed
 * k-mers (MakeBloomFilter.cpp:79-83). Needs no GPU. stats (optional, 3 values): junctions, joints,
 * straights. */
int p3_walk_table(const char *read_path, uint32_t k, const uint64_t *h_kmers, const uint8_t *h_adj, uint64_t n,
                  const uint64_t *h_seeds, uint64_t n_seeds, const char *gfa_path, uint64_t *stats);

/* device milliseconds of the last run of each stage (CUDA events on the context stream):
 * ms[0]=count21 ms[1]=coverage flags ms[2]=solid+bloom ms[3]=seeds ms[4]=adjacency */
int p3_stage_ms(p3_ctx *ctx, float ms[5]);
/* finer timing: ms[0]=histogram+scan ms[1]=scatter into partition bins ms[2]=insert sweep of the last
 * p3_count_short_kmers call (0 in direct mode), ms[3]=dense BF.add passes of the last p3_make_bf
 * (part of p3_stage_ms[2]); *parts = table partitions, *chunks = read chunks (0 = direct mode) */
int p3_count_substage_ms(p3_ctx *ctx, float ms[4], uint32_t *parts, uint64_t *chunks);
/* table load and probe statistics for the k x threshold sweep (BASELINE.json configs[4]). p3_probe_stats: out[0], out[1] =
 * mean / longest number of 32-byte buckets an insert of the last p3_count_short_kmers touched (when it ran with the
 * environment variable P3_PROBE_STATS set, else 0); out[2], out[3] = the same for a lookup of every distinct solid k-mer
 * in the solid set. p3_table_capacity: count table slots and partitions, solid set slots and partitions. */
int p3_probe_stats(p3_ctx *ctx, double out[4]);
/* number of partitions the binned count table of this capacity is cut into (what the multi-GPU driver plans key-range rounds with) */
uint32_t p3_table_partitions(uint64_t table_slots);
int p3_table_capacity(p3_ctx *ctx, uint64_t out[4]);
/* kernels launched by this context since creation (for bench.py's gpu_launches) */
uint64_t p3_launch_count(p3_ctx *ctx);
/* filter parameters in effect */
int p3_bf_params(p3_ctx *ctx, uint64_t *filter_size, uint32_t *num_hashes, uint32_t *k);

#ifdef __cplusplus
}
#endif
#endif
