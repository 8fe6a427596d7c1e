// examples/reference_binding.cpp — the binding a platanus3 maintainer would write, as a program that
// compiles against include/platanus3_b200.h alone (tests/test_abi_cpu.py builds it on every CPU run
// so that INTEGRATION.md cannot drift from the header). It is Assemble<> (reference
// src/Assemble.cpp:7-28) with the hot path on the GPU and the rest through the C ABI:
//
//   ReadFile::LoadFile                     -> p3_load_file            (host)
//   Options::EstimateBloomfilter           -> p3_estimate_bloomfilter (host)
//   CountShortKmer + MakeBF + every CheckDirections answer
//                                          -> p3_assemble_hot_path    (GPU)
//   the Bloom false positives the walk can step on
//                                          -> p3_dbg_close            (GPU, k <= 32)
//   MakeDBG(-t 1) + CountNodeCoverage + PrintGraph
//                                          -> p3_walk_table           (host, any closed table)
//
// usage: reference_binding reads.fastq k [out.gfa]
#include "platanus3_b200.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

static int die(const char *what, int rc) {
    fprintf(stderr, "%s: error %d: %s\n", what, rc, p3_last_error());
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s reads.fa|fq k [out.gfa]\n", argv[0]); return 2; }
    const char *path = argv[1];
    const uint32_t k = (uint32_t)atoi(argv[2]);
    const char *gfa = argc > 3 ? argv[3] : "de_bruijn_graph.gfa";
    if (k > P3_MAX_K_WALK) { fprintf(stderr, "this example closes the table on the device: k <= %d (p3_assemble_file does any k)\n", P3_MAX_K_WALK); return 2; }

    // Load: reads >= k, all_bases, pinned 2-bit staging
    p3_reads *rd = nullptr;
    int rc = p3_load_file(path, k, &rd);
    if (rc) return die("p3_load_file", rc);
    const uint64_t n_reads = p3_reads_count(rd), total = p3_reads_total_bases(rd);

    // main.cpp:22-23: size the filter from all_bases (or take -m)
    uint64_t filter_size = 0; uint32_t num_hashes = 0;
    rc = p3_estimate_bloomfilter(p3_reads_all_bases(rd), k, &filter_size, &num_hashes);
    if (rc) return die("p3_estimate_bloomfilter", rc);

    // the hot path: upload + CountShortKmer + MakeBF + CheckDirections of every solid k-mer
    p3_ctx *ctx = p3_create(/*device*/ 0, /*stream*/ nullptr);
    if (!ctx) return die("p3_create", P3_ERR_CUDA);              // there is no CPU fallback
    rc = p3_assemble_hot_path(ctx, p3_reads_packed(rd), total, p3_reads_offsets(rd), n_reads, p3_reads_nmask(rd),
                              p3_reads_all_bases(rd), k, filter_size, num_hashes, /*table_slots*/ 0, /*solid_slots*/ 0);
    if (rc) return die("p3_assemble_hot_path", rc);

    // seeds: first solid k-mer of every read (MakeBloomFilter.cpp:79-83), as oriented k-mer words
    std::vector<int64_t> seed_pos(n_reads);
    rc = p3_seed_export(ctx, seed_pos.data());
    if (rc) return die("p3_seed_export", rc);
    const char *ascii = p3_reads_ascii(rd);
    const uint64_t *off = p3_reads_offsets(rd);
    const uint64_t kmask = k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
    std::vector<uint64_t> seeds;
    for (uint64_t r = 0; r < n_reads; r++) {
        if (seed_pos[r] < 0) continue;
        uint64_t v = 0;
        for (uint32_t i = 0; i < k; i++) {
            const char c = ascii[off[r] + (uint64_t)seed_pos[r] + i];
            v = ((v << 2) | (uint64_t)(c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0)) & kmask;   // anything else reads as A
        }
        seeds.push_back(v);
    }

    // every k-mer the walk can reach: solid k-mers + the Bloom false positives next to them + the seeds
    uint64_t n_table = 0;
    rc = p3_dbg_close(ctx, seeds.data(), seeds.size(), &n_table);
    if (rc) return die("p3_dbg_close", rc);
    std::vector<uint64_t> kmers(n_table ? n_table : 1);
    std::vector<uint8_t> adj(n_table ? n_table : 1);
    uint64_t got = 0;
    rc = p3_dbg_export(ctx, kmers.data(), adj.data(), kmers.size(), &got);
    if (rc) return die("p3_dbg_export", rc);
    p3_destroy(ctx);
    p3_reads_free(rd);

    // MakeDBG (-t 1 order), CountNodeCoverage, PrintGraph on the host
    uint64_t nodes[3] = {0, 0, 0};
    rc = p3_walk_table(path, k, kmers.data(), adj.data(), got, seeds.data(), seeds.size(), gfa, nodes);
    if (rc) return die("p3_walk_table", rc);
    printf("%llu reads, %llu table k-mers, %llu junctions, %llu joints, %llu straights -> %s\n",
           (unsigned long long)n_reads, (unsigned long long)got, (unsigned long long)nodes[0],
           (unsigned long long)nodes[1], (unsigned long long)nodes[2], gfa);
    return 0;
}
