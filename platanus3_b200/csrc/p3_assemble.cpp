// p3_assemble.cpp — host side of the drop-in: Load (FASTA/FASTQ -> 2-bit staging), the unitig walk
// over the GPU-built adjacency table, node coverage, GFA output and the whole-run entry point
// p3_assemble_file that the platanus3-compatible CLI (p3_cli.cpp) calls.
//
// What runs where: CountShortKmer, MakeBF and every CheckDirections/IsRecorded answer come from
// the CUDA library (p3_gpu.cu). The walk itself (reference src/DeBruijnGraph.cpp:94-297) is
// sequential pointer chasing with order-dependent node ids, so it stays on the host, exactly as
// BASELINE.json's north_star says; it only reads the adjacency table exported by the GPU.
// Semantics follow the reference run with -t 1 (its multi-threaded walk is racy).
#include "../../include/platanus3_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

extern "C" void p3_internal_set_error(const char *msg);   // p3_gpu.cu: feeds p3_last_error()

namespace {

struct HostErr {   // assignment forwards to the library's thread-local error string
    HostErr &operator=(const std::string &m) { p3_internal_set_error(m.c_str()); return *this; }
    HostErr &operator=(const char *m) { p3_internal_set_error(m); return *this; }
} g_host_err;

// ---------------------------------------------------------------------------------------------
// Load: reference src/Load.cpp:32-103
// ---------------------------------------------------------------------------------------------
}  // namespace

struct p3_reads {
    std::string seq;                 // concatenated ASCII of the kept reads
    std::vector<uint64_t> off;       // offsets in bases, size n+1
    uint64_t all_bases = 0;          // counts every record >= k, duplicates included (Load.cpp:61)
    uint64_t *packed = nullptr;      // pinned when a GPU is present
    uint32_t *nmask = nullptr;
    bool pinned = false;
    int has_non_acgt = 0;
};

namespace {

void add_record(std::unordered_map<std::string, size_t> &index, std::vector<std::string> &names,
                std::vector<std::string> &seqs, const std::string &name, std::string &seq, uint32_t k,
                uint64_t &all_bases) {
    if (seq.size() >= k) {
        auto it = index.find(name);
        if (it == index.end()) {
            index.emplace(name, seqs.size());
            names.push_back(name);
            seqs.push_back(seq);
        } else {
            seqs[it->second] = seq;   // same name line: the later record replaces the earlier
        }
        all_bases += seq.size();
    }
}

}  // namespace

extern "C" {

int p3_load_file(const char *path, uint32_t k, p3_reads **out) {
    if (!path || !out) return P3_ERR_ARG;
    // reference src/Load.cpp:26: file_name.substr(size-5, 5) throws for names shorter than 5
    if (strlen(path) < 5) { g_host_err = "read file name shorter than 5 characters"; return P3_ERR_ARG; }
    std::ifstream in(path);
    if (!in.is_open()) { g_host_err = std::string("cannot open ") + path; return P3_ERR_IO; }
    std::unordered_map<std::string, size_t> index;
    std::vector<std::string> names, seqs;
    uint64_t all_bases = 0;
    std::string line, seq, name;
    bool first = true;
    int mode = 0;   // 1 FASTA, 2 FASTQ, decided by the first byte of the first line (Load.cpp:40-48)
    uint64_t line_cnt = 0;
    while (std::getline(in, line)) {
        if (first) {
            first = false;
            if (!line.empty() && line[0] == '>') mode = 1;
            else if (!line.empty() && line[0] == '@') mode = 2;
            else break;   // neither: nothing is loaded
        }
        bool header = mode == 1 ? (!line.empty() && line[0] == '>') : (line_cnt % 4 == 0);
        if (header) {
            if (!name.empty()) {
                add_record(index, names, seqs, name, seq, k, all_bases);
                seq.clear();
            }
            name = line;
        } else if (mode == 1 || line_cnt % 4 == 1) {
            seq += line;
        }
        line_cnt++;
    }
    if (mode && seq.size() >= k) add_record(index, names, seqs, name, seq, k, all_bases);   // Load.cpp:71-74
    p3_reads *r = new p3_reads();
    r->all_bases = all_bases;
    r->off.push_back(0);
    size_t total = 0;
    for (auto &s : seqs) total += s.size();
    r->seq.reserve(total);
    for (auto &s : seqs) { r->seq += s; r->off.push_back(r->seq.size()); }
    uint64_t words = p3_packed_words(r->seq.size());
    r->packed = (uint64_t *)p3_host_alloc(words * sizeof(uint64_t));
    r->nmask = (uint32_t *)p3_host_alloc(words * sizeof(uint32_t));
    r->pinned = r->packed && r->nmask;
    if (!r->pinned) {   // no CUDA device: plain host memory is still fine for the packer
        if (r->packed) p3_host_free(r->packed);
        if (r->nmask) p3_host_free(r->nmask);
        r->packed = (uint64_t *)malloc(words * sizeof(uint64_t));
        r->nmask = (uint32_t *)malloc(words * sizeof(uint32_t));
    }
    p3_pack_reads(r->seq.data(), r->off.data(), r->off.size() - 1, r->packed, r->nmask, &r->has_non_acgt);
    *out = r;
    return P3_OK;
}
void p3_reads_free(p3_reads *r) {
    if (!r) return;
    if (r->pinned) { p3_host_free(r->packed); p3_host_free(r->nmask); }
    else { free(r->packed); free(r->nmask); }
    delete r;
}
uint64_t p3_reads_count(const p3_reads *r) { return r->off.size() - 1; }
uint64_t p3_reads_all_bases(const p3_reads *r) { return r->all_bases; }
uint64_t p3_reads_total_bases(const p3_reads *r) { return r->seq.size(); }
const uint64_t *p3_reads_offsets(const p3_reads *r) { return r->off.data(); }
const uint64_t *p3_reads_packed(const p3_reads *r) { return r->packed; }
const uint32_t *p3_reads_nmask(const p3_reads *r) { return r->has_non_acgt ? r->nmask : nullptr; }
const char *p3_reads_ascii(const p3_reads *r) { return r->seq.data(); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// unitig walk over the exported adjacency table (k <= 32: one word per k-mer)
// ---------------------------------------------------------------------------------------------
namespace {

inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
inline uint8_t rev8(uint8_t b) {
    b = (uint8_t)((b & 0xF0) >> 4 | (b & 0x0F) << 4);
    b = (uint8_t)((b & 0xCC) >> 2 | (b & 0x33) << 2);
    b = (uint8_t)((b & 0xAA) >> 1 | (b & 0x55) << 1);
    return b;
}

// canonical k-mer -> adjacency byte, open addressing (built once from p3_dbg_export)
struct AdjTable {
    std::vector<uint64_t> keys;
    std::vector<uint8_t> vals;
    uint64_t mask = 0;
    void build(const std::vector<uint64_t> &k, const std::vector<uint8_t> &a) {
        uint64_t cap = 16;
        while (cap < 2 * k.size() + 16) cap <<= 1;
        keys.assign(cap, ~0ULL); vals.assign(cap, 0); mask = cap - 1;
        for (size_t i = 0; i < k.size(); i++) {
            uint64_t s = fmix64(k[i]) & mask;
            while (keys[s] != ~0ULL) s = (s + 1) & mask;
            keys[s] = k[i]; vals[s] = a[i];
        }
    }
    bool find(uint64_t key, uint8_t *v) const {
        uint64_t s = fmix64(key) & mask;
        while (keys[s] != ~0ULL) {
            if (keys[s] == key) { *v = vals[s]; return true; }
            s = (s + 1) & mask;
        }
        return false;
    }
};

struct Junction { int id = 0; int coverage = 0; int left_cov[4] = {0, 0, 0, 0}; int right_cov[4] = {0, 0, 0, 0}; };
struct Joint { int id = 0; int coverage = 0; int straight = 0; };
struct Straight { int id = 0; std::string sequence; };

struct Walker {
    int k;
    uint64_t kmask;
    const AdjTable &adj;
    std::unordered_map<uint64_t, Junction> junctions;
    std::unordered_map<uint64_t, Joint> joints;
    std::vector<Straight> straights;   // id = index + 1
    std::deque<uint64_t> visiting;
    int junction_id = 0, joint_id = 0;
    uint64_t missing = 0;              // CheckDirections on a k-mer the table lacks (must stay 0)

    Walker(int k_, const AdjTable &a) : k(k_), kmask(k_ >= 32 ? ~0ULL : ((1ULL << (2 * k_)) - 1)), adj(a) {}

    uint64_t revcomp(uint64_t v) const {   // GetComplementKmer, reference src/BitCalc.cpp:36-45
        uint64_t x = ~v;
        x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
        x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
        x = __builtin_bswap64(x);
        return x >> (64 - 2 * k);
    }
    std::string str(uint64_t v) const {   // GetStringKmer, reference src/BitCalc.cpp:57-65
        std::string s(k, 'A');
        for (int i = 0; i < k; i++) s[i] = "ACGT"[(v >> (2 * (k - 1 - i))) & 3];
        return s;
    }
    // 8 CheckDirections bits of an ORIENTED k-mer: the table stores the canonical orientation;
    // for the other strand direction i of K is direction 7-i of revcomp(K)
    uint8_t directions(uint64_t km) {
        uint64_t rc = revcomp(km);
        uint8_t a = 0;
        if (km <= rc) { if (!adj.find(km, &a)) missing++; return a; }
        if (!adj.find(rc, &a)) missing++;
        return rev8(a);
    }
    uint64_t neighbour(uint64_t km, int d) const {   // reference src/DeBruijnGraph.cpp:327-339
        return d < 4 ? ((km >> 2) | ((uint64_t)d << (2 * k - 2))) : (((km << 2) | (uint64_t)(d - 4)) & kmask);
    }
    // CheckDirections, reference src/DeBruijnGraph.cpp:326-345
    void check_directions(std::vector<uint64_t> &left, std::vector<uint64_t> &right, uint64_t km, int ignored) {
        left.clear(); right.clear();
        uint8_t a = directions(km);
        for (int i = 0; i < 8; i++) {
            if (i == ignored || !((a >> i) & 1)) continue;
            (i < 4 ? left : right).push_back(neighbour(km, i));
        }
    }
    bool is_visited(uint64_t km) {   // reference src/DeBruijnGraph.cpp:300-315
        uint64_t rc = revcomp(km);
        return junctions.count(km) || junctions.count(rc) || joints.count(km) || joints.count(rc);
    }
    void add_junction(uint64_t km) {   // :348-357
        if (is_visited(km)) return;
        junctions[km].id = ++junction_id;
    }
    void add_joint(uint64_t km) {      // :360-369
        if (is_visited(km)) return;
        joints[km].id = ++joint_id;
    }
    void add_straight(const std::string &seq, uint64_t lj, uint64_t rj) {   // :374-391
        if (is_visited(lj)) return;
        add_joint(lj);
        add_joint(rj);
        Straight s; s.id = (int)straights.size() + 1; s.sequence = seq;
        straights.push_back(s);
        joints[lj].straight = s.id;   // operator[] semantics: the entry exists afterwards even if
        joints[rj].straight = s.id;   // add_joint refused it (reference :385-389)
    }
    void push_all(const std::vector<uint64_t> &l, const std::vector<uint64_t> &r) {
        for (uint64_t v : l) visiting.push_back(v);
        for (uint64_t v : r) visiting.push_back(v);
    }
    uint64_t extend_left(uint64_t target, uint64_t previous, std::vector<char> &ext, int previous_base) {   // :229-260
        std::vector<uint64_t> l, r;
        check_directions(l, r, target, 4 + previous_base);
        while (l.size() == 1 && r.empty()) {
            if (is_visited(target)) { ext.clear(); return target; }
            ext.push_back("ACGT"[(target >> (2 * k - 2)) & 3]);
            previous_base = (int)(target & 3);
            previous = target;
            target = l[0];
            check_directions(l, r, target, 4 + previous_base);
        }
        if (!is_visited(target)) { push_all(l, r); add_junction(target); }
        return previous;
    }
    uint64_t extend_right(uint64_t target, uint64_t previous, std::vector<char> &ext, int previous_base) {  // :264-297
        std::vector<uint64_t> l, r;
        check_directions(l, r, target, previous_base);
        while (l.empty() && r.size() == 1) {
            if (is_visited(target)) { ext.clear(); return target; }
            ext.push_back("ACGT"[target & 3]);
            previous_base = (int)((target >> (2 * k - 2)) & 3);
            previous = target;
            target = r[0];
            check_directions(l, r, target, previous_base);
        }
        if (!is_visited(target)) { push_all(l, r); add_junction(target); }
        return previous;
    }
    void search_node(uint64_t target) {   // :158-225
        if (is_visited(target)) return;
        std::vector<uint64_t> l, r;
        check_directions(l, r, target, -1);
        if (l.size() != 1 || r.size() != 1) { push_all(l, r); add_junction(target); return; }
        std::vector<char> ext_l, ext_r;
        uint64_t left_end = extend_left(l[0], target, ext_l, (int)(target & 3));
        std::string left_part(ext_l.rbegin(), ext_l.rend());
        if (is_visited(left_end)) return;
        uint64_t right_end = extend_right(r[0], target, ext_r, (int)((target >> (2 * k - 2)) & 3));
        std::string right_part(ext_r.begin(), ext_r.end());
        if (is_visited(right_end)) return;
        if (left_end == right_end) { add_junction(left_end); return; }
        if (left_part.size() + right_part.size() >= 1) add_straight(left_part + str(target) + right_part, left_end, right_end);
    }
    // MakeDBG with threads_num = 1, reference src/DeBruijnGraph.cpp:94-155. seeds sorted ascending
    // (== std::set<std::string> order for ACGT strings of equal length).
    int make_dbg(const std::vector<uint64_t> &seeds, uint64_t node_limit) {
        uint64_t cnt = 0;
        for (uint64_t s : seeds) {
            cnt++;
            if (is_visited(s)) continue;
            visiting.push_back(s);
            if (cnt % 20 != 0) continue;
            while (!visiting.empty()) {
                uint64_t v = visiting.front(); visiting.pop_front();
                search_node(v);
                if (junctions.size() > node_limit) return -1;
            }
        }
        while (!visiting.empty()) {
            uint64_t v = visiting.front(); visiting.pop_front();
            search_node(v);
            if (junctions.size() > node_limit) return -1;
        }
        return 0;
    }
};

inline int fcode(unsigned char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }

// CountNodeCoverage (reference src/DeBruijnGraph.cpp:394-439) runs on the GPU (p3_node_coverage):
// the node k-mers go down in id order, the counters come back into the maps.
int count_node_coverage(Walker &w, p3_ctx *ctx) {
    std::vector<uint64_t> jk(w.junctions.size()), tk;
    std::vector<uint64_t> tkeys;
    for (auto &kv : w.junctions) jk[kv.second.id - 1] = kv.first;   // ids are 1..n without gaps
    tkeys.reserve(w.joints.size());
    for (auto &kv : w.joints) tkeys.push_back(kv.first);
    std::vector<int32_t> jc(9 * jk.size() + 1), tc(tkeys.size() + 1);
    int rc = p3_node_coverage(ctx, (uint32_t)w.k, jk.data(), jk.size(), tkeys.data(), tkeys.size(), jc.data(), tc.data());
    if (rc) return rc;
    for (size_t i = 0; i < jk.size(); i++) {
        Junction &J = w.junctions[jk[i]];
        J.coverage = jc[9 * i];
        for (int b = 0; b < 4; b++) { J.left_cov[b] = jc[9 * i + 1 + b]; J.right_cov[b] = jc[9 * i + 5 + b]; }
    }
    for (size_t i = 0; i < tkeys.size(); i++) w.joints[tkeys[i]].coverage = tc[i];
    return P3_OK;
}

// PrintGraph, reference src/DeBruijnGraph.cpp:452-544. The reference iterates unordered_maps, so its
// line order is unspecified; this writes straights and junctions by id.
int print_graph(Walker &w, const char *path) {
    FILE *f = fopen(path, "w");
    if (!f) return P3_ERR_IO;
    const int k = w.k;
    fprintf(f, "H\tVN:Z:1.0\n");
    for (auto &s : w.straights) fprintf(f, "S\tStraight_%d\t%s\tKC:i:%zu\n", s.id, s.sequence.c_str(), s.sequence.size());
    std::vector<std::pair<int, uint64_t>> js;
    for (auto &kv : w.junctions) js.push_back({kv.second.id, kv.first});
    std::sort(js.begin(), js.end());
    for (auto &p : js) fprintf(f, "S\tJunction_%d\t%s\tKC:i:%d\n", p.first, w.str(p.second).c_str(), w.junctions[p.second].coverage * k);
    for (auto &p : js) {
        const uint64_t km = p.second;
        const Junction &J = w.junctions[km];
        const uint8_t a = w.directions(km);
        for (int i = 0; i < 4; i++) {
            if (J.left_cov[i] == 0 || !((a >> i) & 1)) continue;   // IsRecorded(output_left_kmer)
            uint64_t n = w.neighbour(km, i), nb = w.revcomp(n);
            auto j1 = w.junctions.find(n);
            if (j1 != w.junctions.end()) { fprintf(f, "L\tJunction_%d\t+\tJunction_%d\t+\t%dM\n", j1->second.id, J.id, k - 1); continue; }
            auto t1 = w.joints.find(n);
            if (t1 != w.joints.end()) { fprintf(f, "L\tStraight_%d\t+\tJunction_%d\t+\t%dM\n", t1->second.straight, J.id, k - 1); continue; }
            auto j2 = w.junctions.find(nb);
            if (j2 != w.junctions.end()) { fprintf(f, "L\tJunction_%d\t-\tJunction_%d\t+\t%dM\n", j2->second.id, J.id, k - 1); continue; }
            auto t2 = w.joints.find(nb);
            if (t2 != w.joints.end()) fprintf(f, "L\tStraight_%d\t-\tJunction_%d\t+\t%dM\n", t2->second.straight, J.id, k - 1);
        }
        for (int i = 0; i < 4; i++) {
            if (J.right_cov[i] == 0 || !((a >> (4 + i)) & 1)) continue;
            uint64_t n = w.neighbour(km, 4 + i), nb = w.revcomp(n);
            auto j1 = w.junctions.find(n);
            if (j1 != w.junctions.end()) { fprintf(f, "L\tJunction_%d\t+\tJunction_%d\t+\t%dM\n", J.id, j1->second.id, k - 1); continue; }
            auto t1 = w.joints.find(n);
            if (t1 != w.joints.end()) { fprintf(f, "L\tJunction_%d\t+\tStraight_%d\t+\t%dM\n", J.id, t1->second.straight, k - 1); continue; }
            auto j2 = w.junctions.find(nb);
            if (j2 != w.junctions.end()) { fprintf(f, "L\tJunction_%d\t+\tJunction_%d\t-\t%dM\n", J.id, j2->second.id, k - 1); continue; }
            auto t2 = w.joints.find(nb);
            if (t2 != w.joints.end()) fprintf(f, "L\tJunction_%d\t+\tStraight_%d\t-\t%dM\n", J.id, t2->second.straight, k - 1);
        }
    }
    fclose(f);
    return P3_OK;
}

struct Log {   // reference src/Logging.cpp: append one line per call
    std::string path;
    void line(const std::string &t) {
        if (path.empty()) return;
        std::ofstream o(path, std::ios::out | std::ios::app);
        if (o.is_open()) o << t << "\n";
    }
};

}  // namespace

extern "C" {

// main.cpp:11-31 + Assemble<>, reference src/Assemble.cpp:7-28, for one read file.
// m = 0 estimates the filter (Options.cpp:50); threads is accepted for command-line compatibility
// (the walk follows the reference's -t 1 order). stats (optional, 8 values): reads, all_bases,
// distinct 21-mers, solid k-mers, table k-mers after closure, junctions, joints, straights.
int p3_assemble_file(const char *read_path, uint32_t k, uint64_t m, int threads, int device,
                     const char *gfa_path, const char *log_path, uint64_t *stats) {
    (void)threads;
    Log log; log.path = log_path ? log_path : "";
    if (k < P3_MIN_K || k > P3_MAX_K_WALK) { g_host_err = "k outside [21,32] is not supported by the host walk of this build"; return P3_ERR_ARG; }
    p3_reads *rd = nullptr;
    int rc = p3_load_file(read_path, k, &rd);
    if (rc) return rc;
    uint64_t n_reads = p3_reads_count(rd);
    uint64_t filter_size = m; uint32_t num_hashes = 10;   // Options.cpp:10-11
    if (filter_size == 0) {
        rc = p3_estimate_bloomfilter(rd->all_bases, k, &filter_size, &num_hashes);
        if (rc) { g_host_err = "cannot size the Bloom filter (too few bases; give -m)"; p3_reads_free(rd); return rc; }
        uint64_t item = (uint64_t)((double)rd->all_bases * 0.0005 * (double)k);
        log.line("all_bases : " + std::to_string(rd->all_bases));
        log.line("item_number :" + std::to_string(item));
        log.line("get filter_size :" + std::to_string(filter_size));
    }
    log.line("read file loaded");
    log.line(std::string("readfile_name : ") + read_path);
    log.line("filter_size : " + std::to_string(filter_size));
    log.line("num_hashes : " + std::to_string((int)num_hashes));
    log.line("kmer_length : " + std::to_string(k));
    log.line("error_rate : " + std::to_string(0.0005));
    log.line("Assemble");
    if (n_reads == 0) { g_host_err = "no reads of length >= k"; p3_reads_free(rd); return P3_ERR_ARG; }

    p3_ctx *ctx = p3_create(device, nullptr);
    if (!ctx) { g_host_err = p3_last_error(); p3_reads_free(rd); return P3_ERR_CUDA; }
    auto bail = [&](int code) { g_host_err = p3_last_error(); p3_destroy(ctx); p3_reads_free(rd); return code; };
    rc = p3_reads_upload(ctx, rd->packed, rd->seq.size(), rd->off.data(), n_reads, p3_reads_nmask(rd));
    if (rc) return bail(rc);
    rc = p3_count_short_kmers(ctx, 0);
    if (rc) return bail(rc);
    log.line("counted short kmer");
    rc = p3_make_bf(ctx, k, filter_size, num_hashes, P3_COV_THRESHOLD, 0);
    if (rc) return bail(rc);
    log.line("bloom filter loaded");
    log.line("get seed kmer");
    std::vector<int64_t> seed_pos(n_reads);
    rc = p3_seed_export(ctx, seed_pos.data());
    if (rc) return bail(rc);
    // seed_kmer: std::set of the forward strings (MakeBloomFilter.cpp:80-81) -> sorted unique values
    std::vector<uint64_t> seeds;
    const uint64_t kmask = k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
    for (uint64_t r = 0; r < n_reads; r++) {
        if (seed_pos[r] < 0) continue;
        const unsigned char *s = (const unsigned char *)rd->seq.data() + rd->off[r] + seed_pos[r];
        uint64_t v = 0;
        for (uint32_t i = 0; i < k; i++) v = ((v << 2) | (uint64_t)fcode(s[i])) & kmask;
        seeds.push_back(v);
    }
    std::sort(seeds.begin(), seeds.end());
    seeds.erase(std::unique(seeds.begin(), seeds.end()), seeds.end());
    log.line("seed kmer num= " + std::to_string(seeds.size()));

    rc = p3_dbg_adjacency(ctx);
    if (rc) return bail(rc);
    uint64_t n_solid = 0, n_edges = 0, n_table = 0, n_pos = 0, n_d21 = 0;
    p3_dbg_stats(ctx, &n_solid, &n_edges);
    p3_short_kmer_stats(ctx, &n_pos, &n_d21);
    rc = p3_dbg_close(ctx, seeds.data(), seeds.size(), &n_table);
    if (rc) return bail(rc);
    std::vector<uint64_t> kmers(n_table ? n_table : 1);
    std::vector<uint8_t> adjb(n_table ? n_table : 1);
    uint64_t got = 0;
    rc = p3_dbg_export(ctx, kmers.data(), adjb.data(), kmers.size(), &got);
    if (rc) return bail(rc);
    kmers.resize(got); adjb.resize(got);

    log.line("start graph extention");
    AdjTable table; table.build(kmers, adjb);
    std::vector<uint64_t>().swap(kmers);
    Walker w((int)k, table);
    if (w.make_dbg(seeds, /*node_limit*/ 4 * (got + 16)) != 0 || w.missing) {
        g_host_err = w.missing ? "internal: walk left the closed adjacency table" : "walk did not terminate";
        p3_destroy(ctx);
        p3_reads_free(rd);
        return P3_ERR_STATE;
    }
    log.line("de bruijn graph loaded");
    rc = count_node_coverage(w, ctx);
    if (rc) return bail(rc);
    p3_destroy(ctx);
    log.line("count node coverage");
    rc = gfa_path ? print_graph(w, gfa_path) : P3_OK;
    if (stats) {
        stats[0] = n_reads; stats[1] = rd->all_bases; stats[2] = n_d21; stats[3] = n_solid; stats[4] = got;
        stats[5] = w.junctions.size(); stats[6] = w.joints.size(); stats[7] = w.straights.size();
    }
    p3_reads_free(rd);
    if (rc) { g_host_err = "cannot write GFA"; return rc; }
    log.line("finish");
    return P3_OK;
}

}  // extern "C"
