// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Builds oracle/_ref/libp3ref.so from the UNMODIFIED reference sources where they lie under
// /root/reference (see oracle/Makefile; nothing from the reference is copied into this repo).
// The reference has no library/FFI surface — main.cpp #includes every .cpp into one
// translation unit — so this driver is that same single TU with main() replaced by a small
// C ABI that lets tests/bench run the reference's own stages one at a time and read back
// their state:
//   ReadFile::LoadFile / CountShortKmer      (reference src/Load.cpp)
//   Options::EstimateBloomfilter             (reference src/Options.cpp)
//   MakeBF<LARGE_BITSET>                     (reference src/MakeBloomFilter.cpp)
//   BF<Key>::add / possiblyContains          (reference src/bloomfilter.cpp)
//   DeBruijnGraph<LARGE_BITSET>::*           (reference src/DeBruijnGraph.cpp)
// The reference's Assemble_k (src/Assemble.cpp:30) only instantiates k in
// {5,21,25,63,101,501,1001,2001,3001}; the templates are generic in the bitset width, so
// this driver also instantiates the widths BASELINE.json's configs name (k=32) plus a few
// neighbours used by the parity tests.
//
// `#define private public` is only there to read BF::m_bits (bloomfilter.cpp:28); no
// reference behaviour changes.

#include <iostream>
#include <unistd.h>
#include <fstream>
#include <sstream>
#include <unordered_map>
#include <array>
#include <vector>
#include <set>
#include <string>
#include <bitset>
#include <tuple>
#include <queue>
#include <cmath>
#include <algorithm>
#include <deque>
#include <mutex>
#include <thread>
#include <cstring>
#include <memory>
#include <omp.h>

#define private public
#include "common.h"
#include "ShowInfo.cpp"
#include "Options.cpp"
#include "Load.cpp"
#include "bloomfilter.cpp"
#include "MakeBloomFilter.cpp"
#include "DeBruijnGraph.cpp"
#include "Assemble.cpp"
#include "Logging.cpp"
#undef private

namespace {

struct StageBase {
    virtual ~StageBase() {}
    virtual void make_bf(ReadFile &rf, Options &opt, Logging &lg) = 0;
    virtual void bf_bits(uint8_t *out) = 0;
    virtual uint64_t bf_size() = 0;
    virtual int check_directions(const char *kmer, int ignored) = 0;
    virtual int is_recorded(const char *kmer) = 0;
    virtual void bf_add(const char *kmer) = 0;
    virtual void bf_reset(uint64_t size, int nh) = 0;
    virtual uint64_t std_hash(const char *kmer) = 0;
    virtual void double_hash(const char *kmer, uint64_t *out2) = 0;
    virtual void make_dbg(Options &opt, Logging &lg) = 0;
    virtual void count_node_coverage(ReadFile &rf) = 0;
    virtual void print_graph() = 0;
    virtual uint64_t n_junctions() = 0;
    virtual uint64_t n_joints() = 0;
    virtual uint64_t n_straights() = 0;
    virtual void canonical(const char *kmer, char *out) = 0;
    std::set<std::string> seeds;
    Logging *lgp = nullptr;
};

template <int K>
struct Stage : StageBase {
    typedef std::bitset<2 * K> BS;
    BF<BS> bf;
    std::unique_ptr<DeBruijnGraph<BS>> dbg;

    void make_bf(ReadFile &rf, Options &opt, Logging &lg) override {
        seeds.clear();
        bf = MakeBF<BS>(rf.reads, rf.shortk_database, opt.filter_size, opt.num_hashes,
                        opt.kmer_length, &seeds, lg);
        lgp = &lg;
    }
    uint64_t bf_size() override { return bf.m_bits.size(); }
    void bf_bits(uint8_t *out) override {
        uint64_t n = bf.m_bits.size();
        memset(out, 0, (n + 7) / 8);
        for (uint64_t i = 0; i < n; i++)
            if (bf.m_bits[i]) out[i >> 3] |= (uint8_t)(1u << (i & 7));
    }
    void bf_reset(uint64_t size, int nh) override { bf.Set_BF(size, (uint8_t)nh); }
    void bf_add(const char *kmer) override {
        BS f = GetFirstKmerForward<BS>(std::string(kmer, K));
        BS b = GetComplementKmer(f);
        BS c = CompareBit(f, b, 2 * K);
        bf.add(&c, 2 * K);
    }
    void ensure_dbg(Logging &lg) {
        if (!dbg) dbg.reset(new DeBruijnGraph<BS>(K, bf, lg));
    }
    int check_directions(const char *kmer, int ignored) override {
        ensure_dbg(*lgp);
        BS t = GetFirstKmerForward<BS>(std::string(kmer, K));
        std::vector<BS> l, r;
        dbg->CheckDirections(&l, &r, t, ignored);
        // re-derive which of the 8 directions answered (CheckDirections only returns k-mers)
        int mask = 0;
        BS back = (t >> 2), front = (t << 2);
        for (int i = 0; i < 8; i++) {
            BS adj = (i < 4) ? (back | dbg->end_bases[i]) : (front | dbg->end_bases[i]);
            const std::vector<BS> &v = (i < 4) ? l : r;
            if (std::find(v.begin(), v.end(), adj) != v.end()) mask |= (1 << i);
        }
        return mask;
    }
    int is_recorded(const char *kmer) override {
        ensure_dbg(*lgp);
        BS t = GetFirstKmerForward<BS>(std::string(kmer, K));
        return dbg->IsRecorded(bf, t) ? 1 : 0;
    }
    uint64_t std_hash(const char *kmer) override {
        BS t = GetFirstKmerForward<BS>(std::string(kmer, K));
        return std::hash<BS>()(t);
    }
    void double_hash(const char *kmer, uint64_t *out2) override {
        BS t = GetFirstKmerForward<BS>(std::string(kmer, K));
        GetDoubleHash_64bit<BS>(&t, out2);
    }
    void canonical(const char *kmer, char *out) override {
        BS f = GetFirstKmerForward<BS>(std::string(kmer, K));
        BS b = GetFirstKmerBackward<BS>(std::string(kmer, K));
        BS c = CompareBit(f, b, 2 * K);
        std::string s = GetStringKmer<BS>(c);
        memcpy(out, s.data(), K);
    }
    void make_dbg(Options &opt, Logging &lg) override {
        dbg.reset(new DeBruijnGraph<BS>(K, bf, lg));
        dbg->MakeDBG(seeds, opt.filter_size, opt.num_hashes, opt.threads_num);
    }
    void count_node_coverage(ReadFile &rf) override { dbg->CountNodeCoverage(rf.reads); }
    void print_graph() override { dbg->PrintGraph(); }
    uint64_t n_junctions() override { return dbg ? dbg->junctions.size() : 0; }
    uint64_t n_joints() override { return dbg ? dbg->joints.size() : 0; }
    uint64_t n_straights() override { return dbg ? dbg->straights.size() : 0; }
};

StageBase *make_stage(int k) {
    switch (k) {
#define P3REF_K(KK) case KK: return new Stage<KK>();
        P3REF_K(21) P3REF_K(22) P3REF_K(25) P3REF_K(27) P3REF_K(31) P3REF_K(32) P3REF_K(33)
        P3REF_K(47) P3REF_K(63) P3REF_K(64) P3REF_K(65) P3REF_K(101) P3REF_K(501)
        P3REF_K(1001) P3REF_K(3001)
#undef P3REF_K
        default: return nullptr;
    }
}

struct RefRun {
    Logging lg;
    Options opt;
    std::unique_ptr<ReadFile> rf;
    std::unique_ptr<StageBase> st;
    std::vector<std::pair<uint64_t, uint64_t>> short_sorted;
};

}  // namespace

extern "C" {

// k values this build instantiates (0-terminated)
const int *p3ref_supported_k() {
    static const int ks[] = {21, 22, 25, 27, 31, 32, 33, 47, 63, 64, 65, 101, 501, 1001, 3001, 0};
    return ks;
}

// Options + ReadFile ctor; log_path NULL -> /dev/null (Logging::WriteLog reopens per line).
void *p3ref_new(const char *readfile, int k, uint64_t m, int threads, const char *log_path) {
    RefRun *r = new RefRun();
    r->lg.log_file = log_path ? log_path : "/dev/null";
    r->opt.readfile_name = readfile ? readfile : "mem.fasta";
    r->opt.kmer_length = (uint32_t)k;
    r->opt.filter_size = m;
    r->opt.threads_num = threads;
    r->st.reset(make_stage(k));
    if (!r->st) { delete r; return nullptr; }
    r->rf.reset(new ReadFile(r->opt));
    return r;
}
void p3ref_free(void *h) { delete (RefRun *)h; }

// main.cpp:22
void p3ref_load_file(void *h) { ((RefRun *)h)->rf->LoadFile(); }
// in-memory equivalent of LoadFasta's effect on `reads`/`all_bases` (Load.cpp:59-62), for the
// timed CPU baseline where no file exists. Names must be unique.
void p3ref_add_read(void *h, const char *name, const char *seq, uint64_t len) {
    RefRun *r = (RefRun *)h;
    if (len >= r->rf->large_kmer_length) {
        r->rf->reads[std::string(name)] = std::string(seq, len);
        r->rf->all_bases += len;
    }
}
// batch form: reads seq[off[i]..off[i+1]) named ">r<i>"
void p3ref_add_reads(void *h, const char *seq, const uint64_t *off, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        std::string name = ">r" + std::to_string(i);
        p3ref_add_read(h, name.c_str(), seq + off[i], off[i + 1] - off[i]);
    }
}
uint64_t p3ref_all_bases(void *h) { return ((RefRun *)h)->rf->all_bases; }
uint64_t p3ref_n_reads(void *h) { return ((RefRun *)h)->rf->reads.size(); }
// dump reads (sorted by name for determinism): lens[i], then concatenated into seq
uint64_t p3ref_reads_export(void *h, uint64_t *lens, char *seq) {
    RefRun *r = (RefRun *)h;
    std::vector<const std::pair<const std::string, std::string> *> v;
    for (auto &kv : r->rf->reads) v.push_back(&kv);
    std::sort(v.begin(), v.end(), [](auto a, auto b) { return a->first < b->first; });
    uint64_t o = 0, i = 0;
    for (auto p : v) {
        if (lens) lens[i] = p->second.size();
        if (seq) memcpy(seq + o, p->second.data(), p->second.size());
        o += p->second.size();
        i++;
    }
    return o;
}

// main.cpp:23
void p3ref_estimate(void *h) {
    RefRun *r = (RefRun *)h;
    r->opt.EstimateBloomfilter(r->rf->all_bases, r->lg);
}
void p3ref_estimate_only(uint64_t all_bases, int k, uint64_t *filter_size, int *num_hashes) {
    Logging lg; lg.log_file = "/dev/null";
    Options o; o.kmer_length = (uint32_t)k;
    o.EstimateBloomfilter(all_bases, lg);
    *filter_size = o.filter_size; *num_hashes = o.num_hashes;
}
uint64_t p3ref_filter_size(void *h) { return ((RefRun *)h)->opt.filter_size; }
int p3ref_num_hashes(void *h) { return ((RefRun *)h)->opt.num_hashes; }
void p3ref_set_filter(void *h, uint64_t m, int nh) {
    ((RefRun *)h)->opt.filter_size = m; ((RefRun *)h)->opt.num_hashes = (uint8_t)nh;
}

// Assemble.cpp:9
void p3ref_count_short(void *h) {
    RefRun *r = (RefRun *)h;
    r->rf->CountShortKmer(r->opt.shortk_length);
}
uint64_t p3ref_short_size(void *h) { return ((RefRun *)h)->rf->shortk_database.size(); }
// sorted by key ascending
void p3ref_short_export(void *h, uint64_t *keys, uint64_t *counts) {
    RefRun *r = (RefRun *)h;
    std::vector<std::pair<uint64_t, uint64_t>> v;
    v.reserve(r->rf->shortk_database.size());
    for (auto &kv : r->rf->shortk_database) v.push_back({kv.first.to_ullong(), kv.second});
    std::sort(v.begin(), v.end());
    for (size_t i = 0; i < v.size(); i++) { keys[i] = v[i].first; counts[i] = v[i].second; }
}

// Assemble.cpp:12
void p3ref_make_bf(void *h) {
    RefRun *r = (RefRun *)h;
    r->st->make_bf(*r->rf, r->opt, r->lg);
}
uint64_t p3ref_bf_size(void *h) { return ((RefRun *)h)->st->bf_size(); }
void p3ref_bf_bits(void *h, uint8_t *out) { ((RefRun *)h)->st->bf_bits(out); }
uint64_t p3ref_seed_count(void *h) { return ((RefRun *)h)->st->seeds.size(); }
// std::set order, k chars each, no separators
void p3ref_seed_export(void *h, char *out) {
    RefRun *r = (RefRun *)h;
    uint64_t o = 0;
    for (auto &s : r->st->seeds) { memcpy(out + o, s.data(), s.size()); o += s.size(); }
}

// stand-alone filter ops (bloomfilter.cpp:46,69) for known-answer tests
void p3ref_bf_reset(void *h, uint64_t size, int nh) {
    RefRun *r = (RefRun *)h; r->st->lgp = &r->lg; r->st->bf_reset(size, nh);
}
void p3ref_bf_add(void *h, const char *kmer) { ((RefRun *)h)->st->bf_add(kmer); }

// DeBruijnGraph.cpp:326 / :318 on the ORIENTED k-mer given as a string
int p3ref_check_directions(void *h, const char *kmer, int ignored) {
    RefRun *r = (RefRun *)h; if (!r->st->lgp) r->st->lgp = &r->lg;
    return r->st->check_directions(kmer, ignored);
}
// CheckDirections over n ORIENTED k-mers given as n*k characters (the per-k-mer neighbour
// queries MakeDBG issues, DeBruijnGraph.cpp:164,232,269, without the walk around them)
void p3ref_check_directions_batch(void *h, const char *kmers, uint64_t n, uint8_t *out) {
    RefRun *r = (RefRun *)h; if (!r->st->lgp) r->st->lgp = &r->lg;
    int k = (int)r->opt.kmer_length;
    for (uint64_t i = 0; i < n; i++) out[i] = (uint8_t)r->st->check_directions(kmers + i * k, -1);
}
int p3ref_is_recorded(void *h, const char *kmer) {
    RefRun *r = (RefRun *)h; if (!r->st->lgp) r->st->lgp = &r->lg;
    return r->st->is_recorded(kmer);
}
uint64_t p3ref_std_hash(void *h, const char *kmer) { return ((RefRun *)h)->st->std_hash(kmer); }
void p3ref_double_hash(void *h, const char *kmer, uint64_t *out2) {
    ((RefRun *)h)->st->double_hash(kmer, out2);
}
void p3ref_canonical(void *h, const char *kmer, char *out) { ((RefRun *)h)->st->canonical(kmer, out); }

// Assemble.cpp:19-26. PrintGraph writes ./de_bruijn_graph.gfa in the CURRENT directory.
void p3ref_make_dbg(void *h) { RefRun *r = (RefRun *)h; r->st->make_dbg(r->opt, r->lg); }
void p3ref_count_node_coverage(void *h) { RefRun *r = (RefRun *)h; r->st->count_node_coverage(*r->rf); }
void p3ref_print_graph(void *h) { ((RefRun *)h)->st->print_graph(); }
uint64_t p3ref_n_junctions(void *h) { return ((RefRun *)h)->st->n_junctions(); }
uint64_t p3ref_n_joints(void *h) { return ((RefRun *)h)->st->n_joints(); }
uint64_t p3ref_n_straights(void *h) { return ((RefRun *)h)->st->n_straights(); }

}  // extern "C"
