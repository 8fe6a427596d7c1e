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
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

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

// read-only view of a whole file (mmap; a plain read() into memory where mmap is refused)
struct FileView {
    const char *data = nullptr; size_t size = 0; bool mapped = false; std::vector<char> owned;
    bool open(const char *path) {
        int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); return false; }
        size = (size_t)st.st_size;
        if (size) {
            void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p != MAP_FAILED) {
                madvise(p, size, MADV_SEQUENTIAL);
                data = (const char *)p; mapped = true;
            } else {
                owned.resize(size);
                size_t got = 0;
                while (got < size) {
                    ssize_t n = ::read(fd, owned.data() + got, size - got);
                    if (n <= 0) break;
                    got += (size_t)n;
                }
                size = got; data = owned.data();
            }
        }
        ::close(fd);
        return true;
    }
    ~FileView() { if (mapped) munmap((void *)data, size); }
};

}  // namespace

extern "C" {

// Load in three phases over the mapped file, two of them parallel (the first version — std::getline
// plus a vector of strings — loaded 60 Mbases/s: 80 s for configs[1]'s 5 Gbp against 0.4 s for the
// GPU hot path; a single memchr pass with an open-addressing name index reached 250):
//   A  (threads) newlines per chunk -> the line number every chunk starts at; then each thread turns
//      the lines that START in its chunk into events: header line / run of sequence lines, with the
//      name hash computed on the spot
//   B  (serial, a few ns per event) the reference's line loop verbatim over the events
//      (src/Load.cpp:51-103): FASTA or single-line FASTQ chosen by the first byte of the first line,
//      a record is added when the NEXT header arrives (only if the current name line is non-empty;
//      otherwise the sequence keeps growing) and once more at the end of the file, reads shorter
//      than k are dropped, a repeated name line replaces the earlier record while all_bases counts both
//   C  (threads) the sequence bytes of the surviving records are copied to their final offsets
//      (lines as std::getline cuts them: the '\n' goes, a '\r' stays), then p3_pack_reads
int p3_load_file(const char *path, uint32_t k, p3_reads **out) {
    if (!path || !out) return P3_ERR_ARG;
    // reference src/Load.cpp:26: file_name.substr(size-5, 5) throws for names shorter than 5
    if (strlen(path) < 5) { g_host_err = "read file name shorter than 5 characters"; return P3_ERR_ARG; }
    const bool timing = getenv("P3_LOAD_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_start = now();
    FileView fv;
    if (!fv.open(path)) { g_host_err = std::string("cannot open ") + path; return P3_ERR_IO; }
    p3_reads *r = new p3_reads();
    const char *const data = fv.data, *const fend = fv.data + fv.size;

    int mode = 0;   // 1 FASTA, 2 FASTQ, decided by the first byte of the first line (Load.cpp:40-48)
    if (fv.size) {
        const char *nl = (const char *)memchr(data, '\n', fv.size);
        const size_t len0 = nl ? (size_t)(nl - data) : fv.size;
        if (len0 && data[0] == '>') mode = 1;
        else if (len0 && data[0] == '@') mode = 2;
    }

    struct Event { const char *p; uint64_t a, b; };   // header: p = line, a = length, b = name hash | 1
                                                      // sequence run: p = first line, a = bytes up to the run's end, b = 0
    auto name_hash = [](const char *q, size_t n) -> uint64_t {
        uint64_t x = 0x9E3779B97F4A7C15ULL ^ (n * 0xff51afd7ed558ccdULL);
        while (n >= 8) { uint64_t v; memcpy(&v, q, 8); x = (x ^ v) * 0xc4ceb9fe1a85ec53ULL; x ^= x >> 29; q += 8; n -= 8; }
        uint64_t v = 0; memcpy(&v, q, n);
        x = (x ^ v) * 0xff51afd7ed558ccdULL; x ^= x >> 32; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 29;
        return x | 1;
    };

    // ---- phase A -------------------------------------------------------------------------------------
    unsigned n_thr = std::thread::hardware_concurrency();
    n_thr = std::max(1u, std::min(n_thr ? n_thr : 1u, 32u));
    size_t chunk = std::max<size_t>((fv.size + n_thr - 1) / n_thr, 1 << 20);
    if (const char *e = getenv("P3_LOAD_CHUNK_BYTES")) chunk = std::max<size_t>(strtoull(e, nullptr, 10), 1);   // tests: tiny chunks
    const size_t n_chunks = mode ? (fv.size + chunk - 1) / chunk : 0;
    if (n_chunks > (1u << 20)) { delete r; g_host_err = "P3_LOAD_CHUNK_BYTES too small for this file"; return P3_ERR_ARG; }
    std::vector<uint64_t> nl_before(n_chunks + 1, 0);
    std::vector<std::vector<Event>> events(n_chunks);
    auto run_chunks = [&](auto fn) {      // chunks are dealt out dynamically to n_thr threads
        std::atomic<size_t> next{0};
        auto body = [&] { for (size_t c; (c = next.fetch_add(1)) < n_chunks;) fn(c); };
        if (n_thr == 1 || n_chunks <= 1) { body(); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < std::min<size_t>(n_thr, n_chunks); t++) th.emplace_back(body);
        for (auto &x : th) x.join();
    };
    run_chunks([&](size_t c) {
        const char *q = data + c * chunk, *e = std::min(fend, q + chunk);
        uint64_t n = 0;
        while (q < e) { const char *nl = (const char *)memchr(q, '\n', (size_t)(e - q)); if (!nl) break; n++; q = nl + 1; }
        nl_before[c + 1] = n;
    });
    for (size_t c = 0; c < n_chunks; c++) nl_before[c + 1] += nl_before[c];
    run_chunks([&](size_t c) {
        const char *cb = data + c * chunk, *ce = std::min(fend, cb + chunk);
        // first line that STARTS in this chunk, and its line number
        const char *q = cb;
        uint64_t line = nl_before[c];
        if (c && cb[-1] != '\n') {
            const char *nl = (const char *)memchr(cb, '\n', (size_t)(fend - cb));
            if (!nl) return;                    // the rest of the file is one line that started earlier
            if (nl >= ce) return;               // ... or that line ends beyond this chunk: no line starts here
            q = nl + 1; line = nl_before[c] + 1;
        }
        std::vector<Event> &ev = events[c];
        const char *run = nullptr;              // open run of sequence lines
        while (q < ce) {
            const char *nl = (const char *)memchr(q, '\n', (size_t)(fend - q));
            const char *le = nl ? nl : fend;
            bool header, is_seq;
            if (mode == 1) { header = le > q && q[0] == '>'; is_seq = !header; }
            else { header = line % 4 == 0; is_seq = line % 4 == 1; }
            if (header) {
                if (run) { ev.push_back({run, (uint64_t)(q - run), 0}); run = nullptr; }
                ev.push_back({q, (uint64_t)(le - q), name_hash(q, (size_t)(le - q))});
            } else if (is_seq) {
                if (!run) run = q;
                if (mode == 2) { ev.push_back({run, (uint64_t)((nl ? nl + 1 : fend) - run), 0}); run = nullptr; }
            }
            q = nl ? nl + 1 : fend;
            line++;
        }
        if (run) ev.push_back({run, (uint64_t)(q - run), 0});
    });
    auto t_events = now();

    // ---- phase B -------------------------------------------------------------------------------------
    struct Rec { uint64_t len; const char *name; uint32_t name_len; uint32_t n_runs; uint64_t run0; bool alive; };
    std::vector<Rec> recs;
    std::vector<std::pair<const char *, uint64_t>> runs;   // (first byte, bytes) of every sequence run of every record
    struct NameIndex {      // name line -> latest record with that name: open addressing over (hash, record)
        std::vector<uint64_t> h; std::vector<uint32_t> rec; uint64_t mask = 0, used = 0;
        void grow() {
            uint64_t cap = mask ? (mask + 1) * 4 : (1u << 16);
            std::vector<uint64_t> nh(cap, 0); std::vector<uint32_t> nr(cap, 0);
            for (uint64_t i = 0; mask && i <= mask; i++)
                if (h[i]) { uint64_t s = h[i] & (cap - 1); while (nh[s]) s = (s + 1) & (cap - 1); nh[s] = h[i]; nr[s] = rec[i]; }
            h.swap(nh); rec.swap(nr); mask = cap - 1;
        }
    } index;
    // bases of a run = its bytes minus its '\n's (a '\r' counts: std::getline keeps it)
    auto run_bases = [&](const char *q, uint64_t bytes) -> uint64_t {
        uint64_t n = bytes;
        const char *e = q + bytes;
        while (q < e) { const char *nl = (const char *)memchr(q, '\n', (size_t)(e - q)); if (!nl) break; n--; q = nl + 1; }
        return n;
    };
    uint64_t all_bases = 0, dead = 0, total = 0;
    bool too_many = false;
    const char *name = nullptr; uint64_t name_len = 0, name_h = 0;   // current header line (empty before the first)
    uint64_t cur_run0 = 0, cur_len = 0;                               // current sequence = runs[cur_run0 ..), cur_len bases
    auto add_record = [&] {
        if (cur_len >= k) {
            if (recs.size() >= 0xFFFFFFFFull) { too_many = true; return; }
            if (!index.mask || 2 * (index.used + 1) > index.mask + 1) index.grow();
            const uint64_t hv = name ? name_h : 1;
            uint64_t s = hv & index.mask;
            for (;; s = (s + 1) & index.mask) {
                if (!index.h[s]) { index.h[s] = hv; index.rec[s] = (uint32_t)recs.size(); index.used++; break; }
                if (index.h[s] != hv) continue;
                Rec &o = recs[index.rec[s]];
                if (o.name_len == name_len && (name_len == 0 || memcmp(o.name, name, name_len) == 0)) {
                    o.alive = false; dead++; total -= o.len;     // the later record replaces the earlier
                    index.rec[s] = (uint32_t)recs.size();
                    break;
                }
            }
            recs.push_back({cur_len, name, (uint32_t)name_len, (uint32_t)(runs.size() - cur_run0), cur_run0, true});
            all_bases += cur_len; total += cur_len;
        } else {
            runs.resize(cur_run0);                               // dropped: shorter than k
        }
        cur_run0 = runs.size(); cur_len = 0;
    };
    {   // one allocation each for the record and run lists, and a name table that will not have to grow
        size_t n_ev = 0;
        for (size_t c = 0; c < n_chunks; c++) n_ev += events[c].size();
        recs.reserve(n_ev / 2 + 16); runs.reserve(n_ev / 2 + 16);
        while (index.mask + 1 < n_ev + 16 && index.mask + 1 < (1ull << 33)) index.grow();
    }
    for (size_t c = 0; c < n_chunks; c++) {
        const std::vector<Event> &evs = events[c];
        for (size_t ei = 0; ei < evs.size(); ei++) {
            const Event &e = evs[ei];
            if (ei + 16 < evs.size() && evs[ei + 16].b) {        // the name table is far larger than the caches
                __builtin_prefetch(&index.h[evs[ei + 16].b & index.mask]);
                __builtin_prefetch(&index.rec[evs[ei + 16].b & index.mask]);
            }
            if (e.b) {                                           // header line
                if (name_len) add_record();                      // (the sequence is NOT cleared when the name is empty)
                name = e.p; name_len = e.a; name_h = e.b;
            } else {
                runs.emplace_back(e.p, e.a);
                cur_len += mode == 2 ? e.a - (e.a && e.p[e.a - 1] == '\n' ? 1 : 0) : run_bases(e.p, e.a);
            }
        }
        std::vector<Event>().swap(events[c]);
    }
    if (mode && cur_len >= k) add_record();                      // Load.cpp:71-74
    if (too_many) { delete r; g_host_err = "more than 2^32 reads in one file"; return P3_ERR_ARG; }
    auto t_parsed = now();

    // ---- phase C -------------------------------------------------------------------------------------
    r->all_bases = all_bases;
    r->off.reserve(recs.size() - dead + 1);
    r->off.push_back(0);
    std::vector<uint32_t> alive;
    alive.reserve(recs.size() - dead);
    for (size_t i = 0; i < recs.size(); i++)
        if (recs[i].alive) { alive.push_back((uint32_t)i); r->off.push_back(r->off.back() + recs[i].len); }
    r->seq.resize(total);
    {
        char *dst0 = &r->seq[0];
        const size_t n_alive = alive.size();
        const size_t per = std::max<size_t>((n_alive + 8 * n_thr - 1) / (8 * n_thr), 1);
        const size_t n_jobs = (n_alive + per - 1) / per;
        std::atomic<size_t> next{0};
        auto body = [&] {
            for (size_t j; (j = next.fetch_add(1)) < n_jobs;) {
                for (size_t i = j * per; i < std::min(n_alive, (j + 1) * per); i++) {
                    const Rec &rc = recs[alive[i]];
                    char *d = dst0 + r->off[i];
                    for (uint32_t u = 0; u < rc.n_runs; u++) {
                        const char *q = runs[rc.run0 + u].first, *e = q + runs[rc.run0 + u].second;
                        while (q < e) {
                            const char *nl = (const char *)memchr(q, '\n', (size_t)(e - q));
                            const char *le = nl ? nl : e;
                            memcpy(d, q, (size_t)(le - q));
                            d += le - q;
                            q = nl ? nl + 1 : e;
                        }
                    }
                }
            }
        };
        if (n_thr == 1 || n_jobs <= 1) body();
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < std::min<size_t>(n_thr, n_jobs); t++) th.emplace_back(body);
            for (auto &x : th) x.join();
        }
    }
    auto t_copied = now();
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
    auto t_alloc = now();
    p3_pack_reads(r->seq.data(), r->off.data(), r->off.size() - 1, r->packed, r->nmask, &r->has_non_acgt);
    if (timing) {
        auto t_end = now();
        auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
        fprintf(stderr, "p3_load_file: events %.3f s, line loop %.3f s, copy %.3f s, staging alloc %.3f s, pack %.3f s "
                        "(%zu bytes, %zu reads, %u threads)\n",
                sec(t_start, t_events), sec(t_events, t_parsed), sec(t_parsed, t_copied), sec(t_copied, t_alloc),
                sec(t_alloc, t_end), fv.size, r->off.size() - 1, n_thr);
    }
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
// unitig walk over the exported adjacency table. One template, two k-mer representations:
//   U64Ops  k <= 32: one 64-bit word per k-mer (first base in the top bits of the 2k-bit value)
//   StrOps  k >  32: the k-mer as its ACGT string (what GetStringKmer prints); the reference's
//           std::bitset<2k> order == the string order, so "canonical" and the seed order carry over
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
inline int fcode(unsigned char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }
// base_to_bit[trans_base[c]] of the reference: complement code, 0 for anything that is not ACGT
inline int rcode(unsigned char c) { return c == 'A' ? 3 : c == 'C' ? 2 : c == 'G' ? 1 : 0; }

// canonical k-mers of every node recorded so far: is_visited (4 map lookups in the reference, :300-315)
// is one probe here. Flat open addressing for uint64 k-mers, a hash set of strings for long ones.
struct VisitedU64 {
    std::vector<uint64_t> keys; uint64_t mask = 0, used = 0;   // ~0 = empty (all-T is never canonical: all-A is smaller)
    void grow() {
        uint64_t cap = mask ? (mask + 1) * 2 : (1u << 12);
        std::vector<uint64_t> nk(cap, ~0ULL);
        for (uint64_t v : keys) if (v != ~0ULL) { uint64_t s = fmix64(v) & (cap - 1); while (nk[s] != ~0ULL) s = (s + 1) & (cap - 1); nk[s] = v; }
        keys.swap(nk); mask = cap - 1;
    }
    void insert(uint64_t v) {
        if (!mask || 2 * (used + 1) > mask + 1) grow();
        uint64_t s = fmix64(v) & mask;
        while (keys[s] != ~0ULL) { if (keys[s] == v) return; s = (s + 1) & mask; }
        keys[s] = v; used++;
    }
    bool contains(uint64_t v) const {
        if (!mask) return false;
        uint64_t s = fmix64(v) & mask;
        while (keys[s] != ~0ULL) { if (keys[s] == v) return true; s = (s + 1) & mask; }
        return false;
    }
};
struct VisitedStr {
    std::unordered_map<std::string, char> m;
    void insert(const std::string &v) { m.emplace(v, 1); }
    bool contains(const std::string &v) const { return m.count(v) != 0; }
};

struct U64Ops {
    using K = uint64_t;
    using Visited = VisitedU64;
    struct Hash { size_t operator()(uint64_t v) const { return (size_t)fmix64(v); } };
    int k; uint64_t kmask;
    explicit U64Ops(int k_) : k(k_), kmask(k_ >= 32 ? ~0ULL : ((1ULL << (2 * k_)) - 1)) {}
    K revcomp(K v) const {   // GetComplementKmer, reference src/BitCalc.cpp:36-45
        uint64_t x = ~v;
        x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
        x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
        x = __builtin_bswap64(x);
        return x >> (64 - 2 * k);
    }
    std::string str(K v) const {   // GetStringKmer, reference src/BitCalc.cpp:57-65
        std::string s(k, 'A');
        for (int i = 0; i < k; i++) s[i] = "ACGT"[(v >> (2 * (k - 1 - i))) & 3];
        return s;
    }
    K neighbour(K km, int d) const {   // reference src/DeBruijnGraph.cpp:327-339
        return d < 4 ? ((km >> 2) | ((uint64_t)d << (2 * k - 2))) : (((km << 2) | (uint64_t)(d - 4)) & kmask);
    }
    int first(K v) const { return (int)((v >> (2 * k - 2)) & 3); }
    int last(K v) const { return (int)(v & 3); }
    K from_read(const unsigned char *s) const {   // GetFirstKmerForward of s[0..k)
        uint64_t v = 0;
        for (int i = 0; i < k; i++) v = ((v << 2) | (uint64_t)fcode(s[i])) & kmask;
        return v;
    }
    K from_read_backward(const unsigned char *s) const {   // GetFirstKmerBackward
        uint64_t v = 0;
        for (int i = k - 1; i >= 0; i--) v = ((v << 2) | (uint64_t)rcode(s[i])) & kmask;
        return v;
    }
    K roll_fw(K v, unsigned char c) const { return ((v << 2) | (uint64_t)fcode(c)) & kmask; }
    K roll_bw(K v, unsigned char c) const { return (v >> 2) | ((uint64_t)rcode(c) << (2 * k - 2)); }
};

struct StrOps {
    using K = std::string;
    using Visited = VisitedStr;
    using Hash = std::hash<std::string>;
    int k;
    explicit StrOps(int k_) : k(k_) {}
    static char comp(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; }
    K revcomp(const K &v) const {
        K r(v.size(), 'A');
        for (size_t i = 0; i < v.size(); i++) r[i] = comp(v[v.size() - 1 - i]);
        return r;
    }
    std::string str(const K &v) const { return v; }
    K neighbour(const K &km, int d) const {
        if (d < 4) return std::string(1, "ACGT"[d]) + km.substr(0, k - 1);
        return km.substr(1) + "ACGT"[d - 4];
    }
    int first(const K &v) const { return fcode((unsigned char)v[0]); }
    int last(const K &v) const { return fcode((unsigned char)v[k - 1]); }
    K from_read(const unsigned char *s) const {
        K v(k, 'A');
        for (int i = 0; i < k; i++) v[i] = "ACGT"[fcode(s[i])];
        return v;
    }
    K from_read_backward(const unsigned char *s) const {
        K v(k, 'A');
        for (int i = 0; i < k; i++) v[i] = "ACGT"[rcode(s[k - 1 - i])];
        return v;
    }
    K roll_fw(const K &v, unsigned char c) const { return v.substr(1) + "ACGT"[fcode(c)]; }
    K roll_bw(const K &v, unsigned char c) const { return std::string(1, "ACGT"[rcode(c)]) + v.substr(0, k - 1); }
    // little-endian words of the bitset<2k> value (what the device and p3_check_directions use)
    void to_words(const K &v, uint64_t *w, int W) const {
        for (int j = 0; j < W; j++) w[j] = 0;
        for (int i = 0; i < k; i++) {
            int bit = 2 * (k - 1 - i);
            w[bit >> 6] |= (uint64_t)fcode((unsigned char)v[i]) << (bit & 63);
        }
    }
    K from_words(const uint64_t *w) const {
        K v(k, 'A');
        for (int i = 0; i < k; i++) {
            int bit = 2 * (k - 1 - i);
            v[i] = "ACGT"[(w[bit >> 6] >> (bit & 63)) & 3];
        }
        return v;
    }
};

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
struct StrAdjTable {
    std::unordered_map<std::string, uint8_t> m;
    bool find(const std::string &key, uint8_t *v) const {
        auto it = m.find(key);
        if (it == m.end()) return false;
        *v = it->second;
        return true;
    }
};

struct Junction { int id = 0; int coverage = 0; int left_cov[4] = {0, 0, 0, 0}; int right_cov[4] = {0, 0, 0, 0}; };
struct Joint { int id = 0; int coverage = 0; int straight = 0; };
struct Straight { int id = 0; std::string sequence; };

template <class O, class A>
struct Walker {
    using K = typename O::K;
    O ops;
    int k;
    const A &adj;
    std::unordered_map<K, Junction, typename O::Hash> junctions;
    std::unordered_map<K, Joint, typename O::Hash> joints;
    std::vector<Straight> straights;   // id = index + 1
    std::deque<K> visiting;
    int junction_id = 0, joint_id = 0;
    uint64_t missing = 0;              // CheckDirections on a k-mer the table lacks (must stay 0)

    Walker(int k_, const A &a) : ops(k_), k(k_), adj(a) {}

    K revcomp(const K &v) const { return ops.revcomp(v); }
    std::string str(const K &v) const { return ops.str(v); }
    K neighbour(const K &km, int d) const { return ops.neighbour(km, d); }
    // 8 CheckDirections bits of an ORIENTED k-mer: the table stores the canonical orientation;
    // for the other strand direction i of K is direction 7-i of revcomp(K)
    uint8_t directions(const K &km) {
        K rc = revcomp(km);
        uint8_t a = 0;
        if (km <= rc) { if (!adj.find(km, &a)) missing++; return a; }
        if (!adj.find(rc, &a)) missing++;
        return rev8(a);
    }
    struct Nb {   // at most 4 neighbours on a side: no heap
        K v[4]; int n = 0;
        void clear() { n = 0; }
        void push_back(const K &x) { v[n++] = x; }
        int size() const { return n; }
        bool empty() const { return n == 0; }
        const K &operator[](int i) const { return v[i]; }
    };
    // CheckDirections, reference src/DeBruijnGraph.cpp:326-345
    void check_directions(Nb &left, Nb &right, const K &km, int ignored) {
        left.clear(); right.clear();
        uint8_t a = directions(km);
        for (int i = 0; i < 8; i++) {
            if (i == ignored || !((a >> i) & 1)) continue;
            (i < 4 ? left : right).push_back(neighbour(km, i));
        }
    }
    // reference src/DeBruijnGraph.cpp:300-315: km or its reverse complement is a junction or a joint.
    // Every key the two maps ever receive is mirrored (in canonical form) in `visited`.
    typename O::Visited visited;
    void mark(const K &km) { K rc = revcomp(km); visited.insert(km <= rc ? km : rc); }
    // k-mers strictly inside a recorded straight. The reference re-walks such a unitig from every
    // later seed that lies on it (every read has a seed) up to its joint, finds the joint visited and
    // returns without having changed anything: all k-mers on the way have exactly one neighbour on
    // each side (which neighbours a k-mer reports depends only on the neighbours' own filter
    // membership, so this holds in either walking direction) and none of them is a node. Remembering
    // them turns that walk into one lookup: ~10x fewer steps at 30x coverage, same graph.
    typename O::Visited covered;
    std::vector<K> path;            // k-mers passed by the current search_node (both extensions)
    bool is_covered(const K &km) const {
        K rc = ops.revcomp(km);
        return covered.contains(km <= rc ? km : rc);
    }
    bool is_visited(const K &km) const {
        K rc = ops.revcomp(km);
        return visited.contains(km <= rc ? km : rc);
    }
    void add_junction(const K &km) {   // :348-357
        if (is_visited(km)) return;
        junctions[km].id = ++junction_id;
        mark(km);
    }
    void add_joint(const K &km) {      // :360-369
        if (is_visited(km)) return;
        joints[km].id = ++joint_id;
        mark(km);
    }
    void add_straight(const std::string &seq, const K &lj, const K &rj) {   // :374-391
        if (is_visited(lj)) return;
        add_joint(lj);
        add_joint(rj);
        Straight s; s.id = (int)straights.size() + 1; s.sequence = seq;
        straights.push_back(s);
        joints[lj].straight = s.id;   // operator[] semantics: the entry exists afterwards even if
        joints[rj].straight = s.id;   // add_joint refused it (reference :385-389)
        mark(lj); mark(rj);
    }
    void push_all(const Nb &l, const Nb &r) {
        for (int i = 0; i < l.size(); i++) visiting.push_back(l[i]);
        for (int i = 0; i < r.size(); i++) visiting.push_back(r[i]);
    }
    K extend_left(K target, K previous, std::vector<char> &ext, int previous_base) {   // :229-260
        Nb l, r;
        check_directions(l, r, target, 4 + previous_base);
        while (l.size() == 1 && r.empty()) {
            if (is_visited(target)) { ext.clear(); return target; }
            path.push_back(target);
            ext.push_back("ACGT"[ops.first(target)]);
            previous_base = ops.last(target);
            previous = target;
            target = l[0];
            check_directions(l, r, target, 4 + previous_base);
        }
        if (!is_visited(target)) { push_all(l, r); add_junction(target); }
        return previous;
    }
    K extend_right(K target, K previous, std::vector<char> &ext, int previous_base) {  // :264-297
        Nb l, r;
        check_directions(l, r, target, previous_base);
        while (l.empty() && r.size() == 1) {
            if (is_visited(target)) { ext.clear(); return target; }
            path.push_back(target);
            ext.push_back("ACGT"[ops.last(target)]);
            previous_base = ops.first(target);
            previous = target;
            target = r[0];
            check_directions(l, r, target, previous_base);
        }
        if (!is_visited(target)) { push_all(l, r); add_junction(target); }
        return previous;
    }
    void search_node(const K &target) {   // :158-225
        if (is_visited(target)) return;
        if (is_covered(target)) return;   // inside a recorded straight: the reference's walk from here changes nothing
        path.clear();
        path.push_back(target);
        Nb l, r;
        check_directions(l, r, target, -1);
        if (l.size() != 1 || r.size() != 1) { push_all(l, r); add_junction(target); return; }
        std::vector<char> ext_l, ext_r;
        K left_end = extend_left(l[0], target, ext_l, ops.last(target));
        std::string left_part(ext_l.rbegin(), ext_l.rend());
        if (is_visited(left_end)) return;
        K right_end = extend_right(r[0], target, ext_r, ops.first(target));
        std::string right_part(ext_r.begin(), ext_r.end());
        if (is_visited(right_end)) return;
        if (left_end == right_end) { add_junction(left_end); return; }
        if (left_part.size() + right_part.size() >= 1) {
            add_straight(left_part + str(target) + right_part, left_end, right_end);
            for (const K &x : path) {          // the straight's k-mers; its two joints are nodes, not interior
                if (x == left_end || x == right_end) continue;
                K rc = revcomp(x);
                covered.insert(x <= rc ? x : rc);
            }
        }
    }
    // MakeDBG with threads_num = 1, reference src/DeBruijnGraph.cpp:94-155. seeds sorted ascending
    // (== std::set<std::string> order for ACGT strings of equal length).
    int make_dbg(const std::vector<K> &seeds, uint64_t node_limit) {
        uint64_t cnt = 0;
        for (const K &s : seeds) {
            cnt++;
            if (is_visited(s)) continue;
            visiting.push_back(s);
            if (cnt % 20 != 0) continue;
            while (!visiting.empty()) {
                K v = visiting.front(); visiting.pop_front();
                search_node(v);
                if (junctions.size() > node_limit) return -1;
            }
        }
        while (!visiting.empty()) {
            K v = visiting.front(); visiting.pop_front();
            search_node(v);
            if (junctions.size() > node_limit) return -1;
        }
        return 0;
    }
};
using Walker64 = Walker<U64Ops, AdjTable>;
using WalkerStr = Walker<StrOps, StrAdjTable>;

// CountNodeCoverage (reference src/DeBruijnGraph.cpp:394-439) for k <= 32 runs on the GPU
// (p3_node_coverage): the node k-mers go down in id order, the counters come back into the maps.
int count_node_coverage(Walker64 &w, p3_ctx *ctx) {
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

// The same on the host, a transcription of the reference loop (one rolling forward and one rolling
// backward k-mer per read). Used where the node k-mers are multi-word (k > 32; the reference does
// this stage on the CPU too) and by the CPU tests of the walk.
template <class W>
void count_node_coverage_host(W &w, const std::string &seq, const std::vector<uint64_t> &off) {
    const int k = w.k;
    auto add_node = [&](const typename W::K &km) {   // AddNodeCoverage, :442-449
        auto j = w.junctions.find(km);
        if (j != w.junctions.end()) j->second.coverage++;
        auto t = w.joints.find(km);
        if (t != w.joints.end()) t->second.coverage++;
    };
    for (size_t r = 0; r + 1 < off.size(); r++) {
        const unsigned char *rd = (const unsigned char *)seq.data() + off[r];
        const size_t len = (size_t)(off[r + 1] - off[r]);
        if (len < (size_t)k) continue;
        auto rdat = [&](size_t i) -> unsigned char { return i < len ? rd[i] : (unsigned char)0; };   // std::string[size()] == '\0'
        typename W::K fw = w.ops.from_read(rd), bw = w.ops.from_read_backward(rd);
        add_node(fw); add_node(bw);
        auto jf = w.junctions.find(fw);
        if (jf != w.junctions.end()) jf->second.right_cov[fcode(rdat(k))]++;
        else {
            auto jb = w.junctions.find(bw);
            if (jb != w.junctions.end()) jb->second.left_cov[rcode(rdat(k))]++;
        }
        for (size_t i = k; i < len; i++) {
            fw = w.ops.roll_fw(fw, rd[i]);
            bw = w.ops.roll_bw(bw, rd[i]);
            // neither k-mer is a node (the common case): one or two probes of the flat visited set
            // instead of four map lookups. (bw is not always revcomp(fw): a non-ACGT base reads as A on both.)
            if (!w.is_visited(fw) && !w.is_visited(bw)) continue;
            add_node(fw); add_node(bw);
            auto j1 = w.junctions.find(fw);
            if (j1 != w.junctions.end()) {
                j1->second.left_cov[fcode(rd[i - k])]++;
                if (i < len - 1) j1->second.right_cov[fcode(rd[i + 1])]++;
            } else {
                auto j2 = w.junctions.find(bw);
                if (j2 != w.junctions.end()) {
                    j2->second.right_cov[rcode(rd[i - k])]++;
                    if (i < len - 1) j2->second.left_cov[rcode(rd[i + 1])]++;
                }
            }
        }
    }
}

// The same for string k-mers without O(k) work per base: a 64-bit rolling hash of the forward and of
// the backward k-mer (polynomial in the 2-bit codes, arithmetic mod 2^64; the backward one slides by
// multiplying with the inverse of the base) is looked up in the set of node hashes first; only a hit
// builds the two strings and runs the reference's bookkeeping (nodes are few, so hits are
// ~nodes x coverage). Equal k-mers have equal hashes, so nothing is missed; a colliding non-node just
// fails the exact lookups.
void count_node_coverage_host(WalkerStr &w, const std::string &seq, const std::vector<uint64_t> &off) {
    const int k = w.k;
    const uint64_t B = 0x9E3779B97F4A7C15ULL;      // odd: invertible mod 2^64
    uint64_t Binv = B;                              // Newton iteration for the inverse
    for (int i = 0; i < 6; i++) Binv *= 2 - B * Binv;
    uint64_t Bk1 = 1;                               // B^(k-1)
    for (int i = 0; i < k - 1; i++) Bk1 *= B;
    auto hash_str = [&](const std::string &s) { uint64_t h = 0; for (char c : s) h = h * B + (uint64_t)fcode((unsigned char)c) + 1; return h; };
    std::unordered_map<uint64_t, char> node_hash;
    node_hash.reserve(2 * (w.junctions.size() + w.joints.size()) + 16);
    for (auto &kv : w.junctions) node_hash.emplace(hash_str(kv.first), 1);
    for (auto &kv : w.joints) node_hash.emplace(hash_str(kv.first), 1);
    auto add_node = [&](const std::string &km) {   // AddNodeCoverage, :442-449
        auto j = w.junctions.find(km);
        if (j != w.junctions.end()) j->second.coverage++;
        auto t = w.joints.find(km);
        if (t != w.joints.end()) t->second.coverage++;
    };
    for (size_t r = 0; r + 1 < off.size(); r++) {
        const unsigned char *rd = (const unsigned char *)seq.data() + off[r];
        const size_t len = (size_t)(off[r + 1] - off[r]);
        if (len < (size_t)k) continue;
        auto rdat = [&](size_t i) -> unsigned char { return i < len ? rd[i] : (unsigned char)0; };
        // digit = code + 1 so that leading 'A's count; fw digits: fcode(read[j]); bw digits: rcode(read[i]) first
        uint64_t hf = 0, hb = 0;
        for (int j = 0; j < k; j++) hf = hf * B + (uint64_t)fcode(rd[j]) + 1;
        for (int j = k - 1; j >= 0; j--) hb = hb * B + (uint64_t)rcode(rd[j]) + 1;
        for (size_t i = (size_t)k - 1; i < len; i++) {   // k-mer ending at base i
            if (i >= (size_t)k) {
                hf = (hf - ((uint64_t)fcode(rd[i - k]) + 1) * Bk1) * B + (uint64_t)fcode(rd[i]) + 1;
                hb = (hb - ((uint64_t)rcode(rd[i - k]) + 1)) * Binv + ((uint64_t)rcode(rd[i]) + 1) * Bk1;
            }
            if (!node_hash.count(hf) && !node_hash.count(hb)) continue;
            const std::string fw = w.ops.from_read(rd + i + 1 - k), bw = w.ops.from_read_backward(rd + i + 1 - k);
            add_node(fw); add_node(bw);
            auto jf = w.junctions.find(fw);
            auto jb = jf == w.junctions.end() ? w.junctions.find(bw) : w.junctions.end();
            if (i == (size_t)k - 1) {   // first k-mer of the read (:407-413)
                if (jf != w.junctions.end()) jf->second.right_cov[fcode(rdat(k))]++;
                else if (jb != w.junctions.end()) jb->second.left_cov[rcode(rdat(k))]++;
            } else if (jf != w.junctions.end()) {   // (:424-435)
                jf->second.left_cov[fcode(rd[i - k])]++;
                if (i < len - 1) jf->second.right_cov[fcode(rd[i + 1])]++;
            } else if (jb != w.junctions.end()) {
                jb->second.right_cov[rcode(rd[i - k])]++;
                if (i < len - 1) jb->second.left_cov[rcode(rd[i + 1])]++;
            }
        }
    }
}

// PrintGraph, reference src/DeBruijnGraph.cpp:452-544. The reference iterates unordered_maps, so its
// line order is unspecified; this writes straights and junctions by id.
template <class W>
int print_graph(W &w, const char *path) {
    using K = typename W::K;
    FILE *f = fopen(path, "w");
    if (!f) return P3_ERR_IO;
    const int k = w.k;
    fprintf(f, "H\tVN:Z:1.0\n");
    for (auto &s : w.straights) fprintf(f, "S\tStraight_%d\t%s\tKC:i:%zu\n", s.id, s.sequence.c_str(), s.sequence.size());
    std::vector<std::pair<int, K>> js;
    for (auto &kv : w.junctions) js.push_back({kv.second.id, kv.first});
    std::sort(js.begin(), js.end(), [](const std::pair<int, K> &a, const std::pair<int, K> &b) { return a.first < b.first; });
    for (auto &p : js) fprintf(f, "S\tJunction_%d\t%s\tKC:i:%d\n", p.first, w.str(p.second).c_str(), w.junctions[p.second].coverage * k);
    for (auto &p : js) {
        const K km = p.second;
        const Junction &J = w.junctions[km];
        const uint8_t a = w.directions(km);
        for (int i = 0; i < 4; i++) {
            if (J.left_cov[i] == 0 || !((a >> i) & 1)) continue;   // IsRecorded(output_left_kmer)
            K n = w.neighbour(km, i), nb = w.revcomp(n);
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
            K n = w.neighbour(km, 4 + i), nb = w.revcomp(n);
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

namespace {

// k > 32: the device exports the distinct solid k-mers as n x W words with their adjacency bytes;
// the walk works on strings. The table is closed under reported neighbours from the host: every
// neighbour that CheckDirections reports but the table lacks (a Bloom false positive, or a seed) is
// sent back to the GPU in batches (p3_check_directions answers the 8 queries of each), wave after
// wave until nothing new appears — the k <= 32 path does the same on the device (p3_dbg_close).
// Returns 0, a negative P3_ERR_* from the library (p3_last_error is set by it) or a POSITIVE
// -P3_ERR_* for host-side failures (message set here).
int close_table_long(p3_ctx *ctx, const StrOps &ops, int W, StrAdjTable &table, std::vector<std::string> frontier) {
    for (int round = 0; !frontier.empty(); round++) {
        if (round > 200) { g_host_err = "closure found no fixed point (filter saturated: the reference walk would not terminate either)"; return -P3_ERR_TABLE_FULL; }
        std::vector<std::string> want;
        std::unordered_map<std::string, char> seen;
        for (const std::string &km : frontier) {
            uint8_t a = 0;
            if (!table.find(km, &a)) continue;
            for (int d = 0; d < 8; d++) {
                if (!((a >> d) & 1)) continue;
                std::string n = ops.neighbour(km, d), rc = ops.revcomp(n);
                const std::string &c = n <= rc ? n : rc;
                uint8_t dummy;
                if (table.find(c, &dummy) || seen.count(c)) continue;
                seen.emplace(c, 1);
                want.push_back(c);
            }
        }
        if (want.empty()) break;
        std::vector<uint64_t> words(want.size() * (size_t)W);
        for (size_t i = 0; i < want.size(); i++) ops.to_words(want[i], words.data() + i * W, W);
        std::vector<uint8_t> masks(want.size());
        int rc = p3_check_directions(ctx, words.data(), want.size(), masks.data());
        if (rc) return rc;
        for (size_t i = 0; i < want.size(); i++) table.m.emplace(want[i], masks[i]);
        frontier.swap(want);
    }
    return 0;
}

int assemble_long(p3_ctx *ctx, p3_reads *rd, uint32_t k, const std::vector<int64_t> &seed_pos, Log &log,
                  const char *gfa_path, uint64_t *stats) {
    const StrOps ops((int)k);
    const int W = (int)((2 * k + 63) / 64);
    const uint64_t n_reads = p3_reads_count(rd);
    std::vector<std::string> seeds;
    for (uint64_t r = 0; r < n_reads; r++)
        if (seed_pos[r] >= 0) seeds.push_back(ops.from_read((const unsigned char *)rd->seq.data() + rd->off[r] + seed_pos[r]));
    std::sort(seeds.begin(), seeds.end());
    seeds.erase(std::unique(seeds.begin(), seeds.end()), seeds.end());
    log.line("seed kmer num= " + std::to_string(seeds.size()));

    int rc = p3_dbg_adjacency(ctx);
    if (rc) return rc;
    uint64_t n_solid = 0, n_edges = 0, n_pos = 0, n_d21 = 0, got = 0;
    p3_dbg_stats(ctx, &n_solid, &n_edges);
    p3_short_kmer_stats(ctx, &n_pos, &n_d21);
    std::vector<uint64_t> words((size_t)std::max<uint64_t>(n_solid, 1) * W);
    std::vector<uint8_t> adjb(std::max<uint64_t>(n_solid, 1));
    rc = p3_dbg_export(ctx, words.data(), adjb.data(), std::max<uint64_t>(n_solid, 1), &got);
    if (rc) return rc;
    StrAdjTable table;
    table.m.reserve((size_t)(got * 1.3) + 16);
    std::vector<std::string> frontier;
    frontier.reserve(got);
    for (uint64_t i = 0; i < got; i++) {
        frontier.push_back(ops.from_words(words.data() + i * W));
        table.m.emplace(frontier.back(), adjb[i]);
    }
    std::vector<uint64_t>().swap(words);
    // walk roots that are not solid (a seed with a non-ACGT base) get an entry of their own
    {
        std::vector<std::string> roots;
        for (const std::string &s : seeds) {
            std::string rcs = ops.revcomp(s);
            const std::string &c = s <= rcs ? s : rcs;
            uint8_t dummy;
            if (!table.find(c, &dummy)) roots.push_back(c);
        }
        std::sort(roots.begin(), roots.end());
        roots.erase(std::unique(roots.begin(), roots.end()), roots.end());
        if (!roots.empty()) {
            std::vector<uint64_t> rw(roots.size() * (size_t)W);
            for (size_t i = 0; i < roots.size(); i++) ops.to_words(roots[i], rw.data() + i * W, W);
            std::vector<uint8_t> masks(roots.size());
            rc = p3_check_directions(ctx, rw.data(), roots.size(), masks.data());
            if (rc) return rc;
            for (size_t i = 0; i < roots.size(); i++) { table.m.emplace(roots[i], masks[i]); frontier.push_back(roots[i]); }
        }
    }
    rc = close_table_long(ctx, ops, W, table, std::move(frontier));
    if (rc) return rc;

    log.line("start graph extention");
    WalkerStr w((int)k, table);
    if (w.make_dbg(seeds, /*node_limit*/ 4 * (table.m.size() + 16)) != 0 || w.missing) {
        g_host_err = w.missing ? "internal: walk left the closed adjacency table" : "walk did not terminate";
        return -P3_ERR_STATE;
    }
    log.line("de bruijn graph loaded");
    count_node_coverage_host(w, rd->seq, rd->off);
    log.line("count node coverage");
    if (gfa_path && print_graph(w, gfa_path) != P3_OK) { g_host_err = "cannot write GFA"; return -P3_ERR_IO; }
    if (stats) {
        stats[0] = n_reads; stats[1] = rd->all_bases; stats[2] = n_d21; stats[3] = n_solid; stats[4] = table.m.size();
        stats[5] = w.junctions.size(); stats[6] = w.joints.size(); stats[7] = w.straights.size();
    }
    return 0;
}

template <class W, class T, class KV>
int walk_and_print(int k, const T &table, const KV &seeds, uint64_t n_table, const p3_reads *rd, const char *gfa_path, uint64_t *stats) {
    const bool timing = getenv("P3_LOAD_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    W w(k, table);
    if (w.make_dbg(seeds, 4 * (n_table + 16)) != 0) { g_host_err = "walk did not terminate"; return P3_ERR_STATE; }
    if (w.missing) { g_host_err = "the adjacency table is not closed: the walk asked for a k-mer it lacks"; return P3_ERR_STATE; }
    auto t1 = std::chrono::steady_clock::now();
    count_node_coverage_host(w, rd->seq, rd->off);
    auto t2 = std::chrono::steady_clock::now();
    if (gfa_path && print_graph(w, gfa_path) != P3_OK) { g_host_err = "cannot write GFA"; return P3_ERR_IO; }
    if (timing) {
        auto t3 = std::chrono::steady_clock::now();
        auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
        fprintf(stderr, "p3_walk_table: walk %.3f s, node coverage (host) %.3f s, GFA %.3f s\n", sec(t0, t1), sec(t1, t2), sec(t2, t3));
    }
    if (stats) { stats[0] = w.junctions.size(); stats[1] = w.joints.size(); stats[2] = w.straights.size(); }
    return P3_OK;
}

}  // namespace

extern "C" {

// Host half of the drop-in on its own: Load + MakeDBG (-t 1 order) + CountNodeCoverage + PrintGraph
// over a CLOSED CheckDirections table from any source (n canonical k-mers of W = ceil(2k/64)
// little-endian words each, one adjacency byte each; closed = every reported neighbour is present)
// and the ORIENTED seed k-mers. Needs no GPU: node coverage is counted on the host here.
// stats (optional, 3 values): junctions, joints, straights.
int p3_walk_table(const char *read_path, uint32_t k, const uint64_t *h_kmers, const uint8_t *h_adj, uint64_t n,
                  const uint64_t *h_seeds, uint64_t n_seeds, const char *gfa_path, uint64_t *stats) {
    if (!read_path || (!h_kmers && n) || (!h_adj && n) || (!h_seeds && n_seeds)) { g_host_err = "p3_walk_table: null argument"; return P3_ERR_ARG; }
    if (k < P3_MIN_K || k > P3_MAX_K) { g_host_err = "k outside [21,3001] is not supported"; return P3_ERR_ARG; }
    p3_reads *rd = nullptr;
    int rc = p3_load_file(read_path, k, &rd);
    if (rc) return rc;
    if (k <= P3_MAX_K_WALK) {
        std::vector<uint64_t> kk(h_kmers, h_kmers + n), seeds(h_seeds, h_seeds + n_seeds);
        std::vector<uint8_t> aa(h_adj, h_adj + n);
        AdjTable table; table.build(kk, aa);
        std::sort(seeds.begin(), seeds.end());
        seeds.erase(std::unique(seeds.begin(), seeds.end()), seeds.end());
        rc = walk_and_print<Walker64>((int)k, table, seeds, n, rd, gfa_path, stats);
    } else {
        const StrOps ops((int)k);
        const int W = (int)((2 * k + 63) / 64);
        StrAdjTable table;
        for (uint64_t i = 0; i < n; i++) table.m.emplace(ops.from_words(h_kmers + i * W), h_adj[i]);
        std::vector<std::string> seeds;
        for (uint64_t i = 0; i < n_seeds; i++) seeds.push_back(ops.from_words(h_seeds + i * W));
        std::sort(seeds.begin(), seeds.end());
        seeds.erase(std::unique(seeds.begin(), seeds.end()), seeds.end());
        rc = walk_and_print<WalkerStr>((int)k, table, seeds, n, rd, gfa_path, stats);
    }
    p3_reads_free(rd);
    return rc;
}

}  // extern "C"

extern "C" {

// main.cpp:11-31 + Assemble<>, reference src/Assemble.cpp:7-28, for one read file.
// m = 0 estimates the filter (Options.cpp:50); threads is accepted for command-line compatibility
// (the walk follows the reference's -t 1 order). stats (optional, 8 values): reads, all_bases,
// distinct 21-mers, solid k-mers, table k-mers after closure, junctions, joints, straights.
int p3_assemble_file(const char *read_path, uint32_t k, uint64_t m, int threads, int device,
                     const char *gfa_path, const char *log_path, uint64_t *stats) {
    (void)threads;
    Log log; log.path = log_path ? log_path : "";
    if (k < P3_MIN_K || k > P3_MAX_K) { g_host_err = "k outside [21,3001] is not supported"; return P3_ERR_ARG; }
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
    if (k > P3_MAX_K_WALK) {   // multi-word k-mers: string walk, closure driven from the host (assemble_long)
        rc = assemble_long(ctx, rd, k, seed_pos, log, gfa_path, stats);
        if (rc > 0) { p3_destroy(ctx); p3_reads_free(rd); return -rc; }   // host-side failure, message already set
        if (rc) return bail(rc);
        p3_destroy(ctx);
        p3_reads_free(rd);
        log.line("finish");
        return P3_OK;
    }
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
    Walker64 w((int)k, table);
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
