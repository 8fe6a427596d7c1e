// p3_long.inc.cu — multi-word k-mers (33 <= k <= 3001, W = ceil(2k/64) words), part of the
// p3_gpu.cu translation unit. The reference instantiates std::bitset<2k> for k up to 3001
// (src/Assemble.cpp:30-53); stage A (21-mers) is unchanged, this file is stages B and C.
//
// Nothing here keeps a k-mer in registers. Word j of the forward or reverse-complement k-mer of
// an occurrence is cut out of the 2-bit stream on demand (two packed words + funnel shift), so:
//   canonical choice = compare words from the top until they differ (usually the first pair)
//   std::hash        = one pass over the canonical words from word 0 up (libstdc++ _Hash_bytes)
//   de-duplication   = set of [hash tag:24 | position:40]; a tag hit is verified by comparing the
//                      two occurrences' canonical words from the stream (exact)
// The distinct k-mers are then materialised as explicit n*W word arrays; BF.add and CheckDirections
// run on those arrays (which is also what p3_check_directions / p3_bf_add take from the host).

constexpr uint64_t kPos40 = (1ULL << 40) - 1;

struct LongK {
    int k, W, nbytes;
    uint64_t topmask;   // valid bits of word W-1
};
static LongK make_longk(uint32_t k) {
    LongK L; L.k = (int)k; L.W = (int)((2 * k + 63) / 64); L.nbytes = (int)((2 * k + 7) / 8);
    int r = (int)((2 * k) & 63);
    L.topmask = r ? ((1ULL << r) - 1) : ~0ULL;
    return L;
}

struct Stream2 { const uint64_t *packed; const uint32_t *nmask; };

__device__ __forceinline__ uint64_t stream_window(const uint64_t *__restrict__ packed, uint64_t s) {
    uint64_t w = s >> 5; int o = (int)(s & 31);
    return window(__ldg(packed + w), __ldg(packed + w + 1), o);
}
__device__ __forceinline__ uint64_t stream_mask2(const uint32_t *__restrict__ nmask, uint64_t s) {
    uint64_t w = s >> 5; int o = (int)(s & 31);
    uint64_t m = (((uint64_t)__ldg(nmask + w) << 32) | __ldg(nmask + w + 1)) << o;
    return spread32((uint32_t)(m >> 32));
}
// word j (little endian) of the forward k-mer starting at stream position p
__device__ __forceinline__ uint64_t fwd_word(const Stream2 &st, uint64_t p, int k, int j) {
    int lo = k - 32 * (j + 1);
    if (lo >= 0) return stream_window(st.packed, p + lo);
    int r = k - 32 * j;   // 1..31 bases in the top word
    return stream_window(st.packed, p) >> (64 - 2 * r);
}
// word j of the rolling "backward" k-mer (GetFirstKmerBackward, reference src/BitCalc.cpp:22-33:
// reverse complement, non-ACGT bases as code 0)
__device__ __forceinline__ uint64_t rc_word(const Stream2 &st, uint64_t p, int k, int j) {
    uint64_t s = p + 32 * (uint64_t)j;
    int r = k - 32 * j; if (r > 32) r = 32;
    uint64_t x = stream_window(st.packed, s);
    uint64_t m2 = st.nmask ? stream_mask2(st.nmask, s) : 0;
    uint64_t v = rev2(~x & ~m2);
    return r < 32 ? (v & ((1ULL << (2 * r)) - 1)) : v;
}
// CompareBit(Fw, Bw): true when the backward k-mer is the smaller one (ties keep Fw)
__device__ __forceinline__ bool occ_use_rc(const Stream2 &st, uint64_t p, const LongK &L) {
    for (int j = L.W - 1; j >= 0; j--) {
        uint64_t f = fwd_word(st, p, L.k, j), r = rc_word(st, p, L.k, j);
        if (f != r) return r < f;
    }
    return false;
}
__device__ __forceinline__ uint64_t occ_word(const Stream2 &st, uint64_t p, const LongK &L, int j, bool rc) {
    return rc ? rc_word(st, p, L.k, j) : fwd_word(st, p, L.k, j);
}
// libstdc++ _Hash_bytes over ceil(2k/8) bytes of little-endian words produced by get(j)
template <typename F>
__device__ __forceinline__ uint64_t hash_words(const LongK &L, F get) {
    const uint64_t mul = 0xc6a4a7935bd1e995ULL;
    uint64_t hash = 0xc70f6907ULL ^ ((uint64_t)L.nbytes * mul);
    const int nfull = L.nbytes >> 3;
    for (int j = 0; j < nfull; j++) {
        uint64_t data = shift_mix(get(j) * mul) * mul;
        hash ^= data; hash *= mul;
    }
    if (L.nbytes & 7) {   // the tail bytes are the whole (zero-extended) last word
        hash ^= get(nfull); hash *= mul;
    }
    hash = shift_mix(hash) * mul;
    hash = shift_mix(hash);
    return hash;
}

// solid plane for any window length x = k-20 (reference RMQ >= 2, src/MakeBloomFilter.cpp:62,75):
// position p is solid iff the run of set coverage bits starting at p is at least x long.
__global__ void __launch_bounds__(256)
solid_long_kernel(const uint32_t *__restrict__ good21, uint64_t n_words, int x, uint32_t *__restrict__ solid, Stats *st) {
    unsigned n = 0;
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    const uint64_t max_ahead = (uint64_t)(x + 31) / 32 + 1;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += stride) {
        uint32_t g = __ldg(good21 + w);
        uint32_t s = 0;
        if (g) {
            // first zero at or after the start of word w+1 (capped: beyond x+32 it cannot matter)
            uint64_t nz = (w + 1) * 32 + max_ahead * 32;
            for (uint64_t a = 1; a <= max_ahead; a++) {
                uint64_t ww = w + a;
                uint32_t v = ww <= n_words ? __ldg(good21 + ww) : 0u;   // good21[n_words] is zero
                if (v != 0xFFFFFFFFu) { nz = ww * 32 + __clz(~v); break; }
            }
            for (int o = 31; o >= 0; o--) {
                uint64_t p = w * 32 + o;
                if (!(g & (0x80000000u >> o))) nz = p;
                else if (nz - p >= (uint64_t)x) s |= 0x80000000u >> o;
            }
        }
        solid[w] = s;
        n += __popc(s);
    }
    unsigned long long a = warp_sum(n);
    if ((threadIdx.x & 31) == 0 && a) atomicAdd(&st->n_adds, a);
}

// canonical k-mers of occurrences p and q are equal
__device__ __forceinline__ bool occ_equal(const Stream2 &st, const LongK &L, uint64_t p, bool rcp, uint64_t q) {
    bool rcq = occ_use_rc(st, q, L);
    for (int j = 0; j < L.W; j++)
        if (occ_word(st, p, L, j, rcp) != occ_word(st, q, L, j, rcq)) return false;
    return true;
}

__global__ void __launch_bounds__(256)
dedupe_long_kernel(Stream2 st, const uint32_t *__restrict__ solid, uint64_t n_words, LongK L,
                   uint64_t *set, uint64_t nbs, Stats *stt) {
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    bool full = false;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += stride) {
        uint32_t s = __ldg(solid + w);
        while (s) {
            int o = __clz(s);
            s &= ~(0x80000000u >> o);
            const uint64_t p = w * 32 + o;
            const bool rcp = occ_use_rc(st, p, L);
            const uint64_t h0 = hash_words(L, [&](int j) { return occ_word(st, p, L, j, rcp); });
            const uint64_t tag = h0 >> 40;
            const uint64_t val = (tag << 40) | p;
            uint64_t b = __umul64hi(fmix64(h0), nbs);
            bool done = false;
            for (uint64_t probe = 0; probe < nbs && probe < kMaxProbe && !done; probe++) {   // a full table fails fast
                uint64_t *bp = set + 4 * b;
                uint64_t sl[4];
                ld_bucket(bp, sl);
#pragma unroll
                for (int i = 0; i < 4 && !done; i++) {
                    uint64_t v = sl[i];
                    if (v == kEmpty) {
                        v = atomicCAS(ull(bp + i), kEmpty, val);
                        if (v == kEmpty) { done = true; break; }
                    }
                    if ((v >> 40) == tag && occ_equal(st, L, p, rcp, v & kPos40)) done = true;
                }
                b = (b + 1 == nbs) ? 0 : b + 1;
            }
            if (!done) full = true;
        }
    }
    if (full) atomicExch(&stt->err_table_full, 1u);
}

// list of set slots -> explicit canonical words, n*W
__global__ void materialise_kernel(Stream2 st, LongK L, const uint64_t *__restrict__ list, uint64_t n, uint64_t *__restrict__ words) {
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t p = list[i] & kPos40;
        bool rc = occ_use_rc(st, p, L);
        for (int j = 0; j < L.W; j++) words[i * L.W + j] = occ_word(st, p, L, j, rc);
    }
}

// ---- explicit word arrays ------------------------------------------------------------------------------
__device__ __forceinline__ void bloom_bits_of_hash(const Bloom &bf, uint64_t h0, uint64_t &h1, uint64_t &h2) { double_hash(h0, h1, h2); }

__global__ void __launch_bounds__(256)
bloom_words_kernel(const uint64_t *__restrict__ words, uint64_t n, LongK L, Bloom bf, uint64_t seg_lo, uint64_t seg_hi) {
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t *kw = words + i * L.W;
        uint64_t h1, h2;
        double_hash(hash_words(L, [&](int j) { return __ldg(kw + j); }), h1, h2);
        uint64_t x = h1;
        for (int q = 0; q < bf.nh; q++, x += h2) {
            uint64_t bit = fastmod(x, bf.fm);
            if (bit < seg_lo || bit >= seg_hi) continue;
            atomicOr(bf.bits + (bit >> 5), 1u << (bit & 31));
        }
    }
}
__global__ void bf_query_words_kernel(const uint64_t *__restrict__ words, uint64_t n, LongK L, Bloom bf, uint8_t *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t *kw = words + i * L.W;
    uint64_t h1, h2;
    double_hash(hash_words(L, [&](int j) { return __ldg(kw + j); }), h1, h2);
    bool ok = true;
    uint64_t x = h1;
    for (int q = 0; q < bf.nh && ok; q++, x += h2) {
        uint64_t bit = fastmod(x, bf.fm);
        ok = (__ldg(bf.bits + (bit >> 5)) >> (bit & 31)) & 1u;
    }
    out[i] = ok ? 1 : 0;
}
__global__ void double_hash_words_kernel(const uint64_t *__restrict__ words, uint64_t n, LongK L, uint64_t *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t *kw = words + i * L.W;
    uint64_t h1, h2;
    double_hash(hash_words(L, [&](int j) { return __ldg(kw + j); }), h1, h2);
    out[2 * i] = h1; out[2 * i + 1] = h2;
}

// CheckDirections (reference src/DeBruijnGraph.cpp:326-345) on explicit ORIENTED k-mers of W words:
// 8 lanes per k-mer, one direction each. K and its true reverse complement (GetComplementKmer) are
// read through word accessors; the neighbour N and rc(N) are one-base shifts of those.
__global__ void __launch_bounds__(256)
adjacency_words_kernel(const uint64_t *__restrict__ words, uint64_t n, LongK L, Bloom bf, uint8_t *__restrict__ adj, Stats *st) {
    const int lane = threadIdx.x & 31;
    const int d = lane & 7, g = lane >> 3;
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t n_warps = (gridDim.x * (uint64_t)blockDim.x) >> 5;
    unsigned edges = 0;
    const int k = L.k, W = L.W;
    for (uint64_t base = warp * 4; base < n; base += n_warps * 4) {
        uint64_t i = base + g;
        bool rec = false;
        if (i < n) {
            const uint64_t *kw = words + i * W;
            auto K = [&](int j) -> uint64_t { return (j >= 0 && j < W) ? __ldg(kw + j) : 0ULL; };
            // true reverse complement of K, word j: bases [32j, 32j+32) of K counted from its first base
            auto R = [&](int j) -> uint64_t {
                if (j < 0 || j >= W) return 0ULL;
                int r = k - 32 * j; if (r > 32) r = 32;
                uint64_t x;
                if (r == 32) {
                    int b0 = 2 * (k - 32 * j - 32);   // lowest bit of the field
                    int wi = b0 >> 6, sh = b0 & 63;
                    x = sh ? ((K(wi) >> sh) | (K(wi + 1) << (64 - sh))) : K(wi);
                } else {
                    x = K(0) << (64 - 2 * r);
                }
                uint64_t v = rev2(~x);
                return r < 32 ? (v & ((1ULL << (2 * r)) - 1)) : v;
            };
            const int tb = 2 * k - 2, twi = tb >> 6, tsh = tb & 63;   // position of the first base
            const bool left = d < 4;
            const uint64_t dn = (uint64_t)(left ? d : d - 4), dc = 3 - dn;
            // N = left ? (K >> 2) | dn << (2k-2) : ((K << 2) | dn) & mask ;  rcN = the mirrored shift of R with dc
            auto shr2 = [&](auto &X, int j, uint64_t top) -> uint64_t {   // ((X >> 2) | top << (2k-2)) word j
                uint64_t v = (X(j) >> 2) | (X(j + 1) << 62);
                if (j == twi) v |= top << tsh;
                return v;
            };
            auto shl2 = [&](auto &X, int j, uint64_t low) -> uint64_t {   // (((X << 2) | low) & mask) word j
                uint64_t v = (X(j) << 2) | (j > 0 ? (X(j - 1) >> 62) : low);
                if (j == W - 1) v &= L.topmask;
                return v;
            };
            auto N = [&](int j) -> uint64_t { return left ? shr2(K, j, dn) : shl2(K, j, dn); };
            auto RN = [&](int j) -> uint64_t { return left ? shl2(R, j, dc) : shr2(R, j, dc); };
            bool use_rc = false;
            for (int j = W - 1; j >= 0; j--) {
                uint64_t a = N(j), b = RN(j);
                if (a != b) { use_rc = b < a; break; }
            }
            uint64_t h1, h2;
            double_hash(hash_words(L, [&](int j) { return use_rc ? RN(j) : N(j); }), h1, h2);
            rec = true;
            uint64_t x = h1;
            for (int q = 0; q < bf.nh && rec; q++, x += h2) {
                uint64_t bit = fastmod(x, bf.fm);
                rec = (__ldg(bf.bits + (bit >> 5)) >> (bit & 31)) & 1u;
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, rec);
        if (d == 0 && i < n) {
            unsigned byte = (m >> (8 * g)) & 0xFFu;
            adj[i] = (uint8_t)byte;
            edges += __popc(byte);
        }
    }
    unsigned long long e = warp_sum(edges);
    if (lane == 0 && e && st) atomicAdd(&st->n_edges, e);
}

// ---- host side of the long-k path --------------------------------------------------------------------------
struct LongState { uint64_t *d_words = nullptr; uint64_t cap_words = 0; };
static CtxStates<LongState> g_long;
static void long_release(p3_ctx *c) {
    LongState *ls = g_long.find(c);
    if (!ls) return;
    dfree(ls->d_words);
    g_long.erase(c);
}

static int bloom_add_words_direct(p3_ctx *c, const uint64_t *d_words, uint64_t n) {
    LongK L = make_longk(c->k);
    uint64_t seg_bits = 40ull << 23;
    uint64_t n_seg = (c->filter_size + seg_bits - 1) / seg_bits;
    if (n_seg > 16 || n * c->num_hashes < (1u << 22)) { n_seg = 1; seg_bits = c->filter_size; }
    else seg_bits = ((c->filter_size + n_seg - 1) / n_seg + 31) / 32 * 32;
    for (uint64_t sg = 0; sg < n_seg && n; sg++) {
        bloom_words_kernel<<<c->grid(), 256, 0, c->stream>>>(d_words, n, L, c->bloom(), sg * seg_bits, std::min<uint64_t>((sg + 1) * seg_bits, c->filter_size));
        c->launches++;
    }
    CU(cudaGetLastError());
    return P3_OK;
}
static int bloom_add_direct_long(p3_ctx *c, uint64_t nd) { return bloom_add_words_direct(c, g_long.get(c).d_words, nd); }
static int bloom_add_words(p3_ctx *c, const uint64_t *d_words, uint64_t n) {
    if (d_words == g_long.get(c).d_words) {   // the context's own k-mer list: binned adds (p3_bloom.inc.cu)
        bool done = false;
        int rcb = bloom_add_binned(c, n, &done);
        if (rcb || done) return rcb;
    }
    return bloom_add_words_direct(c, d_words, n);
}

// stage B for k > 32; the coverage plane (good21) is ready, planes are allocated, filter allocated
static int make_bf_long(p3_ctx *c, uint32_t k, uint64_t solid_slots, uint64_t est_distinct) {
    if (c->total_bases >= kPos40) return fail(P3_ERR_ARG, "k > 32 supports up to 2^40 bases per context");
    LongK L = make_longk(k);
    Stream2 st; st.packed = c->d_packed; st.nmask = c->d_nmask;
    solid_long_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_good21, c->n_words, (int)k - kShortK + 1, c->d_solid, c->d_stats);
    c->launches++;
    int rc = pull_stats(c);
    if (rc) return rc;
    // distinct solid k-mers: never more than the solid positions, normally about the distinct 21-mers that passed the
    // coverage test (est_distinct); an underestimate fails fast (bounded probes) and the set grows
    if (solid_slots == 0) solid_slots = std::max<uint64_t>(2 * std::min<uint64_t>(c->h_stats.n_adds, est_distinct), 1024);
    CU(cudaEventRecord(c->ev[4], c->stream));
    for (int attempt = 0;; attempt++) {
        uint64_t nbs = (solid_slots + 3) / 4;
        if (!c->d_set || c->nbs != nbs) {
            dfree(c->d_set); dfree(c->d_list);
            if (cudaMalloc(&c->d_set, nbs * 32) != cudaSuccess || cudaMalloc(&c->d_list, nbs * 32) != cudaSuccess) {
                cudaGetLastError();
                return fail(P3_ERR_NOMEM, "solid k-mer set allocation failed");
            }
            c->nbs = nbs; c->list_cap = nbs * 4;
        }
        c->set_parts = 1;
        CU(cudaMemsetAsync(c->d_set, 0xFF, nbs * 32, c->stream));
        CU(cudaMemsetAsync(&c->d_stats->n_distinct_solid, 0, sizeof(unsigned long long), c->stream));
        CU(cudaMemsetAsync(&c->d_stats->err_table_full, 0, sizeof(unsigned), c->stream));
        dedupe_long_kernel<<<c->grid(), 256, 0, c->stream>>>(st, c->d_solid, c->n_words, L, c->d_set, nbs, c->d_stats);
        compact_set_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_set, nbs * 4, c->d_list, c->list_cap, c->d_stats);
        c->launches += 2;
        CU(cudaGetLastError());
        rc = pull_stats(c);
        if (rc) return rc;
        if (!c->h_stats.err_table_full) break;
        if (attempt >= 16) return fail(P3_ERR_TABLE_FULL, "solid k-mer set full after growing");
        solid_slots = std::max<uint64_t>(solid_slots * 4, 1024);
    }
    uint64_t nd = c->h_stats.n_distinct_solid;
    LongState &ls = g_long.get(c);
    CU(ensure(ls.d_words, ls.cap_words, sizeof(uint64_t) * std::max<uint64_t>(nd * L.W, 1)));
    if (nd) {
        materialise_kernel<<<c->grid(), 256, 0, c->stream>>>(st, L, c->d_list, nd, ls.d_words);
        c->launches++;
    }
    CU(cudaMemsetAsync(c->d_bloom, 0, sizeof(uint32_t) * c->bloom_words, c->stream));
    CU(cudaEventRecord(c->ev[14], c->stream));
    rc = bloom_add_words(c, ls.d_words, nd);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[15], c->stream));
    CU(cudaEventRecord(c->ev[5], c->stream));
    seeds_kernel<<<c->grid(4), 256, 0, c->stream>>>(c->d_off, c->n_reads, c->d_solid, (int)k, c->d_seed);
    c->launches++;
    CU(cudaEventRecord(c->ev[6], c->stream));
    rc = pull_stats(c);
    if (rc) return rc;
    if (c->h_stats.err_bin_overflow) {   // a filter segment outgrew its bin: add directly (OR is idempotent)
        CU(cudaMemsetAsync(&c->d_stats->err_bin_overflow, 0, sizeof(unsigned), c->stream));
        rc = bloom_add_words_direct(c, ls.d_words, nd);
        if (!rc) rc = pull_stats(c);
        if (rc) return rc;
    }
    return P3_OK;
}

static int adjacency_long(p3_ctx *c, const uint64_t *d_words, uint64_t n, uint8_t *d_adj, Stats *st) {
    if (!n) return P3_OK;
    LongK L = make_longk(c->k);
    uint64_t warps = (n + 3) / 4;
    unsigned blocks = (unsigned)std::min<uint64_t>((warps + 7) / 8, (uint64_t)c->grid());
    adjacency_words_kernel<<<blocks, 256, 0, c->stream>>>(d_words, n, L, c->bloom(), d_adj, st);
    c->launches++;
    CU(cudaGetLastError());
    return P3_OK;
}

static const uint64_t *long_words(p3_ctx *c) { return g_long.get(c).d_words; }

// host-array batch entry points for W-word k-mers: 0 = BF.add, 1 = possiblyContains,
// 2 = GetDoubleHash_64bit, 3 = CheckDirections
static int long_batch(p3_ctx *c, int op, uint32_t k, const uint64_t *h_kmers, uint64_t n, void *h_out) {
    CU(cudaSetDevice(c->device));
    LongK L = make_longk(k);
    uint64_t *dk = nullptr; void *dout = nullptr;
    TmpFree tmp; tmp.own(&dk); tmp.own(&dout);
    CU(cudaMalloc(&dk, sizeof(uint64_t) * n * L.W));
    CU(cudaMemcpyAsync(dk, h_kmers, sizeof(uint64_t) * n * L.W, cudaMemcpyHostToDevice, c->stream));
    const unsigned nblk = (unsigned)((n + 255) / 256);
    size_t out_bytes = 0;
    if (op == 0) {
        bloom_words_kernel<<<c->grid(), 256, 0, c->stream>>>(dk, n, L, c->bloom(), 0, c->filter_size);
    } else if (op == 1) {
        out_bytes = n; CU(cudaMalloc(&dout, out_bytes));
        bf_query_words_kernel<<<nblk, 256, 0, c->stream>>>(dk, n, L, c->bloom(), (uint8_t *)dout);
    } else if (op == 2) {
        out_bytes = sizeof(uint64_t) * 2 * n; CU(cudaMalloc(&dout, out_bytes));
        double_hash_words_kernel<<<nblk, 256, 0, c->stream>>>(dk, n, L, (uint64_t *)dout);
    } else {
        out_bytes = n; CU(cudaMalloc(&dout, out_bytes));
        int rc = adjacency_long(c, dk, n, (uint8_t *)dout, nullptr);
        if (rc) return rc;
    }
    c->launches++;
    CU(cudaGetLastError());
    if (out_bytes) CU(cudaMemcpyAsync(h_out, dout, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}
