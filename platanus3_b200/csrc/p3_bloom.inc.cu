// p3_bloom.inc.cu — binned BF.add (reference src/bloomfilter.cpp:69-74, called from
// src/MakeBloomFilter.cpp:75-77), part of the p3_gpu.cu translation unit.
//
// BF.add of n distinct k-mers is n * num_hashes single-bit ORs at hashed positions of a filter that
// is far larger than L2 (0.3 GB at configs[1], 2.3 GB at 8 ranks, 5.3 GB at human scale). Done
// directly that is one DRAM-random RED per bit (21.8 G/s, profiles/r01_randacc.md). Instead:
//
//   bloom_bin_kernel    every k-mer's num_hashes bit indices are computed ONCE (incremental
//                       (h1 + n*h2) mod filter_size: one add/compare per index instead of a 64-bit
//                       multiply-high), tile-sorted in shared memory by filter SEGMENT (2^27 bits =
//                       16 MB) and written as coalesced runs of 4-byte in-segment offsets
//   bloom_apply_kernel  one launch per segment: RED.OR of its records; the 16 MB window stays in L2
//
// Multi-GPU: the segments are dealt out to the ranks in contiguous shards; the bin kernel stores a
// segment's runs straight into the shard owner's buffer over NVLink peer memory (the same fused
// bin + exchange as the 21-mer count), every rank applies its own shard and the shards are
// all-gathered — no replicated adds and no OR-reduce of whole filter copies.

constexpr int kBinThreads = 256;
constexpr int kBinKpt = 2;                 // k-mers per thread per tile
constexpr int kBinMaxHashes = 32;          // more hash functions than this: direct path

static int bloom_seg_shift() {
    int sh = 27;
    if (const char *e = getenv("P3_BLOOM_SEG_BITS")) {
        unsigned long long b = strtoull(e, nullptr, 10);
        int s = 0;
        while ((1ull << (s + 1)) <= b && s < 40) s++;
        sh = s;
    }
    return std::min(std::max(sh, 10), 31);
}

// hh = (h1, h2) pairs of GetDoubleHash_64bit(std::hash(kmer)); seg_base[s] = where THIS source's
// records of segment s go (may be peer memory); cursor[s] = records written so far (starts at 0)
__global__ void __launch_bounds__(kBinThreads)
bloom_bin_kernel(const uint64_t *__restrict__ hh, uint64_t n, FastMod fm, uint64_t wrap, int nh, int shift, uint32_t P,
                 uint32_t *const *__restrict__ seg_base, unsigned long long *cursor, uint64_t cap, Stats *st) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t T = kBinThreads * kBinKpt * (uint32_t)nh;
    uint32_t *s_rec = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned long long *s_gbase = reinterpret_cast<unsigned long long *>(s_rec + T + (T & 1));
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_gbase + P);
    uint32_t *s_offs = s_hist + P;
    uint32_t *s_cur = s_offs + P;
    uint16_t *s_seg = reinterpret_cast<uint16_t *>(s_cur + P);
    __shared__ uint32_t s_wtot[kBinThreads / 32];
    __shared__ uint32_t s_total;
    const int tid = threadIdx.x;
    const uint32_t mask = (1u << shift) - 1u;
    const uint64_t tile_k = (uint64_t)kBinThreads * kBinKpt;
    const uint64_t n_tiles = (n + tile_k - 1) / tile_k;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t i = tid; i < P; i += kBinThreads) s_hist[i] = 0;
        __syncthreads();
        uint64_t h1[kBinKpt], h2[kBinKpt];
        bool live[kBinKpt];
#pragma unroll
        for (int kk = 0; kk < kBinKpt; kk++) {
            uint64_t i = tile * tile_k + (uint64_t)kk * kBinThreads + tid;
            live[kk] = i < n;
            h1[kk] = live[kk] ? __ldcs(hh + 2 * i) : 0;
            h2[kk] = live[kk] ? __ldcs(hh + 2 * i + 1) : 0;
            if (live[kk]) {
                BloomStep b; b.init(h1[kk], h2[kk], fm, wrap);
                for (int q = 0; q < nh; q++, b.next()) atomicAdd(&s_hist[(uint32_t)(b.r >> shift)], 1u);
            }
        }
        __syncthreads();
        {   // exclusive scan of the histogram, one global claim per non-empty segment
            const uint32_t per = (P + kBinThreads - 1) / kBinThreads;
            const uint32_t b0 = tid * per;
            uint32_t local = 0;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < P) local += s_hist[b0 + j];
            uint32_t incl = local;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if ((tid & 31) >= d) incl += v; }
            if ((tid & 31) == 31) s_wtot[tid >> 5] = incl;
            __syncthreads();
            uint32_t wbase = 0;
            for (int q = 0; q < (tid >> 5); q++) wbase += s_wtot[q];
            uint32_t run = wbase + incl - local;
            for (uint32_t j = 0; j < per; j++) {
                uint32_t i = b0 + j;
                if (i < P) {
                    uint32_t h = s_hist[i];
                    s_offs[i] = run; s_cur[i] = run;
                    if (h) s_gbase[i] = atomicAdd(&cursor[i], (unsigned long long)h);
                    run += h;
                }
            }
            if (tid == kBinThreads - 1) s_total = wbase + incl;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kBinKpt; kk++) {
            if (live[kk]) {
                BloomStep b; b.init(h1[kk], h2[kk], fm, wrap);
                for (int q = 0; q < nh; q++, b.next()) {
                    uint32_t sg = (uint32_t)(b.r >> shift);
                    uint32_t idx = atomicAdd(&s_cur[sg], 1u);
                    s_rec[idx] = (uint32_t)b.r & mask;
                    s_seg[idx] = (uint16_t)sg;
                }
            }
        }
        __syncthreads();
        const uint32_t total = s_total;
        for (uint32_t i = tid; i < total; i += kBinThreads) {
            uint32_t sg = s_seg[i];
            unsigned long long dst = s_gbase[sg] + (i - s_offs[sg]);
            if (dst < cap) seg_base[sg][dst] = s_rec[i];   // overflow: flagged here, and visible in the cursors
            else atomicExch(&st->err_bin_overflow, 1u);
        }
        __syncthreads();
    }
}

constexpr int kMaxRegions = 16;
struct ApplyRegions { const uint32_t *ptr[kMaxRegions]; unsigned long long n[kMaxRegions]; int count; };
// RED.OR of one segment's records (in-segment bit offsets) into its 2^shift-bit window of the filter
__global__ void __launch_bounds__(256)
bloom_apply_kernel(ApplyRegions rg, uint32_t *__restrict__ seg_words) {
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    const uint64_t t0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (int g = 0; g < rg.count; g++) {
        const uint32_t *p = rg.ptr[g];
        const uint64_t n = rg.n[g];
        for (uint64_t i = t0; i < n; i += stride) {
            uint32_t off = __ldcs(p + i);
            atomicOr(seg_words + (off >> 5), 1u << (off & 31));
        }
    }
}

// All segments in ONE launch: the [n_seg x cap] bin space is handed out in chunks from a global
// counter (like insert_bins), so the chunks in flight lie in one or two neighbouring segments and
// their 16 MB windows stay in L2; the per-segment record counts are read from the device cursors
// (no host round trip between binning and applying). CLEAR = false: RED.OR of filter bits
// (bit i = word i>>5 bit i&31); CLEAR = true: RED.AND of plane bits (0x80000000 >> (p & 31)).
template <bool CLEAR>
__global__ void __launch_bounds__(256)
apply_bins_kernel(const uint32_t *__restrict__ bins, uint64_t cap, const unsigned long long *__restrict__ count, uint32_t n_seg,
                  int shift, uint32_t *__restrict__ target, Stats *st) {
    __shared__ unsigned long long s_base;
    const uint64_t n = (uint64_t)n_seg * cap;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_base = atomicAdd(&st->work, (unsigned long long)kSweepChunk);
        __syncthreads();
        const uint64_t cbase = s_base;
        if (cbase >= n) break;
        const uint64_t sg = cbase / cap;
        const uint64_t lim = sg * cap + min((uint64_t)__ldg(count + sg), cap);
        if (cbase >= lim) continue;
        uint32_t *seg_words = target + (sg << (shift - 5));
#pragma unroll
        for (int it = 0; it < kSweepPer; it++) {
            const uint64_t i = cbase + it * 256 + threadIdx.x;
            if (i < lim) {
                const uint32_t off = __ldcs(bins + i);
                if (CLEAR) atomicAnd(seg_words + (off >> 5), ~(0x80000000u >> (off & 31)));
                else atomicOr(seg_words + (off >> 5), 1u << (off & 31));
            }
        }
    }
}

struct BloomBinState {
    uint64_t *d_hh = nullptr; uint64_t cap_hh = 0;
    uint32_t *d_bins = nullptr; uint64_t cap_bins = 0;     // PEER-SHARED (p3_mg_bloom_buffer): never reallocated here
    uint32_t *d_local = nullptr; uint64_t cap_local = 0;   // local scratch of the single-context binned paths
    uint32_t **d_segbase = nullptr; uint64_t cap_segbase = 0;
    size_t smem_set = 0, smem_pos[4] = {0, 0, 0, 0};   // dynamic shared memory opted into so far (per context = per device)
    std::vector<void *> graveyard;   // outgrown peer-shared buffers: freed with the context, never while peers may map them
};
static CtxStates<BloomBinState> g_bbin;
static void bloom_release(p3_ctx *c) {
    BloomBinState *b = g_bbin.find(c);
    if (!b) return;
    dfree(b->d_hh); dfree(b->d_bins); dfree(b->d_local); dfree(b->d_segbase);
    for (void *p : b->graveyard) cudaFree(p);
    g_bbin.erase(c);
}

static size_t bloom_bin_smem(uint32_t nh, uint32_t P) {
    size_t T = (size_t)kBinThreads * kBinKpt * nh;
    return (T + (T & 1)) * 4 + (size_t)P * (8 + 4 + 4 + 4) + T * 2;
}

// (h1, h2) of the context's current distinct k-mer list (single word or W-word arrays)
static int bloom_hash_list(p3_ctx *c, uint64_t n, uint64_t **d_hh) {
    BloomBinState &b = g_bbin.get(c);
    CU(ensure(b.d_hh, b.cap_hh, sizeof(uint64_t) * 2 * std::max<uint64_t>(n, 1)));
    if (n) {
        unsigned blocks = (unsigned)((n + 255) / 256);
        if (c->k > 32) double_hash_words_kernel<<<blocks, 256, 0, c->stream>>>(long_words(c), n, make_longk(c->k), b.d_hh);
        else double_hash_kernel<<<blocks, 256, 0, c->stream>>>((int)((2 * c->k + 7) / 8), c->d_list, n, b.d_hh);
        c->launches++;
        CU(cudaGetLastError());
    }
    *d_hh = b.d_hh;
    return P3_OK;
}

// bins the n k-mers' bit indices into n_seg destinations (h_segbase: device addresses, cap records
// each); h_counts receives the records written per segment (> cap = overflow, nothing lost yet:
// the caller falls back to the direct path)
static int bloom_bin_launch(p3_ctx *c, const uint64_t *d_hh, uint64_t n, uint32_t n_seg, int shift,
                            const uint64_t *h_segbase, uint64_t cap, uint64_t *h_counts) {
    BloomBinState &b = g_bbin.get(c);
    int rc0 = hist_buffers(c);
    if (rc0) return rc0;
    CU(ensure(b.d_segbase, b.cap_segbase, sizeof(uint32_t *) * kMaxParts));
    CU(cudaMemcpyAsync(b.d_segbase, h_segbase, sizeof(uint64_t) * n_seg, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_cursor, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    const size_t smem = bloom_bin_smem(c->num_hashes, n_seg);
    if (smem > b.smem_set) {
        CU(cudaFuncSetAttribute(bloom_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        b.smem_set = smem;
    }
    if (n) {
        const uint64_t tile_k = (uint64_t)kBinThreads * kBinKpt;
        unsigned blocks = (unsigned)std::min<uint64_t>((n + tile_k - 1) / tile_k, (uint64_t)c->n_sm * 2);
        FastMod fm = make_fastmod(c->filter_size);
        uint64_t wrap = ((~0ULL % c->filter_size) + 1) % c->filter_size;
        bloom_bin_kernel<<<blocks, kBinThreads, smem, c->stream>>>(d_hh, n, fm, wrap, (int)c->num_hashes, shift, n_seg,
                                                                   b.d_segbase, c->d_cursor, cap, c->d_stats);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (h_counts) {   // multi-GPU: the caller routes the per-segment counts to the shard owners
        std::vector<unsigned long long> h(n_seg);
        CU(cudaMemcpyAsync(h.data(), c->d_cursor, sizeof(unsigned long long) * n_seg, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (uint32_t i = 0; i < n_seg; i++) h_counts[i] = h[i];
    }
    return P3_OK;
}

static int bloom_apply_launch(p3_ctx *c, uint64_t seg, int shift, const ApplyRegions &rg) {
    unsigned long long tot = 0;
    for (int g = 0; g < rg.count; g++) tot += rg.n[g];
    if (!tot) return P3_OK;
    unsigned blocks = (unsigned)std::min<uint64_t>((tot + 255) / 256, (uint64_t)c->grid());
    bloom_apply_kernel<<<blocks, 256, 0, c->stream>>>(rg, c->d_bloom + (seg << (shift - 5)));
    c->launches++;
    CU(cudaGetLastError());
    return P3_OK;
}

static bool bloom_binned_wanted(p3_ctx *c, uint64_t n, uint32_t n_seg) {
    if (c->num_hashes == 0 || c->num_hashes > kBinMaxHashes || n_seg > kMaxParts) return false;
    if (const char *e = getenv("P3_BLOOM_BINNED")) return atoi(e) != 0;
    return n_seg >= 2 && n * c->num_hashes >= (1u << 22);
}

// single GPU: dense BF.add of the first n k-mers of the current list. *done = false: not applicable
// (tiny job, one segment, overflow) — the caller runs the direct path, which is always correct
// because OR is idempotent.
static int bloom_add_binned(p3_ctx *c, uint64_t n, bool *done) {
    *done = false;
    const int shift = bloom_seg_shift();
    const uint64_t n_seg = (c->filter_size + (1ull << shift) - 1) >> shift;
    if (!n || !bloom_binned_wanted(c, n, (uint32_t)std::min<uint64_t>(n_seg, 1u << 30))) return P3_OK;
    uint64_t *d_hh = nullptr;
    int rc = bloom_hash_list(c, n, &d_hh);
    if (rc) return rc;
    // hashed indices are uniform over the filter: a FULL segment receives n * num_hashes * seg_bits /
    // filter_size of them (the last segment is usually partial, so this is more than 1 / n_seg);
    // 5 % + 64 K slack, rounded to the sweep chunk
    const double share = std::min(1.0, (double)(1ull << shift) / (double)c->filter_size);
    uint64_t cap = (uint64_t)((double)n * c->num_hashes * share * 1.05) + 65536;
    cap = (cap + kSweepChunk - 1) / kSweepChunk * kSweepChunk;
    const uint64_t need = sizeof(uint32_t) * cap * n_seg;
    BloomBinState &b = g_bbin.get(c);
    uint32_t *bins = nullptr;
    if (b.d_local && b.cap_local >= need) bins = b.d_local;
    else if (c->d_bkeys && c->cap_bkeys >= need) { bins = reinterpret_cast<uint32_t *>(c->d_bkeys); c->bins_valid = false; }   // count-stage bins are idle now
    else { CU(ensure(b.d_local, b.cap_local, need)); bins = b.d_local; }
    std::vector<uint64_t> base(n_seg);
    for (uint64_t s = 0; s < n_seg; s++) base[s] = (uint64_t)(uintptr_t)(bins + s * cap);
    rc = bloom_bin_launch(c, d_hh, n, (uint32_t)n_seg, shift, base.data(), cap, nullptr);
    if (rc) return rc;
    // an overflowing segment (impossible for uniform hashes within the slack) raises err_bin_overflow; the
    // records that fit are applied anyway (OR is idempotent) and p3_make_bf re-adds directly when it sees the flag
    CU(cudaMemsetAsync(&c->d_stats->work, 0, sizeof(unsigned long long), c->stream));
    apply_bins_kernel<false><<<c->grid(), 256, 0, c->stream>>>(bins, cap, c->d_cursor, (uint32_t)n_seg, shift, c->d_bloom, c->d_stats);
    c->launches++;
    CU(cudaGetLastError());
    *done = true;
    return P3_OK;
}

// ---- binned coverage-bit clears --------------------------------------------------------------------
// MakeBF's coverage test (reference src/MakeBloomFilter.cpp:52-58) clears one bit per occurrence of a key
// whose count stayed below the threshold: ~0.75 G single-bit RED.ANDs at random positions of a 0.54 GB
// plane at configs[1], 21.8 G/s from DRAM. Same cure as for BF.add: the positions are tile-sorted by plane
// segment (2^27 positions = 16 MB) and applied segment by segment with the window L2 resident. The bins
// have a fixed capacity (the expected share of a segment + 10 % + slack: error k-mers are spread evenly
// over the reads); a segment that overflows raises err_bin_overflow and the caller clears directly.
constexpr int kPosKpt = 8;   // inputs per thread per tile

// KC[key] < thr, the first bucket of the probe sequence already loaded (see count_insert_pre)
__device__ __forceinline__ bool count_below_pre(const Table &t, uint64_t key, uint64_t base, uint64_t b, uint64_t s[4],
                                                uint64_t thr, const Ovf &ovf, unsigned n_overflow) {
    for (uint64_t probe = 0; probe < t.nbp; probe++) {
        if (probe) ld_bucket(t.slots + 4 * (base + b), s);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint64_t v = s[i];
            if ((v & kKey42) == key) {
                uint64_t c = v >> 42;
                if (c < thr && n_overflow) c += ovf_get(ovf, key) << 22;
                return c < thr;
            }
            if (v == kEmpty) return true;   // absent: count 0
        }
        b = (b + 1 == t.nbp) ? 0 : b + 1;
    }
    return true;
}

// "count < thr" of every table slot as ONE BIT (streaming pass over the table, 32 slots per thread): the verdict sweep then
// asks a 0.75 MB window per 48 MB partition instead of the slots themselves — always an L2 (mostly L1) hit
__global__ void __launch_bounds__(256)
below_bits_kernel(const uint64_t *__restrict__ slots, uint64_t n_slots, uint64_t thr, Ovf ovf, const Stats *st, uint32_t *__restrict__ bits) {
    const unsigned n_overflow = st->n_overflow;
    const int lane = threadIdx.x & 31;
    const uint64_t n_words = (n_slots + 31) / 32;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * (uint64_t)blockDim.x) >> 5;
    // a warp turns 32 x 32 consecutive slots into 32 words: lane l reads slot 32 j + l of word j (coalesced), the ballot is
    // word j, lane j keeps it
    for (uint64_t w0 = warp * 32; w0 < n_words; w0 += n_warps * 32) {
        uint32_t mine = 0;
#pragma unroll 4
        for (int j = 0; j < 32; j++) {
            const uint64_t i = (w0 + j) * 32 + lane;
            bool below = false;
            if (i < n_slots) {
                const uint64_t v = __ldcs(slots + i);
                if (v != kEmpty) {
                    uint64_t cnt = v >> 42;
                    if (cnt < thr && n_overflow) cnt += ovf_get(ovf, v & kKey42) << 22;
                    below = cnt < thr;
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, below);
            if (lane == j) mine = m;
        }
        if (w0 + lane < n_words) bits[w0 + lane] = mine;
    }
}

// MODE 0: the verdict sweep — the input is the partition bins of the count (records + word indices, laid
//         out as in_cap-sized bins ending at in_end[p], or contiguous when in_cap == 0); every record whose
//         key's final count < thr emits its position. Tiles are handed out from a global counter so that
//         the grid stays inside one or two table partitions (L2 resident), like insert_bins.
//         out_list != nullptr: no tile sort; the position records (with their rank byte) are appended to
//         out_list instead (multi-GPU owner: they go back to their source ranks).
// MODE 1: a plain list of position records, or (in_cap > 0) a row of regions of in_cap records holding
//         in_end[r] records each (received from the owner ranks)
template <int MODE, bool IDX>
__global__ void __launch_bounds__(kBinThreads, IDX ? 6 : 2)
pos_bin_kernel(Table table, const uint64_t *__restrict__ in, const uint32_t *__restrict__ in_word, const uint32_t *__restrict__ in_idx,
               const uint32_t *__restrict__ below,
               uint64_t n, uint64_t in_cap, const unsigned long long *__restrict__ in_end, const unsigned long long *__restrict__ n_dev,
               uint64_t thr, Ovf ovf, Stats *st, int shift, uint32_t P, uint32_t *__restrict__ bins,
               uint64_t cap, unsigned long long *cursor, uint64_t *__restrict__ out_list) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t T = kBinThreads * kPosKpt;
    uint32_t *s_rec = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned long long *s_gbase = reinterpret_cast<unsigned long long *>(s_rec + T);
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_gbase + P);
    uint32_t *s_offs = s_hist + P;
    uint32_t *s_cur = s_offs + P;
    uint16_t *s_seg = reinterpret_cast<uint16_t *>(s_cur + P);
    __shared__ uint32_t s_wtot[kBinThreads / 32];
    __shared__ uint32_t s_total;
    __shared__ unsigned long long s_tile, s_lbase;
    const int tid = threadIdx.x;
    const unsigned n_overflow = MODE == 0 ? st->n_overflow : 0u;
    const uint64_t pmask = (1ULL << kPosRankShift) - 1, smask = (1ULL << shift) - 1;
    if (n_dev) n = min(n, (uint64_t)*n_dev);
    bool over = false;
    uint64_t t0 = (uint64_t)blockIdx.x * T;
    for (;;) {
        if (MODE == 0) {   // dynamic hand-out (L2 residency of the table partition)
            __syncthreads();
            if (tid == 0) s_tile = atomicAdd(&st->work, (unsigned long long)T);
            __syncthreads();
            t0 = s_tile;
        }
        if (t0 >= n) break;
        uint64_t lim = n;
        if (in_cap) { const uint64_t r = t0 / in_cap; lim = min(n, r * in_cap + min((uint64_t)__ldcg(in_end + r) - (MODE == 0 ? r * in_cap : 0), in_cap)); }
        if (t0 < lim) {
            for (uint32_t i = tid; i < P; i += kBinThreads) s_hist[i] = 0;
            __syncthreads();
            uint64_t pos[kPosKpt];
            if (MODE == 0 && IDX) {
                // the insert left the partition-relative slot of every record: its count is one 8-byte load away. When the
                // index also carries the record's offset and rank (table.sb), the position needs the word index only.
                const uint64_t pbase = in_cap ? 4 * (t0 / in_cap) * table.nbp : 0;
                const uint32_t smask = table.sb ? (1u << table.sb) - 1 : 0xFFFFFFFFu;
#pragma unroll
                for (int half = 0; half < 2; half++) {     // four records at a time keeps the kernel at 6 blocks per SM
                    constexpr int H = kPosKpt / 2;
                    uint32_t r[H], wd[H], v[H];
#pragma unroll
                    for (int q = 0; q < H; q++) {
                        const uint64_t i = t0 + (uint64_t)(half * H + q) * kBinThreads + tid;
                        r[q] = i < lim ? __ldcs(in_idx + i) : kNoSlot;
                        wd[q] = (table.sb && i < lim) ? __ldcs(in_word + i) : 0u;
                    }
#pragma unroll
                    for (int q = 0; q < H; q++) {
                        v[q] = 0;
                        if (r[q] != kNoSlot) {
                            uint64_t base = pbase;
                            if (!in_cap) base = 4 * (uint64_t)part_of(fmix64(__ldcs(in + t0 + (uint64_t)(half * H + q) * kBinThreads + tid) & kKey42), table.P) * table.nbp;
                            const uint64_t slot = base + (r[q] & smask);
                            v[q] = (__ldg(below + (slot >> 5)) >> (slot & 31)) & 1u;     // final count of the record's key < thr
                        }
                    }
#pragma unroll
                    for (int q = 0; q < H; q++) {
                        const int qq = half * H + q;
                        pos[qq] = ~0ULL;
                        if (v[q]) {
                            if (table.sb) {
                                const uint32_t tag = r[q] >> table.sb;   // offset:5 | rank:4
                                pos[qq] = ((uint64_t)wd[q] * 32 + (tag & 31)) | ((uint64_t)(tag >> 5) << kPosRankShift);
                            } else {
                                const uint64_t i = t0 + (uint64_t)qq * kBinThreads + tid;
                                pos[qq] = posrec_of(__ldcs(in + i), __ldcs(in_word + i));
                            }
                        }
                    }
                }
            } else if (MODE == 0 && !IDX) {
                uint64_t rec[kPosKpt / 2], s[kPosKpt / 2][4];
#pragma unroll
                for (int half = 0; half < 2; half++) {
#pragma unroll
                    for (int q = 0; q < kPosKpt / 2; q++) {
                        const uint64_t i = t0 + (uint64_t)(half * (kPosKpt / 2) + q) * kBinThreads + tid;
                        rec[q] = i < lim ? __ldcs(in + i) : ~0ULL;
                        if (rec[q] != ~0ULL) {
                            const uint64_t h = fmix64(rec[q] & kKey42);
                            ld_bucket(table.slots + 4 * ((uint64_t)part_of(h, table.P) * table.nbp + sub_of(h, table.nbp)), s[q]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < kPosKpt / 2; q++) {
                        const int qq = half * (kPosKpt / 2) + q;
                        const uint64_t i = t0 + (uint64_t)qq * kBinThreads + tid;
                        pos[qq] = ~0ULL;
                        if (rec[q] != ~0ULL) {
                            const uint64_t h = fmix64(rec[q] & kKey42);
                            if (count_below_pre(table, rec[q] & kKey42, (uint64_t)part_of(h, table.P) * table.nbp, sub_of(h, table.nbp), s[q], thr, ovf, n_overflow))
                                pos[qq] = posrec_of(rec[q], __ldcs(in_word + i));
                        }
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < kPosKpt; q++) {
                    const uint64_t i = t0 + (uint64_t)q * kBinThreads + tid;
                    pos[q] = i < lim ? (__ldcs(in + i) & pmask) : ~0ULL;
                }
            }
            if (MODE == 0 && out_list) {   // block-level compaction into the list (one global atomic per tile)
                unsigned mine = 0;
#pragma unroll
                for (int q = 0; q < kPosKpt; q++) mine += pos[q] != ~0ULL;
                unsigned incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if ((tid & 31) >= d) incl += v; }
                if ((tid & 31) == 31) s_wtot[tid >> 5] = incl;
                __syncthreads();
                unsigned wbase = 0, total = 0;
#pragma unroll
                for (int q = 0; q < kBinThreads / 32; q++) { unsigned t = s_wtot[q]; if (q < (tid >> 5)) wbase += t; total += t; }
                if (tid == 0 && total) s_lbase = atomicAdd(&st->n_export, (unsigned long long)total);
                __syncthreads();
                unsigned long long j = s_lbase + wbase + incl - mine;
#pragma unroll
                for (int q = 0; q < kPosKpt; q++) if (pos[q] != ~0ULL) { if (j < cap) out_list[j] = pos[q]; else over = true; j++; }
            } else {
#pragma unroll
                for (int q = 0; q < kPosKpt; q++) {
                    if (pos[q] != ~0ULL) { pos[q] &= pmask; atomicAdd(&s_hist[(uint32_t)(pos[q] >> shift)], 1u); }
                }
                __syncthreads();
                {
                    const uint32_t per = (P + kBinThreads - 1) / kBinThreads;
                    const uint32_t b0 = tid * per;
                    uint32_t local = 0;
                    for (uint32_t j = 0; j < per; j++) if (b0 + j < P) local += s_hist[b0 + j];
                    uint32_t incl = local;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if ((tid & 31) >= d) incl += v; }
                    if ((tid & 31) == 31) s_wtot[tid >> 5] = incl;
                    __syncthreads();
                    uint32_t wbase = 0;
                    for (int q = 0; q < (tid >> 5); q++) wbase += s_wtot[q];
                    uint32_t run = wbase + incl - local;
                    for (uint32_t j = 0; j < per; j++) {
                        uint32_t i = b0 + j;
                        if (i < P) {
                            uint32_t h = s_hist[i];
                            s_offs[i] = run; s_cur[i] = run;
                            if (h) s_gbase[i] = atomicAdd(&cursor[i], (unsigned long long)h);
                            run += h;
                        }
                    }
                    if (tid == kBinThreads - 1) s_total = wbase + incl;
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < kPosKpt; q++) {
                    if (pos[q] != ~0ULL) {
                        uint32_t sg = (uint32_t)(pos[q] >> shift);
                        uint32_t idx = atomicAdd(&s_cur[sg], 1u);
                        s_rec[idx] = (uint32_t)(pos[q] & smask);
                        s_seg[idx] = (uint16_t)sg;
                    }
                }
                __syncthreads();
                const uint32_t total = s_total;
                for (uint32_t i = tid; i < total; i += kBinThreads) {
                    uint32_t sg = s_seg[i];
                    unsigned long long dst = s_gbase[sg] + (i - s_offs[sg]);
                    if (dst < cap) bins[(uint64_t)sg * cap + dst] = s_rec[i];
                    else over = true;
                }
                __syncthreads();
            }
        }
        if (MODE != 0) { t0 += (uint64_t)gridDim.x * T; }
    }
    if (over) atomicExch(&st->err_bin_overflow, 1u);
}

// direct (un-binned) form of the same two jobs, for small inputs and as the fallback when a segment bin overflows
template <int MODE>
__global__ void __launch_bounds__(256)
pos_clear_direct_kernel(Table table, const uint64_t *__restrict__ in, const uint32_t *__restrict__ in_word, const uint32_t *__restrict__ in_idx, uint64_t n,
                        uint64_t in_cap, const unsigned long long *__restrict__ in_end, const unsigned long long *__restrict__ n_dev,
                        uint64_t thr, Ovf ovf, const Stats *st, uint32_t *plane) {
    const unsigned n_overflow = MODE == 0 ? st->n_overflow : 0u;
    if (n_dev) n = min(n, (uint64_t)*n_dev);
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        if (in_cap) {
            const uint64_t r = i / in_cap;
            if (i - r * in_cap >= min((uint64_t)__ldcg(in_end + r) - (MODE == 0 ? r * in_cap : 0), in_cap)) continue;
        }
        uint64_t pos;
        if (MODE == 0) {
            const uint64_t rec = __ldcs(in + i);
            if (in_idx) {
                const uint32_t r = __ldcs(in_idx + i);
                if (r == kNoSlot) continue;
                const uint64_t base = 4 * (in_cap ? (i / in_cap) : (uint64_t)part_of(fmix64(rec & kKey42), table.P)) * table.nbp;
                const uint64_t v = __ldcg(table.slots + base + (table.sb ? (r & ((1u << table.sb) - 1)) : r));
                uint64_t cnt = v >> 42;
                if (cnt < thr && n_overflow) cnt += ovf_get(ovf, v & kKey42) << 22;
                if (cnt >= thr) continue;
            } else if (count_lookup(table, rec & kKey42, ovf, n_overflow) >= thr) continue;
            pos = posrec_of(rec, __ldcs(in_word + i));
        } else pos = __ldcs(in + i);
        pos &= (1ULL << kPosRankShift) - 1;
        atomicAnd(plane + (pos >> 5), ~(0x80000000u >> (pos & 31)));
    }
}

// the input of a clear job (see pos_bin_kernel)
struct ClearInput {
    const uint64_t *rec = nullptr; const uint32_t *word = nullptr;    // MODE 0: count records + word indices; MODE 1: position records
    const uint32_t *idx = nullptr;                                    // MODE 0, optional: partition-relative slot of every record (insert_find)
    const uint32_t *below = nullptr;                                  // with idx: one bit per table slot, final count < thr (below_bits_kernel)
    uint64_t n = 0;                                                   // inputs (capacity space when in_cap > 0)
    uint64_t in_cap = 0; const unsigned long long *in_end = nullptr;  // binned / regioned layout
    const unsigned long long *n_dev = nullptr;                        // contiguous layout with the count on the device
};

// clears, in `plane` (n_pos positions), the bit of every emitted position; n_expect = how many positions the
// input is expected to emit at most (sizes the bins). Small jobs, unfitting scratch and P3_DIRECT_CLEARS clear
// directly. A segment that outgrows its bin raises Stats::err_bin_overflow: the caller checks it at its next
// stats read and calls again with force_direct (clearing twice is harmless).
template <int MODE>
static int plane_clear_job(p3_ctx *c, const ClearInput &in, uint64_t n_expect, uint64_t thr, uint32_t *plane, uint64_t n_pos,
                           bool force_direct, bool *binned) {
    if (binned) *binned = false;
    if (in.n == 0) return P3_OK;
    const int shift = bloom_seg_shift();
    const uint64_t n_seg = (n_pos + (1ull << shift) - 1) >> shift;
    bool want = !force_direct && !getenv("P3_DIRECT_CLEARS") && n_seg <= (uint64_t)kMaxParts &&
                ((n_expect >= (1u << 22) && n_seg >= 2) || getenv("P3_BINNED_CLEARS"));
    BloomBinState &b = g_bbin.get(c);
    uint64_t cap = 0;
    if (want) {
        // a full segment's expected share of the emitted positions + 10 % + 64 K, never more than it has positions
        const double share = std::min(1.0, (double)(1ull << shift) / (double)std::max<uint64_t>(n_pos, 1));
        cap = std::min<uint64_t>(1ull << shift, (uint64_t)((double)n_expect * share * 1.10) + 65536);
        cap = (cap + kSweepChunk - 1) / kSweepChunk * kSweepChunk;
        const uint64_t need = sizeof(uint32_t) * cap * n_seg;
        if (b.cap_local < need) {
            size_t fr = 0, tot = 0;
            CU(cudaMemGetInfo(&fr, &tot));
            if (need > (uint64_t)(0.5 * (double)(fr + b.cap_local))) want = false;
            else CU(ensure(b.d_local, b.cap_local, need));
        }
    }
    int rc0 = hist_buffers(c);
    if (rc0) return rc0;
    if (MODE == 0) CU(cudaMemsetAsync(&c->d_stats->work, 0, sizeof(unsigned long long), c->stream));
    if (!want) {
        pos_clear_direct_kernel<MODE><<<c->grid(), 256, 0, c->stream>>>(c->table(), in.rec, in.word, in.idx, in.n, in.in_cap, in.in_end, in.n_dev,
                                                                       thr, c->ovf(), c->d_stats, plane);
        c->launches++;
        CU(cudaGetLastError());
        return P3_OK;
    }
    CU(cudaMemsetAsync(c->d_cursor, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    const size_t smem = (size_t)kBinThreads * kPosKpt * 6 + (size_t)n_seg * 20;
    const bool idx = MODE == 0 && in.idx != nullptr && in.below != nullptr;
    if (smem > b.smem_pos[MODE + (idx ? 2 : 0)]) {
        if (idx) CU(cudaFuncSetAttribute(pos_bin_kernel<MODE, MODE == 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else CU(cudaFuncSetAttribute(pos_bin_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        b.smem_pos[MODE + (idx ? 2 : 0)] = smem;
    }
    const uint64_t T = (uint64_t)kBinThreads * kPosKpt;
    unsigned blocks = (unsigned)std::min<uint64_t>((in.n + T - 1) / T, (uint64_t)c->n_sm * (MODE == 0 ? 4 : 8));
    if (idx) pos_bin_kernel<MODE, MODE == 0><<<blocks, kBinThreads, smem, c->stream>>>(c->table(), in.rec, in.word, in.idx, in.below, in.n, in.in_cap, in.in_end, in.n_dev, thr, c->ovf(),
                                                                                       c->d_stats, shift, (uint32_t)n_seg, b.d_local, cap, c->d_cursor, nullptr);
    else pos_bin_kernel<MODE, false><<<blocks, kBinThreads, smem, c->stream>>>(c->table(), in.rec, in.word, in.idx, nullptr, in.n, in.in_cap, in.in_end, in.n_dev, thr, c->ovf(),
                                                                               c->d_stats, shift, (uint32_t)n_seg, b.d_local, cap, c->d_cursor, nullptr);
    CU(cudaMemsetAsync(&c->d_stats->work, 0, sizeof(unsigned long long), c->stream));
    apply_bins_kernel<true><<<c->grid(), 256, 0, c->stream>>>(b.d_local, cap, c->d_cursor, (uint32_t)n_seg, shift, plane, c->d_stats);
    c->launches += 2;
    CU(cudaGetLastError());
    if (binned) *binned = true;
    return P3_OK;
}

// MakeBF's coverage flags from the binned count (reference src/MakeBloomFilter.cpp:52-58, any threshold):
// good21 := valid, then every occurrence of a key whose final count < thr clears its bit. The bins of a
// one-chunk count are still there; otherwise the reads are binned again chunk by chunk.
// one bit per table slot: final count < thr (input of the verdict sweep over the insert's index stream)
static int below_bits(p3_ctx *c, uint64_t thr) {
    const uint64_t n_slots = c->nb * 4;
    CU(ensure(c->d_below, c->cap_below, sizeof(uint32_t) * ((n_slots + 31) / 32 + 1)));
    below_bits_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_table, n_slots, thr, c->ovf(), c->d_stats, c->d_below);
    c->launches++;
    CU(cudaGetLastError());
    return P3_OK;
}
static int verdict_sweep(p3_ctx *c, uint64_t thr, bool force_direct, bool *binned_any) {
    *binned_any = false;
    CU(cudaMemcpyAsync(c->d_good21, c->d_valid, sizeof(uint32_t) * c->n_words, cudaMemcpyDeviceToDevice, c->stream));
    if (!c->h_stats.n_cand) return P3_OK;
    // occurrences below the threshold: at most (thr - 1) per distinct key, and never more than the positions
    const uint64_t n_expect = std::min<uint64_t>(c->h_stats.n_pos21, c->h_stats.n_cand * std::max<uint64_t>(thr - 1, 1));
    auto one = [&](uint64_t share_num, uint64_t share_den) -> int {
        ClearInput in;
        in.rec = c->d_bkeys; in.word = c->d_bword; in.n = c->bin_n; in.in_cap = c->bin_cap; in.in_end = c->d_binmeta;
        in.idx = c->bins_valid ? c->d_bidx : nullptr;   // re-binned chunks have no index stream: bucket probes instead
        in.below = c->bins_valid ? c->d_below : nullptr;
        in.n_dev = c->bin_cap ? nullptr : c->d_binmeta + kMaxParts;
        bool b = false;
        int rc = plane_clear_job<0>(c, in, n_expect / share_den * share_num + (1u << 16), thr, c->d_good21, c->n_words * 32, force_direct, &b);
        *binned_any = *binned_any || b;
        return rc;
    };
    if (c->bins_valid) {
        int rcb = below_bits(c, thr);
        if (rcb) return rcb;
        return one(1, 1);
    }
    BinPlan pl;
    int rc = plan_bins(c, c->bin_upper, c->bin_exact, &pl);
    if (rc) return rc;
    const uint64_t n_ch = (c->n_words + pl.chunk_words - 1) / pl.chunk_words;
    for (uint64_t w0 = 0; w0 < c->n_words; w0 += pl.chunk_words) {
        const uint64_t w1 = std::min<uint64_t>(w0 + pl.chunk_words, c->n_words);
        rc = bin_chunk(c, pl, w0, w1, nullptr);
        if (!rc) rc = one(2, std::max<uint64_t>(n_ch, 1));   // a chunk's share of the clears, with a factor 2 of slack
        if (rc) return rc;
    }
    c->bins_valid = false;
    return P3_OK;
}
