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
                 uint32_t *const *__restrict__ seg_base, unsigned long long *cursor, uint64_t cap) {
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
            if (dst < cap) seg_base[sg][dst] = s_rec[i];   // overflow is detected on the host from the cursors
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

struct BloomBinState {
    uint64_t *d_hh = nullptr; uint64_t cap_hh = 0;
    uint32_t *d_bins = nullptr; uint64_t cap_bins = 0;     // PEER-SHARED (p3_mg_bloom_buffer): never reallocated here
    uint32_t *d_local = nullptr; uint64_t cap_local = 0;   // local scratch of the single-context binned paths
    uint32_t **d_segbase = nullptr; uint64_t cap_segbase = 0;
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
    if (!c->d_ghist) {
        CU(cudaMalloc(&c->d_ghist, sizeof(unsigned long long) * (kMaxParts + 1)));
        CU(cudaMalloc(&c->d_cursor, sizeof(unsigned long long) * (kMaxParts + 1)));
    }
    CU(ensure(b.d_segbase, b.cap_segbase, sizeof(uint32_t *) * kMaxParts));
    CU(cudaMemcpyAsync(b.d_segbase, h_segbase, sizeof(uint64_t) * n_seg, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_cursor, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    const size_t smem = bloom_bin_smem(c->num_hashes, n_seg);
    CU(cudaFuncSetAttribute(bloom_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n) {
        const uint64_t tile_k = (uint64_t)kBinThreads * kBinKpt;
        unsigned blocks = (unsigned)std::min<uint64_t>((n + tile_k - 1) / tile_k, (uint64_t)c->n_sm * 2);
        FastMod fm = make_fastmod(c->filter_size);
        uint64_t wrap = ((~0ULL % c->filter_size) + 1) % c->filter_size;
        bloom_bin_kernel<<<blocks, kBinThreads, smem, c->stream>>>(d_hh, n, fm, wrap, (int)c->num_hashes, shift, n_seg,
                                                                   b.d_segbase, c->d_cursor, cap);
        c->launches++;
        CU(cudaGetLastError());
    }
    std::vector<unsigned long long> h(n_seg);
    CU(cudaMemcpyAsync(h.data(), c->d_cursor, sizeof(unsigned long long) * n_seg, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < n_seg; i++) h_counts[i] = h[i];
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
    // 5 % + 64 K slack
    const double share = std::min(1.0, (double)(1ull << shift) / (double)c->filter_size);
    const uint64_t cap = (uint64_t)((double)n * c->num_hashes * share * 1.05) + 65536;
    const uint64_t need = sizeof(uint32_t) * cap * n_seg;
    BloomBinState &b = g_bbin.get(c);
    uint32_t *bins = nullptr;
    if (c->d_bkeys && c->cap_bkeys >= need) bins = reinterpret_cast<uint32_t *>(c->d_bkeys);   // count-stage bins are idle now
    else { CU(ensure(b.d_local, b.cap_local, need)); bins = b.d_local; }
    std::vector<uint64_t> base(n_seg), counts(n_seg);
    for (uint64_t s = 0; s < n_seg; s++) base[s] = (uint64_t)(uintptr_t)(bins + s * cap);
    rc = bloom_bin_launch(c, d_hh, n, (uint32_t)n_seg, shift, base.data(), cap, counts.data());
    if (rc) return rc;
    for (uint64_t s = 0; s < n_seg; s++) if (counts[s] > cap) return P3_OK;   // direct path instead
    for (uint64_t s = 0; s < n_seg; s++) {
        ApplyRegions rg; rg.count = 1; rg.ptr[0] = bins + s * cap; rg.n[0] = counts[s];
        rc = bloom_apply_launch(c, s, shift, rg);
        if (rc) return rc;
    }
    *done = true;
    return P3_OK;
}

// ---- binned coverage-bit clears --------------------------------------------------------------------
// MakeBF's coverage test (reference src/MakeBloomFilter.cpp:52-58) clears one bit per count-1 key:
// ~1.05 G single-bit RED.ANDs at random positions of a 0.54 GB plane at configs[1], 21.8 G/s from DRAM.
// Same cure as for BF.add: the positions are tile-sorted by plane segment (2^27 positions = 16 MB) and
// applied segment by segment with the window L2 resident. A segment can never receive more records
// than it has positions, so the bins have a hard upper bound and need neither a histogram nor an
// overflow path.
constexpr int kPosKpt = 8;   // inputs per thread per tile
// MODE 0: candidates (slot, position) of the binned count — emitted when the key's final count < thr
// MODE 1: a plain list of position records
template <int MODE>
__global__ void __launch_bounds__(kBinThreads)
pos_bin_kernel(const uint64_t *__restrict__ slots, const uint64_t *__restrict__ cand_slot, const uint64_t *__restrict__ pos_in,
               uint64_t n, uint64_t thr, Ovf ovf, const Stats *st, int shift, uint32_t P, uint32_t *__restrict__ bins,
               uint64_t cap, unsigned long long *cursor) {
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
    const int tid = threadIdx.x;
    const unsigned n_overflow = MODE == 0 ? st->n_overflow : 0u;
    const uint64_t pmask = (1ULL << kPosRankShift) - 1, smask = (1ULL << shift) - 1;
    const uint64_t n_tiles = (n + T - 1) / T;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t i = tid; i < P; i += kBinThreads) s_hist[i] = 0;
        __syncthreads();
        uint64_t pos[kPosKpt];
#pragma unroll
        for (int q = 0; q < kPosKpt; q++) {
            const uint64_t i = tile * T + (uint64_t)q * kBinThreads + tid;
            pos[q] = ~0ULL;
            if (i < n) {
                if (MODE == 0) {
                    uint64_t v = __ldcg(slots + __ldcs(cand_slot + i));
                    uint64_t c = v >> 42;
                    if (c < thr && n_overflow) c += ovf_get(ovf, v & kKey42) << 22;
                    if (c < thr) pos[q] = __ldcs(pos_in + i) & pmask;
                } else {
                    pos[q] = __ldcs(pos_in + i) & pmask;
                }
            }
            if (pos[q] != ~0ULL) atomicAdd(&s_hist[(uint32_t)(pos[q] >> shift)], 1u);
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
        }
        __syncthreads();
    }
}
// RED.AND of one segment's records into its window of a plane whose bit for position p is
// 0x80000000 >> (p & 31) of word p >> 5 (the layout of the valid / coverage / solid planes)
__global__ void __launch_bounds__(256)
plane_clear_kernel(const uint32_t *__restrict__ rec, uint64_t n, uint32_t *__restrict__ seg_words) {
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t off = __ldcs(rec + i);
        atomicAnd(seg_words + (off >> 5), ~(0x80000000u >> (off & 31)));
    }
}

// clears, in `plane` (n_pos positions), the bit of every emitted position; *done = false when the job is
// too small to be worth it or the scratch does not fit (the caller then clears directly)
template <int MODE>
static int binned_plane_clear(p3_ctx *c, const uint64_t *cand_slot, const uint64_t *pos_in, uint64_t n, uint64_t thr,
                              uint32_t *plane, uint64_t n_pos, bool *done) {
    *done = false;
    const int shift = bloom_seg_shift();
    const uint64_t n_seg = (n_pos + (1ull << shift) - 1) >> shift;
    if (n < (1u << 22) || n_seg < 2 || n_seg > (uint64_t)kMaxParts || getenv("P3_DIRECT_CLEARS")) {
        if (!(getenv("P3_BINNED_CLEARS") && n && n_seg <= (uint64_t)kMaxParts)) return P3_OK;
    }
    const uint64_t cap = std::min<uint64_t>(1ull << shift, n);      // a segment has 2^shift positions: a hard bound
    const uint64_t need = sizeof(uint32_t) * cap * n_seg;
    BloomBinState &b = g_bbin.get(c);
    uint32_t *bins = nullptr;
    if (c->d_bkeys && c->cap_bkeys >= need) bins = reinterpret_cast<uint32_t *>(c->d_bkeys);   // count-stage bins are idle now
    else {
        size_t fr = 0, tot = 0;
        CU(cudaMemGetInfo(&fr, &tot));
        if (b.cap_local < need && need > (uint64_t)(0.5 * (double)fr)) return P3_OK;
        CU(ensure(b.d_local, b.cap_local, need));
        bins = b.d_local;
    }
    if (!c->d_ghist) {
        CU(cudaMalloc(&c->d_ghist, sizeof(unsigned long long) * (kMaxParts + 1)));
        CU(cudaMalloc(&c->d_cursor, sizeof(unsigned long long) * (kMaxParts + 1)));
    }
    CU(cudaMemsetAsync(c->d_cursor, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    const size_t smem = (size_t)kBinThreads * kPosKpt * 6 + (size_t)n_seg * 20;
    CU(cudaFuncSetAttribute(pos_bin_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t T = (uint64_t)kBinThreads * kPosKpt;
    unsigned blocks = (unsigned)std::min<uint64_t>((n + T - 1) / T, (uint64_t)c->n_sm * 8);
    pos_bin_kernel<MODE><<<blocks, kBinThreads, smem, c->stream>>>(c->d_table, cand_slot, pos_in, n, thr, c->ovf(), c->d_stats, shift,
                                                                   (uint32_t)n_seg, bins, cap, c->d_cursor);
    c->launches++;
    CU(cudaGetLastError());
    std::vector<unsigned long long> h(n_seg);
    CU(cudaMemcpyAsync(h.data(), c->d_cursor, sizeof(unsigned long long) * n_seg, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (uint64_t s = 0; s < n_seg; s++) {
        if (h[s] > cap) return fail(P3_ERR_STATE, "binned_plane_clear: more records than positions in a segment (duplicate positions?)");
        if (!h[s]) continue;
        unsigned ab = (unsigned)std::min<uint64_t>((h[s] + 255) / 256, (uint64_t)c->grid());
        plane_clear_kernel<<<ab, 256, 0, c->stream>>>(bins + s * cap, h[s], plane + (s << (shift - 5)));
        c->launches++;
    }
    CU(cudaGetLastError());
    *done = true;
    return P3_OK;
}
