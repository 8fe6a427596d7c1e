// p3_gpu.cu — kernels and the C ABI (include/platanus3_b200.h) of the B200-native
// k-mer-to-graph hot path: CountShortKmer -> MakeBF -> CheckDirections.
//
// Thread mapping for the three streaming kernels: one thread per packed word (32 consecutive
// k-mer start positions). A warp reads 32 consecutive words (256 B, coalesced) plus the next
// word as hand-off, so every k-mer is two registers and a funnel shift away; there is no
// rolling state to carry along a read and no divergence on read boundaries (a read-end bit
// plane turns "k-mer crosses a read end" into one AND per position).
#include "../../include/platanus3_b200.h"
#include "p3_device.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

using namespace p3;

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

// read-end plane: bit (31 - j%32) of rend[j/32] set iff stream position j is the last base of a
// read or lies in the zero padding. "k-mer at p is inside one read" == no end bit in [p, p+k-2].
__global__ void rend_kernel(const uint64_t *__restrict__ off, uint64_t n_reads, uint64_t total_bases,
                            uint64_t n_bits, uint32_t *__restrict__ rend) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t r = t; r < n_reads; r += stride) {
        uint64_t e = off[r + 1];
        if (e == 0) continue;
        uint64_t p = e - 1;
        atomicOr(rend + (p >> 5), 0x80000000u >> (p & 31));
    }
    for (uint64_t p = total_bases + t; p < n_bits; p += stride) atomicOr(rend + (p >> 5), 0x80000000u >> (p & 31));
}

// ---- direct count (P3_COUNT_MODE=direct): every position goes straight to the DRAM-resident table
template <bool HAS_MASK>
__global__ void __launch_bounds__(256)
count21_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ rend,
               const uint32_t *__restrict__ nmask, uint64_t n_words, Table table,
               Ovf ovf, Stats *st, uint32_t *__restrict__ proven2) {
    const uint64_t W21 = ~0ULL << (64 - (kShortK - 1));  // positions p .. p+19
    unsigned n_pos = 0, n_new = 0;
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += stride) {
        uint64_t hi = __ldg(packed + w), lo = __ldg(packed + w + 1);
        uint64_t E = ((uint64_t)__ldg(rend + w) << 32) | __ldg(rend + w + 1);
        uint64_t mhi = 0, mlo = 0;
        if (HAS_MASK) { mhi = spread32(__ldg(nmask + w)); mlo = spread32(__ldg(nmask + w + 1)); }
        uint32_t g = 0;
#pragma unroll 4
        for (int o = 0; o < 32; o++) {
            if (((E << o) & W21) != 0) continue;
            uint64_t x = window(hi, lo, o);
            uint64_t m2 = HAS_MASK ? window(mhi, mlo, o) : 0;
            uint64_t key = canonical_from_window(x, m2, kShortK);
            uint64_t created;
            uint64_t reached = count_insert(table, key, ovf, st, &created);
            n_pos++;
            n_new += (created != ~0ULL) ? 1u : 0u;
            if (reached >= 2) g |= 0x80000000u >> o;
        }
        proven2[w] = g;  // bit set: this 21-mer's final count is certainly >= 2
    }
    unsigned long long a = warp_sum(n_pos), b = warp_sum(n_new);
    if ((threadIdx.x & 31) == 0) {
        if (a) atomicAdd(&st->n_pos21, a);
        if (b) atomicAdd(&st->n_distinct21, b);
    }
}

// shortk_cov >= threshold per position (reference src/MakeBloomFilter.cpp:52-58), one bit each
template <bool HAS_MASK>
__global__ void __launch_bounds__(256)
flags21_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ rend,
               const uint32_t *__restrict__ nmask, uint64_t n_words, Table table,
               Ovf ovf, const Stats *st, uint64_t thr, const uint32_t *__restrict__ proven2,
               uint32_t *__restrict__ good21) {
    const uint64_t W21 = ~0ULL << (64 - (kShortK - 1));
    const unsigned n_overflow = st->n_overflow;
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += stride) {
        uint64_t hi = __ldg(packed + w), lo = __ldg(packed + w + 1);
        uint64_t E = ((uint64_t)__ldg(rend + w) << 32) | __ldg(rend + w + 1);
        uint64_t mhi = 0, mlo = 0;
        if (HAS_MASK) { mhi = spread32(__ldg(nmask + w)); mlo = spread32(__ldg(nmask + w + 1)); }
        uint32_t valid = 0;
#pragma unroll
        for (int o = 0; o < 32; o++) valid |= (((E << o) & W21) == 0) ? (0x80000000u >> o) : 0u;
        // positions count21 already proved (count >= 2) need no second table access
        uint32_t g = proven2 ? (__ldg(proven2 + w) & valid) : 0u;
        uint32_t need = valid & ~g;
        while (need) {
            int o = __clz(need);
            need &= ~(0x80000000u >> o);
            uint64_t x = window(hi, lo, o);
            uint64_t m2 = HAS_MASK ? window(mhi, mlo, o) : 0;
            uint64_t key = canonical_from_window(x, m2, kShortK);
            if (count_lookup(table, key, ovf, n_overflow) >= thr) g |= 0x80000000u >> o;
        }
        good21[w] = g;
    }
}

// ---- binned count (default): bin -> L2-resident insert sweep -> creator positions ------------------
//
// K1 hist21:    per-partition record counts of a chunk of words (smem histogram per block; only for
//               exact bins — the default is fixed-capacity bins, no histogram)
// K2 scan:      exclusive scan -> bin bases / cursors
// K3 scatter21: block-local counting sort of a 4096-position tile by partition, then coalesced
//               runs into the bins. Record = [o:5 | key:42] (uint64) + word index (uint32).
// K4 insert_bins: sweep over the binned records. Records are ordered by partition, so the whole grid
//               works inside one ~24 MB table partition at a time (L2 resident).
// K5 verdict sweep (p3_bloom.inc.cu pos_bin<0>): once every count is final, the SAME bins are swept a
//               second time (again partition by partition out of L2, loads only): a record whose key's
//               count stayed below the threshold clears its position's bit in the coverage plane
//               (reference src/MakeBloomFilter.cpp:52-58). Works for any threshold, needs no side table.
constexpr int kTileWords = 128;                 // words per scatter tile
constexpr int kTilePos = kTileWords * 32;       // 4096 positions

template <bool HAS_MASK, int PMODE>
__global__ void __launch_bounds__(256)
hist21_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ rend,
              const uint32_t *__restrict__ nmask, uint64_t w0, uint64_t w1, uint32_t P,
              unsigned long long *__restrict__ ghist) {
    __shared__ unsigned int sh[kMaxParts];
    const uint64_t W21 = ~0ULL << (64 - (kShortK - 1));
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t w = w0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < w1; w += stride) {
        uint64_t hi = __ldg(packed + w), lo = __ldg(packed + w + 1);
        uint64_t E = ((uint64_t)__ldg(rend + w) << 32) | __ldg(rend + w + 1);
        uint64_t mhi = 0, mlo = 0;
        if (HAS_MASK) { mhi = spread32(__ldg(nmask + w)); mlo = spread32(__ldg(nmask + w + 1)); }
#pragma unroll 4
        for (int o = 0; o < 32; o++) {
            if (((E << o) & W21) != 0) continue;
            uint64_t key = canonical_from_window(window(hi, lo, o), HAS_MASK ? window(mhi, mlo, o) : 0, kShortK);
            atomicAdd(&sh[pid_of<PMODE>(key, P)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
        if (sh[i]) atomicAdd(&ghist[i], (unsigned long long)sh[i]);
}

// one block: cursor[p] = exclusive prefix of ghist; ghist[P] receives the total
__global__ void scan_parts_kernel(const unsigned long long *__restrict__ ghist, uint32_t P,
                                  unsigned long long *__restrict__ cursor, unsigned long long *__restrict__ total) {
    __shared__ unsigned long long s[kMaxParts];
    for (uint32_t i = threadIdx.x; i < kMaxParts; i += blockDim.x) s[i] = i < P ? ghist[i] : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (uint32_t i = 0; i < P; i++) { unsigned long long v = s[i]; s[i] = acc; acc += v; }
        *total = acc;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) cursor[i] = s[i];
}

// fixed-capacity bins: cursor[p] = p * cap (no histogram pass)
__global__ void init_cursors_kernel(unsigned long long *cursor, uint32_t P, uint64_t cap) {
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) cursor[i] = (unsigned long long)i * cap;
}
// deferred overflow check of fixed-capacity bins (no host round trip per chunk): a partition whose
// cursor ran past its capacity dropped records; the flag makes the host redo the stage with exact bins
__global__ void check_cursors_kernel(const unsigned long long *__restrict__ cursor, uint32_t P, uint64_t cap, Stats *st) {
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
        if (cursor[i] - (unsigned long long)i * cap > cap) atomicExch(&st->err_bin_overflow, 1u);
}

// Shared memory of the tile sorts. The sorted records of the 21-mer sort carry their partition and
// their word-in-tile in the bits the key leaves free ([pt:10 @54 | loc:7 @47 | o:5 @42 | key:42]), so
// that no side arrays are needed: 44 KB per block = 5 blocks per SM (72 KB = 3 blocks before). The
// generic record sort (arbitrary 64-bit records) keeps a 16-bit partition array, and a 16-bit source
// index when an auxiliary word travels with the record.
template <bool PART, bool LOC, int AUXB = 0>
struct ScatterSmemT {
    uint64_t key[kTilePos];               // 32 KB  records sorted by partition
    uint32_t hist[kMaxParts];             //  4 KB  per-partition counts, then (in place) exclusive offsets
    unsigned long long gbase[kMaxParts];  //  8 KB
    uint16_t part[PART ? kTilePos : 1];   //  8 KB
    uint16_t loc[LOC ? kTilePos : 1];     //  8 KB
    uint32_t aux4[AUXB == 4 ? kTilePos : 1];   // 16 KB  the auxiliary word of every sorted record (generic record sort)
    uint8_t aux1[AUXB == 1 ? kTilePos : 1];    //  4 KB
    uint32_t warp_tot[16];
    uint32_t total;
};
using ScatterSmem21 = ScatterSmemT<false, false>;
constexpr int kSmPtShift = 54, kSmLocShift = 47;

// exclusive scan of sm.hist[0..P) in place, one global claim per non-empty partition into sm.gbase,
// total into sm.total. Called by all NT threads of the block between two __syncthreads().
template <int NT, class SM>
__device__ __forceinline__ void tile_scan_and_claim(SM &sm, uint32_t P, unsigned long long *cursor, int tid) {
    const uint32_t per_thread = (P + NT - 1) / NT;
    uint32_t local = 0;
    const uint32_t b0 = tid * per_thread;
    for (uint32_t j = 0; j < per_thread; j++) { uint32_t i = b0 + j; if (i < P) local += sm.hist[i]; }
    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if ((tid & 31) >= d) incl += v; }
    if ((tid & 31) == 31) sm.warp_tot[tid >> 5] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int q = 0; q < (tid >> 5); q++) wbase += sm.warp_tot[q];
    uint32_t run = wbase + incl - local;
    for (uint32_t j = 0; j < per_thread; j++) {
        uint32_t i = b0 + j;
        if (i < P) {
            uint32_t h = sm.hist[i];
            sm.hist[i] = run;
            if (h) sm.gbase[i] = atomicAdd(&cursor[i], (unsigned long long)h);
            run += h;
        }
    }
    if (tid == NT - 1) sm.total = wbase + incl;
}

// PEER mode (multi-GPU): partition p's records go to a DIFFERENT buffer per partition — the
// receive buffer of owner rank p, mapped into this process through NVLink peer memory (CUDA IPC,
// p3_ipc_open). The tile's runs are stored there directly, so the exchange happens inside the
// binning kernel, tile by tile, instead of as a separate all-to-all over a send buffer.
constexpr int kMaxPeers = 16;
struct PeerOut { uint64_t *keys[kMaxPeers]; uint32_t *words[kMaxPeers]; };

// 512 threads per 128-word tile: thread t handles 8 of the 32 offsets of word t>>2 (16 per thread
// needed ~70 registers = 3 blocks of 256 per SM; 8 per thread fit 40 = 3 blocks of 512).
// The reverse strand is taken from the bit-reversed, complemented word pair (computed once per
// thread): the reverse complement of the k-mer at offset o is a funnel shift of that pair, like the
// forward k-mer is a funnel shift of (hi, lo) — no per-position bit reversal.
constexpr int kScatterThreads = 4 * kTileWords;
constexpr int kScatterPer = kTilePos / kScatterThreads;   // 8
template <bool HAS_MASK, int PMODE, bool PEER = false, bool FILTER = false>
__global__ void __launch_bounds__(kScatterThreads, 3)
scatter21_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ rend,
                 const uint32_t *__restrict__ nmask, uint64_t w0, uint64_t w1, uint32_t P,
                 unsigned long long *cursor, uint64_t *__restrict__ bkeys, uint32_t *__restrict__ bword,
                 uint32_t *__restrict__ valid_plane, uint64_t tag, Stats *st, PeerOut peer = PeerOut(), uint64_t cap = 0,
                 uint32_t flt_lo = 0, uint32_t flt_hi = 0, uint32_t flt_P = 0) {
    // FILTER (multi-GPU key-range rounds; its own instantiation so that the unfiltered kernels keep their register budget): only the
    // keys whose TABLE partition (of flt_P) lies in [flt_lo, flt_hi) are binned in this pass; the valid plane is written for every
    // position as always
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmem21 &sm = *reinterpret_cast<ScatterSmem21 *>(smem_raw);
    const uint64_t W21 = ~0ULL << (64 - (kShortK - 1));
    const int tid = threadIdx.x;
    const int wt = tid >> 2;                    // word of the tile
    const int o0 = (tid & 3) * kScatterPer;     // first offset this thread handles
    const uint64_t n_tiles = (w1 - w0 + kTileWords - 1) / kTileWords;
    unsigned long long n_pos = 0;               // last thread only: positions binned by this block
    bool over = false;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t i = tid; i < P; i += kScatterThreads) sm.hist[i] = 0;
        __syncthreads();
        const uint64_t w = w0 + tile * kTileWords + wt;
        uint64_t hi = 0, lo = 0, rlo = 0, rhi = 0;
        uint32_t valid = 0;                     // bit (7 - q): offset o0 + q starts a 21-mer inside one read
        if (w < w1) {
            hi = __ldg(packed + w); lo = __ldg(packed + w + 1);
            const uint64_t E = (((uint64_t)__ldg(rend + w) << 32) | __ldg(rend + w + 1)) << o0;
            uint64_t chi = ~hi, clo = ~lo;
            if (HAS_MASK) { chi &= ~spread32(__ldg(nmask + w)); clo &= ~spread32(__ldg(nmask + w + 1)); }
            rlo = rev2(chi); rhi = rev2(clo);      // complemented bases in reverse order: base j at bits [2j+1, 2j]
#pragma unroll
            for (int q = 0; q < kScatterPer; q++) valid |= (((E << q) & W21) == 0) ? (0x80u >> q) : 0u;
            reinterpret_cast<uint8_t *>(valid_plane)[4 * w + (3 - (tid & 3))] = (uint8_t)valid;   // plane bit of offset o = 0x80000000 >> o
        }
        // pass 1: canonical key + partition id once per position, kept in registers as
        // [pid:10 @54 | key:42]; arrival rank inside (tile, partition) from the shared histogram
        uint64_t kp[kScatterPer];
        uint32_t rk[kScatterPer / 2];           // two 16-bit arrival ranks per register
#pragma unroll
        for (int q = 0; q < kScatterPer; q++) {
            const int o = o0 + q;
            kp[q] = ~0ULL;
            if ((q & 1) == 0) rk[q >> 1] = 0;
            if (valid & (0x80u >> q)) {
                const uint64_t f = window(hi, lo, o) >> (64 - 2 * kShortK);
                const uint64_t r = (o ? ((rlo >> (2 * o)) | (rhi << (64 - 2 * o))) : rlo) & kKey42;
                const uint64_t key = f <= r ? f : r;
                if (FILTER) {
                    const uint32_t tp = part_of(fmix64(key), flt_P);
                    if (tp < flt_lo || tp >= flt_hi) continue;
                }
                const uint32_t pt = pid_of<PMODE>(key, P);
                kp[q] = key | ((uint64_t)pt << kSmPtShift);
                rk[q >> 1] |= atomicAdd(&sm.hist[pt], 1u) << (16 * (q & 1));
            }
        }
        __syncthreads();
        tile_scan_and_claim<kScatterThreads>(sm, P, cursor, tid);
        __syncthreads();
        // pass 2: place records sorted by partition
#pragma unroll
        for (int q = 0; q < kScatterPer; q++) {
            const int o = o0 + q;
            if (kp[q] != ~0ULL) {
                const uint32_t pt = (uint32_t)(kp[q] >> kSmPtShift);
                sm.key[sm.hist[pt] + ((rk[q >> 1] >> (16 * (q & 1))) & 0xFFFFu)] = kp[q] | ((uint64_t)o << kRecOffShift) | ((uint64_t)wt << kSmLocShift);
            }
        }
        __syncthreads();
        const uint32_t total = sm.total;
        const uint64_t tile_w0 = w0 + tile * kTileWords;
        if (tid == kScatterThreads - 1) n_pos += total;
        for (uint32_t i = tid; i < total; i += kScatterThreads) {
            const uint64_t v = sm.key[i];
            const uint32_t pt = (uint32_t)(v >> kSmPtShift);
            const unsigned long long dst = sm.gbase[pt] + (i - sm.hist[pt]);
            const uint64_t rec = (v & ((1ULL << kSmLocShift) - 1)) | tag;
            const uint32_t wd = (uint32_t)(tile_w0 + ((v >> kSmLocShift) & (kTileWords - 1)));
            if (PEER) {   // remote (or local) stores over NVLink into owner pt's receive region (cap records, 0 = unbounded)
                if (cap == 0 || dst < cap) { peer.keys[pt][dst] = rec; peer.words[pt][dst] = wd; }
                else over = true;
            } else if (cap == 0 || dst < (uint64_t)(pt + 1) * cap) {   // fixed-capacity bins: an overflowing record is dropped, the cursor tells
                bkeys[dst] = rec;
                bword[dst] = wd;
            }
        }
        __syncthreads();
    }
    if (tid == kScatterThreads - 1 && n_pos && st) atomicAdd(&st->n_pos21, n_pos);
    if (over && st) atomicExch(&st->err_bin_overflow, 1u);
}

// the same tile sort for records that already sit in an array (received from other ranks, k-mer
// lists, position lists): 4096 records per tile, thread t takes records j*512+t (coalesced)
template <int PMODE>
__global__ void __launch_bounds__(256)
hist_rec_kernel(const uint64_t *__restrict__ in, uint64_t n, uint32_t P, unsigned long long *__restrict__ ghist) {
    __shared__ unsigned int sh[kMaxParts];
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(&sh[pid_of<PMODE>(__ldcs(in + i), P)], 1u);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
        if (sh[i]) atomicAdd(&ghist[i], (unsigned long long)sh[i]);
}

// AUXB: bytes of the auxiliary value that travels with each record (0 none, 4 = uint32 word index, 1 = uint8 hint).
// n_dev (optional): the record count lives on the device (written by an earlier kernel of the stream).
// Segmented input (in_cap > 0, a multiple of kTilePos): the input is a row of regions of in_cap records
// each, region r holding in_counts[r] records (what source rank r stored into this rank's receive buffer;
// the counts arrive through the peer mailbox, so they are read on the device) and n = regions * in_cap.
template <int PMODE, int AUXB>
__global__ void __launch_bounds__(kScatterThreads, 3)
scatter_rec_kernel(const uint64_t *__restrict__ in, const void *__restrict__ aux_in_, uint64_t n, uint32_t P,
                   unsigned long long *cursor, uint64_t *__restrict__ out, void *__restrict__ aux_out_, uint64_t cap = 0,
                   const unsigned long long *__restrict__ n_dev = nullptr,
                   uint64_t in_cap = 0, const unsigned long long *__restrict__ in_counts = nullptr, Stats *st = nullptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using SM = ScatterSmemT<true, false, AUXB>;
    SM &sm = *reinterpret_cast<SM *>(smem_raw);
    constexpr int NT = kScatterThreads, PER = kTilePos / NT;   // 512 threads x 8 records
    const int tid = threadIdx.x;
    if (n_dev) n = min(n, (uint64_t)*n_dev);
    const uint64_t n_tiles = (n + kTilePos - 1) / kTilePos;
    bool over = false;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t t0 = tile * kTilePos;
        uint64_t lim = n;
        if (in_cap) { const uint64_t r = t0 / in_cap; lim = r * in_cap + min((uint64_t)__ldcg(in_counts + r), in_cap); }
        if (t0 >= lim) continue;    // block-uniform
        for (uint32_t i = tid; i < P; i += NT) sm.hist[i] = 0;
        __syncthreads();
        // records, their auxiliary values and their partition ids are read / computed ONCE, coalesced, and kept in registers
        uint64_t rec[PER];
        uint32_t rk[PER / 2], pt2[PER / 2], aux[AUXB == 4 ? PER : (AUXB == 1 ? PER / 4 : 1)];
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint64_t i = t0 + j * NT + tid;
            rec[j] = i < lim ? __ldcs(in + i) : 0;
            if (AUXB == 4) aux[j] = i < lim ? __ldcs(static_cast<const uint32_t *>(aux_in_) + i) : 0u;
            if (AUXB == 1) {
                if ((j & 3) == 0) aux[j >> 2] = 0;
                if (i < lim) aux[j >> 2] |= (uint32_t)__ldcs(static_cast<const uint8_t *>(aux_in_) + i) << (8 * (j & 3));
            }
        }
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint64_t i = t0 + j * NT + tid;
            if ((j & 1) == 0) { rk[j >> 1] = 0; pt2[j >> 1] = 0; }
            if (i < lim) {
                const uint32_t pt = pid_of<PMODE>(rec[j], P);
                pt2[j >> 1] |= pt << (16 * (j & 1));
                rk[j >> 1] |= atomicAdd(&sm.hist[pt], 1u) << (16 * (j & 1));
            }
        }
        __syncthreads();
        tile_scan_and_claim<NT>(sm, P, cursor, tid);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint64_t i = t0 + j * NT + tid;
            if (i < lim) {
                const uint32_t pt = (pt2[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
                const uint32_t idx = sm.hist[pt] + ((rk[j >> 1] >> (16 * (j & 1))) & 0xFFFFu);
                sm.key[idx] = rec[j];
                sm.part[idx] = (uint16_t)pt;
                if (AUXB == 4) sm.aux4[idx] = aux[j];
                if (AUXB == 1) sm.aux1[idx] = (uint8_t)(aux[j >> 2] >> (8 * (j & 3)));
            }
        }
        __syncthreads();
        const uint32_t total = sm.total;
        for (uint32_t i = tid; i < total; i += NT) {
            uint32_t pt = sm.part[i];
            unsigned long long dst = sm.gbase[pt] + (i - sm.hist[pt]);
            if (cap && dst >= (uint64_t)(pt + 1) * cap) { over = true; continue; }
            out[dst] = sm.key[i];
            if (AUXB == 4) static_cast<uint32_t *>(aux_out_)[dst] = sm.aux4[i];
            if (AUXB == 1) static_cast<uint8_t *>(aux_out_)[dst] = sm.aux1[i];
        }
        __syncthreads();
    }
    if (over && st) atomicExch(&st->err_bin_overflow, 1u);
}

// Finds the slot of `key` in its partition or claims an empty one (with count 0), the first bucket of the probe
// sequence (bucket b of partition `base`) already loaded into s[]. Returns the GLOBAL slot index, ~0 when the
// partition is full; *created = this call claimed the slot. The pre-loaded bucket may be stale by the time it is
// used: slots only ever go empty -> key, an occupied slot never changes its key, and an empty-looking slot is
// claimed with a CAS, so a stale view is still a correct starting point.
template <bool PROBE_STATS>
__device__ __forceinline__ uint64_t find_or_claim_pre(const Table &t, uint64_t key, uint64_t base, uint64_t b, uint64_t s[4],
                                                      Stats *st, bool *created, unsigned *n_probes) {
    *created = false;
    for (uint64_t probe = 0; probe < t.nbp && probe < kMaxProbe; probe++) {
        uint64_t *bp = t.slots + 4 * (base + b);
        if (probe) ld_bucket(bp, s);
        if (PROBE_STATS) (*n_probes)++;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint64_t v = s[i];
            if ((v & kKey42) == key) return 4 * (base + b) + i;
            if (v == kEmpty) {
                uint64_t old = atomicCAS(ull(bp + i), kEmpty, key);
                if (old == kEmpty) { *created = true; return 4 * (base + b) + i; }
                if ((old & kKey42) == key) return 4 * (base + b) + i;
            }
        }
        b = (b + 1 == t.nbp) ? 0 : b + 1;
    }
    atomicExch(&st->err_table_full, 1u);
    return ~0ULL;
}

// The insert sweep is TWO kernels over the partition bins (profiles/microbench/insert_variants.cu: a kernel
// that mixes bucket loads and atomics on an L2-resident table runs at 67-73 G records/s whatever the batching,
// occupancy or dependence; loads alone 288 G/s, atomics alone 198 G/s, the two as separate kernels 103 G/s):
//   insert_find   bucket load(s) -> slot of the key (claimed with count 0 when new) -> 4-byte partition-relative
//                 slot index per record, streamed out beside the bins
//   insert_add    index stream -> one atomic add per record (returning: the 22-bit count field's wrap-around
//                 into the overflow side table must be seen)
// and MakeBF's verdict sweep reads the same index stream: count of record i = slots[base + idx[i]], no hashing.
// Work is handed out in chunks from ONE global counter instead of a static grid-stride loop: blocks run at
// different speeds, and with a static assignment the fast ones drift many partitions ahead, so that several
// hundred MB of table are live at once and L2 thrashes (measured: 22 G rec/s). With the shared counter all
// in-flight chunks lie within gridDim * kSweepChunk records of each other, i.e. inside one or two partitions.
constexpr int kSweepChunk = 2048;   // records per block per grab (8 per thread)
constexpr int kSweepPer = kSweepChunk / 256;
constexpr int kSweepBatch = 1;   // records in flight per thread: 1 at full occupancy beat 2 / 4 / 8 with fewer warps (84 / 94 / 113 / 166 ms)
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
template <bool PROBE_STATS, bool EXACT, int BATCH = kSweepBatch>
__global__ void __launch_bounds__(256, BATCH == 8 ? 2 : (BATCH == 4 ? 4 : (BATCH == 2 ? 6 : 8)))
insert_find_kernel(const uint64_t *__restrict__ bkeys, uint64_t n, Table table, Stats *st, uint32_t *__restrict__ bidx,
                   uint64_t cap = 0, const unsigned long long *__restrict__ bin_end = nullptr,
                   const unsigned long long *__restrict__ n_dev = nullptr) {
    __shared__ unsigned long long s_base;
    if (n_dev) n = min(n, (uint64_t)*n_dev);
    unsigned created = 0, probes = 0, longest = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_base = atomicAdd(&st->work, (unsigned long long)kSweepChunk);
        __syncthreads();
        const uint64_t cbase = s_base;
        if (cbase >= n) break;
        // fixed-capacity bins (cap > 0, a multiple of kSweepChunk): partition p's records are
        // [p*cap, min(bin_end[p], (p+1)*cap)); a chunk never straddles two partitions
        uint64_t lim = n, pbase = 0;
        if (!EXACT) { const uint64_t p = cbase / cap; lim = min((uint64_t)__ldg(bin_end + p), (p + 1) * cap); pbase = p * table.nbp; }
        if (cbase >= lim) continue;
#pragma unroll
        for (int half = 0; half < kSweepPer / BATCH; half++) {
            uint64_t rec[BATCH], s[BATCH][4];
            uint32_t b32[BATCH];
#pragma unroll
            for (int it = 0; it < BATCH; it++) {
                const uint64_t i = cbase + (uint64_t)(half * BATCH + it) * 256 + threadIdx.x;
                rec[it] = i < lim ? __ldcs(bkeys + i) : ~0ULL;
            }
#pragma unroll
            for (int it = 0; it < BATCH; it++) {
                if (rec[it] != ~0ULL) {
                    const uint64_t h = fmix64(rec[it] & kKey42);
                    b32[it] = (uint32_t)sub_of(h, table.nbp);
                    const uint64_t base = EXACT ? (uint64_t)part_of(h, table.P) * table.nbp : pbase;
                    ld_bucket(table.slots + 4 * (base + b32[it]), s[it]);
                }
            }
#pragma unroll
            for (int it = 0; it < BATCH; it++) {
                if (rec[it] != ~0ULL) {
                    const uint64_t i = cbase + (uint64_t)(half * BATCH + it) * 256 + threadIdx.x;
                    const uint64_t key = rec[it] & kKey42;
                    const uint64_t base = EXACT ? (uint64_t)part_of(fmix64(key), table.P) * table.nbp : pbase;
                    bool cr;
                    unsigned np = 0;
                    const uint64_t slot = find_or_claim_pre<PROBE_STATS>(table, key, base, b32[it], s[it], st, &cr, &np);
                    created += cr;
                    // the index stream also carries the record's offset-in-word and rank when they fit above the slot bits
                    // streaming store: the index stream must not push the table partition out of L2
                    __stcs(bidx + i, slot == ~0ULL ? kNoSlot
                                                   : (uint32_t)(slot - 4 * base) | (table.sb ? (uint32_t)((rec[it] >> kRecOffShift) & 0x1FF) << table.sb : 0u));
                    if (PROBE_STATS) { probes += np; longest = max(longest, np); }
                }
            }
        }
    }
    unsigned long long tot = warp_sum(created);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&st->n_cand, tot);
    if (PROBE_STATS) {
        unsigned long long tp = warp_sum(probes);
        unsigned mx = __reduce_max_sync(0xffffffffu, longest);
        if ((threadIdx.x & 31) == 0) { if (tp) atomicAdd(&st->probes, tp); atomicMax(&st->max_probe, mx); }
    }
}
// EXACT: contiguous (histogram-sized) bins — the partition of a record comes from its key, not from its place
template <bool EXACT>
__global__ void __launch_bounds__(256, 8)
insert_add_kernel(const uint32_t *__restrict__ bidx, const uint64_t *__restrict__ bkeys, uint64_t n, Table table, Ovf ovf, Stats *st,
                  uint64_t cap, const unsigned long long *__restrict__ bin_end, const unsigned long long *__restrict__ n_dev) {
    __shared__ unsigned long long s_base;
    if (n_dev) n = min(n, (uint64_t)*n_dev);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_base = atomicAdd(&st->work, (unsigned long long)kSweepChunk);
        __syncthreads();
        const uint64_t cbase = s_base;
        if (cbase >= n) break;
        uint64_t lim = n, pbase = 0;
        if (!EXACT) { const uint64_t p = cbase / cap; lim = min((uint64_t)__ldg(bin_end + p), (p + 1) * cap); pbase = 4 * p * table.nbp; }
        if (cbase >= lim) continue;
        uint32_t r[kSweepPer];
#pragma unroll
        for (int it = 0; it < kSweepPer; it++) {
            const uint64_t i = cbase + (uint64_t)it * 256 + threadIdx.x;
            r[it] = i < lim ? __ldcs(bidx + i) : kNoSlot;
        }
#pragma unroll
        for (int it = 0; it < kSweepPer; it++) {
            if (r[it] == kNoSlot) continue;
            uint64_t base = pbase;
            if (EXACT) {
                const uint64_t i = cbase + (uint64_t)it * 256 + threadIdx.x;
                base = 4 * (uint64_t)part_of(fmix64(__ldcs(bkeys + i) & kKey42), table.P) * table.nbp;
            }
            const uint32_t rel = table.sb ? (r[it] & ((1u << table.sb) - 1)) : r[it];
            const uint64_t old = atomicAdd(ull(table.slots + base + rel), kCntOne);
            if ((old >> 42) == kCntFieldMax) ovf_add(ovf, old & kKey42, st);   // the 22-bit field wrapped to 0
        }
    }
}

// position record [rank:8 @56 | stream position:56] of a count record and its word index
__device__ __forceinline__ uint64_t posrec_of(uint64_t rec, uint32_t wd) {
    return ((uint64_t)wd * 32 + ((rec >> kRecOffShift) & 31)) | (((rec >> kRecRankShift) & 0xFF) << kPosRankShift);
}

// RMQ window minimum >= threshold  <=>  every 21-mer flag in the window is set
// (reference src/MakeBloomFilter.cpp:62,75). Window length x = k-20 <= 12 for k <= 32.
__device__ __forceinline__ uint32_t solid_bits(uint32_t g0, uint32_t g1, int x) {
    uint64_t a = ((uint64_t)g0 << 32) | g1;
    int w = 1;
    while (2 * w <= x) { a &= a << w; w *= 2; }
    a &= a << (x - w);
    return (uint32_t)(a >> 32);
}

// pass 1: solid-position bit plane + number of BF.add calls
__global__ void __launch_bounds__(256)
solid_kernel(const uint32_t *__restrict__ good21, uint64_t n_words, int k, uint32_t *__restrict__ solid, Stats *st) {
    unsigned n = 0;
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += stride) {
        uint32_t s = solid_bits(__ldg(good21 + w), __ldg(good21 + w + 1), k - kShortK + 1);
        solid[w] = s;
        n += __popc(s);
    }
    unsigned long long a = warp_sum(n);
    if ((threadIdx.x & 31) == 0 && a) atomicAdd(&st->n_adds, a);
}

// pass 2: BF.add for every solid position (reference src/MakeBloomFilter.cpp:75-77). Positions of
// the same canonical k-mer set identical bits, so each distinct k-mer is added once: the solid
// set de-duplicates and its insertion winner does the num_hashes atomicOr's.
template <bool HAS_MASK>
__global__ void __launch_bounds__(256, 8)   // 32 registers: DRAM-latency bound, lives on occupancy
makebf_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ nmask,
              const uint32_t *__restrict__ solid, uint64_t n_words, int k, KSet set, Stats *st) {
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    bool full = false;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += stride) {
        uint32_t s = __ldg(solid + w);
        if (!s) continue;
        uint64_t hi = __ldg(packed + w), lo = __ldg(packed + w + 1);
        uint64_t mhi = 0, mlo = 0;
        if (HAS_MASK) { mhi = spread32(__ldg(nmask + w)); mlo = spread32(__ldg(nmask + w + 1)); }
        while (s) {
            int o = __clz(s);
            s &= ~(0x80000000u >> o);
            uint64_t x = window(hi, lo, o);
            uint64_t m2 = HAS_MASK ? window(mhi, mlo, o) : 0;
            if (set_insert(set, canonical_from_window(x, m2, k)) < 0) full = true;
        }
    }
    if (full) atomicExch(&st->err_table_full, 1u);
}

// ---- binned de-duplication (default for large inputs) ---------------------------------------------------
// The direct kernel above pays one DRAM-random set access per solid position (2.97 G at configs[1],
// 33 G/s). Binned: K1 tile-sorts the canonical k-mers of the solid positions by SET partition (the
// scatter21 machinery, 8-byte records, fixed-capacity bins: partitions are hash-uniform), K2 sweeps
// the bins with the work-counter hand-out of insert_bins so that the whole grid probes one ~24 MB
// partition at a time out of L2. 95 % of the probes are plain hits (load + compare, no atomic).
// Every solid occurrence also carries an adjacency HINT: when the position before / after it in the
// read is solid too, that neighbouring k-mer was added to the filter, so the corresponding direction of
// CheckDirections (reference src/DeBruijnGraph.cpp:326-345) is certainly "recorded" and needs no query
// later. Bit d of the hint = direction d of the CANONICAL k-mer (0-3 left extension by A,C,G,T; 4-7
// right extension); for an occurrence whose canonical form is the reverse strand, left and right swap
// and the base is complemented. Occurrences near a non-ACGT character carry no hint (the reference adds
// such k-mers with its both-strands-read-as-A quirk, so the neighbour relation need not hold).
struct PeerOutK { uint64_t *keys[kMaxPeers]; uint8_t *hints[kMaxPeers]; };
template <bool HAS_MASK, bool PEER>
__global__ void __launch_bounds__(kScatterThreads, 3)
scatter_kmer_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ nmask,
                    const uint32_t *__restrict__ solid, uint64_t w0, uint64_t w1, int k, uint32_t P,
                    unsigned long long *cursor, uint64_t *__restrict__ bins, uint8_t *__restrict__ hbins, uint64_t cap,
                    Stats *st, PeerOutK peer = PeerOutK()) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using SM = ScatterSmemT<true, true>;    // loc[] carries the hint of the sorted record
    SM &sm = *reinterpret_cast<SM *>(smem_raw);
    const int tid = threadIdx.x;
    const int wt = tid >> 2;                    // word of the tile
    const int o0 = (tid & 3) * kScatterPer;     // first offset this thread handles
    const uint64_t n_tiles = (w1 - w0 + kTileWords - 1) / kTileWords;
    const uint64_t km = kmask(k);
    bool over = false;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t i = tid; i < P; i += kScatterThreads) sm.hist[i] = 0;
        __syncthreads();
        const uint64_t w = w0 + tile * kTileWords + wt;
        uint64_t hi = 0, lo = 0, rlo = 0, rhi = 0, S = 0, mbits = 0;
        uint32_t s = 0, prev_base = 0;          // s bit (7 - q): offset o0 + q is a solid position
        bool pm = false;
        if (w < w1) {
            const uint32_t sc = __ldg(solid + w);
            s = (sc >> (24 - o0)) & 0xFFu;
            if (s) {
                hi = __ldg(packed + w); lo = __ldg(packed + w + 1);
                uint64_t chi = ~hi, clo = ~lo;
                if (HAS_MASK) {
                    const uint32_t ma = __ldg(nmask + w), mb = __ldg(nmask + w + 1);
                    chi &= ~spread32(ma); clo &= ~spread32(mb);
                    mbits = ((uint64_t)ma << 32) | mb;
                    pm = w ? (__ldg(nmask + w - 1) & 1u) : false;
                }
                rlo = rev2(chi); rhi = rev2(clo);
                // S bit (32 - j): position 32w + j is solid, j = -1 .. 32
                S = ((uint64_t)(w ? (__ldg(solid + w - 1) & 1u) : 0u) << 33) | ((uint64_t)sc << 1) | (__ldg(solid + w + 1) >> 31);
                prev_base = w ? (uint32_t)(__ldg(packed + w - 1) & 3) : 0u;
            }
        }
        uint64_t key[kScatterPer];
        uint32_t rk[kScatterPer / 2], hint[kScatterPer / 4];   // 16-bit arrival ranks, 8-bit hints
#pragma unroll
        for (int q = 0; q < kScatterPer; q++) {
            const int o = o0 + q;
            key[q] = kEmpty;
            if ((q & 1) == 0) rk[q >> 1] = 0;
            if ((q & 3) == 0) hint[q >> 2] = 0;
            if (s & (0x80u >> q)) {
                const uint64_t f = window(hi, lo, o) >> (64 - 2 * k);
                const uint64_t r = (o ? ((rlo >> (2 * o)) | (rhi << (64 - 2 * o))) : rlo) & km;
                const bool fwd = f <= r;
                key[q] = fwd ? f : r;
                rk[q >> 1] |= atomicAdd(&sm.hist[PEER ? owner_of(key[q], P) : kset_part(key[q], P)], 1u) << (16 * (q & 1));
                bool lh = (S >> (33 - o)) & 1, rh = (S >> (31 - o)) & 1;   // positions p - 1 and p + 1
                if (HAS_MASK) {
                    if (lh) lh = o ? (((mbits << (o - 1)) >> (63 - k)) == 0) : (!pm && (mbits >> (64 - k)) == 0);
                    if (rh) rh = ((mbits << o) >> (63 - k)) == 0;
                }
                uint32_t h = 0;
                if (lh) { const uint32_t lb = o ? (uint32_t)((hi >> (64 - 2 * o)) & 3) : prev_base; h |= fwd ? (1u << lb) : (16u << (3 - lb)); }
                if (rh) {
                    const int j = o + k;
                    const uint32_t rb = (uint32_t)((j < 32 ? (hi >> (62 - 2 * j)) : (lo >> (126 - 2 * j))) & 3);
                    h |= fwd ? (16u << rb) : (1u << (3 - rb));
                }
                hint[q >> 2] |= h << (8 * (q & 3));
            }
        }
        __syncthreads();
        tile_scan_and_claim<kScatterThreads>(sm, P, cursor, tid);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kScatterPer; q++) {
            if (s & (0x80u >> q)) {
                uint32_t pt = PEER ? owner_of(key[q], P) : kset_part(key[q], P);
                uint32_t idx = sm.hist[pt] + ((rk[q >> 1] >> (16 * (q & 1))) & 0xFFFFu);
                sm.key[idx] = key[q];
                sm.part[idx] = (uint16_t)pt;
                sm.loc[idx] = (uint16_t)((hint[q >> 2] >> (8 * (q & 3))) & 0xFFu);
            }
        }
        __syncthreads();
        const uint32_t total = sm.total;
        for (uint32_t i = tid; i < total; i += kScatterThreads) {
            uint32_t pt = sm.part[i];
            unsigned long long dst = sm.gbase[pt] + (i - sm.hist[pt]);
            if (PEER) {   // owner pt's receive region for this source (cap records)
                if (dst < cap) { peer.keys[pt][dst] = sm.key[i]; peer.hints[pt][dst] = (uint8_t)sm.loc[i]; }
                else over = true;
            } else if (dst < (uint64_t)(pt + 1) * cap) { bins[dst] = sm.key[i]; hbins[dst] = (uint8_t)sm.loc[i]; }
            else over = true;
        }
        __syncthreads();
    }
    if (over) atomicExch(PEER ? &st->err_bin_overflow : &st->err_kbin_overflow, 1u);   // deferred check on the host
}
// hints (optional): one byte per record, OR-ed into the byte of the set slot that holds the k-mer
// (a plain load first: 95 % of the records repeat a hint that is already there)
__global__ void __launch_bounds__(256, 6)   // 42 registers; capping at 32 spills 250 bytes and costs 19 ms
set_sweep_kernel(const uint64_t *__restrict__ bins, const uint8_t *__restrict__ hbins, uint64_t n, uint64_t cap,
                 const unsigned long long *__restrict__ bin_end, KSet set, uint32_t *__restrict__ slot_hint, Stats *st) {
    __shared__ unsigned long long s_base;
    bool full = false;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_base = atomicAdd(&st->work, (unsigned long long)kSweepChunk);
        __syncthreads();
        const uint64_t cbase = s_base;
        if (cbase >= n) break;
        const uint64_t pq = cbase / cap;
        const uint64_t lim = min(min((uint64_t)__ldg(bin_end + pq), (pq + 1) * cap), n);
        if (cbase >= lim) continue;
        uint64_t rec[kSweepPer];
        uint32_t hb[kSweepPer / 4];
#pragma unroll
        for (int it = 0; it < kSweepPer; it++) {
            uint64_t i = cbase + it * 256 + threadIdx.x;
            rec[it] = i < lim ? __ldcs(bins + i) : kEmpty;
            if ((it & 3) == 0) hb[it >> 2] = 0;
            if (hbins && i < lim) hb[it >> 2] |= (uint32_t)__ldcs(hbins + i) << (8 * (it & 3));
        }
#pragma unroll
        for (int it = 0; it < kSweepPer; it++) {
            if (rec[it] == kEmpty) continue;
            uint64_t slot;
            const int ins = set_insert_slot(set, rec[it], &slot);
            if (ins < 0) { full = true; continue; }
            // Hints are optional knowledge (a direction without one is simply queried later). The occurrence that creates the
            // k-mer always leaves its hint; of the repeats only every 4th record looks at the slot's hint byte (an extra L2
            // load per record otherwise): a third of the solid set occurs just twice, so plain sampling would lose too many.
            if (ins == 0 && (it & 3)) continue;
            const uint32_t h = ((hb[it >> 2] >> (8 * (it & 3))) & 0xFFu) << (8 * (slot & 3));
            if (h && (ins || (__ldcg(slot_hint + (slot >> 2)) & h) != h)) atomicOr(slot_hint + (slot >> 2), h);
        }
    }
    if (full) atomicExch(&st->err_table_full, 1u);
}

// distinct solid k-mers = occupied slots of the set, compacted into a dense list. One global
// atomic per block-iteration (a per-winner atomic on one list cursor serialises in L2).
__global__ void __launch_bounds__(256)
compact_set_kernel(const uint64_t *__restrict__ set, uint64_t n_slots, uint64_t *__restrict__ list,
                   uint64_t list_cap, Stats *st, const uint8_t *__restrict__ slot_hint = nullptr, uint8_t *__restrict__ adj = nullptr) {
    __shared__ unsigned s_wtot[8];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    uint64_t n_round = (n_slots + stride - 1) / stride;
    for (uint64_t r = 0; r < n_round; r++) {
        uint64_t i = r * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
        uint64_t v = i < n_slots ? __ldcs(set + i) : kEmpty;
        unsigned m = __ballot_sync(0xffffffffu, v != kEmpty);
        if (lane == 0) s_wtot[wid] = __popc(m);
        __syncthreads();
        unsigned wbase = 0, total = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { unsigned t = s_wtot[q]; if (q < wid) wbase += t; total += t; }
        if (threadIdx.x == 0 && total) s_base = atomicAdd(&st->n_distinct_solid, (unsigned long long)total);
        __syncthreads();
        if (v != kEmpty) {
            unsigned long long j = s_base + wbase + __popc(m & ((1u << lane) - 1));
            if (j < list_cap) { list[j] = v; if (slot_hint) adj[j] = slot_hint[i]; }   // adjacency bytes start as the hints
        }
        __syncthreads();
    }
}

// BF.add (reference src/bloomfilter.cpp:69-74) for every distinct solid k-mer, dense over the list.
// The filter is usually larger than L2, so it is processed in `n_seg` passes over segments of
// seg_bits bits (<= ~24 MB): each pass recomputes the num_hashes bit indices (a few dozen ALU ops
// each) and only sets the bits that fall into its segment, so the atomics of a pass all hit an
// L2-resident window instead of DRAM (random RED from DRAM: 21.8 G/s, from L2: ~215 G/s).
__global__ void __launch_bounds__(256)
bloom_list_kernel(const uint64_t *__restrict__ kmers, uint64_t n, Bloom bf, uint64_t seg_lo, uint64_t seg_hi) {
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t h1, h2;
        double_hash(std_hash_kmer1(__ldg(kmers + i), bf.nbytes), h1, h2);
        uint64_t x = h1;
        for (int q = 0; q < bf.nh; q++, x += h2) {
            uint64_t bit = fastmod(x, bf.fm);
            if (bit < seg_lo || bit >= seg_hi) continue;
            uint32_t m = 1u << (bit & 31);
            uint32_t *wp = bf.bits + (bit >> 5);
            atomicOr(wp, m);   // RED.OR: the segment is L2 resident, a test-load first would only add a dependent trip
        }
    }
}

// first solid k-mer of each read (reference src/MakeBloomFilter.cpp:79-83)
__global__ void seeds_kernel(const uint64_t *__restrict__ off, uint64_t n_reads, const uint32_t *__restrict__ solid,
                             int k, int64_t *__restrict__ seed_pos) {
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        uint64_t s = off[r], e = off[r + 1];
        int64_t found = -1;
        if (e - s >= (uint64_t)k) {
            uint64_t last = e - k;
            for (uint64_t w = s >> 5; w <= (last >> 5); w++) {
                uint32_t v = __ldg(solid + w);
                if (w == (s >> 5)) v &= 0xFFFFFFFFu >> (s & 31);
                if (v) {
                    uint64_t p = (w << 5) + __clz(v);
                    if (p <= last) found = (int64_t)(p - s);
                    break;
                }
            }
        }
        seed_pos[r] = found;
    }
}

// CheckDirections (reference src/DeBruijnGraph.cpp:326-345): 8 lanes per k-mer, one direction each
// `set` (may be null): the distinct solid k-mers. Every member was added to the filter, so
// possiblyContains is certainly true for it and its num_hashes probes are skipped; only
// non-members (which mostly fail after a few probes) walk the filter.
// HINT: adj[] comes in holding the hint byte of every k-mer (directions known to be recorded because the
// neighbouring k-mer was seen solid right beside this one in a read): those lanes skip their query.
template <bool HINT>
__global__ void __launch_bounds__(256, 8)   // 32 registers: the kernel lives on occupancy (34 registers cost 40 %: 97 -> 135 ms)
adjacency_kernel(const uint64_t *__restrict__ kmers, uint64_t n, int k, Bloom bf, KSet set, KSet set_b,
                 uint8_t *__restrict__ adj, Stats *st) {
    const int lane = threadIdx.x & 31;
    const int d = lane & 7, g = lane >> 3;
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t n_warps = (gridDim.x * (uint64_t)blockDim.x) >> 5;
    unsigned edges = 0;
    for (uint64_t base = warp * 4; base < n; base += n_warps * 4) {
        uint64_t i = base + g;
        bool rec = false;
        if (HINT && i < n) rec = (adj[i] >> d) & 1;   // read by all 8 lanes before the ballot below, rewritten after it
        if (i < n && !rec) {
            uint64_t nb = neighbour(__ldg(kmers + i), d, k);
            uint64_t rc = revcomp(nb, k);
            uint64_t c = nb <= rc ? nb : rc;   // IsRecorded canonicalises (DeBruijnGraph.cpp:320-321)
            // set_b (multi-GPU): the rank's locally seen solid k-mers, which their owners added. It holds
            // nearly every solid k-mer (each rank samples the whole genome), so it is asked INSTEAD of the
            // owned set: the rare owned-but-not-seen member just walks its num_hashes probes.
            rec = (set_b.slots ? set_contains(set_b, c) : (set.slots && set_contains(set, c))) || bloom_query(bf, c);
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

// Closure of the k-mer set under recorded neighbours (for the host unitig walk): the walk only
// ever moves to a neighbour that CheckDirections reported, i.e. one whose possiblyContains is true.
// Solid k-mers are in the set already; a reported neighbour that is NOT in the set is a Bloom
// false positive ("phantom"). Phantoms are inserted and appended to the list so that the next
// adjacency pass covers them too; iterating to a fixed point gives the walk a table that answers
// every CheckDirections it can possibly ask, with the reference's false positives included.
__global__ void __launch_bounds__(256)
closure_kernel(uint64_t *list, const uint8_t *__restrict__ adj, uint64_t lo, uint64_t hi, int k,
               KSet set, uint64_t list_cap, Stats *st) {
    __shared__ unsigned s_wtot[8];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    uint64_t n_round = (hi - lo + stride - 1) / stride;
    for (uint64_t r = 0; r < n_round; r++) {
        uint64_t i = lo + r * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
        uint64_t fresh[8];
        unsigned mine = 0;
        if (i < hi) {
            uint64_t km = list[i];
            unsigned a = adj[i];
            while (a) {
                int d = __ffs(a) - 1;
                a &= a - 1;
                uint64_t nb = neighbour(km, d, k);
                uint64_t rc = revcomp(nb, k);
                uint64_t c = nb <= rc ? nb : rc;
                int ins = set_insert(set, c);
                if (ins > 0) fresh[mine++] = c;
                else if (ins < 0) atomicExch(&st->err_table_full, 1u);
            }
        }
        unsigned incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        if (lane == 31) s_wtot[wid] = incl;
        __syncthreads();
        unsigned wbase = 0, total = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { unsigned t = s_wtot[q]; if (q < wid) wbase += t; total += t; }
        if (threadIdx.x == 0 && total) s_base = atomicAdd(&st->n_distinct_solid, (unsigned long long)total);
        __syncthreads();
        unsigned long long j = s_base + wbase + incl - mine;
        for (unsigned q = 0; q < mine; q++, j++)
            if (j < list_cap) list[j] = fresh[q];
        __syncthreads();
    }
}

// walk roots (seed k-mers, oriented): make sure their canonical form has a table entry
__global__ void roots_kernel(const uint64_t *__restrict__ roots, uint64_t n, int k, KSet set,
                             uint64_t *list, uint64_t list_cap, Stats *st) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t km = roots[i], rc = revcomp(km, k);
    uint64_t c = km <= rc ? km : rc;
    int ins = set_insert(set, c);
    if (ins > 0) {
        unsigned long long j = atomicAdd(&st->n_distinct_solid, 1ULL);
        if (j < list_cap) list[j] = c;
    } else if (ins < 0) atomicExch(&st->err_table_full, 1u);
}

// probe length of a successful lookup of every listed k-mer in the solid set (bench.py --config 4)
__global__ void set_probe_stats_kernel(const uint64_t *__restrict__ kmers, uint64_t n, KSet t, Stats *st) {
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    unsigned long long tot = 0;
    unsigned mx = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t key = __ldg(kmers + i), h = fmix64(key);
        const uint64_t base = (uint64_t)part_of(h, t.P) * t.nbp;
        uint64_t b = sub_of(h, t.nbp);
        unsigned np = 0;
        for (uint64_t probe = 0; probe < t.nbp; probe++) {
            uint64_t s[4];
            ld_bucket64(t.slots + 4 * (base + b), s);
            np++;
            if (s[0] == key || s[1] == key || s[2] == key || s[3] == key) break;
            if (s[0] == kEmpty || s[1] == kEmpty || s[2] == kEmpty || s[3] == kEmpty) break;
            b = (b + 1 == t.nbp) ? 0 : b + 1;
        }
        tot += np; mx = max(mx, np);
    }
    tot = __reduce_add_sync(0xffffffffu, (unsigned)tot);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&st->probes, tot); atomicMax(&st->max_probe, mx); }
}

// ---- small batch / export kernels -------------------------------------------------------------------
__global__ void export_counts_kernel(const uint64_t *__restrict__ table, uint64_t n_slots, Ovf ovf, Stats *st,
                                     uint64_t thr, uint64_t *keys, uint64_t *counts, uint64_t cap) {
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    const unsigned n_overflow = st->n_overflow;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        uint64_t v = table[i];
        if (v == kEmpty) continue;
        uint64_t key = v & kKey42, c = v >> 42;
        if (n_overflow) c += ovf_get(ovf, key) << 22;
        if (keys) {
            unsigned long long idx = atomicAdd(&st->n_export, 1ULL);
            if (idx < cap) { keys[idx] = key; counts[idx] = c; }
        } else if (c >= thr) {
            atomicAdd(&st->n_good21, 1ULL);
        }
    }
}
__global__ void lookup_counts_kernel(Table table, Ovf ovf, const Stats *st,
                                     const uint64_t *__restrict__ keys, uint64_t n, uint64_t *counts) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) counts[i] = count_lookup(table, keys[i], ovf, st->n_overflow);
}
__global__ void bf_add_kernel(Bloom bf, const uint64_t *__restrict__ kmers, uint64_t n) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) bloom_add(bf, kmers[i]);
}
__global__ void bf_query_kernel(Bloom bf, const uint64_t *__restrict__ kmers, uint64_t n, uint8_t *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = bloom_query(bf, kmers[i]) ? 1 : 0;
}
__global__ void double_hash_kernel(int nbytes, const uint64_t *__restrict__ kmers, uint64_t n, uint64_t *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        uint64_t h1, h2;
        double_hash(std_hash_kmer1(kmers[i], nbytes), h1, h2);
        out[2 * i] = h1; out[2 * i + 1] = h2;
    }
}

// ------------------------------------------------------------------------------------------------
// host side: context + C ABI
// ------------------------------------------------------------------------------------------------

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(P3_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

struct p3_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    // reads
    const uint64_t *d_packed = nullptr; const uint64_t *d_off = nullptr; const uint32_t *d_nmask = nullptr;
    uint64_t *own_packed = nullptr; uint64_t *own_off = nullptr; uint32_t *own_nmask = nullptr;
    uint64_t cap_packed = 0, cap_off = 0, cap_nmask = 0, cap_rend = 0, cap_planes = 0, cap_seed = 0;  // bytes
    uint64_t total_bases = 0, n_reads = 0, n_words = 0;
    uint32_t *d_rend = nullptr;
    bool have_reads = false;
    // count table
    uint64_t *d_table = nullptr; uint64_t nb = 0;     // nb = total buckets = P * nbp
    uint32_t parts = 1; uint64_t nbp = 0;
    bool binned = true;                                // P3_COUNT_MODE=direct switches it off
    uint64_t *d_bkeys = nullptr; uint32_t *d_bword = nullptr; uint64_t cap_bkeys = 0, cap_bword = 0;
    uint32_t *d_below = nullptr; uint64_t cap_below = 0;   // one bit per table slot: final count < threshold (verdict sweep)
    uint32_t *d_bidx = nullptr; uint64_t cap_bidx = 0;   // partition-relative slot index of every binned record (insert_find -> insert_add -> verdict sweep)
    uint32_t *d_valid = nullptr; uint64_t cap_valid = 0;
    uint64_t bin_cap = 0, bin_n = 0; bool bins_valid = false;   // the partition bins of the last count (one chunk) are still there for the verdict sweep
    uint64_t bin_upper = 0; bool bin_exact = false;
    unsigned long long *d_binmeta = nullptr;   // [kMaxParts] bin ends + [kMaxParts] record total of the current bins
    unsigned long long *d_ghist = nullptr, *d_cursor = nullptr;
    std::vector<cudaEvent_t> evpool;       // per-chunk timing events of the binned count (no sync inside the chunk loop)
    // host-buffer path (p3_assemble_hot_path*): the 2-bit staging arrives in pieces on a copy stream while the binning
    // kernel already works on the pieces that are there; results leave on the copy stream while CheckDirections runs
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> up_ev; uint64_t up_pieces = 0, up_piece_words = 0;   // pending upload (consumed by the next count)
    cudaEvent_t ev_main = nullptr, ev_copy = nullptr;
    bool probe_stats = false;              // P3_PROBE_STATS: instrumented insert kernels (bench.py --config 4)
    unsigned long long count_probes = 0; unsigned count_max_probe = 0;
    bool attrs_set = false;
    float ms_sub[4] = {0, 0, 0, 0};
    float ms_bloom = 0;                    // hist, scatter, insert, cand_check
    uint64_t n_chunks = 0, binned_pos = 0; bool pos_on_host = false;
    Table table() const {
        Table t; t.slots = d_table; t.nbp = nbp; t.P = parts;
        // slot:sb | offset-in-word:5 | rank:4 in the 32-bit index stream. (1 << sb) is strictly larger than the slots of a
        // partition, so the all-ones word (kNoSlot) is never a valid record even when all 32 bits are in use (48 MB partitions: sb = 23)
        uint32_t sb = 1;
        while ((1ull << sb) <= nbp * 4) sb++;
        t.sb = sb + 9 <= 32 ? sb : 0;
        return t;
    }
    uint64_t *d_ovf_keys = nullptr; unsigned long long *d_ovf_wraps = nullptr;
    bool have_counts = false;
    // make_bf
    uint32_t *d_good21 = nullptr, *d_solid = nullptr;
    uint32_t *d_proven2 = nullptr; uint64_t cap_proven = 0;
    uint64_t *d_set = nullptr; uint64_t nbs = 0; uint32_t set_parts = 1;   // nbs = total buckets = set_parts * buckets per partition
    bool set_valid = false;   // d_set holds a subset of what the current filter contains
    uint32_t *d_hint = nullptr; uint64_t cap_hint = 0;   // one hint byte per set slot (binned de-duplication)
    bool hints_valid = false; // d_adj[0..n_distinct) holds the hint bytes of the current list
    const uint64_t *d_set_b = nullptr; uint64_t nbs_b = 0; uint32_t parts_b = 1;   // multi-GPU: second (locally seen) solid set, not owned here
    KSet kset() const { KSet t; t.slots = d_set; t.P = set_parts ? set_parts : 1; t.nbp = nbs / t.P; return t; }
    KSet kset_b() const { KSet t; t.slots = const_cast<uint64_t *>(d_set_b); t.P = parts_b ? parts_b : 1; t.nbp = nbs_b / t.P; return t; }
    static KSet no_set() { KSet t; t.slots = nullptr; t.P = 1; t.nbp = 0; return t; }
    uint64_t *d_list = nullptr; uint64_t list_cap = 0;
    uint32_t *d_bloom = nullptr; uint64_t bloom_words = 0;
    int64_t *d_seed = nullptr;
    uint64_t filter_size = 0; uint32_t num_hashes = 0; uint32_t k = 0;
    bool have_bf = false, have_solid = false;
    // adjacency
    uint8_t *d_adj = nullptr; uint64_t adj_cap = 0; bool have_adj = false;
    uint64_t n_solid = 0, n_closed = 0; bool closed = false;
    Stats *d_stats = nullptr; Stats h_stats;
    cudaEvent_t ev[16];
    float ms[5] = {0, 0, 0, 0, 0};
    Ovf ovf() const { Ovf o; o.keys = d_ovf_keys; o.wraps = d_ovf_wraps; return o; }
    Bloom bloom() const {
        Bloom b; b.bits = d_bloom; b.fm = make_fastmod(filter_size); b.nh = (int)num_hashes; b.nbytes = (int)((2 * k + 7) / 8);
        return b;
    }
    uint64_t wrap() const { return filter_size ? ((~0ULL % filter_size) + 1) % filter_size : 0; }
    int grid(int blocks_per_sm = 8) const { return n_sm * blocks_per_sm; }
};

// Per-context state of the optional paths (multi-GPU, multi-word k, binned Bloom adds) lives beside
// the context in small registries. A context is used by one host thread at a time, but different
// threads may create / use / destroy DIFFERENT contexts concurrently, so the registries lock around
// lookup, insertion and erasure (references to unordered_map elements stay valid across rehashes).
template <class S> struct CtxStates {
    std::mutex mu;
    std::unordered_map<struct p3_ctx *, S> m;
    S &get(struct p3_ctx *c) { std::lock_guard<std::mutex> g(mu); return m[c]; }
    S *find(struct p3_ctx *c) { std::lock_guard<std::mutex> g(mu); auto it = m.find(c); return it == m.end() ? nullptr : &it->second; }
    void erase(struct p3_ctx *c) { std::lock_guard<std::mutex> g(mu); m.erase(c); }
};
static void mg_release(struct p3_ctx *c);   // p3_multi.inc.cu
static void long_release(struct p3_ctx *c); // p3_long.inc.cu
static void bloom_release(struct p3_ctx *c);                                  // p3_bloom.inc.cu
static int bloom_add_binned(struct p3_ctx *c, uint64_t n, bool *done);
static int make_bf_long(struct p3_ctx *c, uint32_t k, uint64_t solid_slots, uint64_t est_distinct);
static int verdict_sweep(struct p3_ctx *c, uint64_t thr, bool force_direct, bool *binned_any);   // p3_bloom.inc.cu
static int bloom_add_direct_long(struct p3_ctx *c, uint64_t nd);
static int adjacency_long(struct p3_ctx *c, const uint64_t *d_words, uint64_t n, uint8_t *d_adj, struct p3::Stats *st);
static const uint64_t *long_words(struct p3_ctx *c);
static int long_batch(struct p3_ctx *c, int op, uint32_t k, const uint64_t *h_kmers, uint64_t n, void *h_out);
template <typename T> static void dfree(T *&p) { if (p) { cudaFree((void *)p); p = nullptr; } }
// device temporaries of the batch entry points: released on EVERY return path (CU() returns early on an error)
struct TmpFree {
    void **slot[6]; int n = 0;
    template <typename T> void own(T **pp) { slot[n++] = reinterpret_cast<void **>(pp); }
    ~TmpFree() { for (int i = 0; i < n; i++) if (*slot[i]) { cudaFree(*slot[i]); *slot[i] = nullptr; } }
};
// grow-only device buffer: reallocates only when the request exceeds the capacity, so repeated
// runs on same-sized inputs never touch cudaMalloc/cudaFree (both synchronise the device)
template <typename T> static cudaError_t ensure(T *&p, uint64_t &cap_bytes, uint64_t need_bytes) {
    if (p && cap_bytes >= need_bytes) return cudaSuccess;
    dfree(p); cap_bytes = 0;
    cudaError_t e = cudaMalloc((void **)&p, need_bytes ? need_bytes : 1);
    if (e == cudaSuccess) cap_bytes = need_bytes;
    return e;
}

static int pull_stats(p3_ctx *c) {
    CU(cudaMemcpyAsync(&c->h_stats, c->d_stats, sizeof(Stats), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->binned) {  // distinct keys == keys created; the multi-GPU owner keeps its position total on the host
        if (c->pos_on_host) c->h_stats.n_pos21 = c->binned_pos;
        c->h_stats.n_distinct21 = c->h_stats.n_cand;
    }
    return P3_OK;
}

extern "C" {

const char *p3_last_error(void) { return g_err.c_str(); }
void p3_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }
int p3_version(void) { return 100; }
int p3_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void *p3_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void p3_host_free(void *p) { if (p) cudaFreeHost(p); }

p3_ctx *p3_create(int device, void *stream) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        fail(P3_ERR_CUDA, "p3_create: no CUDA device (this library has no CPU fallback)");
        return nullptr;
    }
    if (device < 0 || device >= n) { fail(P3_ERR_ARG, "p3_create: bad device index"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail(P3_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { fail(P3_ERR_CUDA, "cudaGetDeviceProperties failed"); return nullptr; }
    if (prop.major < 10) { fail(P3_ERR_CUDA, "p3_create: device is not sm_100 (B200) class"); return nullptr; }
    if (const char *g = getenv("P3_L2_FETCH_GRANULARITY")) {  // experiment knob: 32/64/128-byte L2 miss fetch
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
        cudaGetLastError();
    }
    p3_ctx *c = new p3_ctx();
    c->device = device;
    c->n_sm = prop.multiProcessorCount;
    if (stream) { c->stream = (cudaStream_t)stream; }
    else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; fail(P3_ERR_CUDA, "cudaStreamCreate failed"); return nullptr; }
        c->own_stream = true;
    }
    for (auto &e : c->ev) cudaEventCreate(&e);
    if (cudaMalloc(&c->d_stats, sizeof(Stats)) != cudaSuccess ||
        cudaMalloc(&c->d_ovf_keys, sizeof(uint64_t) * kOvfCap) != cudaSuccess ||
        cudaMalloc(&c->d_ovf_wraps, sizeof(unsigned long long) * kOvfCap) != cudaSuccess) {
        fail(P3_ERR_NOMEM, "p3_create: cudaMalloc failed"); delete c; return nullptr;
    }
    cudaMemsetAsync(c->d_stats, 0, sizeof(Stats), c->stream);
    memset(&c->h_stats, 0, sizeof(Stats));
    return c;
}

static void free_reads(p3_ctx *c) {
    dfree(c->own_packed); dfree(c->own_off); dfree(c->own_nmask); dfree(c->d_rend);
    c->cap_packed = c->cap_off = c->cap_nmask = c->cap_rend = 0;
    c->d_packed = nullptr; c->d_off = nullptr; c->d_nmask = nullptr; c->have_reads = false;
}
static void free_bf(p3_ctx *c) {
    dfree(c->d_good21); dfree(c->d_solid); dfree(c->d_set); dfree(c->d_list); dfree(c->d_bloom);
    dfree(c->d_seed); dfree(c->d_adj); dfree(c->d_hint); c->cap_hint = 0; c->adj_cap = 0; c->cap_planes = c->cap_seed = 0;
    c->nbs = 0; c->bloom_words = 0;
    c->have_bf = c->have_solid = c->have_adj = false;
}

void p3_destroy(p3_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    mg_release(c);
    long_release(c);
    bloom_release(c);
    free_reads(c); free_bf(c);
    dfree(c->d_table); dfree(c->d_proven2); dfree(c->d_bkeys); dfree(c->d_bword); dfree(c->d_bidx); dfree(c->d_below); dfree(c->d_valid);
    dfree(c->d_ghist); dfree(c->d_binmeta); dfree(c->d_cursor); dfree(c->d_ovf_keys); dfree(c->d_ovf_wraps); dfree(c->d_stats);
    for (auto &e : c->ev) cudaEventDestroy(e);
    for (auto &e : c->evpool) cudaEventDestroy(e);
    for (auto &e : c->up_ev) cudaEventDestroy(e);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int p3_synchronize(p3_ctx *c) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

static int finish_reads(p3_ctx *c) {
    // positions travel as a 32-bit word index + 5-bit offset: 2^32 words = 137 Gbases per context
    if (c->n_words >= (1ull << 32)) return fail(P3_ERR_ARG, "more than 2^32 packed words (137 Gbases) per context: split the reads over contexts");
    // read-end plane, built on the device from the offsets
    uint64_t plane_words = c->n_words + 1;
    CU(ensure(c->d_rend, c->cap_rend, sizeof(uint32_t) * plane_words));
    CU(cudaMemsetAsync(c->d_rend, 0, sizeof(uint32_t) * plane_words, c->stream));
    rend_kernel<<<c->grid(4), 256, 0, c->stream>>>(c->d_off, c->n_reads, c->total_bases, plane_words * 32, c->d_rend);
    c->launches++;
    CU(cudaGetLastError());
    c->have_reads = true; c->have_counts = false;
    c->have_bf = c->have_solid = c->have_adj = false;
    return P3_OK;
}

int p3_reads_upload(p3_ctx *c, const uint64_t *h_packed, uint64_t total_bases, const uint64_t *h_off,
                    uint64_t n_reads, const uint32_t *h_nmask) {
    if (!c || !h_packed || !h_off) return fail(P3_ERR_ARG, "p3_reads_upload: null argument");
    CU(cudaSetDevice(c->device));
    c->have_reads = false;
    c->total_bases = total_bases; c->n_reads = n_reads; c->n_words = (total_bases + 31) / 32;
    uint64_t pw = c->n_words + 1;
    CU(ensure(c->own_packed, c->cap_packed, sizeof(uint64_t) * pw));
    CU(ensure(c->own_off, c->cap_off, sizeof(uint64_t) * (n_reads + 1)));
    CU(cudaMemcpyAsync(c->own_packed, h_packed, sizeof(uint64_t) * pw, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->own_off, h_off, sizeof(uint64_t) * (n_reads + 1), cudaMemcpyHostToDevice, c->stream));
    if (h_nmask) {
        CU(ensure(c->own_nmask, c->cap_nmask, sizeof(uint32_t) * pw));
        CU(cudaMemcpyAsync(c->own_nmask, h_nmask, sizeof(uint32_t) * pw, cudaMemcpyHostToDevice, c->stream));
    }
    c->d_packed = c->own_packed; c->d_off = c->own_off; c->d_nmask = h_nmask ? c->own_nmask : nullptr;
    return finish_reads(c);
}

int p3_reads_attach(p3_ctx *c, const uint64_t *d_packed, uint64_t total_bases, const uint64_t *d_off,
                    uint64_t n_reads, const uint32_t *d_nmask) {
    if (!c || !d_packed || !d_off) return fail(P3_ERR_ARG, "p3_reads_attach: null argument");
    CU(cudaSetDevice(c->device));
    c->have_reads = false;
    c->total_bases = total_bases; c->n_reads = n_reads; c->n_words = (total_bases + 31) / 32;
    c->d_packed = d_packed; c->d_off = d_off; c->d_nmask = d_nmask;
    return finish_reads(c);
}

// ---- stage A -----------------------------------------------------------------------------------
static int count_direct(p3_ctx *c) {
    CU(ensure(c->d_proven2, c->cap_proven, sizeof(uint32_t) * (c->n_words + 1)));
    CU(cudaEventRecord(c->ev[0], c->stream));
    if (c->d_nmask)
        count21_kernel<true><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_rend, c->d_nmask, c->n_words, c->table(), c->ovf(), c->d_stats, c->d_proven2);
    else
        count21_kernel<false><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_rend, nullptr, c->n_words, c->table(), c->ovf(), c->d_stats, c->d_proven2);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev[1], c->stream));
    return P3_OK;
}

// the tile-sort kernels need > 48 KB of dynamic shared memory when they carry side arrays. The
// attribute is per DEVICE (per context), so it is set once per context, not once per process.
static int scatter_attrs(p3_ctx *c) {
    if (c->attrs_set) return P3_OK;
    const int s21 = (int)sizeof(ScatterSmem21), sP = (int)sizeof(ScatterSmemT<true, false>), sPL = (int)sizeof(ScatterSmemT<true, true>);
    const int sA4 = (int)sizeof(ScatterSmemT<true, false, 4>), sA1 = (int)sizeof(ScatterSmemT<true, false, 1>);
    CU(cudaFuncSetAttribute(scatter21_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, s21));
    CU(cudaFuncSetAttribute(scatter21_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, s21));
    CU(cudaFuncSetAttribute(scatter21_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, s21));
    CU(cudaFuncSetAttribute(scatter21_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, s21));
    CU(cudaFuncSetAttribute(scatter21_kernel<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s21));
    CU(cudaFuncSetAttribute(scatter21_kernel<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s21));
    CU(cudaFuncSetAttribute(scatter21_kernel<true, 1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s21));
    CU(cudaFuncSetAttribute(scatter21_kernel<false, 1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s21));
    CU(cudaFuncSetAttribute(scatter_kmer_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sPL));
    CU(cudaFuncSetAttribute(scatter_kmer_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sPL));
    CU(cudaFuncSetAttribute(scatter_kmer_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sPL));
    CU(cudaFuncSetAttribute(scatter_kmer_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sPL));
    CU(cudaFuncSetAttribute(scatter_rec_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sA4));
    CU(cudaFuncSetAttribute(scatter_rec_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sP));
    CU(cudaFuncSetAttribute(scatter_rec_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sP));
    CU(cudaFuncSetAttribute(scatter_rec_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sA1));
    c->attrs_set = true;
    return P3_OK;
}
static int hist_buffers(p3_ctx *c) {
    if (!c->d_ghist) {
        CU(cudaMalloc(&c->d_ghist, sizeof(unsigned long long) * (kMaxParts + 1)));
        CU(cudaMalloc(&c->d_cursor, sizeof(unsigned long long) * (kMaxParts + 1)));
    }
    return scatter_attrs(c);
}
static cudaEvent_t pool_event(p3_ctx *c, size_t i) {
    while (c->evpool.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); c->evpool.push_back(e); }
    return c->evpool[i];
}
constexpr size_t kSmem21 = sizeof(ScatterSmem21), kSmemP = sizeof(ScatterSmemT<true, false>), kSmemPL = sizeof(ScatterSmemT<true, true>);
constexpr size_t kSmemA4 = sizeof(ScatterSmemT<true, false, 4>), kSmemA1 = sizeof(ScatterSmemT<true, false, 1>);

extern "C++" {
template <bool HAS_MASK>
static void launch_scatter21_local(p3_ctx *c, unsigned sblocks, uint64_t w0, uint64_t w1, uint32_t P, uint64_t cap) {
    scatter21_kernel<HAS_MASK, 0><<<sblocks, kScatterThreads, kSmem21, c->stream>>>(
        c->d_packed, c->d_rend, HAS_MASK ? c->d_nmask : nullptr, w0, w1, P, c->d_cursor, c->d_bkeys, c->d_bword, c->d_valid, 0, c->d_stats, PeerOut(), cap);
}
}  // extern "C++"
// bidx (optional): index stream to use instead of c->d_bidx; first: where the sweep starts in the bins' index space
// (multi-GPU key-range rounds: the bins of partitions [p_lo, p_hi) only, addressed with their absolute indices)
static int launch_insert_bins(p3_ctx *c, const uint64_t *keys, uint64_t n, uint64_t cap,
                              const unsigned long long *bin_end, const unsigned long long *n_dev,
                              uint32_t *bidx = nullptr, unsigned long long first = 0) {
    if (!bidx) bidx = c->d_bidx;
    auto set_work = [&]() -> cudaError_t {
        return first ? cudaMemcpyAsync(&c->d_stats->work, &first, sizeof(first), cudaMemcpyHostToDevice, c->stream)
                     : cudaMemsetAsync(&c->d_stats->work, 0, sizeof(unsigned long long), c->stream);
    };
    CU(set_work());
    if (c->probe_stats) {
        if (cap) insert_find_kernel<true, false><<<c->grid(8), 256, 0, c->stream>>>(keys, n, c->table(), c->d_stats, bidx, cap, bin_end, n_dev);
        else insert_find_kernel<true, true><<<c->grid(8), 256, 0, c->stream>>>(keys, n, c->table(), c->d_stats, bidx, cap, bin_end, n_dev);
    } else if (cap) {
        static const int variant = getenv("P3_FIND_BATCH") ? atoi(getenv("P3_FIND_BATCH")) : kSweepBatch;   // experiment knob
        if (variant == 8) insert_find_kernel<false, false, 8><<<c->grid(2), 256, 0, c->stream>>>(keys, n, c->table(), c->d_stats, bidx, cap, bin_end, n_dev);
        else if (variant == 2) insert_find_kernel<false, false, 2><<<c->grid(6), 256, 0, c->stream>>>(keys, n, c->table(), c->d_stats, bidx, cap, bin_end, n_dev);
        else if (variant == 4) insert_find_kernel<false, false, 4><<<c->grid(4), 256, 0, c->stream>>>(keys, n, c->table(), c->d_stats, bidx, cap, bin_end, n_dev);
        else insert_find_kernel<false, false><<<c->grid(8), 256, 0, c->stream>>>(keys, n, c->table(), c->d_stats, bidx, cap, bin_end, n_dev);
    } else {
        insert_find_kernel<false, true><<<c->grid(8), 256, 0, c->stream>>>(keys, n, c->table(), c->d_stats, bidx, cap, bin_end, n_dev);
    }
    CU(set_work());
    if (cap) insert_add_kernel<false><<<c->grid(8), 256, 0, c->stream>>>(bidx, keys, n, c->table(), c->ovf(), c->d_stats, cap, bin_end, n_dev);
    else insert_add_kernel<true><<<c->grid(8), 256, 0, c->stream>>>(bidx, keys, n, c->table(), c->ovf(), c->d_stats, cap, bin_end, n_dev);
    c->launches += 2;
    CU(cudaGetLastError());
    return P3_OK;
}

// Plan of the binned count: how many words per chunk so that the bins (12 B per record) fit, and the
// fixed bin capacity. exact = false: fixed-capacity bins — the partition of a key is a hash, so a chunk's
// records spread evenly over the partitions; each gets room for its expected share + 3 % + 8192 and no
// histogram pass is needed. A partition that overflows anyway (heavy-hitter keys: satellites,
// homopolymers) raises Stats::err_bin_overflow on the device; the caller then redoes the stage with exact
// (histogram-sized) bins. Nothing inside the chunk loop waits for the device.
struct BinPlan { uint64_t chunk_words = 0; double pos_per_word = 32; uint32_t P = 1; bool exact = false; };
static uint64_t plan_fixed_cap(const BinPlan &pl, uint64_t words) {
    uint64_t share = (uint64_t)((double)words * pl.pos_per_word / (double)pl.P * 1.03) + 8192;
    return (share + kSweepChunk - 1) / kSweepChunk * kSweepChunk;
}
static int plan_bins(p3_ctx *c, uint64_t upper, bool exact, BinPlan *pl) {
    int rc0 = hist_buffers(c);
    if (rc0) return rc0;
    if (!c->d_binmeta) CU(cudaMalloc(&c->d_binmeta, sizeof(unsigned long long) * (kMaxParts + 2)));
    pl->P = c->parts; pl->exact = exact;
    pl->pos_per_word = std::min(32.0, c->n_words ? (double)upper / (double)c->n_words : 32.0);
    auto records_for = [&](uint64_t words) -> uint64_t {
        return std::max<uint64_t>(words * 32, exact ? 0 : plan_fixed_cap(*pl, words) * pl->P);
    };
    // the device is only asked for its free memory when the bins have to grow
    const uint64_t all_words = std::max<uint64_t>((c->n_words + kTileWords - 1) / kTileWords * kTileWords, kTileWords);
    uint64_t chunk_words = all_words;
    const char *benv = getenv("P3_BIN_BUDGET_BYTES");
    if (benv || records_for(std::min(all_words, c->n_words)) * 12 > c->cap_bkeys + c->cap_bword) {
        uint64_t budget;
        if (benv) budget = strtoull(benv, nullptr, 10);
        else {
            size_t fr = 0, tot = 0;
            CU(cudaMemGetInfo(&fr, &tot));
            budget = (uint64_t)(0.7 * (double)(fr + c->cap_bkeys + c->cap_bword));
        }
        chunk_words = std::max<uint64_t>(budget / (13 * 32), kTileWords) / kTileWords * kTileWords;
        chunk_words = std::min<uint64_t>(std::max<uint64_t>(chunk_words, kTileWords), all_words);
    }
    pl->chunk_words = chunk_words;
    const uint64_t rec_cap = records_for(std::min(chunk_words, c->n_words));
    CU(ensure(c->d_bkeys, c->cap_bkeys, sizeof(uint64_t) * rec_cap));
    CU(ensure(c->d_bword, c->cap_bword, sizeof(uint32_t) * rec_cap));
    CU(ensure(c->d_bidx, c->cap_bidx, sizeof(uint32_t) * rec_cap));
    CU(ensure(c->d_valid, c->cap_valid, sizeof(uint32_t) * (c->n_words + 1)));
    return P3_OK;
}
// bins the 21-mers of words [w0, w1) by table partition into d_bkeys / d_bword; the bin ends (and, for
// exact bins, the record total) are copied to d_binmeta, which no later kernel touches
static int bin_chunk(p3_ctx *c, const BinPlan &pl, uint64_t w0, uint64_t w1, cudaEvent_t ev_hist_done) {
    const uint32_t P = pl.P;
    const unsigned sblocks = (unsigned)std::min<uint64_t>((w1 - w0 + kTileWords - 1) / kTileWords, (uint64_t)c->n_sm * 3);
    const uint64_t cap = pl.exact ? 0 : plan_fixed_cap(pl, w1 - w0);
    if (cap && c->up_pieces) {
        // the reads are still arriving (p3_assemble_hot_path): bin piece by piece as the uploads complete, into the same bins
        init_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, cap);
        if (ev_hist_done) CU(cudaEventRecord(ev_hist_done, c->stream));
        for (uint64_t i = 0; i < c->up_pieces; i++) {
            CU(cudaStreamWaitEvent(c->stream, c->up_ev[i], 0));
            const uint64_t a = std::max<uint64_t>(w0, i * c->up_piece_words), b = std::min<uint64_t>(w1, (i + 1) * c->up_piece_words);
            if (b <= a) continue;
            const unsigned sb = (unsigned)std::min<uint64_t>((b - a + kTileWords - 1) / kTileWords, (uint64_t)c->n_sm * 3);
            if (c->d_nmask) launch_scatter21_local<true>(c, sb, a, b, P, cap);
            else launch_scatter21_local<false>(c, sb, a, b, P, cap);
            c->launches++;
        }
        c->up_pieces = 0;
        check_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, cap, c->d_stats);
    } else if (cap) {
        init_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, cap);
        if (ev_hist_done) CU(cudaEventRecord(ev_hist_done, c->stream));
        if (c->d_nmask) launch_scatter21_local<true>(c, sblocks, w0, w1, P, cap);
        else launch_scatter21_local<false>(c, sblocks, w0, w1, P, cap);
        check_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, cap, c->d_stats);
    } else {
        CU(cudaMemsetAsync(c->d_ghist, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
        if (c->d_nmask) hist21_kernel<true, 0><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_rend, c->d_nmask, w0, w1, P, c->d_ghist);
        else hist21_kernel<false, 0><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_rend, nullptr, w0, w1, P, c->d_ghist);
        scan_parts_kernel<<<1, 256, 0, c->stream>>>(c->d_ghist, P, c->d_cursor, c->d_ghist + kMaxParts);
        if (ev_hist_done) CU(cudaEventRecord(ev_hist_done, c->stream));
        if (c->d_nmask) launch_scatter21_local<true>(c, sblocks, w0, w1, P, 0);
        else launch_scatter21_local<false>(c, sblocks, w0, w1, P, 0);
        CU(cudaMemcpyAsync(c->d_binmeta + kMaxParts, c->d_ghist + kMaxParts, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaMemcpyAsync(c->d_binmeta, c->d_cursor, sizeof(unsigned long long) * P, cudaMemcpyDeviceToDevice, c->stream));
    c->launches += 3;
    CU(cudaGetLastError());
    c->bin_cap = cap;
    c->bin_n = cap ? (uint64_t)P * cap : (w1 - w0) * 32;
    return P3_OK;
}
// the main stream waits for every piece of a pending upload (paths that cannot consume it piece by piece)
static void upload_wait_all(p3_ctx *c) {
    for (uint64_t i = 0; i < c->up_pieces; i++) cudaStreamWaitEvent(c->stream, c->up_ev[i], 0);
    c->up_pieces = 0;
}
static int count_binned(p3_ctx *c, uint64_t upper, bool exact) {
    BinPlan pl;
    int rc = plan_bins(c, upper, exact, &pl);
    if (rc) return rc;
    if (c->up_pieces && pl.chunk_words < c->n_words) upload_wait_all(c);   // several chunks: no piecewise binning
    c->n_chunks = 0; c->bins_valid = false;
    for (int i = 0; i < 4; i++) c->ms_sub[i] = 0;
    CU(cudaEventRecord(c->ev[0], c->stream));
    for (uint64_t w0 = 0; w0 < c->n_words; w0 += pl.chunk_words) {
        const uint64_t w1 = std::min<uint64_t>(w0 + pl.chunk_words, c->n_words);
        const size_t e = 4 * (size_t)c->n_chunks;
        CU(cudaEventRecord(pool_event(c, e), c->stream));
        rc = bin_chunk(c, pl, w0, w1, pool_event(c, e + 1));
        if (rc) return rc;
        CU(cudaEventRecord(pool_event(c, e + 2), c->stream));
        rc = launch_insert_bins(c, c->d_bkeys, c->bin_n, c->bin_cap, c->d_binmeta, c->bin_cap ? nullptr : c->d_binmeta + kMaxParts);
        if (rc) return rc;
        CU(cudaEventRecord(pool_event(c, e + 3), c->stream));
        c->n_chunks++;
    }
    c->bins_valid = c->n_chunks == 1;   // the verdict sweep of MakeBF can reuse them
    c->bin_upper = upper; c->bin_exact = exact;
    CU(cudaEventRecord(c->ev[1], c->stream));
    return P3_OK;
}
// sub-stage times of the chunks of the last count_binned (after the stream has been synchronised)
static void count_binned_times(p3_ctx *c) {
    for (uint64_t ci = 0; ci < c->n_chunks && 4 * ci + 3 < c->evpool.size(); ci++) {
        float a = 0, b = 0, d = 0;
        cudaEventElapsedTime(&a, c->evpool[4 * ci], c->evpool[4 * ci + 1]);
        cudaEventElapsedTime(&b, c->evpool[4 * ci + 1], c->evpool[4 * ci + 2]);
        cudaEventElapsedTime(&d, c->evpool[4 * ci + 2], c->evpool[4 * ci + 3]);
        c->ms_sub[0] += a; c->ms_sub[1] += b; c->ms_sub[2] += d;
    }
}

// allocate (or reuse) and clear the count table: partitions of ~24 MB each in binned mode so that
// one partition plus the streaming bins stay in L2
// partitions of the binned count table for a given capacity (also what p3_table_partitions tells the multi-GPU driver)
static uint32_t table_partitions(uint64_t table_slots) {
    // ~48 MB per partition: the insert sweep is insensitive to the partition size between 17 and 50 MB (measured), while
    // every tile sort pays a scan + one global claim per partition and tile (scatter21: 33 / 42 / 53 / 72 ms at 350 / 520 /
    // 696 / 1000 partitions)
    uint64_t part_bytes = 48ull << 20;
    if (const char *e = getenv("P3_PART_MB")) part_bytes = std::max<uint64_t>(strtoull(e, nullptr, 10), 1) << 20;
    uint64_t want = (table_slots * 8 + part_bytes - 1) / part_bytes;
    if (const char *e = getenv("P3_PARTS")) want = strtoull(e, nullptr, 10);
    return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 1), kMaxParts);
}
static int setup_table(p3_ctx *c, uint64_t table_slots) {
    uint32_t P = c->binned ? table_partitions(table_slots) : 1;
    uint64_t nbp = ((table_slots + 3) / 4 + P - 1) / P;
    if (nbp == 0) nbp = 1;
    if (nbp >= (1ull << 32)) return fail(P3_ERR_ARG, "count table partition too large");
    uint64_t nb = nbp * P;
    if (!c->d_table || c->nb != nb) {
        dfree(c->d_table);
        if (cudaMalloc(&c->d_table, nb * 32) != cudaSuccess) { cudaGetLastError(); return fail(P3_ERR_NOMEM, "count table allocation failed"); }
        c->nb = nb;
    }
    c->parts = P; c->nbp = nbp;
    c->probe_stats = getenv("P3_PROBE_STATS") != nullptr;
    CU(cudaMemsetAsync(c->d_table, 0xFF, nb * 32, c->stream));
    CU(cudaMemsetAsync(c->d_ovf_keys, 0xFF, sizeof(uint64_t) * kOvfCap, c->stream));
    CU(cudaMemsetAsync(c->d_ovf_wraps, 0, sizeof(unsigned long long) * kOvfCap, c->stream));
    CU(cudaMemsetAsync(c->d_stats, 0, sizeof(Stats), c->stream));
    memset(&c->h_stats, 0, sizeof(Stats));
    return P3_OK;
}

int p3_count_short_kmers(p3_ctx *c, uint64_t table_slots) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    if (!c->have_reads) return fail(P3_ERR_STATE, "p3_count_short_kmers: no reads attached");
    CU(cudaSetDevice(c->device));
    const char *mode = getenv("P3_COUNT_MODE");
    c->binned = !(mode && strcmp(mode, "direct") == 0);
    uint64_t upper = c->total_bases > (kShortK - 1) * c->n_reads ? c->total_bases - (kShortK - 1) * c->n_reads : 0;
    if (table_slots == 0) {
        table_slots = std::max<uint64_t>(2 * upper, 1024);
        size_t fr = 0, tot = 0;
        CU(cudaMemGetInfo(&fr, &tot));
        uint64_t lim = (uint64_t)(0.25 * (double)(fr + (c->d_table ? c->nb * 32 : 0))) / 8;
        if (table_slots > lim) table_slots = lim;
    }
    int rc = setup_table(c, table_slots);
    if (rc) return rc;
    c->pos_on_host = false;
    bool exact = getenv("P3_EXACT_BINS") != nullptr;
    if (c->up_pieces && (!c->binned || exact)) upload_wait_all(c);
    rc = c->binned ? count_binned(c, upper, exact) : count_direct(c);
    if (rc) return rc;
    rc = pull_stats(c);
    if (rc) return rc;
    if (c->binned && c->h_stats.err_bin_overflow && !exact) {   // a heavy-hitter partition overflowed its fixed share: exact bins
        rc = setup_table(c, table_slots);
        if (!rc) rc = count_binned(c, upper, true);
        if (!rc) rc = pull_stats(c);
        if (rc) return rc;
    }
    if (c->binned) count_binned_times(c);
    c->count_probes = c->h_stats.probes; c->count_max_probe = c->h_stats.max_probe;
    CU(cudaEventElapsedTime(&c->ms[0], c->ev[0], c->ev[1]));
    if (c->h_stats.err_table_full) return fail(P3_ERR_TABLE_FULL, "21-mer count table full: raise table_slots");
    if (c->h_stats.err_ovf_full) return fail(P3_ERR_TABLE_FULL, "count overflow side table full");
    c->have_counts = true;
    c->have_bf = c->have_solid = c->have_adj = false;  // buffers are kept for reuse
    return P3_OK;
}

int p3_short_kmer_stats(p3_ctx *c, uint64_t *n_positions, uint64_t *n_distinct) {
    if (!c || !c->have_counts) return fail(P3_ERR_STATE, "no counts");
    if (n_positions) *n_positions = c->h_stats.n_pos21;
    if (n_distinct) *n_distinct = c->h_stats.n_distinct21;
    return P3_OK;
}

int p3_short_kmer_export(p3_ctx *c, uint64_t *h_keys, uint64_t *h_counts, uint64_t cap, uint64_t *n) {
    if (!c || !c->have_counts || !h_keys || !h_counts) return fail(P3_ERR_STATE, "p3_short_kmer_export: no counts / null output");
    CU(cudaSetDevice(c->device));
    uint64_t nd = c->h_stats.n_distinct21;
    if (n) *n = nd;
    if (cap < nd) return fail(P3_ERR_ARG, "p3_short_kmer_export: capacity too small");
    if (nd == 0) return P3_OK;
    uint64_t *dk = nullptr, *dc = nullptr;
    TmpFree tmp; tmp.own(&dk); tmp.own(&dc);
    CU(cudaMalloc(&dk, sizeof(uint64_t) * nd));
    CU(cudaMalloc(&dc, sizeof(uint64_t) * nd));
    CU(cudaMemsetAsync(&c->d_stats->n_export, 0, sizeof(unsigned long long), c->stream));
    export_counts_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_table, c->nb * 4, c->ovf(), c->d_stats, 0, dk, dc, nd);
    c->launches++;
    CU(cudaMemcpyAsync(h_keys, dk, sizeof(uint64_t) * nd, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(h_counts, dc, sizeof(uint64_t) * nd, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

int p3_short_kmer_lookup(p3_ctx *c, const uint64_t *h_keys, uint64_t n, uint64_t *h_counts) {
    if (!c || !c->have_counts) return fail(P3_ERR_STATE, "no counts");
    if (n == 0) return P3_OK;
    CU(cudaSetDevice(c->device));
    uint64_t *dk = nullptr, *dc = nullptr;
    TmpFree tmp; tmp.own(&dk); tmp.own(&dc);
    CU(cudaMalloc(&dk, sizeof(uint64_t) * n));
    CU(cudaMalloc(&dc, sizeof(uint64_t) * n));
    CU(cudaMemcpyAsync(dk, h_keys, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, c->stream));
    lookup_counts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->table(), c->ovf(), c->d_stats, dk, n, dc);
    c->launches++;
    CU(cudaMemcpyAsync(h_counts, dc, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

// ---- stage B -----------------------------------------------------------------------------------
static int alloc_bloom(p3_ctx *c, uint32_t k, uint64_t filter_size, uint32_t num_hashes, uint64_t min_words = 0) {
    if (k < P3_MIN_K || k > P3_MAX_K) return fail(P3_ERR_ARG, "k outside [21,3001] is not supported");
    if (filter_size == 0) return fail(P3_ERR_ARG, "filter_size == 0 (the reference divides by zero here)");
    if (num_hashes > 255) return fail(P3_ERR_ARG, "num_hashes > 255 (uint8_t in the reference)");
    uint64_t words = std::max<uint64_t>((filter_size + 31) / 32, min_words);   // min_words: room for whole shards (multi-GPU)
    if (!c->d_bloom || c->bloom_words != words) {
        dfree(c->d_bloom);
        if (cudaMalloc(&c->d_bloom, sizeof(uint32_t) * words) != cudaSuccess) { cudaGetLastError(); return fail(P3_ERR_NOMEM, "bloom allocation failed"); }
        c->bloom_words = words;
    }
    c->k = k; c->filter_size = filter_size; c->num_hashes = num_hashes;
    return P3_OK;
}

// distinct canonical k-mers of the solid positions -> c->d_set / c->d_list (grows on overflow);
// leaves n_distinct_solid in h_stats
static int dedupe_solid_positions(p3_ctx *c, uint32_t k, uint64_t solid_slots) {
    int rc0 = hist_buffers(c);
    if (rc0) return rc0;
    bool allow_binned = true;
    for (int attempt = 0;; attempt++) {
        // partitions of ~24 MB (one stays L2 resident under the binned sweep)
        uint64_t buckets = (solid_slots + 3) / 4;
        uint64_t want = (buckets * 32 + (24ull << 20) - 1) / (24ull << 20);
        if (const char *e = getenv("P3_SET_PARTS")) want = strtoull(e, nullptr, 10);
        // At most 96 partitions: the tile sort of 4096 positions needs runs of a few dozen records per
        // partition to write coalesced (measured at configs[1]: 77 partitions 95 ms, 128 partitions 127 ms).
        // A larger set is left unpartitioned and filled directly.
        uint32_t P = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 1), kMaxParts);
        if (!getenv("P3_SET_PARTS") && P > 96) P = 1;
        uint64_t nbp = std::max<uint64_t>((buckets + P - 1) / P, 1);
        if (nbp >= (1ull << 32)) return fail(P3_ERR_ARG, "solid k-mer set partition too large");
        uint64_t nbs = nbp * P;
        if (!c->d_set || c->nbs != nbs) {
            dfree(c->d_set); dfree(c->d_list);
            if (cudaMalloc(&c->d_set, nbs * 32) != cudaSuccess || cudaMalloc(&c->d_list, nbs * 32) != cudaSuccess) {
                cudaGetLastError();
                return fail(P3_ERR_NOMEM, "solid k-mer set allocation failed");
            }
            c->nbs = nbs; c->list_cap = nbs * 4;
        }
        c->set_parts = P;
        CU(cudaMemsetAsync(c->d_set, 0xFF, nbs * 32, c->stream));
        CU(cudaMemsetAsync(&c->d_stats->n_distinct_solid, 0, sizeof(unsigned long long), c->stream));
        CU(cudaMemsetAsync(&c->d_stats->err_table_full, 0, sizeof(unsigned), c->stream));
        CU(cudaMemsetAsync(&c->d_stats->err_kbin_overflow, 0, sizeof(unsigned), c->stream));
        // binned path: needs room for n_adds 8-byte records (+5 %) in the idle count-stage bins
        const uint64_t n_occ = c->h_stats.n_adds;
        bool binned = false;
        const char *force = getenv("P3_DEDUPE_BINNED");
        if (allow_binned && (force ? atoi(force) != 0 : (P >= 2 && n_occ >= (1u << 22)))) {
            uint64_t cap = (uint64_t)((double)n_occ / (double)P * 1.05) + 8192;
            cap = (cap + kSweepChunk - 1) / kSweepChunk * kSweepChunk;
            const uint64_t need = sizeof(uint64_t) * cap * P;
            bool fits = c->cap_bkeys >= need;
            if (!fits) {
                size_t fr = 0, tot = 0;
                CU(cudaMemGetInfo(&fr, &tot));
                fits = need < (uint64_t)(0.5 * (double)(fr + c->cap_bkeys));
            }
            if (fits) {
                c->bins_valid = false;                                          // the count bins become the k-mer bins
                CU(ensure(c->d_bkeys, c->cap_bkeys, need));
                CU(ensure(c->d_bword, c->cap_bword, cap * P));                 // one hint byte per binned k-mer
                CU(ensure(c->d_hint, c->cap_hint, nbs * 4));                   // one hint byte per set slot
                CU(cudaMemsetAsync(c->d_hint, 0, nbs * 4, c->stream));
                CU(cudaMemsetAsync(c->d_solid + c->n_words, 0, sizeof(uint32_t), c->stream));   // the hint of the last word looks one word ahead
                init_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, cap);
                unsigned sblocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>((c->n_words + kTileWords - 1) / kTileWords, 1), (uint64_t)c->n_sm * 3);
                uint8_t *hb = reinterpret_cast<uint8_t *>(c->d_bword);
                if (c->d_nmask) scatter_kmer_kernel<true, false><<<sblocks, kScatterThreads, kSmemPL, c->stream>>>(c->d_packed, c->d_nmask, c->d_solid, 0, c->n_words, (int)k, P, c->d_cursor, c->d_bkeys, hb, cap, c->d_stats);
                else scatter_kmer_kernel<false, false><<<sblocks, kScatterThreads, kSmemPL, c->stream>>>(c->d_packed, nullptr, c->d_solid, 0, c->n_words, (int)k, P, c->d_cursor, c->d_bkeys, hb, cap, c->d_stats);
                // a heavy-hitter partition that outgrew its bin dropped records: err_kbin_overflow, seen below
                CU(cudaMemsetAsync(&c->d_stats->work, 0, sizeof(unsigned long long), c->stream));
                set_sweep_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_bkeys, hb, (uint64_t)P * cap, cap, c->d_cursor, c->kset(), c->d_hint, c->d_stats);
                c->launches += 3;
                CU(cudaGetLastError());
                binned = true;
            }
        }
        if (!binned) {
            if (c->d_nmask)
                makebf_kernel<true><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_nmask, c->d_solid, c->n_words, (int)k, c->kset(), c->d_stats);
            else
                makebf_kernel<false><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, nullptr, c->d_solid, c->n_words, (int)k, c->kset(), c->d_stats);
            c->launches++;
        }
        if (binned && (!c->d_adj || c->adj_cap < c->list_cap)) {   // the adjacency bytes start as the hints
            dfree(c->d_adj);
            c->adj_cap = c->list_cap;
            CU(cudaMalloc(&c->d_adj, c->adj_cap));
        }
        compact_set_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_set, nbs * 4, c->d_list, c->list_cap, c->d_stats,
                                                             binned ? reinterpret_cast<const uint8_t *>(c->d_hint) : nullptr, c->d_adj);
        c->launches++;
        CU(cudaGetLastError());
        int rc = pull_stats(c);
        if (rc) return rc;
        if (binned && c->h_stats.err_kbin_overflow) { allow_binned = false; continue; }   // direct path, same capacity
        c->hints_valid = binned;
        if (!c->h_stats.err_table_full) return P3_OK;
        if (attempt >= 16) return fail(P3_ERR_TABLE_FULL, "solid k-mer set full after growing");
        solid_slots = std::max<uint64_t>(solid_slots * 4, 1024);
    }
}

// dense BF.add over the first nd k-mers of c->d_list: binned by filter segment (p3_bloom.inc.cu);
// tiny jobs and single-segment filters take the direct kernel, one pass per L2-sized segment
static int bloom_add_direct(p3_ctx *c, uint64_t nd) {
    if (c->k > 32) return bloom_add_direct_long(c, nd);
    uint64_t seg_bits = 40ull << 23;                       // 40 MB of filter per pass
    uint64_t n_seg = (c->filter_size + seg_bits - 1) / seg_bits;
    if (n_seg > 16 || nd * c->num_hashes < (1u << 22)) { n_seg = 1; seg_bits = c->filter_size; }   // huge filter / tiny job: one pass
    else seg_bits = ((c->filter_size + n_seg - 1) / n_seg + 31) / 32 * 32;
    for (uint64_t sg = 0; sg < n_seg && nd; sg++) {
        bloom_list_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_list, nd, c->bloom(), sg * seg_bits, std::min<uint64_t>((sg + 1) * seg_bits, c->filter_size));
        c->launches++;
    }
    CU(cudaGetLastError());
    return P3_OK;
}
static int bloom_add_list(p3_ctx *c, uint64_t nd) {
    bool done = false;
    int rcb = bloom_add_binned(c, nd, &done);
    if (rcb || done) return rcb;
    return bloom_add_direct(c, nd);
}

int p3_make_bf(p3_ctx *c, uint32_t k, uint64_t filter_size, uint32_t num_hashes, uint32_t cov_threshold,
               uint64_t solid_slots) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    if (!c->have_counts) return fail(P3_ERR_STATE, "p3_make_bf: run p3_count_short_kmers first");
    CU(cudaSetDevice(c->device));
    c->set_valid = false; c->hints_valid = false;
    int rc = alloc_bloom(c, k, filter_size, num_hashes);
    if (rc) return rc;
    uint64_t pw = c->n_words + 1;
    if (!c->d_good21 || !c->d_solid || c->cap_planes < sizeof(uint32_t) * pw) {
        dfree(c->d_good21); dfree(c->d_solid);
        CU(cudaMalloc(&c->d_good21, sizeof(uint32_t) * pw));
        CU(cudaMalloc(&c->d_solid, sizeof(uint32_t) * pw));
        c->cap_planes = sizeof(uint32_t) * pw;
    }
    CU(ensure(c->d_seed, c->cap_seed, sizeof(int64_t) * std::max<uint64_t>(c->n_reads, 1)));
    // reset the stage-B part of the stats
    CU(cudaMemsetAsync(&c->d_stats->n_adds, 0, sizeof(unsigned long long) * 5, c->stream));
    CU(cudaMemsetAsync(c->d_good21 + c->n_words, 0, sizeof(uint32_t), c->stream));

    // B1: coverage flags
    CU(cudaMemsetAsync(&c->d_stats->err_bin_overflow, 0, sizeof(unsigned), c->stream));
    CU(cudaEventRecord(c->ev[2], c->stream));
    bool clears_binned = false;
    if (c->binned) {
        rc = verdict_sweep(c, cov_threshold, false, &clears_binned);
        if (rc) return rc;
    } else {
        const uint32_t *proven = cov_threshold == 2 ? c->d_proven2 : nullptr;
        if (c->d_nmask)
            flags21_kernel<true><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_rend, c->d_nmask, c->n_words, c->table(), c->ovf(), c->d_stats, cov_threshold, proven, c->d_good21);
        else
            flags21_kernel<false><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_rend, nullptr, c->n_words, c->table(), c->ovf(), c->d_stats, cov_threshold, proven, c->d_good21);
        c->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev[3], c->stream));

    // B2a: solid plane + n_adds; number of distinct good 21-mers sizes the solid set
    auto solid_pass = [&]() -> int {
        CU(cudaMemsetAsync(&c->d_stats->n_adds, 0, sizeof(unsigned long long), c->stream));
        if (k > 32) return P3_OK;   // the multi-word path builds its own solid plane (p3_long.inc.cu)
        solid_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_good21, c->n_words, (int)k, c->d_solid, c->d_stats);
        c->launches++;
        return P3_OK;
    };
    rc = solid_pass();
    if (rc) return rc;
    if (solid_slots == 0) {
        export_counts_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_table, c->nb * 4, c->ovf(), c->d_stats, cov_threshold, nullptr, nullptr, 0);
        c->launches++;
    }
    rc = pull_stats(c);
    if (rc) return rc;
    if (clears_binned && c->h_stats.err_bin_overflow) {
        // a plane segment outgrew its bin (positions of count-1 keys bunched up): clear directly — clearing a
        // bit twice is harmless — and redo the solid plane
        CU(cudaMemsetAsync(&c->d_stats->err_bin_overflow, 0, sizeof(unsigned), c->stream));
        rc = verdict_sweep(c, cov_threshold, true, &clears_binned);
        if (rc) return rc;
        rc = solid_pass();
        if (!rc) rc = pull_stats(c);
        if (rc) return rc;
    }

    if (k > 32) {   // multi-word k-mers: p3_long.inc.cu
        rc = make_bf_long(c, k, solid_slots, (uint64_t)(1.25 * (double)c->h_stats.n_good21) + 1024);
        if (rc) return rc;
        CU(cudaEventElapsedTime(&c->ms_bloom, c->ev[14], c->ev[15]));
        CU(cudaEventElapsedTime(&c->ms[1], c->ev[2], c->ev[3]));
        CU(cudaEventElapsedTime(&c->ms[2], c->ev[4], c->ev[5]));
        CU(cudaEventElapsedTime(&c->ms[3], c->ev[5], c->ev[6]));
        c->have_bf = true; c->have_solid = true; c->have_adj = false; c->set_valid = false;
        c->d_set_b = nullptr; c->nbs_b = 0; c->parts_b = 1;
        return P3_OK;
    }
    if (solid_slots == 0) {
        uint64_t est = std::min<uint64_t>(c->h_stats.n_adds, (uint64_t)(1.25 * (double)c->h_stats.n_good21) + 1024);
        solid_slots = std::max<uint64_t>(2 * est, 1024);
    }
    // B2b: de-duplicate the solid positions into the set/list, then dense BF.add over the list
    CU(cudaEventRecord(c->ev[4], c->stream));
    rc = dedupe_solid_positions(c, k, solid_slots);
    if (rc) return rc;
    CU(cudaMemsetAsync(c->d_bloom, 0, sizeof(uint32_t) * c->bloom_words, c->stream));
    CU(cudaEventRecord(c->ev[14], c->stream));
    rc = bloom_add_list(c, c->h_stats.n_distinct_solid);
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
        rc = bloom_add_direct(c, c->h_stats.n_distinct_solid);
        if (!rc) rc = pull_stats(c);
        if (rc) return rc;
    }
    CU(cudaEventElapsedTime(&c->ms_bloom, c->ev[14], c->ev[15]));
    CU(cudaEventElapsedTime(&c->ms[1], c->ev[2], c->ev[3]));
    CU(cudaEventElapsedTime(&c->ms[2], c->ev[4], c->ev[5]));
    CU(cudaEventElapsedTime(&c->ms[3], c->ev[5], c->ev[6]));
    c->have_bf = true; c->have_solid = true; c->have_adj = false; c->set_valid = true;
    c->d_set_b = nullptr; c->nbs_b = 0; c->parts_b = 1;
    return P3_OK;
}

int p3_make_bf_stats(p3_ctx *c, uint64_t *n_adds, uint64_t *n_distinct_solid) {
    if (!c || !c->have_solid) return fail(P3_ERR_STATE, "no make_bf result");
    if (n_adds) *n_adds = c->h_stats.n_adds;
    if (n_distinct_solid) *n_distinct_solid = c->h_stats.n_distinct_solid;
    return P3_OK;
}

int p3_bf_export(p3_ctx *c, uint8_t *h_bits) {
    if (!c || !c->have_bf || !h_bits) return fail(P3_ERR_STATE, "p3_bf_export: no filter");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(h_bits, c->d_bloom, (c->filter_size + 7) / 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

int p3_bf_import(p3_ctx *c, uint32_t k, uint64_t filter_size, uint32_t num_hashes, const uint8_t *h_bits) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    int rc = alloc_bloom(c, k, filter_size, num_hashes);
    if (rc) return rc;
    c->set_valid = false; c->hints_valid = false; c->have_solid = false; c->have_adj = false;
    CU(cudaMemsetAsync(c->d_bloom, 0, sizeof(uint32_t) * c->bloom_words, c->stream));
    if (h_bits) CU(cudaMemcpyAsync(c->d_bloom, h_bits, (filter_size + 7) / 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->have_bf = true;
    return P3_OK;
}

int p3_seed_export(p3_ctx *c, int64_t *h_seed_pos) {
    if (!c || !c->have_solid || !h_seed_pos) return fail(P3_ERR_STATE, "p3_seed_export: no make_bf result");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(h_seed_pos, c->d_seed, sizeof(int64_t) * c->n_reads, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

int p3_solid_flags_export(p3_ctx *c, uint32_t *h_bitmap) {
    if (!c || !c->have_solid || !h_bitmap) return fail(P3_ERR_STATE, "p3_solid_flags_export: no make_bf result");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(h_bitmap, c->d_solid, sizeof(uint32_t) * c->n_words, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

static int with_kmers(p3_ctx *c, const uint64_t *h_kmers, uint64_t n, uint64_t **dk) {
    CU(cudaMalloc(dk, sizeof(uint64_t) * n));
    CU(cudaMemcpyAsync(*dk, h_kmers, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, c->stream));
    return P3_OK;
}

int p3_bf_add(p3_ctx *c, const uint64_t *h_kmers, uint64_t n) {
    if (!c || !c->have_bf) return fail(P3_ERR_STATE, "p3_bf_add: no filter");
    if (n == 0) return P3_OK;
    if (c->k > 32) return long_batch(c, 0, c->k, h_kmers, n, nullptr);
    CU(cudaSetDevice(c->device));
    uint64_t *dk = nullptr;
    TmpFree tmp; tmp.own(&dk);
    int rc = with_kmers(c, h_kmers, n, &dk);
    if (rc) return rc;
    bf_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->bloom(), dk, n);
    c->launches++;
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

int p3_bf_possibly_contains(p3_ctx *c, const uint64_t *h_kmers, uint64_t n, uint8_t *h_out) {
    if (!c || !c->have_bf) return fail(P3_ERR_STATE, "p3_bf_possibly_contains: no filter");
    if (n == 0) return P3_OK;
    if (c->k > 32) return long_batch(c, 1, c->k, h_kmers, n, h_out);
    CU(cudaSetDevice(c->device));
    uint64_t *dk = nullptr; uint8_t *dout = nullptr;
    TmpFree tmp; tmp.own(&dk); tmp.own(&dout);
    int rc = with_kmers(c, h_kmers, n, &dk);
    if (rc) return rc;
    CU(cudaMalloc(&dout, n));
    bf_query_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->bloom(), dk, n, dout);
    c->launches++;
    CU(cudaMemcpyAsync(h_out, dout, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

int p3_double_hash(p3_ctx *c, uint32_t k, const uint64_t *h_kmers, uint64_t n, uint64_t *h_out) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    if (k < 1 || k > P3_MAX_K) return fail(P3_ERR_ARG, "p3_double_hash: k must be <= 3001");
    if (n == 0) return P3_OK;
    if (k > 32) return long_batch(c, 2, k, h_kmers, n, h_out);
    CU(cudaSetDevice(c->device));
    uint64_t *dk = nullptr, *dout = nullptr;
    TmpFree tmp; tmp.own(&dk); tmp.own(&dout);
    int rc = with_kmers(c, h_kmers, n, &dk);
    if (rc) return rc;
    CU(cudaMalloc(&dout, sizeof(uint64_t) * 2 * n));
    double_hash_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((int)((2 * k + 7) / 8), dk, n, dout);
    c->launches++;
    CU(cudaMemcpyAsync(h_out, dout, sizeof(uint64_t) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

// ---- stage C -----------------------------------------------------------------------------------
int p3_dbg_adjacency(p3_ctx *c) {
    if (!c || !c->have_solid) return fail(P3_ERR_STATE, "p3_dbg_adjacency: run p3_make_bf first");
    CU(cudaSetDevice(c->device));
    uint64_t n = c->h_stats.n_distinct_solid;
    if (!c->d_adj || c->adj_cap < n) {
        dfree(c->d_adj);
        c->adj_cap = std::max<uint64_t>(c->list_cap, std::max<uint64_t>(n, 1));
        CU(cudaMalloc(&c->d_adj, c->adj_cap));
    }
    CU(cudaMemsetAsync(&c->d_stats->n_edges, 0, sizeof(unsigned long long), c->stream));
    CU(cudaEventRecord(c->ev[7], c->stream));
    if (n && c->k > 32) {
        int rcl = adjacency_long(c, long_words(c), n, c->d_adj, c->d_stats);
        if (rcl) return rcl;
    } else if (n) {
        KSet sa = c->set_valid ? c->kset() : p3_ctx::no_set(), sb = (c->set_valid && c->d_set_b) ? c->kset_b() : p3_ctx::no_set();
        if (const char *e = getenv("P3_ADJ_SET")) {   // experiment knob: which shortcut set the neighbour queries may use
            if (!strcmp(e, "owned")) sb = p3_ctx::no_set();
            else if (!strcmp(e, "none")) sa = sb = p3_ctx::no_set();
        }
        if (c->hints_valid) {   // hinted directions need no query; the rest go straight to the filter (a set lookup would only add an access)
            if (getenv("P3_ADJ_SET") && !strcmp(getenv("P3_ADJ_SET"), "hint+set")) adjacency_kernel<true><<<c->grid(), 256, 0, c->stream>>>(c->d_list, n, (int)c->k, c->bloom(), sa, sb, c->d_adj, c->d_stats);
            else adjacency_kernel<true><<<c->grid(), 256, 0, c->stream>>>(c->d_list, n, (int)c->k, c->bloom(), p3_ctx::no_set(), p3_ctx::no_set(), c->d_adj, c->d_stats);
        } else adjacency_kernel<false><<<c->grid(), 256, 0, c->stream>>>(c->d_list, n, (int)c->k, c->bloom(), sa, sb, c->d_adj, c->d_stats);
        c->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(c->ev[8], c->stream));
    int rc = pull_stats(c);
    if (rc) return rc;
    CU(cudaEventElapsedTime(&c->ms[4], c->ev[7], c->ev[8]));
    c->have_adj = true;
    c->n_solid = n; c->n_closed = n; c->closed = false;
    return P3_OK;
}

// fixed point of "add every recorded neighbour" (see closure_kernel), seeded with the solid k-mers
// and the given walk roots; *n_total = solid + root + phantom k-mers
int p3_dbg_close(p3_ctx *c, const uint64_t *h_roots, uint64_t n_roots, uint64_t *n_total) {
    if (!c || !c->have_adj) return fail(P3_ERR_STATE, "p3_dbg_close: run p3_dbg_adjacency first");
    if (c->k > 32) return fail(P3_ERR_ARG, "p3_dbg_close: k > 32 is not supported yet (the host walk is single-word)");
    CU(cudaSetDevice(c->device));
    // from here on the set also holds k-mers that need not answer possiblyContains (roots), so it
    // must not short-cut Bloom queries any more
    c->set_valid = false;
    auto adjacency_range = [&](uint64_t from, uint64_t to) -> int {
        if (to <= from) return P3_OK;
        uint64_t nn = to - from, warps = (nn + 3) / 4;
        unsigned blocks = (unsigned)std::min<uint64_t>((warps + 7) / 8, (uint64_t)c->grid());
        adjacency_kernel<false><<<blocks, 256, 0, c->stream>>>(c->d_list + from, nn, (int)c->k, c->bloom(), p3_ctx::no_set(), p3_ctx::no_set(), c->d_adj + from, nullptr);
        c->launches++;
        CU(cudaGetLastError());
        return P3_OK;
    };
    auto check_full = [&](uint64_t n2) -> int {
        if (c->h_stats.err_table_full || n2 > c->list_cap || n2 > c->adj_cap)
            return fail(P3_ERR_TABLE_FULL, "p3_dbg_close: k-mer set full (false-positive closure too large; raise solid_slots or -m)");
        return P3_OK;
    };
    uint64_t hi = c->n_closed;
    if (h_roots && n_roots) {
        uint64_t *dr = nullptr;
        TmpFree tmp; tmp.own(&dr);
        CU(cudaMalloc(&dr, sizeof(uint64_t) * n_roots));
        CU(cudaMemcpyAsync(dr, h_roots, sizeof(uint64_t) * n_roots, cudaMemcpyHostToDevice, c->stream));
        roots_kernel<<<(unsigned)((n_roots + 255) / 256), 256, 0, c->stream>>>(dr, n_roots, (int)c->k, c->kset(), c->d_list, c->list_cap, c->d_stats);
        c->launches++;
        int rc = pull_stats(c);
        if (rc) return rc;
        uint64_t n1 = c->h_stats.n_distinct_solid;
        if ((rc = check_full(n1))) return rc;
        if ((rc = adjacency_range(hi, n1))) return rc;
        hi = n1;
    }
    uint64_t lo = c->closed ? c->n_closed : 0;
    for (int round = 0; lo < hi; round++) {
        if (round > 200) return fail(P3_ERR_TABLE_FULL, "p3_dbg_close: no fixed point (filter saturated: the reference walk would not terminate either)");
        closure_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_list, c->d_adj, lo, hi, (int)c->k, c->kset(), c->list_cap, c->d_stats);
        c->launches++;
        CU(cudaGetLastError());
        int rc = pull_stats(c);
        if (rc) return rc;
        uint64_t n2 = c->h_stats.n_distinct_solid;
        if ((rc = check_full(n2))) return rc;
        if ((rc = adjacency_range(hi, n2))) return rc;
        lo = hi; hi = n2;
    }
    CU(cudaStreamSynchronize(c->stream));
    c->n_closed = hi; c->closed = true;
    if (n_total) *n_total = hi;
    return P3_OK;
}

int p3_dbg_stats(p3_ctx *c, uint64_t *n_kmers, uint64_t *n_edges) {
    if (!c || !c->have_adj) return fail(P3_ERR_STATE, "no adjacency result");
    if (n_kmers) *n_kmers = c->n_solid;
    if (n_edges) *n_edges = c->h_stats.n_edges;
    return P3_OK;
}

int p3_dbg_export(p3_ctx *c, uint64_t *h_kmers, uint8_t *h_adj, uint64_t cap, uint64_t *n) {
    if (!c || !c->have_solid) return fail(P3_ERR_STATE, "p3_dbg_export: no make_bf result");
    CU(cudaSetDevice(c->device));
    uint64_t nd = c->have_adj ? c->n_closed : c->h_stats.n_distinct_solid;
    if (n) *n = nd;
    if (cap < nd) return fail(P3_ERR_ARG, "p3_dbg_export: capacity too small");
    if (nd == 0) return P3_OK;
    const uint64_t W = (2 * (uint64_t)c->k + 63) / 64;
    if (h_kmers) CU(cudaMemcpyAsync(h_kmers, c->k > 32 ? long_words(c) : c->d_list, sizeof(uint64_t) * nd * W, cudaMemcpyDeviceToHost, c->stream));
    if (h_adj) {
        if (!c->have_adj) return fail(P3_ERR_STATE, "p3_dbg_export: run p3_dbg_adjacency first");
        CU(cudaMemcpyAsync(h_adj, c->d_adj, nd, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

int p3_check_directions(p3_ctx *c, const uint64_t *h_kmers, uint64_t n, uint8_t *h_mask) {
    if (!c || !c->have_bf) return fail(P3_ERR_STATE, "p3_check_directions: no filter");
    if (n == 0) return P3_OK;
    if (c->k > 32) return long_batch(c, 3, c->k, h_kmers, n, h_mask);
    CU(cudaSetDevice(c->device));
    uint64_t *dk = nullptr; uint8_t *dout = nullptr;
    TmpFree tmp; tmp.own(&dk); tmp.own(&dout);
    int rc = with_kmers(c, h_kmers, n, &dk);
    if (rc) return rc;
    CU(cudaMalloc(&dout, n));
    uint64_t warps = (n + 3) / 4;
    unsigned blocks = (unsigned)std::min<uint64_t>((warps + 7) / 8, (uint64_t)c->grid());
    adjacency_kernel<false><<<blocks, 256, 0, c->stream>>>(dk, n, (int)c->k, c->bloom(), c->set_valid ? c->kset() : p3_ctx::no_set(), (c->set_valid && c->d_set_b) ? c->kset_b() : p3_ctx::no_set(), dout, nullptr);
    c->launches++;
    CU(cudaMemcpyAsync(h_mask, dout, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

// ---- whole path ----------------------------------------------------------------------------------
// upload for the whole-path entry points: offsets (and the non-ACGT plane) on the main stream, the 2-bit staging in
// pieces on the copy stream, one event per piece; the count bins a piece as soon as it is there
static int upload_pieces(p3_ctx *c, const uint64_t *h_packed, uint64_t total_bases, const uint64_t *h_off, uint64_t n_reads, const uint32_t *h_nmask) {
    if (!h_packed || !h_off) return fail(P3_ERR_ARG, "p3_assemble_hot_path: null argument");
    CU(cudaSetDevice(c->device));
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    }
    c->have_reads = false;
    c->total_bases = total_bases; c->n_reads = n_reads; c->n_words = (total_bases + 31) / 32;
    const uint64_t pw = c->n_words + 1;
    CU(ensure(c->own_packed, c->cap_packed, sizeof(uint64_t) * pw));
    CU(ensure(c->own_off, c->cap_off, sizeof(uint64_t) * (n_reads + 1)));
    CU(cudaMemcpyAsync(c->own_off, h_off, sizeof(uint64_t) * (n_reads + 1), cudaMemcpyHostToDevice, c->stream));
    if (h_nmask) {
        CU(ensure(c->own_nmask, c->cap_nmask, sizeof(uint32_t) * pw));
        CU(cudaMemcpyAsync(c->own_nmask, h_nmask, sizeof(uint32_t) * pw, cudaMemcpyHostToDevice, c->stream));
    }
    // the copy stream may only overwrite the staging once the main stream's earlier kernels are done with it
    CU(cudaEventRecord(c->ev_main, c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
    const uint64_t pieces = std::min<uint64_t>(std::max<uint64_t>(c->n_words >> 22, 1), 16);
    const uint64_t piece_words = ((c->n_words + pieces - 1) / pieces + kTileWords - 1) / kTileWords * kTileWords;
    while (c->up_ev.size() < pieces) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->up_ev.push_back(e); }
    for (uint64_t i = 0; i < pieces; i++) {
        const uint64_t a = std::min<uint64_t>(i * piece_words, pw), b = std::min<uint64_t>((i + 1) * piece_words + 1, pw);   // + the hand-off word
        if (b > a) CU(cudaMemcpyAsync(c->own_packed + a, h_packed + a, sizeof(uint64_t) * (b - a), cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->up_ev[i], c->copy_stream));
    }
    c->up_pieces = pieces; c->up_piece_words = std::max<uint64_t>(piece_words, 1);
    c->d_packed = c->own_packed; c->d_off = c->own_off; c->d_nmask = h_nmask ? c->own_nmask : nullptr;
    return finish_reads(c);   // the read-end plane only needs the offsets
}
static int hot_path_impl(p3_ctx *c, const uint64_t *h_packed, uint64_t total_bases, const uint64_t *h_off,
                         uint64_t n_reads, const uint32_t *h_nmask, uint64_t all_bases, uint32_t k,
                         uint64_t filter_size, uint32_t num_hashes, uint64_t table_slots, uint64_t solid_slots,
                         uint8_t *h_bits, int64_t *h_seed_pos, uint64_t *h_kmers, uint8_t *h_adj, uint64_t cap, uint64_t *n_out) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    if (filter_size == 0) {
        int rc = p3_estimate_bloomfilter(all_bases, k, &filter_size, &num_hashes);
        if (rc) return rc;
    }
    int rc = upload_pieces(c, h_packed, total_bases, h_off, n_reads, h_nmask);
    if (!rc) rc = p3_count_short_kmers(c, table_slots);
    if (c->up_pieces) upload_wait_all(c);
    if (rc) return rc;
    rc = p3_make_bf(c, k, filter_size, num_hashes, P3_COV_THRESHOLD, solid_slots);
    if (rc) return rc;
    const bool want_out = h_bits || h_seed_pos || h_kmers || h_adj;
    const uint64_t nd = c->h_stats.n_distinct_solid, W = (2 * (uint64_t)k + 63) / 64;
    if (n_out) *n_out = nd;
    if (want_out && (h_kmers || h_adj) && cap < nd) return fail(P3_ERR_ARG, "p3_assemble_hot_path_to_host: capacity too small");
    if (want_out) {   // filter, seeds and k-mers are final now: they leave on the copy stream while CheckDirections runs
        CU(cudaEventRecord(c->ev_main, c->stream));
        CU(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
        if (h_bits) CU(cudaMemcpyAsync(h_bits, c->d_bloom, (c->filter_size + 7) / 8, cudaMemcpyDeviceToHost, c->copy_stream));
        if (h_seed_pos) CU(cudaMemcpyAsync(h_seed_pos, c->d_seed, sizeof(int64_t) * c->n_reads, cudaMemcpyDeviceToHost, c->copy_stream));
        if (h_kmers && nd) CU(cudaMemcpyAsync(h_kmers, k > 32 ? long_words(c) : c->d_list, sizeof(uint64_t) * nd * W, cudaMemcpyDeviceToHost, c->copy_stream));
        CU(cudaEventRecord(c->ev_copy, c->copy_stream));
    }
    rc = p3_dbg_adjacency(c);
    if (rc) return rc;
    if (want_out) {
        if (h_adj && nd) CU(cudaMemcpyAsync(h_adj, c->d_adj, nd, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
        CU(cudaStreamSynchronize(c->stream));
    }
    return P3_OK;
}
int p3_assemble_hot_path(p3_ctx *c, const uint64_t *h_packed, uint64_t total_bases, const uint64_t *h_off,
                         uint64_t n_reads, const uint32_t *h_nmask, uint64_t all_bases, uint32_t k,
                         uint64_t filter_size, uint32_t num_hashes, uint64_t table_slots, uint64_t solid_slots) {
    return hot_path_impl(c, h_packed, total_bases, h_off, n_reads, h_nmask, all_bases, k, filter_size, num_hashes, table_slots, solid_slots,
                         nullptr, nullptr, nullptr, nullptr, 0, nullptr);
}
int p3_assemble_hot_path_to_host(p3_ctx *c, const uint64_t *h_packed, uint64_t total_bases, const uint64_t *h_off,
                                 uint64_t n_reads, const uint32_t *h_nmask, uint64_t all_bases, uint32_t k,
                                 uint64_t filter_size, uint32_t num_hashes, uint64_t table_slots, uint64_t solid_slots,
                                 uint8_t *h_bits, int64_t *h_seed_pos, uint64_t *h_kmers, uint8_t *h_adj, uint64_t cap, uint64_t *n) {
    return hot_path_impl(c, h_packed, total_bases, h_off, n_reads, h_nmask, all_bases, k, filter_size, num_hashes, table_slots, solid_slots,
                         h_bits, h_seed_pos, h_kmers, h_adj, cap, n);
}

// Probe statistics for the k x threshold sweep (BASELINE.json configs[4]). out[0], out[1]: mean and longest number of
// 32-byte buckets an insert of the last p3_count_short_kmers touched (only when it ran with P3_PROBE_STATS set, else 0);
// out[2], out[3]: the same for a successful lookup of every distinct solid k-mer in the solid set (measured now).
int p3_probe_stats(p3_ctx *c, double out[4]) {
    if (!c || !out) return fail(P3_ERR_ARG, "p3_probe_stats: null argument");
    CU(cudaSetDevice(c->device));
    out[0] = out[1] = out[2] = out[3] = 0;
    if (c->have_counts && c->probe_stats && c->h_stats.n_pos21) { out[0] = (double)c->count_probes / (double)c->h_stats.n_pos21; out[1] = c->count_max_probe; }
    if (c->have_solid && c->k <= 32 && c->d_set && c->h_stats.n_distinct_solid) {
        const uint64_t n = c->n_solid ? c->n_solid : c->h_stats.n_distinct_solid;
        CU(cudaMemsetAsync(&c->d_stats->probes, 0, sizeof(unsigned long long), c->stream));
        CU(cudaMemsetAsync(&c->d_stats->max_probe, 0, sizeof(unsigned), c->stream));
        set_probe_stats_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_list, n, c->kset(), c->d_stats);
        c->launches++;
        Stats h;
        CU(cudaMemcpyAsync(&h, c->d_stats, sizeof(Stats), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        out[2] = (double)h.probes / (double)n; out[3] = h.max_probe;
    }
    return P3_OK;
}
// out[0..3] = count table slots, count table partitions, solid set slots, solid set partitions
uint32_t p3_table_partitions(uint64_t table_slots) { return table_partitions(table_slots); }
int p3_table_capacity(p3_ctx *c, uint64_t out[4]) {
    if (!c || !out) return fail(P3_ERR_ARG, "p3_table_capacity: null argument");
    out[0] = c->nb * 4; out[1] = c->parts; out[2] = c->nbs * 4; out[3] = c->set_parts;
    return P3_OK;
}

int p3_count_substage_ms(p3_ctx *c, float ms[4], uint32_t *parts, uint64_t *chunks) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    for (int i = 0; i < 3; i++) ms[i] = c->binned ? c->ms_sub[i] : 0.f;
    ms[3] = c->ms_bloom;
    if (parts) *parts = c->parts;
    if (chunks) *chunks = c->binned ? c->n_chunks : 0;
    return P3_OK;
}
int p3_stage_ms(p3_ctx *c, float ms[5]) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    for (int i = 0; i < 5; i++) ms[i] = c->ms[i];
    return P3_OK;
}
uint64_t p3_launch_count(p3_ctx *c) { return c ? c->launches : 0; }
int p3_bf_params(p3_ctx *c, uint64_t *filter_size, uint32_t *num_hashes, uint32_t *k) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    if (filter_size) *filter_size = c->filter_size;
    if (num_hashes) *num_hashes = c->num_hashes;
    if (k) *k = c->k;
    return P3_OK;
}

}  // extern "C"

#include "p3_long.inc.cu"
#include "p3_bloom.inc.cu"
#include "p3_cover.inc.cu"
#include "p3_multi.inc.cu"
