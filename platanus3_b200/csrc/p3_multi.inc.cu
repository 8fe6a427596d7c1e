// p3_multi.inc.cu — multi-GPU hot path (DESIGN.md row e), part of the p3_gpu.cu translation unit.
// One process per GPU. Canonical k-mers are hash-partitioned: the OWNER of a key counts it / de-duplicates
// it. Per-rank work does not grow with the number of ranks:
//
//   A   every rank bins its 21-mers by owner; the binning kernel stores owner j's records STRAIGHT INTO
//       rank j's receive region over NVLink peer memory (bin + exchange are one kernel). The owner sorts
//       what it received into its table-partition bins (which persist over the chunks) and, after the
//       last chunk, counts everything in ONE L2-resident insert sweep.
//   B1  the owner sweeps its bins a second time: records of keys whose count stayed below the threshold
//       go back to their source rank (sorted by rank, peer stores), which clears those coverage bits.
//   B2  solid plane and seeds are local. Every solid OCCURRENCE (canonical k-mer + adjacency hint, 9 B)
//       goes to the owner of the k-mer the same way; the owner sorts them by set partition and
//       de-duplicates them in one L2-resident sweep that also ORs the hints together.
//   B3  Bloom adds: sharded filter (p3_bloom.inc.cu), all-gather of the shards.
//   C   CheckDirections of the owned k-mers: hinted directions are known, the others probe the filter.
//
// Every rank owns ONE peer-visible arena: [PeerCtl | receive set 0 | receive set 1]. A set is a row of
// n_ranks regions (region r = what source rank r sent). Stages alternate between the two sets, so one
// cross-rank barrier per step is enough (see p3_mg_sync). The barrier is a one-block kernel that writes
// an epoch into every peer's PeerCtl and spins on its own — no host round trip, no NCCL call.
// Transport 1 ("staged") writes the same regions into a local staging buffer instead and leaves the
// movement to the caller's all-to-all (NCCL): the baseline the fused path is measured against.

constexpr int kCtlBytes = 4096;
struct PeerCtl {
    unsigned long long flag[kMaxPeers];         // flag[r]: last barrier epoch rank r announced to this rank
    unsigned long long count[2][kMaxPeers];     // count[set][r]: records source r stored into region r of receive set `set`
};
struct PeerLinks { unsigned char *arena[kMaxPeers]; };
struct PeerOut64 { uint64_t *p[kMaxPeers]; };

__global__ void peer_sync_kernel(PeerLinks L, int me, int n, unsigned long long epoch, Stats *st) {
    const int j = threadIdx.x;
    if (j >= n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&reinterpret_cast<PeerCtl *>(L.arena[j])->flag[me]), "l"(epoch) : "memory");
    const unsigned long long *mine = &reinterpret_cast<PeerCtl *>(L.arena[me])->flag[j];
    unsigned long long v = 0;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > 60000000000LL) { atomicExch(&st->err_peer_timeout, 1u); break; }   // ~30 s: a peer died
        __nanosleep(500);
    }
}
// count[set][me] of every destination rank := what this rank sent there (peer stores; the barrier that follows publishes them)
__global__ void publish_counts_kernel(PeerLinks L, int me, int n, int set, const unsigned long long *__restrict__ sent) {
    const int j = threadIdx.x;
    if (j < n) reinterpret_cast<PeerCtl *>(L.arena[j])->count[set][me] = sent[j];
}
// staged transport: the counts travel with the caller's all-to-all; this only lines them up
__global__ void stage_counts_kernel(unsigned long long *dst, const unsigned long long *__restrict__ sent, int n) {
    const int j = threadIdx.x;
    if (j < n) dst[j] = sent[j];
}

// records in the fixed-capacity bins (sum over the partitions), added to st->owned_pos: the owner's position
// total when the insert runs in several rounds over re-used bins
__global__ void accum_cursors_kernel(const unsigned long long *__restrict__ cursor, uint32_t P, uint64_t cap, Stats *st) {
    unsigned long long mine = 0;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) mine += min(cursor[i] - (unsigned long long)i * cap, (unsigned long long)cap);
    if (mine) atomicAdd(&st->owned_pos, mine);     // one block, at most 256 adds
}

// tile sort of a position list by source rank (the top byte) into the ranks' regions
__global__ void __launch_bounds__(kScatterThreads, 3)
scatter_pos_peer_kernel(const uint64_t *__restrict__ in, uint64_t n, const unsigned long long *__restrict__ n_dev, uint32_t P,
                        unsigned long long *cursor, PeerOut64 out, uint64_t cap, Stats *st) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using SM = ScatterSmemT<true, false>;
    SM &sm = *reinterpret_cast<SM *>(smem_raw);
    constexpr int NT = kScatterThreads, PER = kTilePos / NT;
    const int tid = threadIdx.x;
    if (n_dev) n = min(n, (uint64_t)*n_dev);
    const uint64_t n_tiles = (n + kTilePos - 1) / kTilePos;
    bool over = false;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t i = tid; i < P; i += NT) sm.hist[i] = 0;
        __syncthreads();
        const uint64_t t0 = tile * kTilePos;
        uint64_t rec[PER];
        uint32_t rk[PER / 2];
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint64_t i = t0 + j * NT + tid;
            rec[j] = i < n ? __ldcs(in + i) : 0;
            if ((j & 1) == 0) rk[j >> 1] = 0;
            if (i < n) rk[j >> 1] |= atomicAdd(&sm.hist[pid_of<2>(rec[j], P)], 1u) << (16 * (j & 1));
        }
        __syncthreads();
        tile_scan_and_claim<NT>(sm, P, cursor, tid);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint64_t i = t0 + j * NT + tid;
            if (i < n) {
                uint32_t pt = pid_of<2>(rec[j], P);
                uint32_t idx = sm.hist[pt] + ((rk[j >> 1] >> (16 * (j & 1))) & 0xFFFFu);
                sm.key[idx] = rec[j];
                sm.part[idx] = (uint16_t)pt;
            }
        }
        __syncthreads();
        const uint32_t total = sm.total;
        for (uint32_t i = tid; i < total; i += NT) {
            uint32_t pt = sm.part[i];
            unsigned long long dst = sm.gbase[pt] + (i - sm.hist[pt]);
            if (dst < cap) out.p[pt][dst] = sm.key[i];
            else over = true;
        }
        __syncthreads();
    }
    if (over) atomicExch(&st->err_bin_overflow, 1u);
}

// host-side copy of owner_of (for tests and for callers that route on the host)
static inline uint32_t owner_of_host(uint64_t key, uint32_t n) {
    return (uint32_t)(((unsigned __int128)owner_mix(key) * n) >> 64);
}

struct MgState {   // per-context state of the multi-GPU path
    uint32_t n_ranks = 0, my_rank = 0; int transport = 0; bool emulated = false, connected = false;
    unsigned char *arena = nullptr; uint64_t arena_bytes = 0, set_bytes = 0;
    unsigned char *staging = nullptr; uint64_t cap_staging = 0;         // transport 1: one set worth of send regions
    unsigned long long *d_stage_cnt = nullptr;                          // transport 1: counts lined up for the all-to-all
    PeerLinks links;
    unsigned long long epoch = 0;
    unsigned long long *d_sent = nullptr;                               // [kMaxParts + 1] per-destination cursors of the current send
    // A
    uint64_t chunk_words = 0, n_chunks = 0, capA = 0, part_cap = 0;
    bool multi_round = false;    // the owner inserted in several rounds: the bins hold the last round only (human-scale inputs)
    // key-range rounds: round r counts the keys of table partitions [p_lo, p_hi) only — ALL their occurrences, so the counts
    // of those partitions are final when the round's insert ends and its verdicts can follow at once. The bins hold the
    // partitions of one round and are addressed with absolute indices through pointers shifted back by p_lo * part_cap.
    uint32_t key_rounds = 1, p_lo = 0, p_hi = 0; bool round_bins_valid = false;
    uint64_t *bk(const p3_ctx *c) const;   // partition bins as the kernels of the current round see them
    uint32_t *bw(const p3_ctx *c) const;
    uint32_t *bi(const p3_ctx *c) const;
    // B1
    uint64_t *d_sing = nullptr; uint64_t cap_sing = 0; uint64_t capB = 0; uint32_t n_slices = 0;
    // B2
    uint64_t capK = 0, kpart_cap = 0; uint32_t k = 0;
    std::vector<void *> graveyard;   // outgrown peer-visible buffers: freed with the context, never while peers may map them
    unsigned char *set_ptr(uint32_t rank, int set) const { return links.arena[rank] + kCtlBytes + (uint64_t)set * set_bytes; }
    PeerCtl *ctl() const { return reinterpret_cast<PeerCtl *>(arena); }
};
static CtxStates<MgState> g_mg;
// (device addresses are plain integers to the host: the shifted base lies before the allocation, only indices of the current
// round's partitions are ever added to it)
template <typename T> static T *shifted_base(T *p, uint64_t elems) { return reinterpret_cast<T *>(reinterpret_cast<uintptr_t>(p) - elems * sizeof(T)); }
uint64_t *MgState::bk(const p3_ctx *c) const { return key_rounds > 1 ? shifted_base(c->d_bkeys, (uint64_t)p_lo * part_cap) : c->d_bkeys; }
uint32_t *MgState::bw(const p3_ctx *c) const { return key_rounds > 1 ? shifted_base(c->d_bword, (uint64_t)p_lo * part_cap) : c->d_bword; }
uint32_t *MgState::bi(const p3_ctx *c) const { return key_rounds > 1 ? shifted_base(c->d_bidx, (uint64_t)p_lo * part_cap) : c->d_bidx; }
// multi-word k-mers (see the end of this file)
struct LongMg {
    uint64_t *d_store = nullptr; uint64_t cap_store = 0;    // bytes
    unsigned long long *d_store_n = nullptr;                // records in the store (device)
    uint64_t store_records = 0;                             // capacity in records
    uint64_t chunk_words = 0, capL = 0; int W = 0;
};
static CtxStates<LongMg> g_longmg;
static void longmg_release(p3_ctx *c) {
    LongMg *l = g_longmg.find(c);
    if (!l) return;
    dfree(l->d_store); dfree(l->d_store_n);
    g_longmg.erase(c);
}


static void mg_release(p3_ctx *c) {
    MgState *mp = g_mg.find(c);
    if (!mp) return;
    MgState &m = *mp;
    dfree(m.arena); dfree(m.staging); dfree(m.d_stage_cnt); dfree(m.d_sent); dfree(m.d_sing);
    for (void *p : m.graveyard) cudaFree(p);
    g_mg.erase(c);
    longmg_release(c);
}
static int mg_ready(p3_ctx *c, MgState **out, const char *who) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    MgState *m = g_mg.find(c);
    if (!m || !m->connected) return fail(P3_ERR_STATE, std::string(who) + ": run p3_mg_arena and p3_mg_connect first");
    CU(cudaSetDevice(c->device));
    *out = m;
    return P3_OK;
}
// records per region of a receive set for records of `bytes` bytes each (keys block + aux block), a multiple of the sort tile
static uint64_t region_cap(const MgState &m, uint64_t bytes) { return m.set_bytes / m.n_ranks / bytes / kTilePos * kTilePos; }
// where this rank's records for destination j go: peer transport = region my_rank of rank j's set, staged = region j of the local staging
static unsigned char *dest_block(const MgState &m, uint32_t j, int set) { return m.transport == 0 ? m.set_ptr(j, set) : m.staging; }
static uint32_t dest_region(const MgState &m, uint32_t j) { return m.transport == 0 ? m.my_rank : j; }

extern "C" {

uint32_t p3_owner_of_key(uint64_t key, uint32_t n_ranks) { return n_ranks ? owner_of_host(key, n_ranks) : 0; }

// CUDA IPC plumbing for the arenas (one process per GPU): export a cudaMalloc'ed buffer of this process as a
// 64-byte handle / map another process's buffer into this one (NVLink peer access is enabled on first use)
int p3_ipc_export(const void *d_ptr, uint8_t handle[64]) {
    if (!d_ptr || !handle) return fail(P3_ERR_ARG, "p3_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    memcpy(handle, &h, 64);
    return P3_OK;
}
int p3_ipc_open(int device, const uint8_t handle[64], void **d_ptr) {
    if (!handle || !d_ptr) return fail(P3_ERR_ARG, "p3_ipc_open: null argument");
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return P3_OK;
}
int p3_ipc_close(int device, void *d_ptr) {
    if (!d_ptr) return P3_OK;
    CU(cudaSetDevice(device));
    CU(cudaIpcCloseMemHandle(d_ptr));
    return P3_OK;
}

// The peer-visible arena of this rank: control block + two receive sets of set_bytes each (the same value on every
// rank). transport 0 = peers store into it over NVLink (export it with p3_ipc_export), 1 = staged: a local send
// buffer of one set is added and the caller moves the regions with an all-to-all. A grown arena is a NEW
// allocation (export it again); the old one is kept until the context is destroyed because peers may still map it.
int p3_mg_arena(p3_ctx *c, uint32_t n_ranks, uint32_t my_rank, uint64_t set_bytes, int transport, void **d_arena) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    if (n_ranks == 0 || n_ranks > (uint32_t)kMaxPeers || my_rank >= n_ranks) return fail(P3_ERR_ARG, "p3_mg_arena: 1..16 ranks");
    if (transport != 0 && transport != 1) return fail(P3_ERR_ARG, "p3_mg_arena: transport 0 (peer) or 1 (staged)");
    CU(cudaSetDevice(c->device));
    MgState &m = g_mg.get(c);
    set_bytes = std::max<uint64_t>((set_bytes + 4095) / 4096 * 4096, (uint64_t)n_ranks * kTilePos * 12);
    const uint64_t need = kCtlBytes + 2 * set_bytes;
    if (!m.arena || m.arena_bytes != need) {
        if (m.arena) m.graveyard.push_back(m.arena);
        m.arena = nullptr;
        if (cudaMalloc((void **)&m.arena, need) != cudaSuccess) { cudaGetLastError(); return fail(P3_ERR_NOMEM, "p3_mg_arena: allocation failed"); }
        m.arena_bytes = need;
        CU(cudaMemsetAsync(m.arena, 0, kCtlBytes, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        m.epoch = 0;
    }
    m.set_bytes = set_bytes; m.n_ranks = n_ranks; m.my_rank = my_rank; m.transport = transport; m.connected = false;
    if (transport == 1) {
        CU(ensure(m.staging, m.cap_staging, set_bytes));
        if (!m.d_stage_cnt) CU(cudaMalloc(&m.d_stage_cnt, sizeof(unsigned long long) * kMaxPeers));
    }
    if (!m.d_sent) CU(cudaMalloc(&m.d_sent, sizeof(unsigned long long) * (kMaxParts + 1)));
    CU(cudaFuncSetAttribute(scatter_pos_peer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemP));
    int rca = scatter_attrs(c);
    if (rca) return rca;
    if (d_arena) *d_arena = m.arena;
    return P3_OK;
}
// arena_ptrs[r] = rank r's arena as mapped into this process (p3_ipc_open; this rank's own pointer at my_rank).
// same_stream != 0: all ranks are contexts of ONE process that launch on ONE stream (the emulated communicator of the
// tests); stream order then replaces the barrier.
int p3_mg_connect(p3_ctx *c, const uint64_t *arena_ptrs, int same_stream) {
    if (!c || !arena_ptrs) return fail(P3_ERR_ARG, "p3_mg_connect: null argument");
    MgState *mp = g_mg.find(c);
    if (!mp || !mp->arena) return fail(P3_ERR_STATE, "p3_mg_connect: run p3_mg_arena first");
    MgState &m = *mp;
    for (uint32_t j = 0; j < (uint32_t)kMaxPeers; j++) m.links.arena[j] = j < m.n_ranks ? (unsigned char *)(uintptr_t)arena_ptrs[j] : nullptr;
    if (m.links.arena[m.my_rank] != m.arena) return fail(P3_ERR_ARG, "p3_mg_connect: arena_ptrs[my_rank] is not this rank's arena");
    m.emulated = same_stream != 0;
    m.connected = true;
    return P3_OK;
}
// Cross-rank barrier on the context's stream: everything this rank stored into its peers before is visible to
// them after, and the other way round. One per step is enough because the stages alternate between the two
// receive sets: a rank can only enter step s+1's barrier after it has consumed step s, so nobody overwrites a
// set that is still being read. Also needed once at the start of every stage. Staged transport / one rank /
// same-stream emulation: nothing to do (the all-to-all resp. the stream orders the steps).
int p3_mg_sync(p3_ctx *c) {
    MgState *m;
    int rc = mg_ready(c, &m, "p3_mg_sync");
    if (rc) return rc;
    if (m->n_ranks == 1 || m->emulated || m->transport == 1) return P3_OK;
    m->epoch++;
    peer_sync_kernel<<<1, 32, 0, c->stream>>>(m->links, (int)m->my_rank, (int)m->n_ranks, m->epoch, c->d_stats);
    c->launches++;
    CU(cudaGetLastError());
    return P3_OK;
}
// staged transport: what the caller's all-to-all has to move for receive set `set` of a stage (0 = A count records,
// 1 = B1 positions, 2 = B2 k-mers, 3 = B2 multi-word k-mers). out[0..2] = send block, receive block, bytes per rank of the 8-byte records;
// out[3..5] the same for the auxiliary block (bytes per rank 0 = none); out[6..7] = send / receive counts (n_ranks uint64 each).
int p3_mg_staged_buffers(p3_ctx *c, int stage, int set, uint64_t out[8]) {
    MgState *m;
    int rc = mg_ready(c, &m, "p3_mg_staged_buffers");
    if (rc) return rc;
    if (m->transport != 1 || set < 0 || set > 1 || !out) return fail(P3_ERR_ARG, "p3_mg_staged_buffers: staged transport only");
    const uint64_t auxb = stage == 0 ? 4 : stage == 2 ? 1 : 0;
    uint64_t cap = stage == 0 ? m->capA : stage == 1 ? m->capB : m->capK;
    if (stage == 3) {    // multi-word k-mer records: W words each, counted here in 8-byte units
        LongMg *lm = g_longmg.find(c);
        if (!lm || !lm->W) return fail(P3_ERR_STATE, "p3_mg_staged_buffers: run p3_mg_long_begin first");
        cap = lm->capL * (uint64_t)lm->W;
    }
    unsigned char *recv = m->set_ptr(m->my_rank, set);
    out[0] = (uint64_t)(uintptr_t)m->staging; out[1] = (uint64_t)(uintptr_t)recv; out[2] = cap * 8;
    out[3] = (uint64_t)(uintptr_t)(m->staging + (uint64_t)m->n_ranks * cap * 8); out[4] = (uint64_t)(uintptr_t)(recv + (uint64_t)m->n_ranks * cap * 8); out[5] = cap * auxb;
    out[6] = (uint64_t)(uintptr_t)m->d_stage_cnt; out[7] = (uint64_t)(uintptr_t)&m->ctl()->count[set][0];
    return P3_OK;
}
static int mg_publish(p3_ctx *c, MgState &m, int set) {
    if (m.transport == 0) publish_counts_kernel<<<1, 32, 0, c->stream>>>(m.links, (int)m.my_rank, (int)m.n_ranks, set, m.d_sent);
    else stage_counts_kernel<<<1, 32, 0, c->stream>>>(m.d_stage_cnt, m.d_sent, (int)m.n_ranks);
    c->launches++;
    CU(cudaGetLastError());
    return P3_OK;
}

// ---- A: CountShortKmer, reference src/Load.cpp:105-127 ----------------------------------------------------------
// table_slots: capacity of this rank's count table. owner_positions: upper estimate of the 21-mer positions this rank
// will own (all ranks' positions / n_ranks for a hash partition) — sizes the partition bins, which hold every
// received record until the one insert sweep. n_chunks / chunk_words: the same on every rank (ranks with fewer words
// send empty chunks).
static int mg_count_begin(p3_ctx *c, uint64_t table_slots, uint64_t owner_positions, uint64_t chunk_words, uint64_t n_chunks, uint32_t key_rounds) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_count_begin");
    if (rc) return rc;
    MgState &m = *mp;
    if (!c->have_reads) return fail(P3_ERR_STATE, "p3_mg_count_begin: no reads attached");
    if (table_slots == 0 || chunk_words == 0) return fail(P3_ERR_ARG, "p3_mg_count_begin: table_slots and chunk_words required");
    c->binned = true;
    rc = setup_table(c, table_slots);
    if (!rc) rc = hist_buffers(c);
    if (rc) return rc;
    if (!c->d_binmeta) CU(cudaMalloc(&c->d_binmeta, sizeof(unsigned long long) * (kMaxParts + 2)));
    m.chunk_words = (chunk_words + kTileWords - 1) / kTileWords * kTileWords; m.n_chunks = n_chunks;
    m.capA = region_cap(m, 12);
    const uint64_t per_region = (uint64_t)((double)std::min<uint64_t>(m.chunk_words, c->n_words + kTileWords) * 32 / m.n_ranks * 1.03) + 8192;
    if (m.capA < per_region) return fail(P3_ERR_ARG, "p3_mg_count_begin: receive set too small for this chunk size (raise set_bytes or lower chunk_words)");
    const uint32_t P = c->parts;
    m.part_cap = ((uint64_t)((double)owner_positions / P * 1.03) + 8192 + kSweepChunk - 1) / kSweepChunk * kSweepChunk;
    m.key_rounds = std::max<uint32_t>(std::min<uint32_t>(key_rounds, P), 1);
    m.p_lo = 0; m.p_hi = P; m.round_bins_valid = false;
    const uint64_t P_bins = m.key_rounds > 1 ? (P + m.key_rounds - 1) / m.key_rounds + 1 : P;   // partitions whose bins exist at a time
    CU(ensure(c->d_bkeys, c->cap_bkeys, sizeof(uint64_t) * m.part_cap * P_bins));
    CU(ensure(c->d_bword, c->cap_bword, sizeof(uint32_t) * m.part_cap * P_bins));
    CU(ensure(c->d_bidx, c->cap_bidx, sizeof(uint32_t) * m.part_cap * P_bins));
    CU(ensure(c->d_valid, c->cap_valid, sizeof(uint32_t) * (c->n_words + 1)));
    init_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, m.part_cap);   // the bins persist over the chunks
    c->launches++;
    CU(cudaEventRecord(c->ev[0], c->stream));
    c->bins_valid = false; c->have_counts = false; c->pos_on_host = true; c->binned_pos = 0; c->n_chunks = 0;
    m.multi_round = m.key_rounds > 1;
    return P3_OK;
}
int p3_mg_count_begin(p3_ctx *c, uint64_t table_slots, uint64_t owner_positions, uint64_t chunk_words, uint64_t n_chunks) {
    return mg_count_begin(c, table_slots, owner_positions, chunk_words, n_chunks, 1);
}
// Key-range rounds for inputs whose records an owner cannot hold at once: owner_positions is the estimate for ALL rounds
// (a partition's bin must hold all occurrences of its keys), the bins exist for one round's partitions at a time. Per
// round r: p3_mg_key_round_begin(r); every chunk: p3_mg_count_send (only the keys of the round's partitions travel), sync,
// p3_mg_count_recv; p3_mg_count_finish; then the round's verdicts (p3_mg_cover_begin_keyed once, p3_mg_cover_key_round,
// slices of p3_mg_cover_send / _recv); p3_mg_count_next_round before the next round; p3_mg_count_end after the last.
int p3_mg_count_begin_keyed(p3_ctx *c, uint64_t table_slots, uint64_t owner_positions, uint64_t chunk_words, uint64_t n_chunks, uint32_t n_rounds) {
    if (n_rounds < 2) return fail(P3_ERR_ARG, "p3_mg_count_begin_keyed: at least 2 rounds (one round: p3_mg_count_begin)");
    return mg_count_begin(c, table_slots, owner_positions, chunk_words, n_chunks, n_rounds);
}
int p3_mg_key_round_begin(p3_ctx *c, uint32_t round) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_key_round_begin");
    if (rc) return rc;
    MgState &m = *mp;
    if (m.key_rounds < 2 || round >= m.key_rounds) return fail(P3_ERR_ARG, "p3_mg_key_round_begin: round out of range");
    const uint32_t P = c->parts;
    m.p_lo = (uint32_t)((uint64_t)P * round / m.key_rounds);
    m.p_hi = (uint32_t)((uint64_t)P * (round + 1) / m.key_rounds);
    m.round_bins_valid = false;
    init_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, m.part_cap);   // absolute: partitions outside the round stay empty
    c->launches++;
    CU(cudaGetLastError());
    return P3_OK;
}
// bin chunk `ch` of this rank's reads by owner, straight into the owners' receive set ch % 2 (peer stores)
int p3_mg_count_send(p3_ctx *c, uint64_t ch) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_count_send");
    if (rc) return rc;
    MgState &m = *mp;
    const int set = (int)(ch & 1);
    const uint64_t w0 = std::min<uint64_t>(ch * m.chunk_words, c->n_words), w1 = std::min<uint64_t>((ch + 1) * m.chunk_words, c->n_words);
    CU(cudaMemsetAsync(m.d_sent, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    if (w1 > w0) {
        PeerOut po;
        for (uint32_t j = 0; j < (uint32_t)kMaxPeers; j++) {
            po.keys[j] = nullptr; po.words[j] = nullptr;
            if (j < m.n_ranks) {
                unsigned char *blk = dest_block(m, j, set);
                const uint64_t reg = dest_region(m, j);
                po.keys[j] = reinterpret_cast<uint64_t *>(blk) + reg * m.capA;
                po.words[j] = reinterpret_cast<uint32_t *>(blk + (uint64_t)m.n_ranks * m.capA * 8) + reg * m.capA;
            }
        }
        const unsigned sblocks = (unsigned)std::min<uint64_t>((w1 - w0 + kTileWords - 1) / kTileWords, (uint64_t)c->n_sm * 3);
        const uint64_t tag = (uint64_t)m.my_rank << kRecRankShift;
        if (m.key_rounds > 1) {   // key-range rounds: this round's partitions only
            if (c->d_nmask) scatter21_kernel<true, 1, true, true><<<sblocks, kScatterThreads, kSmem21, c->stream>>>(c->d_packed, c->d_rend, c->d_nmask, w0, w1, m.n_ranks, m.d_sent, nullptr, nullptr, c->d_valid, tag, c->d_stats, po, m.capA, m.p_lo, m.p_hi, c->parts);
            else scatter21_kernel<false, 1, true, true><<<sblocks, kScatterThreads, kSmem21, c->stream>>>(c->d_packed, c->d_rend, nullptr, w0, w1, m.n_ranks, m.d_sent, nullptr, nullptr, c->d_valid, tag, c->d_stats, po, m.capA, m.p_lo, m.p_hi, c->parts);
        } else if (c->d_nmask) scatter21_kernel<true, 1, true><<<sblocks, kScatterThreads, kSmem21, c->stream>>>(c->d_packed, c->d_rend, c->d_nmask, w0, w1, m.n_ranks, m.d_sent, nullptr, nullptr, c->d_valid, tag, c->d_stats, po, m.capA);
        else scatter21_kernel<false, 1, true><<<sblocks, kScatterThreads, kSmem21, c->stream>>>(c->d_packed, c->d_rend, nullptr, w0, w1, m.n_ranks, m.d_sent, nullptr, nullptr, c->d_valid, tag, c->d_stats, po, m.capA);
        c->launches++;
        CU(cudaGetLastError());
    }
    return mg_publish(c, m, set);
}
// owner side: sort what arrived in receive set ch % 2 into the table-partition bins
int p3_mg_count_recv(p3_ctx *c, uint64_t ch) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_count_recv");
    if (rc) return rc;
    MgState &m = *mp;
    const int set = (int)(ch & 1);
    unsigned char *blk = m.set_ptr(m.my_rank, set);
    const uint64_t n = (uint64_t)m.n_ranks * m.capA;
    const unsigned sblocks = (unsigned)std::min<uint64_t>((n + kTilePos - 1) / kTilePos, (uint64_t)c->n_sm * 3);
    scatter_rec_kernel<0, 4><<<sblocks, kScatterThreads, kSmemA4, c->stream>>>(
        reinterpret_cast<const uint64_t *>(blk), blk + n * 8, n, c->parts, c->d_cursor, m.bk(c), m.bw(c), m.part_cap, nullptr,
        m.capA, &m.ctl()->count[set][0], c->d_stats);
    c->launches++;
    CU(cudaGetLastError());
    c->n_chunks++;
    return P3_OK;
}
// after the last chunk: one L2-resident insert sweep over everything this rank owns (enqueue only)
int p3_mg_count_finish(p3_ctx *c) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_count_finish");
    if (rc) return rc;
    MgState &m = *mp;
    const uint32_t P = c->parts;
    check_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, m.part_cap, c->d_stats);
    CU(cudaMemcpyAsync(c->d_binmeta, c->d_cursor, sizeof(unsigned long long) * P, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaEventRecord(c->ev[10], c->stream));
    c->bin_cap = m.part_cap; c->bin_n = (uint64_t)P * m.part_cap;
    if (m.key_rounds > 1) rc = launch_insert_bins(c, m.bk(c), (uint64_t)m.p_hi * m.part_cap, c->bin_cap, c->d_binmeta, nullptr, m.bi(c), (unsigned long long)m.p_lo * m.part_cap);
    else rc = launch_insert_bins(c, c->d_bkeys, c->bin_n, c->bin_cap, c->d_binmeta, nullptr);
    if (rc) return rc;
    c->launches++;
    CU(cudaEventRecord(c->ev[1], c->stream));
    return P3_OK;
}
// Human-scale inputs: the owner cannot hold every received record until one insert sweep (16 B per record). The
// count then runs in ROUNDS of chunks — p3_mg_count_finish after the last chunk of every round, this call between two
// rounds: the inserted records are forgotten (only their total is kept) and the bins start empty again. The table pays
// one pass through L2 per round; the verdict stage re-bins the reads round by round (p3_mg_cover_rebin_*).
int p3_mg_count_next_round(p3_ctx *c) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_count_next_round");
    if (rc) return rc;
    MgState &m = *mp;
    // the bin ends as p3_mg_count_finish saved them (d_cursor itself is scratch of the verdict stage's clears in between)
    accum_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_binmeta, c->parts, m.part_cap, c->d_stats);
    init_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, c->parts, m.part_cap);
    c->launches += 2;
    CU(cudaGetLastError());
    m.multi_round = true; m.round_bins_valid = false;
    return P3_OK;
}
// waits for the stage and checks it; the counts of the owned keys are final afterwards
int p3_mg_count_end(p3_ctx *c) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_count_end");
    if (rc) return rc;
    std::vector<unsigned long long> ends(c->parts);
    CU(cudaMemcpyAsync(ends.data(), c->d_binmeta, sizeof(unsigned long long) * c->parts, cudaMemcpyDeviceToHost, c->stream));
    rc = pull_stats(c);
    if (rc) return rc;
    if (c->h_stats.err_peer_timeout) return fail(P3_ERR_CUDA, "multi-GPU barrier timed out waiting for a peer rank");
    if (c->h_stats.err_bin_overflow) return fail(P3_ERR_TABLE_FULL, "multi-GPU count: a receive region or partition bin overflowed (skewed keys: raise set_bytes / owner_positions)");
    if (c->h_stats.err_table_full) return fail(P3_ERR_TABLE_FULL, "21-mer count table full: raise table_slots");
    if (c->h_stats.err_ovf_full) return fail(P3_ERR_TABLE_FULL, "count overflow side table full");
    c->binned_pos = c->h_stats.owned_pos;     // earlier rounds (p3_mg_count_next_round)
    for (uint32_t p = 0; p < c->parts; p++) c->binned_pos += ends[p] - (unsigned long long)p * mp->part_cap;
    c->h_stats.n_pos21 = c->binned_pos;
    CU(cudaEventElapsedTime(&c->ms[0], c->ev[0], c->ev[1]));
    CU(cudaEventElapsedTime(&c->ms_sub[2], c->ev[10], c->ev[1]));
    c->ms_sub[0] = 0; c->ms_sub[1] = c->ms[0] - c->ms_sub[2];
    c->bins_valid = !mp->multi_round;   // several rounds: the verdict stage bins the records again, round by round
    c->have_counts = true;
    c->have_bf = c->have_solid = c->have_adj = false;
    return P3_OK;
}

// ---- B1: MakeBF's coverage test, reference src/MakeBloomFilter.cpp:52-58 -----------------------------------------
static int ensure_planes(p3_ctx *c) {
    uint64_t pw = c->n_words + 1;
    if (!c->d_good21 || !c->d_solid || c->cap_planes < sizeof(uint32_t) * pw) {
        dfree(c->d_good21); dfree(c->d_solid);
        CU(cudaMalloc(&c->d_good21, sizeof(uint32_t) * pw));
        CU(cudaMalloc(&c->d_solid, sizeof(uint32_t) * pw));
        c->cap_planes = sizeof(uint32_t) * pw;
    }
    CU(ensure(c->d_seed, c->cap_seed, sizeof(int64_t) * std::max<uint64_t>(c->n_reads, 1)));
    return P3_OK;
}
// coverage plane := every valid 21-mer position. The owners send their verdicts in *n_slices rounds so that a round
// fits the receive regions; owner_distinct = the largest number of distinct owned 21-mers of any rank (so that every
// rank computes the same number of rounds).
static int mg_cover_begin(p3_ctx *c, uint32_t cov_threshold, uint64_t owner_distinct, uint32_t *n_slices, bool keyed) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_cover_begin");
    if (rc) return rc;
    MgState &m = *mp;
    if (keyed ? m.key_rounds < 2 : (!c->have_counts || (!c->bins_valid && !m.multi_round))) return fail(P3_ERR_STATE, "p3_mg_cover_begin: run the count stage first");
    rc = ensure_planes(c);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_good21, c->d_valid, sizeof(uint32_t) * c->n_words, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_good21 + c->n_words, 0, sizeof(uint32_t), c->stream));
    m.capB = region_cap(m, 8);
    // occurrences below the threshold per owner: at most (thr - 1) per distinct key; a source's share of a slice must fit a region
    const uint64_t worst = std::max<uint64_t>(owner_distinct * std::max<uint32_t>(cov_threshold > 1 ? cov_threshold - 1 : 1, 1), 1);
    const uint64_t per_slice = std::max<uint64_t>(m.capB * m.n_ranks / 5 * 4, 1);   // 25 % slack for the spread over the sources
    uint64_t want_slices = std::max<uint64_t>((worst + per_slice - 1) / per_slice, 1);
    want_slices = std::max<uint64_t>(want_slices, (worst + (1ull << 28) - 1) >> 28);   // the owner's verdict list stays below 2^28 records (2 GB)
    if (const char *e = getenv("P3_MG_COVER_SLICES")) want_slices = std::max<uint64_t>(want_slices, strtoull(e, nullptr, 10));   // test knob
    m.n_slices = (uint32_t)std::min<uint64_t>(want_slices, keyed ? std::max<uint32_t>(c->parts / m.key_rounds, 1) : c->parts);
    CU(ensure(m.d_sing, m.cap_sing, sizeof(uint64_t) * std::min<uint64_t>(worst / m.n_slices * 5 / 4 + (1u << 20), m.capB * m.n_ranks)));
    if (n_slices) *n_slices = m.n_slices;
    if (c->bins_valid) {
        rc = below_bits(c, cov_threshold);      // "count < threshold" of every owned slot as one bit: what the verdict sweep asks
        if (rc) return rc;
    }
    if (!keyed) CU(cudaMemsetAsync(&c->d_stats->err_bin_overflow, 0, sizeof(unsigned), c->stream));   // keyed: the count is still running, p3_mg_count_end reads the flag
    CU(cudaEventRecord(c->ev[2], c->stream));
    return P3_OK;
}
int p3_mg_cover_begin(p3_ctx *c, uint32_t cov_threshold, uint64_t owner_distinct, uint32_t *n_slices) {
    return mg_cover_begin(c, cov_threshold, owner_distinct, n_slices, false);
}
// key-range rounds: once, after the FIRST round's p3_mg_count_finish (the valid plane is complete then: every round scans
// all reads). owner_distinct = upper estimate of the distinct keys ONE round inserts on any rank.
int p3_mg_cover_begin_keyed(p3_ctx *c, uint32_t cov_threshold, uint64_t owner_distinct, uint32_t *n_slices) {
    return mg_cover_begin(c, cov_threshold, owner_distinct, n_slices, true);
}
// key-range rounds, after every round's p3_mg_count_finish: the counts of the round's partitions are final; their
// "count < threshold" bits are taken and the round's bins (records + index stream) feed the verdict slices
int p3_mg_cover_key_round(p3_ctx *c, uint32_t cov_threshold) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_cover_key_round");
    if (rc) return rc;
    if (mp->key_rounds < 2) return fail(P3_ERR_STATE, "p3_mg_cover_key_round: key-range rounds only");
    rc = below_bits(c, cov_threshold);
    if (rc) return rc;
    mp->round_bins_valid = true;
    return P3_OK;
}
// Several insert rounds (p3_mg_count_next_round): the bins no longer hold the records, so every round of chunks is sent
// and sorted into the bins AGAIN (p3_mg_count_send / _recv between these two calls), and the round's verdict slices
// (p3_mg_cover_send / _recv) look the keys up in the finished table instead of following the insert's index stream.
int p3_mg_cover_rebin_begin(p3_ctx *c) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_cover_rebin_begin");
    if (rc) return rc;
    if (!c->have_counts || !mp->multi_round) return fail(P3_ERR_STATE, "p3_mg_cover_rebin_begin: only after a count in several rounds");
    init_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, c->parts, mp->part_cap);
    c->launches++;
    CU(cudaGetLastError());
    return P3_OK;
}
int p3_mg_cover_rebin_end(p3_ctx *c) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_cover_rebin_end");
    if (rc) return rc;
    if (!c->have_counts || !mp->multi_round) return fail(P3_ERR_STATE, "p3_mg_cover_rebin_end: only after a count in several rounds");
    check_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, c->parts, mp->part_cap, c->d_stats);
    CU(cudaMemcpyAsync(c->d_binmeta, c->d_cursor, sizeof(unsigned long long) * c->parts, cudaMemcpyDeviceToDevice, c->stream));
    c->launches++;
    CU(cudaGetLastError());
    return P3_OK;
}
// owner side, round `slice`: sweep the bins of its share of the table partitions; the positions of keys whose
// final count < cov_threshold go to the regions of their source ranks (receive set slice % 2)
int p3_mg_cover_send(p3_ctx *c, uint32_t cov_threshold, uint32_t slice) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_cover_send");
    if (rc) return rc;
    MgState &m = *mp;
    if (slice >= m.n_slices) return fail(P3_ERR_ARG, "p3_mg_cover_send: slice out of range");
    const int set = (int)(slice & 1);
    // the slice's share of the partitions whose bins are there: all of them, or the current key-range round's
    const uint32_t pa = m.key_rounds > 1 ? m.p_lo : 0, pn = m.key_rounds > 1 ? m.p_hi - m.p_lo : c->parts;
    const uint32_t p0 = pa + (uint32_t)((uint64_t)pn * slice / m.n_slices), p1 = pa + (uint32_t)((uint64_t)pn * (slice + 1) / m.n_slices);
    CU(cudaMemsetAsync(m.d_sent, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    CU(cudaMemsetAsync(&c->d_stats->n_export, 0, sizeof(unsigned long long), c->stream));
    if (p1 > p0) {
        // verdict sweep over partitions [p0, p1): tiles are handed out from st->work, which starts at the slice's first record
        const unsigned long long first = (unsigned long long)p0 * m.part_cap;
        CU(cudaMemcpyAsync(&c->d_stats->work, &first, sizeof(first), cudaMemcpyHostToDevice, c->stream));
        const uint64_t n_end = (uint64_t)p1 * m.part_cap;
        const size_t smem = (size_t)kBinThreads * kPosKpt * 6 + 20;
        const uint64_t T = (uint64_t)kBinThreads * kPosKpt;
        unsigned blocks = (unsigned)std::min<uint64_t>((n_end - first + T - 1) / T, (uint64_t)c->n_sm * 4);
        if (c->bins_valid || m.round_bins_valid)
            pos_bin_kernel<0, true><<<blocks, kBinThreads, smem, c->stream>>>(c->table(), m.bk(c), m.bw(c), m.bi(c), c->d_below, n_end, m.part_cap, c->d_binmeta, nullptr,
                                                                       cov_threshold, c->ovf(), c->d_stats, 27, 1, nullptr, m.cap_sing / sizeof(uint64_t), nullptr, m.d_sing);
        else   // re-binned round: no index stream, the keys are looked up in the table
            pos_bin_kernel<0, false><<<blocks, kBinThreads, smem, c->stream>>>(c->table(), c->d_bkeys, c->d_bword, nullptr, nullptr, n_end, m.part_cap, c->d_binmeta, nullptr,
                                                                        cov_threshold, c->ovf(), c->d_stats, 27, 1, nullptr, m.cap_sing / sizeof(uint64_t), nullptr, m.d_sing);
        PeerOut64 po;
        for (uint32_t j = 0; j < (uint32_t)kMaxPeers; j++)
            po.p[j] = j < m.n_ranks ? reinterpret_cast<uint64_t *>(dest_block(m, j, set)) + dest_region(m, j) * m.capB : nullptr;
        const uint64_t nmax = m.cap_sing / sizeof(uint64_t);
        unsigned sblocks = (unsigned)std::min<uint64_t>((nmax + kTilePos - 1) / kTilePos, (uint64_t)c->n_sm * 3);
        scatter_pos_peer_kernel<<<sblocks, kScatterThreads, kSmemP, c->stream>>>(m.d_sing, nmax, &c->d_stats->n_export, m.n_ranks, m.d_sent, po, m.capB, c->d_stats);
        c->launches += 2;
        CU(cudaGetLastError());
    }
    return mg_publish(c, m, set);
}
// source side: clear the bits of the positions that arrived in receive set slice % 2
int p3_mg_cover_recv(p3_ctx *c, uint32_t slice) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_cover_recv");
    if (rc) return rc;
    MgState &m = *mp;
    const int set = (int)(slice & 1);
    ClearInput in;
    in.rec = reinterpret_cast<const uint64_t *>(m.set_ptr(m.my_rank, set));
    in.n = (uint64_t)m.n_ranks * m.capB; in.in_cap = m.capB; in.in_end = &m.ctl()->count[set][0];
    // what a round may deliver: the regions' capacity at most, normally this rank's share of the owners' verdicts
    // distinct keys the verdicts can come from: known after the count, or (key-range rounds: the count is still running) what one
    // round's share of the table can hold
    const uint64_t n_keys = m.key_rounds > 1 ? (uint64_t)((double)c->nb * 4 * 0.7 / m.key_rounds) : c->h_stats.n_cand;
    const uint64_t n_expect = std::min<uint64_t>(in.n, (n_keys / std::max<uint32_t>(m.n_slices, 1)) * 3 / 2 + (1u << 20));
    rc = plane_clear_job<1>(c, in, n_expect, 0, c->d_good21, c->n_words * 32, false, nullptr);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[3], c->stream));
    return P3_OK;
}

// ---- B2: solid k-mers to their owners, reference src/MakeBloomFilter.cpp:60-83 ------------------------------------
// solid plane (RMQ test as a window AND of the coverage plane) and seeds are local; sizes the owned set (owned_slots)
// and turns the count bins into k-mer bins. k <= 32 (multi-word k-mers are single-GPU in this build).
int p3_mg_solid_begin(p3_ctx *c, uint32_t k, uint64_t owned_slots) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_solid_begin");
    if (rc) return rc;
    MgState &m = *mp;
    if (!c->d_good21) return fail(P3_ERR_STATE, "p3_mg_solid_begin: run the coverage stage first");
    if (k < P3_MIN_K || k > 32) return fail(P3_ERR_ARG, "p3_mg_solid_begin: k outside [21,32] (multi-word k-mers: p3_mg_long_*)");
    c->k = k; m.k = k; c->set_valid = false; c->hints_valid = false; c->d_set_b = nullptr; c->nbs_b = 0; c->parts_b = 1;
    CU(cudaEventRecord(c->ev[4], c->stream));
    CU(cudaMemsetAsync(&c->d_stats->n_adds, 0, sizeof(unsigned long long) * 5, c->stream));
    CU(cudaMemsetAsync(&c->d_stats->err_table_full, 0, sizeof(unsigned), c->stream));
    CU(cudaMemsetAsync(c->d_solid + c->n_words, 0, sizeof(uint32_t), c->stream));
    solid_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_good21, c->n_words, (int)k, c->d_solid, c->d_stats);
    seeds_kernel<<<c->grid(4), 256, 0, c->stream>>>(c->d_off, c->n_reads, c->d_solid, (int)k, c->d_seed);
    c->launches += 2;
    // the owned set: partitions of ~24 MB (at most 96, see dedupe_solid_positions), list, hints, adjacency
    uint64_t buckets = (std::max<uint64_t>(owned_slots, 1024) + 3) / 4;
    uint64_t want = (buckets * 32 + (24ull << 20) - 1) / (24ull << 20);
    if (const char *e = getenv("P3_SET_PARTS")) want = strtoull(e, nullptr, 10);
    uint32_t P = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 1), 96);
    uint64_t nbp = std::max<uint64_t>((buckets + P - 1) / P, 1);
    uint64_t nbs = nbp * P;
    if (!c->d_set || c->nbs != nbs) {
        dfree(c->d_set); dfree(c->d_list);
        if (cudaMalloc(&c->d_set, nbs * 32) != cudaSuccess || cudaMalloc(&c->d_list, nbs * 32) != cudaSuccess) {
            cudaGetLastError();
            return fail(P3_ERR_NOMEM, "owned k-mer set allocation failed");
        }
        c->nbs = nbs; c->list_cap = nbs * 4;
    }
    c->set_parts = P;
    CU(ensure(c->d_hint, c->cap_hint, nbs * 4));
    if (!c->d_adj || c->adj_cap < c->list_cap) { dfree(c->d_adj); c->adj_cap = c->list_cap; CU(cudaMalloc(&c->d_adj, c->adj_cap)); }
    CU(cudaMemsetAsync(c->d_set, 0xFF, nbs * 32, c->stream));
    CU(cudaMemsetAsync(c->d_hint, 0, nbs * 4, c->stream));
    // k-mer bins in the (now idle) count bins: 8-byte k-mers in d_bkeys, one hint byte each in d_bword
    c->bins_valid = false;
    const uint64_t room = std::min<uint64_t>(c->cap_bkeys / 8, c->cap_bword);
    m.kpart_cap = room / P / kSweepChunk * kSweepChunk;
    if (m.kpart_cap == 0) return fail(P3_ERR_STATE, "p3_mg_solid_begin: no bins (run the count stage first)");
    m.capK = region_cap(m, 9);
    init_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, m.kpart_cap);
    c->launches++;
    CU(cudaGetLastError());
    return P3_OK;
}
// the solid occurrences of chunk `ch` (canonical k-mer + adjacency hint) to the owners' receive set ch % 2
int p3_mg_solid_send(p3_ctx *c, uint64_t ch) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_solid_send");
    if (rc) return rc;
    MgState &m = *mp;
    const int set = (int)(ch & 1);
    const uint64_t w0 = std::min<uint64_t>(ch * m.chunk_words, c->n_words), w1 = std::min<uint64_t>((ch + 1) * m.chunk_words, c->n_words);
    CU(cudaMemsetAsync(m.d_sent, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    if (w1 > w0) {
        PeerOutK po;
        for (uint32_t j = 0; j < (uint32_t)kMaxPeers; j++) {
            po.keys[j] = nullptr; po.hints[j] = nullptr;
            if (j < m.n_ranks) {
                unsigned char *blk = dest_block(m, j, set);
                const uint64_t reg = dest_region(m, j);
                po.keys[j] = reinterpret_cast<uint64_t *>(blk) + reg * m.capK;
                po.hints[j] = blk + (uint64_t)m.n_ranks * m.capK * 8 + reg * m.capK;
            }
        }
        const unsigned sblocks = (unsigned)std::min<uint64_t>((w1 - w0 + kTileWords - 1) / kTileWords, (uint64_t)c->n_sm * 3);
        if (c->d_nmask) scatter_kmer_kernel<true, true><<<sblocks, kScatterThreads, kSmemPL, c->stream>>>(c->d_packed, c->d_nmask, c->d_solid, w0, w1, (int)m.k, m.n_ranks, m.d_sent, nullptr, nullptr, m.capK, c->d_stats, po);
        else scatter_kmer_kernel<false, true><<<sblocks, kScatterThreads, kSmemPL, c->stream>>>(c->d_packed, nullptr, c->d_solid, w0, w1, (int)m.k, m.n_ranks, m.d_sent, nullptr, nullptr, m.capK, c->d_stats, po);
        c->launches++;
        CU(cudaGetLastError());
    }
    return mg_publish(c, m, set);
}
// owner side: sort the occurrences that arrived in receive set ch % 2 into the set-partition bins
int p3_mg_solid_recv(p3_ctx *c, uint64_t ch) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_solid_recv");
    if (rc) return rc;
    MgState &m = *mp;
    const int set = (int)(ch & 1);
    unsigned char *blk = m.set_ptr(m.my_rank, set);
    const uint64_t n = (uint64_t)m.n_ranks * m.capK;
    const unsigned sblocks = (unsigned)std::min<uint64_t>((n + kTilePos - 1) / kTilePos, (uint64_t)c->n_sm * 3);
    scatter_rec_kernel<4, 1><<<sblocks, kScatterThreads, kSmemA1, c->stream>>>(
        reinterpret_cast<const uint64_t *>(blk), blk + n * 8, n, c->set_parts, c->d_cursor, c->d_bkeys, c->d_bword, m.kpart_cap, nullptr,
        m.capK, &m.ctl()->count[set][0], c->d_stats);
    c->launches++;
    CU(cudaGetLastError());
    return P3_OK;
}
// between two rounds of chunks (human-scale inputs: the k-mer bins live in the count bins, which hold one round): the
// round's occurrences are de-duplicated into the owned set and the bins start empty again
int p3_mg_solid_next_round(p3_ctx *c) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_solid_next_round");
    if (rc) return rc;
    MgState &m = *mp;
    const uint32_t P = c->set_parts;
    check_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, m.kpart_cap, c->d_stats);
    CU(cudaMemsetAsync(&c->d_stats->work, 0, sizeof(unsigned long long), c->stream));
    set_sweep_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_bkeys, reinterpret_cast<const uint8_t *>(c->d_bword), (uint64_t)P * m.kpart_cap, m.kpart_cap,
                                                      c->d_cursor, c->kset(), c->d_hint, c->d_stats);
    init_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, m.kpart_cap);
    c->launches += 3;
    CU(cudaGetLastError());
    return P3_OK;
}
// after the last chunk: one L2-resident de-duplication sweep (hints OR-ed per k-mer), then set -> list + hint bytes
int p3_mg_solid_finish(p3_ctx *c) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_solid_finish");
    if (rc) return rc;
    MgState &m = *mp;
    const uint32_t P = c->set_parts;
    check_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, m.kpart_cap, c->d_stats);
    CU(cudaMemsetAsync(&c->d_stats->work, 0, sizeof(unsigned long long), c->stream));
    set_sweep_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_bkeys, reinterpret_cast<const uint8_t *>(c->d_bword), (uint64_t)P * m.kpart_cap, m.kpart_cap,
                                                      c->d_cursor, c->kset(), c->d_hint, c->d_stats);
    CU(cudaMemsetAsync(&c->d_stats->n_distinct_solid, 0, sizeof(unsigned long long), c->stream));
    compact_set_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_set, c->nbs * 4, c->d_list, c->list_cap, c->d_stats,
                                                         reinterpret_cast<const uint8_t *>(c->d_hint), c->d_adj);
    c->launches += 3;
    CU(cudaGetLastError());
    return P3_OK;
}
// waits for the stage; this context then holds its OWNED distinct solid k-mers (list + hints) and an empty filter of at
// least filter_words_cap words (room for whole shards), ready for p3_mg_bloom_bin / p3_mg_bloom_direct
int p3_mg_solid_end(p3_ctx *c, uint64_t filter_size, uint32_t num_hashes, uint64_t filter_words_cap, uint64_t *n_adds, uint64_t *n_owned) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_solid_end");
    if (rc) return rc;
    rc = pull_stats(c);
    if (rc) return rc;
    if (c->h_stats.err_peer_timeout) return fail(P3_ERR_CUDA, "multi-GPU barrier timed out waiting for a peer rank");
    if (c->h_stats.err_bin_overflow) return fail(P3_ERR_TABLE_FULL, "multi-GPU MakeBF: a receive region or bin overflowed (raise set_bytes)");
    if (c->h_stats.err_table_full) return fail(P3_ERR_TABLE_FULL, "owned k-mer set full: raise owned_slots");
    rc = alloc_bloom(c, mp->k, filter_size, num_hashes, filter_words_cap);
    if (rc) return rc;
    CU(cudaMemsetAsync(c->d_bloom, 0, sizeof(uint32_t) * c->bloom_words, c->stream));
    CU(cudaEventRecord(c->ev[5], c->stream));
    CU(cudaEventRecord(c->ev[14], c->stream));
    c->have_bf = true; c->have_solid = true; c->have_adj = false;
    c->set_valid = c->hints_valid = mp->k <= 32;     // multi-word k-mers: the set holds store indices, there are no hints
    if (n_adds) *n_adds = c->h_stats.n_adds;
    if (n_owned) *n_owned = c->h_stats.n_distinct_solid;
    return P3_OK;
}

// ---- B3: sharded BF.add, reference src/bloomfilter.cpp:69-74 ---------------------------------------------------------
uint64_t p3_bloom_seg_bits(void) { return 1ull << bloom_seg_shift(); }

// buffer that receives the binned bit indices of the segments this rank owns (peers store into it;
// export with p3_ipc_export). Grow-only; an outgrown buffer is kept until the context is destroyed
// because peers may still have it mapped.
int p3_mg_bloom_buffer(p3_ctx *c, uint64_t n_u32, uint32_t **d_buf) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    BloomBinState &b = g_bbin.get(c);
    uint64_t need = sizeof(uint32_t) * std::max<uint64_t>(n_u32, 1);
    if (!b.d_bins || b.cap_bins < need) {
        if (b.d_bins) b.graveyard.push_back(b.d_bins);
        b.d_bins = nullptr; b.cap_bins = 0;
        if (cudaMalloc((void **)&b.d_bins, need) != cudaSuccess) { cudaGetLastError(); return fail(P3_ERR_NOMEM, "bloom bin buffer allocation failed"); }
        b.cap_bins = need;
    }
    if (d_buf) *d_buf = b.d_bins;
    return P3_OK;
}
// source side: the num_hashes bit indices of every owned k-mer, binned by filter segment and stored to
// h_segbase[s] (device addresses, possibly peer memory; cap records each). h_counts[s] = records
// written (a count above cap means that segment overflowed: use p3_mg_bloom_direct on all ranks).
int p3_mg_bloom_bin(p3_ctx *c, uint32_t n_seg, const uint64_t *h_segbase, uint64_t cap, uint64_t *h_counts) {
    return p3_mg_bloom_bin_range(c, n_seg, h_segbase, cap, h_counts, 0, ~0ULL);
}
// the same for k-mers [first, first + count) of the owned list only: the shard owners' buffers then need room for one
// PASS over the list at a time (human scale: 19 indices of 3.2 G k-mers are 240 GB over all ranks when binned at once)
int p3_mg_bloom_bin_range(p3_ctx *c, uint32_t n_seg, const uint64_t *h_segbase, uint64_t cap, uint64_t *h_counts, uint64_t first, uint64_t count) {
    if (!c || !c->have_bf || !h_segbase || !h_counts) return fail(P3_ERR_STATE, "p3_mg_bloom_bin: run p3_mg_solid_end first");
    if (n_seg == 0 || n_seg > (uint32_t)kMaxParts || c->num_hashes > (uint32_t)kBinMaxHashes)
        return fail(P3_ERR_ARG, "p3_mg_bloom_bin: too many segments / hash functions for the binned path");
    CU(cudaSetDevice(c->device));
    const uint64_t n = c->h_stats.n_distinct_solid;
    first = std::min(first, n);
    count = std::min(count, n - first);
    uint64_t *d_hh = nullptr;
    int rc = bloom_hash_list(c, n, &d_hh);
    if (rc) return rc;
    rc = bloom_bin_launch(c, d_hh + 2 * first, count, n_seg, bloom_seg_shift(), h_segbase, cap, h_counts);
    CU(cudaMemsetAsync(&c->d_stats->err_bin_overflow, 0, sizeof(unsigned), c->stream));   // the caller judges overflow from the counts
    return rc;
}
// owner side: OR the received records of local segments [seg_first, seg_first + n_local) into this
// context's filter; h_ptr / h_n [n_local][n_src] = where source r's records of the segment are and how many
int p3_mg_bloom_apply(p3_ctx *c, uint64_t seg_first, uint32_t n_local, uint32_t n_src, const uint64_t *h_ptr, const uint64_t *h_n) {
    if (!c || !c->have_bf) return fail(P3_ERR_STATE, "p3_mg_bloom_apply: no filter");
    if (n_src > (uint32_t)kMaxRegions) return fail(P3_ERR_ARG, "p3_mg_bloom_apply: at most 16 sources");
    CU(cudaSetDevice(c->device));
    const int shift = bloom_seg_shift();
    for (uint32_t s = 0; s < n_local; s++) {
        if (((seg_first + s + 1) << (shift - 5)) > c->bloom_words) return fail(P3_ERR_ARG, "p3_mg_bloom_apply: segment outside the filter allocation");
        ApplyRegions rg; rg.count = (int)n_src;
        for (uint32_t r = 0; r < n_src; r++) { rg.ptr[r] = (const uint32_t *)(uintptr_t)h_ptr[s * n_src + r]; rg.n[r] = h_n[s * n_src + r]; }
        int rc = bloom_apply_launch(c, seg_first + s, shift, rg);
        if (rc) return rc;
    }
    return P3_OK;
}
// fallback of the sharded adds: every owned k-mer straight into this rank's full copy (then OR-reduce)
int p3_mg_bloom_direct(p3_ctx *c) {
    if (!c || !c->have_bf) return fail(P3_ERR_STATE, "p3_mg_bloom_direct: no filter");
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->d_bloom, 0, sizeof(uint32_t) * c->bloom_words, c->stream));
    return bloom_add_direct(c, c->h_stats.n_distinct_solid);
}
int p3_mg_filter(p3_ctx *c, uint32_t **d_bits, uint64_t *n_words) {
    if (!c || !c->d_bloom) return fail(P3_ERR_STATE, "p3_mg_filter: no filter");
    if (d_bits) *d_bits = c->d_bloom;
    if (n_words) *n_words = c->bloom_words;
    return P3_OK;
}
// marks the end of the MakeBF stage for p3_stage_ms (call once the filter is complete, before p3_dbg_adjacency)
int p3_mg_makebf_done(p3_ctx *c) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->ev[15], c->stream));
    CU(cudaEventRecord(c->ev[6], c->stream));
    return P3_OK;
}
// bytes of device memory in use on this context's GPU (total - free), for the benchmark's hbm_peak_bytes
uint64_t p3_device_mem_used(p3_ctx *c) {
    if (!c) return 0;
    size_t fr = 0, tot = 0;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return 0; }
    return (uint64_t)(tot - fr);
}

}  // extern "C"

// ---- B2 for multi-word k-mers (33 <= k <= 3001, W = ceil(2k/64) words), reference src/MakeBloomFilter.cpp:60-83 with
// std::bitset<2k> k-mers (src/Assemble.cpp:30-53) ---------------------------------------------------------------------
// A solid occurrence travels as its W canonical words (8W bytes; the owner is a hash of std::hash(k-mer), the value the
// Bloom hashes start from). The owner appends what arrives to a STORE of W-word records and, after the last chunk,
// de-duplicates the store the way the single-GPU path de-duplicates read positions (p3_long.inc.cu): set slots hold
// [hash tag:24 | store index:40]; a tag hit is verified by comparing the two records' words — exact. The distinct
// k-mers become the context's n x W word array, on which the sharded BF.add (its (h1, h2) pairs) and CheckDirections
// run unchanged. No adjacency hints for multi-word k-mers.
// source side: every solid occurrence of words [w0, w1) -> W canonical words into region `me` of its owner's receive
// set. One word per thread; a block counts its occurrences per destination, claims room once per destination, then
// writes (the canonical choice and the hash are computed in both passes: nothing per occurrence is kept in registers).
__global__ void __launch_bounds__(256)
long_scatter_kernel(Stream2 st, const uint32_t *__restrict__ solid, uint64_t w0, uint64_t w1, LongK L, uint32_t n_ranks,
                    unsigned long long *sent, PeerOut64 out, uint64_t cap, Stats *stt) {
    __shared__ unsigned s_cnt[kMaxPeers];
    __shared__ unsigned long long s_base[kMaxPeers];
    const int tid = threadIdx.x;
    const uint64_t n_tiles = (w1 - w0 + 255) / 256;
    bool over = false;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (tid < kMaxPeers) s_cnt[tid] = 0;
        __syncthreads();
        const uint64_t w = w0 + tile * 256 + tid;
        const uint32_t sbits = w < w1 ? __ldg(solid + w) : 0u;
        for (uint32_t s = sbits; s;) {
            const int o = __clz(s);
            s &= ~(0x80000000u >> o);
            const uint64_t p = w * 32 + o;
            const bool rc = occ_use_rc(st, p, L);
            const uint64_t h0 = hash_words(L, [&](int j) { return occ_word(st, p, L, j, rc); });
            atomicAdd(&s_cnt[owner_of(h0, n_ranks)], 1u);
        }
        __syncthreads();
        if (tid < (int)n_ranks) {
            const unsigned cnt = s_cnt[tid];
            s_base[tid] = cnt ? atomicAdd(&sent[tid], (unsigned long long)cnt) : 0ULL;
            s_cnt[tid] = 0;
        }
        __syncthreads();
        for (uint32_t s = sbits; s;) {
            const int o = __clz(s);
            s &= ~(0x80000000u >> o);
            const uint64_t p = w * 32 + o;
            const bool rc = occ_use_rc(st, p, L);
            const uint64_t h0 = hash_words(L, [&](int j) { return occ_word(st, p, L, j, rc); });
            const uint32_t dst = owner_of(h0, n_ranks);
            const unsigned long long slot = s_base[dst] + atomicAdd(&s_cnt[dst], 1u);
            if (slot < cap) {
                uint64_t *q = out.p[dst] + slot * (uint64_t)L.W;
                for (int j = 0; j < L.W; j++) q[j] = occ_word(st, p, L, j, rc);
            } else over = true;
        }
        __syncthreads();
    }
    if (over) atomicExch(&stt->err_bin_overflow, 1u);
}

// owner side: append the records of the n_ranks regions (cap records of W words each, counts[r] filled) to the store.
// Every block reads the store size as it was when the kernel started; long_advance_kernel adds the total afterwards.
__global__ void __launch_bounds__(256)
long_gather_kernel(const uint64_t *__restrict__ recv, uint64_t cap, int W, int n_ranks, const unsigned long long *__restrict__ counts,
                   uint64_t *__restrict__ store, const unsigned long long *__restrict__ store_n, uint64_t store_cap, Stats *stt) {
    const uint64_t base = *store_n;
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x, t0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t pre = 0;
    bool over = false;
    for (int r = 0; r < n_ranks; r++) {
        const uint64_t n = min((uint64_t)__ldcg(counts + r), cap);
        const uint64_t room = base + pre < store_cap ? store_cap - (base + pre) : 0;
        const uint64_t m = min(n, room);
        if (m < n) over = true;
        const uint64_t *src = recv + (uint64_t)r * cap * W;
        uint64_t *dst = store + (base + pre) * W;
        for (uint64_t i = t0; i < m * W; i += stride) dst[i] = __ldcs(src + i);
        pre += n;
    }
    if (over && t0 == 0) atomicExch(&stt->err_bin_overflow, 1u);
}
__global__ void long_advance_kernel(int n_ranks, uint64_t cap, const unsigned long long *__restrict__ counts, unsigned long long *store_n, uint64_t store_cap) {
    unsigned long long tot = *store_n;
    for (int r = 0; r < n_ranks; r++) tot += min((unsigned long long)counts[r], (unsigned long long)cap);
    *store_n = min(tot, (unsigned long long)store_cap);
}

// de-duplication of the store: slot = [tag:24 | record index:40]
__global__ void __launch_bounds__(256)
long_dedupe_store_kernel(const uint64_t *__restrict__ store, const unsigned long long *__restrict__ n_dev, LongK L, uint64_t *set, uint64_t nbs, Stats *stt) {
    const uint64_t n = *n_dev;
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    bool full = false;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t *kw = store + i * L.W;
        const uint64_t h0 = hash_words(L, [&](int j) { return __ldg(kw + j); });
        const uint64_t tag = h0 >> 40;
        const uint64_t val = (tag << 40) | i;
        uint64_t b = __umul64hi(fmix64(h0), nbs);
        bool done = false;
        for (uint64_t probe = 0; probe < nbs && probe < kMaxProbe && !done; probe++) {
            uint64_t *bp = set + 4 * b;
            uint64_t sl[4];
            ld_bucket(bp, sl);
#pragma unroll
            for (int q = 0; q < 4 && !done; q++) {
                uint64_t v = sl[q];
                if (v == kEmpty) {
                    v = atomicCAS(ull(bp + q), kEmpty, val);
                    if (v == kEmpty) { done = true; break; }
                }
                if ((v >> 40) == tag) {
                    const uint64_t *ow = store + (v & kPos40) * L.W;
                    bool eq = true;
                    for (int j = 0; j < L.W && eq; j++) eq = __ldg(kw + j) == __ldg(ow + j);
                    if (eq) done = true;
                }
            }
            b = (b + 1 == nbs) ? 0 : b + 1;
        }
        if (!done) full = true;
    }
    if (full) atomicExch(&stt->err_table_full, 1u);
}
__global__ void long_materialise_store_kernel(const uint64_t *__restrict__ store, int W, const uint64_t *__restrict__ list, const unsigned long long *__restrict__ n_dev,
                                              uint64_t n_cap, uint64_t *__restrict__ words) {
    const uint64_t n = min((uint64_t)*n_dev, n_cap);
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n * W; i += stride) {
        const uint64_t r = i / W, j = i - r * W;
        words[i] = __ldg(store + (list[r] & kPos40) * W + j);
    }
}

extern "C" {

// local part of B2 for k > 32: solid plane (window of k-20 set coverage bits) and seeds; *n_adds = this rank's solid
// occurrences = the BF.add calls the reference would make for its reads (sizes the owners' stores: sum over ranks / n_ranks)
int p3_mg_long_solid(p3_ctx *c, uint32_t k, uint64_t *n_adds) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_long_solid");
    if (rc) return rc;
    if (!c->d_good21) return fail(P3_ERR_STATE, "p3_mg_long_solid: run the coverage stage first");
    if (k <= 32 || k > P3_MAX_K) return fail(P3_ERR_ARG, "p3_mg_long_solid: 33 <= k <= 3001");
    if (c->total_bases >= kPos40) return fail(P3_ERR_ARG, "k > 32 supports up to 2^40 bases per context");
    c->k = k; mp->k = k; c->set_valid = false; c->hints_valid = false; c->d_set_b = nullptr; c->nbs_b = 0; c->parts_b = 1;
    CU(cudaEventRecord(c->ev[4], c->stream));
    CU(cudaMemsetAsync(&c->d_stats->n_adds, 0, sizeof(unsigned long long) * 5, c->stream));
    CU(cudaMemsetAsync(&c->d_stats->err_table_full, 0, sizeof(unsigned), c->stream));
    CU(cudaMemsetAsync(&c->d_stats->err_bin_overflow, 0, sizeof(unsigned), c->stream));
    solid_long_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_good21, c->n_words, (int)k - kShortK + 1, c->d_solid, c->d_stats);
    seeds_kernel<<<c->grid(4), 256, 0, c->stream>>>(c->d_off, c->n_reads, c->d_solid, (int)k, c->d_seed);
    c->launches += 2;
    CU(cudaGetLastError());
    rc = pull_stats(c);
    if (rc) return rc;
    if (n_adds) *n_adds = c->h_stats.n_adds;
    return P3_OK;
}
// sizes the owner's store (owner_occurrences records) and its set (owned_slots), and fixes the chunking of the sends:
// occ_per_word = upper estimate of the solid occurrences per packed word of any rank (<= 32). *chunk_words / *n_chunks:
// what p3_mg_long_send takes (the same on every rank when max_words is the largest n_words of any rank).
int p3_mg_long_begin(p3_ctx *c, uint64_t owned_slots, uint64_t owner_occurrences, double occ_per_word, uint64_t max_words,
                     uint64_t *chunk_words, uint64_t *n_chunks) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_long_begin");
    if (rc) return rc;
    MgState &m = *mp;
    if (m.k <= 32) return fail(P3_ERR_STATE, "p3_mg_long_begin: run p3_mg_long_solid first");
    LongK L = make_longk(m.k);
    LongMg &lm = g_longmg.get(c);
    lm.W = L.W;
    lm.capL = m.set_bytes / m.n_ranks / (8ull * L.W);
    if (lm.capL < 64) return fail(P3_ERR_ARG, "p3_mg_long_begin: receive set too small for k-mers of this length (raise set_bytes)");
    // a chunk's occurrences spread evenly over the owners (a hash): a region must hold its share + 10 % + slack
    occ_per_word = std::min(32.0, std::max(occ_per_word, 0.001));
    const double room = (double)lm.capL - std::min<double>(8192.0, (double)lm.capL / 4);
    uint64_t cw = (uint64_t)(room * m.n_ranks / (occ_per_word * 1.10));
    if (cw < 32) return fail(P3_ERR_ARG, "p3_mg_long_begin: receive set too small for k-mers of this length (raise set_bytes)");
    cw = cw >= 256 ? cw / 256 * 256 : cw / 32 * 32;
    cw = std::min<uint64_t>(cw, std::max<uint64_t>((max_words + 255) / 256 * 256, 256));
    lm.chunk_words = cw;
    if (chunk_words) *chunk_words = cw;
    if (n_chunks) *n_chunks = std::max<uint64_t>((max_words + cw - 1) / cw, 1);
    lm.store_records = std::max<uint64_t>(owner_occurrences, 1024);
    CU(ensure(lm.d_store, lm.cap_store, lm.store_records * 8ull * L.W));
    if (!lm.d_store_n) CU(cudaMalloc(&lm.d_store_n, sizeof(unsigned long long)));
    CU(cudaMemsetAsync(lm.d_store_n, 0, sizeof(unsigned long long), c->stream));
    uint64_t nbs = (std::max<uint64_t>(owned_slots, 1024) + 3) / 4;
    if (!c->d_set || c->nbs != nbs) {
        dfree(c->d_set); dfree(c->d_list);
        if (cudaMalloc(&c->d_set, nbs * 32) != cudaSuccess || cudaMalloc(&c->d_list, nbs * 32) != cudaSuccess) {
            cudaGetLastError();
            return fail(P3_ERR_NOMEM, "owned k-mer set allocation failed");
        }
        c->nbs = nbs; c->list_cap = nbs * 4;
    }
    c->set_parts = 1;
    CU(cudaMemsetAsync(c->d_set, 0xFF, nbs * 32, c->stream));
    c->bins_valid = false;
    return P3_OK;
}
int p3_mg_long_send(p3_ctx *c, uint64_t ch) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_long_send");
    if (rc) return rc;
    MgState &m = *mp;
    LongMg *lm = g_longmg.find(c);
    if (!lm || !lm->chunk_words) return fail(P3_ERR_STATE, "p3_mg_long_send: run p3_mg_long_begin first");
    const int set = (int)(ch & 1);
    LongK L = make_longk(m.k);
    const uint64_t w0 = std::min<uint64_t>(ch * lm->chunk_words, c->n_words), w1 = std::min<uint64_t>((ch + 1) * lm->chunk_words, c->n_words);
    CU(cudaMemsetAsync(m.d_sent, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    if (w1 > w0) {
        PeerOut64 po;
        for (uint32_t j = 0; j < (uint32_t)kMaxPeers; j++)
            po.p[j] = j < m.n_ranks ? reinterpret_cast<uint64_t *>(dest_block(m, j, set)) + dest_region(m, j) * lm->capL * L.W : nullptr;
        Stream2 st; st.packed = c->d_packed; st.nmask = c->d_nmask;
        const unsigned blocks = (unsigned)std::min<uint64_t>((w1 - w0 + 255) / 256, (uint64_t)c->grid());
        long_scatter_kernel<<<blocks, 256, 0, c->stream>>>(st, c->d_solid, w0, w1, L, m.n_ranks, m.d_sent, po, lm->capL, c->d_stats);
        c->launches++;
        CU(cudaGetLastError());
    }
    return mg_publish(c, m, set);
}
int p3_mg_long_recv(p3_ctx *c, uint64_t ch) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_long_recv");
    if (rc) return rc;
    MgState &m = *mp;
    LongMg *lm = g_longmg.find(c);
    if (!lm || !lm->chunk_words) return fail(P3_ERR_STATE, "p3_mg_long_recv: run p3_mg_long_begin first");
    const int set = (int)(ch & 1);
    const uint64_t *recv = reinterpret_cast<const uint64_t *>(m.set_ptr(m.my_rank, set));
    const unsigned long long *counts = &m.ctl()->count[set][0];
    long_gather_kernel<<<c->grid(4), 256, 0, c->stream>>>(recv, lm->capL, lm->W, (int)m.n_ranks, counts, lm->d_store, lm->d_store_n, lm->store_records, c->d_stats);
    long_advance_kernel<<<1, 1, 0, c->stream>>>((int)m.n_ranks, lm->capL, counts, lm->d_store_n, lm->store_records);
    c->launches += 2;
    CU(cudaGetLastError());
    return P3_OK;
}
// after the last chunk: de-duplicate the store, materialise the distinct k-mers as this context's n x W word array
// (enqueue only; p3_mg_solid_end waits, checks and leaves the empty filter)
int p3_mg_long_finish(p3_ctx *c) {
    MgState *mp;
    int rc = mg_ready(c, &mp, "p3_mg_long_finish");
    if (rc) return rc;
    MgState &m = *mp;
    LongMg *lm = g_longmg.find(c);
    if (!lm || !lm->d_store) return fail(P3_ERR_STATE, "p3_mg_long_finish: run p3_mg_long_begin first");
    LongK L = make_longk(m.k);
    long_dedupe_store_kernel<<<c->grid(), 256, 0, c->stream>>>(lm->d_store, lm->d_store_n, L, c->d_set, c->nbs, c->d_stats);
    CU(cudaMemsetAsync(&c->d_stats->n_distinct_solid, 0, sizeof(unsigned long long), c->stream));
    compact_set_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_set, c->nbs * 4, c->d_list, c->list_cap, c->d_stats);
    LongState &ls = g_long.get(c);
    // the distinct k-mers are at most the store's records and at most the list's capacity
    CU(ensure(ls.d_words, ls.cap_words, sizeof(uint64_t) * std::max<uint64_t>(std::min<uint64_t>(lm->store_records, c->list_cap) * L.W, 1)));
    long_materialise_store_kernel<<<c->grid(), 256, 0, c->stream>>>(lm->d_store, L.W, c->d_list, &c->d_stats->n_distinct_solid,
                                                                   std::min<uint64_t>(lm->store_records, c->list_cap), ls.d_words);
    c->launches += 3;
    CU(cudaGetLastError());
    return P3_OK;
}

}  // extern "C"
