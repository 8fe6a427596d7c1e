// p3_multi.inc.cu — multi-GPU building blocks (DESIGN.md row e), part of the p3_gpu.cu translation
// unit. One process per GPU; the exchanges themselves (all-to-all, filter OR-reduce) are done by
// the caller with torch.distributed/NCCL on the device buffers these entry points fill or read:
//
//   every rank : bin its 21-mers by OWNER rank          p3_mg_owner_hist / p3_mg_owner_scatter
//   all-to-all of the count records (12 B each)
//   owner      : binned, L2-resident insert             p3_mg_count_begin / _records / _end
//   owner      : singleton verdicts by source rank      p3_mg_singletons
//   all-to-all of the positions (8 B each)
//   every rank : coverage plane, solid plane, seeds,    p3_mg_cover_begin / _clear, p3_mg_solid_local
//                locally distinct solid k-mers
//   every rank : bin those k-mers by owner              p3_mg_kmer_owner_hist / _scatter
//   all-to-all of the k-mers (8 B each)
//   owner      : de-duplicate, BF.add into its copy     p3_mg_owned_begin / _insert / _end
//   OR-reduce of the filter copies                      p3_mg_filter gives the buffer
//   owner      : CheckDirections of its k-mers          p3_dbg_adjacency (filter is complete & local)

__global__ void __launch_bounds__(256)
singleton_list_kernel(const uint64_t *__restrict__ slots, const uint64_t *__restrict__ cand_slot,
                      const uint64_t *__restrict__ cand_pos, uint64_t n_cand, uint64_t thr, Ovf ovf,
                      Stats *st, uint64_t *__restrict__ out) {
    __shared__ unsigned s_wtot[8];
    __shared__ unsigned long long s_base;
    const unsigned n_overflow = st->n_overflow;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    uint64_t n_round = (n_cand + stride - 1) / stride;
    for (uint64_t r = 0; r < n_round; r++) {
        uint64_t j = r * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
        bool single = false;
        uint64_t pos = 0;
        if (j < n_cand) {
            uint64_t v = __ldcg(slots + __ldcs(cand_slot + j));
            uint64_t c = v >> 42;
            if (c < thr && n_overflow) c += ovf_get(ovf, v & kKey42) << 22;
            single = c < thr;
            if (single) pos = __ldcs(cand_pos + j);
        }
        unsigned m = __ballot_sync(0xffffffffu, single);
        if (lane == 0) s_wtot[wid] = __popc(m);
        __syncthreads();
        unsigned wbase = 0, total = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { unsigned t = s_wtot[q]; if (q < wid) wbase += t; total += t; }
        if (threadIdx.x == 0 && total) s_base = atomicAdd(&st->n_export, (unsigned long long)total);
        __syncthreads();
        if (single) out[s_base + wbase + __popc(m & ((1u << lane) - 1))] = pos;
        __syncthreads();
    }
}

__global__ void clear_positions_kernel(const uint64_t *__restrict__ pos, uint64_t n, uint32_t *good21) {
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t p = __ldcs(pos + i) & ((1ULL << kPosRankShift) - 1);
        atomicAnd(good21 + (p >> 5), ~(0x80000000u >> (p & 31)));
    }
}

// Coverage verdicts without a list, a sort and an all-to-all: the owner looks at each of its
// first-occurrence candidates and, when the key's final count stayed below the threshold, clears the
// position's bit straight in the SOURCE rank's coverage plane (RED.AND over NVLink peer memory; the
// rank is the top byte of the position record).
struct PeerPlanes { uint32_t *p[kMaxPeers]; };
__global__ void __launch_bounds__(256)
cand_check_peer_kernel(const uint64_t *__restrict__ slots, const uint64_t *__restrict__ cand_slot,
                       const uint64_t *__restrict__ cand_pos, uint64_t n_cand, uint64_t thr, Ovf ovf,
                       const Stats *st, PeerPlanes planes) {
    const unsigned n_overflow = st->n_overflow;
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n_cand; j += stride) {
        uint64_t v = __ldcg(slots + __ldcs(cand_slot + j));
        uint64_t c = v >> 42;
        if (c < thr && n_overflow) c += ovf_get(ovf, v & kKey42) << 22;
        if (c < thr) {
            uint64_t rec = __ldcs(cand_pos + j);
            uint64_t pos = rec & ((1ULL << kPosRankShift) - 1);
            atomicAnd(planes.p[(rec >> kPosRankShift) & (kMaxPeers - 1)] + (pos >> 5), ~(0x80000000u >> (pos & 31)));
        }
    }
}

__global__ void set_insert_list_kernel(const uint64_t *__restrict__ kmers, uint64_t n, KSet set, Stats *st) {
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    bool full = false;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride)
        if (set_insert(set, __ldcs(kmers + i)) < 0) full = true;
    if (full) atomicExch(&st->err_table_full, 1u);
}

// host-side copy of owner_of (for tests and for callers that route on the host)
static inline uint32_t owner_of_host(uint64_t key, uint32_t n) {
    return (uint32_t)(((unsigned __int128)owner_mix(key) * n) >> 64);
}

struct MgState {   // extra per-context state of the multi-GPU path
    uint64_t *d_sing = nullptr, *d_sing2 = nullptr; uint64_t cap_sing = 0, cap_sing2 = 0;
    uint64_t *d_set2 = nullptr, *d_list2 = nullptr; uint64_t nbs2 = 0; uint32_t parts2 = 1;
    uint64_t rec_cap = 0;
    uint64_t n_local = 0;
    // receive buffers of the fused bin + exchange (peers store into them over NVLink); plain
    // cudaMalloc allocations so that cudaIpcGetMemHandle can export them to the other processes
    // (two sets: chunk c+1 is received while chunk c is being inserted)
    uint64_t *d_rkeys[2] = {nullptr, nullptr}; uint32_t *d_rwords[2] = {nullptr, nullptr};
    uint64_t cap_rkeys[2] = {0, 0}, cap_rwords[2] = {0, 0};
    // the fused bin + exchange kernel of the NEXT chunk runs on its own stream with its own cursors
    cudaStream_t stream2 = nullptr; unsigned long long *d_cursor2 = nullptr;
    bool swapped = false;   // c->d_set/d_list currently hold the OWNED set (p3_mg_owned_end swapped them in)
};
static CtxStates<MgState> g_mg;

static void mg_release(p3_ctx *c) {
    MgState *mp = g_mg.find(c);
    if (!mp) return;
    MgState &m = *mp;
    dfree(m.d_sing); dfree(m.d_sing2); dfree(m.d_set2); dfree(m.d_list2);
    for (int i = 0; i < 2; i++) { dfree(m.d_rkeys[i]); dfree(m.d_rwords[i]); }
    dfree(m.d_cursor2);
    if (m.stream2) cudaStreamDestroy(m.stream2);
    g_mg.erase(c);
}

static int mg_scan(p3_ctx *c, uint32_t P, uint64_t *h_counts) {
    scan_parts_kernel<<<1, 256, 0, c->stream>>>(c->d_ghist, P, c->d_cursor, c->d_ghist + kMaxParts);
    c->launches++;
    if (h_counts) {
        std::vector<unsigned long long> h(P);
        CU(cudaMemcpyAsync(h.data(), c->d_ghist, sizeof(unsigned long long) * P, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (uint32_t i = 0; i < P; i++) h_counts[i] = h[i];
    }
    return P3_OK;
}
static int mg_hist_buffers(p3_ctx *c) {
    if (!c->d_ghist) {
        CU(cudaMalloc(&c->d_ghist, sizeof(unsigned long long) * (kMaxParts + 1)));
        CU(cudaMalloc(&c->d_cursor, sizeof(unsigned long long) * (kMaxParts + 1)));
    }
    return scatter_attrs();
}

extern "C" {

uint32_t p3_owner_of_key(uint64_t key, uint32_t n_ranks) { return n_ranks ? owner_of_host(key, n_ranks) : 0; }

int p3_mg_owner_hist(p3_ctx *c, uint32_t n_ranks, uint64_t w0, uint64_t w1, uint64_t *h_counts) {
    if (!c || !c->have_reads) return fail(P3_ERR_STATE, "p3_mg_owner_hist: no reads attached");
    if (n_ranks == 0 || n_ranks > 256 || w1 > c->n_words || w0 > w1) return fail(P3_ERR_ARG, "p3_mg_owner_hist: bad arguments");
    CU(cudaSetDevice(c->device));
    int rc = mg_hist_buffers(c);
    if (rc) return rc;
    CU(cudaMemsetAsync(c->d_ghist, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    if (c->d_nmask) hist21_kernel<true, 1><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_rend, c->d_nmask, w0, w1, n_ranks, c->d_ghist);
    else hist21_kernel<false, 1><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_rend, nullptr, w0, w1, n_ranks, c->d_ghist);
    c->launches++;
    CU(cudaGetLastError());
    return mg_scan(c, n_ranks, h_counts);
}

// must follow p3_mg_owner_hist with the same arguments (it consumes the cursors that call set up)
int p3_mg_owner_scatter(p3_ctx *c, uint32_t n_ranks, uint32_t my_rank, uint64_t w0, uint64_t w1,
                        uint64_t *d_keys, uint32_t *d_words) {
    if (!c || !c->have_reads || !d_keys || !d_words) return fail(P3_ERR_STATE, "p3_mg_owner_scatter: no reads / null buffers");
    if (my_rank >= n_ranks || n_ranks > 256) return fail(P3_ERR_ARG, "p3_mg_owner_scatter: bad rank");
    CU(cudaSetDevice(c->device));
    CU(ensure(c->d_valid, c->cap_valid, sizeof(uint32_t) * (c->n_words + 1)));
    unsigned sblocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>((w1 - w0 + kTileWords - 1) / kTileWords, 1), (uint64_t)c->n_sm * 3);
    const uint64_t tag = (uint64_t)my_rank << kRecRankShift;
    if (c->d_nmask) scatter21_kernel<true, 1><<<sblocks, kScatterThreads, sizeof(ScatterSmem), c->stream>>>(c->d_packed, c->d_rend, c->d_nmask, w0, w1, n_ranks, c->d_cursor, d_keys, d_words, c->d_valid, tag);
    else scatter21_kernel<false, 1><<<sblocks, kScatterThreads, sizeof(ScatterSmem), c->stream>>>(c->d_packed, c->d_rend, nullptr, w0, w1, n_ranks, c->d_cursor, d_keys, d_words, c->d_valid, tag);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

// Fused bin + exchange: like p3_mg_owner_scatter, but owner j's records are stored straight to
// keys_base[j] / words_base[j] — device pointers into rank j's receive buffer that are mapped into
// this process (NVLink peer memory: p3_ipc_open of the owner's p3_mg_recv_buffers), already offset to
// the region reserved for this source rank (sizes from the all-gathered p3_mg_owner_hist counts). The
// caller synchronises all ranks before the owners read their buffers.
int p3_mg_owner_scatter_peer(p3_ctx *c, uint32_t n_ranks, uint32_t my_rank, uint64_t w0, uint64_t w1,
                             const uint64_t *keys_base, const uint64_t *words_base, int async) {
    if (!c || !c->have_reads || !keys_base || !words_base) return fail(P3_ERR_STATE, "p3_mg_owner_scatter_peer: no reads / null buffers");
    if (my_rank >= n_ranks || n_ranks > kMaxPeers) return fail(P3_ERR_ARG, "p3_mg_owner_scatter_peer: at most 16 ranks");
    CU(cudaSetDevice(c->device));
    int rc = mg_hist_buffers(c);
    if (rc) return rc;
    MgState &m = g_mg.get(c);
    if (!m.stream2) {
        CU(cudaStreamCreateWithFlags(&m.stream2, cudaStreamNonBlocking));
        CU(cudaMalloc(&m.d_cursor2, sizeof(unsigned long long) * (kMaxParts + 1)));
    }
    if (!c->d_valid || c->cap_valid < sizeof(uint32_t) * (c->n_words + 1)) {
        CU(cudaStreamSynchronize(m.stream2));   // nothing in flight may still write the old plane
        CU(ensure(c->d_valid, c->cap_valid, sizeof(uint32_t) * (c->n_words + 1)));
    }
    PeerOut po;
    for (uint32_t j = 0; j < (uint32_t)kMaxPeers; j++) {
        po.keys[j] = j < n_ranks ? (uint64_t *)(uintptr_t)keys_base[j] : nullptr;
        po.words[j] = j < n_ranks ? (uint32_t *)(uintptr_t)words_base[j] : nullptr;
    }
    // positions inside each destination region are relative: cursors start at zero. The kernel runs
    // on the second stream so that it can overlap the owner-side insert of the previous chunk.
    cudaStream_t st = m.stream2;
    CU(cudaMemsetAsync(m.d_cursor2, 0, sizeof(unsigned long long) * (kMaxParts + 1), st));
    unsigned sblocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>((w1 - w0 + kTileWords - 1) / kTileWords, 1), (uint64_t)c->n_sm * 3);
    const uint64_t tag = (uint64_t)my_rank << kRecRankShift;
    if (c->d_nmask) scatter21_kernel<true, 1, true><<<sblocks, kScatterThreads, sizeof(ScatterSmem), st>>>(c->d_packed, c->d_rend, c->d_nmask, w0, w1, n_ranks, m.d_cursor2, nullptr, nullptr, c->d_valid, tag, po);
    else scatter21_kernel<false, 1, true><<<sblocks, kScatterThreads, sizeof(ScatterSmem), st>>>(c->d_packed, c->d_rend, nullptr, w0, w1, n_ranks, m.d_cursor2, nullptr, nullptr, c->d_valid, tag, po);
    c->launches++;
    CU(cudaGetLastError());
    if (!async) CU(cudaStreamSynchronize(st));
    return P3_OK;
}
// waits for an async p3_mg_owner_scatter_peer of this context
int p3_mg_scatter_wait(p3_ctx *c) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    MgState *mp = g_mg.find(c);
    if (!mp || !mp->stream2) return P3_OK;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(mp->stream2));
    return P3_OK;
}

// receive buffers for n_records count records (grow-only; a grown buffer is a NEW allocation, whose
// handle has to be exported again)
int p3_mg_recv_buffers(p3_ctx *c, uint64_t n_records, uint32_t which, uint64_t **d_keys, uint32_t **d_words) {
    if (!c || which > 1) return fail(P3_ERR_ARG, "p3_mg_recv_buffers: null ctx / buffer index > 1");
    CU(cudaSetDevice(c->device));
    MgState &m = g_mg.get(c);
    n_records = std::max<uint64_t>(n_records, 1);
    CU(ensure(m.d_rkeys[which], m.cap_rkeys[which], sizeof(uint64_t) * n_records));
    CU(ensure(m.d_rwords[which], m.cap_rwords[which], sizeof(uint32_t) * n_records));
    if (d_keys) *d_keys = m.d_rkeys[which];
    if (d_words) *d_words = m.d_rwords[which];
    return P3_OK;
}

// CUDA IPC plumbing for the peer buffers (one process per GPU): export a cudaMalloc'ed buffer of this
// process as a 64-byte handle / map another process's buffer into this one (NVLink peer access is
// enabled on first use)
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

int p3_mg_count_begin(p3_ctx *c, uint64_t table_slots, uint64_t max_records_per_call) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    if (table_slots == 0) return fail(P3_ERR_ARG, "p3_mg_count_begin: table_slots required");
    CU(cudaSetDevice(c->device));
    int rc = mg_hist_buffers(c);
    if (rc) return rc;
    c->binned = true;
    rc = setup_table(c, table_slots);
    if (rc) return rc;
    MgState &m = g_mg.get(c);
    m.rec_cap = std::max<uint64_t>(max_records_per_call, 1);
    CU(ensure(c->d_bkeys, c->cap_bkeys, sizeof(uint64_t) * m.rec_cap));
    CU(ensure(c->d_bword, c->cap_bword, sizeof(uint32_t) * m.rec_cap));
    c->cand_cap = c->nb * 4;
    CU(ensure(c->d_cand_slot, c->cap_cand_slot, sizeof(uint64_t) * c->cand_cap));
    CU(ensure(c->d_cand_pos, c->cap_cand_pos, sizeof(uint64_t) * c->cand_cap));
    c->binned_pos = 0; c->n_chunks = 0;
    for (int i = 0; i < 4; i++) c->ms_sub[i] = 0;
    c->have_counts = false;
    return P3_OK;
}

int p3_mg_count_records(p3_ctx *c, const uint64_t *d_keys, const uint32_t *d_words, uint64_t n) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    if (n == 0) return P3_OK;
    CU(cudaSetDevice(c->device));
    const uint32_t P = c->parts;
    // fixed-capacity partition bins, as in the single-GPU count (p3_gpu.cu count_binned): no histogram
    // pass over the received records unless a partition overflows its share + 3 % + 8192
    uint64_t cap = getenv("P3_EXACT_BINS") ? 0 : (((uint64_t)((double)n / (double)P * 1.03) + 8192 + kSweepChunk - 1) / kSweepChunk * kSweepChunk);
    const uint64_t rec_cap = std::max<uint64_t>(n, cap * P);
    CU(ensure(c->d_bkeys, c->cap_bkeys, sizeof(uint64_t) * rec_cap));   // local bins grow with the largest batch
    CU(ensure(c->d_bword, c->cap_bword, sizeof(uint32_t) * rec_cap));
    unsigned sblocks = (unsigned)std::min<uint64_t>((n + kTilePos - 1) / kTilePos, (uint64_t)c->n_sm * 3);
    CU(cudaEventRecord(c->ev[10], c->stream));
    CU(cudaEventRecord(c->ev[11], c->stream));
    if (cap) {
        init_cursors_kernel<<<1, 256, 0, c->stream>>>(c->d_cursor, P, cap);
        scatter_rec_kernel<0, true><<<sblocks, kScatterThreads, sizeof(ScatterSmem), c->stream>>>(d_keys, d_words, n, P, c->d_cursor, c->d_bkeys, c->d_bword, cap);
        c->launches += 2;
        CU(cudaGetLastError());
        std::vector<unsigned long long> h_cur(P);
        CU(cudaMemcpyAsync(h_cur.data(), c->d_cursor, sizeof(unsigned long long) * P, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (uint32_t q = 0; q < P; q++)
            if (h_cur[q] - (unsigned long long)q * cap > cap) { cap = 0; break; }
    }
    if (!cap) {
        CU(cudaMemsetAsync(c->d_ghist, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
        CU(cudaEventRecord(c->ev[10], c->stream));
        hist_rec_kernel<0><<<c->grid(), 256, 0, c->stream>>>(d_keys, n, P, c->d_ghist);
        int rc = mg_scan(c, P, nullptr);
        if (rc) return rc;
        CU(cudaEventRecord(c->ev[11], c->stream));
        scatter_rec_kernel<0, true><<<sblocks, kScatterThreads, sizeof(ScatterSmem), c->stream>>>(d_keys, d_words, n, P, c->d_cursor, c->d_bkeys, c->d_bword);
        c->launches += 3;
    }
    CU(cudaEventRecord(c->ev[12], c->stream));
    CU(cudaMemsetAsync(&c->d_stats->work, 0, sizeof(unsigned long long), c->stream));
    if (cap) insert_bins_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_bkeys, c->d_bword, (uint64_t)P * cap, c->table(), c->ovf(), c->d_stats, c->d_cand_slot, c->d_cand_pos, c->cand_cap, cap, c->d_cursor);
    else insert_bins_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_bkeys, c->d_bword, n, c->table(), c->ovf(), c->d_stats, c->d_cand_slot, c->d_cand_pos, c->cand_cap);
    CU(cudaEventRecord(c->ev[13], c->stream));
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    {   // owner-side sub-stage times: partition histogram, tile sort by partition, L2-resident insert sweep
        float a = 0, b = 0, d = 0;
        cudaEventElapsedTime(&a, c->ev[10], c->ev[11]);
        cudaEventElapsedTime(&b, c->ev[11], c->ev[12]);
        cudaEventElapsedTime(&d, c->ev[12], c->ev[13]);
        c->ms_sub[0] += a; c->ms_sub[1] += b; c->ms_sub[2] += d;
    }
    c->binned_pos += n; c->n_chunks++;
    return P3_OK;
}

int p3_mg_count_end(p3_ctx *c) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    int rc = pull_stats(c);
    if (rc) return rc;
    if (c->h_stats.err_table_full) return fail(P3_ERR_TABLE_FULL, "21-mer count table full: raise table_slots");
    if (c->h_stats.err_ovf_full) return fail(P3_ERR_TABLE_FULL, "count overflow side table full");
    if (c->h_stats.n_cand > c->cand_cap) return fail(P3_ERR_TABLE_FULL, "candidate list overflow");
    c->have_counts = true;
    c->have_bf = c->have_solid = c->have_adj = false;
    return P3_OK;
}

// positions (with their source rank in the top byte) of every key this rank owns whose final
// count is below the reference's cov_threshold of 2, grouped by source rank
int p3_mg_singletons(p3_ctx *c, uint32_t n_ranks, uint64_t *h_counts, const uint64_t **d_pos) {
    if (!c || !c->have_counts) return fail(P3_ERR_STATE, "p3_mg_singletons: no counts");
    CU(cudaSetDevice(c->device));
    MgState &m = g_mg.get(c);
    uint64_t nc = c->h_stats.n_cand;
    CU(ensure(m.d_sing, m.cap_sing, sizeof(uint64_t) * std::max<uint64_t>(nc, 1)));
    CU(ensure(m.d_sing2, m.cap_sing2, sizeof(uint64_t) * std::max<uint64_t>(nc, 1)));
    CU(cudaMemsetAsync(&c->d_stats->n_export, 0, sizeof(unsigned long long), c->stream));
    if (nc) {
        singleton_list_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_table, c->d_cand_slot, c->d_cand_pos, nc, P3_COV_THRESHOLD, c->ovf(), c->d_stats, m.d_sing);
        c->launches++;
    }
    int rc = pull_stats(c);
    if (rc) return rc;
    uint64_t ns = c->h_stats.n_export;
    CU(cudaMemsetAsync(c->d_ghist, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    if (ns) hist_rec_kernel<2><<<c->grid(), 256, 0, c->stream>>>(m.d_sing, ns, n_ranks, c->d_ghist);
    rc = mg_scan(c, n_ranks, h_counts);
    if (rc) return rc;
    if (ns) {
        unsigned sblocks = (unsigned)std::min<uint64_t>((ns + kTilePos - 1) / kTilePos, (uint64_t)c->n_sm * 3);
        scatter_rec_kernel<2, false><<<sblocks, kScatterThreads, sizeof(ScatterSmem), c->stream>>>(m.d_sing, nullptr, ns, n_ranks, c->d_cursor, m.d_sing2, nullptr);
        c->launches += 2;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    if (d_pos) *d_pos = m.d_sing2;
    return P3_OK;
}

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

// coverage plane := every valid 21-mer position (p3_mg_owner_scatter filled the valid plane)
int p3_mg_cover_begin(p3_ctx *c) {
    if (!c || !c->have_reads || !c->d_valid) return fail(P3_ERR_STATE, "p3_mg_cover_begin: run p3_mg_owner_scatter first");
    CU(cudaSetDevice(c->device));
    int rc = ensure_planes(c);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_good21, c->d_valid, sizeof(uint32_t) * c->n_words, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_good21 + c->n_words, 0, sizeof(uint32_t), c->stream));
    return P3_OK;
}
int p3_mg_cover_clear(p3_ctx *c, const uint64_t *d_pos, uint64_t n) {
    if (!c || !c->d_good21) return fail(P3_ERR_STATE, "p3_mg_cover_clear: run p3_mg_cover_begin first");
    if (n == 0) return P3_OK;
    CU(cudaSetDevice(c->device));
    bool cleared = false;
    int rcc = binned_plane_clear<1>(c, nullptr, d_pos, n, 0, c->d_good21, c->n_words * 32, &cleared);
    if (rcc) return rcc;
    if (cleared) { CU(cudaStreamSynchronize(c->stream)); return P3_OK; }
    clear_positions_kernel<<<c->grid(), 256, 0, c->stream>>>(d_pos, n, c->d_good21);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

int p3_mg_cover_plane(p3_ctx *c, uint32_t **d_plane) {
    if (!c || !c->d_good21) return fail(P3_ERR_STATE, "p3_mg_cover_plane: run p3_mg_cover_begin first");
    if (d_plane) *d_plane = c->d_good21;
    return P3_OK;
}
// owner side, fused with the exchange: clear the bits of this rank's count-1 keys directly in the
// source ranks' planes (planes[j] = rank j's p3_mg_cover_plane, mapped into this process). All ranks
// must have run p3_mg_cover_begin before any rank calls this, and must synchronise afterwards.
int p3_mg_cover_peer(p3_ctx *c, uint32_t n_ranks, const uint64_t *planes) {
    if (!c || !c->have_counts || !planes) return fail(P3_ERR_STATE, "p3_mg_cover_peer: no counts / null planes");
    if (n_ranks == 0 || n_ranks > kMaxPeers) return fail(P3_ERR_ARG, "p3_mg_cover_peer: at most 16 ranks");
    CU(cudaSetDevice(c->device));
    PeerPlanes pp;
    for (uint32_t j = 0; j < (uint32_t)kMaxPeers; j++) pp.p[j] = (uint32_t *)(uintptr_t)planes[j < n_ranks ? j : 0];
    uint64_t nc = c->h_stats.n_cand;
    if (nc) {
        cand_check_peer_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_table, c->d_cand_slot, c->d_cand_pos, nc, P3_COV_THRESHOLD, c->ovf(), c->d_stats, pp);
        c->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

// solid plane (window AND of the coverage plane), seeds, and the locally distinct solid k-mers
int p3_mg_solid_local(p3_ctx *c, uint32_t k, uint64_t solid_slots, uint64_t *n_adds, uint64_t *n_local) {
    if (!c || !c->d_good21) return fail(P3_ERR_STATE, "p3_mg_solid_local: run p3_mg_cover_begin first");
    if (k < P3_MIN_K || k > 32) return fail(P3_ERR_ARG, "multi-GPU path: k outside [21,32] is not supported yet");
    CU(cudaSetDevice(c->device));
    c->k = k; c->set_valid = false; c->d_set_b = nullptr; c->nbs_b = 0; c->parts_b = 1;
    {   // give the local-set buffers of the previous run back to their role so that they are reused
        MgState &m = g_mg.get(c);
        if (m.swapped) {
            std::swap(c->d_set, m.d_set2); std::swap(c->d_list, m.d_list2); std::swap(c->nbs, m.nbs2); std::swap(c->set_parts, m.parts2);
            c->list_cap = c->nbs * 4;
            m.swapped = false;
        }
    }
    CU(cudaMemsetAsync(&c->d_stats->n_adds, 0, sizeof(unsigned long long) * 5, c->stream));
    solid_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_good21, c->n_words, (int)k, c->d_solid, c->d_stats);
    c->launches++;
    int rc = pull_stats(c);
    if (rc) return rc;
    if (solid_slots == 0) solid_slots = std::max<uint64_t>(2 * std::min<uint64_t>(c->h_stats.n_adds, 1ull << 26), 1024);
    rc = dedupe_solid_positions(c, k, solid_slots);
    if (rc) return rc;
    seeds_kernel<<<c->grid(4), 256, 0, c->stream>>>(c->d_off, c->n_reads, c->d_solid, (int)k, c->d_seed);
    c->launches++;
    CU(cudaStreamSynchronize(c->stream));
    g_mg.get(c).n_local = c->h_stats.n_distinct_solid;
    c->have_solid = true;
    if (n_adds) *n_adds = c->h_stats.n_adds;
    if (n_local) *n_local = c->h_stats.n_distinct_solid;
    return P3_OK;
}

int p3_mg_kmer_owner_hist(p3_ctx *c, uint32_t n_ranks, uint64_t *h_counts) {
    if (!c || !c->have_solid) return fail(P3_ERR_STATE, "p3_mg_kmer_owner_hist: run p3_mg_solid_local first");
    CU(cudaSetDevice(c->device));
    uint64_t n = g_mg.get(c).n_local;
    CU(cudaMemsetAsync(c->d_ghist, 0, sizeof(unsigned long long) * (kMaxParts + 1), c->stream));
    if (n) { hist_rec_kernel<3><<<c->grid(), 256, 0, c->stream>>>(c->d_list, n, n_ranks, c->d_ghist); c->launches++; }
    CU(cudaGetLastError());
    return mg_scan(c, n_ranks, h_counts);
}
int p3_mg_kmer_owner_scatter(p3_ctx *c, uint32_t n_ranks, uint64_t *d_out) {
    if (!c || !c->have_solid || !d_out) return fail(P3_ERR_STATE, "p3_mg_kmer_owner_scatter: bad state");
    CU(cudaSetDevice(c->device));
    uint64_t n = g_mg.get(c).n_local;
    if (n) {
        unsigned sblocks = (unsigned)std::min<uint64_t>((n + kTilePos - 1) / kTilePos, (uint64_t)c->n_sm * 3);
        scatter_rec_kernel<3, false><<<sblocks, kScatterThreads, sizeof(ScatterSmem), c->stream>>>(c->d_list, nullptr, n, n_ranks, c->d_cursor, d_out, nullptr);
        c->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

int p3_mg_owned_begin(p3_ctx *c, uint64_t owned_slots) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    MgState &m = g_mg.get(c);
    if (m.swapped) return fail(P3_ERR_STATE, "p3_mg_owned_begin: run p3_mg_solid_local first");
    uint64_t nbs = (std::max<uint64_t>(owned_slots, 1024) + 3) / 4;
    if (!m.d_set2 || m.nbs2 != nbs) {
        dfree(m.d_set2); dfree(m.d_list2);
        if (cudaMalloc(&m.d_set2, nbs * 32) != cudaSuccess || cudaMalloc(&m.d_list2, nbs * 32) != cudaSuccess) {
            cudaGetLastError();
            return fail(P3_ERR_NOMEM, "owned k-mer set allocation failed");
        }
        m.nbs2 = nbs;
    }
    m.parts2 = 1;
    CU(cudaMemsetAsync(m.d_set2, 0xFF, nbs * 32, c->stream));
    CU(cudaMemsetAsync(&c->d_stats->err_table_full, 0, sizeof(unsigned), c->stream));
    return P3_OK;
}
int p3_mg_owned_insert(p3_ctx *c, const uint64_t *d_kmers, uint64_t n) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    if (n == 0) return P3_OK;
    CU(cudaSetDevice(c->device));
    MgState &m = g_mg.get(c);
    KSet owned; owned.slots = m.d_set2; owned.P = 1; owned.nbp = m.nbs2;
    set_insert_list_kernel<<<c->grid(), 256, 0, c->stream>>>(d_kmers, n, owned, c->d_stats);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

// the owned set becomes THE set/list of this context; the filter copy (at least min_words words) is
// cleared and, with do_adds, receives BF.add of every owned k-mer
static int owned_finish(p3_ctx *c, uint32_t k, uint64_t filter_size, uint32_t num_hashes, uint64_t min_words,
                        bool do_adds, uint64_t *n_owned) {
    if (!c) return fail(P3_ERR_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    MgState &m = g_mg.get(c);
    if (!m.d_set2) return fail(P3_ERR_STATE, "p3_mg_owned_end: run p3_mg_owned_begin first");
    int rc = alloc_bloom(c, k, filter_size, num_hashes, min_words);
    if (rc) return rc;
    std::swap(c->d_set, m.d_set2); std::swap(c->d_list, m.d_list2); std::swap(c->nbs, m.nbs2); std::swap(c->set_parts, m.parts2);
    c->list_cap = c->nbs * 4;
    m.swapped = true;
    CU(cudaMemsetAsync(&c->d_stats->n_distinct_solid, 0, sizeof(unsigned long long), c->stream));
    compact_set_kernel<<<c->grid(), 256, 0, c->stream>>>(c->d_set, c->nbs * 4, c->d_list, c->list_cap, c->d_stats);
    c->launches++;
    rc = pull_stats(c);
    if (rc) return rc;
    if (c->h_stats.err_table_full) return fail(P3_ERR_TABLE_FULL, "owned k-mer set full: raise owned_slots");
    CU(cudaMemsetAsync(c->d_bloom, 0, sizeof(uint32_t) * c->bloom_words, c->stream));
    if (do_adds) {
        rc = bloom_add_list(c, c->h_stats.n_distinct_solid);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(c->stream));
    c->have_bf = true; c->have_solid = true; c->have_adj = false; c->set_valid = true;
    c->d_set_b = m.d_set2; c->nbs_b = m.nbs2; c->parts_b = m.parts2;   // after the swap: the locally seen solid k-mers
    if (n_owned) *n_owned = c->h_stats.n_distinct_solid;
    return P3_OK;
}
// replicated-filter variant: adds into this rank's full copy (the caller then OR-reduces the copies)
int p3_mg_owned_end(p3_ctx *c, uint32_t k, uint64_t filter_size, uint32_t num_hashes, uint64_t *n_owned) {
    return owned_finish(c, k, filter_size, num_hashes, 0, true, n_owned);
}
// sharded-filter variant: list only; the adds follow as p3_mg_bloom_bin / _apply (or _direct)
int p3_mg_owned_list(p3_ctx *c, uint32_t k, uint64_t filter_size, uint32_t num_hashes, uint64_t filter_words_cap, uint64_t *n_owned) {
    return owned_finish(c, k, filter_size, num_hashes, filter_words_cap, false, n_owned);
}

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
    if (!c || !c->have_bf || !h_segbase || !h_counts) return fail(P3_ERR_STATE, "p3_mg_bloom_bin: run p3_mg_owned_list first");
    if (n_seg == 0 || n_seg > (uint32_t)kMaxParts || c->num_hashes > (uint32_t)kBinMaxHashes)
        return fail(P3_ERR_ARG, "p3_mg_bloom_bin: too many segments / hash functions for the binned path");
    CU(cudaSetDevice(c->device));
    uint64_t n = c->h_stats.n_distinct_solid;
    uint64_t *d_hh = nullptr;
    int rc = bloom_hash_list(c, n, &d_hh);
    if (rc) return rc;
    return bloom_bin_launch(c, d_hh, n, n_seg, bloom_seg_shift(), h_segbase, cap, h_counts);
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
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}
// fallback of the sharded adds: every owned k-mer straight into this rank's full copy (then OR-reduce)
int p3_mg_bloom_direct(p3_ctx *c) {
    if (!c || !c->have_bf) return fail(P3_ERR_STATE, "p3_mg_bloom_direct: no filter");
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->d_bloom, 0, sizeof(uint32_t) * c->bloom_words, c->stream));
    int rc = bloom_add_list(c, c->h_stats.n_distinct_solid);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return P3_OK;
}

int p3_mg_filter(p3_ctx *c, uint32_t **d_bits, uint64_t *n_words) {
    if (!c || !c->d_bloom) return fail(P3_ERR_STATE, "p3_mg_filter: no filter");
    if (d_bits) *d_bits = c->d_bloom;
    if (n_words) *n_words = c->bloom_words;
    return P3_OK;
}

}  // extern "C"
