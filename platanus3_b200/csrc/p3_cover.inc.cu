// p3_cover.inc.cu — DeBruijnGraph::CountNodeCoverage (reference src/DeBruijnGraph.cpp:394-449) on
// the GPU, part of the p3_gpu.cu translation unit. The reference walks every k-mer of every read
// (forward and backward orientation) under `omp critical` and bumps counters of the junction /
// joint nodes it hits. Here: the node k-mers go into two small open-addressing tables (a few MB,
// L2 resident), one thread per packed word streams the reads, counters are int32 atomics.
//   junction counters: [coverage, left_kmers_cov[4], right_kmers_cov[4]]  (9 per node)
//   joint counters   : [coverage]

// At k = 32 the all-T k-mer equals the table's empty marker (kEmpty = ~0), and the keys here are
// ORIENTED k-mers, so it is a legal key (a read with 32 T's forward, 32 A's backward). It never enters
// the table: its node index travels beside it in `ones_idx` (-1 = no such node).
struct NodeTable { const uint64_t *keys; const uint32_t *idx; uint64_t mask; int ones_idx; };

__global__ void node_table_build_kernel(const uint64_t *__restrict__ kmers, uint64_t n, uint64_t *keys, uint32_t *idx, uint64_t mask) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = kmers[i];
    if (key == kEmpty) return;   // kept outside the table (NodeTable::ones_idx)
    uint64_t s = fmix64(key) & mask;
    for (;;) {
        uint64_t old = atomicCAS(ull(keys + s), kEmpty, key);
        if (old == kEmpty || old == key) { idx[s] = (uint32_t)i; return; }
        s = (s + 1) & mask;
    }
}
__device__ __forceinline__ int node_find(const NodeTable &t, uint64_t key) {
    if (key == kEmpty) return t.ones_idx;
    if (!t.keys) return -1;
    uint64_t s = fmix64(key) & t.mask;
    for (;;) {
        uint64_t v = __ldg(t.keys + s);
        if (v == key) return (int)__ldg(t.idx + s);
        if (v == kEmpty) return -1;
        s = (s + 1) & t.mask;
    }
}

template <bool HAS_MASK>
__global__ void __launch_bounds__(256)
node_coverage_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ rend,
                     const uint32_t *__restrict__ nmask, uint64_t n_words, int k, NodeTable J, NodeTable T,
                     int *jcov, int *tcov) {
    const uint64_t Wk = ~0ULL << (64 - (k - 1));   // k >= 21
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += stride) {
        const uint64_t hi = __ldg(packed + w), lo = __ldg(packed + w + 1);
        const uint64_t E = ((uint64_t)__ldg(rend + w) << 32) | __ldg(rend + w + 1);
        const uint64_t pw = w ? __ldg(packed + w - 1) : 0;
        const bool prev_end = w ? (__ldg(rend + w - 1) & 1u) : true;     // position 32w-1 ends a read
        uint64_t mhi = 0, mlo = 0, mfl = 0;
        bool pmask = false;
        if (HAS_MASK) {
            uint32_t a = __ldg(nmask + w), b = __ldg(nmask + w + 1);
            mhi = spread32(a); mlo = spread32(b); mfl = ((uint64_t)a << 32) | b;
            pmask = w ? (__ldg(nmask + w - 1) & 1u) : false;
        }
        for (int o = 0; o < 32; o++) {
            if (((E << o) & Wk) != 0) continue;            // k-mer crosses a read end
            const uint64_t x = window(hi, lo, o);
            const uint64_t m2 = HAS_MASK ? window(mhi, mlo, o) : 0;
            const uint64_t fw = x >> (64 - 2 * k);
            const uint64_t bw = rev2(~x & ~m2) & kmask(k);  // rolling kmer_Bw incl. the non-ACGT quirk
            // AddNodeCoverage(kmer_Fw); AddNodeCoverage(kmer_Bw)   (:442-449)
            const int jf = node_find(J, fw), jb = node_find(J, bw);
            const int tf = node_find(T, fw), tb = node_find(T, bw);
            if (jf >= 0) atomicAdd(jcov + 9 * jf, 1);
            if (tf >= 0) atomicAdd(tcov + tf, 1);
            if (jb >= 0) atomicAdd(jcov + 9 * jb, 1);
            if (tb >= 0) atomicAdd(tcov + tb, 1);
            if (jf < 0 && jb < 0) continue;
            // path coverage (:407-413 for the first k-mer of a read, :424-435 for the others)
            const bool first = o ? ((E >> (64 - o)) & 1) : prev_end;
            const int jn = o + k;                           // base after the k-mer, inside hi:lo (<= 63)
            const bool has_next = !((E >> (63 - (jn - 1))) & 1);
            int nf = 0, nr = 0;                              // fcode / rcode of that base ('\0' -> 0 when absent)
            if (has_next) {
                nf = (int)((jn < 32 ? (hi >> (62 - 2 * jn)) : (lo >> (62 - 2 * (jn - 32)))) & 3);
                bool nm = HAS_MASK && ((mfl >> (63 - jn)) & 1);
                nr = nm ? 0 : 3 - nf;
            }
            int pf = 0, pr = 0;
            if (!first) {
                pf = (int)((o ? (hi >> (64 - 2 * o)) : pw) & 3);
                bool pm = HAS_MASK && (o ? ((mfl >> (64 - o)) & 1) : pmask);
                pr = pm ? 0 : 3 - pf;
            }
            if (first) {
                if (jf >= 0) atomicAdd(jcov + 9 * jf + 5 + nf, 1);          // right_kmers_cov[fcode(read[k])]
                else atomicAdd(jcov + 9 * jb + 1 + nr, 1);                  // left_kmers_cov[rcode(read[k])]
            } else if (jf >= 0) {
                atomicAdd(jcov + 9 * jf + 1 + pf, 1);                       // left[fcode(read[i-k])]
                if (has_next) atomicAdd(jcov + 9 * jf + 5 + nf, 1);         // right[fcode(read[i+1])]
            } else {
                atomicAdd(jcov + 9 * jb + 5 + pr, 1);                       // right[rcode(read[i-k])]
                if (has_next) atomicAdd(jcov + 9 * jb + 1 + nr, 1);         // left[rcode(read[i+1])]
            }
        }
    }
}

static int build_node_table(p3_ctx *c, const uint64_t *h_kmers, uint64_t n, uint64_t **d_keys, uint32_t **d_idx, uint64_t *mask, int *ones_idx) {
    *d_keys = nullptr; *d_idx = nullptr; *mask = 0; *ones_idx = -1;
    if (n == 0) return P3_OK;
    for (uint64_t i = 0; i < n; i++) if (h_kmers[i] == kEmpty) *ones_idx = (int)i;   // last one wins, like the table
    uint64_t cap = 16;
    while (cap < 2 * n + 16) cap <<= 1;
    uint64_t *dk = nullptr;
    TmpFree tmp; tmp.own(&dk);
    CU(cudaMalloc(d_keys, sizeof(uint64_t) * cap));
    CU(cudaMalloc(d_idx, sizeof(uint32_t) * cap));
    CU(cudaMalloc(&dk, sizeof(uint64_t) * n));
    CU(cudaMemsetAsync(*d_keys, 0xFF, sizeof(uint64_t) * cap, c->stream));
    CU(cudaMemcpyAsync(dk, h_kmers, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, c->stream));
    node_table_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(dk, n, *d_keys, *d_idx, cap - 1);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    *mask = cap - 1;
    return P3_OK;
}

extern "C" int p3_node_coverage(p3_ctx *c, uint32_t k, const uint64_t *h_junctions, uint64_t nj,
                                const uint64_t *h_joints, uint64_t nt, int32_t *h_jcov, int32_t *h_tcov) {
    if (!c || !c->have_reads) return fail(P3_ERR_STATE, "p3_node_coverage: no reads attached");
    if (k < P3_MIN_K || k > P3_MAX_K_WALK) return fail(P3_ERR_ARG, "p3_node_coverage: k outside [21,32]");
    CU(cudaSetDevice(c->device));
    uint64_t *jk = nullptr, *tk = nullptr; uint32_t *ji = nullptr, *ti = nullptr; uint64_t jm = 0, tm = 0; int jo = -1, to = -1;
    int *djc = nullptr, *dtc = nullptr;
    TmpFree tmp; tmp.own(&jk); tmp.own(&ji); tmp.own(&tk); tmp.own(&ti); tmp.own(&djc); tmp.own(&dtc);
    int rc = build_node_table(c, h_junctions, nj, &jk, &ji, &jm, &jo);
    if (!rc) rc = build_node_table(c, h_joints, nt, &tk, &ti, &tm, &to);
    if (!rc) {
        CU(cudaMalloc(&djc, sizeof(int) * std::max<uint64_t>(9 * nj, 1)));
        CU(cudaMalloc(&dtc, sizeof(int) * std::max<uint64_t>(nt, 1)));
        CU(cudaMemsetAsync(djc, 0, sizeof(int) * std::max<uint64_t>(9 * nj, 1), c->stream));
        CU(cudaMemsetAsync(dtc, 0, sizeof(int) * std::max<uint64_t>(nt, 1), c->stream));
        NodeTable J{jk, ji, jm, jo}, T{tk, ti, tm, to};
        if (nj + nt) {
            if (c->d_nmask) node_coverage_kernel<true><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_rend, c->d_nmask, c->n_words, (int)k, J, T, djc, dtc);
            else node_coverage_kernel<false><<<c->grid(), 256, 0, c->stream>>>(c->d_packed, c->d_rend, nullptr, c->n_words, (int)k, J, T, djc, dtc);
            c->launches++;
        }
        CU(cudaGetLastError());
        if (nj) CU(cudaMemcpyAsync(h_jcov, djc, sizeof(int) * 9 * nj, cudaMemcpyDeviceToHost, c->stream));
        if (nt) CU(cudaMemcpyAsync(h_tcov, dtc, sizeof(int) * nt, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return rc;
}
