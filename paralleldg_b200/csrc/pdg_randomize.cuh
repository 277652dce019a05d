// Junction-tree randomisation, one warp per chain.  Restates JunctionMap.randomize_by_jt
// (junction_tree.py:235-241): graph from cliques (:201-219) -> maximal cliques + a junction tree
// (decomposable.py:124-150, here by maximum-cardinality search) -> uniform random equivalent
// junction tree (junction_tree.py:698-774; every distinct separator handled from the original
// tree, Thomas & Green 2009) -> random relabelling + empty-node padding (:19-65).  All draws come
// from addressable Philox streams, so lanes generate them in parallel.
#pragma once
#include "pdg_kernels.cuh"

struct RandSmem {
    int off_card, off_order, off_cliqueof, off_parent, off_keys, off_jdraw, off_comp, off_nodes, off_rank,
        off_cstart, off_canon, off_compidx, off_vind, off_wj, off_cnt, off_fill, off_off, off_newe,
        off_done, off_inT, off_inw, off_isroot, off_numbered, off_used, off_X, off_wmask, off_sepsz, off_deg32,
        off_sets, total;      // off_sets < 0: adj / K / sep stay in the global scratch
};

__host__ __device__ inline RandSmem rand_smem_layout(int p, int W)
{
    RandSmem s;
    int o = 0;
    const int a2 = (((p + 2) * 2 + 15) / 16) * 16, a1 = ((p + 2 + 15) / 16) * 16;
    s.off_card = o; o += a2;   s.off_order = o; o += a2;  s.off_cliqueof = o; o += a2;
    s.off_parent = o; o += a2; s.off_keys = o; o += a2;   s.off_jdraw = o; o += a2;
    s.off_comp = o; o += a2;   s.off_nodes = o; o += a2;  s.off_rank = o; o += a2;
    s.off_cstart = o; o += a2; s.off_canon = o; o += a2;  s.off_compidx = o; o += a2;
    s.off_vind = o; o += a2;   s.off_wj = o; o += a2;     s.off_cnt = o; o += a2;
    s.off_fill = o; o += a2;   s.off_off = o; o += a2;    s.off_newe = o; o += 2 * a2;
    s.off_done = o; o += a1;   s.off_inT = o; o += a1;    s.off_inw = o; o += a1;  s.off_isroot = o; o += a1;
    s.off_numbered = o; o += W * 8; s.off_used = o; o += W * 8; s.off_X = o; o += W * 8;
    s.off_wmask = o; o += a1; s.off_sepsz = o; o += a2;
    s.off_deg32 = o; o += ((p + 2) * 4 + 15) / 16 * 16;
    // up to 128 vertices the three p x W work tables (graph adjacency, cliques, separators) live in
    // shared memory too (<= 6 KB): the MCS walks them with one dependent access per step
    if (p * W <= 256) { s.off_sets = o; o += 3 * p * W * 8; } else s.off_sets = -1;
    s.total = o;
    return s;
}

struct RandScratch {
    u64 *adj, *K, *sep;   // [chain][p][W] each
};

__device__ __forceinline__ bool subset_w(const u64 *a, const u64 *b, int W)
{
    for (int w = 0; w < W; ++w) if (a[w] & ~b[w]) return false;
    return true;
}
// the same over the word range [lo, hi] that holds all bits of a
__device__ __forceinline__ bool subset_r(const u64 *a, const u64 *b, int lo, int hi)
{
    for (int w = lo; w <= hi; ++w) if (a[w] & ~b[w]) return false;
    return true;
}
__device__ __forceinline__ bool equal_w(const u64 *a, const u64 *b, int W)
{
    for (int w = 0; w < W; ++w) if (a[w] != b[w]) return false;
    return true;
}

// derived state of one chain: N = transpose of M, CSR of the tree.  Called by one warp; s_deg is
// (p + 1) ints of shared memory.  O(p) per chain: degrees and fill cursors through shared atomics
// over the edge list (the order inside an adjacency list carries no meaning).
template <typename PT>
__device__ __forceinline__ void build_derived(const PT &P, long long chain, int lane, int *s_deg)
{
    const int p = P.p, W = P.W;
    u64 *Mc = P.M + (size_t)chain * p * W;
    u64 *Nc = P.N + (size_t)chain * p * W;
    const int *E = P.edges + (size_t)chain * 2 * (p - 1);
    unsigned short *g_off = P.csr_off + (size_t)chain * (p + 1);
    unsigned short *g_adj = P.csr_adj + (size_t)chain * 2 * p;
    for (int i = lane; i < p * W; i += 32) Nc[i] = 0ull;
    for (int i = lane; i <= p; i += 32) s_deg[i] = 0;
    __threadfence();
    __syncwarp();
    for (int i0 = lane; i0 < p * W; i0 += 128) {
        u64 xs[4];                                   // four independent loads in flight per lane
#pragma unroll
        for (int q = 0; q < 4; ++q) xs[q] = (i0 + 32 * q < p * W) ? Mc[i0 + 32 * q] : 0ull;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + 32 * q, t = i / W, w = i - t * W;
            u64 x = xs[q];
            while (x) {
                const int b = __ffsll((long long)x) - 1;
                x &= x - 1;
                const int g = w * 64 + b;
                atomicOr(&Nc[(size_t)g * W + (t >> 6)], 1ull << (t & 63));
            }
        }
    }
    for (int e = lane; e < p - 1; e += 32) { atomicAdd(&s_deg[E[2 * e]], 1); atomicAdd(&s_deg[E[2 * e + 1]], 1); }
    __syncwarp();
    // exclusive prefix over p degrees: 32 per pass with a warp scan
    int carry = 0;
    for (int base = 0; base <= p; base += 32) {
        const int t = base + lane;
        const int d = (t < p) ? s_deg[t] : 0;
        int incl = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(PDG_FULL, incl, o); if (lane >= o) incl += v; }
        const int excl = carry + incl - d;
        if (t <= p) { g_off[t] = (unsigned short)excl; if (t < p) s_deg[t] = excl; }
        carry += __shfl_sync(PDG_FULL, incl, 31);
    }
    __syncwarp();
    for (int e = lane; e < p - 1; e += 32) {
        const int a = E[2 * e], b = E[2 * e + 1];
        g_adj[atomicAdd(&s_deg[a], 1)] = (unsigned short)b;
        g_adj[atomicAdd(&s_deg[b], 1)] = (unsigned short)a;
    }
    __threadfence();              // N (atomic ORs) and the CSR are read back by other lanes right after
    __syncwarp();
}

#ifdef PDG_DEFINE_GLOBAL_KERNELS
__global__ void __launch_bounds__(32) k_build_derived(PdgParams P, long long chain_begin, long long chain_count)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const long long c = blockIdx.x;
    if (c >= chain_count) return;
    build_derived(P, chain_begin + c, threadIdx.x, (int *)smem);
}
#endif

// The fields of PdgParams the randomiser needs, passed BY VALUE: randomize_chain is out of line (it
// is cold inside the MH kernel and would bloat its loop body) and taking the address of the kernel
// parameter block would force the whole block into local memory.
struct RandArgs {
    int p, W;
    u64 *M, *N;
    int *edges;
    unsigned short *csr_off, *csr_adj;
    int *err;
    u32 chain0, seed_lo, seed_hi;
};
__device__ __forceinline__ RandArgs rand_args(const PdgParams &P)
{
    RandArgs a;
    a.p = P.p; a.W = P.W; a.M = P.M; a.N = P.N; a.edges = P.edges; a.csr_off = P.csr_off; a.csr_adj = P.csr_adj;
    a.err = P.err; a.chain0 = P.chain0; a.seed_lo = P.seed_lo; a.seed_hi = P.seed_hi;
    return a;
}

// Maximum cardinality search over the graph adjacency `adj` -> maximal cliques K[0..S), their clique-tree
// parents and separators (replaces nx.find_cliques + the maximum-weight spanning tree of
// decomposable.py:124-150: any junction tree is a valid start for the randomisation).  Returns S.
// key of vertex u = (cardinality << 16) | (0xFFFF - u), 0 once numbered; vertex u = lane + 32 j lives in
// register j of its lane, so a step is 2 NW max operations, one warp reduction and 2 NW bit tests on the
// broadcast-loaded adjacency row of the new vertex.  NW = bitset words handled (2: p <= 128, rows in
// shared memory; 8: rows in global memory, built with L2 atomics, hence read with ld.cg).
template <int NW, bool ADJ_L2>
__device__ __forceinline__ int mcs_cliques(const u64 *adj, u64 *K, u64 *sep, u64 *numbered, unsigned short *order,
                                           unsigned short *cliqueof, short *parent, int p, int W, int lane)
{
    constexpr int NJ = 2 * NW;
    int S = 0, prev_card = 0;
    u32 key[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { const int u = lane + 32 * j; key[j] = (u < p) ? (u32)(0xFFFFu - u) : 0u; }
    // (vertex 65535 - u never collides with key 0: u < 512)
    for (int step = 0; step < p; ++step) {
        u32 best = 0u;
#pragma unroll
        for (int j = 0; j < NJ; ++j) best = key[j] > best ? key[j] : best;
        best = __reduce_max_sync(PDG_FULL, best);
        const int v = (int)(0xFFFFu - (best & 0xFFFFu)), new_card = (int)(best >> 16);
        // the row of v, without v itself and split into numbered / not yet numbered neighbours
        u64 xw[NW], nb[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            xw[w] = (w < W) ? (ADJ_L2 ? __ldcg(&adj[(size_t)v * W + w]) : adj[(size_t)v * W + w]) : 0ull;
            nb[w] = (w < W) ? numbered[w] : ~0ull;
            xw[w] &= ~((u64)((v >> 6) == w) << (v & 63));       // (arithmetic: keeps xw[] in registers)
        }
        if (new_card <= prev_card) {
            u32 lastkey = 0u;
            u64 xl = 0ull;
#pragma unroll
            for (int w = 0; w < NW; ++w) xl |= xw[w] & (0ull - (u64)(lane == w));     // this lane's word of the row
            if (lane < W) {
                const u64 x0 = xl & numbered[lane];                                  // the separator: numbered neighbours
                K[(size_t)S * W + lane] = x0; sep[(size_t)S * W + lane] = x0;
                u64 x = x0;
                while (x) {
                    const int b = __ffsll((long long)x) - 1;
                    x &= x - 1;
                    const int g = lane * 64 + b;
                    const u32 k2 = ((u32)order[g] << 16) | (u32)g;
                    lastkey = k2 > lastkey ? k2 : lastkey;
                }
            }
            lastkey = __reduce_max_sync(PDG_FULL, lastkey);
            if (lane == 0) parent[S] = (new_card != 0) ? (short)cliqueof[lastkey & 0xFFFFu] : (short)(S == 0 ? -1 : 0);
            ++S;
        }
        __syncwarp();
        if (lane == 0) {
            cliqueof[v] = (unsigned short)(S - 1);
            K[(size_t)(S - 1) * W + (v >> 6)] |= 1ull << (v & 63);
            order[v] = (unsigned short)step;
        }
        // cardinalities of the not yet numbered neighbours, v leaves the candidates
        const int vslot = (lane == (v & 31)) ? (v >> 5) : 99;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const u64 xu = xw[j >> 1] & ~nb[j >> 1];
            const u32 inc = (u32)(xu >> (lane + 32 * (j & 1))) & 1u;
            key[j] += inc << 16;                        // (xu holds unnumbered vertices below p only)
            key[j] &= 0u - (u32)(vslot != j);           // (arithmetic mask: no dynamic register indexing)
        }
        __syncwarp();
        if (lane == (v >> 6)) numbered[lane] |= 1ull << (v & 63);
        prev_card = new_card;
        __syncwarp();
    }
    return S;
}

#ifdef PDG_PHASE_TIMERS
static __device__ unsigned long long g_rz[8];     // developer probe: cycles of the randomiser's stages (adjacency, search, separators, relabel, derived)
#define RZ(i) { const long long n_ = clock64(); if (lane == 0) atomicAdd(&g_rz[i], (unsigned long long)(n_ - rz_t)); rz_t = n_; }
#else
#define RZ(i)
#endif
// randomisation of ONE chain by ONE warp; `smem` = rand_smem_layout(p, W).total bytes of warp-private
// shared memory.  Works on the global-memory state and rebuilds N and the CSR at the end.
static __device__ __noinline__ void randomize_chain(RandArgs P, RandScratch Z, long long chain, u32 iter,
                                             unsigned char *smem, int lane)
{
    const int p = P.p, W = P.W;
    const RandSmem L = rand_smem_layout(p, W);
    unsigned short *card = (unsigned short *)(smem + L.off_card), *order = (unsigned short *)(smem + L.off_order);
    unsigned short *cliqueof = (unsigned short *)(smem + L.off_cliqueof), *keys = (unsigned short *)(smem + L.off_keys);
    short *parent = (short *)(smem + L.off_parent);
    unsigned short *jdraw = (unsigned short *)(smem + L.off_jdraw), *comp = (unsigned short *)(smem + L.off_comp);
    unsigned short *nodes = (unsigned short *)(smem + L.off_nodes), *rank = (unsigned short *)(smem + L.off_rank);
    unsigned short *cstart = (unsigned short *)(smem + L.off_cstart), *canon = (unsigned short *)(smem + L.off_canon);
    unsigned short *compidx = (unsigned short *)(smem + L.off_compidx), *vind = (unsigned short *)(smem + L.off_vind);
    unsigned short *wj = (unsigned short *)(smem + L.off_wj), *cnt = (unsigned short *)(smem + L.off_cnt);
    unsigned short *fill = (unsigned short *)(smem + L.off_fill);
    unsigned short *newe = (unsigned short *)(smem + L.off_newe);
    unsigned char *done = smem + L.off_done, *inT = smem + L.off_inT, *inw = smem + L.off_inw, *isroot = smem + L.off_isroot;
    u64 *numbered = (u64 *)(smem + L.off_numbered), *used = (u64 *)(smem + L.off_used), *X = (u64 *)(smem + L.off_X);
    unsigned char *wmask = smem + L.off_wmask;
    unsigned short *sepsz = (unsigned short *)(smem + L.off_sepsz);
    u64 *Mc = P.M + (size_t)chain * p * W;
    const u64 *Nc = P.N + (size_t)chain * p * W;
    u64 *adj = Z.adj + (size_t)chain * p * W, *K = Z.K + (size_t)chain * p * W, *sep = Z.sep + (size_t)chain * p * W;
    if (L.off_sets >= 0) { adj = (u64 *)(smem + L.off_sets); K = adj + (size_t)p * W; sep = K + (size_t)p * W; }
    int *E = P.edges + (size_t)chain * 2 * (p - 1);
    const u32 gchain = P.chain0 + (u32)chain, k0 = P.seed_lo, k1 = P.seed_hi;

#ifdef PDG_PHASE_TIMERS
    long long rz_t = clock64();
#endif
    int S = 0;
    // Two codings of steps (a) and (b) with identical results: up to 128 vertices the gather /
    // shared-memory form is the shorter one, above it the scatter / register form (1.4 -> 1.0 ms
    // per call at p = 500).
    if (W <= 2) {
        // ---- (a) graph adjacency: adj[v] = union of the cliques holding v, minus v (:201-219) ----
        for (int i = lane; i < p * W; i += 32) {
            const int v = i / W, w = i - v * W;
            u64 acc = 0ull;
            for (int tw = 0; tw < W; ++tw) {
                u64 x = Nc[(size_t)v * W + tw];
                while (x) {
                    const int b = __ffsll((long long)x) - 1;
                    x &= x - 1;
                    acc |= Mc[(size_t)(tw * 64 + b) * W + w];
                }
            }
            if ((v >> 6) == w) acc &= ~(1ull << (v & 63));
            adj[i] = acc;
            K[i] = 0ull; sep[i] = 0ull;
        }
        for (int i = lane; i < p; i += 32) { card[i] = 0; done[i] = 0; inT[i] = 0; }
        if (lane < W) { numbered[lane] = 0ull; used[lane] = 0ull; }
        __syncwarp();

        S = mcs_cliques<2, false>(adj, K, sep, numbered, order, cliqueof, parent, p, W, lane);

    } else {
        // ---- (a) graph adjacency: adj[v] = union of the cliques holding v (:201-219; v's own bit is
        // masked where the row is used).  Scatter form: lane <-> tree node, one fire-and-forget OR per
        // (member, non-empty word) instead of a chain of dependent loads per (vertex, word).
        for (int i = lane; i < p * W; i += 32) { adj[i] = 0ull; K[i] = 0ull; sep[i] = 0ull; }
        for (int i = lane; i < p; i += 32) { done[i] = 0; inT[i] = 0; }
        if (lane < W) { numbered[lane] = 0ull; used[lane] = 0ull; }
        __threadfence();          // the zeroes must be in L2 before another lane's OR reaches the same word
        __syncwarp();
        for (int t = lane; t < p; t += 32) {
            u64 c[8];
    #pragma unroll
            for (int w = 0; w < 8; ++w) c[w] = (w < W) ? Mc[(size_t)t * W + w] : 0ull;
    #pragma unroll
            for (int w1 = 0; w1 < 8; ++w1) {
                u64 x = c[w1];
                while (x) {
                    const int v = w1 * 64 + __ffsll((long long)x) - 1;
                    x &= x - 1;
    #pragma unroll
                    for (int w2 = 0; w2 < 8; ++w2) if (c[w2]) atomicOr(&adj[(size_t)v * W + w2], c[w2]);
                }
            }
        }
        __threadfence();          // fire-and-forget reductions: performed before any lane reads a row back
        __syncwarp();
        RZ(0)
        S = mcs_cliques<8, true>(adj, K, sep, numbered, order, cliqueof, parent, p, W, lane);
    }

    RZ(1)
    // per clique: which words are non-empty (W <= 8 -> one byte) and the separator size, in shared
    // memory, so that the per-separator scans below reject almost every clique without a global load
    for (int d = lane; d < S; d += 32) {
        unsigned int m = 0u; int sz = 0;
        for (int w = 0; w < W; ++w) { if (K[(size_t)d * W + w]) m |= 1u << w; sz += __popcll(sep[(size_t)d * W + w]); }
        wmask[d] = (unsigned char)m; sepsz[d] = (unsigned short)sz;
    }
    __syncwarp();
    // ---- (c) every distinct separator: components of the induced forest + generalised Pruefer
    int ne = 0;
    for (int c = 1; c < S; ++c) {
        if (done[c]) continue;           // warp-uniform (shared memory)
        if (lane < W) X[lane] = sep[(size_t)c * W + lane];
        __syncwarp();
        // cliques of an AR-like graph are local: the separator's bits sit in one or two words
        int wlo = W, whi = -1, xsz = 0;
        unsigned int xm = 0u;
        for (int w = 0; w < W; ++w) if (X[w]) { if (wlo == W) wlo = w; whi = w; xm |= 1u << w; xsz += __popcll(X[w]); }
        int P_ = 0, q = 0;
        for (int base = 0; base < S; base += 32) {
            const int d = base + lane;
            bool in = false, root = false, dn = false;
            if (d < S) {
                in = ((wmask[d] & xm) == xm) && subset_r(X, K + (size_t)d * W, wlo, whi);
                if (in) {
                    const int pd = parent[d];
                    const bool top = (pd < 0) || ((wmask[pd] & xm) != xm) || !subset_r(X, K + (size_t)pd * W, wlo, whi);
                    // sep[d] == X  <=>  same size and X inside sep[d]
                    const bool eq = (d > 0) && sepsz[d] == xsz && subset_r(X, sep + (size_t)d * W, wlo, whi);
                    root = top || eq;
                    dn = !top && eq;
                }
            }
            const u32 bin = __ballot_sync(PDG_FULL, in), broot = __ballot_sync(PDG_FULL, root);
            if (in) {
                nodes[P_ + __popc(bin & ((1u << lane) - 1u))] = (unsigned short)d;
                inT[d] = 1; isroot[d] = root ? 1 : 0;
                if (dn) done[d] = 1;
                if (root) rank[d] = (unsigned short)(q + __popc(broot & ((1u << lane) - 1u)));
            }
            P_ += __popc(bin); q += __popc(broot);
        }
        __syncwarp();
        for (int i = lane; i <= q; i += 32) { cstart[i] = 0; }
        for (int i = lane; i < q; i += 32) { cnt[i] = 0; inw[i] = 1; fill[i] = 0; }
        __syncwarp();
        if (lane == 0) {
            for (int a = 0; a < P_; ++a) {
                const int d = nodes[a];
                comp[d] = isroot[d] ? (unsigned short)d : comp[parent[d]];
                cstart[rank[comp[d]] + 1] += 1;
            }
            for (int i = 0; i < q; ++i) cstart[i + 1] += cstart[i];
            for (int a = 0; a < P_; ++a) {
                const int d = nodes[a], i = rank[comp[d]];
                const int pos = cstart[i] + fill[i]++;
                canon[pos] = (unsigned short)d; compidx[pos] = (unsigned short)i;
            }
        }
        __syncwarp();
        for (int t = lane; t < q - 2; t += 32)
            vind[t] = (unsigned short)draw_int(k0, k1, gchain, iter, TAG_PRUF_V, ((u32)c << 16) | (u32)t, (u32)P_);
        for (int i = lane; i < q; i += 32)
            wj[i] = (unsigned short)draw_int(k0, k1, gchain, iter, TAG_PRUF_W, ((u32)c << 16) | (u32)i, (u32)(cstart[i + 1] - cstart[i]));
        __syncwarp();
        if (lane == 0) {
            // junction_tree.py:749-769 with a descending pointer instead of the rescans
            for (int t = 0; t < q - 2; ++t) cnt[compidx[vind[t]]] += 1;
            int ptr = q - 1, pending = -1;
            for (int t = q - 3; t >= 0; --t) {
                int x;
                if (pending >= 0) { x = pending; pending = -1; }
                else { while (!(inw[ptr] && cnt[ptr] == 0)) --ptr; x = ptr; }
                const int y = vind[t], cy = compidx[y];
                newe[2 * ne] = canon[cstart[x] + wj[x]]; newe[2 * ne + 1] = canon[y]; ++ne;
                inw[x] = 0;
                cnt[cy] -= 1;
                if (cnt[cy] == 0 && inw[cy] && cy > ptr) pending = cy;
            }
            int a = -1, b = -1;
            for (int i = 0; i < q; ++i) if (inw[i]) { if (a < 0) a = i; else if (b < 0) b = i; }
            newe[2 * ne] = canon[cstart[a] + wj[a]]; newe[2 * ne + 1] = canon[cstart[b] + wj[b]]; ++ne;
        }
        ne = __shfl_sync(PDG_FULL, ne, 0);
        for (int a = lane; a < P_; a += 32) inT[nodes[a]] = 0;
        __syncwarp();
    }
    if (ne != (S > 0 ? S - 1 : 0)) { if (lane == 0) atomicOr(&P.err[chain], ERR_NUMERIC); }
    RZ(2)

    // ---- (d) relabel + pad (junction_tree.py:36-64) ------------------------------------------
    for (int i = lane; i < p; i += 32) {
        keys[i] = (unsigned short)i;
        jdraw[i] = (i >= 1) ? (unsigned short)draw_int(k0, k1, gchain, iter, TAG_PERM, (u32)i, (u32)(i + 1)) : 0;
    }
    __syncwarp();
    if (lane == 0) {
        for (int i = p - 1; i >= 1; --i) { const int j = jdraw[i]; const unsigned short t = keys[i]; keys[i] = keys[j]; keys[j] = t; }
        for (int c = 0; c < S; ++c) used[keys[c] >> 6] |= 1ull << (keys[c] & 63);
    }
    for (int i = lane; i < p * W; i += 32) Mc[i] = 0ull;
    __syncwarp();
    for (int i = lane; i < S * W; i += 32) { const int c = i / W, w = i - c * W; Mc[(size_t)keys[c] * W + w] = K[i]; }
    for (int e = lane; e < ne; e += 32) { E[2 * e] = keys[newe[2 * e]]; E[2 * e + 1] = keys[newe[2 * e + 1]]; }
    for (int t = lane; t < p; t += 32) {
        if (bit64(used, t)) continue;
        int rk = 0;                                   // unused tree nodes below t
        for (int w = 0; w < (t >> 6); ++w) rk += 64 - __popcll(used[w]);
        rk += (t & 63) - __popcll(used[t >> 6] & ((1ull << (t & 63)) - 1ull));
        const int n = S + rk;
        const int j = (int)draw_int(k0, k1, gchain, iter, TAG_PAD, (u32)t, (u32)n);
        int target;
        if (j < S) target = keys[j];
        else {
            int r = j - S; target = -1;
            for (int w = 0; w < W && target < 0; ++w) {
                u64 x = ~used[w];
                if (w == W - 1 && (p & 63)) x &= (1ull << (p & 63)) - 1ull;
                const int cfree = __popcll(x);
                if (r < cfree) target = w * 64 + select64(x, r); else r -= cfree;
            }
        }
        E[2 * (ne + rk)] = t; E[2 * (ne + rk) + 1] = target;
    }
    __syncwarp();
    __threadfence_block();
    RZ(3)
    build_derived(P, chain, lane, (int *)(smem + L.off_deg32));
    RZ(4)
#ifdef PDG_PHASE_TIMERS
    if (lane == 0) { atomicAdd(&g_rz[5], 1ull); atomicAdd(&g_rz[6], (unsigned long long)S); }
#endif
}

#ifdef PDG_DEFINE_GLOBAL_KERNELS
__global__ void __launch_bounds__(32) k_randomize(PdgParams P, RandScratch Z, u32 iter)
{
    extern __shared__ __align__(16) unsigned char smem[];
    randomize_chain(rand_args(P), Z, blockIdx.x, iter, smem, threadIdx.x);
}
#endif
