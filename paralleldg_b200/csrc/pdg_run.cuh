// The fused Metropolis-Hastings kernel: proposal enumeration on the latent tree, clique scoring,
// Philox accept/reject, state update and update-record append, for iterations [it_begin, it_end)
// of every chain (one CTA per chain).  Restates mh_parallel.py:225-304 (parallel moves) and
// :66-168 (single move) with python sets replaced by ballot-built bitsets.
#pragma once
#include "pdg_kernels.cuh"

// shared-memory carve-up, identical on host and device
struct RunSmem {
    int off_tri, off_idx, off_delta, off_dl, off_dp, off_u, off_sub, off_leaf, off_bnd, off_partner,
        off_mvU, off_mvUadj, off_acc, off_csr_off, off_csr_adj, off_work, total;
};

__host__ __device__ inline RunSmem run_smem_layout(int p, int W, int nwarps)
{
    RunSmem s;
    int o = 0;
    const int np32 = ((p + 31) / 32) + 1;
    s.off_tri = o;      o += nwarps * 528 * 8;
    s.off_work = o;     o += nwarps * 4 * 8;
    s.off_delta = o;    o += p * 8;
    s.off_dl = o;       o += p * 8;
    s.off_dp = o;       o += p * 8;
    s.off_u = o;        o += p * 8;
    s.off_sub = o;      o += ((W * 8 + 15) / 16) * 16;
    s.off_leaf = o;     o += ((np32 * 4 + 15) / 16) * 16;
    s.off_bnd = o;      o += ((np32 * 4 + 15) / 16) * 16;
    s.off_idx = o;      o += nwarps * (((p + 1) * 2 + 15) / 16) * 16;
    s.off_partner = o;  o += ((p * 2 + 15) / 16) * 16;
    s.off_mvU = o;      o += ((p * 2 + 15) / 16) * 16;
    s.off_mvUadj = o;   o += ((p * 2 + 15) / 16) * 16;
    s.off_csr_off = o;  o += (((p + 1) * 2 + 15) / 16) * 16;
    s.off_csr_adj = o;  o += ((2 * p * 2 + 15) / 16) * 16;
    s.off_acc = o;      o += ((p + 15) / 16) * 16;
    s.total = o;
    return s;
}

// n-th set bit of a multi-word (32-bit words) bitset
__device__ __forceinline__ int select_multi32(const u32 *w, int nwords, int n)
{
    for (int i = 0; i < nwords; ++i) {
        const int c = __popc(w[i]);
        if (n < c) return i * 32 + select32(w[i], n);
        n -= c;
    }
    return -1;
}

template <int MODEL, int KIND>
__global__ void __launch_bounds__(128) k_mh_run(PdgParams P, long long it_begin, long long it_end, PdgReplayDev R)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int p = P.p, W = P.W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, nw = NT >> 5;
    const long long chain = blockIdx.x;
    const RunSmem L = run_smem_layout(p, W, nw);
    double *s_delta = (double *)(smem + L.off_delta);
    double *s_dl = (double *)(smem + L.off_dl);
    double *s_dp = (double *)(smem + L.off_dp);
    double *s_u = (double *)(smem + L.off_u);
    u64 *s_sub = (u64 *)(smem + L.off_sub);
    u32 *s_sub32 = (u32 *)(smem + L.off_sub);
    u32 *s_leaf = (u32 *)(smem + L.off_leaf);
    u32 *s_bnd = (u32 *)(smem + L.off_bnd);
    short *s_partner = (short *)(smem + L.off_partner);
    short *s_mvU = (short *)(smem + L.off_mvU);
    short *s_mvUadj = (short *)(smem + L.off_mvUadj);
    unsigned char *s_acc = (unsigned char *)(smem + L.off_acc);
    unsigned short *s_off = (unsigned short *)(smem + L.off_csr_off);
    unsigned short *s_adj = (unsigned short *)(smem + L.off_csr_adj);
    WarpScratch ws;
    ws.tri32 = (double *)(smem + L.off_tri) + warp * 528;
    ws.idx = (short *)(smem + L.off_idx) + warp * ((((p + 1) * 2 + 15) / 16) * 8);
    ws.kbig = P.kbig;
    ws.big = P.big ? P.big + ((size_t)chain * nw + warp) * ((size_t)(P.kbig + 1) * (P.kbig + 2) / 2) : nullptr;

    u64 *Mc = P.M + (size_t)chain * p * W;
    u64 *Nc = P.N + (size_t)chain * p * W;
    for (int i = tid; i <= p; i += NT) s_off[i] = P.csr_off[(size_t)chain * (p + 1) + i];
    for (int i = tid; i < 2 * (p - 1); i += NT) s_adj[i] = P.csr_adj[(size_t)chain * 2 * p + i];
    const int nw32 = (p + 31) >> 5;
    const u32 gchain = P.chain0 + (u32)chain;
    long long k = P.kcount[chain];
    long long nrec = P.nrec[chain];
    double logl_prev = P.logl[(size_t)chain * P.n_samples + it_begin - 1];
    int errflag = 0;
    const int npass = (p + NT - 1) / NT;
    unsigned long long *s_work = (unsigned long long *)(smem + L.off_work) + warp * 4;
    if (lane < 4) s_work[lane] = 0ull;

    for (long long it = it_begin; it < it_end; ++it) {
        // ---- draws (mh_parallel.py:230-231) -------------------------------------------
        int node, mtype, aux = -1;
        u32 r[4] = {0u, 0u, 0u, 0u};
        if (R.on) {
            const size_t o = (size_t)chain * R.n_samples + it;
            node = R.it_node[o]; mtype = R.it_mtype[o]; aux = R.it_aux[o];
        } else {
            philox4x32_10(gchain, (u32)it, TAG_ITER, 0u, P.seed_lo, P.seed_hi, r);
            node = (int)__umulhi(r[0], (u32)p);
            mtype = (int)(r[1] >> 31);
        }
        if (tid < W) s_sub[tid] = Nc[(size_t)node * W + tid];
        __syncthreads();
        int msub = 0;
        for (int w = 0; w < W; ++w) msub += __popcll(s_sub[w]);
        if (msub == 0) mtype = 0;          // :237-240 forced move types
        if (msub == p) mtype = 1;
        // ---- one scan of the tree: leaves of the induced subtree and its outer boundary ----
        // (parallel_moves.py:193-208 and :253-272)
        for (int ps = 0; ps < npass; ++ps) {
            const int t = ps * NT + tid;
            bool leaf = false, bnd = false;
            if (t < p) {
                const int inT = bit32(s_sub32, t);
                int cnt = 0, last = -1;
                for (int e = s_off[t]; e < s_off[t + 1]; ++e) {
                    const int a = s_adj[e];
                    if (bit32(s_sub32, a)) { ++cnt; last = a; }
                }
                leaf = inT && cnt == 1;
                bnd = !inT && cnt >= 1;
                s_partner[t] = (short)last;
            }
            const u32 bl = __ballot_sync(PDG_FULL, leaf), bb = __ballot_sync(PDG_FULL, bnd);
            if (lane == 0 && (ps * NT + warp * 32) < nw32 * 32) { s_leaf[(ps * NT >> 5) + warp] = bl; s_bnd[(ps * NT >> 5) + warp] = bb; }
        }
        __syncthreads();
        // ---- ordered proposal list ------------------------------------------------------
        int nm, nfwd;
        double logq = 0.0;
        const u32 *set = (mtype == 0) ? s_bnd : s_leaf;
        int nset = 0;
        for (int i = 0; i < nw32; ++i) nset += __popc(set[i]);
        const bool special = (mtype == 0 && msub == 0) || (mtype == 1 && msub <= 2);
        if (special) {
            nm = 1;
            if (tid == 0) {
                if (mtype == 0) {                       // pm:138-139 one random tree node
                    s_mvU[0] = (short)(R.on ? aux : (int)__umulhi(r[2], (u32)p));
                    s_mvUadj[0] = -1;
                } else if (msub == 1) {                 // pm:186-187
                    s_mvU[0] = (short)select_multi32(s_sub32, nw32, 0);
                    s_mvUadj[0] = -1;
                } else {                                // pm:184-185 one of the two leaves
                    const int a = select_multi32(s_sub32, nw32, 0), b = select_multi32(s_sub32, nw32, 1);
                    int pick;
                    if (R.on) { pick = (aux == a) ? 0 : 1; if (aux != a && aux != b) errflag |= ERR_REPLAY; }
                    else pick = (int)(r[2] >> 31);
                    s_mvU[0] = (short)(pick ? b : a);
                    s_mvUadj[0] = (short)(pick ? a : b);
                }
            }
            nfwd = (mtype == 1 && msub == 2) ? 2 : 1;
            if (KIND == PDG_KIND_PARALLEL && mtype == 1 && msub == 2) logq = -0.6931471805599453;   // :277
        } else {
            nm = nset; nfwd = nset;
            if (KIND == PDG_KIND_PARALLEL) {
                if (R.on) {
                    // the reference's ORDER (python set iteration) with the SET checked here
                    const long long o = R.it_off[(size_t)chain * (R.n_samples + 1) + it];
                    const long long cnt = R.it_off[(size_t)chain * (R.n_samples + 1) + it + 1] - o;
                    if (cnt != nm) { errflag |= ERR_REPLAY; nm = (int)(cnt < nm ? cnt : nm); }
                    for (int j = tid; j < nm; j += NT) {
                        const int U = R.mv_U[o + j], Ua = R.mv_Uadj[o + j];
                        if (U < 0 || U >= p || !bit32(set, U) || s_partner[U] != Ua) errflag |= ERR_REPLAY;
                        s_mvU[j] = (short)U; s_mvUadj[j] = (short)Ua;
                    }
                } else {
                    for (int j = tid; j < nm; j += NT) {
                        const int U = select_multi32(set, nw32, j);
                        s_mvU[j] = (short)U; s_mvUadj[j] = s_partner[U];
                    }
                }
            } else {
                // single move: moves[0] after np.random.shuffle (:87,:95 / :128,:138)
                if (tid == 0) {
                    int U, Ua;
                    if (R.on) {
                        const long long o = R.it_off[(size_t)chain * (R.n_samples + 1) + it];
                        U = R.mv_U[o]; Ua = R.mv_Uadj[o];
                        if (U < 0 || U >= p || !bit32(set, U) || s_partner[U] != Ua) errflag |= ERR_REPLAY;
                    } else {
                        const int pick = nset > 1 ? (int)__umulhi(r[3], (u32)nset) : 0;
                        U = select_multi32(set, nw32, pick); Ua = s_partner[U];
                    }
                    s_mvU[0] = (short)U; s_mvUadj[0] = (short)Ua;
                }
                nm = 1;
            }
        }
        __syncthreads();
        if (KIND == PDG_KIND_SINGLE) {
            // Hastings term log_q2 - log_q1 = -log(num_moves) + log(num_reverse)  (:111-118, :154-160)
            const int U = s_mvU[0], Ua = s_mvUadj[0];
            int nrev;
            if (mtype == 0) {
                int nl = 0;
                if (msub >= 2) for (int i = 0; i < nw32; ++i) nl += __popc(s_leaf[i]);
                const int notin = (Ua >= 0 && msub >= 2 && bit32(s_leaf, Ua)) ? 0 : 1;   // pm:276-281
                nrev = nl + notin + (msub == 1 ? 1 : 0);
            } else {
                if (msub == 1) nrev = p;                                                   // :133-134
                else {
                    int nb = 0;
                    for (int i = 0; i < nw32; ++i) nb += __popc(s_bnd[i]);
                    nrev = nb + 2 - ((int)s_off[U + 1] - (int)s_off[U]);                   // pm:283-284
                }
            }
            const double log_q2 = -log((double)nfwd), log_q1 = -log((double)nrev);
            logq = log_q2 - log_q1;
        }
        // ---- score + accept: one warp per proposal ---------------------------------------
        for (int j = warp; j < nm; j += nw) {
            const int U = s_mvU[j], Ua = s_mvUadj[j];
            const u64 *Crow = Mc + (size_t)U * W;
            const u64 *Arow = (Ua >= 0) ? Mc + (size_t)Ua * W : nullptr;
            int kb = 0, ks = 0;
            double sc[4] = {0.0, 0.0, 0.0, 0.0};
            if (MODEL == PDG_MODEL_GGM) {
                const int e = ggm_move_scores(P, ws, Crow, Arow, node, lane, kb, ks, sc);
                if (e) errflag |= e;
            } else {
                for (int w = 0; w < W; ++w) {
                    u64 b = Crow[w];
                    if ((node >> 6) == w) b &= ~(1ull << (node & 63));
                    kb += __popcll(b);
                    if (Arow) ks += __popcll(b & Arow[w]);
                }
            }
            if (lane == 0) {
                const bool adj = Ua >= 0;
                if (MODEL != PDG_MODEL_UNIFORM) {
                    // algorithmic work of this proposal: the complete sets the reference scores
                    const int ksz[4] = {kb, ks, kb + 1, adj ? ks + 1 : 0};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const unsigned long long kk_ = (unsigned long long)ksz[q];
                        if (!kk_) continue;
                        s_work[0] += 8ull * kk_ * kk_ * (P.dmode == 2 ? 2ull : 1ull);
                        s_work[1] += (2ull * kk_ * kk_ * kk_ + 3ull * kk_ * kk_ + 6ull * kk_) * (P.dmode == 2 ? 2ull : 1ull);
                        s_work[2] += 1ull; s_work[3] += kk_;
                    }
                }
                double log_p2, log_p1, log_g2, log_g1;
                if (mtype == 0) {
                    const int cC = kb, cS = ks, cCn = kb + 1, cSn = adj ? ks + 1 : 0;
                    log_p2 = sc[2] - (adj ? sc[3] : 0.0);
                    log_p1 = sc[0] - sc[1];
                    log_g2 = prior_partial(P, cCn, cSn, 0);
                    log_g1 = prior_partial(P, cC, cS, 1);
                    if (msub == 0) { log_p1 += sc[3]; log_g1 += prior_partial(P, 1, cS, 1); }   // :261-263
                } else {
                    const int cC = kb + 1, cS = adj ? ks + 1 : 0, cCn = kb, cSn = ks;
                    log_p2 = sc[0] - sc[1];
                    log_p1 = sc[2] - (adj ? sc[3] : 0.0);
                    log_g2 = prior_partial(P, cCn, cSn, 1);
                    log_g1 = prior_partial(P, cC, cS, 0);
                    if (msub == 1) { log_p2 += sc[3]; log_g2 += prior_partial(P, 1, cSn, 1); }  // :293-295
                }
                const double dl = log_p2 - log_p1, dp = log_g2 - log_g1;
                const double tot = (dl + dp) + logq;
                double u;
                if (R.on) u = R.mv_u[R.it_off[(size_t)chain * (R.n_samples + 1) + it] + j];
                else u = draw_unif(P.seed_lo, P.seed_hi, gchain, (u32)it, TAG_MOVE, (u32)U);
                const double e = exp(tot);
                const bool acc = u <= (e < 1.0 ? e : 1.0);                                        // :268,:300
                s_delta[j] = dl + dp; s_acc[j] = acc ? 1 : 0;
                s_dl[j] = dl; s_dp[j] = dp; s_u[j] = u;
            }
        }
        __syncthreads();
        // ---- apply in proposal order: records, bit flips, running log-probability ---------
        if (warp == 0) {
            double log_p = 0.0;
            for (int j = 0; j < nm; ++j) {
                const long long kk = k + j + 1;
                const int U = s_mvU[j], Ua = s_mvUadj[j];
                if (P.mvlog_cap && lane == 0 && kk - 1 < P.mvlog_cap) {
                    const size_t o = (size_t)chain * P.mvlog_cap + (kk - 1);
                    P.ml_dlik[o] = s_dl[j]; P.ml_dprior[o] = s_dp[j]; P.ml_logq[o] = logq; P.ml_u[o] = s_u[j];
                    P.ml_acc[o] = s_acc[j]; P.ml_U[o] = U;
                }
                if (!s_acc[j]) continue;
                log_p += s_delta[j];
                if (nrec < P.rec_cap) {
                    const size_t ro = (size_t)chain * P.rec_cap + nrec;
                    if (lane == 0) {
                        pdg_update h;
                        h.i = (int)it; h.k = (int)kk; h.node = (unsigned short)node; h.U = (unsigned short)U;
                        h.Uadj = (short)Ua; h.type = (unsigned char)mtype; h.pad = 0;
                        P.rec[ro] = h;
                    }
                    for (int w = lane; w < 2 * W; w += 32) {
                        u64 v;
                        if (w < W) v = Mc[(size_t)U * W + w];
                        else v = (Ua >= 0) ? Mc[(size_t)Ua * W + (w - W)] : 0ull;
                        P.rec_bits[ro * 2 * W + w] = v;
                    }
                } else errflag |= ERR_OVERFLOW;
                ++nrec;
                __syncwarp();
                if (lane == 0) {                         // junction_tree.py:78-100 connect / disconnect
                    Mc[(size_t)U * W + (node >> 6)] ^= 1ull << (node & 63);
                    Nc[(size_t)node * W + (U >> 6)] ^= 1ull << (U & 63);
                }
                __syncwarp();
            }
            logl_prev = logl_prev + log_p;               // :304 running sum
            if (lane == 0) P.logl[(size_t)chain * P.n_samples + it] = logl_prev;
        }
        k += nm;
        __syncthreads();
    }
    if (tid == 0) { P.kcount[chain] = k; P.nrec[chain] = nrec; }
    __syncwarp();
    if (lane < 4 && s_work[2]) atomicAdd(&P.work[chain * 4 + lane], s_work[lane]);
    if (errflag) atomicOr(&P.err[chain], errflag);
}
