// Kernels of the parallelDG B200 hot path.  One CTA owns one chain for the whole launch; the
// per-chain junction-tree state (clique bit-matrix M, its transpose N, tree CSR) lives in
// global memory laid out chain-major (L2 resident: 1024 chains x 64 KB at p = 500) and the
// per-iteration scratch in shared memory.
#pragma once
#include "pdg_device.cuh"
#include "../../include/pdg_b200.h"

struct PdgParams {
    int p, W;
    int model, prior;
    double prior_a, prior_b;
    long long n_chains;
    u32 chain0, seed_lo, seed_hi;
    long long n_samples;
    // chain state
    u64 *M, *N;                 // [chain][p][W]
    int *edges;                 // [chain][p-1][2]
    unsigned short *csr_off;    // [chain][p+1]
    unsigned short *csr_adj;    // [chain][2p]
    // GGM
    const double *A, *Dm, *logDdiag, *gconst;
    int dmode;                  // 0 identity, 1 diagonal, 2 general
    double delta, nobs;
    double *big;                // [chain][warp] packed triangles for cliques > 32
    int kbig;                   // largest clique the scratch supports
    // outputs
    double *logl;               // [chain][n_samples]
    long long *kcount, *nrec;   // [chain]
    pdg_update *rec;            // [chain][rec_cap]
    u64 *rec_bits;              // [chain][rec_cap][2][W]
    long long rec_cap;
    int *err;                   // [chain]
    unsigned long long *work;   // [chain][4]: bytes, 6*flops, scored sets, sum k
    u64 *G0;                    // [chain][p][W] graph adjacency at MCMC index 0 (for edge tallies)
    // per-proposal diagnostics (optional)
    long long mvlog_cap;
    double *ml_dlik, *ml_dprior, *ml_logq, *ml_u;
    int *ml_acc, *ml_U;
};

struct PdgReplayDev {
    int on;
    long long n_samples;
    const int *it_node, *it_mtype, *it_aux;
    const long long *it_off;
    const int *mv_U, *mv_Uadj;
    const double *mv_u;
};

enum { ERR_REPLAY = 1, ERR_OVERFLOW = 2, ERR_NUMERIC = 4, ERR_BIGCLIQUE = 8 };

// seqdist:31-151 log_prior_partial(clq, sep, extra_arg)
__device__ __forceinline__ double prior_partial(const PdgParams &P, int csize, int ssize, int extra)
{
    switch (P.prior) {
    case PDG_PRIOR_MBC: {
        double lp = 0.0;
        lp += P.prior_a * ((double)csize - 1.0);
        lp -= P.prior_b * (double)ssize;
        return lp;
    }
    case PDG_PRIOR_EDGEPENALTY: {
        const double a = (double)csize, b = (double)ssize;
        return (-P.prior_a * a * (a - 1.0) / 2.0) - (-P.prior_a * b * (b - 1.0) / 2.0);
    }
    case PDG_PRIOR_JUMPPENALTY:
        return -P.prior_a * (double)(1 - extra);
    default:
        return 0.0;
    }
}

// ggm:26-39 with wishart.py:17-37 folded: score(X) = gconst[k] - t1 ld(A_XX) + t2 ld(D_XX)
__device__ __forceinline__ double ggm_score_from(const PdgParams &P, int k, double ldA, double ldD)
{
    if (k == 0) return 0.0;
    const double t1 = (P.delta + P.nobs + (double)k - 1.0) * 0.5;
    const double t2 = (P.delta + (double)k - 1.0) * 0.5;
    return P.gconst[k] - t1 * ldA + t2 * ldD;
}

// per-warp scratch: packed 32x32 triangle + index list
struct WarpScratch {
    double *tri32;            // 528 doubles (shared)
    short *idx;               // p + 1 entries (shared)
    double *big;              // global scratch for k > 32 (may be null)
    int kbig;
};

// fills idx[0..k) = [S0 | B\S0 | v?] from the two bit rows; returns kb, ks through refs
__device__ __forceinline__ void build_index(const u64 *Crow, const u64 *Adjrow, int node, bool with_v,
                                            int W, int lane, short *idx, int &kb, int &ks)
{
    int cb = 0, cs = 0;
    for (int w = 0; w < W; ++w) {
        u64 b = Crow[w];
        if (node >= 0 && (node >> 6) == w) b &= ~(1ull << (node & 63));
        const u64 s = Adjrow ? (b & Adjrow[w]) : 0ull;
        cb += __popcll(b);
        cs += __popcll(s);
    }
    kb = cb; ks = cs;
    const int k = cb + (with_v ? 1 : 0);
    for (int i = lane; i < k; i += 32) {
        int g;
        if (i >= cb) g = node;
        else {
            const bool inS = i < cs;
            int n = inS ? i : i - cs;
            g = -1;
            for (int w = 0; w < W; ++w) {
                u64 b = Crow[w];
                if (node >= 0 && (node >> 6) == w) b &= ~(1ull << (node & 63));
                const u64 s = Adjrow ? (b & Adjrow[w]) : 0ull;
                const u64 x = inS ? s : (b & ~s);
                const int c = __popcll(x);
                if (n < c) { g = w * 64 + select64(x, n); break; }
                n -= c;
            }
        }
        idx[i] = (short)g;
    }
    __syncwarp();
}

// gathers the packed lower triangle of mat[idx, idx] and factorises it
__device__ __forceinline__ bool gather_ldl(const double *__restrict__ mat, int p, double *a, const short *idx,
                                           int lane, int k, int ks, int kb,
                                           double &sumS, double &sumB, double &ldv, double &lschur)
{
    for (int i = lane; i < k; i += 32) {
        const double *row = mat + (size_t)idx[i] * p;
        double *dst = a + tri(i, 0);
        for (int l = 0; l <= i; ++l) dst[l] = __ldg(row + idx[l]);
    }
    return ldl_logdets(a, lane, k, ks, kb, sumS, sumB, ldv, lschur);
}

// Scores the four complete sets of one move with a single factorisation (see ldl_logdets).
// out[0] = score(B), out[1] = score(S0), out[2] = score(B+v), out[3] = score(S0+v).
// Returns 0 or an ERR_* bit.
__device__ __forceinline__ int ggm_move_scores(const PdgParams &P, const WarpScratch &ws, const u64 *Crow,
                                               const u64 *Adjrow, int node, int lane, int &kb, int &ks,
                                               double out[4])
{
    build_index(Crow, Adjrow, node, true, P.W, lane, ws.idx, kb, ks);
    const int k = kb + 1;
    double *a = ws.tri32;
    if (k > 32) {
        if (!ws.big || k > ws.kbig) return ERR_BIGCLIQUE;
        a = ws.big;
    }
    double sS, sB, ldv, lsc;
    if (!gather_ldl(P.A, P.p, a, ws.idx, lane, k, ks, kb, sS, sB, ldv, lsc)) return ERR_NUMERIC;
    double dS = 0.0, dB = 0.0, dv = 0.0, dsc = 0.0;
    if (P.dmode == 1) {
        double accS = 0.0, accB = 0.0;
        for (int j = lane; j < kb; j += 32) { const double l = P.logDdiag[ws.idx[j]]; accB += l; if (j < ks) accS += l; }
        dS = warp_sum(accS); dB = warp_sum(accB); dv = P.logDdiag[node]; dsc = dv;
    } else if (P.dmode == 2) {
        if (!gather_ldl(P.Dm, P.p, a, ws.idx, lane, k, ks, kb, dS, dB, dv, dsc)) return ERR_NUMERIC;
    }
    out[0] = ggm_score_from(P, kb, sB, dB);
    out[1] = ggm_score_from(P, ks, sS, dS);
    out[2] = ggm_score_from(P, kb + 1, sB + ldv, dB + dv);
    out[3] = ggm_score_from(P, ks + 1, sS + lsc, dS + dsc);
    return 0;
}

// score of one complete set given as a bit row (used for logl[0] and the batch scorer)
__device__ __forceinline__ int ggm_set_score(const PdgParams &P, const WarpScratch &ws, const u64 *row, int lane,
                                             double &score)
{
    int kb, ks;
    build_index(row, nullptr, -1, false, P.W, lane, ws.idx, kb, ks);
    const int k = kb;
    if (k == 0) { score = 0.0; return 0; }
    double *a = ws.tri32;
    if (k > 32) {
        if (!ws.big || k > ws.kbig) return ERR_BIGCLIQUE;
        a = ws.big;
    }
    double sS, sB, ldv, lsc;
    if (!gather_ldl(P.A, P.p, a, ws.idx, lane, k, 0, k, sS, sB, ldv, lsc)) return ERR_NUMERIC;
    double dB = 0.0;
    if (P.dmode == 1) {
        double acc = 0.0;
        for (int j = lane; j < k; j += 32) acc += P.logDdiag[ws.idx[j]];
        dB = warp_sum(acc);
    } else if (P.dmode == 2) {
        double t0, t1, t2;
        if (!gather_ldl(P.Dm, P.p, a, ws.idx, lane, k, 0, k, t0, dB, t1, t2)) return ERR_NUMERIC;
    }
    score = ggm_score_from(P, k, sB, dB);
    return 0;
}
