// Shared definitions of the parallelDG B200 kernels: launch parameters, the device-wide memo
// table of clique scores, priors and the hyper-inverse-Wishart clique scorer.
//
// Memory layout in HBM (chain-major, SoA over chains):
//   M  [chain][p][W] u64   clique bit-matrix  (t2clique, junction_tree.py:36-42)
//   N  [chain][p][W] u64   its transpose       (node2t,  junction_tree.py:68-76)
//   edges [chain][p-1][2] i32, csr_off [chain][p+1] u16, csr_adj [chain][2p] u16   the tree `t`
//   logl [chain][n_samples] f64, rec/rec_bits [chain][rec_cap]  per-chain append buffers
// and one memo table per GPU shared by every chain (all chains score the same data set).
#pragma once
#include "pdg_device.cuh"
#include "../../include/pdg_b200.h"

#define PDG_WMAX_LIMIT 8            // p <= 512
#define MEMO_SENT 0x7FF8DEADBEEF0001ull
#define MEMO_PROBES 16

// Open-addressing memo of score(X), keyed by the bitset of the complete set X.  Values are a
// canonical function of the key (ascending-index factorisation / cell order), so concurrent
// writers store identical bits and a reader that races a writer merely recomputes.
struct MemoTable {
    ulonglong2 *kv;     // W == 1: {key, value bits}
    u64 *tag;           // W  > 1: 64-bit hash tag (0 = empty), key words and value arrays
    u64 *keyw;
    u64 *val;
    u32 mask;           // slots - 1
    int W;
};

// pool of hashed cell tables for log-linear sets with more cells than the dense per-warp table
struct LoglinSparse {
    int *lock;              // [nsp]
    u64 *hkeys;             // [nsp][hslots]
    unsigned int *hcnt;     // [nsp][hslots]
    unsigned int *cc;       // [nsp][nrows + 2] count-of-counts
    int *list;              // [nsp][nrows + 1] slots taken by the scan in progress; entry 0 = how many
    int nsp, hslots;
};

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
    double *big;                // [warp slot] packed triangles for cliques > 32
    int kbig;                   // largest clique the scratch supports
    int tri_extra, ktri_run;    // MH kernel: bytes added to the warp's shared triangle and the largest clique it holds
    // log-linear
    const unsigned char *data;  // [p][nrows] level codes, column-major
    const int *levels;          // [p]
    long long nrows;
    long long data_ld;        // column stride of `data` (nrows rounded up to 16), zero column at index p
    int ll_cap;               // log-linear: cells of the per-warp shared count table (0 for other models)
    double cell_alpha;
    int n_ktab;                 // lgamma tables: ktab[j] = cells, alphatab[j], lgtab[j][c]
    const double *ktab, *alphatab, *lgtab;
    double ll_const;            // lgamma(N + cell_alpha) - lgamma(cell_alpha)
    unsigned int *hist;         // [warp slot][hist_cells] dense cell counters
    int hist_cells;
    LoglinSparse sp;
    MemoTable memo;
    int memo_small;             // GGM: also memoise the sets of the lane-private path
    // outputs
    double *logl;               // [chain][n_samples]
    long long *kcount, *nrec;   // [chain]
    pdg_update *rec;            // [chain][rec_cap]
    u64 *rec_bits;              // [chain][rec_cap][2][W]
    long long rec_cap;
    int *err;                   // [chain]
    unsigned long long *work;   // [chain][8]: see pdg_get_work
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

// ---- memo table --------------------------------------------------------------------------
__device__ __forceinline__ u64 mix64(u64 k)
{
    k ^= k >> 30; k *= 0xBF58476D1CE4E5B9ull;
    k ^= k >> 27; k *= 0x94D049BB133111EBull;
    k ^= k >> 31;
    return k;
}

template <int WMAX>
__device__ __forceinline__ u64 key_hash(const u64 (&X)[WMAX], int W)
{
    // multilinear combination of the words (independent multiplies by unrelated odd constants), one
    // finaliser: a serial mix64 per word was 6 % of the p = 500 kernel.  (Multipliers in arithmetic
    // progression would let structured bitsets collide: sum(d_w) = sum(w d_w) = 0.)
    const u64 MUL[8] = {0xBF58476D1CE4E5B9ull, 0x94D049BB133111EBull, 0xD6E8FEB86659FD93ull, 0xFF51AFD7ED558CCDull,
                        0xC4CEB9FE1A85EC53ull, 0x9FB21C651E98DF25ull, 0xE7037ED1A0B428DBull, 0xA0761D6478BD642Full};
    u64 h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) if (w < W) h += (X[w] + 0x2545F4914F6CDD1Dull * (u64)(w + 1)) * MUL[w & 7];
    return mix64(h);
}

template <int WMAX>
__device__ __forceinline__ bool memo_lookup(const MemoTable &T, const u64 (&X)[WMAX], int W, double &val)
{
    if (WMAX == 1) {
        u32 h = (u32)mix64(X[0]) & T.mask;
#pragma unroll 1
        for (int pr = 0; pr < MEMO_PROBES; ++pr) {
            const ulonglong2 e = __ldcg(&T.kv[h]);
            if (e.x == X[0]) {
                if (e.y == MEMO_SENT) return false;
                val = __longlong_as_double((long long)e.y);
                return true;
            }
            if (e.x == 0ull) return false;
            h = (h + 1) & T.mask;
        }
        return false;
    } else {
        const u64 hh = key_hash<WMAX>(X, W);
        const u64 tag = hh | 1ull;
        u32 h = (u32)(hh >> 32) & T.mask;
#pragma unroll 1
        for (int pr = 0; pr < MEMO_PROBES; ++pr) {
            const u64 t = __ldcg(&T.tag[h]);
            if (t == 0ull) return false;
            if (t == tag) {
                // key words and value in flight together: one round trip, not W + 1
                u64 kw[WMAX];
#pragma unroll
                for (int w = 0; w < WMAX; ++w) kw[w] = (w < W) ? __ldcg(&T.keyw[(size_t)h * W + w]) : 0ull;
                const u64 v = __ldcg(&T.val[h]);
                bool same = true;
#pragma unroll
                for (int w = 0; w < WMAX; ++w) same = same && (kw[w] == X[w]);
                if (same) {
                    if (v == MEMO_SENT) return false;
                    val = __longlong_as_double((long long)v);
                    return true;
                }
            }
            h = (h + 1) & T.mask;
        }
        return false;
    }
}

template <int WMAX>
__device__ __forceinline__ void memo_insert(const MemoTable &T, const u64 (&X)[WMAX], int W, double val)
{
    const u64 vb = (u64)__double_as_longlong(val);
    if (WMAX == 1) {
        u32 h = (u32)mix64(X[0]) & T.mask;
#pragma unroll 1
        for (int pr = 0; pr < MEMO_PROBES; ++pr) {
            const u64 old = atomicCAS(&T.kv[h].x, 0ull, X[0]);
            if (old == 0ull || old == X[0]) { __stcg(&T.kv[h].y, vb); return; }
            h = (h + 1) & T.mask;
        }
    } else {
        const u64 hh = key_hash<WMAX>(X, W);
        const u64 tag = hh | 1ull;
        u32 h = (u32)(hh >> 32) & T.mask;
#pragma unroll 1
        for (int pr = 0; pr < MEMO_PROBES; ++pr) {
            const u64 old = atomicCAS(&T.tag[h], 0ull, tag);
            if (old == 0ull) {
#pragma unroll
                for (int w = 0; w < WMAX; ++w) if (w < W) __stcg(&T.keyw[(size_t)h * W + w], X[w]);
                __threadfence();
                __stcg(&T.val[h], vb);
                return;
            }
            if (old == tag) {
                bool same = true;
#pragma unroll
                for (int w = 0; w < WMAX; ++w) if (w < W) same = same && (__ldcg(&T.keyw[(size_t)h * W + w]) == X[w]);
                if (same) { __stcg(&T.val[h], vb); return; }
            }
            h = (h + 1) & T.mask;
        }
    }
}

// ---- hyper-inverse-Wishart clique score ------------------------------------------------------
// ggm:26-39 with wishart.py:17-37 folded: score(X) = gconst[k] - t1 ld(A_XX) + t2 ld(D_XX)
__device__ __forceinline__ double ggm_score_from(const PdgParams &P, int k, double ldA, double ldD)
{
    if (k == 0) return 0.0;
    const double t1 = (P.delta + P.nobs + (double)k - 1.0) * 0.5;
    const double t2 = (P.delta + (double)k - 1.0) * 0.5;
    return P.gconst[k] - t1 * ldA + t2 * ldD;
}

// bytes of the per-warp scratch region: the packed 32x32 triangle (GGM) or the log-linear count
// table (`cap` counters + 48 words of column info, pdg_loglin.cuh)
__host__ __device__ inline int tri_bytes_for(int ll_cap) { return ll_cap > 0 ? (ll_cap + 48) * 4 : 528 * 8; }

// per-warp scratch: packed 32x32 triangle + index list
struct HelpBox;
struct WarpScratch {
    HelpBox *help;            // log-linear MH kernel: the CTA's scan-sharing mailbox (pdg_loglin.cuh), else null
    int ktri;                 // largest clique the shared triangle tri32 holds (32 unless the MH kernel has room to spare)
    double *tri32;            // 528 doubles, or the log-linear count table (shared)
    short *idx;               // p + 1 entries (shared)
    double *big;              // global scratch for k > 32 (may be null)
    int kbig;
};

// idx[rank] = g for every set bit g of the key, ascending (rank by popcount of lower bits)
template <int WMAX>
__device__ __forceinline__ int build_index_sorted(const u64 (&X)[WMAX], int W, int lane, short *idx)
{
    int base = 0;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
        if (w < W) {
            const u64 x = X[w];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int b = half * 32 + lane;
                if ((x >> b) & 1ull) idx[base + __popcll(x & ((1ull << b) - 1ull))] = (short)(w * 64 + b);
            }
            base += __popcll(x);
        }
    }
    __syncwarp();
    return base;
}

// gathers the packed lower triangle of mat[idx, idx] and factorises it (LDL^T, see pdg_device.cuh)
__device__ __forceinline__ bool gather_ldl(const double *__restrict__ mat, int p, double *a, const short *idx,
                                           int lane, int k, double &logdet)
{
    if (k <= 32) {
        // lane <-> column: one coalescing-free but INDEPENDENT load per lane and row (the row index
        // comes by shuffle, so no shared-memory read sits between the loads and they pipeline)
        const int mycol = (lane < k) ? (int)idx[lane] : 0;
#pragma unroll 4
        for (int i = 0; i < k; ++i) {
            const int ri = __shfl_sync(PDG_FULL, mycol, i);
            if (lane <= i) a[tri(i, lane)] = __ldg(mat + (size_t)ri * p + mycol);
        }
    } else {
        // more than 32 vertices: lanes sweep the columns of one row after the other, so the loads of a
        // row -- and, unrolled, of four rows -- are independent (a lane per ROW walked its columns one
        // dependent L2 round trip at a time)
#pragma unroll 4
        for (int i = 0; i < k; ++i) {
            const double *row = mat + (size_t)idx[i] * p;
            double *dst = a + tri(i, 0);
            for (int l = lane; l <= i; l += 32) dst[l] = __ldg(row + idx[l]);
        }
        __syncwarp();
    }
    double sS, ldv, lsc;
    return ldl_logdets(a, lane, k, 0, k, sS, logdet, ldv, lsc);
}

// canonical score of one complete set (ascending vertex order); whole warp cooperates.
// Returns 0 or an ERR_* bit.
template <int WMAX>
__device__ __forceinline__ int ggm_set_score(const PdgParams &P, const WarpScratch &ws, const u64 (&X)[WMAX], int lane,
                                             double &score)
{
    const int k = build_index_sorted<WMAX>(X, P.W, lane, ws.idx);
    if (k == 0) { score = 0.0; return 0; }
    double *a = ws.tri32;
    if (k > ws.ktri) {
        if (!ws.big || k > ws.kbig) return ERR_BIGCLIQUE;
        a = ws.big;
    }
    double ldA, ldD = 0.0;
    if (!gather_ldl(P.A, P.p, a, ws.idx, lane, k, ldA)) return ERR_NUMERIC;
    if (P.dmode == 1) {
        double acc = 0.0;
        for (int j = lane; j < k; j += 32) acc += P.logDdiag[ws.idx[j]];
        ldD = warp_sum(acc);
    } else if (P.dmode == 2) {
        if (!gather_ldl(P.Dm, P.p, a, ws.idx, lane, k, ldD)) return ERR_NUMERIC;
    }
    score = ggm_score_from(P, k, ldA, ldD);
    return 0;
}

// ---- lane-private scorer for small cliques ---------------------------------------------------
// Cliques met by the sampler are small (mean 4-5 vertices), so the batched scorer gives every
// LANE its own complete set and factorises it in REGISTERS: the packed lower triangle of
// A[X, X] is gathered with fully unrolled, statically indexed code and reduced by a right-looking
// LDL^T in ascending vertex order.  All lanes of the warp run the same straight-line code for
// K = 4, 8 or 12, whichever covers the warp-wide largest set; smaller sets are padded with identity rows,
// which leaves every operation on the real block -- and therefore the result bit pattern --
// unchanged, so score(X) stays a pure function of X.  32 scores per warp pass, no shared memory.
#define PDG_KT 12                       // largest clique of the lane-private path

#define PDG_KTRI (PDG_KT * (PDG_KT + 1) / 2)
struct LaneScratch {
    int unused;
    double *a;                          // [PDG_KTRI][32] lane-interleaved triangles (PDG_LANE_SMEM build)
    unsigned short *idx;                // [PDG_KT][32]
};

// compact alternative (-DPDG_LANE_SMEM): the same ascending-order LDL^T with rolled loops over a
// lane-private triangle in shared memory; ~10x fewer instructions of CODE than the unrolled
// register variants (the MH kernel is instruction-fetch bound), more instructions EXECUTED.
template <int WMAX>
__device__ __forceinline__ int ggm_lane_score_smem(const PdgParams &P, const LaneScratch &ls, int lane, const u64 (&X)[WMAX],
                                                   int k, double &score)
{
    if (k == 0) { score = 0.0; return 0; }
    double *a = ls.a + lane;
    unsigned short *idx = ls.idx + lane;
    {
        int t = 0;
#pragma unroll
        for (int w = 0; w < WMAX; ++w) {
            u64 x = X[w];
            while (x) { idx[t * 32] = (unsigned short)(w * 64 + __ffsll((long long)x) - 1); x &= x - 1; ++t; }
        }
    }
    double ld[2] = {0.0, 0.0};
    bool ok = true;
    const int npass = (P.dmode == 2) ? 2 : 1, p = P.p;
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
        const double *__restrict__ mat = pass ? P.Dm : P.A;
        for (int i = 0, e = 0; i < k; ++i) {
            const double *row = mat + (size_t)idx[i * 32] * p;
            for (int l = 0; l <= i; ++l, ++e) a[e * 32] = __ldg(row + idx[l * 32]);
        }
        double prod = 1.0, lacc = 0.0;
        for (int j = 0; j < k; ++j) {
            const int jj = (j * (j + 1)) >> 1;
            const double d = a[(jj + j) * 32];
            ok = ok && (d > 0.0);
            prod *= d;
            if (prod > 1e150 || prod < 1e-150) { lacc += log(prod); prod = 1.0; }
            const double inv = 1.0 / d;
            for (int i = j + 1; i < k; ++i) {
                double *ri = a + ((i * (i + 1)) >> 1) * 32;
                const double f = ri[j * 32] * inv;
                for (int l = j + 1; l <= i; ++l) ri[l * 32] -= f * a[(((l * (l + 1)) >> 1) + j) * 32];
            }
        }
        ld[pass] = lacc + log(prod);
    }
    if (P.dmode == 1) for (int j = 0; j < k; ++j) ld[1] += P.logDdiag[idx[j * 32]];
    score = ggm_score_from(P, k, ld[0], ld[1]);
    return ok ? 0 : ERR_NUMERIC;
}

// One inlined copy per K: the (rare) general-D case reuses the same code through a rolled
// two-pass loop, which keeps the hot loop small (the MH kernel is instruction-cache sensitive:
// 32 KB L1.5 I-cache, several warps per SM at different program counters) and every array in
// registers.
#ifdef PDG_PHASE_TIMERS
static __device__ unsigned long long g_pt[8];
#define LPT(i) { const long long n_ = clock64(); lp[i] += (unsigned long long)(n_ - l0); l0 = n_; }
#else
#define LPT(i)
#endif

#ifndef PDG_NO_K4
#define PDG_NO_K4 0
#endif

template <int K, int WMAX>
__device__ __forceinline__ int ggm_lane_score_k(const PdgParams &P, const LaneScratch &ls, int lane, const u64 (&X)[WMAX], int k,
                                                double &score)
{
#ifdef PDG_PHASE_TIMERS
    long long l0 = clock64(); unsigned long long lp[5] = {0, 0, 0, 0, 0};
#endif
    int idx[K];
    if (WMAX == 1) {
        u64 y = X[0];
#pragma unroll
        for (int t = 0; t < K; ++t) { idx[t] = y ? __ffsll((long long)y) - 1 : 0; y &= y - 1; }
    } else {
        // multi-word keys: walk the words once with a dynamic counter into a lane-private column of
        // shared memory (ls.idx[t * 32 + lane]), then read it back with static indices
        unsigned short *stk = ls.idx + lane;
        int t = 0;
#pragma unroll
        for (int w = 0; w < WMAX; ++w) {
            u64 y = X[w];
            while (y && t < K) { stk[t * 32] = (unsigned short)(w * 64 + __ffsll((long long)y) - 1); ++t; y &= y - 1; }
        }
#pragma unroll
        for (int q = 0; q < K; ++q) idx[q] = (q < t) ? (int)stk[q * 32] : 0;
    }
    LPT(0)
    double ld[2] = {0.0, 0.0};
    bool ok = true;
    const int npass = (P.dmode == 2) ? 2 : 1;
    const int p = P.p;
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
        const double *__restrict__ mat = pass ? P.Dm : P.A;
        double a[K * (K + 1) / 2];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const double *row = mat + (size_t)idx[i] * p;
#pragma unroll
            for (int l = 0; l <= i; ++l) {
                double v = (i == l) ? 1.0 : 0.0;
                if (i < k) v = __ldg(row + idx[l]);
                a[i * (i + 1) / 2 + l] = v;
            }
        }
        LPT(1)
        double prod = 1.0, piv[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double d = a[j * (j + 1) / 2 + j];
            piv[j] = d;
            ok = ok && (d > 0.0);
            prod *= d;
            const double inv = 1.0 / d;
#pragma unroll
            for (int i = j + 1; i < K; ++i) {
                const double f = a[i * (i + 1) / 2 + j] * inv;
#pragma unroll
                for (int l = j + 1; l <= i; ++l) a[i * (i + 1) / 2 + l] -= f * a[l * (l + 1) / 2 + j];
            }
        }
        LPT(2)
        double v;
        if (prod > 1e-280 && prod < 1e280) v = log(prod);
        else {                               // product left the double range: sum the logs instead
            v = 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) v += log(piv[j]);
        }
        ld[pass] = v;
        LPT(3)
    }
    if (P.dmode == 1) {
#pragma unroll
        for (int j = 0; j < K; ++j) if (j < k) ld[1] += P.logDdiag[idx[j]];
    }
    score = ggm_score_from(P, k, ld[0], ld[1]);
    LPT(4)
#ifdef PDG_PHASE_TIMERS
    if ((threadIdx.x & 31) == 0) { for (int q_ = 0; q_ < 5; ++q_) atomicAdd(&g_pt[q_], lp[q_]); atomicAdd(&g_pt[5 + (K == 8)], 1ull); }
#endif
    return (ok || k == 0) ? 0 : ERR_NUMERIC;
}

// scores of the lanes' own sets; k = |X| for lanes that need one (<= PDG_KT), 0 for the others.
// Warp-uniform dispatch on the largest k.
template <int WMAX>
__device__ __forceinline__ int ggm_lane_scores(const PdgParams &P, const LaneScratch &ls, int lane, const u64 (&X)[WMAX],
                                               int k, double &score)
{
#ifdef PDG_LANE_SMEM
    return ggm_lane_score_smem<WMAX>(P, ls, lane, X, k, score);
#endif
    const int kmax = (int)__reduce_max_sync(PDG_FULL, (unsigned)k);
    if (kmax == 0) { score = 0.0; return 0; }
    if (kmax <= 4 && !PDG_NO_K4) return ggm_lane_score_k<4, WMAX>(P, ls, lane, X, k, score);
    if (kmax <= 8) return ggm_lane_score_k<8, WMAX>(P, ls, lane, X, k, score);
    return ggm_lane_score_k<12, WMAX>(P, ls, lane, X, k, score);
}

// ---- lane-private scorer, rolled (the one the MH kernel and the batch scorer use) ---------------
// Same arithmetic as ggm_lane_score_k -- right-looking LDL^T in ascending vertex order, log of the
// product of the pivots -- hence the same bits for every set, but with ROLLED loops over a
// lane-private triangle in shared memory: ~150 instructions of code whatever the set size, and the
// number executed follows the actual |X| (the unrolled register variants run K = 4 / 8 / 12 padded,
// 800 / 1800 / 3100 instructions each, and need ~200 registers: with them inlined the MH loop neither
// fits the instruction cache nor keeps its own state in registers).  The lanes that need a score are
// served PDG_SL at a time; element e of slot s sits at tri[e * PDG_SL + s] (the warp's 32-row scratch
// triangle: 528 doubles >= PDG_SL * PDG_KTRI).
#define PDG_SL 6
template <int WMAX>
__device__ __forceinline__ int ggm_lane_scores_rolled(const PdgParams &P, double *tri, unsigned short *lidx, int lane,
                                                      const u64 (&X)[WMAX], int k, double &score)
{
    const u32 m = __ballot_sync(PDG_FULL, k > 0);
    const int rank = __popc(m & ((1u << lane) - 1u)), total = __popc(m);
    const int p = P.p, npass = (P.dmode == 2) ? 2 : 1;
    int err = 0;
#pragma unroll 1
    for (int r0 = 0; r0 < total; r0 += PDG_SL) {
        if (k > 0 && rank >= r0 && rank < r0 + PDG_SL) {
            double *a = tri + (rank - r0);
            unsigned short *idx = lidx + (rank - r0);
            {
                int t = 0;
#pragma unroll
                for (int w = 0; w < WMAX; ++w) {
                    u64 x = X[w];
                    while (x) { idx[t * PDG_SL] = (unsigned short)(w * 64 + __ffsll((long long)x) - 1); x &= x - 1; ++t; }
                }
            }
            double ld[2] = {0.0, 0.0};
            bool ok = true;
#pragma unroll 1
            for (int pass = 0; pass < npass; ++pass) {
                const double *__restrict__ mat = pass ? P.Dm : P.A;
                {
                    double *dst = a;
#pragma unroll 1
                    for (int i = 0; i < k; ++i) {
                        const double *row = mat + (size_t)idx[i * PDG_SL] * p;
                        const unsigned short *il = idx;
#pragma unroll 2
                        for (int l = 0; l <= i; ++l, dst += PDG_SL, il += PDG_SL) *dst = __ldg(row + *il);
                    }
                }
                double prod = 1.0;
                // (running pointers instead of triangle index arithmetic: dj -> (j, j), rij -> (i, j), clj -> (l, j))
                double *dj = a;
#pragma unroll 1
                for (int j = 0; j < k; ++j) {
                    const double d = *dj;
                    ok = ok && (d > 0.0);
                    prod *= d;
                    const double inv = 1.0 / d;
                    double *rij = dj + (j + 1) * PDG_SL;                     // (j + 1, j)
#pragma unroll 1
                    for (int i = j + 1; i < k; ++i) {
                        const double f = *rij * inv;
                        double *ril = rij + PDG_SL;                              // (i, j + 1)
                        const double *clj = dj + (j + 1) * PDG_SL;               // (j + 1, j)
#pragma unroll 2
                        for (int l = j + 1; l <= i; ++l) { *ril -= f * *clj; ril += PDG_SL; clj += (l + 1) * PDG_SL; }
                        rij += (i + 1) * PDG_SL;                                 // (i + 1, j)
                    }
                    dj += (j + 2) * PDG_SL;                                  // (j + 1, j + 1)
                }
                double v;
                if (prod > 1e-280 && prod < 1e280) v = log(prod);
                else {                           // product left the double range: sum the logs of the pivots instead
                    v = 0.0;
                    for (int j = 0; j < k; ++j) v += log(a[(((j * (j + 1)) >> 1) + j) * PDG_SL]);
                }
                ld[pass] = v;
            }
            if (P.dmode == 1) for (int j = 0; j < k; ++j) ld[1] += P.logDdiag[idx[j * PDG_SL]];
            score = ggm_score_from(P, k, ld[0], ld[1]);
            if (!ok) err = ERR_NUMERIC;
        }
        __syncwarp();
    }
    return err;
}

// ---- small sets in registers: eight lanes per set, four sets per pass --------------------------------
// The scorer the MH kernel and the batch scorer use for sets of at most 8 vertices.  Lane r of a group of
// eight holds ROW r of its set's lower triangle in registers (statically indexed: the loops are fully
// unrolled); a pivot's diagonal and column reach the other rows by shuffles.  The arithmetic is that of
// ggm_lane_scores_rolled -- right-looking LDL^T in ascending vertex order, f = a_ij / d_j as a multiply by
// the reciprocal, a_il -= f a_lj, log of the product of the pivots -- so the bits of a score do not depend on
// which of the two computed it; rows beyond |X| are identity rows and leave every operation on the real
// block unchanged.  ~330 instructions per pass of four sets with ~1.2 k cycles of dependent latency, where
// the rolled lane-private code runs ~1500 instructions per 7-vertex set one shared-memory round trip at a time.
template <int WMAX>
__device__ __forceinline__ int ggm_group8_scores(const PdgParams &P, int lane, const u64 (&X)[WMAX], int k, double &score)
{
    const u32 m = __ballot_sync(PDG_FULL, k > 0);
    if (m == 0u) return 0;
    const int rank = __popc(m & ((1u << lane) - 1u)), total = __popc(m);
    const int g = lane >> 3, r = lane & 7, gbase = lane & ~7;
    const int npass = (P.dmode == 2) ? 2 : 1, p = P.p, W = P.W;
    int err = 0;
#pragma unroll 1
    for (int r0 = 0; r0 < total; r0 += 4) {
        const bool have = r0 + g < total;
        const int src = have ? (int)__fns(m, 0u, r0 + g + 1) : 0;        // the lane whose set this group factorises
        u64 Y[WMAX];
#pragma unroll
        for (int w = 0; w < WMAX; ++w) Y[w] = __shfl_sync(PDG_FULL, X[w], src);
        int kk = __shfl_sync(PDG_FULL, k, src);
        if (!have) kk = 0;
        // this lane's vertex: the r-th member of the set
        int idx = 0;
        {
            int need = r;
            bool found = false;
#pragma unroll
            for (int w = 0; w < WMAX; ++w) {
                if (w < W && !found) {
                    const int c = __popcll(Y[w]);
                    if (need < c) { idx = w * 64 + select64(Y[w], need); found = true; }
                    else need -= c;
                }
            }
        }
        double ld0 = 0.0, ld1 = 0.0;
        bool ok = true;
#pragma unroll 1
        for (int pass = 0; pass < npass; ++pass) {
            const double *__restrict__ mat = pass ? P.Dm : P.A;
            double a[8], piv[8];
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                const int il = __shfl_sync(PDG_FULL, idx, gbase + l);
                double v = (l == r) ? 1.0 : 0.0;
                if (r < kk && l <= r) v = __ldg(mat + (size_t)idx * p + il);
                a[l] = v;
            }
            double prod = 1.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double d = __shfl_sync(PDG_FULL, a[j], gbase + j);
                piv[j] = d;
                ok = ok && (d > 0.0);
                prod *= d;
                const double inv = 1.0 / d;
                const double f = a[j] * inv;
#pragma unroll
                for (int l = j + 1; l < 8; ++l) {
                    const double c = __shfl_sync(PDG_FULL, a[j], gbase + l);
                    if (r >= l) a[l] -= f * c;
                }
            }
            double v;
            if (prod > 1e-280 && prod < 1e280) v = log(prod);
            else {                               // product left the double range: sum the logs of the pivots instead
                v = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) if (j < kk) v += log(piv[j]);
            }
            if (pass) ld1 = v; else ld0 = v;
        }
        if (P.dmode == 1) {
            const double mine = (r < kk) ? P.logDdiag[idx] : 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { const double t = __shfl_sync(PDG_FULL, mine, gbase + j); if (j < kk) ld1 += t; }
        }
        const double sc = ggm_score_from(P, kk, ld0, ld1);
        const int bad = (!ok && kk > 0) ? 1 : 0;
        // hand the scores back to the lanes that asked
        const int mygrp = rank - r0;
        const bool mine_now = k > 0 && mygrp >= 0 && mygrp < 4;
        const double got = __shfl_sync(PDG_FULL, sc, mine_now ? 8 * mygrp : 0);
        const int gotbad = __shfl_sync(PDG_FULL, bad, mine_now ? 8 * mygrp : 0);
        if (mine_now) { score = got; if (gotbad) err = ERR_NUMERIC; }
    }
    return err;
}

template <int WMAX>
__device__ __forceinline__ void load_row(u64 (&X)[WMAX], const u64 *row, int W)
{
#pragma unroll
    for (int w = 0; w < WMAX; ++w) X[w] = (w < W) ? row[w] : 0ull;
}

// ---- shared-memory layout of one warp of the MH kernel (pdg_run.cuh) --------------------------
struct RunSmem {
    int off_tri, off_idx, off_sub, off_leaf, off_bnd, off_partner, off_mvU, off_mvUadj, off_csr_off, off_csr_adj,
        off_M, off_N, off_T, off_lidx, off_tmpU, off_tmpA, off_cnt, off_pair, total;
};

__host__ __device__ inline RunSmem run_smem_layout(int p, int W, bool smem_state, int tri_bytes)
{
    RunSmem s;
    int o = 0;
    const int np32 = 2 * W + 2;
    const int a2 = ((p + 2) * 2 + 15) / 16 * 16;
#ifdef PDG_LANE_SMEM
    s.off_tri = o;      o += PDG_KTRI * 32 * 8;          // lane-private triangles (alias the 32x32 warp triangle)
    s.off_lidx = o;     o += PDG_KT * 32 * 2;
#else
    s.off_tri = o;      o += tri_bytes;
    s.off_lidx = o;     o += PDG_KT * 32 * 2;            // lane-private index columns (multi-word keys)
#endif
    s.off_sub = o;      o += W * 8;
    s.off_M = o;        o += smem_state ? p * W * 8 : 0;
    s.off_N = o;        o += smem_state ? p * W * 8 : 0;
    s.off_T = o;        o += smem_state ? p * W * 8 : 0;     // adjacency bit-matrix of the latent tree
    s.off_leaf = o;     o += (np32 * 4 + 15) / 16 * 16;
    s.off_bnd = o;      o += (np32 * 4 + 15) / 16 * 16;
    s.off_idx = o;      o += a2;
    s.off_partner = o;  o += a2;
    s.off_mvU = o;      o += a2;
    s.off_mvUadj = o;   o += a2;
    s.off_csr_off = o;  o += a2;
    s.off_csr_adj = o;  o += 2 * a2;
    s.off_tmpU = o;     o += smem_state ? 0 : a2;      // unsorted boundary of the neighbourhood-driven scan
    s.off_tmpA = o;     o += smem_state ? 0 : a2;
    s.off_cnt = o;      o += 16;
    s.off_pair = o;     o += 96;                           // two PairBox mailboxes (pdg_run.cuh, SPEC kernels)
    s.total = (o + 15) / 16 * 16;
    return s;
}

