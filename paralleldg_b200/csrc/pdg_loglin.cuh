// Log-linear (Dirichlet / contingency table) clique scorer, one warp per complete set.
// Restates discrete_dec_log_linear.py:12-70 + dirichlet.py:63-85 + auxiliary_functions.py:134-151:
//   score(X) = sum over occupied cells [lgamma(n_cell + a) - lgamma(a)] - [lgamma(K a + N) - lgamma(K a)]
// with K = prod of levels over X and a = cell_alpha * prod(levels outside X) / prod(all levels).
// The scan reads the k level-code columns of X (column-major uint8, coalesced, 16 rows per lane
// and load) once and counts cells with warp-wide atomics: into a dense per-warp table when
// K <= hist_cells (<= 65536), else into a hashed table from a small lock-protected pool.  The lgamma terms
// come from host-built tables (bit-identical to the reference's math.lgamma) when available.
#pragma once
#include "pdg_kernels.cuh"

// Everything the scan needs, by value: the function below is deliberately OUT OF LINE (its own
// register allocation -- the unrolled loads would otherwise spill the MH kernel -- and no address
// of the kernel parameter block is taken).
struct LoglinArgs {
    const unsigned char *data; const int *levels; long long nrows, ld; int zero_col; double cell_alpha;
    int n_ktab; const double *ktab, *alphatab, *lgtab;
    unsigned int *hist; int hist_cells; LoglinSparse sp; unsigned long long *work; int W;
};
template <int WMAX> struct KeyW { u64 w[WMAX]; };
struct ScoreRes { double score; int err; };

// Dense counting pass.  The table is column-major uint8 with a 16-byte-aligned column stride `ld`,
// zero padding below row n, one all-zero column behind the last real one and 4 KB of slack, so the
// pass needs no column or load predicates: every lane reads 16 rows per load (uint4), a warp pass
// covers 2048 rows of four columns (16 loads in flight per lane), missing columns of the last
// chunk point at the zero column with multiplier 1.  Cell codes are built with PACKED arithmetic:
// four rows per 32-bit multiply-add while codes fit a byte (K <= 256), two per multiply-add above.
// The loads of the next chunk are issued before the current pass is counted.
// Counting: shared-memory atomics into R interleaved replicas of the table (R * K <= 1024; lanes
// l and l + R share one, so same-cell collisions inside a warp mostly disappear), or this warp's
// table in global memory above 1024 cells.
template <bool WIDE, bool GLB>
__device__ __forceinline__ void loglin_count_dense(const unsigned char *D, long long ld, long long n, int k,
                                                   const unsigned int *colinfo, unsigned int *tab, int rshift, int lane)
{
    constexpr int NC = WIDE ? 2 : 1;
    uint4 x[4][4];
    auto issue = [&](long long base, int j0) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const unsigned char *col = D + (size_t)colinfo[16 + j0 + jj] * (size_t)ld + base + lane * 16;
#pragma unroll
            for (int u = 0; u < 4; ++u) x[jj][u] = __ldg((const uint4 *)(col + 512 * u));
        }
    };
    auto count1 = [&](unsigned int code) {
        if (GLB) atomicAdd(&tab[code], 1u); else atomicAdd(&tab[code << rshift], 1u);
    };
    issue(0, 0);
    for (long long base = 0; base < n; base += 2048) {
        unsigned int c[4][4][NC];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int h = 0; h < NC; ++h) c[u][q][h] = 0u;
        for (int j0 = 0; j0 < k; j0 += 4) {
            unsigned int L4[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) L4[jj] = colinfo[j0 + jj];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned int w[4] = {x[jj][u].x, x[jj][u].y, x[jj][u].z, x[jj][u].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (WIDE) {
                            c[u][q][0] = c[u][q][0] * L4[jj] + __byte_perm(w[q], 0u, 0x4240);   // rows 4q, 4q+2
                            c[u][q][NC - 1] = c[u][q][NC - 1] * L4[jj] + __byte_perm(w[q], 0u, 0x4341);   // rows 4q+1, 4q+3
                        } else {
                            c[u][q][0] = c[u][q][0] * L4[jj] + w[q];
                        }
                    }
                }
            }
            const int nj0 = (j0 + 4 < k) ? j0 + 4 : 0;
            const long long nbase = nj0 ? base : base + 2048;
            if (nbase < n) issue(nbase, nj0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long r = base + 512 * u + lane * 16;
            if (r < ld) {      // ld is a multiple of 16: the padding rows land in cell 0 and are subtracted by the caller
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (WIDE) {
                        count1(c[u][q][0] & 0xffffu); count1(c[u][q][NC - 1] & 0xffffu);
                        count1(c[u][q][0] >> 16); count1(c[u][q][NC - 1] >> 16);
                    } else {
                        count1(c[u][q][0] & 0xffu); count1((c[u][q][0] >> 8) & 0xffu);
                        count1((c[u][q][0] >> 16) & 0xffu); count1(c[u][q][0] >> 24);
                    }
                }
            }
        }
    }
}

template <int WMAX>
__device__ __forceinline__ ScoreRes loglin_scan(LoglinArgs P, WarpScratch ws, int slot, KeyW<WMAX> XK, int lane)
{
    ScoreRes res; res.score = 0.0; res.err = 0;
    double score = 0.0;
    const int k = build_index_sorted<WMAX>(XK.w, P.W, lane, ws.idx);
    if (k == 0) return res;
    const long long n = P.nrows;
    double Kd = 1.0;
    u64 Ki = 1ull;
    bool big = false;
    for (int j = 0; j < k; ++j) {
        const u64 Lj = (u64)P.levels[ws.idx[j]];
        Kd *= (double)Lj;
        if (Ki > (1ull << 62) / Lj) big = true; else Ki *= Lj;
    }
    int tj = -1;
    for (int j = 0; j < P.n_ktab; ++j) if (P.ktab[j] == Kd) tj = j;
    const double *LG = tj >= 0 ? P.lgtab + (size_t)tj * (size_t)(n + 2) : nullptr;
    const double alpha = tj >= 0 ? P.alphatab[tj] : P.cell_alpha / Kd;
    const double constK = tj >= 0 ? LG[n + 1] : lgamma(Kd * alpha + (double)n) - lgamma(Kd * alpha);
    const double lga = LG ? 0.0 : lgamma(alpha);
    const unsigned char *D = P.data;
    double acc = 0.0;
    if (!big && Ki <= (u64)P.hist_cells && k <= 16) {
        // shared layout (the 528-double triangle buffer): 1024 counters, then 16 level counts and
        // 16 column ids (slots past k: multiplier 1, the zero column)
        unsigned int *sh = (unsigned int *)ws.tri32;
        unsigned int *colinfo = sh + 1024;
        if (lane < 16) {
            const int g = (lane < k) ? (int)ws.idx[lane] : P.zero_col;
            colinfo[lane] = (lane < k) ? (unsigned int)P.levels[g] : 1u;
            colinfo[16 + lane] = (unsigned int)g;
        }
        const bool glb = Ki > 1024ull;
        int rshift = 0;
        if (!glb) while (rshift < 5 && (Ki << (rshift + 1)) <= 1024ull) ++rshift;
        const int R = 1 << rshift;
        unsigned int *hist = P.hist + (size_t)slot * P.hist_cells;      // all zero on entry
        if (!glb) for (int c = lane; c < ((int)Ki << rshift); c += 32) sh[c] = 0u;
        __syncwarp();
        unsigned int *tab = glb ? hist : sh + (lane & (R - 1));
        if (glb) loglin_count_dense<true, true>(D, P.ld, n, k, colinfo, tab, 0, lane);
        else if (Ki > 256ull) loglin_count_dense<true, false>(D, P.ld, n, k, colinfo, tab, rshift, lane);
        else loglin_count_dense<false, false>(D, P.ld, n, k, colinfo, tab, rshift, lane);
        if (glb) __threadfence();
        __syncwarp();
        for (u64 c = lane; c < Ki; c += 32) {
            unsigned int cnt;
            if (!glb) { cnt = 0u; for (int q = 0; q < R; ++q) cnt += sh[((int)c << rshift) + ((q + lane) & (R - 1))]; }
            else { cnt = __ldcg(&hist[c]); if (cnt) hist[c] = 0u; }
            if (c == 0ull) cnt -= (unsigned int)(P.ld - n);      // zero padding rows of the last block
            if (cnt) acc += LG ? LG[cnt] : (lgamma((double)cnt + alpha) - lga);
        }
        if (glb) __threadfence();
        __syncwarp();
        score = warp_sum(acc) - constK;
        res.score = score;
        return res;
    }
    if (big) { res.err = ERR_BIGCLIQUE; return res; }            // cell codes beyond 62 bits
    if (lane == 0) atomicAdd(&P.work[(size_t)slot * 16 + 6], 1ull);       // sets that needed the hashed-cell pool
    // ---- sparse path: hashed cells, then a count-of-counts pass for an order-free sum ----------
    const LoglinSparse S = P.sp;
    if (S.nsp == 0) { res.err = ERR_BIGCLIQUE; return res; }
    int s = 0;
    if (lane == 0) {
        s = slot % S.nsp;
        while (atomicCAS(&S.lock[s], 0, 1) != 0) { s = (s + 1) % S.nsp; __nanosleep(200); }
    }
    s = __shfl_sync(PDG_FULL, s, 0);
    u64 *hk = S.hkeys + (size_t)s * S.hslots;
    unsigned int *hc = S.hcnt + (size_t)s * S.hslots, *cc = S.cc + (size_t)s * (size_t)(n + 2);
    const u32 hmask = (u32)S.hslots - 1u;
    for (long long r = lane; r < n; r += 32) {
        u64 c = 0ull;
        for (int j = 0; j < k; ++j) { const int g = ws.idx[j]; c = c * (u64)P.levels[g] + D[(size_t)g * (size_t)P.ld + r]; }
        const u64 key = c + 1ull;
        u32 h = (u32)mix64(key) & hmask;
        for (;;) {
            const u64 old = atomicCAS(&hk[h], 0ull, key);
            if (old == 0ull || old == key) { atomicAdd(&hc[h], 1u); break; }
            h = (h + 1) & hmask;
        }
    }
    __threadfence();
    __syncwarp();
    unsigned int maxc = 0;
    for (int h = lane; h < S.hslots; h += 32) {
        if (__ldcg(&hk[h]) != 0ull) {
            const unsigned int c = __ldcg(&hc[h]);
            atomicAdd(&cc[c], 1u);
            maxc = c > maxc ? c : maxc;
            hk[h] = 0ull; hc[h] = 0u;
        }
    }
    maxc = __reduce_max_sync(PDG_FULL, maxc);
    __threadfence();
    __syncwarp();
    for (unsigned int c = 1 + lane; c <= maxc; c += 32) {
        const unsigned int m = __ldcg(&cc[c]);
        if (m) {
            acc += (double)m * (LG ? LG[c] : (lgamma((double)c + alpha) - lga));
            cc[c] = 0u;
        }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(&S.lock[s], 0);
    score = warp_sum(acc) - constK;
    res.score = score;
    return res;
}


// inlined shim used by the kernels
template <int WMAX>
__device__ __forceinline__ int loglin_set_score(const PdgParams &P, const WarpScratch &ws, int slot,
                                                const u64 (&X)[WMAX], int lane, double &score)
{
    LoglinArgs A;
    A.data = P.data; A.levels = P.levels; A.nrows = P.nrows; A.ld = P.data_ld; A.zero_col = P.p; A.cell_alpha = P.cell_alpha;
    A.n_ktab = P.n_ktab; A.ktab = P.ktab; A.alphatab = P.alphatab; A.lgtab = P.lgtab;
    A.hist = P.hist; A.hist_cells = P.hist_cells; A.sp = P.sp; A.work = P.work; A.W = P.W;
    KeyW<WMAX> K;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) K.w[w] = X[w];
    const ScoreRes r = loglin_scan<WMAX>(A, ws, slot, K, lane);
    score = r.score;
    return r.err;
}
