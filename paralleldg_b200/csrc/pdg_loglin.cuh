// Log-linear (Dirichlet / contingency table) clique scorer, one warp per complete set.
// Restates discrete_dec_log_linear.py:12-70 + dirichlet.py:63-85 + auxiliary_functions.py:134-151:
//   score(X) = sum over occupied cells [lgamma(n_cell + a) - lgamma(a)] - [lgamma(K a + N) - lgamma(K a)]
// with K = prod of levels over X and a = cell_alpha * prod(levels outside X) / prod(all levels).
// The scan reads the k level-code columns of X (column-major uint8, coalesced, 4 rows per lane
// and load) once and counts cells with warp-wide atomics: into a dense per-warp table when
// K <= hist_cells, else into a hashed table from a small lock-protected pool.  The lgamma terms
// come from host-built tables (bit-identical to the reference's math.lgamma) when available.
#pragma once
#include "pdg_kernels.cuh"

// Everything the scan needs, by value: the function below is deliberately OUT OF LINE (its own
// register allocation -- the unrolled loads would otherwise spill the MH kernel -- and no address
// of the kernel parameter block is taken).
struct LoglinArgs {
    const unsigned char *data; const int *levels; long long nrows; double cell_alpha;
    int n_ktab; const double *ktab, *alphatab, *lgtab;
    unsigned int *hist; int hist_cells; LoglinSparse sp; unsigned long long *work; int W;
};
template <int WMAX> struct KeyW { u64 w[WMAX]; };
struct ScoreRes { double score; int err; };

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
        // column ids and level counts in registers; every row block issues ALL its column loads
        // before the first use (static unroll, predicated), so a block costs one L2 round trip
        // instead of k dependent ones
        int gc[16];
        unsigned int Lv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { gc[j] = (j < k) ? (int)ws.idx[j] : 0; Lv[j] = (j < k) ? (unsigned int)P.levels[gc[j]] : 1u; }
        // where the cells are counted: lane-private columns (<= 32 cells, no atomics), shared
        // atomics (<= 1024 cells), or this warp's table in global memory
        // (same-address global atomics serialise in L2 at ~100 ns each: a 64-cell table would cost
        // ~10 ms per scan, so small tables are counted in shared memory)
        const int mode = (Ki <= 32ull) ? 0 : (Ki <= 1024ull) ? 1 : 2;
        unsigned int *sh = (unsigned int *)ws.tri32;
        unsigned int *hist = P.hist + (size_t)slot * P.hist_cells;      // all zero on entry
        if (mode == 0) for (int c = lane; c < 32 * (int)Ki; c += 32) sh[c] = 0u;
        if (mode == 1) for (int c = lane; c < (int)Ki; c += 32) sh[c] = 0u;
        __syncwarp();
        const bool vec = (n & 3) == 0;
        auto count = [&](unsigned int c) {
            if (mode == 0) sh[c * 32 + lane] += 1u;
            else if (mode == 1) atomicAdd(&sh[c], 1u);
            else atomicAdd(&hist[c], 1u);
        };
        if (vec) {
            // 4 rows per lane and load (uchar4), EIGHT 128-row blocks in flight; the columns are taken
            // four at a time in a warp-uniform loop, so one L2 round trip covers 32 loads and only
            // ceil(k/4)*4 columns are ever touched.  Cell codes accumulate in registers.
            for (long long r0 = (long long)lane * 4; r0 < n; r0 += 1024) {
                unsigned int c[8][4];
#pragma unroll
                for (int u = 0; u < 8; ++u) { c[u][0] = 0u; c[u][1] = 0u; c[u][2] = 0u; c[u][3] = 0u; }
                for (int j0 = 0; j0 < k; j0 += 4) {
                    uchar4 x[8][4];
                    unsigned int L4[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = j0 + jj;
                        const int g = (j < k) ? (int)ws.idx[j] : (int)ws.idx[0];
                        L4[jj] = (j < k) ? (unsigned int)P.levels[g] : 1u;
                        const unsigned char *col = D + (size_t)g * n + r0;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            x[u][jj] = make_uchar4(0, 0, 0, 0);
                            if (j < k && r0 + 128 * u < n) x[u][jj] = __ldg((const uchar4 *)(col + 128 * u));
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            c[u][0] = c[u][0] * L4[jj] + x[u][jj].x; c[u][1] = c[u][1] * L4[jj] + x[u][jj].y;
                            c[u][2] = c[u][2] * L4[jj] + x[u][jj].z; c[u][3] = c[u][3] * L4[jj] + x[u][jj].w;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (r0 + 128 * u < n) { count(c[u][0]); count(c[u][1]); count(c[u][2]); count(c[u][3]); }
                }
            }
        } else {
            for (long long r = lane; r < n; r += 32) {
                unsigned char x[16];
                unsigned int c0 = 0u;
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < k) x[j] = __ldg(D + (size_t)gc[j] * n + r);
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < k) c0 = c0 * Lv[j] + x[j];
                count(c0);
            }
        }
        if (mode == 2) __threadfence();
        __syncwarp();
        for (u64 c = lane; c < Ki; c += 32) {
            unsigned int cnt;
            if (mode == 0) { cnt = 0u; for (int q = 0; q < 32; ++q) cnt += sh[(int)c * 32 + q]; }
            else if (mode == 1) cnt = sh[c];
            else { cnt = __ldcg(&hist[c]); if (cnt) hist[c] = 0u; }
            if (cnt) acc += LG ? LG[cnt] : (lgamma((double)cnt + alpha) - lga);
        }
        if (mode == 2) __threadfence();
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
        for (int j = 0; j < k; ++j) { const int g = ws.idx[j]; c = c * (u64)P.levels[g] + D[(size_t)g * n + r]; }
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
    A.data = P.data; A.levels = P.levels; A.nrows = P.nrows; A.cell_alpha = P.cell_alpha;
    A.n_ktab = P.n_ktab; A.ktab = P.ktab; A.alphatab = P.alphatab; A.lgtab = P.lgtab;
    A.hist = P.hist; A.hist_cells = P.hist_cells; A.sp = P.sp; A.work = P.work; A.W = P.W;
    KeyW<WMAX> K;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) K.w[w] = X[w];
    const ScoreRes r = loglin_scan<WMAX>(A, ws, slot, K, lane);
    score = r.score;
    return r.err;
}
