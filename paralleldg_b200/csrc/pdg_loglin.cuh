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

template <int WMAX>
__device__ __forceinline__ int loglin_set_score(const PdgParams &P, const WarpScratch &ws, int slot,
                                                const u64 (&X)[WMAX], int lane, double &score)
{
    const int k = build_index_sorted<WMAX>(X, P.W, lane, ws.idx);
    if (k == 0) { score = 0.0; return 0; }
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
    if (!big && Ki <= (u64)P.hist_cells) {
        unsigned int *hist = P.hist + (size_t)slot * P.hist_cells;      // all zero on entry
        if ((n & 3) == 0) {
            for (long long r = (long long)lane * 4; r < n; r += 128) {
                unsigned int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                for (int j = 0; j < k; ++j) {
                    const int g = ws.idx[j];
                    const unsigned int Lj = (unsigned int)P.levels[g];
                    const uchar4 x = __ldg((const uchar4 *)(D + (size_t)g * n + r));
                    c0 = c0 * Lj + x.x; c1 = c1 * Lj + x.y; c2 = c2 * Lj + x.z; c3 = c3 * Lj + x.w;
                }
                atomicAdd(&hist[c0], 1u); atomicAdd(&hist[c1], 1u); atomicAdd(&hist[c2], 1u); atomicAdd(&hist[c3], 1u);
            }
        } else {
            for (long long r = lane; r < n; r += 32) {
                unsigned int c = 0;
                for (int j = 0; j < k; ++j) { const int g = ws.idx[j]; c = c * (unsigned int)P.levels[g] + D[(size_t)g * n + r]; }
                atomicAdd(&hist[c], 1u);
            }
        }
        __threadfence();
        __syncwarp();
        for (u64 c = lane; c < Ki; c += 32) {
            const unsigned int cnt = __ldcg(&hist[c]);
            if (cnt) {
                acc += LG ? LG[cnt] : (lgamma((double)cnt + alpha) - lga);
                hist[c] = 0u;
            }
        }
        __threadfence();
        __syncwarp();
        score = warp_sum(acc) - constK;
        return 0;
    }
    if (big) return ERR_BIGCLIQUE;            // cell codes beyond 62 bits
    // ---- sparse path: hashed cells, then a count-of-counts pass for an order-free sum ----------
    const LoglinSparse S = P.sp;
    if (S.nsp == 0) return ERR_BIGCLIQUE;
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
    return 0;
}
