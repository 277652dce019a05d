// Log-linear (Dirichlet / contingency table) clique scorer, one warp per complete set.
// Restates discrete_dec_log_linear.py:12-70 + dirichlet.py:63-85 + auxiliary_functions.py:134-151:
//   score(X) = sum over occupied cells [lgamma(n_cell + a) - lgamma(a)] - [lgamma(K a + N) - lgamma(K a)]
// with K = prod of levels over X and a = cell_alpha * prod(levels outside X) / prod(all levels).
// The scan reads the k level-code columns of X (column-major uint8, coalesced, 16 rows per lane
// and load) once and counts cells with warp-wide atomics: into a dense per-warp table when
// K <= hist_cells (<= 65536), else into a hashed table from a small lock-protected pool.  The lgamma terms
// come from host-built tables (bit-identical to the reference's math.lgamma) when available.
#pragma once
#include <type_traits>

#include "pdg_kernels.cuh"

// Everything the scan needs, by value, so that no address of the kernel parameter block is taken
// (that would move the whole block to local memory).  The scan is inlined into its callers: as an
// out-of-line function its by-value arguments went through a 320-byte stack frame and the kernel
// kept 255 registers anyway.
struct LoglinArgs {
    const unsigned char *data; const int *levels; long long nrows, ld; int zero_col; double cell_alpha;
    int n_ktab; const double *ktab, *alphatab, *lgtab;
    unsigned int *hist; int hist_cells, cap; LoglinSparse sp; unsigned long long *work; int W;
};
template <int WMAX> struct KeyW { u64 w[WMAX]; };
#ifdef PDG_PHASE_TIMERS
static __device__ unsigned long long g_kh[2][32];
static __device__ unsigned long long g_sp[16];       // developer probe: cycles of the hashed path's stages
#define SPT(i) { const long long n_ = clock64(); if (lane == 0) atomicAdd(&g_sp[i], (unsigned long long)(n_ - sp_t)); sp_t = n_; }      // developer probe: scans and cycles by set size
#endif
#ifndef PDG_PHASE_TIMERS
#define SPT(i)
#endif
struct ScoreRes { double score; int err; };

// Dense counting pass.  The table is column-major uint8 with a 16-byte-aligned column stride `ld`,
// zero padding below row n, one all-zero column behind the last real one and 4 KB of slack, so the
// pass needs no column or load predicates: every lane reads 16 rows per load (uint4), a warp pass
// covers 2048 rows of four column SLOTS (16 loads in flight per lane); empty slots point at the zero
// column with multiplier 1.  Cell codes are built with PACKED arithmetic, four rows per 32-bit
// multiply-add:
//   MODE 0  K <= 256: one byte-packed accumulator per word;
//   MODE 1  the columns split into two groups of <= 256 cells each (slots [0, slotsA) and the rest):
//           two byte-packed accumulators, code = codeA * KB + codeB formed when counting;
//   MODE 2  anything else up to 65536 cells: two 16-bit-packed accumulators per word.
// The loads of the next chunk are issued before the current pass is counted.
// Counting: shared-memory reductions into R interleaved replicas of the table (R * K <= cap; lanes
// l and l + R share one, so same-cell collisions inside a warp mostly disappear), or this warp's
// table in global memory above `cap` cells.
// A scan offered to the idle warps of the CTA (warps whose chain has ended): the rows are dealt out in
// chunks through a ticket that carries the request number, every participant counts into its own
// table (same layout) and adds it to the requester's, and the requester waits until all chunks are
// accounted for.  Counts are integers, so the table -- and the score -- do not depend on who counted what.
#define PDG_HELP_ROWS 4096          // rows per chunk (a multiple of 512 * NB)
struct HelpBox {
    int lock;                       // 1 while a request is open
    int ticket;                     // (request number << 16) | next chunk
    int n_chunks, done;             // chunks of the open request / chunks counted AND merged
    int finished;                   // chains of the CTA that have ended
    int mode, glb, nslots, slotsA, rshift;
    unsigned int mulA, mulB, cells; // cells: counters of the (replicated) shared table
    unsigned int *req_table;        // the requester's shared table
    unsigned int *hist;             // the requester's table in global memory (glb)
    long long n, ld;
    unsigned int colinfo[32];       // multipliers and column ids by slot
};

struct LlCount {
    int nslots, slotsA;         // column slots in use (multiple of 4), slots of the first group
    unsigned int mulA, mulB;    // shared: byte strides of codeA / codeB; global: KB, 1
    unsigned int sbase;         // shared-window byte address of this lane's replica (shared counting)
    unsigned int *hist;         // this warp's table in global memory (global counting)
};

__device__ __forceinline__ void ll_red_shared(unsigned int addr)
{
    asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory");
}

// NB = 512-row blocks per warp pass (4: 16 loads in flight per lane; 1 for the light MH kernel)
template <int MODE, bool GLB, int NB>
__device__ __forceinline__ void loglin_count_dense(const unsigned char *D, long long ld, long long b0, long long n, LlCount q,
                                                   const unsigned int *colinfo, int lane)
{
    // rows [b0, n); b0 is a multiple of 512 * NB
    constexpr int NC = MODE == 0 ? 1 : 2;
    uint4 x[4][NB];
    auto issue = [&](long long base, int j0) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const unsigned char *col = D + (size_t)colinfo[16 + j0 + jj] * (size_t)ld + base + lane * 16;
#pragma unroll
            for (int u = 0; u < NB; ++u) x[jj][u] = __ldg((const uint4 *)(col + 512 * u));
        }
    };
    // a, b: this row's codeA / codeB (MODE 1) or the code and nothing (MODE 0, 2)
    auto count1 = [&](unsigned int a, unsigned int b) {
        if (GLB) {
            // (merging the lanes of one cell with match.any before the reduction was measured: the match costs
            // more per row than the serialised reductions on a hot cell save -- 66.9 -> 59.5 M proposals/s)
            if (MODE == 1) atomicAdd(&q.hist[a * q.mulA + b], 1u); else atomicAdd(&q.hist[a], 1u);
        } else {
            if (MODE == 1) ll_red_shared(a * q.mulA + (b * q.mulB + q.sbase)); else ll_red_shared(a * q.mulA + q.sbase);
        }
    };
    issue(b0, 0);
    for (long long base = b0; base < n; base += 512 * NB) {
        unsigned int c[NB][4][NC];
#pragma unroll
        for (int u = 0; u < NB; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int h = 0; h < NC; ++h) c[u][w][h] = 0u;
        for (int j0 = 0; j0 < q.nslots; j0 += 4) {
            unsigned int L4[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) L4[jj] = colinfo[j0 + jj];
            // (MODE 1: ONE warp-uniform branch per group of four slots -- as a per-statement condition it was
            // if-converted and both accumulators' multiply-adds were issued for every slot)
            auto accumulate = [&](auto HSEL) {
                constexpr int H = decltype(HSEL)::value;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                    for (int u = 0; u < NB; ++u) {
                        const unsigned int w4[4] = {x[jj][u].x, x[jj][u].y, x[jj][u].z, x[jj][u].w};
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            if (MODE == 2) {
                                c[u][w][0] = c[u][w][0] * L4[jj] + __byte_perm(w4[w], 0u, 0x4240);            // rows 4w, 4w+2
                                c[u][w][NC - 1] = c[u][w][NC - 1] * L4[jj] + __byte_perm(w4[w], 0u, 0x4341);  // rows 4w+1, 4w+3
                            } else {
                                c[u][w][H] = c[u][w][H] * L4[jj] + w4[w];
                            }
                        }
                    }
                }
            };
            if (MODE == 1 && j0 >= q.slotsA) accumulate(std::integral_constant<int, NC - 1>());
            else accumulate(std::integral_constant<int, 0>());
            const int nj0 = (j0 + 4 < q.nslots) ? j0 + 4 : 0;
            const long long nbase = nj0 ? base : base + 512 * NB;
            if (nbase < n) issue(nbase, nj0);
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const long long r = base + 512 * u + lane * 16;
            if (r < ld) {      // ld is a multiple of 16: the padding rows land in cell 0 and are subtracted by the caller
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const unsigned int a = c[u][w][0], b = c[u][w][NC - 1];
                    if (MODE == 2) {
                        count1(a & 0xffffu, 0u); count1(b & 0xffffu, 0u); count1(a >> 16, 0u); count1(b >> 16, 0u);
                    } else {
                        count1(a & 0xffu, b & 0xffu); count1((a >> 8) & 0xffu, (b >> 8) & 0xffu);
                        count1((a >> 16) & 0xffu, (b >> 16) & 0xffu); count1(a >> 24, b >> 24);
                    }
                }
            }
        }
    }
}

// sum over the cells of a count table of [lgamma(cnt + alpha) - lgamma(alpha)], minus the constant of
// the set: the order (cells dealt to lanes, then a butterfly) is the same for scanned and derived tables
struct LlConst { const double *LG; double alpha, constK, lga; };

__device__ __forceinline__ LlConst ll_constants(const LoglinArgs &P, double Kd)
{
    LlConst r;
    const long long n = P.nrows;
    int tj = -1;
    for (int j = 0; j < P.n_ktab; ++j) if (P.ktab[j] == Kd) tj = j;
    r.LG = tj >= 0 ? P.lgtab + (size_t)tj * (size_t)(n + 2) : nullptr;
    r.alpha = tj >= 0 ? P.alphatab[tj] : P.cell_alpha / Kd;
    r.constK = tj >= 0 ? r.LG[n + 1] : lgamma(Kd * r.alpha + (double)n) - lgamma(Kd * r.alpha);
    r.lga = r.LG ? 0.0 : lgamma(r.alpha);
    return r;
}

__device__ __forceinline__ double ll_table_score(const LlConst &C, const unsigned int *tab, unsigned int cells, int lane)
{
    double acc = 0.0;
    // eight table look-ups in flight per lane; the sum keeps the ascending cell order
    if (C.LG) {
        for (unsigned int c0 = 0u; c0 < cells; c0 += 256u) {
            unsigned int cn[8];
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const unsigned int c = c0 + 32u * u + lane; cn[u] = c < cells ? tab[c] : 0u; }
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = cn[u] ? __ldg(&C.LG[cn[u]]) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) if (cn[u]) acc += t[u];
        }
    } else {
        for (unsigned int c = lane; c < cells; c += 32) {
            const unsigned int cnt = tab[c];
            if (cnt) acc += lgamma((double)cnt + C.alpha) - C.lga;
        }
    }
    return warp_sum(acc) - C.constK;
}

// Sum over a count table in global memory (cells dealt to lanes, ascending, like ll_table_score; the butterfly
// is left to the caller) that also restores the all-zero state of the table.  A table of 2^16 cells is 256 KB
// that has usually left L2 since this warp last used it: one dependent load per round trip made the sweep
// 1.4 M cycles of a 2.3 M-cycle scan, so sixteen rounds of loads are in flight.  `pad`: rows of zero padding
// that were counted into cell 0.
static __device__ __noinline__ double ll_sweep_global(const LlConst C, unsigned int *tab, unsigned int ncell, unsigned int pad, int lane)
{
    double acc = 0.0;
    constexpr int UB = 16;
    if (C.LG) {
        for (unsigned int c0 = 0u; c0 < ncell; c0 += 32u * UB) {
            unsigned int cn[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) { const unsigned int c = c0 + 32u * u + lane; cn[u] = c < ncell ? __ldcg(&tab[c]) : 0u; }
#pragma unroll
            for (int u = 0; u < UB; ++u) if (cn[u]) tab[c0 + 32u * u + lane] = 0u;
            if (c0 == 0u && lane == 0) cn[0] -= pad;
            double t[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) t[u] = cn[u] ? __ldg(&C.LG[cn[u]]) : 0.0;
#pragma unroll
            for (int u = 0; u < UB; ++u) if (cn[u]) acc += t[u];
        }
    } else {
        for (unsigned int c = lane; c < ncell; c += 32u) {
            unsigned int cnt = __ldcg(&tab[c]);
            if (cnt) tab[c] = 0u;
            if (c == 0u) cnt -= pad;
            if (cnt) acc += lgamma((double)cnt + C.alpha) - C.lga;
        }
    }
    return acc;
}

// What the last dense scan left behind: when `cells` > 0 the exact count table of the set whose
// sorted column list is still in ws.idx sits compact in shared memory at sh[0 .. cells).
struct LlTable { unsigned int cells; int k; };

// Score of a proper subset Z of the set whose table `t` describes, by summing the other columns out
// of the table one at a time (ping-pong buffers behind the table; cells <= cap / 2).  The counts are
// exact integers, so the score is bit-identical to what a scan of Z's own columns gives.
template <int WMAX>
__device__ __forceinline__ double loglin_derive(const LoglinArgs &P, const WarpScratch &ws, const LlTable &t,
                                                const KeyW<WMAX> &Z, int lane)
{
    unsigned int *sh = (unsigned int *)ws.tri32;
    const unsigned int *colinfo = sh + P.cap + 32;         // levels of the table's columns by rank
    const int k = t.k;
    // which ranks stay
    bool keep = false;
    if (lane < k) {
        const int g = ws.idx[lane];
        u64 zw = 0ull;
#pragma unroll
        for (int w = 0; w < WMAX; ++w) zw |= Z.w[w] & (0ull - (u64)((g >> 6) == w));     // (no dynamic register indexing)
        keep = (zw >> (g & 63)) & 1ull;
    }
    const u32 zmask = __ballot_sync(PDG_FULL, keep);
    const unsigned int *src = sh;
    unsigned int *bufA = sh + t.cells, *bufB = sh + t.cells + (t.cells >> 1);
    unsigned int *dst = bufA;
    unsigned int hi = 1u;            // cells of the kept columns before rank j
    unsigned int lo = t.cells;       // cells of the columns from rank j on (all still present)
    double Kd = 1.0;
    for (int j = 0; j < k; ++j) {
        const unsigned int L = colinfo[j];
        lo /= L;                     // now: cells of the columns after rank j
        if ((zmask >> j) & 1u) { hi *= L; Kd *= (double)L; continue; }
        if (L == 1u) continue;
        const unsigned int nout = hi * lo;
        for (unsigned int o = lane; o < nout; o += 32) {
            const unsigned int h = o / lo, l = o - h * lo;
            const unsigned int *in = src + (size_t)h * L * lo + l;
            unsigned int sum = 0u;
            for (unsigned int d = 0; d < L; ++d) sum += in[(size_t)d * lo];
            dst[o] = sum;
        }
        __syncwarp();
        src = dst;
        dst = (dst == bufA) ? bufB : bufA;
    }
    const LlConst C = ll_constants(P, Kd);
    const double sc = ll_table_score(C, src, hi, lane);
    __syncwarp();
    return sc;
}

// counts the chunks of request `seq` this warp manages to take, into the table `q` describes; returns how many
template <int NB>
static __device__ __noinline__ int help_chunks(HelpBox *hb, int seq, const unsigned char *D, LlCount q,
                                               const unsigned int *colinfo, int lane)
{
    const int mode = hb->mode, glb = hb->glb, n_chunks = hb->n_chunks;
    const long long n = hb->n, ld = hb->ld;
    int mine = 0;
    for (;;) {
        int c = -1;
        if (lane == 0) {
            for (;;) {
                const int old = *(volatile int *)&hb->ticket;
                if ((old >> 16) != seq || (old & 0xffff) >= n_chunks) break;
                if (atomicCAS(&hb->ticket, old, old + 1) == old) { c = old & 0xffff; break; }
            }
        }
        c = __shfl_sync(PDG_FULL, c, 0);
        if (c < 0) break;
        const long long b0 = (long long)c * PDG_HELP_ROWS;
        const long long b1 = (b0 + PDG_HELP_ROWS < n) ? b0 + PDG_HELP_ROWS : n;
        if (mode == 0) loglin_count_dense<0, false, NB>(D, ld, b0, b1, q, colinfo, lane);
        else if (!glb) loglin_count_dense<1, false, NB>(D, ld, b0, b1, q, colinfo, lane);
        else loglin_count_dense<1, true, NB>(D, ld, b0, b1, q, colinfo, lane);
        ++mine;
    }
    return mine;
}

// what a warp does once its own chain has ended: serve the scans of the CTA's other chains until all
// of them have ended (`live_warps` chains in this CTA)
template <int NB>
__device__ __forceinline__ void help_until_done(HelpBox *hb, const unsigned char *D, unsigned int *sh, int cap, int live_warps, int lane)
{
    if (lane == 0) atomicAdd(&hb->finished, 1);
    for (;;) {
        int fin = 0, t = 0, open = 0;
        if (lane == 0) { fin = *(volatile int *)&hb->finished; t = *(volatile int *)&hb->ticket; open = *(volatile int *)&hb->lock; }
        fin = __shfl_sync(PDG_FULL, fin, 0); t = __shfl_sync(PDG_FULL, t, 0); open = __shfl_sync(PDG_FULL, open, 0);
        if (fin >= live_warps) break;
        const int seq = t >> 16;
        if (!open || seq == 0 || (t & 0xffff) >= *(volatile int *)&hb->n_chunks) { __nanosleep(1000); continue; }
        // (the request's fields were written before its ticket; a stale read fails the ticket's compare-and-swap)
        const int glb = *(volatile int *)&hb->glb, rshift = *(volatile int *)&hb->rshift;
        const unsigned int cells = *(volatile unsigned int *)&hb->cells;
        if (!glb && (cells == 0u || cells > (unsigned int)cap)) { __nanosleep(1000); continue; }
        LlCount q;
        q.nslots = *(volatile int *)&hb->nslots; q.slotsA = *(volatile int *)&hb->slotsA;
        q.mulA = *(volatile unsigned int *)&hb->mulA; q.mulB = *(volatile unsigned int *)&hb->mulB;
        q.hist = *(unsigned int * volatile *)&hb->hist;
        q.sbase = (unsigned int)__cvta_generic_to_shared(sh) + 4u * (unsigned int)(lane & ((1 << rshift) - 1));
        if (!glb) for (unsigned int c = lane; c < cells; c += 32) sh[c] = 0u;
        __syncwarp();
        const int mine = help_chunks<NB>(hb, seq, D, q, hb->colinfo, lane);
        __syncwarp();
        if (mine) {
            if (!glb) {
                unsigned int *dst = *(unsigned int * volatile *)&hb->req_table;
                for (unsigned int c = lane; c < cells; c += 32) { const unsigned int v = sh[c]; if (v) atomicAdd(&dst[c], v); }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) atomicAdd(&hb->done, mine);
        }
    }
}

template <int WMAX, int NB>
__device__ __forceinline__ ScoreRes loglin_scan(LoglinArgs P, WarpScratch ws, int slot, KeyW<WMAX> XK, int lane, LlTable &tout)
{
    ScoreRes res; res.score = 0.0; res.err = 0;
    tout.cells = 0u; tout.k = 0;
    double score = 0.0;
#ifdef PDG_PHASE_TIMERS
    const long long kh_t0 = clock64();
#endif
    const int k = build_index_sorted<WMAX>(XK.w, P.W, lane, ws.idx);
    if (k == 0) return res;
    const long long n = P.nrows;
    double Kd = 1.0;
    u64 Ki = 1ull;
    bool big = false;
    // greedy split into two column groups of <= 256 cells each (byte-packed codes)
    int nA = 0;
    unsigned int prodA = 1u, prodB = 1u;
    bool splitB = false, twoOk = true;
    for (int j = 0; j < k; ++j) {
        const u64 Lj = (u64)P.levels[ws.idx[j]];
        Kd *= (double)Lj;
        if (Ki > (1ull << 62) / Lj) big = true; else Ki *= Lj;
        if (!splitB && (u64)prodA * Lj <= 256ull) { prodA *= (unsigned int)Lj; ++nA; }
        else { splitB = true; if ((u64)prodB * Lj <= 256ull) prodB *= (unsigned int)Lj; else twoOk = false; }
    }
    const LlConst C = ll_constants(P, Kd);
    const unsigned char *D = P.data;
    double acc = 0.0;
    if (!big && Ki <= (u64)P.hist_cells && k <= 16) {
        // shared layout: `cap` counters, then 16 multipliers and 16 column ids by SLOT and the 16
        // levels of the columns by RANK (for loglin_derive)
        unsigned int *sh = (unsigned int *)ws.tri32;
        unsigned int *colinfo = sh + P.cap;
        const int slotsA = (nA + 3) & ~3, slotsB = (k - nA + 3) & ~3;
        twoOk = twoOk && slotsA + slotsB <= 16;
        const int mode = (Ki <= 256ull) ? 0 : twoOk ? 1 : 2;
        if (lane < 16) {
            int j = lane;                                   // rank held by this slot, -1: empty
            if (mode == 1) j = lane < slotsA ? (lane < nA ? lane : -1) : (nA + lane - slotsA < k ? nA + lane - slotsA : -1);
            else if (lane >= k) j = -1;
            const int g = j >= 0 ? (int)ws.idx[j] : P.zero_col;
            colinfo[lane] = j >= 0 ? (unsigned int)P.levels[g] : 1u;
            colinfo[16 + lane] = (unsigned int)g;
            colinfo[32 + lane] = lane < k ? (unsigned int)P.levels[ws.idx[lane]] : 1u;
        }
        const bool glb = Ki > (u64)P.cap;
        int rshift = 0;
        if (!glb) while (rshift < 5 && (Ki << (rshift + 1)) <= (u64)P.cap) ++rshift;
        const int R = 1 << rshift;
        unsigned int *hist = P.hist + (size_t)slot * P.hist_cells;      // all zero on entry
        if (!glb) for (int c = lane; c < ((int)Ki << rshift); c += 32) sh[c] = 0u;
        __syncwarp();
        LlCount q;
        q.nslots = mode == 1 ? slotsA + slotsB : (k + 3) & ~3;
        q.slotsA = mode == 1 ? slotsA : q.nslots;
        q.hist = hist;
        q.sbase = (unsigned int)__cvta_generic_to_shared(sh) + 4u * (unsigned int)(lane & (R - 1));
        if (glb) { q.mulA = mode == 1 ? prodB : 1u; q.mulB = 1u; }
        else { q.mulB = 4u << rshift; q.mulA = mode == 1 ? prodB * q.mulB : q.mulB; }
        // large scans are offered to the warps of the CTA whose chains have ended (MH kernel only)
        HelpBox *hb = ws.help;
        bool coop = false;
        if (hb && mode != 2 && n >= 4 * PDG_HELP_ROWS) {
            int got = 0;
            // (only once some chain of the CTA has ended: the chunked pass gives up the prefetch across chunk ends)
            if (lane == 0 && *(volatile int *)&hb->finished > 0) got = atomicCAS(&hb->lock, 0, 1) == 0 ? 1 : 0;
            coop = __shfl_sync(PDG_FULL, got, 0) != 0;
        }
        if (coop) {
            hb->colinfo[lane] = colinfo[lane];
            __syncwarp();
            int seq = 0;
            if (lane == 0) {
                hb->mode = mode; hb->glb = glb ? 1 : 0; hb->nslots = q.nslots; hb->slotsA = q.slotsA; hb->rshift = rshift;
                hb->mulA = q.mulA; hb->mulB = q.mulB; hb->cells = glb ? 0u : ((unsigned int)Ki << rshift);
                hb->req_table = sh; hb->hist = hist; hb->n = n; hb->ld = P.ld;
                hb->n_chunks = (int)((n + PDG_HELP_ROWS - 1) / PDG_HELP_ROWS); hb->done = 0;
                seq = ((*(volatile int *)&hb->ticket >> 16) + 1) & 0x7fff;
                if (seq == 0) seq = 1;
                __threadfence_block();
                *(volatile int *)&hb->ticket = seq << 16;            // published last
            }
            seq = __shfl_sync(PDG_FULL, seq, 0);
            __syncwarp();
            const int mine = help_chunks<NB>(hb, seq, D, q, colinfo, lane);
            if (lane == 0) {
                atomicAdd(&hb->done, mine);
                while (*(volatile int *)&hb->done < hb->n_chunks) __nanosleep(200);
                __threadfence_block();
                *(volatile int *)&hb->lock = 0;
            }
            __syncwarp();
        } else if (mode == 0) loglin_count_dense<0, false, NB>(D, P.ld, 0, n, q, colinfo, lane);
        else if (mode == 1 && !glb) loglin_count_dense<1, false, NB>(D, P.ld, 0, n, q, colinfo, lane);
        else if (mode == 1) loglin_count_dense<1, true, NB>(D, P.ld, 0, n, q, colinfo, lane);
        else if (!glb) loglin_count_dense<2, false, NB>(D, P.ld, 0, n, q, colinfo, lane);
        else loglin_count_dense<2, true, NB>(D, P.ld, 0, n, q, colinfo, lane);
        __syncwarp();
        if (!glb) {
            // fold the replicas into a compact table sh[0 .. Ki) (32 cells at a time: all reads of a
            // round come before its writes, and a round only overwrites replicas already folded)
            const unsigned int pad = (unsigned int)(P.ld - n);      // zero padding rows of the last block
            for (unsigned int c0 = 0; c0 < (unsigned int)Ki; c0 += 32) {
                const unsigned int c = c0 + lane;
                unsigned int cnt = 0u;
                if (c < (unsigned int)Ki) for (int r = 0; r < R; ++r) cnt += sh[(c << rshift) + ((r + lane) & (R - 1))];
                if (c == 0u) cnt -= pad;
                __syncwarp();
                if (c < (unsigned int)Ki) sh[c] = cnt;
                __syncwarp();
            }
            score = ll_table_score(C, sh, (unsigned int)Ki, lane);
            if (2ull * Ki <= (u64)P.cap) { tout.cells = (unsigned int)Ki; tout.k = k; }
        } else {
            __threadfence();
            __syncwarp();
            acc = ll_sweep_global(C, hist, (unsigned int)Ki, (unsigned int)(P.ld - n), lane);
            __threadfence();
            __syncwarp();
            score = warp_sum(acc) - C.constK;
        }
        res.score = score;
#ifdef PDG_PHASE_TIMERS
        if (lane == 0) { atomicAdd(&g_kh[0][k < 31 ? k : 31], 1ull); atomicAdd(&g_kh[1][k < 31 ? k : 31], (unsigned long long)(clock64() - kh_t0)); }
        if (lane == 0 && k >= 13) atomicAdd(&P.work[(size_t)slot * 16 + 15], (unsigned long long)(clock64() - kh_t0));    // per chain: cycles in scans of 13+ columns
#endif
        return res;
    }
    const double *LG = C.LG;
    const double alpha = C.alpha, constK = C.constK, lga = C.lga;
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
    // slots taken by this scan, one list per lane (a lane meets at most its own rows: 16 per 512-row block)
    int *mylist = S.list + (size_t)s * (size_t)(n + 512) + (size_t)lane * (size_t)(16 * ((n + 511) / 512));
    int nlist = 0;
    const u32 hmask = (u32)S.hslots - 1u;
    // Keys.  The sorted columns are split greedily into groups of <= 256 cells each (<= 8 groups: the code fits
    // 62 bits); every lane covers 16 consecutive rows with one 16-byte load per column, and a group's code
    // is built with the dense pass's packed arithmetic -- four rows per 32-bit multiply-add -- into one
    // byte-packed accumulator per four rows; only the few group codes of a row are then combined in 64 bits.
    // Groups are padded to four column slots with the zero column and multiplier 1.  (One row per lane, one
    // byte per load and a 64-bit multiply-add per column and row took 14 ms per 17-column scan; 64-bit
    // multiply-adds over 16 rows per load still 2.2 ms.)
    // Insertion: a slot that already holds the key is found by a plain L2 read -- eight in flight -- with the
    // count added fire-and-forget; only a key's first insertion waits for a compare-and-swap and records its
    // slot in the list.
#ifdef PDG_PHASE_TIMERS
    long long sp_t = clock64();
    if (lane == 0) atomicAdd(&g_sp[7], (unsigned long long)(sp_t - kh_t0));      // wait for a pool
#endif
    unsigned int *sh = (unsigned int *)ws.tri32;          // the count table is idle on this path: slot lists live there
    unsigned int *slot_mul = sh, *slot_col = sh + 64, *grp_lo = sh + 128, *grp_cells = sh + 144;   // [64] [64] [9] [8]
    int G = 0;
    if (lane == 0) {
        int nslot = 0, g = 0;
        unsigned int prod = 1u;
        grp_lo[0] = 0u;
        for (int j = 0; j < k; ++j) {
            const unsigned int gcol = (unsigned int)ws.idx[j], Lj = (unsigned int)P.levels[gcol];
            if (prod * Lj > 256u) {                            // close the group: pad to four slots
                while (nslot & 3) { slot_mul[nslot] = 1u; slot_col[nslot] = (unsigned int)P.zero_col; ++nslot; }
                grp_cells[g] = prod; ++g; grp_lo[g] = (unsigned int)nslot; prod = 1u;
            }
            slot_mul[nslot] = Lj; slot_col[nslot] = gcol; ++nslot; prod *= Lj;
        }
        while (nslot & 3) { slot_mul[nslot] = 1u; slot_col[nslot] = (unsigned int)P.zero_col; ++nslot; }
        grp_cells[g] = prod; ++g; grp_lo[g] = (unsigned int)nslot;
        G = g;
    }
    G = __shfl_sync(PDG_FULL, G, 0);
    __syncwarp();
    if (G > 8) { if (lane == 0) atomicExch(&S.lock[s], 0); res.err = ERR_BIGCLIQUE; return res; }
    // Tables that fit the pool's key array taken as 32-bit counters (2 * hslots cells: 17-19 binary columns at
    // n = 10^5) are counted DENSELY there -- one fire-and-forget reduction per row on the mixed-radix code,
    // then the sweep of the global dense path; a hashed insertion is a compare-and-swap whose result the
    // warp waits for plus a reduction per row (8.7 M cycles per scan in the sampler).
    const bool pool_dense = !big && Ki <= 2ull * (u64)S.hslots;
    unsigned int *ptab = (unsigned int *)hk;
    // a lane that leads the same key again and again (the table's heaviest cell, as a rule) keeps that cell's
    // count in registers and adds it once: repeated reductions on one address were most of the insertion time
    u64 pend_key = 0ull;
    u32 pend_h = 0u, pend_cnt = 0u;
    // The key-building part is rolled over the column groups (the fully unrolled form of the loop body was
    // 28 KB of straight-line code next to the MH loop of the SM's other warps).
    for (long long base = 0; base < n; base += 512) {
        const long long r0 = base + lane * 16;
        u64 keys[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) keys[i] = 0ull;
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            unsigned int c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
            const int lo = (int)grp_lo[g], hi = (int)grp_lo[g + 1];
            // (a group of binary columns is eight slots: all of its loads are in flight together)
#pragma unroll 1
            for (int s0 = lo; s0 < hi; s0 += 8) {
                uint4 x[8];
                unsigned int L8[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const bool in = s0 + jj < hi;
                    L8[jj] = in ? slot_mul[s0 + jj] : 1u;
                    x[jj] = in ? __ldg((const uint4 *)(D + (size_t)slot_col[s0 + jj] * (size_t)P.ld + r0)) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    c0 = c0 * L8[jj] + x[jj].x; c1 = c1 * L8[jj] + x[jj].y;
                    c2 = c2 * L8[jj] + x[jj].z; c3 = c3 * L8[jj] + x[jj].w;
                }
            }
            const u64 Kg = (u64)grp_cells[g];            // (keys are 0 before the first group)
            const unsigned int cw[4] = {c0, c1, c2, c3};
#pragma unroll
            for (int i = 0; i < 16; ++i) keys[i] = keys[i] * Kg + (u64)((cw[i >> 2] >> (8 * (i & 3))) & 0xffu);
        }
        SPT(0)
        if (pool_dense) {
#pragma unroll
            for (int u = 0; u < 16; ++u) if (r0 + u < n) atomicAdd(&ptab[(u32)keys[u]], 1u);
            continue;
        }
        // All sixteen rows of the lane in ONE wave of compare-and-swaps: an empty slot is taken (a new cell: the
        // lane notes the slot in its own list), a slot holding the key is a hit, anything else walks on.  The
        // tables of real cliques hold thousands of thinly filled cells, so nearly every batch of rows opens new
        // cells and meets an occupied home slot somewhere; with a read before the compare-and-swap, a shared
        // list counter and one probe at a time a batch was six dependent L2 round trips (12.4 M cycles a scan).
        u32 hh[16];
        u64 old[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            keys[u] = (r0 + u < n) ? keys[u] + 1ull : 0ull;                       // 0: no row
            hh[u] = (u32)((keys[u] * 0x9E3779B97F4A7C15ull) >> 32) & hmask;      // (Fibonacci hashing: one multiply)
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (keys[u] == 0ull || keys[u] == pend_key) { old[u] = keys[u]; if (keys[u]) hh[u] = pend_h; }      // no row / slot known
            else old[u] = atomicCAS(&hk[hh[u]], 0ull, keys[u]);
        }
        u32 restm = 0u;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (keys[u] == 0ull) continue;
            if (old[u] == 0ull) mylist[nlist++] = (int)hh[u];
            else if (old[u] != keys[u]) { restm |= 1u << u; continue; }
            if (keys[u] == pend_key) ++pend_cnt;
            else {
                if (pend_cnt) atomicAdd(&hc[pend_h], pend_cnt);
                pend_key = keys[u]; pend_h = hh[u]; pend_cnt = 1u;
            }
        }
        // linear probing, every unresolved row of the lane one step per round
        while (__any_sync(PDG_FULL, restm != 0u)) {
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if ((restm >> u) & 1u) { hh[u] = (hh[u] + 1u) & hmask; old[u] = atomicCAS(&hk[hh[u]], 0ull, keys[u]); }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                if (!((restm >> u) & 1u)) continue;
                if (old[u] == 0ull) mylist[nlist++] = (int)hh[u];
                else if (old[u] != keys[u]) continue;
                atomicAdd(&hc[hh[u]], 1u);
                restm &= ~(1u << u);
            }
        }
        SPT(1)
    }
    if (pend_cnt) atomicAdd(&hc[pend_h], pend_cnt);
    __threadfence();
    __syncwarp();
    if (pool_dense) {
        acc = ll_sweep_global(C, ptab, (unsigned int)Ki, 0u, lane);
        __threadfence();
        __syncwarp();
        SPT(1)
#ifdef PDG_PHASE_TIMERS
        if (lane == 0) atomicAdd(&g_sp[4], 1ull);
#endif
        if (lane == 0) atomicExch(&S.lock[s], 0);
        res.score = warp_sum(acc) - constK;
#ifdef PDG_PHASE_TIMERS
        if (lane == 0) { atomicAdd(&g_kh[0][k < 31 ? k : 31], 1ull); atomicAdd(&g_kh[1][k < 31 ? k : 31], (unsigned long long)(clock64() - kh_t0)); }
        if (lane == 0) atomicAdd(&P.work[(size_t)slot * 16 + 15], (unsigned long long)(clock64() - kh_t0));
#endif
        return res;
    }
    // the occupied slots come off the lanes' lists (a few thousand cells), not off a pass over the whole table
    const int ncell = (int)__reduce_add_sync(PDG_FULL, (unsigned int)nlist);
    const int nlmax = (int)__reduce_max_sync(PDG_FULL, (unsigned int)nlist);
    unsigned int maxc = 0;
    for (int i0 = 0; i0 < nlmax; i0 += 4) {
        int hx[4];
        unsigned int cx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) hx[u] = i0 + u < nlist ? mylist[i0 + u] : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) cx[u] = hx[u] >= 0 ? __ldcg(&hc[hx[u]]) : 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (hx[u] < 0) continue;
            atomicAdd(&cc[cx[u]], 1u);
            maxc = cx[u] > maxc ? cx[u] : maxc;
            hk[hx[u]] = 0ull; hc[hx[u]] = 0u;
        }
    }
    maxc = __reduce_max_sync(PDG_FULL, maxc);
    __threadfence();
    __syncwarp();
    SPT(2)
    // count-of-counts in ascending count order (order-free sum), eight loads in flight per lane
    for (unsigned int c0 = 1; c0 <= maxc; c0 += 256) {
        unsigned int mx[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const unsigned int c = c0 + 32 * u + lane; mx[u] = c <= maxc ? __ldcg(&cc[c]) : 0u; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned int c = c0 + 32 * u + lane;
            if (mx[u]) {
                acc += (double)mx[u] * (LG ? LG[c] : (lgamma((double)c + alpha) - lga));
                cc[c] = 0u;
            }
        }
    }
    __threadfence();
    __syncwarp();
    SPT(3)
#ifdef PDG_PHASE_TIMERS
    if (lane == 0) { atomicAdd(&g_sp[4], 1ull); atomicAdd(&g_sp[5], (unsigned long long)ncell); atomicAdd(&g_sp[6], (unsigned long long)maxc); }
#endif
    if (lane == 0) atomicExch(&S.lock[s], 0);
    score = warp_sum(acc) - constK;
    res.score = score;
#ifdef PDG_PHASE_TIMERS
    if (lane == 0) { atomicAdd(&g_kh[0][k < 31 ? k : 31], 1ull); atomicAdd(&g_kh[1][k < 31 ? k : 31], (unsigned long long)(clock64() - kh_t0)); }
    if (lane == 0) atomicAdd(&P.work[(size_t)slot * 16 + 15], (unsigned long long)(clock64() - kh_t0));
#endif
    return res;
}


// inlined shims used by the kernels
__device__ __forceinline__ LoglinArgs loglin_args(const PdgParams &P)
{
    LoglinArgs A;
    A.data = P.data; A.levels = P.levels; A.nrows = P.nrows; A.ld = P.data_ld; A.zero_col = P.p; A.cell_alpha = P.cell_alpha;
    A.n_ktab = P.n_ktab; A.ktab = P.ktab; A.alphatab = P.alphatab; A.lgtab = P.lgtab;
    A.hist = P.hist; A.hist_cells = P.hist_cells; A.cap = P.ll_cap; A.sp = P.sp; A.work = P.work; A.W = P.W;
    return A;
}

template <int WMAX, int NB>
__device__ __forceinline__ int loglin_set_score(const PdgParams &P, const WarpScratch &ws, int slot,
                                                const u64 (&X)[WMAX], int lane, double &score, LlTable &t)
{
    KeyW<WMAX> K;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) K.w[w] = X[w];
    const ScoreRes r = loglin_scan<WMAX, NB>(loglin_args(P), ws, slot, K, lane, t);
    score = r.score;
    return r.err;
}

template <int WMAX>
__device__ __forceinline__ double loglin_subset_score(const PdgParams &P, const WarpScratch &ws, const LlTable &t,
                                                      const u64 (&Z)[WMAX], int lane)
{
    KeyW<WMAX> K;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) K.w[w] = Z[w];
    return loglin_derive<WMAX>(loglin_args(P), ws, t, K, lane);
}
