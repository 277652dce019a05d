// Kernels and launchers that are instantiated per likelihood model.  Included by exactly one
// translation unit per model (pdg_model_<name>.cu defines PDG_TU_MODEL and PDG_TU_NAME), so that
// the three sets of template instantiations compile in parallel.
#include <stdlib.h>

#include "pdg_ctx.h"
#include "pdg_run.cuh"
#include "pdg_ggm_big.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// logl[0] (mh_parallel.py:59-60 / :217-218): clique scores in tree-node order minus
// nu_s * score(s) over the DISTINCT non-empty separators (first occurrence in edge order), plus
// log_prior over cliques and distinct separators (seqdist:44-52).  One warp per chain.
struct InitSmem { int off_tri, off_lidx, off_idx, off_sc, off_ss, off_cs, off_ssz, off_nu, off_first, total; };
__host__ __device__ inline InitSmem init_smem_layout(int p, int W, int tri_bytes)
{
    InitSmem s; int o = 0;
    const int a2 = ((p + 2) * 2 + 15) / 16 * 16;
#ifdef PDG_LANE_SMEM
    s.off_tri = o; o += PDG_KTRI * 32 * 8;
    s.off_lidx = o; o += PDG_KT * 32 * 2;
#else
    s.off_tri = o; o += tri_bytes;
    s.off_lidx = o; o += PDG_KT * 32 * 2;
#endif
    s.off_sc = o; o += p * 8;
    s.off_ss = o; o += p * 8;
    s.off_idx = o; o += a2;
    s.off_cs = o; o += a2;
    s.off_ssz = o; o += a2;
    s.off_nu = o; o += a2;
    s.off_first = o; o += (p + 15) / 16 * 16;
    s.total = o;
    return s;
}

// one set, same value on every lane: lane 0 carries the key through warp_scores so that logl[0]
// uses exactly the score function of the sampler
template <int MODEL, int WMAX>
__device__ __forceinline__ int set_score_any(const PdgParams &P, const WarpScratch &ws, const LaneScratch &ls, int slot,
                                             const u64 (&X)[WMAX], int lane, double &sc)
{
    u32 rs = 0u, rk = 0u;
    double v = 0.0;
    const int e = warp_scores<MODEL, WMAX>(P, ws, ls, slot, lane, X, lane == 0, v, rs, rk);
    sc = __shfl_sync(PDG_FULL, v, 0);
    return __reduce_or_sync(PDG_FULL, (unsigned)e);
}

template <int MODEL, int WMAX>
__global__ void __launch_bounds__(32) k_init_logl(PdgParams P, int kind)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int p = P.p, W = P.W, lane = threadIdx.x;
    const long long chain = blockIdx.x;
    const InitSmem L = init_smem_layout(p, W, tri_bytes_for(P.ll_cap));
    double *s_sc = (double *)(smem + L.off_sc), *s_ss = (double *)(smem + L.off_ss);
    unsigned short *s_cs = (unsigned short *)(smem + L.off_cs), *s_ssz = (unsigned short *)(smem + L.off_ssz);
    unsigned short *s_nu = (unsigned short *)(smem + L.off_nu);
    unsigned char *s_first = smem + L.off_first;
    WarpScratch ws;
    ws.help = nullptr;
    ws.ktri = 32;
    ws.tri32 = (double *)(smem + L.off_tri);
    ws.idx = (short *)(smem + L.off_idx);
    ws.kbig = P.kbig;
    const int slot = (int)chain;
    ws.big = P.big ? P.big + (size_t)slot * ((size_t)(P.kbig + 1) * (P.kbig + 2) / 2) : nullptr;
    LaneScratch ls;
    ls.unused = 0;
    ls.a = (double *)(smem + L.off_tri);
    ls.idx = (unsigned short *)(smem + L.off_lidx);
    const u64 *Mc = P.M + (size_t)chain * p * W;
    const int *E = P.edges + (size_t)chain * 2 * (p - 1);
    int errflag = 0;
    for (int t = 0; t < p; ++t) {
        u64 X[WMAX];
        load_row<WMAX>(X, Mc + (size_t)t * W, W);
        int cs = 0;
#pragma unroll
        for (int w = 0; w < WMAX; ++w) cs += __popcll(X[w]);
        double sc = 0.0;
        if (MODEL != PDG_MODEL_UNIFORM && cs > 0) errflag |= set_score_any<MODEL, WMAX>(P, ws, ls, slot, X, lane, sc);
        if (lane == 0) { s_sc[t] = sc; s_cs[t] = (unsigned short)cs; }
    }
    for (int e = 0; e < p - 1; ++e) {
        const int a = E[2 * e], b = E[2 * e + 1];
        u64 X[WMAX];
        int sz = 0;
#pragma unroll
        for (int w = 0; w < WMAX; ++w) { X[w] = (w < W) ? (Mc[(size_t)a * W + w] & Mc[(size_t)b * W + w]) : 0ull; sz += __popcll(X[w]); }
        int nu = 0; bool first = true;
        for (int e2 = lane; e2 < p - 1; e2 += 32) {
            const int a2 = E[2 * e2], b2 = E[2 * e2 + 1];
            bool same = true;
#pragma unroll
            for (int w = 0; w < WMAX; ++w) if (w < W) same = same && ((Mc[(size_t)a2 * W + w] & Mc[(size_t)b2 * W + w]) == X[w]);
            if (same) { ++nu; if (e2 < e) first = false; }
        }
        nu = warp_sum_i(nu);
        first = __all_sync(PDG_FULL, first);
        double sc = 0.0;
        if (MODEL != PDG_MODEL_UNIFORM && first && sz > 0) errflag |= set_score_any<MODEL, WMAX>(P, ws, ls, slot, X, lane, sc);
        if (lane == 0) { s_ss[e] = sc; s_ssz[e] = (unsigned short)sz; s_nu[e] = (unsigned short)nu; s_first[e] = first ? 1 : 0; }
    }
    __syncwarp();
    if (lane == 0) {
        double cl = 0.0, lclq = 0.0;
        for (int t = 0; t < p; ++t) {
            if (s_cs[t] == 0) continue;
            cl += s_sc[t];
            const double a = (double)s_cs[t];
            if (P.prior == PDG_PRIOR_MBC) lclq += P.prior_a * (a - 1.0);
            else if (P.prior == PDG_PRIOR_EDGEPENALTY) lclq += -P.prior_a * a * (a - 1.0) / 2.0;
        }
        double sc = 0.0, lsep = 0.0;
        for (int e = 0; e < p - 1; ++e) {
            if (!s_first[e]) continue;
            const double b = (double)s_ssz[e];
            if (P.prior == PDG_PRIOR_MBC) lsep += P.prior_b * b;
            else if (P.prior == PDG_PRIOR_EDGEPENALTY) lsep += -P.prior_a * b * (b - 1.0) / 2.0;
            if (s_ssz[e] == 0) continue;
            sc += (double)s_nu[e] * s_ss[e];
        }
        double prior;
        if (P.prior == PDG_PRIOR_JUMPPENALTY) prior = -P.prior_a * (double)(1 - (kind == PDG_KIND_SINGLE ? 0 : 1));
        else if (P.prior == PDG_PRIOR_UNIFORM) prior = 0.0;
        else prior = lclq - lsep;
        P.logl[(size_t)chain * P.n_samples] = (cl - sc) + prior;
        P.kcount[chain] = (kind == PDG_KIND_SINGLE) ? 1 : 0;
        P.nrec[chain] = 0;
    }
    if (errflag) atomicOr(&P.err[chain], errflag);
}

// batch scorer (sd.log_likelihood_partial([c], {})): every LANE takes one complete set, so a
// warp scores 32 sets per pass through the same routine the sampler uses (warp_scores); sets
// above PDG_KT vertices and log-linear sets are resolved cooperatively.  grid <= n_chains because
// the per-warp scratch (large triangles, cell tables) is sized per chain.
template <int MODEL, int WMAX>
__global__ void __launch_bounds__(32) k_score_sets(PdgParams P, const u64 *bits, long long n_sets, double *out, int *err,
                                                   int *big_list, int *big_count, int big_kmax, int group)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int W = P.W, lane = threadIdx.x;
    WarpScratch ws;
    ws.help = nullptr;
    ws.ktri = 32;
    LaneScratch ls;
    ws.tri32 = (double *)smem;
    ls.unused = 0;
#ifdef PDG_LANE_SMEM
    ls.a = (double *)smem;
    ls.idx = (unsigned short *)(smem + PDG_KTRI * 32 * 8);
    ws.idx = (short *)(smem + PDG_KTRI * 32 * 8 + PDG_KT * 32 * 2);
#else
    ls.a = nullptr;
    ls.idx = (unsigned short *)(smem + tri_bytes_for(P.ll_cap));
    ws.idx = (short *)(smem + tri_bytes_for(P.ll_cap) + PDG_KT * 32 * 2);
#endif
    ws.kbig = P.kbig;
    const int slot = blockIdx.x;
    ws.big = P.big ? P.big + (size_t)slot * ((size_t)(P.kbig + 1) * (P.kbig + 2) / 2) : nullptr;
    u32 rs = 0u, rk = 0u;
    // `group` sets per warp pass (1 ... 32, lanes 0 ... group - 1): sets that miss the memo and need the whole
    // warp (contingency-table scans, cliques of 13 ... 32 vertices) are resolved one after the other, so a
    // batch is spread over as many warps as the launch has before any warp takes a second set
    for (long long s0 = (long long)blockIdx.x * group; s0 < n_sets; s0 += (long long)gridDim.x * group) {
        const long long s = (lane < group) ? s0 + lane : n_sets;
        u64 X[WMAX];
        int k = 0;
#pragma unroll
        for (int w = 0; w < WMAX; ++w) { X[w] = (s < n_sets && w < W) ? bits[(size_t)s * W + w] : 0ull; k += __popcll(X[w]); }
        // GGM sets above 32 vertices go to the CTA-per-set kernel (pdg_ggm_big.cuh): listed here, scored there
        bool mine = s < n_sets && k > 0;
        if (MODEL == PDG_MODEL_GGM && big_list && mine && k >= PDG_BIG_MIN && k <= big_kmax) {
            big_list[atomicAdd(big_count, 1)] = (int)s;
            atomicMax(big_count + 1, k);                  // largest listed set: sizes the CTA kernel's shared memory
            mine = false;
        }
        double sc = 0.0;
        int e = 0;
        if (MODEL != PDG_MODEL_UNIFORM) e = warp_scores<MODEL, WMAX>(P, ws, ls, slot, lane, X, mine, sc, rs, rk);
        if (e) atomicOr(err, e);
        if (s < n_sets && (mine || k == 0)) out[s] = sc;
        __syncwarp();
    }
}

template <typename F>
int set_smem(pdg_ctx *ctx, F f, int bytes)
{
    if (bytes > 48 * 1024) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return PDG_OK;
}

#define PDG_SPEC_NO_FIT (-100)
template <int MODEL, int WMAX, bool SS, bool LIGHT, bool DIAG, bool SPEC>
int launch_run_k(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long randomize)
{
    const RunSmem L = run_smem_layout(ctx->P.p, ctx->P.W, SS, tri_bytes_for(ctx->P.ll_cap) + ctx->P.tri_extra);
    const RandSmem LR = rand_smem_layout(ctx->P.p, ctx->P.W);
    // shared memory per chain: one warp region, or the two regions of a pair back to back; the in-kernel
    // randomiser may use all of it
    const int run_bytes = L.total * (SPEC ? 2 : 1);
    const int per_warp = (randomize > 0 && LR.total > run_bytes) ? LR.total : run_bytes;
    const int wpc = pdg_chains_per_cta(ctx, LIGHT);          // chains per CTA; SPEC: two warps each
    const int warps = wpc * (SPEC ? 2 : 1);
    const int sm = per_warp * wpc;
    if (SPEC && sm > 227 * 1024) return PDG_SPEC_NO_FIT;     // the caller falls back to one warp per chain
    int rc;
    if ((rc = set_smem(ctx, k_mh_run<MODEL, WMAX, SS, LIGHT, DIAG, SPEC>, sm))) return rc;
    const unsigned grid = (unsigned)((ctx->P.n_chains + wpc - 1) / wpc);
    pdg_time_begin(ctx, ctx->ev_run);
    // Instruction-cache policy.  The kernel is ~14 k instructions; the loop body alone is 28 KB (ncu), the
    // 32 KB instruction cache is shared by every warp of the SM, and the in-kernel randomiser (run by each
    // warp every `randomize` iterations) streams another ~90 KB through it.  Two remedies exist: `rsync` --
    // the chains of a CTA enter every randomisation together (one CTA barrier per interval), so its code
    // passes through the cache once per interval; `lockstep` -- the chains re-align every N iterations
    // (power of two >= 2) and share the fetches of the loop body.  Measured (1024 chains, M proposals/s):
    //   p = 50 GGM, pairs   : neither 465, rsync 526            -> rsync with pairs
    //   p = 50 GGM, no pairs: neither 495, rsync 483
    //   p = 500 GGM         : lockstep-16 50.4, neither 43.9, rsync 55.1, both 50.6   -> rsync for large p
    //   log-linear p = 100  : neither 43.4, rsync 42.2
    int lockstep = 0;
    if (const char *v = getenv("PDG_LOCKSTEP")) lockstep = atoi(v);
    if (wpc <= 1 || lockstep < 2 || (lockstep & (lockstep - 1))) lockstep = 0;
    int rsync = (SPEC || (!SS && MODEL != PDG_MODEL_LOGLIN)) ? 1 : 0;
    if (const char *v = getenv("PDG_RSYNC")) rsync = atoi(v);
    if (rsync && wpc > 1 && randomize > 0) lockstep |= 0x10000;
    k_mh_run<MODEL, WMAX, SS, LIGHT, DIAG, SPEC><<<grid, warps * 32, sm, ctx->stream>>>(ctx->P, kind, b, e, R, lockstep, ctx->Z, randomize, per_warp);
    pdg_time_end(ctx, ctx->ev_run);
    ctx->launches++;
    CK(cudaGetLastError());
    return PDG_OK;
}

template <int MODEL, int WMAX, bool SS>
int launch_run_t(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long randomize)
{
    // replay logs and the per-proposal diagnostics go through the DIAG build of the kernel (one warp per chain)
    const bool diag = R.on || ctx->P.mvlog_cap > 0;
    // Two warps per chain pay where the loop body is small enough for two warps per chain to share the
    // instruction cache: chain state in shared memory, no table scans (measured, 1024 chains: p = 50 GGM
    // 495 -> 526 M proposals/s; p = 500 GGM 50 -> 41, log-linear p = 100 43 -> 35).  PDG_SPEC overrides.
    const bool spec = ctx->spec && (ctx->spec_forced || (SS && MODEL != PDG_MODEL_LOGLIN));
    if constexpr (MODEL == PDG_MODEL_LOGLIN) {
        if (ctx->light) {
            if (diag) return launch_run_k<MODEL, WMAX, SS, true, true, false>(ctx, kind, b, e, R, randomize);
            if (spec) {
                const int rc = launch_run_k<MODEL, WMAX, SS, true, false, true>(ctx, kind, b, e, R, randomize);
                if (rc != PDG_SPEC_NO_FIT) return rc;
            }
            return launch_run_k<MODEL, WMAX, SS, true, false, false>(ctx, kind, b, e, R, randomize);
        }
    }
    if (diag) return launch_run_k<MODEL, WMAX, SS, false, true, false>(ctx, kind, b, e, R, randomize);
    if (spec) {
        const int rc = launch_run_k<MODEL, WMAX, SS, false, false, true>(ctx, kind, b, e, R, randomize);
        if (rc != PDG_SPEC_NO_FIT) return rc;
    }
    return launch_run_k<MODEL, WMAX, SS, false, false, false>(ctx, kind, b, e, R, randomize);
}

// The chain state lives in shared memory whenever p * W * 16 <= 8192 bytes (pdg_create): always for
// one- and two-word bitsets (p <= 128), never for more than four words (p >= 257), so only the
// four-word kernels exist in both forms.
template <int MODEL, int WMAX>
int launch_run_w(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long rz)
{
    if constexpr (WMAX <= 2) return launch_run_t<MODEL, WMAX, true>(ctx, kind, b, e, R, rz);
    else if constexpr (WMAX == 8) return launch_run_t<MODEL, WMAX, false>(ctx, kind, b, e, R, rz);
    else {
        if (ctx->smem_state) return launch_run_t<MODEL, WMAX, true>(ctx, kind, b, e, R, rz);
        return launch_run_t<MODEL, WMAX, false>(ctx, kind, b, e, R, rz);
    }
}

template <int MODEL>
int launch_run_m(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long rz)
{
    switch (ctx->wmax) {
    case 1: return launch_run_w<MODEL, 1>(ctx, kind, b, e, R, rz);
    case 2: return launch_run_w<MODEL, 2>(ctx, kind, b, e, R, rz);
    case 4: return launch_run_w<MODEL, 4>(ctx, kind, b, e, R, rz);
    default: return launch_run_w<MODEL, 8>(ctx, kind, b, e, R, rz);
    }
}

template <int MODEL, int WMAX>
int launch_init_t(pdg_ctx *ctx, int kind)
{
    const InitSmem L = init_smem_layout(ctx->P.p, ctx->P.W, tri_bytes_for(ctx->P.ll_cap));
    int rc;
    if ((rc = set_smem(ctx, k_init_logl<MODEL, WMAX>, L.total))) return rc;
    k_init_logl<MODEL, WMAX><<<(unsigned)ctx->P.n_chains, 32, L.total, ctx->stream>>>(ctx->P, kind);
    ctx->launches++;
    CK(cudaGetLastError());
    return PDG_OK;
}

template <int MODEL>
int launch_init_m(pdg_ctx *ctx, int kind)
{
    switch (ctx->wmax) {
    case 1: return launch_init_t<MODEL, 1>(ctx, kind);
    case 2: return launch_init_t<MODEL, 2>(ctx, kind);
    case 4: return launch_init_t<MODEL, 4>(ctx, kind);
    default: return launch_init_t<MODEL, 8>(ctx, kind);
    }
}

template <int MODEL, int WMAX>
int launch_score_t(pdg_ctx *ctx, const u64 *d_bits, long long n_sets, double *d_out, int *d_err)
{
    long long grid = n_sets < ctx->P.n_chains ? n_sets : ctx->P.n_chains;
    int group = (int)((n_sets + grid - 1) / grid);
    group = group < 1 ? 1 : group > 32 ? 32 : group;
    if (const char *v = getenv("PDG_SCORE_GROUP")) { const int x = atoi(v); if (x >= 1 && x <= 32) group = x; }
    grid = (n_sets + group - 1) / group < grid ? (n_sets + group - 1) / group : grid;
#ifdef PDG_LANE_SMEM
    const int sm = PDG_KTRI * 32 * 8 + PDG_KT * 32 * 2 + ((ctx->P.p + 2) * 2 + 15) / 16 * 16;
#else
    const int sm = tri_bytes_for(ctx->P.ll_cap) + PDG_KT * 32 * 2 + ((ctx->P.p + 2) * 2 + 15) / 16 * 16;
#endif
    int *d_list = nullptr, *d_count = nullptr;
    int big_kmax = 0;
    const int big_bytes = 220 * 1024;
    if constexpr (MODEL == PDG_MODEL_GGM) {
        // sets above 32 vertices: listed by the warp kernel, scored one CTA each (TMA-staged rows, blocked LDL^T)
        if (ctx->P.p >= PDG_BIG_MIN && !getenv("PDG_NO_BIG_KERNEL")) {
            if (n_sets + 2 > ctx->big_cap) {
                if (ctx->d_big_list) cudaFree(ctx->d_big_list);
                ctx->d_big_list = nullptr; ctx->big_cap = 0;
                CK(cudaMalloc((void **)&ctx->d_big_list, (size_t)(n_sets + 2) * sizeof(int)));
                ctx->big_cap = n_sets + 2;
            }
            d_count = ctx->d_big_list; d_list = ctx->d_big_list + 2;
            CK(cudaMemsetAsync(d_count, 0, 2 * sizeof(int), ctx->stream));
            big_kmax = big_smem_layout(ctx->P.p, big_bytes).kmax;
            if (big_kmax > ctx->P.kbig) big_kmax = ctx->P.kbig;
        }
    }
    k_score_sets<MODEL, WMAX><<<(unsigned)grid, 32, sm, ctx->stream>>>(ctx->P, d_bits, n_sets, d_out, d_err, d_list, d_count, big_kmax, group);
    ctx->launches++;
    CK(cudaGetLastError());
    if constexpr (MODEL == PDG_MODEL_GGM) {
        if (d_list) {
            int hb[2] = {0, 0};
            CK(cudaMemcpyAsync(hb, d_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            const int n_big = hb[0];
            if (n_big > 0) {
                // shared memory for the largest set of THIS batch: sets of <= ~100 vertices leave room for
                // several CTAs per SM
                int bytes = big_smem_bytes_for(ctx->P.p, hb[1]);
                if (bytes > big_bytes) bytes = big_bytes;
                CK(cudaFuncSetAttribute(k_score_big<WMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_bytes));
                int per_sm = (227 * 1024) / (bytes + 1024);
                per_sm = per_sm < 1 ? 1 : per_sm > 8 ? 8 : per_sm;
                const int cap = ctx->sm_count * per_sm;
                const int g = n_big < cap ? n_big : cap;
                k_score_big<WMAX><<<g, 256, bytes, ctx->stream>>>(ctx->P, d_bits, d_list, n_big, d_out, d_err, bytes);
                ctx->launches++;
                CK(cudaGetLastError());
            }
        }
    }
    return PDG_OK;
}

template <int MODEL>
int launch_score_m(pdg_ctx *ctx, const u64 *d_bits, long long n_sets, double *d_out, int *d_err)
{
    switch (ctx->wmax) {
    case 1: return launch_score_t<MODEL, 1>(ctx, d_bits, n_sets, d_out, d_err);
    case 2: return launch_score_t<MODEL, 2>(ctx, d_bits, n_sets, d_out, d_err);
    case 4: return launch_score_t<MODEL, 4>(ctx, d_bits, n_sets, d_out, d_err);
    default: return launch_score_t<MODEL, 8>(ctx, d_bits, n_sets, d_out, d_err);
    }
}


}  // namespace

#if defined(PDG_PHASE_TIMERS) && defined(PDG_TU_IS_LOGLIN)
// developer probe (PDG_PHASE_TIMERS builds only): scans and cycles by set size, see loglin_scan
extern "C" int pdg_debug_loglin_sp(unsigned long long *out)
{
    return cudaMemcpyFromSymbol(out, g_sp, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -2;
}
extern "C" int pdg_debug_loglin_kh(unsigned long long *out)
{
    return cudaMemcpyFromSymbol(out, g_kh, sizeof(unsigned long long) * 64) == cudaSuccess ? 0 : -2;
}
#endif

#if defined(PDG_PHASE_TIMERS) && defined(PDG_TU_IS_GGM)
extern "C" int pdg_debug_ggm_rz(unsigned long long *out)
{
    return cudaMemcpyFromSymbol(out, g_rz, sizeof(unsigned long long) * 8) == cudaSuccess ? 0 : -2;
}
extern "C" int pdg_debug_ggm_kg(unsigned long long *out)
{
    return cudaMemcpyFromSymbol(out, g_kg, sizeof(unsigned long long) * 128) == cudaSuccess ? 0 : -2;
}
#endif

#define PDG_CAT2(a, b) a##b
#define PDG_CAT(a, b) PDG_CAT2(a, b)

int PDG_CAT(pdg_launch_run_, PDG_TU_NAME)(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long rz)
{
    return launch_run_m<PDG_TU_MODEL>(ctx, kind, b, e, R, rz);
}
int PDG_CAT(pdg_launch_init_, PDG_TU_NAME)(pdg_ctx *ctx, int kind) { return launch_init_m<PDG_TU_MODEL>(ctx, kind); }
int PDG_CAT(pdg_launch_score_, PDG_TU_NAME)(pdg_ctx *ctx, const u64 *d_bits, long long n_sets, double *d_out, int *d_err)
{
    return launch_score_m<PDG_TU_MODEL>(ctx, d_bits, n_sets, d_out, d_err);
}
