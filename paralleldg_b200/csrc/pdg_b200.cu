// libpdg_b200.so: C-ABI (include/pdg_b200.h) over the sm_100a kernels of the parallelDG
// parallel Metropolis-Hastings hot path.  No torch types cross this boundary.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "pdg_run.cuh"
#include "pdg_tally.cuh"

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
    unsigned long long wk[8] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
    double v = 0.0;
    const int e = warp_scores<MODEL, WMAX>(P, ws, ls, slot, lane, X, lane == 0, v, wk);
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
__global__ void __launch_bounds__(32) k_score_sets(PdgParams P, const u64 *bits, long long n_sets, double *out, int *err)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int W = P.W, lane = threadIdx.x;
    WarpScratch ws;
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
    unsigned long long wk[8] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
    for (long long s0 = (long long)blockIdx.x * 32; s0 < n_sets; s0 += (long long)gridDim.x * 32) {
        const long long s = s0 + lane;
        u64 X[WMAX];
        int k = 0;
#pragma unroll
        for (int w = 0; w < WMAX; ++w) { X[w] = (s < n_sets && w < W) ? bits[(size_t)s * W + w] : 0ull; k += __popcll(X[w]); }
        double sc = 0.0;
        int e = 0;
        if (MODEL != PDG_MODEL_UNIFORM) e = warp_scores<MODEL, WMAX>(P, ws, ls, slot, lane, X, s < n_sets && k > 0, sc, wk);
        if (e) atomicOr(err, e);
        if (s < n_sets) out[s] = sc;
        __syncwarp();
    }
}

__global__ void k_memo_clear(MemoTable T)
{
    const size_t n = (size_t)T.mask + 1, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (size_t i = i0; i < n; i += st) {
        if (T.W == 1) T.kv[i] = make_ulonglong2(0ull, MEMO_SENT);
        else { T.tag[i] = 0ull; T.val[i] = MEMO_SENT; }
    }
    // multi-word keys: the key words are cleared as well.  A writer that finds its own tag in a slot
    // publishes a value there once the key words match; with words left over from an earlier run
    // (and two keys sharing a tag) that match could be against a key the slot no longer holds.
    if (T.W > 1) for (size_t i = i0; i < n * (size_t)T.W; i += st) T.keyw[i] = 0ull;
}

// trajectory compaction: exclusive scan of the per-chain record counts, then a packed copy
__global__ void __launch_bounds__(1024) k_scan_counts(const long long *nrec, long long rec_cap, long long n, long long *offsets)
{
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const long long chunk = (n + 1023) / 1024, b = tid * chunk, e = (b + chunk < n) ? b + chunk : n;
    long long s = 0;
    for (long long i = b; i < e; ++i) s += (nrec[i] < rec_cap ? nrec[i] : rec_cap);
    part[tid] = s;
    __syncthreads();
    if (tid == 0) { long long acc = 0; for (int i = 0; i < 1024; ++i) { const long long t = part[i]; part[i] = acc; acc += t; } offsets[n] = acc; }
    __syncthreads();
    long long acc = part[tid];
    for (long long i = b; i < e; ++i) { offsets[i] = acc; acc += (nrec[i] < rec_cap ? nrec[i] : rec_cap); }
}

__global__ void __launch_bounds__(128) k_compact(PdgParams P, const long long *offsets, pdg_update *out, u64 *out_bits)
{
    const long long chain = blockIdx.x;
    const long long n = offsets[chain + 1] - offsets[chain], o = offsets[chain];
    const int W2 = 2 * P.W;
    const pdg_update *src = P.rec + (size_t)chain * P.rec_cap;
    const u64 *sb = P.rec_bits + (size_t)chain * P.rec_cap * W2;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) out[o + i] = src[i];
    for (long long i = threadIdx.x; i < n * W2; i += blockDim.x) out_bits[(size_t)o * W2 + i] = sb[i];
}

// ---------------------------------------------------------------------------------------------
struct pdg_ctx {
    int device = 0;
    PdgParams P;
    RandScratch Z;
    cudaStream_t stream = 0;
    std::string err;
    bool model_set = false, state_set = false;
    long long launches = 0;
    std::vector<void *> owned;
    // model buffers (replaced on re-set)
    double *d_A = nullptr, *d_D = nullptr, *d_logD = nullptr, *d_gconst = nullptr;
    std::vector<void *> model_owned;      // log-linear buffers, freed on re-set / destroy
    int wmax = 1;                         // compile-time word count the kernels are dispatched on
    int wpc = 8;                          // warps (= chains) per CTA of the MH kernel (8 -> one CTA per SM at 1024 chains)
    bool smem_state = false;
    bool light = false;                   // log-linear model with few rows: the 128-register MH kernel
    int memo_bits = 24;
    // compaction
    long long *d_offsets = nullptr;
    pdg_update *d_packed = nullptr;
    u64 *d_packed_bits = nullptr;
    long long packed_total = -1, packed_cap = 0;
    int *d_since = nullptr;               // edge-tally scratch, kept between calls
    unsigned long long *d_tally = nullptr;
    // replay staging
    std::vector<void *> replay_tmp;
    // CUDA-event timing of the MH and randomisation launches
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_run, ev_rand;
};

static void time_begin(pdg_ctx *ctx, std::vector<std::pair<cudaEvent_t, cudaEvent_t>> &v)
{
    if (!ctx->timing) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, ctx->stream);
    v.push_back(std::make_pair(a, b));
}
static void time_end(pdg_ctx *ctx, std::vector<std::pair<cudaEvent_t, cudaEvent_t>> &v)
{
    if (!ctx->timing) return;
    cudaEventRecord(v.back().second, ctx->stream);
}

static int fail(pdg_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, PDG_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
static cudaError_t dalloc(pdg_ctx *ctx, T **ptr, size_t count)
{
    void *v = nullptr;
    cudaError_t e = cudaMalloc(&v, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
    if (e == cudaSuccess) { ctx->owned.push_back(v); *ptr = (T *)v; }
    return e;
}

template <typename F>
static int set_smem(pdg_ctx *ctx, F f, int bytes)
{
    if (bytes > 48 * 1024) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return PDG_OK;
}

template <int MODEL, int WMAX, bool SS, bool LIGHT>
static int launch_run_k(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long randomize)
{
    const RunSmem L = run_smem_layout(ctx->P.p, ctx->P.W, SS, tri_bytes_for(ctx->P.ll_cap));
    const RandSmem LR = rand_smem_layout(ctx->P.p, ctx->P.W);
    const int per_warp = (randomize > 0 && LR.total > L.total) ? LR.total : L.total;
    const int wpc = ctx->wpc, sm = per_warp * wpc;
    int rc;
    if ((rc = set_smem(ctx, k_mh_run<MODEL, WMAX, SS, LIGHT>, sm))) return rc;
    const unsigned grid = (unsigned)((ctx->P.n_chains + wpc - 1) / wpc);
    time_begin(ctx, ctx->ev_run);
    // chains of one CTA re-align every `lockstep` iterations (power of two, 0 = never)
    int lockstep = 0;
    // (not for log-linear models with many rows: a table scan then costs ~100x an iteration and the
    // other chains of the CTA must not wait for it)
    if (wpc > 1 && !(MODEL == PDG_MODEL_LOGLIN && ctx->P.nrows > 8192)) {
        lockstep = 16;
        if (const char *v = getenv("PDG_LOCKSTEP")) lockstep = atoi(v);
        if (lockstep & (lockstep - 1)) lockstep = 0;
    }
    k_mh_run<MODEL, WMAX, SS, LIGHT><<<grid, wpc * 32, sm, ctx->stream>>>(ctx->P, kind, b, e, R, lockstep, ctx->Z, randomize, per_warp);
    time_end(ctx, ctx->ev_run);
    ctx->launches++;
    CK(cudaGetLastError());
    return PDG_OK;
}

template <int MODEL, int WMAX, bool SS>
static int launch_run_t(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long randomize)
{
    if constexpr (MODEL == PDG_MODEL_LOGLIN) {
        if (ctx->light) return launch_run_k<MODEL, WMAX, SS, true>(ctx, kind, b, e, R, randomize);
    }
    return launch_run_k<MODEL, WMAX, SS, false>(ctx, kind, b, e, R, randomize);
}

template <int MODEL, int WMAX>
static int launch_run_w(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long rz)
{
    if (ctx->smem_state) return launch_run_t<MODEL, WMAX, true>(ctx, kind, b, e, R, rz);
    return launch_run_t<MODEL, WMAX, false>(ctx, kind, b, e, R, rz);
}

template <int MODEL>
static int launch_run_m(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long rz)
{
    switch (ctx->wmax) {
    case 1: return launch_run_w<MODEL, 1>(ctx, kind, b, e, R, rz);
    case 2: return launch_run_w<MODEL, 2>(ctx, kind, b, e, R, rz);
    case 4: return launch_run_w<MODEL, 4>(ctx, kind, b, e, R, rz);
    default: return launch_run_w<MODEL, 8>(ctx, kind, b, e, R, rz);
    }
}

// randomize > 0: the kernel randomises every chain's junction tree at every multiple of it
static int run_dispatch(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long randomize = 0)
{
    switch (ctx->P.model) {
    case PDG_MODEL_GGM: return launch_run_m<PDG_MODEL_GGM>(ctx, kind, b, e, R, randomize);
    case PDG_MODEL_LOGLIN: return launch_run_m<PDG_MODEL_LOGLIN>(ctx, kind, b, e, R, randomize);
    default: return launch_run_m<PDG_MODEL_UNIFORM>(ctx, kind, b, e, R, randomize);
    }
}

template <int MODEL, int WMAX>
static int launch_init_t(pdg_ctx *ctx, int kind)
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
static int launch_init_m(pdg_ctx *ctx, int kind)
{
    switch (ctx->wmax) {
    case 1: return launch_init_t<MODEL, 1>(ctx, kind);
    case 2: return launch_init_t<MODEL, 2>(ctx, kind);
    case 4: return launch_init_t<MODEL, 4>(ctx, kind);
    default: return launch_init_t<MODEL, 8>(ctx, kind);
    }
}

template <int MODEL, int WMAX>
static int launch_score_t(pdg_ctx *ctx, const u64 *d_bits, long long n_sets, double *d_out, int *d_err)
{
    long long grid = n_sets < ctx->P.n_chains ? n_sets : ctx->P.n_chains;
#ifdef PDG_LANE_SMEM
    const int sm = PDG_KTRI * 32 * 8 + PDG_KT * 32 * 2 + ((ctx->P.p + 2) * 2 + 15) / 16 * 16;
#else
    const int sm = tri_bytes_for(ctx->P.ll_cap) + PDG_KT * 32 * 2 + ((ctx->P.p + 2) * 2 + 15) / 16 * 16;
#endif
    k_score_sets<MODEL, WMAX><<<(unsigned)grid, 32, sm, ctx->stream>>>(ctx->P, d_bits, n_sets, d_out, d_err);
    ctx->launches++;
    CK(cudaGetLastError());
    return PDG_OK;
}

template <int MODEL>
static int launch_score_m(pdg_ctx *ctx, const u64 *d_bits, long long n_sets, double *d_out, int *d_err)
{
    switch (ctx->wmax) {
    case 1: return launch_score_t<MODEL, 1>(ctx, d_bits, n_sets, d_out, d_err);
    case 2: return launch_score_t<MODEL, 2>(ctx, d_bits, n_sets, d_out, d_err);
    case 4: return launch_score_t<MODEL, 4>(ctx, d_bits, n_sets, d_out, d_err);
    default: return launch_score_t<MODEL, 8>(ctx, d_bits, n_sets, d_out, d_err);
    }
}

static int memo_clear(pdg_ctx *ctx)
{
    k_memo_clear<<<296, 256, 0, ctx->stream>>>(ctx->P.memo);
    ctx->launches++;
    CK(cudaGetLastError());
    return PDG_OK;
}

template <typename T>
static int stage(pdg_ctx *ctx, const T *src, size_t count, const T **dst)
{
    *dst = nullptr;
    if (!src || count == 0) return PDG_OK;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeDevice) { *dst = src; return PDG_OK; }
    cudaGetLastError();
    void *d = nullptr;
    CK(cudaMalloc(&d, count * sizeof(T)));
    ctx->replay_tmp.push_back(d);
    CK(cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *dst = (const T *)d;
    return PDG_OK;
}

extern "C" {

int pdg_abi_version(void) { return PDG_ABI_VERSION; }

const char *pdg_last_error(pdg_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int64_t pdg_launch_count(pdg_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pdg_create(pdg_ctx **out, int device, int p, int64_t n_chains, uint32_t chain0, uint64_t seed,
               int64_t n_samples, int64_t rec_cap)
{
    if (!out) return PDG_E_ARG;
    *out = nullptr;
    if (p < 2 || p > 64 * PDG_WMAX_LIMIT || n_chains < 1 || n_samples < 1 || rec_cap < 0) return PDG_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PDG_E_CUDA;   // no CPU fallback
    if (device < 0 || device >= ndev) return PDG_E_ARG;
    pdg_ctx *ctx = new pdg_ctx();
    ctx->device = device;
    ctx->wpc = n_chains >= 8 ? 8 : (int)n_chains;      // the last CTA may be partial (its barrier counts live warps)
    if (const char *v = getenv("PDG_WPC")) { const int x = atoi(v); if (x >= 1 && x <= 8) ctx->wpc = x; }
    // 16M slots (256 MB) for one- and two-word keys: the p = 50 hit rate saturates at 2^23; wider
    // keys come with far more distinct cliques (p = 500: 21M per run), 32M slots = 2.7 GB there
    ctx->memo_bits = (p > 128) ? 25 : 24;
    if (const char *v = getenv("PDG_MEMO_BITS")) { const int x = atoi(v); if (x >= 8 && x <= 26) ctx->memo_bits = x; }
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return PDG_E_CUDA; }
    PdgParams &P = ctx->P;
    memset(&P, 0, sizeof(P));
    P.p = p; P.W = (p + 63) / 64;
    P.model = PDG_MODEL_UNIFORM; P.prior = PDG_PRIOR_MBC; P.prior_a = 2.0; P.prior_b = 4.0;
    P.n_chains = n_chains; P.chain0 = chain0; P.seed_lo = (u32)seed; P.seed_hi = (u32)(seed >> 32);
    P.n_samples = n_samples; P.rec_cap = rec_cap;
    P.memo_small = getenv("PDG_NO_MEMO_SMALL") ? 0 : 1;
    const size_t C = (size_t)n_chains, pw = (size_t)p * P.W;
    cudaError_t e = cudaSuccess;
#define TRY(x) if (e == cudaSuccess) e = (x)
    TRY(dalloc(ctx, &P.M, C * pw));
    TRY(dalloc(ctx, &P.N, C * pw));
    TRY(dalloc(ctx, &P.edges, C * 2 * (p - 1)));
    TRY(dalloc(ctx, &P.csr_off, C * (p + 1)));
    TRY(dalloc(ctx, &P.csr_adj, C * 2 * p));
    TRY(dalloc(ctx, &P.logl, C * (size_t)n_samples));
    TRY(dalloc(ctx, &P.kcount, C));
    TRY(dalloc(ctx, &P.nrec, C));
    TRY(dalloc(ctx, &P.err, C));
    TRY(dalloc(ctx, &P.work, C * 16));
    TRY(dalloc(ctx, &P.G0, C * pw));
    TRY(dalloc(ctx, &P.rec, C * (size_t)rec_cap));
    TRY(dalloc(ctx, &P.rec_bits, C * (size_t)rec_cap * 2 * P.W));
    TRY(dalloc(ctx, &ctx->Z.adj, C * pw));
    TRY(dalloc(ctx, &ctx->Z.K, C * pw));
    TRY(dalloc(ctx, &ctx->Z.sep, C * pw));
    TRY(dalloc(ctx, &ctx->d_offsets, C + 1));
    // scratch for cliques above the 32-row shared-memory triangle
    P.kbig = p + 1 < 256 ? p + 1 : 256;
    if (P.kbig > 32) TRY(dalloc(ctx, &P.big, C * ((size_t)(P.kbig + 1) * (P.kbig + 2) / 2)));
    else { P.kbig = 0; P.big = nullptr; }
    TRY(cudaMemset(P.err, 0, C * sizeof(int)));
    TRY(cudaMemset(P.work, 0, C * 16 * sizeof(unsigned long long)));
    {
        const int W = P.W;
        ctx->wmax = W <= 1 ? 1 : W <= 2 ? 2 : W <= 4 ? 4 : 8;
        ctx->smem_state = (size_t)p * W * 16 <= 8192;
        MemoTable &T = P.memo;
        T.W = (ctx->wmax == 1) ? 1 : W;
        T.mask = (1u << ctx->memo_bits) - 1u;
        const size_t slots = (size_t)1 << ctx->memo_bits;
        T.kv = nullptr; T.tag = nullptr; T.keyw = nullptr; T.val = nullptr;
        if (ctx->wmax == 1) { TRY(dalloc(ctx, &T.kv, slots)); }
        else { TRY(dalloc(ctx, &T.tag, slots)); TRY(dalloc(ctx, &T.val, slots)); TRY(dalloc(ctx, &T.keyw, slots * W)); }
    }
    TRY(cudaMemset(P.kcount, 0, C * sizeof(long long)));
    TRY(cudaMemset(P.nrec, 0, C * sizeof(long long)));
    TRY(cudaMemset(P.logl, 0, C * (size_t)n_samples * sizeof(double)));
#undef TRY
    if (e != cudaSuccess) {
        for (void *v : ctx->owned) cudaFree(v);
        delete ctx;
        return e == cudaErrorMemoryAllocation ? PDG_E_NOMEM : PDG_E_CUDA;
    }
    *out = ctx;
    return PDG_OK;
}

int pdg_destroy(pdg_ctx *ctx)
{
    if (!ctx) return PDG_E_ARG;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (void *v : ctx->owned) cudaFree(v);
    for (void *v : ctx->replay_tmp) cudaFree(v);
    if (ctx->d_A) cudaFree(ctx->d_A);
    if (ctx->d_D) cudaFree(ctx->d_D);
    if (ctx->d_logD) cudaFree(ctx->d_logD);
    if (ctx->d_gconst) cudaFree(ctx->d_gconst);
    for (void *v : ctx->model_owned) cudaFree(v);
    if (ctx->d_since) { cudaFree(ctx->d_since); cudaFree(ctx->d_tally); }
    if (ctx->d_packed) cudaFree(ctx->d_packed);
    if (ctx->d_packed_bits) cudaFree(ctx->d_packed_bits);
    if (ctx->P.ml_dlik) { cudaFree(ctx->P.ml_dlik); cudaFree(ctx->P.ml_dprior); cudaFree(ctx->P.ml_logq); cudaFree(ctx->P.ml_u); cudaFree(ctx->P.ml_acc); cudaFree(ctx->P.ml_U); }
    delete ctx;
    return PDG_OK;
}

int pdg_set_seed(pdg_ctx *ctx, uint64_t seed, uint32_t chain0)
{
    if (!ctx) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    P.seed_lo = (u32)seed; P.seed_hi = (u32)(seed >> 32); P.chain0 = chain0;
    const size_t C = (size_t)P.n_chains;
    CK(cudaMemsetAsync(P.err, 0, C * sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(P.kcount, 0, C * sizeof(long long), ctx->stream));
    CK(cudaMemsetAsync(P.nrec, 0, C * sizeof(long long), ctx->stream));
    ctx->packed_total = -1;
    ctx->state_set = false;
    return PDG_OK;
}

int pdg_set_stream(pdg_ctx *ctx, void *stream)
{
    if (!ctx) return PDG_E_ARG;
    ctx->stream = (cudaStream_t)stream;
    return PDG_OK;
}

int pdg_set_prior(pdg_ctx *ctx, int prior_kind, double a, double b)
{
    if (!ctx) return PDG_E_ARG;
    if (prior_kind < 0 || prior_kind > 3) return fail(ctx, PDG_E_ARG, "unknown prior kind");
    ctx->P.prior = prior_kind; ctx->P.prior_a = a; ctx->P.prior_b = b;
    return PDG_OK;
}

int pdg_set_model_uniform(pdg_ctx *ctx)
{
    if (!ctx) return PDG_E_ARG;
    ctx->P.model = PDG_MODEL_UNIFORM; ctx->P.ll_cap = 0;
    ctx->model_set = true;
    return PDG_OK;
}

int pdg_set_model_ggm(pdg_ctx *ctx, const void *A, const void *D, int64_t n, double delta, const void *gconst)
{
    if (!ctx || !A || !D || !gconst) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    const size_t pp = (size_t)P.p * P.p;
    if (!ctx->d_A) {
        CK(cudaMalloc((void **)&ctx->d_A, pp * 8));
        CK(cudaMalloc((void **)&ctx->d_D, pp * 8));
        CK(cudaMalloc((void **)&ctx->d_logD, (size_t)P.p * 8));
        CK(cudaMalloc((void **)&ctx->d_gconst, (size_t)(P.p + 2) * 8));
    }
    CK(cudaMemcpyAsync(ctx->d_A, A, pp * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_D, D, pp * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_gconst, gconst, (size_t)(P.p + 1) * 8, cudaMemcpyDefault, ctx->stream));
    // classify D (identity / diagonal / general) on the host: one-off, p*p doubles
    std::vector<double> hD(pp), hl(P.p);
    CK(cudaMemcpyAsync(hD.data(), ctx->d_D, pp * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    bool diag = true, ident = true;
    for (int i = 0; i < P.p; ++i)
        for (int j = 0; j < P.p; ++j) {
            const double v = hD[(size_t)i * P.p + j];
            if (i != j && v != 0.0) diag = false;
            if (i == j && v != 1.0) ident = false;
        }
    P.dmode = diag ? (ident ? 0 : 1) : 2;
    for (int i = 0; i < P.p; ++i) hl[i] = log(hD[(size_t)i * P.p + i]);
    CK(cudaMemcpyAsync(ctx->d_logD, hl.data(), (size_t)P.p * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    P.A = ctx->d_A; P.Dm = ctx->d_D; P.logDdiag = ctx->d_logD; P.gconst = ctx->d_gconst;
    P.delta = delta; P.nobs = (double)n;
    P.model = PDG_MODEL_GGM; P.ll_cap = 0;
    ctx->model_set = true;
    return PDG_OK;
}

int pdg_set_model_loglin(pdg_ctx *ctx, const void *data, const void *levels, int64_t nrows, double cell_alpha,
                         int n_ktab, const void *ktab, const void *alphatab, const void *lgtab)
{
    if (!ctx || !data || !levels || nrows < 1 || n_ktab < 0) return PDG_E_ARG;
    if (n_ktab > 0 && (!ktab || !alphatab || !lgtab)) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    for (void *v : ctx->model_owned) cudaFree(v);
    ctx->model_owned.clear();
    auto own = [&](size_t bytes, void **out) -> cudaError_t {
        cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
        if (e == cudaSuccess) ctx->model_owned.push_back(*out);
        return e;
    };
    void *d_data = nullptr, *d_lv = nullptr, *d_k = nullptr, *d_a = nullptr, *d_lg = nullptr;
    // columns padded to a 16-byte stride, an all-zero column behind the last one and 4 KB of slack:
    // the dense scan (pdg_loglin.cuh) then needs no load predicates
    const size_t ld = ((size_t)nrows + 15) & ~(size_t)15;
    CK(own((size_t)(P.p + 1) * ld + 4096, &d_data));
    CK(own((size_t)P.p * sizeof(int), &d_lv));
    CK(cudaMemsetAsync(d_data, 0, (size_t)(P.p + 1) * ld + 4096, ctx->stream));
    CK(cudaMemcpy2DAsync(d_data, ld, data, (size_t)nrows, (size_t)nrows, (size_t)P.p, cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync(d_lv, levels, (size_t)P.p * sizeof(int), cudaMemcpyDefault, ctx->stream));
    if (n_ktab > 0) {
        CK(own((size_t)n_ktab * 8, &d_k));
        CK(own((size_t)n_ktab * 8, &d_a));
        CK(own((size_t)n_ktab * (size_t)(nrows + 2) * 8, &d_lg));
        CK(cudaMemcpyAsync(d_k, ktab, (size_t)n_ktab * 8, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync(d_a, alphatab, (size_t)n_ktab * 8, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync(d_lg, lgtab, (size_t)n_ktab * (size_t)(nrows + 2) * 8, cudaMemcpyDefault, ctx->stream));
    }
    // dense cell counters: one table per warp slot (= chain); larger sets go through a pool of
    // hashed tables
    int hist_cells = 1 << 16;
    if (const char *v = getenv("PDG_HIST_CELLS")) { const int x = atoi(v); if (x >= 2 && x <= (1 << 16)) hist_cells = x; }
    void *d_hist = nullptr;
    CK(own((size_t)P.n_chains * hist_cells * 4, &d_hist));
    CK(cudaMemsetAsync(d_hist, 0, (size_t)P.n_chains * hist_cells * 4, ctx->stream));
    LoglinSparse sp;
    sp.nsp = (int)(P.n_chains < 64 ? P.n_chains : 64);
    int hs = 1;
    while (hs < 2 * nrows) hs <<= 1;
    sp.hslots = hs;
    void *d_lock = nullptr, *d_hk = nullptr, *d_hc = nullptr, *d_cc = nullptr;
    CK(own((size_t)sp.nsp * 4, &d_lock));
    CK(own((size_t)sp.nsp * hs * 8, &d_hk));
    CK(own((size_t)sp.nsp * hs * 4, &d_hc));
    CK(own((size_t)sp.nsp * (size_t)(nrows + 2) * 4, &d_cc));
    CK(cudaMemsetAsync(d_lock, 0, (size_t)sp.nsp * 4, ctx->stream));
    CK(cudaMemsetAsync(d_hk, 0, (size_t)sp.nsp * hs * 8, ctx->stream));
    CK(cudaMemsetAsync(d_hc, 0, (size_t)sp.nsp * hs * 4, ctx->stream));
    CK(cudaMemsetAsync(d_cc, 0, (size_t)sp.nsp * (size_t)(nrows + 2) * 4, ctx->stream));
    sp.lock = (int *)d_lock; sp.hkeys = (u64 *)d_hk; sp.hcnt = (unsigned int *)d_hc; sp.cc = (unsigned int *)d_cc;
    CK(cudaStreamSynchronize(ctx->stream));
    P.data = (const unsigned char *)d_data; P.data_ld = (long long)ld; P.levels = (const int *)d_lv; P.nrows = nrows; P.cell_alpha = cell_alpha;
    P.n_ktab = n_ktab; P.ktab = (const double *)d_k; P.alphatab = (const double *)d_a; P.lgtab = (const double *)d_lg;
    P.hist = (unsigned int *)d_hist; P.hist_cells = hist_cells; P.sp = sp;
    P.model = PDG_MODEL_LOGLIN;
    // few rows: a scan is cheap, so the light kernel (two CTAs per SM, 1024-cell tables) runs the chains
    ctx->light = nrows <= 8192;
    if (const char *v = getenv("PDG_LIGHT")) ctx->light = atoi(v) != 0;
    // per-warp shared count table: 4096 cells when a CTA of the MH kernel still fits in shared memory
    P.ll_cap = ctx->light ? 1024 : 4096;
    if (const char *v = getenv("PDG_LL_CAP")) { const int x = atoi(v); if (x >= 1024 && x <= 8192 && !(x & (x - 1))) P.ll_cap = x; }
    for (;;) {
        const RunSmem L = run_smem_layout(P.p, P.W, ctx->smem_state, tri_bytes_for(P.ll_cap));
        const RandSmem LR = rand_smem_layout(P.p, P.W);
        const int per_warp = LR.total > L.total ? LR.total : L.total;
        if ((long long)per_warp * ctx->wpc <= 200 * 1024 || P.ll_cap <= 1024) break;
        P.ll_cap >>= 1;
    }
    ctx->model_set = true;
    return PDG_OK;
}

static int launch_derived(pdg_ctx *ctx, long long b, long long n)
{
    const int sm = (ctx->P.p + 2) * 4 + 64;
    k_build_derived<<<(unsigned)n, 32, sm, ctx->stream>>>(ctx->P, b, n);
    ctx->launches++;
    CK(cudaGetLastError());
    return PDG_OK;
}

int pdg_set_state(pdg_ctx *ctx, int64_t chain_begin, int64_t chain_count, const void *edges, const void *bits)
{
    if (!ctx || !edges || !bits) return PDG_E_ARG;
    PdgParams &P = ctx->P;
    if (chain_begin < 0 || chain_count < 1 || chain_begin + chain_count > P.n_chains) return fail(ctx, PDG_E_ARG, "chain range");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(P.edges + (size_t)chain_begin * 2 * (P.p - 1), edges, (size_t)chain_count * 2 * (P.p - 1) * sizeof(int), cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync(P.M + (size_t)chain_begin * P.p * P.W, bits, (size_t)chain_count * P.p * P.W * 8, cudaMemcpyDefault, ctx->stream));
    ctx->state_set = true;
    return launch_derived(ctx, chain_begin, chain_count);
}

int pdg_get_state(pdg_ctx *ctx, int64_t chain_begin, int64_t chain_count, void *edges, void *bits)
{
    if (!ctx) return PDG_E_ARG;
    PdgParams &P = ctx->P;
    if (chain_begin < 0 || chain_count < 1 || chain_begin + chain_count > P.n_chains) return fail(ctx, PDG_E_ARG, "chain range");
    CK(cudaSetDevice(ctx->device));
    if (edges) CK(cudaMemcpyAsync(edges, P.edges + (size_t)chain_begin * 2 * (P.p - 1), (size_t)chain_count * 2 * (P.p - 1) * sizeof(int), cudaMemcpyDefault, ctx->stream));
    if (bits) CK(cudaMemcpyAsync(bits, P.M + (size_t)chain_begin * P.p * P.W, (size_t)chain_count * P.p * P.W * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PDG_OK;
}

int pdg_set_cliques(pdg_ctx *ctx, const void *bits, int n_cliques)
{
    if (!ctx) return PDG_E_ARG;
    PdgParams &P = ctx->P;
    const int p = P.p, W = P.W;
    if (bits && (n_cliques < 1 || n_cliques > p)) return fail(ctx, PDG_E_ARG, "n_cliques");
    CK(cudaSetDevice(ctx->device));
    std::vector<u64> row((size_t)p * W, 0ull);
    if (bits) {
        CK(cudaMemcpy(row.data(), bits, (size_t)n_cliques * W * 8, cudaMemcpyDefault));
    } else {
        for (int v = 0; v < p; ++v) row[(size_t)v * W + (v >> 6)] = 1ull << (v & 63);
    }
    std::vector<int> ed((size_t)2 * (p - 1));
    for (int e = 0; e < p - 1; ++e) { ed[2 * e] = e; ed[2 * e + 1] = e + 1; }
    // replicate to every chain (device-side broadcast from chain 0)
    CK(cudaMemcpyAsync(P.M, row.data(), (size_t)p * W * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(P.edges, ed.data(), (size_t)2 * (p - 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    size_t have = 1;
    while (have < (size_t)P.n_chains) {
        const size_t n = (have < (size_t)P.n_chains - have) ? have : (size_t)P.n_chains - have;
        CK(cudaMemcpyAsync(P.M + have * p * W, P.M, n * p * W * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(P.edges + have * 2 * (p - 1), P.edges, n * 2 * (p - 1) * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        have += n;
    }
    ctx->state_set = true;
    return launch_derived(ctx, 0, P.n_chains);
}

int pdg_randomize(pdg_ctx *ctx, uint32_t iter)
{
    if (!ctx) return PDG_E_ARG;
    if (!ctx->state_set) return fail(ctx, PDG_E_STATE, "state not set");
    CK(cudaSetDevice(ctx->device));
    const RandSmem L = rand_smem_layout(ctx->P.p, ctx->P.W);
    if (L.total > 48 * 1024) CK(cudaFuncSetAttribute(k_randomize, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    time_begin(ctx, ctx->ev_rand);
    k_randomize<<<(unsigned)ctx->P.n_chains, 32, L.total, ctx->stream>>>(ctx->P, ctx->Z, iter);
    time_end(ctx, ctx->ev_rand);
    ctx->launches++;
    CK(cudaGetLastError());
    return PDG_OK;
}

int pdg_init_logl(pdg_ctx *ctx, int kind)
{
    if (!ctx) return PDG_E_ARG;
    if (!ctx->state_set || !ctx->model_set) return fail(ctx, PDG_E_STATE, "model/state not set");
    CK(cudaSetDevice(ctx->device));
    int rc;
    // a run starts with an empty memo: nothing scored in an earlier run is reused
    if ((rc = memo_clear(ctx))) return rc;
    switch (ctx->P.model) {
    case PDG_MODEL_GGM: rc = launch_init_m<PDG_MODEL_GGM>(ctx, kind); break;
    case PDG_MODEL_LOGLIN: rc = launch_init_m<PDG_MODEL_LOGLIN>(ctx, kind); break;
    default: rc = launch_init_m<PDG_MODEL_UNIFORM>(ctx, kind); break;
    }
    if (rc) return rc;
    k_snapshot_graph<<<(unsigned)ctx->P.n_chains, 32, 0, ctx->stream>>>(ctx->P);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->packed_total = -1;
    return PDG_OK;
}

int pdg_run(pdg_ctx *ctx, int kind, int64_t it_begin, int64_t it_end, const pdg_replay *replay)
{
    if (!ctx) return PDG_E_ARG;
    if (!ctx->state_set || !ctx->model_set) return fail(ctx, PDG_E_STATE, "model/state not set");
    if (kind != PDG_KIND_PARALLEL && kind != PDG_KIND_SINGLE) return fail(ctx, PDG_E_ARG, "kernel kind");
    if (it_begin < 1 || it_end > ctx->P.n_samples || it_begin > it_end) return fail(ctx, PDG_E_ARG, "iteration range");
    CK(cudaSetDevice(ctx->device));
    PdgReplayDev R;
    memset(&R, 0, sizeof(R));
    if (replay) {
        const size_t C = (size_t)ctx->P.n_chains, ns = (size_t)replay->n_samples;
        int rc;
        R.on = 1; R.n_samples = replay->n_samples;
        if ((rc = stage(ctx, replay->it_node, C * ns, &R.it_node))) return rc;
        if ((rc = stage(ctx, replay->it_mtype, C * ns, &R.it_mtype))) return rc;
        if ((rc = stage(ctx, replay->it_aux, C * ns, &R.it_aux))) return rc;
        if ((rc = stage(ctx, (const long long *)replay->it_off, C * (ns + 1), &R.it_off))) return rc;
        if ((rc = stage(ctx, replay->mv_U, (size_t)replay->n_moves_total, &R.mv_U))) return rc;
        if ((rc = stage(ctx, replay->mv_Uadj, (size_t)replay->n_moves_total, &R.mv_Uadj))) return rc;
        if ((rc = stage(ctx, replay->mv_u, (size_t)replay->n_moves_total, &R.mv_u))) return rc;
    }
    ctx->packed_total = -1;
    int rc = PDG_OK;
    if (it_end > it_begin) rc = run_dispatch(ctx, kind, it_begin, it_end, R);
    if (replay && !ctx->replay_tmp.empty()) {
        CK(cudaStreamSynchronize(ctx->stream));
        for (void *v : ctx->replay_tmp) cudaFree(v);
        ctx->replay_tmp.clear();
    }
    return rc;
}

int pdg_sample(pdg_ctx *ctx, int kind, int64_t randomize)
{
    if (!ctx) return PDG_E_ARG;
    if (randomize < 1) return fail(ctx, PDG_E_ARG, "randomize interval");
    if (kind != PDG_KIND_PARALLEL && kind != PDG_KIND_SINGLE) return fail(ctx, PDG_E_ARG, "kernel kind");
    int rc = pdg_init_logl(ctx, kind);
    if (rc) return rc;
    const long long M = ctx->P.n_samples;
    if (M < 2) return PDG_OK;
    PdgReplayDev R;
    memset(&R, 0, sizeof(R));
    // One launch per randomisation interval (with the randomiser as its own kernel in between) or
    // the whole run fused into ONE launch?  Separate launches re-align the chains every `randomize`
    // iterations, which is what the instruction-fetch-bound GGM / small-table workloads want
    // (p=50 GGM: 286 vs 279 M proposals/s, p=15 log-linear: 1196 vs 632 M/s); but every launch also
    // waits for its slowest chain, and when a memo miss costs ~100 iterations (table scans over
    // many rows) that tail dominates: p=100, n=1e5 log-linear runs 2.4x faster fused.
    bool fused = ctx->P.model == PDG_MODEL_LOGLIN && ctx->P.nrows > 8192;
    if (const char *v = getenv("PDG_FUSED")) fused = atoi(v) != 0;
    if (!fused) {
        long long it = 1;
        while (it < M) {
            if (it % randomize == 0 && (rc = pdg_randomize(ctx, (uint32_t)it))) return rc;
            long long end = (it / randomize + 1) * randomize;
            if (end > M) end = M;
            if ((rc = run_dispatch(ctx, kind, it, end, R))) return rc;
            it = end;
        }
        return PDG_OK;
    }
    // fused: every warp carries its chain through all iterations and randomises its own junction
    // tree every `randomize` iterations
    return run_dispatch(ctx, kind, 1, M, R, randomize);
}

int pdg_sync(pdg_ctx *ctx)
{
    if (!ctx) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<int> h((size_t)ctx->P.n_chains);
    CK(cudaMemcpy(h.data(), ctx->P.err, h.size() * sizeof(int), cudaMemcpyDeviceToHost));
    int any = 0; long long first = -1;
    for (size_t i = 0; i < h.size(); ++i) if (h[i]) { if (first < 0) first = (long long)i; any |= h[i]; }
    if (!any) return PDG_OK;
    char buf[160];
    snprintf(buf, sizeof(buf), "chain %lld: device error flags 0x%x (1 replay mismatch, 2 record overflow, 4 non-SPD pivot, 8 clique too large)", first, any);
    ctx->err = buf;
    if (any & ERR_REPLAY) return PDG_E_REPLAY;
    if (any & (ERR_NUMERIC | ERR_BIGCLIQUE)) return PDG_E_NUMERIC;
    return PDG_E_OVERFLOW;
}

int pdg_get_counts(pdg_ctx *ctx, void *n_proposals, void *n_accepted)
{
    if (!ctx) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t C = (size_t)ctx->P.n_chains;
    if (n_proposals) CK(cudaMemcpyAsync(n_proposals, ctx->P.kcount, C * 8, cudaMemcpyDefault, ctx->stream));
    if (n_accepted) CK(cudaMemcpyAsync(n_accepted, ctx->P.nrec, C * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PDG_OK;
}

int pdg_get_logl(pdg_ctx *ctx, void *logl)
{
    if (!ctx || !logl) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(logl, ctx->P.logl, (size_t)ctx->P.n_chains * ctx->P.n_samples * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PDG_OK;
}

int pdg_compact_updates(pdg_ctx *ctx, int64_t *total, void *offsets)
{
    if (!ctx || !total) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    k_scan_counts<<<1, 1024, 0, ctx->stream>>>(P.nrec, P.rec_cap, P.n_chains, ctx->d_offsets);
    ctx->launches++;
    CK(cudaGetLastError());
    long long tot = 0;
    CK(cudaMemcpyAsync(&tot, ctx->d_offsets + P.n_chains, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (tot > ctx->packed_cap) {
        if (ctx->d_packed) { cudaFree(ctx->d_packed); cudaFree(ctx->d_packed_bits); ctx->d_packed = nullptr; ctx->d_packed_bits = nullptr; }
        CK(cudaMalloc((void **)&ctx->d_packed, (size_t)tot * sizeof(pdg_update)));
        CK(cudaMalloc((void **)&ctx->d_packed_bits, (size_t)tot * 2 * P.W * 8));
        ctx->packed_cap = tot;
    }
    if (tot > 0) {
        k_compact<<<(unsigned)P.n_chains, 128, 0, ctx->stream>>>(P, ctx->d_offsets, ctx->d_packed, ctx->d_packed_bits);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    ctx->packed_total = tot;
    *total = tot;
    if (offsets) CK(cudaMemcpyAsync(offsets, ctx->d_offsets, (size_t)(P.n_chains + 1) * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PDG_OK;
}

int pdg_get_updates(pdg_ctx *ctx, void *updates, void *bits)
{
    if (!ctx) return PDG_E_ARG;
    if (ctx->packed_total < 0) return fail(ctx, PDG_E_STATE, "call pdg_compact_updates first");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->packed_total;
    if (n && updates) CK(cudaMemcpyAsync(updates, ctx->d_packed, n * sizeof(pdg_update), cudaMemcpyDefault, ctx->stream));
    if (n && bits) CK(cudaMemcpyAsync(bits, ctx->d_packed_bits, n * 2 * ctx->P.W * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PDG_OK;
}

int pdg_enable_move_log(pdg_ctx *ctx, int64_t cap)
{
    if (!ctx || cap < 1) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    if (P.ml_dlik) { cudaFree(P.ml_dlik); cudaFree(P.ml_dprior); cudaFree(P.ml_logq); cudaFree(P.ml_u); cudaFree(P.ml_acc); cudaFree(P.ml_U); P.ml_dlik = nullptr; }
    const size_t n = (size_t)P.n_chains * (size_t)cap;
    CK(cudaMalloc((void **)&P.ml_dlik, n * 8));
    CK(cudaMalloc((void **)&P.ml_dprior, n * 8));
    CK(cudaMalloc((void **)&P.ml_logq, n * 8));
    CK(cudaMalloc((void **)&P.ml_u, n * 8));
    CK(cudaMalloc((void **)&P.ml_acc, n * 4));
    CK(cudaMalloc((void **)&P.ml_U, n * 4));
    CK(cudaMemset(P.ml_acc, 0xff, n * 4));
    P.mvlog_cap = cap;
    return PDG_OK;
}

int pdg_get_move_log(pdg_ctx *ctx, void *dlik, void *dprior, void *logq, void *u, void *acc, void *U)
{
    if (!ctx) return PDG_E_ARG;
    PdgParams &P = ctx->P;
    if (!P.mvlog_cap) return fail(ctx, PDG_E_STATE, "move log not enabled");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)P.n_chains * (size_t)P.mvlog_cap;
    if (dlik) CK(cudaMemcpyAsync(dlik, P.ml_dlik, n * 8, cudaMemcpyDefault, ctx->stream));
    if (dprior) CK(cudaMemcpyAsync(dprior, P.ml_dprior, n * 8, cudaMemcpyDefault, ctx->stream));
    if (logq) CK(cudaMemcpyAsync(logq, P.ml_logq, n * 8, cudaMemcpyDefault, ctx->stream));
    if (u) CK(cudaMemcpyAsync(u, P.ml_u, n * 8, cudaMemcpyDefault, ctx->stream));
    if (acc) CK(cudaMemcpyAsync(acc, P.ml_acc, n * 4, cudaMemcpyDefault, ctx->stream));
    if (U) CK(cudaMemcpyAsync(U, P.ml_U, n * 4, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PDG_OK;
}

int pdg_score_sets(pdg_ctx *ctx, const void *bits, int64_t n_sets, void *scores)
{
    if (!ctx || !bits || !scores || n_sets < 0) return PDG_E_ARG;
    if (!ctx->model_set) return fail(ctx, PDG_E_STATE, "model not set");
    if (n_sets == 0) return PDG_OK;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    u64 *d_bits = nullptr; double *d_out = nullptr; int *d_err = nullptr;
    CK(cudaMalloc((void **)&d_bits, (size_t)n_sets * P.W * 8));
    CK(cudaMalloc((void **)&d_out, (size_t)n_sets * 8));
    CK(cudaMalloc((void **)&d_err, 4));
    CK(cudaMemsetAsync(d_err, 0, 4, ctx->stream));
    CK(cudaMemcpyAsync(d_bits, bits, (size_t)n_sets * P.W * 8, cudaMemcpyDefault, ctx->stream));
    int rc;
    switch (P.model) {
    case PDG_MODEL_GGM: rc = launch_score_m<PDG_MODEL_GGM>(ctx, d_bits, n_sets, d_out, d_err); break;
    case PDG_MODEL_LOGLIN: rc = launch_score_m<PDG_MODEL_LOGLIN>(ctx, d_bits, n_sets, d_out, d_err); break;
    default: rc = launch_score_m<PDG_MODEL_UNIFORM>(ctx, d_bits, n_sets, d_out, d_err); break;
    }
    if (rc) return rc;
    int herr = 0;
    CK(cudaMemcpyAsync(scores, d_out, (size_t)n_sets * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bits); cudaFree(d_out); cudaFree(d_err);
    if (herr) return fail(ctx, PDG_E_NUMERIC, "scorer: non-SPD pivot or clique above the supported size");
    return PDG_OK;
}

int pdg_edge_tallies(pdg_ctx *ctx, int64_t burnin, void *tallies)
{
    if (!ctx || !tallies || burnin < 0) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    const size_t pp = (size_t)P.p * P.p;
    if (!ctx->d_since) {
        CK(cudaMalloc((void **)&ctx->d_since, (size_t)P.n_chains * pp * sizeof(int)));
        CK(cudaMalloc((void **)&ctx->d_tally, pp * 8));
    }
    int *since = ctx->d_since; unsigned long long *d_t = ctx->d_tally;
    CK(cudaMemsetAsync(d_t, 0, pp * 8, ctx->stream));
    k_edge_tallies<<<(unsigned)P.n_chains, 32, 0, ctx->stream>>>(P, since, burnin, d_t);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<long long> h(pp);
    CK(cudaMemcpyAsync(h.data(), d_t, pp * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int a = 0; a < P.p; ++a)
        for (int b = a + 1; b < P.p; ++b) h[(size_t)b * P.p + a] = h[(size_t)a * P.p + b];
    CK(cudaMemcpy(tallies, h.data(), pp * 8, cudaMemcpyDefault));
    return PDG_OK;
}

int pdg_host_alloc(void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return PDG_E_ARG;
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, bytes > 0 ? (size_t)bytes : 16, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? PDG_E_NOMEM : PDG_E_CUDA; }
    return PDG_OK;
}

int pdg_host_free(void *ptr)
{
    if (!ptr) return PDG_E_ARG;
    return cudaFreeHost(ptr) == cudaSuccess ? PDG_OK : PDG_E_CUDA;
}

int pdg_enable_timing(pdg_ctx *ctx, int on)
{
    if (!ctx) return PDG_E_ARG;
    ctx->timing = on != 0;
    for (auto &e : ctx->ev_run) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto &e : ctx->ev_rand) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    ctx->ev_run.clear(); ctx->ev_rand.clear();
    return PDG_OK;
}

int pdg_get_timing(pdg_ctx *ctx, double *ms_run, double *ms_randomize, int64_t *n_run, int64_t *n_randomize)
{
    if (!ctx) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    double a = 0.0, b = 0.0;
    for (auto &e : ctx->ev_run) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, e.first, e.second)); a += ms; }
    for (auto &e : ctx->ev_rand) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, e.first, e.second)); b += ms; }
    if (ms_run) *ms_run = a;
    if (ms_randomize) *ms_randomize = b;
    if (n_run) *n_run = (int64_t)ctx->ev_run.size();
    if (n_randomize) *n_randomize = (int64_t)ctx->ev_rand.size();
    return PDG_OK;
}

int pdg_get_work(pdg_ctx *ctx, int64_t work[16])
{
    if (!ctx || !work) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<unsigned long long> h((size_t)ctx->P.n_chains * 16);
    CK(cudaMemcpy(h.data(), ctx->P.work, h.size() * 8, cudaMemcpyDeviceToHost));
    for (int q = 0; q < 16; ++q) work[q] = 0;
    for (size_t c = 0; c < (size_t)ctx->P.n_chains; ++c) for (int q = 0; q < 16; ++q) work[q] += (int64_t)h[c * 16 + q];
#ifdef PDG_PHASE_TIMERS
    { unsigned long long g[8]; cudaMemcpyFromSymbol(g, g_pt, sizeof(g));
      fprintf(stderr, "lane scorer cycles: idx %llu gather %llu ldl %llu log %llu score %llu | calls K4 %llu K8 %llu\n", g[0], g[1], g[2], g[3], g[4], g[5], g[6]);
      unsigned long long kh[2][32]; cudaMemcpyFromSymbol(kh, g_kh, sizeof(kh));
      for (int k = 0; k < 32; ++k) if (kh[0][k]) fprintf(stderr, "  scans k=%d: %llu, %.0f cycles each\n", k, kh[0][k], (double)kh[1][k] / (double)kh[0][k]); }
#endif
    return PDG_OK;
}

}  // extern "C"
