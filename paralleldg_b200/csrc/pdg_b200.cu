// libpdg_b200.so: C-ABI (include/pdg_b200.h) over the sm_100a kernels of the parallelDG
// parallel Metropolis-Hastings hot path.  No torch types cross this boundary.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#define PDG_DEFINE_GLOBAL_KERNELS
#include "pdg_ctx.h"
#include "pdg_tally.cuh"
#include "pdg_post.cuh"

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
#define fail pdg_fail
#define time_begin pdg_time_begin
#define time_end pdg_time_end

template <typename T>
static cudaError_t dalloc(pdg_ctx *ctx, T **ptr, size_t count)
{
    void *v = nullptr;
    cudaError_t e = cudaMalloc(&v, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
    if (e == cudaSuccess) { ctx->owned.push_back(v); *ptr = (T *)v; }
    return e;
}

// randomize > 0: the kernel randomises every chain's junction tree at every multiple of it
static int run_dispatch(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long randomize = 0)
{
    switch (ctx->P.model) {
    case PDG_MODEL_GGM: return pdg_launch_run_ggm(ctx, kind, b, e, R, randomize);
    case PDG_MODEL_LOGLIN: return pdg_launch_run_loglin(ctx, kind, b, e, R, randomize);
    default: return pdg_launch_run_uniform(ctx, kind, b, e, R, randomize);
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

// packed record buffers: grown with headroom and never shrunk -- the total changes a little from run to
// run, and a cudaFree + cudaMalloc of hundreds of MB per call costs tens of milliseconds
static int ensure_packed(pdg_ctx *ctx, long long tot)
{
    if (tot <= ctx->packed_cap) return PDG_OK;
    if (ctx->d_packed) { cudaFree(ctx->d_packed); cudaFree(ctx->d_packed_bits); ctx->d_packed = nullptr; ctx->d_packed_bits = nullptr; }
    ctx->packed_cap = 0;
    const long long cap = tot + tot / 4 + 1024;
    CK(cudaMalloc((void **)&ctx->d_packed, (size_t)cap * sizeof(pdg_update)));
    CK(cudaMalloc((void **)&ctx->d_packed_bits, (size_t)cap * 2 * ctx->P.W * 8));
    ctx->packed_cap = cap;
    return PDG_OK;
}

// tallies are accumulated for a < b only; mirror them
__global__ void k_symmetrize(unsigned long long *t, int p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p * p) return;
    const int a = i / p, b = i - a * p;
    if (a > b) t[i] = t[b * p + a];
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
    if (p < 2 || p > 64 * PDG_WMAX_LIMIT || n_chains < 1 || n_samples < 1 || n_samples > 2147483647LL || rec_cap < 0) return PDG_E_ARG;   // pdg_update.i is 32 bits wide
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PDG_E_CUDA;   // no CPU fallback
    if (device < 0 || device >= ndev) return PDG_E_ARG;
    pdg_ctx *ctx = new pdg_ctx();
    ctx->device = device;
    ctx->wpc = n_chains >= 8 ? 8 : (int)n_chains;      // upper bound; the launcher spreads the chains over all SMs
    if (const char *v = getenv("PDG_WPC")) { const int x = atoi(v); if (x >= 1 && x <= 8) { ctx->wpc = x; ctx->wpc_forced = true; } }
    { int n = 0; if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && n > 0) ctx->sm_count = n; }
    // two warps per chain (speculative pairs) while that still leaves every chain resident at once
    ctx->spec = n_chains <= (int64_t)ctx->sm_count * 8;
    if (const char *v = getenv("PDG_SPEC")) { ctx->spec = atoi(v) != 0; ctx->spec_forced = true; }
    // 16M slots (256 MB) for one- and two-word keys: the p = 50 hit rate saturates at 2^23; wider
    // keys come with far more distinct cliques (p = 500: 21M per run): 64M slots = 5.4 GB there -- the hit
    // rate is that of 32M slots (81 %) but the probe sequences are shorter (61.7 -> 64.9 M proposals/s)
    ctx->memo_bits = (p > 128) ? 26 : 24;
    if (const char *v = getenv("PDG_MEMO_BITS")) { const int x = atoi(v); if (x >= 8 && x <= 28) ctx->memo_bits = x; }
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return PDG_E_CUDA; }
    PdgParams &P = ctx->P;
    memset(&P, 0, sizeof(P));
    P.p = p; P.W = (p + 63) / 64;
    P.model = PDG_MODEL_UNIFORM; P.prior = PDG_PRIOR_MBC; P.prior_a = 2.0; P.prior_b = 4.0;
    P.n_chains = n_chains; P.chain0 = chain0; P.seed_lo = (u32)seed; P.seed_hi = (u32)(seed >> 32);
    P.n_samples = n_samples; P.rec_cap = rec_cap;
    P.memo_small = getenv("PDG_NO_MEMO_SMALL") ? 0 : 1;
    P.tri_extra = 0; P.ktri_run = 32;
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
    TRY(dalloc(ctx, &P.work, 2 * C * 16));              // (two scratch / work slots per chain: one per warp of a pair)
    TRY(dalloc(ctx, &P.G0, C * pw));
    TRY(dalloc(ctx, &P.rec, C * (size_t)rec_cap));
    TRY(dalloc(ctx, &P.rec_bits, C * (size_t)rec_cap * 2 * P.W));
    TRY(dalloc(ctx, &ctx->Z.adj, C * pw));
    TRY(dalloc(ctx, &ctx->Z.K, C * pw));
    TRY(dalloc(ctx, &ctx->Z.sep, C * pw));
    TRY(dalloc(ctx, &ctx->d_offsets, C + 1));
    // scratch for cliques above the 32-row shared-memory triangle
    P.kbig = p + 1 < 256 ? p + 1 : 256;
    if (P.kbig > 32) TRY(dalloc(ctx, &P.big, 2 * C * ((size_t)(P.kbig + 1) * (P.kbig + 2) / 2)));
    else { P.kbig = 0; P.big = nullptr; }
    TRY(cudaMemset(P.err, 0, C * sizeof(int)));
    TRY(cudaMemset(P.work, 0, 2 * C * 16 * sizeof(unsigned long long)));
    {
        const int W = P.W;
        ctx->wmax = W <= 1 ? 1 : W <= 2 ? 2 : W <= 4 ? 4 : 8;
        ctx->smem_state = (size_t)p * W * 16 <= 8192;
        MemoTable &T = P.memo;
        T.W = (ctx->wmax == 1) ? 1 : W;
        // the table never takes more than half of the memory that is still free on the device
        {
            size_t mfree = 0, mtotal = 0;
            if (e == cudaSuccess && cudaMemGetInfo(&mfree, &mtotal) == cudaSuccess) {
                const size_t per_slot = (ctx->wmax == 1) ? 16 : (size_t)16 + 8 * (size_t)W;
                while (ctx->memo_bits > 8 && (per_slot << ctx->memo_bits) > mfree / 2) --ctx->memo_bits;
            }
        }
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
    // the memo table must never be read as allocated: cudaMalloc hands back the previous owner's
    // bytes (possibly another model's {clique -> score} entries)
    if (memo_clear(ctx) != PDG_OK || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        for (void *v : ctx->owned) cudaFree(v);
        delete ctx;
        return PDG_E_CUDA;
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
    if (ctx->d_sbits) { cudaFree(ctx->d_sbits); cudaFree(ctx->d_sout); }
    if (ctx->d_serr) cudaFree(ctx->d_serr);
    if (ctx->d_big_list) cudaFree(ctx->d_big_list);
    for (auto *v : {&ctx->ev_run, &ctx->ev_rand, &ctx->ev_free})
        for (auto &e : *v) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->ev_copy); }
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
    ctx->P.model = PDG_MODEL_UNIFORM; ctx->P.ll_cap = 0; ctx->P.tri_extra = 0; ctx->P.ktri_run = 32;
    ctx->model_set = true;
    return memo_clear(ctx);                 // cached scores belong to the model that produced them
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
    // Large p: the in-kernel randomiser needs more shared memory per warp than the MH loop, so the loop's
    // scratch triangle takes the difference and holds cliques well above 32 vertices (p = 500: 59) that
    // would otherwise be factorised in the global scratch.
    P.tri_extra = 0; P.ktri_run = 32;
    if (!ctx->smem_state && !getenv("PDG_NO_TRI_EXTRA")) {
        const RunSmem L0 = run_smem_layout(P.p, P.W, false, tri_bytes_for(0));
        const RandSmem LR = rand_smem_layout(P.p, P.W);
        int extra = ((LR.total - L0.total) / 16) * 16;
        if (extra > 0) {
            int k = 32;
            while ((long long)(k + 1) * (k + 2) / 2 * 8 <= (long long)tri_bytes_for(0) + extra && k + 1 <= P.kbig) ++k;
            P.tri_extra = extra; P.ktri_run = k;
        }
    }
    ctx->model_set = true;
    return memo_clear(ctx);                 // cached scores belong to the model that produced them
}

int pdg_set_model_loglin(pdg_ctx *ctx, const void *data, const void *levels, int64_t nrows, double cell_alpha,
                         int n_ktab, const void *ktab, const void *alphatab, const void *lgtab)
{
    if (!ctx || !data || !levels || nrows < 1 || n_ktab < 0) return PDG_E_ARG;
    if (n_ktab > 0 && (!ktab || !alphatab || !lgtab)) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    // dense cell counters: one table per warp slot (= chain); larger sets go through a pool of
    // hashed tables
    int hist_cells = 1 << 16;
    if (const char *v = getenv("PDG_HIST_CELLS")) { const int x = atoi(v); if (x >= 2 && x <= (1 << 16)) hist_cells = x; }
    // columns padded to a 16-byte stride, an all-zero column behind the last one and 4 KB of slack:
    // the dense scan (pdg_loglin.cuh) then needs no load predicates
    const size_t ld = ((size_t)nrows + 15) & ~(size_t)15;
    const size_t data_bytes = (size_t)(P.p + 1) * ld + 4096;
    LoglinSparse sp;
    sp.nsp = (int)(P.n_chains < 64 ? P.n_chains : 64);
    int hs = 1;
    while (hs < 2 * nrows) hs <<= 1;
    sp.hslots = hs;
    // A pooled context that is re-armed with a model of the same shape keeps its buffers: the count
    // tables (0.5 GB at 1024 chains) are left all-zero by every scan, so only the data and the
    // lgamma tables are uploaded again -- no free / malloc / memset per call.
    const long long shape[4] = {(long long)nrows, (long long)n_ktab, (long long)hist_cells, P.n_chains};
    const bool reuse = !ctx->model_owned.empty() && ctx->P.model == PDG_MODEL_LOGLIN &&
                       shape[0] == ctx->ll_shape[0] && shape[1] == ctx->ll_shape[1] &&
                       shape[2] == ctx->ll_shape[2] && shape[3] == ctx->ll_shape[3];
    if (!reuse) {
        for (void *v : ctx->model_owned) cudaFree(v);
        ctx->model_owned.clear();
        ctx->ll_shape[0] = -1;
        auto own = [&](size_t bytes, void **out) -> cudaError_t {
            cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
            if (e == cudaSuccess) ctx->model_owned.push_back(*out);
            return e;
        };
        void *d_data = nullptr, *d_lv = nullptr, *d_k = nullptr, *d_a = nullptr, *d_lg = nullptr;
        void *d_hist = nullptr, *d_lock = nullptr, *d_hk = nullptr, *d_hc = nullptr, *d_cc = nullptr, *d_list = nullptr;
        CK(own(data_bytes, &d_data));
        CK(own((size_t)P.p * sizeof(int), &d_lv));
        CK(own((size_t)(n_ktab > 0 ? n_ktab : 1) * 8, &d_k));
        CK(own((size_t)(n_ktab > 0 ? n_ktab : 1) * 8, &d_a));
        CK(own((size_t)(n_ktab > 0 ? n_ktab : 1) * (size_t)(nrows + 2) * 8, &d_lg));
        CK(own((size_t)2 * P.n_chains * hist_cells * 4, &d_hist));
        CK(own((size_t)sp.nsp * 4, &d_lock));
        CK(own((size_t)sp.nsp * hs * 8, &d_hk));
        CK(own((size_t)sp.nsp * hs * 4, &d_hc));
        CK(own((size_t)sp.nsp * (size_t)(nrows + 2) * 4, &d_cc));
        CK(own((size_t)sp.nsp * (size_t)(nrows + 512) * 4, &d_list));
        CK(cudaMemsetAsync(d_data, 0, data_bytes, ctx->stream));
        CK(cudaMemsetAsync(d_hist, 0, (size_t)2 * P.n_chains * hist_cells * 4, ctx->stream));
        CK(cudaMemsetAsync(d_lock, 0, (size_t)sp.nsp * 4, ctx->stream));
        CK(cudaMemsetAsync(d_hk, 0, (size_t)sp.nsp * hs * 8, ctx->stream));
        CK(cudaMemsetAsync(d_hc, 0, (size_t)sp.nsp * hs * 4, ctx->stream));
        CK(cudaMemsetAsync(d_cc, 0, (size_t)sp.nsp * (size_t)(nrows + 2) * 4, ctx->stream));
        CK(cudaMemsetAsync(d_list, 0, (size_t)sp.nsp * (size_t)(nrows + 512) * 4, ctx->stream));
        P.data = (const unsigned char *)d_data; P.levels = (const int *)d_lv;
        P.ktab = (const double *)d_k; P.alphatab = (const double *)d_a; P.lgtab = (const double *)d_lg;
        P.hist = (unsigned int *)d_hist;
        sp.lock = (int *)d_lock; sp.hkeys = (u64 *)d_hk; sp.hcnt = (unsigned int *)d_hc; sp.cc = (unsigned int *)d_cc; sp.list = (int *)d_list;
        P.sp = sp;
        for (int q = 0; q < 4; ++q) ctx->ll_shape[q] = shape[q];
    }
    // (padding rows / the zero column were cleared when the buffer was created and are never written)
    CK(cudaMemcpy2DAsync((void *)P.data, ld, data, (size_t)nrows, (size_t)nrows, (size_t)P.p, cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync((void *)P.levels, levels, (size_t)P.p * sizeof(int), cudaMemcpyDefault, ctx->stream));
    if (n_ktab > 0) {
        CK(cudaMemcpyAsync((void *)P.ktab, ktab, (size_t)n_ktab * 8, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync((void *)P.alphatab, alphatab, (size_t)n_ktab * 8, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync((void *)P.lgtab, lgtab, (size_t)n_ktab * (size_t)(nrows + 2) * 8, cudaMemcpyDefault, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));       // the host arrays may go away after the call
    P.data_ld = (long long)ld; P.nrows = nrows; P.cell_alpha = cell_alpha;
    P.n_ktab = n_ktab; P.hist_cells = hist_cells;
    P.model = PDG_MODEL_LOGLIN; P.tri_extra = 0; P.ktri_run = 32;
    // few rows: a scan is cheap, so the light kernel (two CTAs per SM, 1024-cell tables) runs the chains
    ctx->light = nrows <= 8192;
    if (const char *v = getenv("PDG_LIGHT")) ctx->light = atoi(v) != 0;
    // per-warp shared count table: 4096 cells when a CTA of the MH kernel still fits in shared memory
    P.ll_cap = ctx->light ? 1024 : 4096;
    if (const char *v = getenv("PDG_LL_CAP")) { const int x = atoi(v); if (x >= 1024 && x <= 8192 && !(x & (x - 1))) P.ll_cap = x; }
    for (;;) {
        const RunSmem L = run_smem_layout(P.p, P.W, ctx->smem_state, tri_bytes_for(P.ll_cap));
        const RandSmem LR = rand_smem_layout(P.p, P.W);
        const int run_bytes = L.total * ((ctx->spec && ctx->spec_forced) ? 2 : 1);    // log-linear: pairs only when forced
        const int per_chain = LR.total > run_bytes ? LR.total : run_bytes;
        if ((long long)per_chain * pdg_chains_per_cta(ctx, ctx->light) <= 224 * 1024 || P.ll_cap <= 1024) break;
        P.ll_cap >>= 1;
    }
    ctx->model_set = true;
    return memo_clear(ctx);                 // cached scores belong to the model that produced them
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
    case PDG_MODEL_GGM: rc = pdg_launch_init_ggm(ctx, kind); break;
    case PDG_MODEL_LOGLIN: rc = pdg_launch_init_loglin(ctx, kind); break;
    default: rc = pdg_launch_init_uniform(ctx, kind); break;
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
    // The whole run is ONE launch: every warp carries its chain through all iterations and randomises
    // its own junction tree every `randomize` iterations, so no chain ever waits for another and there is
    // a single launch tail per run instead of one per randomisation interval (round 1: a third of every
    // 100-iteration launch was the tail of its slowest CTA).  PDG_FUSED=0 restores one launch per interval
    // with the randomiser as its own kernel in between.
    // (the light log-linear kernel -- thousands of short-table chains, two CTAs per SM -- keeps one launch
    // per interval: 1340 vs 1070 M proposals/s on config 3, its chains are cheap enough for the tail not to
    // matter and the separate randomiser kernel runs at full occupancy)
    bool fused = !(ctx->P.model == PDG_MODEL_LOGLIN && ctx->light);
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
    int rc = ensure_packed(ctx, tot);
    if (rc) return rc;
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

int pdg_collect(pdg_ctx *ctx, int64_t *total, void *n_proposals, void *n_accepted, void *offsets, void *logl,
                int64_t burnin, void *tallies)
{
    if (!ctx || !total || burnin < 0) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    const size_t C = (size_t)P.n_chains, pp = (size_t)P.p * P.p;
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
    }
    // the log-probability traces leave on the copy stream as soon as the sampling kernels are done,
    // while this stream scans the record counts, packs the records and tallies the edges
    CK(cudaEventRecord(ctx->ev_copy, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
    if (logl) CK(cudaMemcpyAsync(logl, P.logl, C * (size_t)P.n_samples * 8, cudaMemcpyDefault, ctx->copy_stream));
    k_scan_counts<<<1, 1024, 0, ctx->stream>>>(P.nrec, P.rec_cap, P.n_chains, ctx->d_offsets);
    ctx->launches++;
    CK(cudaGetLastError());
    long long tot = 0;
    CK(cudaMemcpyAsync(&tot, ctx->d_offsets + P.n_chains, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_proposals) CK(cudaMemcpyAsync(n_proposals, P.kcount, C * 8, cudaMemcpyDefault, ctx->stream));
    if (n_accepted) CK(cudaMemcpyAsync(n_accepted, P.nrec, C * 8, cudaMemcpyDefault, ctx->stream));
    if (offsets) CK(cudaMemcpyAsync(offsets, ctx->d_offsets, (C + 1) * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));              // (the caller sizes its record buffers from `total`)
    int rc = ensure_packed(ctx, tot);
    if (rc) return rc;
    if (tot > 0) {
        k_compact<<<(unsigned)P.n_chains, 128, 0, ctx->stream>>>(P, ctx->d_offsets, ctx->d_packed, ctx->d_packed_bits);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    ctx->packed_total = tot;
    *total = tot;
    if (tallies) {
        if (!ctx->d_since) {
            CK(cudaMalloc((void **)&ctx->d_since, C * pp * sizeof(int)));
            CK(cudaMalloc((void **)&ctx->d_tally, pp * 8));
        }
        CK(cudaMemsetAsync(ctx->d_tally, 0, pp * 8, ctx->stream));
        k_edge_tallies<<<(unsigned)P.n_chains, 32, 0, ctx->stream>>>(P, ctx->d_since, burnin, ctx->d_tally);
        ctx->launches++;
        CK(cudaGetLastError());
        k_symmetrize<<<(unsigned)((pp + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_tally, P.p);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(tallies, ctx->d_tally, pp * 8, cudaMemcpyDefault, ctx->stream));
    }
    return PDG_OK;                                       // pdg_get_updates / pdg_get_state finish the collection
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
    if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));      // a pdg_collect in flight ends here
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
    if (n_sets > ctx->score_cap) {           // staging kept by the context, freed by pdg_destroy
        if (ctx->d_sbits) { cudaFree(ctx->d_sbits); cudaFree(ctx->d_sout); ctx->d_sbits = nullptr; ctx->d_sout = nullptr; }
        ctx->score_cap = 0;
        CK(cudaMalloc((void **)&ctx->d_sbits, (size_t)n_sets * P.W * 8));
        CK(cudaMalloc((void **)&ctx->d_sout, (size_t)n_sets * 8));
        ctx->score_cap = n_sets;
    }
    if (!ctx->d_serr) CK(cudaMalloc((void **)&ctx->d_serr, 4));
    CK(cudaMemsetAsync(ctx->d_serr, 0, 4, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_sbits, bits, (size_t)n_sets * P.W * 8, cudaMemcpyDefault, ctx->stream));
    int rc;
    switch (P.model) {
    case PDG_MODEL_GGM: rc = pdg_launch_score_ggm(ctx, ctx->d_sbits, n_sets, ctx->d_sout, ctx->d_serr); break;
    case PDG_MODEL_LOGLIN: rc = pdg_launch_score_loglin(ctx, ctx->d_sbits, n_sets, ctx->d_sout, ctx->d_serr); break;
    default: rc = pdg_launch_score_uniform(ctx, ctx->d_sbits, n_sets, ctx->d_sout, ctx->d_serr); break;
    }
    if (rc) return rc;
    int herr = 0;
    CK(cudaMemcpyAsync(scores, ctx->d_sout, (size_t)n_sets * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync(&herr, ctx->d_serr, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
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

// RAII for the temporaries of the post-processing calls
namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};
}

int pdg_size_series(pdg_ctx *ctx, void *sizes)
{
    if (!ctx || !sizes) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    const size_t bytes = (size_t)P.n_chains * (size_t)P.n_samples * sizeof(int);
    cudaPointerAttributes at;
    const bool on_device = cudaPointerGetAttributes(&at, sizes) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    DevBuf tmp;
    int *dst = (int *)sizes;
    if (!on_device) { CK(tmp.alloc(bytes)); dst = (int *)tmp.p; }
    k_size_series<<<(unsigned)P.n_chains, 32, 0, ctx->stream>>>(P, dst);
    ctx->launches++;
    CK(cudaGetLastError());
    if (!on_device) CK(cudaMemcpyAsync(sizes, dst, bytes, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PDG_OK;
}

int pdg_edge_ess(pdg_ctx *ctx, int64_t burnin, double lo, double hi, void *ess, void *n_edges)
{
    if (!ctx || !ess || burnin < 0 || burnin >= ctx->P.n_samples) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    PdgParams &P = ctx->P;
    const size_t C = (size_t)P.n_chains, pp = (size_t)P.p * P.p;
    const long long len = P.n_samples - burnin;
    const int words = (int)((len + 63) / 64);
    const size_t smem = (size_t)ESS_WARPS * ((size_t)(words + 2) * 8 + (size_t)(words + 1) * 4);
    if (smem > 200 * 1024) return fail(ctx, PDG_E_ARG, "pdg_edge_ess: series longer than 400 000 positions");
    if (!ctx->d_since) {
        CK(cudaMalloc((void **)&ctx->d_since, C * pp * sizeof(int)));
        CK(cudaMalloc((void **)&ctx->d_tally, pp * 8));
    }
    DevBuf slot, nq, bm, vals, out;
    CK(slot.alloc(C * pp * sizeof(int)));
    CK(nq.alloc(C * sizeof(int)));
    CK(out.alloc(C * 8));
    k_edge_presence<<<(unsigned)C, 32, 0, ctx->stream>>>(P, ctx->d_since, (int *)slot.p, burnin, lo, hi, (int *)nq.p);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<int> hq(C);
    CK(cudaMemcpyAsync(hq.data(), nq.p, C * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int qcap = 1;
    for (size_t c = 0; c < C; ++c) qcap = hq[c] > qcap ? hq[c] : qcap;
    long long grid = (long long)ctx->sm_count * 2;
    if (grid > (long long)C) grid = (long long)C;
    CK(bm.alloc((size_t)grid * qcap * words * 8));
    CK(vals.alloc((size_t)grid * qcap * 8));
    CK(cudaFuncSetAttribute(k_edge_ess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_edge_ess<<<(unsigned)grid, 32 * ESS_WARPS, smem, ctx->stream>>>(P, (const int *)slot.p, (const int *)nq.p, burnin, qcap,
                                                                     words, (u64 *)bm.p, (double *)vals.p, (double *)out.p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ess, out.p, C * 8, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_edges) CK(cudaMemcpy(n_edges, hq.data(), C * sizeof(int), cudaMemcpyDefault));
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
    // the recorded pairs go back to the pool (destroyed with the context)
    for (auto &e : ctx->ev_run) ctx->ev_free.push_back(e);
    for (auto &e : ctx->ev_rand) ctx->ev_free.push_back(e);
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
    std::vector<unsigned long long> h((size_t)ctx->P.n_chains * 2 * 16);
    CK(cudaMemcpy(h.data(), ctx->P.work, h.size() * 8, cudaMemcpyDeviceToHost));
    for (int q = 0; q < 16; ++q) work[q] = 0;
    for (size_t c = 0; c < (size_t)ctx->P.n_chains * 2; ++c) for (int q = 0; q < 16; ++q) work[q] += (int64_t)h[c * 16 + q];
    return PDG_OK;
}

#ifdef PDG_PHASE_TIMERS
/* developer probe (PDG_PHASE_TIMERS builds only): the raw per-slot work counters, [2 * n_chains][16] */
int pdg_debug_work_per_chain(pdg_ctx *ctx, void *out)
{
    if (!ctx || !out) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, ctx->P.work, (size_t)ctx->P.n_chains * 2 * 16 * 8, cudaMemcpyDeviceToHost));
    return PDG_OK;
}
#endif

}  // extern "C"
