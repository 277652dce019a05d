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
#include "pdg_randomize.cuh"
#include "pdg_tally.cuh"

// ---------------------------------------------------------------------------------------------
// logl[0] (mh_parallel.py:59-60 / :217-218): clique scores in tree-node order minus
// nu_s * score(s) over the DISTINCT non-empty separators (first occurrence in edge order), plus
// log_prior over cliques and distinct separators (seqdist:44-52).
struct InitSmem { int off_tri, off_idx, off_buf, off_sc, off_ss, off_cs, off_ssz, off_nu, off_first, total; };
__host__ __device__ inline InitSmem init_smem_layout(int p, int W, int nw)
{
    InitSmem s; int o = 0;
    s.off_tri = o; o += nw * 528 * 8;
    s.off_sc = o; o += p * 8;
    s.off_ss = o; o += p * 8;
    s.off_buf = o; o += nw * W * 8;
    s.off_idx = o; o += nw * ((((p + 1) * 2 + 15) / 16) * 16);
    s.off_cs = o; o += ((p * 2 + 15) / 16) * 16;
    s.off_ssz = o; o += ((p * 2 + 15) / 16) * 16;
    s.off_nu = o; o += ((p * 2 + 15) / 16) * 16;
    s.off_first = o; o += ((p + 15) / 16) * 16;
    s.total = o;
    return s;
}

template <int MODEL>
__global__ void __launch_bounds__(128) k_init_logl(PdgParams P, int kind)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int p = P.p, W = P.W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const long long chain = blockIdx.x;
    const InitSmem L = init_smem_layout(p, W, nw);
    double *s_sc = (double *)(smem + L.off_sc), *s_ss = (double *)(smem + L.off_ss);
    unsigned short *s_cs = (unsigned short *)(smem + L.off_cs), *s_ssz = (unsigned short *)(smem + L.off_ssz);
    unsigned short *s_nu = (unsigned short *)(smem + L.off_nu);
    unsigned char *s_first = smem + L.off_first;
    u64 *buf = (u64 *)(smem + L.off_buf) + warp * W;
    WarpScratch ws;
    ws.tri32 = (double *)(smem + L.off_tri) + warp * 528;
    ws.idx = (short *)(smem + L.off_idx) + warp * ((((p + 1) * 2 + 15) / 16) * 8);
    ws.kbig = P.kbig;
    ws.big = P.big ? P.big + ((size_t)chain * nw + warp) * ((size_t)(P.kbig + 1) * (P.kbig + 2) / 2) : nullptr;
    const u64 *Mc = P.M + (size_t)chain * p * W;
    const int *E = P.edges + (size_t)chain * 2 * (p - 1);
    int errflag = 0;
    for (int t = warp; t < p; t += nw) {
        const u64 *row = Mc + (size_t)t * W;
        int cs = 0;
        for (int w = 0; w < W; ++w) cs += __popcll(row[w]);
        double sc = 0.0;
        if (MODEL == PDG_MODEL_GGM && cs > 0) { const int e = ggm_set_score(P, ws, row, lane, sc); if (e) errflag |= e; }
        if (lane == 0) { s_sc[t] = sc; s_cs[t] = (unsigned short)cs; }
    }
    for (int e = warp; e < p - 1; e += nw) {
        const int a = E[2 * e], b = E[2 * e + 1];
        if (lane < W) buf[lane] = Mc[(size_t)a * W + lane] & Mc[(size_t)b * W + lane];
        __syncwarp();
        int sz = 0;
        for (int w = 0; w < W; ++w) sz += __popcll(buf[w]);
        int nu = 0; bool first = true;
        for (int e2 = lane; e2 < p - 1; e2 += 32) {
            const int a2 = E[2 * e2], b2 = E[2 * e2 + 1];
            bool same = true;
            for (int w = 0; w < W; ++w) same = same && ((Mc[(size_t)a2 * W + w] & Mc[(size_t)b2 * W + w]) == buf[w]);
            if (same) { ++nu; if (e2 < e) first = false; }
        }
        nu = warp_sum_i(nu);
        first = __all_sync(PDG_FULL, first);
        double sc = 0.0;
        if (MODEL == PDG_MODEL_GGM && first && sz > 0) { const int er = ggm_set_score(P, ws, buf, lane, sc); if (er) errflag |= er; }
        if (lane == 0) { s_ss[e] = sc; s_ssz[e] = (unsigned short)sz; s_nu[e] = (unsigned short)nu; s_first[e] = first ? 1 : 0; }
        __syncwarp();
    }
    __syncthreads();
    if (tid == 0) {
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

// batch scorer: one warp per complete set (sd.log_likelihood_partial([c], {}))
template <int MODEL>
__global__ void __launch_bounds__(128) k_score_sets(PdgParams P, const u64 *bits, long long n_sets, double *out, int *err)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int p = P.p, W = P.W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    WarpScratch ws;
    ws.tri32 = (double *)smem + warp * 528;
    ws.idx = (short *)(smem + nw * 528 * 8) + warp * ((((p + 1) * 2 + 15) / 16) * 8);
    ws.kbig = P.kbig;
    ws.big = P.big ? P.big + ((size_t)blockIdx.x * nw + warp) * ((size_t)(P.kbig + 1) * (P.kbig + 2) / 2) : nullptr;
    for (long long s = (long long)blockIdx.x * nw + warp; s < n_sets; s += (long long)gridDim.x * nw) {
        double sc = 0.0;
        if (MODEL == PDG_MODEL_GGM) { const int e = ggm_set_score(P, ws, bits + (size_t)s * W, lane, sc); if (e && lane == 0) atomicOr(err, e); }
        if (lane == 0) out[s] = sc;
        __syncwarp();
    }
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
    // compaction
    long long *d_offsets = nullptr;
    pdg_update *d_packed = nullptr;
    u64 *d_packed_bits = nullptr;
    long long packed_total = -1, packed_cap = 0;
    // replay staging
    std::vector<void *> replay_tmp;
    int nwarps = 4;
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

template <int MODEL, int KIND>
static int launch_run(pdg_ctx *ctx, long long b, long long e, const PdgReplayDev &R)
{
    const int NT = ctx->nwarps * 32;
    const RunSmem L = run_smem_layout(ctx->P.p, ctx->P.W, ctx->nwarps);
    int rc;
    if ((rc = set_smem(ctx, k_mh_run<MODEL, KIND>, L.total))) return rc;
    time_begin(ctx, ctx->ev_run);
    k_mh_run<MODEL, KIND><<<(unsigned)ctx->P.n_chains, NT, L.total, ctx->stream>>>(ctx->P, b, e, R);
    time_end(ctx, ctx->ev_run);
    ctx->launches++;
    CK(cudaGetLastError());
    return PDG_OK;
}

static int run_dispatch(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R)
{
    const int m = ctx->P.model;
    if (kind == PDG_KIND_PARALLEL) {
        if (m == PDG_MODEL_GGM) return launch_run<PDG_MODEL_GGM, PDG_KIND_PARALLEL>(ctx, b, e, R);
        return launch_run<PDG_MODEL_UNIFORM, PDG_KIND_PARALLEL>(ctx, b, e, R);
    }
    if (m == PDG_MODEL_GGM) return launch_run<PDG_MODEL_GGM, PDG_KIND_SINGLE>(ctx, b, e, R);
    return launch_run<PDG_MODEL_UNIFORM, PDG_KIND_SINGLE>(ctx, b, e, R);
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
    if (p < 2 || p > 4096 || n_chains < 1 || n_samples < 1 || rec_cap < 0) return PDG_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PDG_E_CUDA;   // no CPU fallback
    if (device < 0 || device >= ndev) return PDG_E_ARG;
    pdg_ctx *ctx = new pdg_ctx();
    ctx->device = device;
    if (const char *nwv = getenv("PDG_NWARPS")) { const int v = atoi(nwv); if (v >= 1 && v <= 4) ctx->nwarps = v; }
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return PDG_E_CUDA; }
    PdgParams &P = ctx->P;
    memset(&P, 0, sizeof(P));
    P.p = p; P.W = (p + 63) / 64;
    P.model = PDG_MODEL_UNIFORM; P.prior = PDG_PRIOR_MBC; P.prior_a = 2.0; P.prior_b = 4.0;
    P.n_chains = n_chains; P.chain0 = chain0; P.seed_lo = (u32)seed; P.seed_hi = (u32)(seed >> 32);
    P.n_samples = n_samples; P.rec_cap = rec_cap;
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
    TRY(dalloc(ctx, &P.work, C * 4));
    TRY(dalloc(ctx, &P.G0, C * pw));
    TRY(dalloc(ctx, &P.rec, C * (size_t)rec_cap));
    TRY(dalloc(ctx, &P.rec_bits, C * (size_t)rec_cap * 2 * P.W));
    TRY(dalloc(ctx, &ctx->Z.adj, C * pw));
    TRY(dalloc(ctx, &ctx->Z.K, C * pw));
    TRY(dalloc(ctx, &ctx->Z.sep, C * pw));
    TRY(dalloc(ctx, &ctx->d_offsets, C + 1));
    // scratch for cliques above the 32-row shared-memory triangle
    P.kbig = p + 1 < 256 ? p + 1 : 256;
    if (P.kbig > 32) TRY(dalloc(ctx, &P.big, C * ctx->nwarps * ((size_t)(P.kbig + 1) * (P.kbig + 2) / 2)));
    else { P.kbig = 0; P.big = nullptr; }
    TRY(cudaMemset(P.err, 0, C * sizeof(int)));
    TRY(cudaMemset(P.work, 0, C * 4 * sizeof(unsigned long long)));
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
    ctx->P.model = PDG_MODEL_UNIFORM;
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
    P.model = PDG_MODEL_GGM;
    ctx->model_set = true;
    return PDG_OK;
}

int pdg_set_model_loglin(pdg_ctx *ctx, const void *data, const void *levels, int64_t nrows, double cell_alpha,
                         int n_ktab, const void *ktab, const void *alphatab, const void *lgtab)
{
    if (!ctx) return PDG_E_ARG;
    return fail(ctx, PDG_E_STATE, "log-linear scorer not built in this revision");
}

static int launch_derived(pdg_ctx *ctx, long long b, long long n)
{
    const int sm = (2 * (ctx->P.p + 2)) * 2 + 64;
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
    const int NT = ctx->nwarps * 32;
    const InitSmem L = init_smem_layout(ctx->P.p, ctx->P.W, ctx->nwarps);
    int rc;
    if (ctx->P.model == PDG_MODEL_GGM) {
        if ((rc = set_smem(ctx, k_init_logl<PDG_MODEL_GGM>, L.total))) return rc;
        k_init_logl<PDG_MODEL_GGM><<<(unsigned)ctx->P.n_chains, NT, L.total, ctx->stream>>>(ctx->P, kind);
    } else {
        if ((rc = set_smem(ctx, k_init_logl<PDG_MODEL_UNIFORM>, L.total))) return rc;
        k_init_logl<PDG_MODEL_UNIFORM><<<(unsigned)ctx->P.n_chains, NT, L.total, ctx->stream>>>(ctx->P, kind);
    }
    ctx->launches++;
    CK(cudaGetLastError());
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
    int rc = pdg_init_logl(ctx, kind);
    if (rc) return rc;
    const long long M = ctx->P.n_samples;
    long long it = 1;
    while (it < M) {
        if (it % randomize == 0 && (rc = pdg_randomize(ctx, (uint32_t)it))) return rc;
        long long end = (it / randomize + 1) * randomize;
        if (end > M) end = M;
        if ((rc = pdg_run(ctx, kind, it, end, nullptr))) return rc;
        it = end;
    }
    return PDG_OK;
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
    const int nw = ctx->nwarps;
    long long grid = (n_sets + nw - 1) / nw;
    if (grid > P.n_chains) grid = P.n_chains;      // the k > 32 scratch is sized per chain
    const int sm = nw * 528 * 8 + nw * ((((P.p + 1) * 2 + 15) / 16) * 16);
    if (P.model == PDG_MODEL_GGM) k_score_sets<PDG_MODEL_GGM><<<(unsigned)grid, nw * 32, sm, ctx->stream>>>(P, d_bits, n_sets, d_out, d_err);
    else k_score_sets<PDG_MODEL_UNIFORM><<<(unsigned)grid, nw * 32, sm, ctx->stream>>>(P, d_bits, n_sets, d_out, d_err);
    ctx->launches++;
    CK(cudaGetLastError());
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
    int *since = nullptr; unsigned long long *d_t = nullptr;
    CK(cudaMalloc((void **)&since, (size_t)P.n_chains * pp * sizeof(int)));
    CK(cudaMalloc((void **)&d_t, pp * 8));
    CK(cudaMemsetAsync(d_t, 0, pp * 8, ctx->stream));
    k_edge_tallies<<<(unsigned)P.n_chains, 32, 0, ctx->stream>>>(P, since, burnin, d_t);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<long long> h(pp);
    CK(cudaMemcpyAsync(h.data(), d_t, pp * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(since); cudaFree(d_t);
    for (int a = 0; a < P.p; ++a)
        for (int b = a + 1; b < P.p; ++b) h[(size_t)b * P.p + a] = h[(size_t)a * P.p + b];
    CK(cudaMemcpy(tallies, h.data(), pp * 8, cudaMemcpyDefault));
    return PDG_OK;
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

int pdg_get_work(pdg_ctx *ctx, int64_t work[4])
{
    if (!ctx || !work) return PDG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<unsigned long long> h((size_t)ctx->P.n_chains * 4);
    CK(cudaMemcpy(h.data(), ctx->P.work, h.size() * 8, cudaMemcpyDeviceToHost));
    for (int q = 0; q < 4; ++q) work[q] = 0;
    for (size_t c = 0; c < (size_t)ctx->P.n_chains; ++c) for (int q = 0; q < 4; ++q) work[q] += (int64_t)h[c * 4 + q];
    return PDG_OK;
}

}  // extern "C"
