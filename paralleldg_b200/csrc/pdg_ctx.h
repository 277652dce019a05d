// Host-side context of libpdg_b200.so, shared by the C-ABI translation unit (pdg_b200.cu) and the
// per-model translation units that instantiate the MH / init / batch-scorer kernels
// (pdg_model_ggm.cu, pdg_model_loglin.cu, pdg_model_uniform.cu; compiled in parallel).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <utility>
#include <vector>

#include "pdg_kernels.cuh"
#include "pdg_randomize.cuh"

struct pdg_ctx {
    int device = 0;
    PdgParams P;
    RandScratch Z;
    cudaStream_t stream = 0;
    std::string err;
    bool model_set = false, state_set = false;
    long long launches = 0;
    std::vector<void *> owned;
    // model buffers (kept across re-sets of the same shape)
    double *d_A = nullptr, *d_D = nullptr, *d_logD = nullptr, *d_gconst = nullptr;
    std::vector<void *> model_owned;      // log-linear buffers, freed on a re-set with another shape / destroy
    long long ll_shape[4] = {-1, -1, -1, -1};   // (nrows, n_ktab, hist_cells, n_chains) the buffers were sized for
    int wmax = 1;                         // compile-time word count the kernels are dispatched on
    int wpc = 8;                          // most warps (= chains) per CTA of the MH kernel
    bool wpc_forced = false;              // PDG_WPC set: use exactly wpc
    int sm_count = 148;
    bool smem_state = false;
    bool light = false;                   // log-linear model with few rows: the 128-register MH kernel
    bool spec = false;                    // two warps per chain (speculative pairs, pdg_run.cuh): chains <= 8 per SM
    bool spec_forced = false;             // PDG_SPEC set: use / do not use them whatever the model
    int memo_bits = 24;
    // compaction
    long long *d_offsets = nullptr;
    pdg_update *d_packed = nullptr;
    u64 *d_packed_bits = nullptr;
    long long packed_total = -1, packed_cap = 0;
    int *d_since = nullptr;               // edge-tally scratch, kept between calls
    unsigned long long *d_tally = nullptr;
    // batch-scorer staging, kept between calls
    u64 *d_sbits = nullptr; double *d_sout = nullptr; int *d_serr = nullptr; long long score_cap = 0;
    int *d_big_list = nullptr; long long big_cap = 0;     // [0] = count, then the indices of the sets for k_score_big
    // replay staging
    std::vector<void *> replay_tmp;
    // CUDA-event timing of the MH and randomisation launches (events come from a pool and are
    // destroyed with the context)
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_run, ev_rand, ev_free;
    // second stream + events for device-to-host copies that overlap the next launch
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy = nullptr;
};

// chains per CTA of the MH kernel: spread over ALL SMs (1024 chains on 148 SMs -> 7 per CTA, 147 CTAs,
// instead of 128 CTAs of 8); with more chains than fit at once, whole waves of equal size
inline int pdg_chains_per_cta(const pdg_ctx *ctx, bool light)
{
    if (ctx->wpc_forced) return ctx->wpc;
    const long long resident = (long long)ctx->sm_count * ((light && !ctx->spec) ? 2 : 1);
    const long long waves = (ctx->P.n_chains + resident * 8 - 1) / (resident * 8);
    int wpc = (int)((ctx->P.n_chains + resident * waves - 1) / (resident * waves));
    return wpc < 1 ? 1 : wpc > 8 ? 8 : wpc;
}

inline int pdg_fail(pdg_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return pdg_fail(ctx, PDG_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

inline void pdg_time_begin(pdg_ctx *ctx, std::vector<std::pair<cudaEvent_t, cudaEvent_t>> &v)
{
    if (!ctx->timing) return;
    std::pair<cudaEvent_t, cudaEvent_t> e;
    if (!ctx->ev_free.empty()) { e = ctx->ev_free.back(); ctx->ev_free.pop_back(); }
    else { cudaEventCreate(&e.first); cudaEventCreate(&e.second); }
    cudaEventRecord(e.first, ctx->stream);
    v.push_back(e);
}
inline void pdg_time_end(pdg_ctx *ctx, std::vector<std::pair<cudaEvent_t, cudaEvent_t>> &v)
{
    if (!ctx->timing) return;
    cudaEventRecord(v.back().second, ctx->stream);
}

// per-model launchers (one translation unit each)
#define PDG_DECLARE_MODEL(NAME)                                                                                         \
    int pdg_launch_run_##NAME(pdg_ctx *ctx, int kind, long long b, long long e, const PdgReplayDev &R, long long rz); \
    int pdg_launch_init_##NAME(pdg_ctx *ctx, int kind);                                                               \
    int pdg_launch_score_##NAME(pdg_ctx *ctx, const u64 *d_bits, long long n_sets, double *d_out, int *d_err);
PDG_DECLARE_MODEL(ggm)
PDG_DECLARE_MODEL(loglin)
PDG_DECLARE_MODEL(uniform)
