/*
 * pdg_b200.h -- C-ABI of libpdg_b200.so: the B200 (sm_100a) implementation of parallelDG's
 * parallel Metropolis-Hastings hot path, many independent chains at once.
 *
 * This is the drop-in boundary.  The reference is pure Python and has no FFI; each entry point
 * below names the reference interface it replaces (paths under /root/reference/parallelDG/).
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every call returns 0 on success or a negative PDG_E_* code; pdg_last_error() gives the
 *     message of the last failure on that context.  Nothing throws across the ABI.
 *   - `const void *` inputs may be HOST or DEVICE pointers (resolved with UVA /
 *     cudaMemcpyDefault); outputs are written to caller-owned HOST or DEVICE buffers.
 *   - the caller owns every buffer it passes in; the library owns only its internal device state.
 *   - one context = one GPU + one stream; calls on a context are not re-entrant.  One context
 *     per GPU, driven by one host thread/process per GPU (replaces process-per-chain,
 *     mh_parallel.py:479-530).
 *   - bitsets are W = ceil(p/64) little-endian 64-bit words, bit g of word g/64.
 *   - a chain's state is the reference's JunctionMap (graph/junction_tree.py:11-132): a tree on
 *     p tree nodes (p-1 edges) and one clique bitset per tree node.
 */
#ifndef PDG_B200_H
#define PDG_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDG_ABI_VERSION 1

enum pdg_error {
    PDG_OK = 0,
    PDG_E_ARG = -1,          /* bad argument */
    PDG_E_CUDA = -2,         /* CUDA runtime error */
    PDG_E_STATE = -3,        /* call order (model/state not set) */
    PDG_E_REPLAY = -4,       /* replay log disagrees with the enumerated move set */
    PDG_E_OVERFLOW = -5,     /* update-record capacity exceeded */
    PDG_E_NUMERIC = -6,      /* non-positive pivot / unsupported clique size */
    PDG_E_NOMEM = -7
};

enum pdg_model_kind { PDG_MODEL_UNIFORM = 0, PDG_MODEL_GGM = 1, PDG_MODEL_LOGLIN = 2 };
/* mh_parallel.py:317-332 get_prior */
enum pdg_prior_kind { PDG_PRIOR_MBC = 0, PDG_PRIOR_EDGEPENALTY = 1, PDG_PRIOR_JUMPPENALTY = 2, PDG_PRIOR_UNIFORM = 3 };
/* mh_parallel.py:182-315 (parallel moves) / :23-180 (single move) */
enum pdg_kernel_kind { PDG_KIND_PARALLEL = 0, PDG_KIND_SINGLE = 1 };

typedef struct pdg_ctx pdg_ctx;

/* one accepted move = one entry of Trajectory.jt_updates (mh_parallel.py:270,302):
 * (i, k, move_type, {node}, (Cnew, C, Cadj)); C and Cadj travel as bitsets next to the header,
 * Cnew = C +/- node. */
typedef struct pdg_update {
    int32_t i;        /* MCMC index */
    int32_t k;        /* proposal sub-index */
    uint16_t node;    /* graph vertex moved */
    uint16_t U;       /* tree node whose clique changed */
    int16_t Uadj;     /* anchor tree node, -1 = none */
    uint8_t type;     /* 0 connect, 1 disconnect */
    uint8_t pad;
} pdg_update;

/* replay of the reference's random streams; arrays for ALL chains of the
 * context, chain-major.  HOST or DEVICE pointers. */
typedef struct pdg_replay {
    int64_t n_samples;        /* stride of the it_* arrays */
    const int32_t *it_node;   /* [n_chains][n_samples] np.random.randint(p)   (mh_parallel.py:230) */
    const int32_t *it_mtype;  /* [n_chains][n_samples] np.random.randint(2)   (:231) */
    const int32_t *it_aux;    /* [n_chains][n_samples] random tree node (pm:138) / random leaf (pm:185), else -1 */
    const int64_t *it_off;    /* [n_chains][n_samples+1] offsets into the mv_* arrays */
    const int32_t *mv_U;      /* ordered proposals of every iteration */
    const int32_t *mv_Uadj;
    const double *mv_u;       /* np.random.uniform() per proposal (:268,:300) */
    int64_t n_moves_total;
} pdg_replay;

int pdg_abi_version(void);

/* replaces: the per-chain set-up of sample_trajectory (mh_parallel.py:188-223).
 * chain ids chain0 .. chain0+n_chains-1 select the Philox streams (so shards on different GPUs
 * draw different chains); rec_cap = update records kept per chain. */
int pdg_create(pdg_ctx **out, int device, int p, int64_t n_chains, uint32_t chain0,
               uint64_t seed, int64_t n_samples, int64_t rec_cap);
int pdg_destroy(pdg_ctx *ctx);
/* re-arms a context for another run of the same shape: new Philox key / first chain id, error
 * flags and counters cleared (lets a host reuse the device allocations across calls) */
int pdg_set_seed(pdg_ctx *ctx, uint64_t seed, uint32_t chain0);
const char *pdg_last_error(pdg_ctx *ctx);
/* the stream every later call is enqueued on (0 = default); a cudaStream_t passed as void* */
int pdg_set_stream(pdg_ctx *ctx, void *stream);

/* replaces: get_prior (mh_parallel.py:317-332) + seqdist priors (:31-151) */
int pdg_set_prior(pdg_ctx *ctx, int prior_kind, double a, double b);
/* replaces: UniformJTDistribution (seqdist:154-176) */
int pdg_set_model_uniform(pdg_ctx *ctx);
/* replaces: GGMJTPosterior.init_model (seqdist:334-343).  A = D + X^T X and D are p*p row-major
 * fp64; gconst[k], k = 0..p, holds the Gamma-function part of the clique score
 * (wishart.py:24-30):  (-t1 k log(pi) + multigammaln(t1,k)) - (-t2 k log(pi) + multigammaln(t2,k)),
 * t1 = (delta+n+k-1)/2, t2 = (delta+k-1)/2, gconst[0] = 0. */
int pdg_set_model_ggm(pdg_ctx *ctx, const void *A, const void *D, int64_t n, double delta,
                      const void *gconst);
/* replaces: LogLinearJTPosterior.init_model (seqdist:239-252).  data: column-major level codes
 * [p][nrows] uint8; levels[p]; optional lgamma tables for n_ktab distinct cell counts
 * ktab[j] = prod of levels over a set (fp64): alphatab[j] = the per-cell pseudo count of such a
 * set (loglin:52-55), lgtab[j][c] = lgamma(c + alpha_j) - lgamma(alpha_j) for c = 0..nrows
 * (dirichlet.py:76-83) and lgtab[j][nrows+1] = lgamma(K alpha_j + nrows) - lgamma(K alpha_j)
 * (row stride nrows + 2).  Sets whose cell count is not in ktab use the device lgamma. */
int pdg_set_model_loglin(pdg_ctx *ctx, const void *data, const void *levels, int64_t nrows,
                         double cell_alpha, int n_ktab, const void *ktab, const void *alphatab,
                         const void *lgtab);

/* replaces: JunctionMap.__init__ from a junction tree (junction_tree.py:12-76).
 * edges [count][p-1][2] int32, bits [count][p][W] uint64 */
int pdg_set_state(pdg_ctx *ctx, int64_t chain_begin, int64_t chain_count, const void *edges,
                  const void *bits);
int pdg_get_state(pdg_ctx *ctx, int64_t chain_begin, int64_t chain_count, void *edges, void *bits);
/* every chain starts from the same decomposable graph given by its maximal cliques
 * (mh_parallel.py:190-192), or from p singletons when bits == NULL (:194-200); follow with
 * pdg_randomize(ctx, 0) to draw the per-chain junction tree + labelling. */
int pdg_set_cliques(pdg_ctx *ctx, const void *bits, int n_cliques);
/* replaces: JunctionMap.randomize_by_jt (junction_tree.py:235-241) for every chain */
int pdg_randomize(pdg_ctx *ctx, uint32_t iter);
/* replaces: logl[0] (mh_parallel.py:59-60 / :217-218) */
int pdg_init_logl(pdg_ctx *ctx, int kind);
/* replaces: the loop body of sample_trajectory / sample_trajectory_single_move for MCMC
 * indices [it_begin, it_end) with no randomisation inside; replay == NULL -> Philox streams */
int pdg_run(pdg_ctx *ctx, int kind, int64_t it_begin, int64_t it_end, const pdg_replay *replay);
/* replaces: sample_trajectory(n_samples, randomize, ...) end to end in Philox mode:
 * init_logl, then for every block of `randomize` iterations randomise + run */
int pdg_sample(pdg_ctx *ctx, int kind, int64_t randomize);
/* waits for the context's stream and reports the first per-chain error, if any */
int pdg_sync(pdg_ctx *ctx);

/* outputs (Trajectory fields: n_updates, logl, jt_updates) */
int pdg_get_counts(pdg_ctx *ctx, void *n_proposals /*int64[n_chains]*/, void *n_accepted /*int64[n_chains]*/);
int pdg_get_logl(pdg_ctx *ctx, void *logl /* double[n_chains][n_samples] */);
/* trajectory compaction: packs the per-chain update records contiguously on the device;
 * total = number of records, offsets[n_chains+1] (nullable) */
int pdg_compact_updates(pdg_ctx *ctx, int64_t *total, void *offsets);
int pdg_get_updates(pdg_ctx *ctx, void *updates /* pdg_update[total] */, void *bits /* uint64[total][2][W] */);
/* replaces: the end of sample_trajectory (gtraj.set_logl / set_nupdates / set_jt_updates,
 * mh_parallel.py:306-311) for all chains in one go: counters, offsets, logl traces (nullable), record
 * compaction and -- when `tallies` is given -- the edge tallies, with the logl copy on a second stream
 * overlapping the compaction and tally kernels.  On return `total`, the counters and the offsets are
 * valid; the logl / tallies copies may still be in flight: the pdg_get_updates call that follows
 * (buffers sized from `total`) waits for everything. */
int pdg_collect(pdg_ctx *ctx, int64_t *total, void *n_proposals, void *n_accepted, void *offsets, void *logl,
                int64_t burnin, void *tallies);
/* per-proposal diagnostics of the LAST pdg_run (only when enabled): k-indexed, chain-major
 * [n_chains][cap]: dlik, dprior, logq, u (fp64) and acc (int32) */
int pdg_enable_move_log(pdg_ctx *ctx, int64_t cap_per_chain);
int pdg_get_move_log(pdg_ctx *ctx, void *dlik, void *dprior, void *logq, void *u, void *acc, void *U);

/* replaces: sd.log_likelihood_partial([c], {}) for a batch of complete sets
 * (ggm:42-86 / loglin:39-70): bits [n_sets][W] -> scores[n_sets] */
int pdg_score_sets(pdg_ctx *ctx, const void *bits, int64_t n_sets, void *scores);
/* edge-inclusion tallies (trajectory.py:126-152 + empirical_graph_distribution.py:35-41):
 * tallies[p][p] int64 += sum over chains and MCMC indices >= burnin of the edge indicator */
int pdg_edge_tallies(pdg_ctx *ctx, int64_t burnin, void *tallies);

/* on-device trajectory post-processing for every chain, from the per-chain update records:
 * sizes[n_chains][n_samples] int32 = number of edges at every list position (trajectory.py:202-211
 * size(), positions as in trajectory.py:126-152) */
int pdg_size_series(pdg_ctx *ctx, void *sizes);
/* edge-marginal effective sample size per chain (BASELINE.json's second metric): over the list
 * positions >= burnin, for every edge with lo < marginal < hi the indicator series' autocorrelation
 * r(h) (auxiliary_functions.py:546-550), ESS = n / (-1 + 2 sum of the initial positive pairs
 * r(2t) + r(2t+1)); ess[n_chains] fp64 = median over those edges (NaN when none), n_edges[n_chains]
 * int32 (nullable) = how many qualified */
int pdg_edge_ess(pdg_ctx *ctx, int64_t burnin, double lo, double hi, void *ess, void *n_edges);

/* measurement support (bench.py): CUDA-event timing of every MH-kernel and randomisation launch
 * on the context's stream, and the algorithmic work the MH kernel performed:
 * for the complete sets actually EVALUATED (memo misses): work[0] = bytes (GGM 8 k^2, log-linear
 * n k per set), work[1] = 6 x flops (GGM 2k^3 + 3k^2 + 6k per set), work[2] = sets, work[3] = sum
 * of k; for the sets REQUESTED by proposals (hits + misses): work[4] = sets, work[5] = sum of k;
 * work[6] = log-linear sets counted through the hashed-cell pool, work[7] = log-linear sets whose
 * score was read off the count table of a superset scanned in the same pass;
 * work[8..15] = developer phase timers (zero unless built with -DPDG_PHASE_TIMERS) */
int pdg_enable_timing(pdg_ctx *ctx, int on);
int pdg_get_timing(pdg_ctx *ctx, double *ms_run, double *ms_randomize, int64_t *n_run, int64_t *n_randomize);
int pdg_get_work(pdg_ctx *ctx, int64_t work[16]);

/* page-locked host memory for the output buffers (logl traces and update records are hundreds of
 * MB per run; pinned destinations make pdg_get_* a single DMA).  Not tied to a context. */
int pdg_host_alloc(void **ptr, int64_t bytes);
int pdg_host_free(void *ptr);

/* machine ceilings for the roofline of the scoring kernels, measured on `device` (bench.py):
 * fp64 DFMA throughput (GFLOP/s) for the log-det path (wishart.py:17-37); L2-resident and HBM read
 * bandwidth (GB/s) and shared-memory reduction rates (G reductions/s) for the contingency-table
 * scans (auxiliary_functions.py:134-151).  Not tied to a context. */
enum pdg_probe_kind { PDG_PROBE_FP64 = 0, PDG_PROBE_L2_READ = 1, PDG_PROBE_HBM_READ = 2,
                      PDG_PROBE_RED_SHARED_FREE = 3, PDG_PROBE_RED_SHARED_RANDOM = 4 };
int pdg_probe(int device, int which, double *out);

/* number of CUDA kernels this context has launched so far (bench.py's gpu_launches) */
int64_t pdg_launch_count(pdg_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
