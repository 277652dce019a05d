/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the parallelDG parallel
 * Metropolis-Hastings hot path.  Nothing under paralleldg_b200/ may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU baseline.
 *
 * Parity status: the reference ships NO golden vectors for this path; the
 * oracle is pinned instead against outputs of the reference itself, traced in the build
 * container by oracle/make_golden.py and committed under tests/golden/.
 *
 * Reference functions restated (all paths under /root/reference/parallelDG/):
 *   mh_parallel.py:182-315            sample_trajectory            -> pdgo_chain_run(kind=0)
 *   mh_parallel.py:23-180             sample_trajectory_single_move-> pdgo_chain_run(kind=1)
 *   graph/parallel_moves.py:121-142,253-272   connect proposals    -> enum_connect()
 *   graph/parallel_moves.py:169-208           disconnect proposals -> enum_disconnect()
 *   graph/parallel_moves.py:276-284           reverse-move counts  -> single-move branch
 *   graph/junction_tree.py:78-132             JunctionMap          -> pdgo_chain state
 *   graph/junction_tree.py:235-241,698-774,19-76 + graph/decomposable.py:124-150
 *                                             randomize_by_jt      -> pdgo_chain_randomize()
 *   distributions/gaussian_graphical_model.py:26-86 + wishart.py:17-37  -> score_ggm()
 *   distributions/discrete_dec_log_linear.py:12-70 + dirichlet.py:63-85
 *     + auxiliary_functions.py:134-151                             -> score_loglin()
 *   distributions/sequential_junction_tree_distributions.py:31-151 -> prior_partial()
 */
#ifndef PDG_ORACLE_H
#define PDG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PDGO_MODEL_UNIFORM = 0, PDGO_MODEL_GGM = 1, PDGO_MODEL_LOGLIN = 2 };
enum { PDGO_PRIOR_MBC = 0, PDGO_PRIOR_EDGEPENALTY = 1, PDGO_PRIOR_JUMPPENALTY = 2, PDGO_PRIOR_UNIFORM = 3 };
enum { PDGO_KIND_PARALLEL = 0, PDGO_KIND_SINGLE = 1 };

typedef struct pdgo_model {
    int32_t p;
    int32_t model;
    int32_t prior;
    int32_t pad0;
    double prior_a, prior_b;
    /* GGM */
    const double *A;     /* p*p row-major: D + X^T X  (seqdist:337, ggm:36) */
    const double *Dm;    /* p*p row-major: D */
    int64_t n;           /* rows of X */
    double delta;
    /* log-linear */
    const uint8_t *data; /* column-major: data[j*nrows + r], level codes */
    const int32_t *levels; /* p level counts */
    int64_t nrows;
    double cell_alpha;
} pdgo_model;

/* update record: one accepted move (mh_parallel.py:270,302) */
typedef struct pdgo_rec {
    int32_t i;       /* MCMC index */
    int32_t k;       /* proposal sub-index */
    int32_t type;    /* 0 connect, 1 disconnect */
    int32_t node;
    int32_t U;
    int32_t Uadj;    /* -1 = none */
} pdgo_rec;

/* replay log of one chain, arrays indexed by absolute MCMC iteration (see oracle/ref_shims.py) */
typedef struct pdgo_replay {
    const int32_t *it_node;
    const int32_t *it_mtype_raw;
    const int32_t *it_aux;
    const int64_t *it_off;   /* n_samples+1 */
    const int32_t *mv_U;
    const int32_t *mv_Uadj;
    const double *mv_u;
} pdgo_replay;

/* per-move diagnostics (optional outputs) */
typedef struct pdgo_moves_out {
    int64_t cap, count;
    int32_t *iter, *U, *Uadj, *acc;
    double *dlik, *dprior, *logq, *u;
} pdgo_moves_out;

void pdgo_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                        uint32_t k0, uint32_t k1, uint32_t out[4]);

/* score of one complete set (bitset of W=ceil(p/64) words); no memoisation */
double pdgo_score(const pdgo_model *m, const uint64_t *bits);
double pdgo_prior_partial(const pdgo_model *m, int csize, int ssize, int extra);

void *pdgo_chain_create(const pdgo_model *m, int use_cache);
void pdgo_chain_destroy(void *ch);
int pdgo_chain_set_state(void *ch, const int32_t *edges /*(p-1)*2*/, const uint64_t *bits /*p*W*/);
int pdgo_chain_get_state(void *ch, int32_t *edges, uint64_t *bits);
/* state from maximal cliques of a decomposable graph (or p singletons when ncl==p): rows of bits */
int pdgo_chain_set_cliques(void *ch, const uint64_t *bits, int ncl);
double pdgo_chain_init_logl(void *ch, int kind);
int pdgo_chain_randomize(void *ch, uint64_t seed, uint32_t chain_id, uint32_t iter);
/* runs iterations [it_begin, it_end); logl[it_begin-1] must hold the previous value.
 * kcount in/out = running proposal counter k.  returns 0 ok, <0 error (replay mismatch...) */
int pdgo_chain_run(void *ch, int kind, int64_t it_begin, int64_t it_end,
                   const pdgo_replay *replay, uint64_t seed, uint32_t chain_id,
                   double *logl, int64_t *kcount,
                   pdgo_rec *recs, uint64_t *rec_bits /* per rec: C[W], Cadj[W] */,
                   int64_t rec_cap, int64_t *n_rec, pdgo_moves_out *mv);

/* Philox production mode, many chains, pthreads: the CPU baseline.
 * init: all-singletons + randomize at iteration 0.  Outputs may be NULL except counters.
 * n_prop[c], n_acc[c] per chain; logl_out (n_chains*n_samples) optional;
 * recs/rec_bits with per-chain capacity rec_cap optional; final state optional */
int pdgo_run_many(const pdgo_model *m, int kind, int64_t n_chains, uint32_t chain0,
                  int64_t n_samples, int64_t randomize, uint64_t seed, int n_threads,
                  int64_t *n_prop, int64_t *n_acc, double *logl_out,
                  pdgo_rec *recs, uint64_t *rec_bits, int64_t rec_cap,
                  int32_t *final_edges, uint64_t *final_bits);

/* generalised-Pruefer reconnection of a forest (junction_tree.py:712-774) with injected draws:
 * comp_of[P] (component index of each node, nodes listed in canonical order), v_ind[q-2], w_j[q];
 * out_edges[(q-1)*2] are indices into the canonical node list */
int pdgo_tree_from_forest(int P, int q, const int32_t *comp_start /*q+1*/,
                          const int32_t *v_ind, const int32_t *w_j, int32_t *out_edges);

#ifdef __cplusplus
}
#endif
#endif
