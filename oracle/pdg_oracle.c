/*
 * TEST INFRASTRUCTURE ONLY -- see pdg_oracle.h for the scope statement and the list of
 * reference functions (file:line) restated here.  Plain sequential C; written for clarity,
 * one chain at a time, python-set semantics replaced by 64-bit-word bitsets.
 */
#define _GNU_SOURCE
#include "pdg_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ Philox4x32-10 -------- */
/* Salmon et al. 2011 (Random123).  Production-mode random streams of BOTH the oracle and the
 * CUDA path are defined on this generator:  key = (seed lo, seed hi),
 * counter = (chain id, MCMC iteration, stream tag, index). */
void pdgo_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                        uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { TAG_ITER = 0, TAG_MOVE = 1, TAG_PRUF_V = 2, TAG_PRUF_W = 3, TAG_PERM = 4, TAG_PAD = 5 };

static inline uint32_t mulhi32(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

static uint32_t draw_int(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t tag, uint32_t idx, uint32_t n)
{
    uint32_t o[4];
    pdgo_philox4x32_10(chain, iter, tag, idx, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return mulhi32(o[0], n);
}

static double draw_unif(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t tag, uint32_t idx)
{
    uint32_t o[4];
    pdgo_philox4x32_10(chain, iter, tag, idx, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    /* 53-bit uniform in [0,1), same construction as numpy's legacy random_sample */
    return ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) / 9007199254740992.0;
}

/* ------------------------------------------------------------------ bitset helpers ------- */
static inline int bs_popc(const uint64_t *a, int W) { int c = 0; for (int w = 0; w < W; ++w) c += __builtin_popcountll(a[w]); return c; }
static inline int bs_test(const uint64_t *a, int i) { return (int)((a[i >> 6] >> (i & 63)) & 1u); }
static inline void bs_set(uint64_t *a, int i) { a[i >> 6] |= (uint64_t)1 << (i & 63); }
static inline void bs_clr(uint64_t *a, int i) { a[i >> 6] &= ~((uint64_t)1 << (i & 63)); }
static inline int bs_empty(const uint64_t *a, int W) { for (int w = 0; w < W; ++w) if (a[w]) return 0; return 1; }
static inline int bs_eq(const uint64_t *a, const uint64_t *b, int W) { for (int w = 0; w < W; ++w) if (a[w] != b[w]) return 0; return 1; }
static inline int bs_subset(const uint64_t *a, const uint64_t *b, int W) { for (int w = 0; w < W; ++w) if (a[w] & ~b[w]) return 0; return 1; }
static int bs_list(const uint64_t *a, int W, int *out)
{
    int k = 0;
    for (int w = 0; w < W; ++w) { uint64_t x = a[w]; while (x) { int b = __builtin_ctzll(x); out[k++] = w * 64 + b; x &= x - 1; } }
    return k;
}

/* ------------------------------------------------------------------ scorers -------------- */
/* scipy.special.multigammaln(a, d) */
static double multigammaln(double a, int d)
{
    double res = (d * (d - 1) * 0.25) * log(M_PI);
    double s = 0.0;
    for (int j = 1; j <= d; ++j) s += lgamma(a - (j - 1.0) / 2.0);
    return res + s;
}

/* log|det| by LU with partial pivoting (what numpy.linalg.slogdet does; wishart.py:34) */
static double logdet_lu(double *a, int k)
{
    double ld = 0.0;
    for (int j = 0; j < k; ++j) {
        int piv = j; double best = fabs(a[j * k + j]);
        for (int i = j + 1; i < k; ++i) if (fabs(a[i * k + j]) > best) { best = fabs(a[i * k + j]); piv = i; }
        if (piv != j) for (int c = 0; c < k; ++c) { double t = a[j * k + c]; a[j * k + c] = a[piv * k + c]; a[piv * k + c] = t; }
        double d = a[j * k + j];
        ld += log(fabs(d));
        for (int i = j + 1; i < k; ++i) {
            double f = a[i * k + j] / d;
            if (f != 0.0) for (int c = j + 1; c < k; ++c) a[i * k + c] -= f * a[j * k + c];
        }
    }
    return ld;
}

/* wishart.py:17-37 log_norm_constant(D, delta) on the k x k sub-block idx of mat */
static double wishart_lnc(const double *mat, int p, const int *idx, int k, double delta, double *work)
{
    double t = (delta + k - 1.0) / 2.0;
    double K = -t * k * log(M_PI);
    K += multigammaln(t, k);
    for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) work[i * k + j] = mat[(size_t)idx[i] * p + idx[j]];
    K -= t * logdet_lu(work, k);
    return K;
}

/* gaussian_graphical_model.py:26-39,64-72 */
static double score_ggm(const pdgo_model *m, const int *idx, int k)
{
    if (k == 0) return 0.0;
    double *work = (double *)malloc(sizeof(double) * (size_t)k * k);
    double c1 = wishart_lnc(m->A, m->p, idx, k, m->delta + (double)m->n, work);
    double c2 = wishart_lnc(m->Dm, m->p, idx, k, m->delta, work);
    free(work);
    return c1 - c2;
}

/* discrete_dec_log_linear.py:12-36,48-55 + dirichlet.py:63-85 + auxiliary_functions.py:134-151.
 * Cells are found by successive refinement (group id x level) so no cell code can overflow;
 * the final pass relabels in first-seen row order = python dict insertion order. */
static double score_loglin(const pdgo_model *m, const int *idx, int k)
{
    if (k == 0) return 0.0;
    int64_t n = m->nrows;
    /* alpha rule loglin:48-55: cell_alpha * prod(levels outside) / prod(all levels); evaluated
     * the reference's way while the products fit int64, else (p >= 63 binary)
     * by the algebraically equal cell_alpha / prod(levels inside). */
    double Kin = 1.0; long double tot = 1.0L; int fits = 1;
    int64_t tot_i = 1, out_i = 1;
    for (int j = 0; j < m->p; ++j) {
        tot *= m->levels[j];
        if (tot > 9.0e18L) fits = 0; else tot_i *= m->levels[j];
    }
    for (int j = 0; j < k; ++j) Kin *= (double)m->levels[idx[j]];
    double alpha;
    if (fits) {
        int jj = 0;
        for (int j = 0; j < m->p; ++j) { if (jj < k && idx[jj] == j) { ++jj; continue; } out_i *= m->levels[j]; }
        alpha = m->cell_alpha * (double)out_i / (double)tot_i;
    } else {
        alpha = m->cell_alpha / Kin;
    }
    int32_t *gid = (int32_t *)calloc((size_t)n, sizeof(int32_t));
    int64_t G = 1;
    int32_t *map = NULL; size_t mapcap = 0;
    for (int j = 0; j < k; ++j) {
        int L = m->levels[idx[j]];
        size_t need = (size_t)G * (size_t)L;
        if (need > mapcap) { free(map); map = (int32_t *)malloc(need * sizeof(int32_t)); mapcap = need; }
        for (size_t q = 0; q < need; ++q) map[q] = -1;
        const uint8_t *col = m->data + (size_t)idx[j] * (size_t)n;
        int64_t Gn = 0;
        for (int64_t r = 0; r < n; ++r) {
            size_t key = (size_t)gid[r] * (size_t)L + col[r];
            if (map[key] < 0) map[key] = (int32_t)Gn++;
            gid[r] = map[key];
        }
        G = Gn;
    }
    int64_t *cnt = (int64_t *)calloc((size_t)G, sizeof(int64_t));
    for (int64_t r = 0; r < n; ++r) cnt[gid[r]]++;
    /* c1 = lnc(counts, alpha), c2 = lnc({}, alpha) */
    double lg_a = lgamma(alpha);
    double num1 = (Kin - (double)G) * lg_a;
    int64_t N = 0;
    for (int64_t g = 0; g < G; ++g) { num1 += lgamma((double)cnt[g] + alpha); N += cnt[g]; }
    double c1 = num1 - lgamma(Kin * alpha + (double)N);
    double c2 = Kin * lg_a - lgamma(Kin * alpha + 0.0);
    free(cnt); free(map); free(gid);
    return c1 - c2;
}

static double score_idx(const pdgo_model *m, const int *idx, int k)
{
    switch (m->model) {
    case PDGO_MODEL_GGM: return score_ggm(m, idx, k);
    case PDGO_MODEL_LOGLIN: return score_loglin(m, idx, k);
    default: return 0.0;
    }
}

double pdgo_score(const pdgo_model *m, const uint64_t *bits)
{
    int W = (m->p + 63) / 64;
    int *idx = (int *)malloc(sizeof(int) * (size_t)(m->p + 1));
    int k = bs_list(bits, W, idx);
    double s = score_idx(m, idx, k);
    free(idx);
    return s;
}

/* sequential_junction_tree_distributions.py:31-151 log_prior_partial(clq, sep, extra_arg) */
double pdgo_prior_partial(const pdgo_model *m, int csize, int ssize, int extra)
{
    switch (m->prior) {
    case PDGO_PRIOR_MBC: {
        double lp = 0.0;
        lp += m->prior_a * ((double)csize - 1.0);
        lp -= m->prior_b * (double)ssize;
        return lp;
    }
    case PDGO_PRIOR_EDGEPENALTY: {
        double a = (double)csize, b = (double)ssize;
        return (-m->prior_a * a * (a - 1.0) / 2.0) - (-m->prior_a * b * (b - 1.0) / 2.0);
    }
    case PDGO_PRIOR_JUMPPENALTY:
        return -m->prior_a * (1 - extra);
    default:
        return 0.0;
    }
}

/* ------------------------------------------------------------------ chain state ---------- */
typedef struct {
    uint64_t *keys; double *vals; uint8_t *used; size_t cap, count; int W;
} score_cache;

typedef struct {
    const pdgo_model *m;
    int p, W;
    uint64_t *M;     /* t2clique: p rows of W words */
    uint64_t *N;     /* node2t  : p rows of W words */
    int32_t *edges;  /* (p-1)*2 */
    int32_t *adj_off, *adj; /* CSR of the tree */
    int use_cache;
    score_cache cache;
    int *idx;        /* scratch p+1 */
    int64_t n_score_calls, n_score_miss;
} chain_t;

static uint64_t hash_bits(const uint64_t *b, int W)
{
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int w = 0; w < W; ++w) { h ^= b[w]; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; }
    return h;
}

static void cache_init(score_cache *c, int W) { c->W = W; c->cap = 1024; c->count = 0; c->keys = (uint64_t *)calloc(c->cap * W, 8); c->vals = (double *)calloc(c->cap, 8); c->used = (uint8_t *)calloc(c->cap, 1); }
static void cache_free(score_cache *c) { free(c->keys); free(c->vals); free(c->used); }
static double *cache_find(score_cache *c, const uint64_t *b, int insert);
static void cache_grow(score_cache *c)
{
    score_cache n; n.W = c->W; n.cap = c->cap * 2; n.count = 0;
    n.keys = (uint64_t *)calloc(n.cap * n.W, 8); n.vals = (double *)calloc(n.cap, 8); n.used = (uint8_t *)calloc(n.cap, 1);
    for (size_t i = 0; i < c->cap; ++i) if (c->used[i]) { double *v = cache_find(&n, c->keys + i * c->W, 1); *v = c->vals[i]; }
    cache_free(c); *c = n;
}
static double *cache_find(score_cache *c, const uint64_t *b, int insert)
{
    size_t mask = c->cap - 1, i = (size_t)hash_bits(b, c->W) & mask;
    while (c->used[i]) { if (bs_eq(c->keys + i * c->W, b, c->W)) return &c->vals[i]; i = (i + 1) & mask; }
    if (!insert) return NULL;
    if ((c->count + 1) * 2 > c->cap) { cache_grow(c); return cache_find(c, b, 1); }
    c->used[i] = 1; memcpy(c->keys + i * c->W, b, 8 * (size_t)c->W); c->count++; c->vals[i] = NAN;
    return &c->vals[i];
}

/* memoised score of a complete set, as ggm:63-72 / loglin:21-36 do per chain */
static double chain_score(chain_t *ch, const uint64_t *bits)
{
    ch->n_score_calls++;
    if (ch->m->model == PDGO_MODEL_UNIFORM) return 0.0;
    if (bs_empty(bits, ch->W)) return 0.0;
    if (ch->use_cache) {
        double *v = cache_find(&ch->cache, bits, 0);
        if (v) return *v;
    }
    ch->n_score_miss++;
    int k = bs_list(bits, ch->W, ch->idx);
    double s = score_idx(ch->m, ch->idx, k);
    if (ch->use_cache) { double *v = cache_find(&ch->cache, bits, 1); *v = s; }
    return s;
}

void *pdgo_chain_create(const pdgo_model *m, int use_cache)
{
    chain_t *ch = (chain_t *)calloc(1, sizeof(chain_t));
    ch->m = m; ch->p = m->p; ch->W = (m->p + 63) / 64;
    int p = ch->p, W = ch->W;
    ch->M = (uint64_t *)calloc((size_t)p * W, 8);
    ch->N = (uint64_t *)calloc((size_t)p * W, 8);
    ch->edges = (int32_t *)calloc((size_t)(p > 1 ? p - 1 : 1) * 2, 4);
    ch->adj_off = (int32_t *)calloc((size_t)p + 1, 4);
    ch->adj = (int32_t *)calloc((size_t)2 * p, 4);
    ch->idx = (int *)calloc((size_t)p + 1, sizeof(int));
    ch->use_cache = use_cache;
    cache_init(&ch->cache, W);
    return ch;
}

void pdgo_chain_destroy(void *h)
{
    chain_t *ch = (chain_t *)h;
    if (!ch) return;
    free(ch->M); free(ch->N); free(ch->edges); free(ch->adj_off); free(ch->adj); free(ch->idx);
    cache_free(&ch->cache);
    free(ch);
}

static void build_csr(chain_t *ch)
{
    int p = ch->p;
    memset(ch->adj_off, 0, sizeof(int32_t) * (size_t)(p + 1));
    for (int e = 0; e < p - 1; ++e) { ch->adj_off[ch->edges[2 * e] + 1]++; ch->adj_off[ch->edges[2 * e + 1] + 1]++; }
    for (int t = 0; t < p; ++t) ch->adj_off[t + 1] += ch->adj_off[t];
    int32_t *fill = (int32_t *)calloc((size_t)p, 4);
    for (int e = 0; e < p - 1; ++e) {
        int a = ch->edges[2 * e], b = ch->edges[2 * e + 1];
        ch->adj[ch->adj_off[a] + fill[a]++] = b;
        ch->adj[ch->adj_off[b] + fill[b]++] = a;
    }
    free(fill);
}

static void build_node2t(chain_t *ch)
{
    int p = ch->p, W = ch->W;
    memset(ch->N, 0, 8 * (size_t)p * W);
    for (int t = 0; t < p; ++t)
        for (int g = 0; g < p; ++g)
            if (bs_test(ch->M + (size_t)t * W, g)) bs_set(ch->N + (size_t)g * W, t);
}

int pdgo_chain_set_state(void *h, const int32_t *edges, const uint64_t *bits)
{
    chain_t *ch = (chain_t *)h;
    int p = ch->p, W = ch->W;
    memcpy(ch->edges, edges, sizeof(int32_t) * 2 * (size_t)(p - 1));
    memcpy(ch->M, bits, 8 * (size_t)p * W);
    build_csr(ch);
    build_node2t(ch);
    return 0;
}

int pdgo_chain_get_state(void *h, int32_t *edges, uint64_t *bits)
{
    chain_t *ch = (chain_t *)h;
    memcpy(edges, ch->edges, sizeof(int32_t) * 2 * (size_t)(ch->p - 1));
    memcpy(bits, ch->M, 8 * (size_t)ch->p * ch->W);
    return 0;
}

int pdgo_chain_set_cliques(void *h, const uint64_t *bits, int ncl)
{
    chain_t *ch = (chain_t *)h;
    int p = ch->p, W = ch->W;
    if (ncl > p) return -1;
    memset(ch->M, 0, 8 * (size_t)p * W);
    memcpy(ch->M, bits, 8 * (size_t)ncl * W);
    for (int e = 0; e < p - 1; ++e) { ch->edges[2 * e] = e; ch->edges[2 * e + 1] = e + 1; }
    build_csr(ch);
    build_node2t(ch);
    return 0;
}

/* logl[0]  (mh_parallel.py:59-60 / 217-218): sum of clique scores (t2clique order, empties
 * skipped, junction_tree.py:102-108) minus nu_s * score(s) over DISTINCT non-empty separators
 * (junction_tree.py:110-132, ggm:73-82) plus log_prior over cliques and DISTINCT separators
 * (seqdist:44-52). */
double pdgo_chain_init_logl(void *h, int kind)
{
    chain_t *ch = (chain_t *)h;
    const pdgo_model *m = ch->m;
    int p = ch->p, W = ch->W;
    double cl = 0.0, lclq = 0.0;
    for (int t = 0; t < p; ++t) {
        const uint64_t *c = ch->M + (size_t)t * W;
        if (bs_empty(c, W)) continue;
        cl += chain_score(ch, c);
        int a = bs_popc(c, W);
        if (m->prior == PDGO_PRIOR_MBC) lclq += m->prior_a * ((double)a - 1.0);
        else if (m->prior == PDGO_PRIOR_EDGEPENALTY) lclq += -m->prior_a * a * (a - 1.0) / 2.0;
    }
    int E = p - 1;
    uint64_t *seps = (uint64_t *)calloc((size_t)(E > 0 ? E : 1) * W, 8);
    int *nu = (int *)calloc((size_t)(E > 0 ? E : 1), sizeof(int));
    int nd = 0;
    for (int e = 0; e < E; ++e) {
        uint64_t s[64];
        const uint64_t *a = ch->M + (size_t)ch->edges[2 * e] * W, *b = ch->M + (size_t)ch->edges[2 * e + 1] * W;
        for (int w = 0; w < W; ++w) s[w] = a[w] & b[w];
        int f = -1;
        for (int d = 0; d < nd; ++d) if (bs_eq(seps + (size_t)d * W, s, W)) { f = d; break; }
        if (f < 0) { memcpy(seps + (size_t)nd * W, s, 8 * (size_t)W); nu[nd] = 1; nd++; } else nu[f]++;
    }
    double sc = 0.0, lsep = 0.0;
    for (int d = 0; d < nd; ++d) {
        const uint64_t *s = seps + (size_t)d * W;
        int b = bs_popc(s, W);
        if (m->prior == PDGO_PRIOR_MBC) lsep += m->prior_b * (double)b;
        else if (m->prior == PDGO_PRIOR_EDGEPENALTY) lsep += -m->prior_a * b * (b - 1.0) / 2.0;
        if (b == 0) continue;
        sc += nu[d] * chain_score(ch, s);
    }
    free(seps); free(nu);
    double prior;
    if (m->prior == PDGO_PRIOR_JUMPPENALTY) prior = -m->prior_a * (1 - (kind == PDGO_KIND_SINGLE ? 0 : 1));
    else if (m->prior == PDGO_PRIOR_UNIFORM) prior = 0.0;
    else prior = lclq - lsep;
    return (cl - sc) + prior;
}

/* ------------------------------------------------------------------ move enumeration ----- */
typedef struct { int U, Uadj; } move_t;

/* parallel_moves.py:253-272: every (U, Uadj) with Uadj in subtree, U adjacent, U outside.
 * Canonical (production) order: ascending U. */
static int enum_connect(chain_t *ch, const uint64_t *sub, move_t *mv)
{
    int p = ch->p, n = 0;
    for (int U = 0; U < p; ++U) {
        if (bs_test(sub, U)) continue;
        for (int e = ch->adj_off[U]; e < ch->adj_off[U + 1]; ++e) {
            int a = ch->adj[e];
            if (bs_test(sub, a)) { mv[n].U = U; mv[n].Uadj = a; n++; }
        }
    }
    return n;
}

/* parallel_moves.py:193-208: leaves of the induced subtree with their in-subtree neighbour.
 * Canonical order: ascending U. */
static int enum_leaves(chain_t *ch, const uint64_t *sub, move_t *mv)
{
    int p = ch->p, n = 0;
    for (int U = 0; U < p; ++U) {
        if (!bs_test(sub, U)) continue;
        int cnt = 0, last = -1;
        for (int e = ch->adj_off[U]; e < ch->adj_off[U + 1]; ++e) if (bs_test(sub, ch->adj[e])) { cnt++; last = ch->adj[e]; }
        if (cnt == 1) { mv[n].U = U; mv[n].Uadj = last; n++; }
    }
    return n;
}

static void mv_push(pdgo_moves_out *o, int it, int U, int Uadj, int acc, double dlik, double dprior, double logq, double u)
{
    if (!o) return;
    if (o->count < o->cap) {
        int64_t j = o->count;
        o->iter[j] = it; o->U[j] = U; o->Uadj[j] = Uadj; o->acc[j] = acc;
        o->dlik[j] = dlik; o->dprior[j] = dprior; o->logq[j] = logq; o->u[j] = u;
    }
    o->count++;
}

int pdgo_chain_run(void *h, int kind, int64_t it_begin, int64_t it_end,
                   const pdgo_replay *rp, uint64_t seed, uint32_t chain_id,
                   double *logl, int64_t *kcount,
                   pdgo_rec *recs, uint64_t *rec_bits, int64_t rec_cap, int64_t *n_rec,
                   pdgo_moves_out *mvout)
{
    chain_t *ch = (chain_t *)h;
    const pdgo_model *m = ch->m;
    int p = ch->p, W = ch->W;
    move_t *mv = (move_t *)malloc(sizeof(move_t) * (size_t)(2 * p + 2));
    move_t *lv = (move_t *)malloc(sizeof(move_t) * (size_t)(2 * p + 2));
    uint64_t C[64], Cadj[64], Cnew[64], S[64], Snew[64], one[64], sub[64];
    int rc = 0;
    int64_t k = *kcount;
    for (int64_t i = it_begin; i < it_end; ++i) {
        int node, mtype, aux = -1; uint32_t r[4] = {0, 0, 0, 0};
        if (rp) { node = rp->it_node[i]; mtype = rp->it_mtype_raw[i]; aux = rp->it_aux[i]; }
        else {
            pdgo_philox4x32_10(chain_id, (uint32_t)i, TAG_ITER, 0, (uint32_t)seed, (uint32_t)(seed >> 32), r);
            node = (int)mulhi32(r[0], (uint32_t)p);
            mtype = (int)(r[1] >> 31);
        }
        memcpy(sub, ch->N + (size_t)node * W, 8 * (size_t)W);
        int msub = bs_popc(sub, W);
        if (msub == 0) mtype = 0;      /* mh_parallel.py:237-240 */
        if (msub == p) mtype = 1;
        double log_p = 0.0;
        int nmv = 0; double logq_par = 0.0;
        if (mtype == 0) {
            if (msub > 0) nmv = enum_connect(ch, sub, mv);
            else {
                /* parallel_moves.py:138-139: one uniformly chosen tree node */
                int U = rp ? aux : (int)mulhi32(r[2], (uint32_t)p);
                mv[0].U = U; mv[0].Uadj = -1; nmv = 1;
            }
        } else {
            nmv = enum_leaves(ch, sub, mv);
            if (msub == 1) { int only = -1; for (int t = 0; t < p; ++t) if (bs_test(sub, t)) only = t; mv[0].U = only; mv[0].Uadj = -1; nmv = 1; }
            else if (msub == 2) {
                /* parallel_moves.py:184-185: ONE of the two leaves */
                int pick;
                if (rp) { pick = (mv[0].U == aux) ? 0 : 1; if (mv[pick].U != aux) { rc = -10; goto done; } }
                else pick = (int)(r[2] >> 31);
                mv[0] = mv[pick]; nmv = 1;
                logq_par = -log(2.0);
            }
        }
        int n_forward = (mtype == 1 && msub == 2) ? 2 : nmv;   /* num_moves, pm:185 returns 2 */
        double logq_single = 0.0;
        if (kind == PDGO_KIND_SINGLE) {
            /* choose moves[0] after np.random.shuffle (mh_parallel.py:87,95 / 128,138) */
            int pick = 0;
            if (rp) {
                int64_t o = rp->it_off[i];
                if (rp->it_off[i + 1] - o != 1) { rc = -11; goto done; }
                pick = -1;
                for (int j = 0; j < nmv; ++j) if (mv[j].U == rp->mv_U[o] && mv[j].Uadj == rp->mv_Uadj[o]) pick = j;
                if (pick < 0) { rc = -12; goto done; }
            } else if (nmv > 1) pick = (int)mulhi32(r[3], (uint32_t)nmv);
            move_t chosen = mv[pick];
            int n_reverse;
            if (mtype == 0) {
                /* mh_parallel.py:92-94 + parallel_moves.py:276-281 */
                int nl = (msub >= 2) ? enum_leaves(ch, sub, lv) : 0;
                int notin = 1;
                for (int j = 0; j < nl; ++j) if (lv[j].U == chosen.Uadj) notin = 0;
                n_reverse = nl + notin + (msub == 1 ? 1 : 0);
            } else {
                /* mh_parallel.py:133-137 + parallel_moves.py:283-284 */
                if (msub == 1) n_reverse = p;
                else { int nb = enum_connect(ch, sub, lv); n_reverse = nb + 2 - (ch->adj_off[chosen.U + 1] - ch->adj_off[chosen.U]); }
            }
            double log_q2 = -log((double)n_forward), log_q1 = -log((double)n_reverse);
            logq_single = log_q2 - log_q1;
            mv[0] = chosen; nmv = 1;
        } else if (rp) {
            /* take the reference's move ORDER (python set iteration), verify the SET */
            int64_t o = rp->it_off[i], cnt = rp->it_off[i + 1] - o;
            if (cnt != nmv) { rc = -13; goto done; }
            for (int j = 0; j < nmv; ++j) {
                int f = -1;
                for (int q = 0; q < nmv; ++q) if (mv[q].U == rp->mv_U[o + j] && mv[q].Uadj == rp->mv_Uadj[o + j]) f = q;
                if (f < 0) { rc = -14; goto done; }
                lv[j] = mv[f];
            }
            memcpy(mv, lv, sizeof(move_t) * (size_t)nmv);
        }
        for (int j = 0; j < nmv; ++j) {
            int U = mv[j].U, Uadj = mv[j].Uadj;
            memcpy(C, ch->M + (size_t)U * W, 8 * (size_t)W);
            if (Uadj >= 0) memcpy(Cadj, ch->M + (size_t)Uadj * W, 8 * (size_t)W); else memset(Cadj, 0, 8 * (size_t)W);
            memcpy(Cnew, C, 8 * (size_t)W);
            if (mtype == 0) bs_set(Cnew, node); else bs_clr(Cnew, node);
            for (int w = 0; w < W; ++w) { S[w] = C[w] & Cadj[w]; Snew[w] = Cnew[w] & Cadj[w]; }
            memset(one, 0, 8 * (size_t)W); bs_set(one, node);
            int cC = bs_popc(C, W), cS = bs_popc(S, W), cCn = bs_popc(Cnew, W), cSn = bs_popc(Snew, W);
            double log_p2 = chain_score(ch, Cnew) - (cSn ? 1 * chain_score(ch, Snew) : 0.0);
            double log_p1 = chain_score(ch, C) - (cS ? 1 * chain_score(ch, S) : 0.0);
            double log_g2, log_g1, logq;
            if (mtype == 0) {
                log_g2 = pdgo_prior_partial(m, cCn, cSn, 0);
                log_g1 = pdgo_prior_partial(m, cC, cS, 1);
                if (msub == 0) { log_p1 += chain_score(ch, one); log_g1 += pdgo_prior_partial(m, 1, cS, 1); }
            } else {
                log_g2 = pdgo_prior_partial(m, cCn, cSn, 1);
                log_g1 = pdgo_prior_partial(m, cC, cS, 0);
                if (msub == 1) { log_p2 += chain_score(ch, one); log_g2 += pdgo_prior_partial(m, 1, cSn, 1); }
            }
            logq = (kind == PDGO_KIND_SINGLE) ? logq_single : (mtype == 1 ? logq_par : 0.0);
            double dlik = log_p2 - log_p1, dpri = log_g2 - log_g1;
            double tot = (dlik + dpri) + logq;
            double u = rp ? rp->mv_u[rp->it_off[i] + j] : draw_unif(seed, chain_id, (uint32_t)i, TAG_MOVE, (uint32_t)U);
            k += 1;
            double e = exp(tot);
            int acc = (u <= (e < 1.0 ? e : 1.0));
            mv_push(mvout, (int)i, U, Uadj, acc, dlik, dpri, logq, u);
            if (acc) {
                log_p += dlik + dpri;
                if (recs && *n_rec < rec_cap) {
                    pdgo_rec *rr = &recs[*n_rec];
                    rr->i = (int32_t)i; rr->k = (int32_t)k; rr->type = mtype; rr->node = node; rr->U = U; rr->Uadj = Uadj;
                    if (rec_bits) { memcpy(rec_bits + (size_t)(*n_rec) * 2 * W, C, 8 * (size_t)W); memcpy(rec_bits + (size_t)(*n_rec) * 2 * W + W, Cadj, 8 * (size_t)W); }
                }
                (*n_rec)++;
                if (mtype == 0) { bs_set(ch->M + (size_t)U * W, node); bs_set(ch->N + (size_t)node * W, U); }
                else { bs_clr(ch->M + (size_t)U * W, node); bs_clr(ch->N + (size_t)node * W, U); }
            }
        }
        logl[i] = logl[i - 1] + log_p;
    }
done:
    *kcount = k;
    free(mv); free(lv);
    return rc;
}

/* ------------------------------------------------------------------ randomisation -------- */
/* junction_tree.py:712-774 random_tree_from_forest with the random draws injected.
 * Nodes 0..P-1 in canonical order, component i = [comp_start[i], comp_start[i+1]). */
int pdgo_tree_from_forest(int P, int q, const int32_t *comp_start, const int32_t *v_ind,
                          const int32_t *w_j, int32_t *out_edges)
{
    if (q < 2) return 0;
    int32_t *comp_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    for (int i = 0; i < q; ++i) for (int a = comp_start[i]; a < comp_start[i + 1]; ++a) comp_of[a] = i;
    int32_t *cnt = (int32_t *)calloc((size_t)q, 4);
    uint8_t *inw = (uint8_t *)malloc((size_t)q);
    memset(inw, 1, (size_t)q);
    for (int t = 0; t < q - 2; ++t) cnt[comp_of[v_ind[t]]]++;
    int ptr = q - 1, pending = -1, ne = 0;
    for (int t = q - 3; t >= 0; --t) {           /* y = v.pop(): last element first */
        int x;
        if (pending >= 0) { x = pending; pending = -1; }
        else { while (!(inw[ptr] && cnt[ptr] == 0)) ptr--; x = ptr; }
        int y = v_ind[t], cy = comp_of[y];
        out_edges[2 * ne] = comp_start[x] + w_j[x]; out_edges[2 * ne + 1] = y; ne++;
        inw[x] = 0;
        cnt[cy]--;
        if (cnt[cy] == 0 && inw[cy] && cy > ptr) pending = cy;
    }
    int a = -1, b = -1;
    for (int i = 0; i < q; ++i) if (inw[i]) { if (a < 0) a = i; else if (b < 0) b = i; }
    out_edges[2 * ne] = comp_start[a] + w_j[a]; out_edges[2 * ne + 1] = comp_start[b] + w_j[b]; ne++;
    free(comp_of); free(cnt); free(inw);
    return ne;
}

static int cmp_i32(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return (x > y) - (x < y); }

/* JunctionMap.randomize_by_jt (junction_tree.py:235-241):
 *   to_graph (:201-219)  -> graph adjacency from the cliques
 *   dlib.junction_tree (decomposable.py:124-150) -> maximal cliques + SOME junction tree; restated
 *        with maximum-cardinality search (Blair & Peyton 1993, fig. 4) because the uniform
 *        randomisation below makes the starting junction tree irrelevant
 *   jtlib.randomize (:698-710): for every distinct separator cut and reconnect uniformly
 *        (:670-694, :522-541, :712-774).  The component sets of separator X do not depend on how
 *        other separators were reconnected (Thomas & Green 2009), so every separator is handled
 *        from the ORIGINAL tree.
 *   create_t_and_t2clique (:19-65): random relabelling onto p tree nodes + empty-node padding.
 * Random draws come from Philox streams (see tags), the canonical orders are ascending ids. */
int pdgo_chain_randomize(void *h, uint64_t seed, uint32_t chain_id, uint32_t iter)
{
    chain_t *ch = (chain_t *)h;
    int p = ch->p, W = ch->W;
    uint64_t *adj = (uint64_t *)calloc((size_t)p * W, 8);
    for (int t = 0; t < p; ++t) {
        const uint64_t *c = ch->M + (size_t)t * W;
        int k = bs_list(c, W, ch->idx);
        for (int a = 0; a < k; ++a) for (int w = 0; w < W; ++w) adj[(size_t)ch->idx[a] * W + w] |= c[w];
    }
    for (int v = 0; v < p; ++v) bs_clr(adj + (size_t)v * W, v);
    /* --- maximum cardinality search -> cliques K[0..S), parent[], sep[] --- */
    uint64_t *K = (uint64_t *)calloc((size_t)p * W, 8);
    uint64_t *sep = (uint64_t *)calloc((size_t)p * W, 8);
    uint64_t *numbered = (uint64_t *)calloc((size_t)W, 8);
    int32_t *card = (int32_t *)calloc((size_t)p, 4), *order = (int32_t *)calloc((size_t)p, 4);
    int32_t *cliqueof = (int32_t *)calloc((size_t)p, 4), *parent = (int32_t *)calloc((size_t)p, 4);
    int S = 0, prev_card = 0;
    for (int step = 0; step < p; ++step) {
        int v = -1, best = -1;
        for (int u = 0; u < p; ++u) if (!bs_test(numbered, u) && card[u] > best) { best = card[u]; v = u; }
        int new_card = best;
        if (new_card <= prev_card) {
            uint64_t *Ks = K + (size_t)S * W;
            for (int w = 0; w < W; ++w) Ks[w] = adj[(size_t)v * W + w] & numbered[w];
            memcpy(sep + (size_t)S * W, Ks, 8 * (size_t)W);
            if (new_card != 0) {
                int k = bs_list(Ks, W, ch->idx), last = ch->idx[0];
                for (int a = 1; a < k; ++a) if (order[ch->idx[a]] > order[last]) last = ch->idx[a];
                parent[S] = cliqueof[last];
            } else parent[S] = (S == 0) ? -1 : 0;
            S++;
        }
        cliqueof[v] = S - 1; bs_set(K + (size_t)(S - 1) * W, v); bs_set(numbered, v); order[v] = step; prev_card = new_card;
        int k = bs_list(adj + (size_t)v * W, W, ch->idx);
        for (int a = 0; a < k; ++a) if (!bs_test(numbered, ch->idx[a])) card[ch->idx[a]]++;
    }
    /* --- children lists of the clique tree --- */
    int32_t *coff = (int32_t *)calloc((size_t)S + 1, 4), *cl = (int32_t *)calloc((size_t)S + 1, 4), *cf = (int32_t *)calloc((size_t)S + 1, 4);
    for (int c = 1; c < S; ++c) coff[parent[c] + 1]++;
    for (int c = 0; c < S; ++c) coff[c + 1] += coff[c];
    for (int c = 1; c < S; ++c) cl[coff[parent[c]] + cf[parent[c]]++] = c;
    /* --- per distinct separator: components + generalised Pruefer --- */
    int32_t *new_edges = (int32_t *)calloc((size_t)(S > 1 ? S - 1 : 1) * 2, 4);
    int ne = 0;
    uint8_t *done = (uint8_t *)calloc((size_t)S + 1, 1), *inT = (uint8_t *)calloc((size_t)S + 1, 1);
    int32_t *nodes = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
    int32_t *comp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
    int32_t *roots = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
    int32_t *cstart = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 2));
    int32_t *canon = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
    int32_t *vind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
    int32_t *wj = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
    int32_t *ebuf = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(S + 1));
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
    int32_t *rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
    for (int c = 1; c < S; ++c) {
        if (done[c]) continue;
        const uint64_t *X = sep + (size_t)c * W;
        /* T_X: cliques containing X, found by local traversal from c */
        int P = 0, sp = 0;
        stack[sp++] = c; inT[c] = 1;
        while (sp) {
            int d = stack[--sp]; nodes[P++] = d;
            int pd = parent[d];
            if (pd >= 0 && !inT[pd] && bs_subset(X, K + (size_t)pd * W, W)) { inT[pd] = 1; stack[sp++] = pd; }
            for (int e = coff[d]; e < coff[d + 1]; ++e) { int g = cl[e]; if (!inT[g] && bs_subset(X, K + (size_t)g * W, W)) { inT[g] = 1; stack[sp++] = g; } }
        }
        qsort(nodes, (size_t)P, sizeof(int32_t), cmp_i32);
        int q = 0;
        for (int a = 0; a < P; ++a) {
            int d = nodes[a], pd = parent[d];
            int top = (pd < 0) || !inT[pd];
            int isroot = top || bs_eq(sep + (size_t)d * W, X, W);
            if (!top && isroot) done[d] = 1;
            if (isroot) { comp[d] = d; roots[q++] = d; } else comp[d] = comp[pd];
        }
        /* canonical node list: components by ascending root id, members ascending */
        for (int i = 0; i < q; ++i) rank[roots[i]] = i;       /* roots[] is ascending already */
        for (int i = 0; i <= q; ++i) cstart[i] = 0;
        for (int a = 0; a < P; ++a) cstart[rank[comp[nodes[a]]] + 1]++;
        for (int i = 0; i < q; ++i) cstart[i + 1] += cstart[i];
        { int32_t *fill = (int32_t *)calloc((size_t)q, 4);
          for (int a = 0; a < P; ++a) { int i = rank[comp[nodes[a]]]; canon[cstart[i] + fill[i]++] = nodes[a]; }
          free(fill); }
        for (int t = 0; t < q - 2; ++t) vind[t] = (int32_t)draw_int(seed, chain_id, iter, TAG_PRUF_V, ((uint32_t)c << 16) | (uint32_t)t, (uint32_t)P);
        for (int i = 0; i < q; ++i) wj[i] = (int32_t)draw_int(seed, chain_id, iter, TAG_PRUF_W, ((uint32_t)c << 16) | (uint32_t)i, (uint32_t)(cstart[i + 1] - cstart[i]));
        int got = pdgo_tree_from_forest(P, q, cstart, vind, wj, ebuf);
        for (int e = 0; e < got; ++e) { new_edges[2 * ne] = canon[ebuf[2 * e]]; new_edges[2 * ne + 1] = canon[ebuf[2 * e + 1]]; ne++; }
        for (int a = 0; a < P; ++a) inT[nodes[a]] = 0;
    }
    int rc = (ne == (S > 0 ? S - 1 : 0)) ? 0 : -20;
    /* --- relabel + pad (junction_tree.py:36-64) --- */
    int32_t *keys = (int32_t *)malloc(sizeof(int32_t) * (size_t)p);
    for (int i = 0; i < p; ++i) keys[i] = i;
    for (int i = p - 1; i >= 1; --i) {
        int j = (int)draw_int(seed, chain_id, iter, TAG_PERM, (uint32_t)i, (uint32_t)(i + 1));
        int32_t t = keys[i]; keys[i] = keys[j]; keys[j] = t;
    }
    memset(ch->M, 0, 8 * (size_t)p * W);
    for (int c = 0; c < S; ++c) memcpy(ch->M + (size_t)keys[c] * W, K + (size_t)c * W, 8 * (size_t)W);
    int E = 0;
    for (int e = 0; e < ne; ++e) { ch->edges[2 * E] = keys[new_edges[2 * e]]; ch->edges[2 * E + 1] = keys[new_edges[2 * e + 1]]; E++; }
    int32_t *pool = (int32_t *)malloc(sizeof(int32_t) * (size_t)p);
    uint8_t *used = (uint8_t *)calloc((size_t)p, 1);
    int np_ = 0;
    for (int c = 0; c < S; ++c) { pool[np_++] = keys[c]; used[keys[c]] = 1; }
    for (int t = 0; t < p; ++t) {
        if (used[t]) continue;
        int j = (int)draw_int(seed, chain_id, iter, TAG_PAD, (uint32_t)t, (uint32_t)np_);
        ch->edges[2 * E] = t; ch->edges[2 * E + 1] = pool[j]; E++;
        pool[np_++] = t;
    }
    if (E != p - 1) rc = -21;
    build_csr(ch);
    build_node2t(ch);
    free(adj); free(K); free(sep); free(numbered); free(card); free(order); free(cliqueof); free(parent);
    free(coff); free(cl); free(cf); free(new_edges); free(done); free(inT); free(nodes); free(comp); free(roots);
    free(cstart); free(canon); free(vind); free(wj); free(ebuf); free(stack); free(rank); free(keys); free(pool); free(used);
    return rc;
}

/* ------------------------------------------------------------------ many chains ---------- */
typedef struct {
    const pdgo_model *m; int kind; int64_t n_chains; uint32_t chain0; int64_t n_samples, randomize; uint64_t seed;
    int64_t *n_prop, *n_acc; double *logl_out; pdgo_rec *recs; uint64_t *rec_bits; int64_t rec_cap;
    int32_t *final_edges; uint64_t *final_bits;
    int64_t next; pthread_mutex_t mu; int err;
} many_t;

static void *many_worker(void *arg)
{
    many_t *j = (many_t *)arg;
    const pdgo_model *m = j->m;
    int p = m->p, W = (p + 63) / 64;
    double *logl = (double *)malloc(sizeof(double) * (size_t)j->n_samples);
    uint64_t *single = (uint64_t *)calloc((size_t)p * W, 8);
    for (int v = 0; v < p; ++v) bs_set(single + (size_t)v * W, v);
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int64_t c = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (c >= j->n_chains) break;
        uint32_t cid = j->chain0 + (uint32_t)c;
        void *ch = pdgo_chain_create(m, 1);
        pdgo_chain_set_cliques(ch, single, p);
        int rc = pdgo_chain_randomize(ch, j->seed, cid, 0);
        logl[0] = pdgo_chain_init_logl(ch, j->kind);
        int64_t k = (j->kind == PDGO_KIND_SINGLE) ? 1 : 0, nrec = 0;
        pdgo_rec *recs = j->recs ? j->recs + (size_t)c * j->rec_cap : NULL;
        uint64_t *rb = j->rec_bits ? j->rec_bits + (size_t)c * j->rec_cap * 2 * W : NULL;
        int64_t it = 1;
        while (it < j->n_samples && rc == 0) {
            if (it % j->randomize == 0) rc = pdgo_chain_randomize(ch, j->seed, cid, (uint32_t)it);
            int64_t end = (it / j->randomize + 1) * j->randomize;
            if (end > j->n_samples) end = j->n_samples;
            if (rc == 0) rc = pdgo_chain_run(ch, j->kind, it, end, NULL, j->seed, cid, logl, &k, recs, rb, j->rec_cap, &nrec, NULL);
            it = end;
        }
        j->n_prop[c] = k; j->n_acc[c] = nrec;
        if (j->logl_out) memcpy(j->logl_out + (size_t)c * j->n_samples, logl, sizeof(double) * (size_t)j->n_samples);
        if (j->final_edges) pdgo_chain_get_state(ch, j->final_edges + (size_t)c * 2 * (p - 1), j->final_bits + (size_t)c * p * W);
        if (rc) j->err = rc;
        pdgo_chain_destroy(ch);
    }
    free(logl); free(single);
    return NULL;
}

int pdgo_run_many(const pdgo_model *m, int kind, int64_t n_chains, uint32_t chain0,
                  int64_t n_samples, int64_t randomize, uint64_t seed, int n_threads,
                  int64_t *n_prop, int64_t *n_acc, double *logl_out,
                  pdgo_rec *recs, uint64_t *rec_bits, int64_t rec_cap,
                  int32_t *final_edges, uint64_t *final_bits)
{
    many_t j; memset(&j, 0, sizeof(j));
    j.m = m; j.kind = kind; j.n_chains = n_chains; j.chain0 = chain0; j.n_samples = n_samples; j.randomize = randomize; j.seed = seed;
    j.n_prop = n_prop; j.n_acc = n_acc; j.logl_out = logl_out; j.recs = recs; j.rec_bits = rec_bits; j.rec_cap = rec_cap;
    j.final_edges = final_edges; j.final_bits = final_bits;
    pthread_mutex_init(&j.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, many_worker, &j);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&j.mu);
    return j.err;
}
