// The fused Metropolis-Hastings kernel.  ONE WARP owns one chain for iterations [it_begin,
// it_end): proposal enumeration on the latent tree (ballot/popc bitsets), clique scoring through
// the device-wide memo table, Philox accept/reject, state update and update-record append.
// Restates mh_parallel.py:225-304 (parallel moves) and :66-168 (single move).
//
// Lane mapping inside an iteration:
//   enumeration : lane <-> tree node (32 per pass)
//   scoring     : lane = 4 * move + set; the four complete sets of a move are
//                 set 0 = B = C - {v}, 1 = S0 = B & Cadj, 2 = B + v, 3 = S0 + v
//                 (connect: C = B, Cnew = B + v; disconnect: C = B + v, Cnew = B), so eight
//                 proposals are looked up / accepted per pass with no idle half-warps
//   memo misses : the whole warp factorises (GGM) or counts (log-linear) one set at a time
#pragma once
#include "pdg_kernels.cuh"
#include "pdg_loglin.cuh"
#include "pdg_randomize.cuh"

#ifdef PDG_PHASE_TIMERS
static __device__ unsigned long long g_kg[2][64];     // developer probe: cooperative GGM scores and cycles by set size; [.][0] = lane-private passes
#endif
#ifndef PDG_KT_RUN
#define PDG_KT_RUN 8   // sets up to this size are factorised lane-privately (<= PDG_KT), larger ones by the whole warp (p = 50: 12 -> 526, 8 -> 559, 5 -> 558 M proposals/s)
#endif

// developer-only latency attribution: nvcc -DPDG_PHASE_TIMERS accumulates clock64() deltas of the
// phases of an iteration into work[6] (phase id in the high bits is not needed: one run per phase
// selection, chosen with the PDG_PHASE environment variable through P.mvlog_cap's sign bit... no:
// simply eight counters in the per-chain `work` slots 8..15).
#ifdef PDG_PHASE_TIMERS
#define PT_DECL long long pt_last = clock64(); unsigned long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PT(i) { const long long pt_now = clock64(); pt[i] += (unsigned long long)(pt_now - pt_last); pt_last = pt_now; }
#define PT_FLUSH if (lane == 0) for (int q_ = 0; q_ < 8; ++q_) atomicAdd(&P.work[chain * 16 + 8 + q_], pt[q_]);
#else
#define PT_DECL
#define PT(i)
#define PT_FLUSH
#endif

// work accounting of one evaluated set (pdg_get_work): rare (memo misses only), fire-and-forget
template <int MODEL>
__device__ __forceinline__ void account_evaluated(const PdgParams &P, int slot, int k)
{
    const unsigned long long kk = (unsigned long long)k;
    unsigned long long *w = P.work + (size_t)slot * 16;
    if (MODEL == PDG_MODEL_GGM) {
        const unsigned long long dm = (P.dmode == 2 ? 2ull : 1ull);
        atomicAdd(&w[0], 8ull * kk * kk * dm);
        atomicAdd(&w[1], (2ull * kk * kk * kk + 3ull * kk * kk + 6ull * kk) * dm);
    } else {
        atomicAdd(&w[0], (unsigned long long)P.nrows * kk);
    }
    atomicAdd(&w[2], 1ull);
    atomicAdd(&w[3], kk);
}

// Scores of up to 32 complete sets, one per lane (`need` lanes carry a key each): ONE parallel round
// trip to the memo table, then
//   GGM, |X| <= PDG_KT : the lanes that missed factorise their own sets (rolled LDL^T over lane-private
//                        triangles in the warp's scratch) and publish them
//   otherwise          : the warp resolves the misses one after the other with the cooperative
//                        scorers (32-row LDL^T / contingency-table scan)
// req_sets / req_k: warp-uniform counters of the sets requested (hits + misses) and their sizes.
// Returns the ERR_* bits raised.
template <int MODEL, int WMAX, int NB = 4>
__device__ __forceinline__ int warp_scores(const PdgParams &P, const WarpScratch &ws, const LaneScratch &ls, int slot,
                                           int lane, const u64 (&X)[WMAX], bool need, double &score,
                                           u32 &req_sets, u32 &req_k)
{
    const int W = P.W;
    int err = 0;
    bool hit = true;
    score = 0.0;
    int kq = 0;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) kq += __popcll(X[w]);
    if (need) hit = memo_lookup<WMAX>(P.memo, X, W, score);
    req_sets += (u32)__popc(__ballot_sync(PDG_FULL, need));
    req_k += __reduce_add_sync(PDG_FULL, need ? (u32)kq : 0u);
    u32 miss = __ballot_sync(PDG_FULL, need && !hit);
    if (miss == 0u) return 0;
    if (MODEL == PDG_MODEL_GGM) {
        const bool smiss = need && !hit && kq <= PDG_KT_RUN;
        if (__any_sync(PDG_FULL, smiss)) {
            double s_new = 0.0;
#ifdef PDG_PHASE_TIMERS
            const long long lp_t0 = clock64();
#endif
            // one- and two-word bitsets: eight lanes per set in registers (p = 50: 551 -> 586 M proposals/s);
            // wider keys keep the rolled lane-private code, whose footprint does not grow with the word
            // count (p = 500: 61.8 with it, 57.5 with the register form).  Same arithmetic, same bits.
            if (WMAX <= 2 && PDG_KT_RUN <= 8) err |= ggm_group8_scores<WMAX>(P, lane, X, smiss ? kq : 0, s_new);
            else err |= ggm_lane_scores_rolled<WMAX>(P, ws.tri32, ls.idx, lane, X, smiss ? kq : 0, s_new);
#ifdef PDG_PHASE_TIMERS
            const u32 lp_m = __ballot_sync(PDG_FULL, smiss);
            if (lane == 0) { atomicAdd(&g_kg[0][0], (unsigned long long)__popc(lp_m)); atomicAdd(&g_kg[1][0], (unsigned long long)(clock64() - lp_t0)); }
#endif
            if (smiss) {
                score = s_new; hit = true;
                memo_insert<WMAX>(P.memo, X, W, s_new);
                account_evaluated<MODEL>(P, slot, kq);
            }
            __syncwarp();
            miss = __ballot_sync(PDG_FULL, need && !hit);
        }
    }
    while (miss) {
        int src = __ffs((int)miss) - 1;
        if (MODEL == PDG_MODEL_LOGLIN) {
            // largest missing set first: the smaller ones that are its subsets are then read off its
            // count table instead of being scanned
            const int kmax = __reduce_max_sync(PDG_FULL, (need && !hit) ? kq : 0);
            src = __ffs((int)__ballot_sync(PDG_FULL, need && !hit && kq == kmax)) - 1;
        }
        u64 Y[WMAX];
#pragma unroll
        for (int w = 0; w < WMAX; ++w) Y[w] = __shfl_sync(PDG_FULL, X[w], src);
        double sc = 0.0;
        int e;
        LlTable tab; tab.cells = 0u; tab.k = 0;
#ifdef PDG_PHASE_TIMERS
        const long long ggm_t0 = clock64();
#endif
        if (MODEL == PDG_MODEL_GGM) e = ggm_set_score<WMAX>(P, ws, Y, lane, sc);
        else e = loglin_set_score<WMAX, NB>(P, ws, slot, Y, lane, sc, tab);
#ifdef PDG_PHASE_TIMERS
        if (MODEL == PDG_MODEL_GGM && lane == 0) {
            int kk = 0;
#pragma unroll
            for (int w = 0; w < WMAX; ++w) kk += __popcll(Y[w]);
            kk = kk < 63 ? kk : 63;
            atomicAdd(&g_kg[0][kk], 1ull); atomicAdd(&g_kg[1][kk], (unsigned long long)(clock64() - ggm_t0));
        }
#endif
        err |= e;
        if (lane == 0) {
            if (!e) memo_insert<WMAX>(P.memo, Y, W, sc);
            int k = 0;
#pragma unroll
            for (int w = 0; w < WMAX; ++w) k += __popcll(Y[w]);
            account_evaluated<MODEL>(P, slot, k);
        }
        // every lane waiting for this very key takes the value
        bool same = need && !hit;
#pragma unroll
        for (int w = 0; w < WMAX; ++w) same = same && (X[w] == Y[w]);
        if (same) { score = sc; hit = true; }
        if (MODEL == PDG_MODEL_LOGLIN && tab.cells && !e) {
            for (;;) {
                bool sub = need && !hit;
#pragma unroll
                for (int w = 0; w < WMAX; ++w) sub = sub && ((X[w] & ~Y[w]) == 0ull);
                const u32 subs = __ballot_sync(PDG_FULL, sub);
                if (!subs) break;
                const int zs = __ffs((int)subs) - 1;
                u64 Zk[WMAX];
#pragma unroll
                for (int w = 0; w < WMAX; ++w) Zk[w] = __shfl_sync(PDG_FULL, X[w], zs);
                const double sz = loglin_subset_score<WMAX>(P, ws, tab, Zk, lane);
                if (lane == 0) { memo_insert<WMAX>(P.memo, Zk, W, sz); atomicAdd(&P.work[(size_t)slot * 16 + 7], 1ull); }
                bool sm = need && !hit;
#pragma unroll
                for (int w = 0; w < WMAX; ++w) sm = sm && (X[w] == Zk[w]);
                if (sm) { score = sz; hit = true; }
            }
        }
        miss = __ballot_sync(PDG_FULL, need && !hit);
    }
    return err;
}

// (re)loads the tree of a chain -- and, for small p, its whole state plus the adjacency bit-matrix of
// the tree -- into the warp's shared memory: at kernel start and after every randomisation
template <bool SMEM_STATE>
static __device__ __noinline__ void reload_chain_state(const PdgParams &P, unsigned char *smem, long long chain, int lane)
{
    const int p = P.p, W = P.W;
    const RunSmem L = run_smem_layout(p, W, SMEM_STATE, tri_bytes_for(P.ll_cap) + P.tri_extra);
    unsigned short *s_off = (unsigned short *)(smem + L.off_csr_off);
    unsigned short *s_adj = (unsigned short *)(smem + L.off_csr_adj);
    for (int i = lane; i <= p; i += 32) s_off[i] = P.csr_off[(size_t)chain * (p + 1) + i];
    for (int i = lane; i < 2 * (p - 1); i += 32) s_adj[i] = P.csr_adj[(size_t)chain * 2 * p + i];
    if (SMEM_STATE) {
        const u64 *Mg = P.M + (size_t)chain * p * W, *Ng = P.N + (size_t)chain * p * W;
        u64 *Mc = (u64 *)(smem + L.off_M), *Nc = (u64 *)(smem + L.off_N), *Tt = (u64 *)(smem + L.off_T);
        for (int i = lane; i < p * W; i += 32) { Mc[i] = Mg[i]; Nc[i] = Ng[i]; Tt[i] = 0ull; }
        __syncwarp();
        for (int t = lane; t < p; t += 32)
            for (int e = s_off[t]; e < s_off[t + 1]; ++e) { const int a = s_adj[e]; Tt[(size_t)t * W + (a >> 6)] |= 1ull << (a & 63); }
    }
    __syncwarp();
}

// LIGHT (log-linear models with few rows): one 512-row block per scan pass and a 128-register budget,
// so that two CTAs share an SM when a GPU runs more than 1184 chains.
// DIAG: the variant that can follow a replay log of the reference and keep the per-proposal
// diagnostics; the production kernels carry neither.
// SPEC: TWO warps per chain.  A chain is one long dependent instruction stream and a GPU that runs
// 1024 of them has fewer than two warps per scheduler, so the kernel is latency bound.  Iterations
// that accept nothing leave the state untouched (three out of four at the bench shapes), hence warp 1
// of a pair works on iteration it + 1 -- enumeration, memo probes, factorisations / table scans,
// accept decisions, all read-only -- while warp 0 runs iteration it as usual.  When warp 0 accepted
// nothing, warp 1's iteration is exactly what a sequential run would have computed and it is committed
// (records, bit flips, counters); otherwise it is discarded and iteration it + 1 is redone from the new
// state (scores it evaluated stay in the memo table).  Every draw is addressed by (chain, iteration), so
// the chains are bit-identical to the one-warp kernel's.
#ifndef PDG_NB
#define PDG_NB 2          // 512-row blocks per pass of the contingency-table scan (8 loads in flight per lane; 4: 45.3, 2: 47.4 M proposals/s on config 5)
#endif
#ifndef PDG_MH_THREADS
#define PDG_MH_THREADS 256
#endif

// pair mailbox (shared memory, 64 bytes at RunSmem::off_pair of the pair's first warp)
struct PairBox {
    double a_logp, b_logp;
    int a_nm, a_nacc, b_nm, b_nacc, b_ok, pad;
};

__device__ __forceinline__ void pair_bar(int id) { asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory"); }

template <int MODEL, int WMAX, bool SMEM_STATE, bool LIGHT, bool DIAG, bool SPEC>
__global__ void __launch_bounds__(SPEC ? 512 : PDG_MH_THREADS, (LIGHT && !SPEC) ? 2 : 1)
k_mh_run(const __grid_constant__ PdgParams P, int kind, long long it_begin, long long it_end,
         const __grid_constant__ PdgReplayDev R, int lockstep, RandScratch Z, long long randomize, int smem_per_warp)
{
    extern __shared__ __align__(16) unsigned char smem_all[];
    // log-linear chains with long tables: warps whose chain has ended serve the scans of the CTA's other
    // chains (pdg_loglin.cuh HelpBox) -- the run ends with its slowest chain, and that chain's scans are
    // the one thing the idle warps can take off its critical path
    constexpr bool HELP = MODEL == PDG_MODEL_LOGLIN && !LIGHT && !SPEC && !DIAG;
    __shared__ HelpBox s_help;
    if (HELP) {
        if (threadIdx.x == 0) { s_help.lock = 0; s_help.ticket = 0; s_help.n_chunks = 0; s_help.done = 0; s_help.finished = 0; }
        __syncthreads();
    }
    const int p = P.p, W = P.W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int role = SPEC ? (warp & 1) : 0;                 // 1 = the speculating warp of a pair
    const int cw = SPEC ? (warp >> 1) : warp;               // chain of this CTA
    const int wpc = SPEC ? (int)(blockDim.x >> 6) : (int)(blockDim.x >> 5);
    const long long chain = (long long)blockIdx.x * wpc + cw;
    if (chain >= P.n_chains) return;                       // (whole pairs leave; the barriers count live warps only)
    const int live_threads = 32 * (int)(((long long)(blockIdx.x + 1) * wpc <= P.n_chains) ? wpc : (P.n_chains - (long long)blockIdx.x * wpc));
    const RunSmem L = run_smem_layout(p, W, SMEM_STATE, tri_bytes_for(P.ll_cap) + P.tri_extra);
    // `smem_per_warp` bytes per CHAIN: one warp region (L.total, or the randomiser's layout if larger), or
    // the two regions of a pair back to back (the randomiser, run by the first warp while the second one
    // waits, may spread over both).  Chain state and pair mailbox sit in the first warp's region.
    unsigned char *smem_c = smem_all + (size_t)cw * smem_per_warp;
    unsigned char *smem = smem_c + (SPEC ? role * L.total : 0);
    u64 *s_sub = (u64 *)(smem + L.off_sub);
    u32 *s_sub32 = (u32 *)(smem + L.off_sub);
    u32 *s_leaf = (u32 *)(smem + L.off_leaf);
    u32 *s_bnd = (u32 *)(smem + L.off_bnd);
    short *s_partner = (short *)(smem + L.off_partner);
    short *s_mvU = (short *)(smem + L.off_mvU);
    short *s_mvUadj = (short *)(smem + L.off_mvUadj);
    unsigned short *s_off = (unsigned short *)(smem_c + L.off_csr_off);
    unsigned short *s_adj = (unsigned short *)(smem_c + L.off_csr_adj);
    short *s_tmpU = (short *)(smem + L.off_tmpU), *s_tmpA = (short *)(smem + L.off_tmpA);
    int *s_cnt = (int *)(smem + L.off_cnt);
    volatile PairBox *box = (volatile PairBox *)(smem_c + L.off_pair);
    const int bar_id = 1 + cw;
    WarpScratch ws;
    ws.help = HELP ? &s_help : nullptr;
    ws.ktri = P.ktri_run;
    ws.tri32 = (double *)(smem + L.off_tri);
    ws.idx = (short *)(smem + L.off_idx);
    ws.kbig = P.kbig;
    const int slot = SPEC ? (int)(2 * chain + role) : (int)chain;      // scratch slot of this warp (2 per chain exist)
    ws.big = P.big ? P.big + (size_t)slot * ((size_t)(P.kbig + 1) * (P.kbig + 2) / 2) : nullptr;
    LaneScratch ls;
    ls.unused = 0;
    ls.a = (double *)(smem + L.off_tri);
    ls.idx = (unsigned short *)(smem + L.off_lidx);

    u64 *Mg = P.M + (size_t)chain * p * W;
    u64 *Ng = P.N + (size_t)chain * p * W;
    u64 *Mc = SMEM_STATE ? (u64 *)(smem_c + L.off_M) : Mg;
    u64 *Nc = SMEM_STATE ? (u64 *)(smem_c + L.off_N) : Ng;
    u64 *Tt = (u64 *)(smem_c + L.off_T);                    // only with SMEM_STATE

    if (role == 0) reload_chain_state<SMEM_STATE>(P, smem, chain, lane);
    if (SPEC) pair_bar(bar_id);

    const u32 gchain = P.chain0 + (u32)chain;
    long long k = P.kcount[chain];
    long long nrec = P.nrec[chain];
    double logl_prev = P.logl[(size_t)chain * P.n_samples + it_begin - 1];
    double lbuf = 0.0;                  // logl of iteration (block start + lane): stored 32 at a time, coalesced
    u32 rb0 = 0u, rb1 = 0u, rb2 = 0u, rb3 = 0u;     // Philox draws of iteration (block start + lane)
    int rb_blk = -1;                    // which block of 32 iterations the draws belong to
    int errflag = 0;
    u32 req_sets = 0u, req_k = 0u;      // sets requested / their sizes, committed iterations (warp-uniform)
    const u32 lt = (1u << lane) - 1u;
    const int nw32 = (p + 31) >> 5;
    const bool replay_on = DIAG && R.on;
    // MCMC indices fit 32 bits (pdg_create); the next randomisation is tracked by a counter instead of a
    // 64-bit modulo per iteration
    const int it0 = (int)it_begin, it1 = (int)it_end, rz = (int)randomize;
    int next_rand = rz > 0 ? (int)(((it_begin + rz - 1) / rz) * rz) : 0x7fffffff;
    PT_DECL

    // stores the running log-probability of iteration x (called for every iteration, in order)
    auto emit_logl = [&](int x, double v) {
        const int s = (x - it0) & 31;
        if (lane == s) lbuf = v;
        if ((s == 31 || x + 1 == it1) && role == 0 && lane <= s) P.logl[(size_t)chain * P.n_samples + (x - s) + lane] = lbuf;
    };
    // accepted proposals in order: record append, bit flips (junction_tree.py:78-100 connect / disconnect)
    auto apply_accepts = [&](u32 am, bool acc, int itx, long long kbase, int j, int node, int mtype, int U, int Ua,
                             const u64 (&Cw)[WMAX], const u64 (&Aw)[WMAX]) {
        if (acc) {
            const long long slot_r = nrec + __popc(am & lt);
            if (slot_r < P.rec_cap) {
                const size_t ro = (size_t)chain * P.rec_cap + slot_r;
                pdg_update h;
                h.i = itx; h.k = (int)(kbase + j + 1); h.node = (unsigned short)node; h.U = (unsigned short)U;
                h.Uadj = (short)Ua; h.type = (unsigned char)mtype; h.pad = 0;
                P.rec[ro] = h;
#pragma unroll
                for (int w = 0; w < WMAX; ++w) if (w < W) { P.rec_bits[ro * 2 * W + w] = Cw[w]; P.rec_bits[ro * 2 * W + W + w] = Aw[w]; }
            } else if (P.rec_cap > 0) errflag |= ERR_OVERFLOW;      // rec_cap == 0: recording is off
            if (kbase + j + 1 > 2147483647LL) errflag |= ERR_OVERFLOW;     // pdg_update.k is 32 bits wide
        }
        nrec += __popc(am);
        __syncwarp();
        if (acc) {
            atomicXor(&Mc[(size_t)U * W + (node >> 6)], 1ull << (node & 63));
            atomicXor(&Nc[(size_t)node * W + (U >> 6)], 1ull << (U & 63));
        }
        if (!SMEM_STATE) __threadfence();           // the bit flips are performed before any lane reads the rows again
        __syncwarp();
    };

    // bit 16 of `lockstep`: the chains of a CTA enter every randomisation together (one CTA barrier per
    // interval), so that the randomiser's code streams through the instruction cache once per interval
    // instead of evicting the loop body whenever any one warp of the SM randomises
    const bool rsync = (lockstep & 0x10000) != 0;
    lockstep &= 0xffff;
    int it = it0;
    int round = 0, lock_epoch = -1;
    while (it < it1) {
        // optional re-alignment of the chains of a CTA every `lockstep` iterations (PDG_LOCKSTEP, a power of
        // two >= 2; on for the large-p kernels, whose per-iteration path does not fit the instruction cache:
        // in step the warps share its fetches).  A pair advances by one or two iterations per round, so the
        // barrier is tied to the epoch (it - it0) / lockstep, which every warp enters exactly once.
        if (lockstep) {
            const int ep = (it - it0) / lockstep;
            if (ep != lock_epoch) { lock_epoch = ep; asm volatile("bar.sync 15, %0;" :: "r"((SPEC ? 2 : 1) * live_threads) : "memory"); }
        }
        ++round;
        const long long nrec_round = nrec;
        // ---- junction-tree randomisation every `randomize` iterations (mh_parallel.py:226-228),
        // fused here so that a whole run is ONE launch and no chain ever waits for another.  The
        // randomiser works on the global-memory state and reuses this warp's shared memory.
        if (it == next_rand) {
            next_rand += rz;
            if (rsync) asm volatile("bar.sync 14, %0;" :: "r"((SPEC ? 2 : 1) * live_threads) : "memory");
            if (role == 0) {
                __syncwarp();
                if (SMEM_STATE) for (int i = lane; i < p * W; i += 32) { Mg[i] = Mc[i]; Ng[i] = Nc[i]; }
                __syncwarp();
                if (!SMEM_STATE) __threadfence();       // rows flipped with L2 atomics: nothing stale in L1 for the randomiser's loads
                randomize_chain(rand_args(P), Z, chain, (u32)it, smem, lane);
                __syncwarp();
                reload_chain_state<SMEM_STATE>(P, smem, chain, lane);
            }
            if (SPEC) pair_bar(bar_id);
        }
        const int my = it + role;          // the iteration this warp works on
        // warp 1 sits a round out when its iteration does not exist or starts with a randomisation
        const bool active = role == 0 || (my < it1 && my != next_rand);
        // results of this warp's iteration
        int nm = 0, node = 0, mtype = 0;
        double log_p = 0.0;
        bool spec_ok = true;
        u32 it_req_sets = 0u, it_req_k = 0u;
        int it_err = 0;
        // warp 1: the accept decisions of its (single) scoring pass, applied only on commit
        u32 pend_am = 0u; bool pend_acc = false; int pend_U = 0, pend_Ua = -1, pend_j = 0;
        u64 pend_C[WMAX], pend_A[WMAX];
#pragma unroll
        for (int w = 0; w < WMAX; ++w) { pend_C[w] = 0ull; pend_A[w] = 0ull; }
        if (active) {
        // ---- draws (mh_parallel.py:230-231): one Philox block per iteration, 32 iterations at a time
        // (lane l draws for iteration block start + l), handed out by shuffle
        const int sl = (my - it0) & 31;
        int aux = -1;
        u32 r2 = 0u, r3 = 0u;
        if (replay_on) {
            const size_t o = (size_t)chain * R.n_samples + my;
            node = R.it_node[o]; mtype = R.it_mtype[o]; aux = R.it_aux[o];
        } else {
            const int blk = (my - it0) >> 5;
            if (blk != rb_blk) {
                rb_blk = blk;
                u32 r[4];
                philox4x32_10(gchain, (u32)(my - sl + lane), TAG_ITER, 0u, P.seed_lo, P.seed_hi, r);
                rb0 = r[0]; rb1 = r[1]; rb2 = r[2]; rb3 = r[3];
            }
            node = (int)__umulhi(__shfl_sync(PDG_FULL, rb0, sl), (u32)p);
            mtype = (int)(__shfl_sync(PDG_FULL, rb1, sl) >> 31);
        }
        // (state in global memory is updated with atomics in L2: read it there, never through L1)
        if (lane < W) s_sub[lane] = SMEM_STATE ? Nc[(size_t)node * W + lane] : __ldcg(&Nc[(size_t)node * W + lane]);
        __syncwarp();
        int msub = 0;
#pragma unroll
        for (int w = 0; w < WMAX; ++w) if (w < W) msub += __popcll(s_sub[w]);
        PT(0)
        if (msub == 0) mtype = 0;            // :237-240 forced move types
        if (msub == p) mtype = 1;
        const bool special = (mtype == 0 && msub == 0) || (mtype == 1 && msub <= 2);
        // ---- one pass over the tree: leaves of the induced subtree / its outer boundary ------
        // (parallel_moves.py:193-208 and :253-272); proposals come out in ascending tree-node order
        int nleaf = 0, nbnd = 0;
        if (!SMEM_STATE) {
            // Large p (state in global memory): neighbourhood-driven enumeration.  List the few tree
            // nodes of the subtree, walk only THEIR adjacency lists (leaves: exactly one neighbour
            // inside; boundary: the neighbours outside), then rank-sort the boundary by tree node.
            // O(|subtree| * degree) instead of a pass over all p tree nodes.
            short *s_list = ws.idx;
            for (int i = lane; i < nw32; i += 32) { s_leaf[i] = 0u; s_bnd[i] = 0u; }
            if (lane == 0) *s_cnt = 0;
            {
                const int mine = (lane < W) ? __popcll(s_sub[lane]) : 0;
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(PDG_FULL, incl, o); if (lane >= o) incl += v; }
                int off = incl - mine;
                if (lane < W) {
                    u64 x = s_sub[lane];
                    while (x) { s_list[off++] = (short)(lane * 64 + __ffsll((long long)x) - 1); x &= x - 1; }
                }
            }
            __syncwarp();
            for (int i0 = 0; i0 < msub; i0 += 32) {
                const int i = i0 + lane;
                bool leaf = false;
                int s = -1, last = -1;
                if (i < msub) {
                    s = s_list[i];
                    int cnt = 0;
                    for (int e = s_off[s]; e < s_off[s + 1]; ++e) {
                        const int a = s_adj[e];
                        if (bit32(s_sub32, a)) { ++cnt; last = a; }
                        else { const int pos = atomicAdd(s_cnt, 1); s_tmpU[pos] = (short)a; s_tmpA[pos] = (short)s; }
                    }
                    leaf = cnt == 1;
                    if (leaf) { s_partner[s] = (short)last; atomicOr(&s_leaf[s >> 5], 1u << (s & 31)); }
                }
                const u32 bl = __ballot_sync(PDG_FULL, leaf);
                if (leaf && mtype == 1 && !special) {
                    const int pos = nleaf + __popc(bl & lt);
                    s_mvU[pos] = (short)s; s_mvUadj[pos] = (short)last;
                }
                nleaf += __popc(bl);
            }
            __syncwarp();
            nbnd = *s_cnt;
            for (int j = lane; j < nbnd; j += 32) {
                const int U = s_tmpU[j], A = s_tmpA[j];
                int rank = 0;
                for (int i = 0; i < nbnd; ++i) rank += (s_tmpU[i] < U) ? 1 : 0;
                s_partner[U] = (short)A;
                atomicOr(&s_bnd[U >> 5], 1u << (U & 31));
                if (mtype == 0 && !special) { s_mvU[rank] = (short)U; s_mvUadj[rank] = (short)A; }
            }
            nm = (mtype == 0) ? nbnd : nleaf;
        } else
        for (int base = 0; base < p; base += 32) {
            const int t = base + lane;
            bool leaf = false, bnd = false;
            int last = -1;
            if (t < p) {
                const int inT = bit32(s_sub32, t);
                int cnt = 0;
                // neighbours inside the subtree = adjacency row & subtree bitset
#pragma unroll
                for (int w = 0; w < WMAX; ++w) {
                    if (w < W) {
                        const u64 x = Tt[(size_t)t * W + w] & s_sub[w];
                        cnt += __popcll(x);
                        if (x && last < 0) last = w * 64 + __ffsll((long long)x) - 1;
                    }
                }
                leaf = inT && cnt == 1;
                bnd = !inT && cnt >= 1;
                s_partner[t] = (short)last;
            }
            const u32 bl = __ballot_sync(PDG_FULL, leaf), bb = __ballot_sync(PDG_FULL, bnd);
            if (lane == 0) { s_leaf[base >> 5] = bl; s_bnd[base >> 5] = bb; }
            nleaf += __popc(bl); nbnd += __popc(bb);
            const u32 fl = mtype == 0 ? bb : bl;
            if (!special && ((fl >> lane) & 1u)) {
                const int pos = nm + __popc(fl & lt);
                s_mvU[pos] = (short)t; s_mvUadj[pos] = (short)last;
            }
            nm += __popc(fl);
        }
        __syncwarp();
        PT(1)
        int nfwd = nm;
        double logq = 0.0;
        const u32 *set = (mtype == 0) ? s_bnd : s_leaf;
        if (special) {
            nm = 1;
            if (!replay_on) r2 = __shfl_sync(PDG_FULL, rb2, sl);
            if (lane == 0) {
                if (mtype == 0) {                         // pm:138-139 one random tree node
                    s_mvU[0] = (short)(replay_on ? aux : (int)__umulhi(r2, (u32)p));
                    s_mvUadj[0] = -1;
                } else {
                    int a = -1, b = -1;
                    for (int i = 0; i < nw32; ++i) {
                        u32 x = s_sub32[i];
                        while (x) { const int t = i * 32 + __ffs((int)x) - 1; x &= x - 1; if (a < 0) a = t; else if (b < 0) b = t; }
                    }
                    if (msub == 1) { s_mvU[0] = (short)a; s_mvUadj[0] = -1; }          // pm:186-187
                    else {                                  // pm:184-185 one of the two leaves
                        int pick;
                        if (replay_on) { pick = (aux == a) ? 0 : 1; if (aux != a && aux != b) it_err |= ERR_REPLAY; }
                        else pick = (int)(r2 >> 31);
                        s_mvU[0] = (short)(pick ? b : a);
                        s_mvUadj[0] = (short)(pick ? a : b);
                    }
                }
            }
            nfwd = (mtype == 1 && msub == 2) ? 2 : 1;
            if (kind == PDG_KIND_PARALLEL && mtype == 1 && msub == 2) logq = -0.6931471805599453;   // :277
        } else if (kind == PDG_KIND_PARALLEL) {
            if (replay_on) {
                // the reference's ORDER (python set iteration); the SET is checked against ours
                const long long o = R.it_off[(size_t)chain * (R.n_samples + 1) + my];
                const long long cnt = R.it_off[(size_t)chain * (R.n_samples + 1) + my + 1] - o;
                if (cnt != nm) { it_err |= ERR_REPLAY; nm = (int)(cnt < nm ? cnt : nm); }
                __syncwarp();
                for (int j = lane; j < nm; j += 32) {
                    const int U = R.mv_U[o + j], Ua = R.mv_Uadj[o + j];
                    if (U < 0 || U >= p || !bit32(set, U) || s_partner[U] != Ua) it_err |= ERR_REPLAY;
                    s_mvU[j] = (short)U; s_mvUadj[j] = (short)Ua;
                }
            }
        } else {
            // single move: moves[0] after np.random.shuffle (:87,:95 / :128,:138)
            int U, Ua;
            if (replay_on) {
                const long long o = R.it_off[(size_t)chain * (R.n_samples + 1) + my];
                U = R.mv_U[o]; Ua = R.mv_Uadj[o];
                if (U < 0 || U >= p || !bit32(set, U) || s_partner[U] != Ua) it_err |= ERR_REPLAY;
            } else {
                r3 = __shfl_sync(PDG_FULL, rb3, sl);
                const int pick = nm > 1 ? (int)__umulhi(r3, (u32)nm) : 0;
                U = s_mvU[pick]; Ua = s_mvUadj[pick];
            }
            __syncwarp();
            if (lane == 0) { s_mvU[0] = (short)U; s_mvUadj[0] = (short)Ua; }
            nm = 1;
        }
        __syncwarp();
        if (kind == PDG_KIND_SINGLE) {
            // Hastings term log_q2 - log_q1 = -log(num_moves) + log(num_reverse)  (:111-118, :154-160)
            const int U = s_mvU[0], Ua = s_mvUadj[0];
            int nrev;
            if (mtype == 0) {
                const int nl = msub >= 2 ? nleaf : 0;
                const int notin = (Ua >= 0 && msub >= 2 && bit32(s_leaf, Ua)) ? 0 : 1;   // pm:276-281
                nrev = nl + notin + (msub == 1 ? 1 : 0);
            } else {
                if (msub == 1) nrev = p;                                                   // :133-134
                else nrev = nbnd + 2 - ((int)s_off[U + 1] - (int)s_off[U]);                // pm:283-284
            }
            const double log_q2 = -log((double)nfwd), log_q1 = -log((double)nrev);
            logq = log_q2 - log_q1;
        }
        PT(2)
        // warp 1 keeps the decisions of ONE pass pending; an iteration with more than eight proposals is
        // left to warp 0
        if (SPEC && role == 1 && nm > 8) { spec_ok = false; nm = 0; }
        // ---- score + accept + apply, eight proposals per pass -----------------------------------
        for (int g0 = 0; g0 < nm; g0 += 8) {
            const int mi = lane >> 2, q = lane & 3, j = g0 + mi;
            const bool valid = j < nm;
            int U = 0, Ua = -1;
            if (valid) { U = s_mvU[j]; Ua = s_mvUadj[j]; }
            const bool adj = Ua >= 0;
            u64 Cw[WMAX], Aw[WMAX], X[WMAX];
            int kb = 0, ks = 0;
#pragma unroll
            for (int w = 0; w < WMAX; ++w) {
                Cw[w] = (valid && w < W) ? (SMEM_STATE ? Mc[(size_t)U * W + w] : __ldcg(&Mc[(size_t)U * W + w])) : 0ull;
                Aw[w] = (valid && adj && w < W) ? (SMEM_STATE ? Mc[(size_t)Ua * W + w] : __ldcg(&Mc[(size_t)Ua * W + w])) : 0ull;
                const u64 vb = ((node >> 6) == w) ? (1ull << (node & 63)) : 0ull;
                const u64 B = Cw[w] & ~vb, S0 = B & Aw[w];
                kb += __popcll(B); ks += __popcll(S0);
                X[w] = (q == 0) ? B : (q == 1) ? S0 : (q == 2) ? (B | vb) : (S0 | vb);
            }
            PT(3)
            double sc = 0.0;
            if (MODEL != PDG_MODEL_UNIFORM) {
                const int kq = (q == 0) ? kb : (q == 1) ? ks : (q == 2) ? kb + 1 : ks + 1;
                it_err |= warp_scores<MODEL, WMAX, LIGHT ? 1 : (SPEC ? 2 : PDG_NB)>(P, ws, ls, slot, lane, X, valid && kq > 0, sc, it_req_sets, it_req_k);
            }
            PT(4)
            const double sc1 = __shfl_down_sync(PDG_FULL, sc, 1), sc2 = __shfl_down_sync(PDG_FULL, sc, 2),
                         sc3 = __shfl_down_sync(PDG_FULL, sc, 3);
            bool acc = false;
            double delta = 0.0;
            if (valid && q == 0) {
                double log_p2, log_p1, log_g2, log_g1;
                if (mtype == 0) {
                    const int cC = kb, cS = ks, cCn = kb + 1, cSn = adj ? ks + 1 : 0;
                    log_p2 = sc2 - (adj ? sc3 : 0.0);
                    log_p1 = sc - sc1;
                    log_g2 = prior_partial(P, cCn, cSn, 0);
                    log_g1 = prior_partial(P, cC, cS, 1);
                    if (msub == 0) { log_p1 += sc3; log_g1 += prior_partial(P, 1, cS, 1); }   // :261-263
                } else {
                    const int cC = kb + 1, cS = adj ? ks + 1 : 0, cCn = kb, cSn = ks;
                    log_p2 = sc - sc1;
                    log_p1 = sc2 - (adj ? sc3 : 0.0);
                    log_g2 = prior_partial(P, cCn, cSn, 1);
                    log_g1 = prior_partial(P, cC, cS, 0);
                    if (msub == 1) { log_p2 += sc3; log_g2 += prior_partial(P, 1, cSn, 1); }  // :293-295
                }
                const double dl = log_p2 - log_p1, dp = log_g2 - log_g1;
                const double tot = (dl + dp) + logq;
                double u;
                if (replay_on) u = R.mv_u[R.it_off[(size_t)chain * (R.n_samples + 1) + my] + j];
                else u = draw_unif_unrolled(P.seed_lo, P.seed_hi, gchain, (u32)my, TAG_MOVE, (u32)U);
                // accept iff u <= min(1, exp(tot))  (:268,:300).  A single-precision exp brackets the
                // threshold to 1e-4 relative (its error stays below 2e-5 wherever it does not underflow,
                // and where it underflows exp(tot) < 2^-53 <= every non-zero u); only a draw inside the
                // bracket needs the double-precision exp, so the decision is the exact one.
                {
                    const double ef = (double)__expf((float)tot);
                    if (u <= ef * 0.9999) acc = true;
                    else if (u > ef * 1.0001) acc = false;
                    else { const double e = exp(tot); acc = u <= (e < 1.0 ? e : 1.0); }
                }
                delta = dl + dp;
                if (DIAG) {
                    const long long kk = k + j + 1;
                    if (P.mvlog_cap && kk - 1 < P.mvlog_cap) {
                        const size_t o = (size_t)chain * P.mvlog_cap + (kk - 1);
                        P.ml_dlik[o] = dl; P.ml_dprior[o] = dp; P.ml_logq[o] = logq; P.ml_u[o] = u;
                        P.ml_acc[o] = acc ? 1 : 0; P.ml_U[o] = U;
                    }
                }
            }
            PT(5)
            const u32 am = __ballot_sync(PDG_FULL, acc);
            if (am) {
                u32 rem = am;
                while (rem) {                                   // :269,:301 log_p += ... in proposal order
                    const int src = __ffs((int)rem) - 1;
                    rem &= rem - 1;
                    log_p += __shfl_sync(PDG_FULL, delta, src);
                }
                if (SPEC && role == 1) {
                    pend_am = am; pend_acc = acc; pend_U = U; pend_Ua = Ua; pend_j = j;
#pragma unroll
                    for (int w = 0; w < WMAX; ++w) { pend_C[w] = Cw[w]; pend_A[w] = Aw[w]; }
                } else {
                    apply_accepts(am, acc, my, k, j, node, mtype, U, Ua, Cw, Aw);
                }
            }
        }
        PT(6)
        }   // active
        if (SPEC) {
            // ---- the pair exchanges what its two iterations did (mailbox double-buffered by round)
            volatile PairBox *bx = box + (round & 1);
            const int my_nacc = role == 0 ? (int)(nrec - nrec_round) : __popc(pend_am);
            if (lane == 0) {
                if (role == 0) { bx->a_nm = nm; bx->a_nacc = my_nacc; bx->a_logp = log_p; }
                else {
                    bx->b_nm = nm; bx->b_nacc = my_nacc; bx->b_logp = log_p; bx->b_ok = (active && spec_ok) ? 1 : 0;
                }
            }
            pair_bar(bar_id);
            const int a_nm = bx->a_nm, a_nacc = bx->a_nacc, b_nm = bx->b_nm, b_nacc = bx->b_nacc;
            const double a_logp = bx->a_logp, b_logp = bx->b_logp;
            const bool commit = bx->b_ok != 0 && a_nacc == 0;
            // iteration `it` (warp 0 has applied it already)
            k += a_nm;
            logl_prev = logl_prev + a_logp;                     // :304 running sum
            emit_logl(it, logl_prev);
            if (role == 0) { errflag |= it_err; req_sets += it_req_sets; req_k += it_req_k; }
            else nrec += a_nacc;
            if (commit) {
                // iteration `it + 1`: warp 1 computed it on the very state a sequential run would have seen
                if (role == 1) {
                    if (pend_am) apply_accepts(pend_am, pend_acc, my, k, pend_j, node, mtype, pend_U, pend_Ua, pend_C, pend_A);
                    errflag |= it_err; req_sets += it_req_sets; req_k += it_req_k;
                } else nrec += b_nacc;
                k += b_nm;
                logl_prev = logl_prev + b_logp;
                emit_logl(it + 1, logl_prev);
                if (b_nacc) pair_bar(bar_id);                   // the flips are in place before anyone reads the state again
                it += 2;
            } else ++it;
        }
        if (!SPEC) {
            k += nm;
            logl_prev = logl_prev + log_p;                      // :304 running sum
            emit_logl(it, logl_prev);
            errflag |= it_err;
            req_sets += it_req_sets; req_k += it_req_k;
            ++it;
        }
    }
    __syncwarp();
    PT_FLUSH
    if (role == 0) {
        if (SMEM_STATE) for (int i = lane; i < p * W; i += 32) { Mg[i] = Mc[i]; Ng[i] = Nc[i]; }
        if (lane == 0) { P.kcount[chain] = k; P.nrec[chain] = nrec; }
    }
    if (lane == 0 && (req_sets | req_k)) {
        atomicAdd(&P.work[chain * 16 + 4], (unsigned long long)req_sets);
        atomicAdd(&P.work[chain * 16 + 5], (unsigned long long)req_k);
    }
    if (errflag) atomicOr(&P.err[chain], errflag);
    if (HELP) {
        __syncwarp();
        help_until_done<LIGHT ? 1 : PDG_NB>(&s_help, P.data, (unsigned int *)ws.tri32, P.ll_cap, live_threads >> 5, lane);
    }
}
