// Trajectory post-processing on the device, for every chain of a context, straight from the
// per-chain update records (no graph per iteration is ever materialised):
//   k_size_series : number of edges per MCMC index            (trajectory.py:202-211 size())
//   k_edge_presence + k_edge_ess : per-chain effective sample size of the edge indicators
//       (the second half of BASELINE.json's metric): for every edge whose marginal after burn-in
//       lies in (lo, hi), the indicator series over the list positions of trajectory.py:126-152,
//       its autocorrelation r(h) with the reference's estimator (auxiliary_functions.py:546-550:
//       sums divided by n and c0), ESS = n / (-1 + 2 sum_t (r(2t) + r(2t+1))) with the sum cut at the
//       first non-positive pair (Geyer), and the chain's figure = median over those edges.
// The indicator series are BIT-PACKED: toggle bits are set with XOR at the list position of every
// accepted move that adds / removes the edge, a prefix-XOR turns them into the presence series, and
// sum_t x_t x_{t+h} is a popcount of (series & series >> h) -- exact integer arithmetic, one lag per
// lane.
#pragma once
#include "pdg_kernels.cuh"

#define PDG_SIZE_SENT (-2147483647 - 1)

// neighbours gained / lost by record r of a chain (parallel_moves.py:287-309): C - Cadj - {node}
__device__ __forceinline__ u64 simplex_word(const u64 *rb, long long r, int W, int w, int node)
{
    u64 x = rb[(size_t)r * 2 * W + w] & ~rb[(size_t)r * 2 * W + W + w];
    if ((node >> 6) == w) x &= ~(1ull << (node & 63));
    return x;
}

__global__ void __launch_bounds__(32) k_size_series(PdgParams P, int *out_all)
{
    const int p = P.p, W = P.W, lane = threadIdx.x;
    const long long chain = blockIdx.x, ns = P.n_samples;
    int *out = out_all + (size_t)chain * ns;
    const u64 *G0 = P.G0 + (size_t)chain * p * W;
    int e0 = 0;
    for (int i = lane; i < p * W; i += 32) e0 += __popcll(G0[i]);
    e0 = warp_sum_i(e0) >> 1;
    for (long long j = lane; j < ns; j += 32) out[j] = PDG_SIZE_SENT;
    __syncwarp();
    long long n = P.nrec[chain];
    if (n > P.rec_cap) n = P.rec_cap;
    const pdg_update *rec = P.rec + (size_t)chain * P.rec_cap;
    const u64 *rb = P.rec_bits + (size_t)chain * P.rec_cap * 2 * W;
    int size = e0;
    for (long long r0 = 0; r0 < n; r0 += 32) {
        const long long r = r0 + lane;
        int delta = 0, pos = -1;
        if (r < n) {
            const pdg_update h = rec[r];
            int c = 0;
            for (int w = 0; w < W; ++w) c += __popcll(simplex_word(rb, r, W, w, h.node));
            delta = h.type == 0 ? c : -c;
            pos = h.i - 1;
        }
        int incl = delta;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(PDG_FULL, incl, o); if (lane >= o) incl += v; }
        int nextpos = __shfl_down_sync(PDG_FULL, pos, 1);
        if (lane == 31) nextpos = (r + 1 < n) ? rec[r + 1].i - 1 : -1;
        // the last record of an MCMC index carries the size of that list position
        if (r < n && pos != nextpos && pos >= 0 && pos < ns) out[pos] = size + incl;
        size += __shfl_sync(PDG_FULL, incl, 31);
    }
    __syncwarp();
    int carry = e0;
    for (long long j0 = 0; j0 < ns; j0 += 32) {
        const long long j = j0 + lane;
        const int v = j < ns ? out[j] : PDG_SIZE_SENT;
        const u32 has = __ballot_sync(PDG_FULL, v != PDG_SIZE_SENT);
        const u32 below = has & (lane == 31 ? 0xffffffffu : ((2u << lane) - 1u));
        const int src = below ? 31 - __clz((int)below) : 0;
        const int got = __shfl_sync(PDG_FULL, v, src);
        const int val = below ? got : carry;
        if (j < ns) out[j] = val;
        carry = __shfl_sync(PDG_FULL, val, 31);
    }
}

// Per-chain presence totals of every edge over the list positions [burnin, n_samples) (the same
// replay as k_edge_tallies, kept per chain), then the edges with lo < marginal < hi get a slot:
// slot_all[chain][a * p + b] (a < b) = slot or -1, nq[chain] = number of slots.
__global__ void __launch_bounds__(32) k_edge_presence(PdgParams P, int *since_all, int *slot_all, long long burnin,
                                                      double lo, double hi, int *nq)
{
    const int p = P.p, W = P.W, lane = threadIdx.x;
    const long long chain = blockIdx.x;
    int *since = since_all + (size_t)chain * p * p;
    int *tot = slot_all + (size_t)chain * p * p;
    for (int i = lane; i < p * p; i += 32) { since[i] = -1; tot[i] = 0; }
    __syncwarp();
    const u64 *G0 = P.G0 + (size_t)chain * p * W;
    for (int i = lane; i < p * W; i += 32) {
        const int v = i / W, w = i - v * W;
        u64 x = G0[i];
        while (x) {
            const int b = __ffsll((long long)x) - 1;
            x &= x - 1;
            const int y = w * 64 + b;
            if (v < y) since[v * p + y] = 0;
        }
    }
    __syncwarp();
    long long n = P.nrec[chain];
    if (n > P.rec_cap) n = P.rec_cap;
    const pdg_update *rec = P.rec + (size_t)chain * P.rec_cap;
    const u64 *rb = P.rec_bits + (size_t)chain * P.rec_cap * 2 * W;
    for (long long r = 0; r < n; ++r) {
        const pdg_update h = rec[r];
        const int node = h.node, pos = h.i - 1;
        for (int w = 0; w < W; ++w) {
            const u64 x = simplex_word(rb, r, W, w, node);
            const int cnt = __popcll(x);
            for (int j = lane; j < cnt; j += 32) {
                const int y = w * 64 + select64(x, j);
                const int a = node < y ? node : y, b = node < y ? y : node;
                int *s = &since[a * p + b];
                if (h.type == 0) { if (*s < 0) *s = pos; }
                else if (*s >= 0) {
                    const long long from = *s > burnin ? *s : burnin;
                    const long long d = (long long)pos - from;
                    if (d > 0) tot[a * p + b] += (int)d;
                    *s = -1;
                }
            }
        }
        __syncwarp();
    }
    const double len = (double)(P.n_samples - burnin);
    int count = 0;
    for (int i0 = 0; i0 < p * p; i0 += 32) {
        const int i = i0 + lane;
        bool q = false;
        if (i < p * p) {
            int t = tot[i];
            const int s = since[i];
            if (s >= 0) {
                const long long from = s > burnin ? s : burnin;
                const long long d = P.n_samples - from;
                if (d > 0) t += (int)d;
            }
            const double m = (double)t / len;
            q = (i / p) < (i % p) && m > lo && m < hi;
        }
        const u32 qm = __ballot_sync(PDG_FULL, q);
        if (i < p * p) tot[i] = q ? count + __popc(qm & ((1u << lane) - 1u)) : -1;
        count += __popc(qm);
    }
    if (lane == 0) nq[chain] = count;
}

// in-word inclusive prefix parity
__device__ __forceinline__ u64 prefix_xor64(u64 x)
{
    x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16; x ^= x << 32;
    return x;
}

// One CTA (ESS_WARPS warps) per chain at a time.  bitmaps: [gridDim.x][qcap][words] u64 scratch;
// vals: [gridDim.x][qcap] f64 scratch; dynamic shared memory: ESS_WARPS * (words + 2) u64 + ESS_WARPS *
// (words + 1) u32.
#define ESS_WARPS 4
__global__ void __launch_bounds__(32 * ESS_WARPS) k_edge_ess(PdgParams P, const int *slot_all, const int *nq_all,
                                                             long long burnin, int qcap, int words, u64 *bitmaps,
                                                             double *vals_all, double *ess_out)
{
    extern __shared__ __align__(16) unsigned char ess_smem[];
    const int p = P.p, W = P.W, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nthr = blockDim.x;
    const long long len = P.n_samples - burnin;                 // series length n
    u64 *bm = bitmaps + (size_t)blockIdx.x * qcap * words;
    double *vals = vals_all + (size_t)blockIdx.x * qcap;
    u64 *s_ser = (u64 *)ess_smem + (size_t)warp * (words + 2);
    u32 *s_cum = (u32 *)(ess_smem + (size_t)ESS_WARPS * (words + 2) * 8) + (size_t)warp * (words + 1);
    for (long long chain = blockIdx.x; chain < P.n_chains; chain += gridDim.x) {
        const int nq = nq_all[chain] < qcap ? nq_all[chain] : qcap;
        const int *slot = slot_all + (size_t)chain * p * p;
        if (nq == 0) { if (threadIdx.x == 0) ess_out[chain] = __longlong_as_double(0x7ff8000000000000ll); continue; }
        for (size_t i = threadIdx.x; i < (size_t)nq * words; i += nthr) bm[i] = 0ull;
        __syncthreads();
        // ---- toggle bits: edges of the initial graph at position 0, then one XOR per edge of every record
        const u64 *G0 = P.G0 + (size_t)chain * p * W;
        for (int i = threadIdx.x; i < p * W; i += nthr) {
            const int v = i / W, w = i - v * W;
            u64 x = G0[i];
            while (x) {
                const int b = __ffsll((long long)x) - 1;
                x &= x - 1;
                const int y = w * 64 + b;
                if (v < y) { const int s = slot[v * p + y]; if (s >= 0 && s < nq) atomicXor(&bm[(size_t)s * words], 1ull); }
            }
        }
        long long n = P.nrec[chain];
        if (n > P.rec_cap) n = P.rec_cap;
        const pdg_update *rec = P.rec + (size_t)chain * P.rec_cap;
        const u64 *rb = P.rec_bits + (size_t)chain * P.rec_cap * 2 * W;
        for (long long r = threadIdx.x; r < n; r += nthr) {
            const pdg_update h = rec[r];
            const int node = h.node;
            long long pos = (long long)h.i - 1 - burnin;
            if (pos < 0) pos = 0;
            if (pos >= len) continue;
            for (int w = 0; w < W; ++w) {
                u64 x = simplex_word(rb, r, W, w, node);
                while (x) {
                    const int y = w * 64 + __ffsll((long long)x) - 1;
                    x &= x - 1;
                    const int a = node < y ? node : y, b = node < y ? y : node;
                    const int s = slot[a * p + b];
                    if (s >= 0 && s < nq) atomicXor(&bm[(size_t)s * words + (pos >> 6)], 1ull << (pos & 63));
                }
            }
        }
        __threadfence();
        __syncthreads();
        // ---- one warp per edge: presence series, autocorrelation, Geyer cut
        for (int s = warp; s < nq; s += ESS_WARPS) {
            const u64 *tg = bm + (size_t)s * words;
            u32 carry = 0u, run = 0u;
            for (int w0 = 0; w0 < words; w0 += 32) {
                const int w = w0 + lane;
                u64 x = w < words ? prefix_xor64(__ldcg(&tg[w])) : 0ull;
                u32 par = (u32)(x >> 63);                      // parity of this word
                u32 incl = par;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(PDG_FULL, incl, o); if (lane >= o) incl ^= v; }
                const u32 before = (incl ^ par) ^ carry;
                if (before) x = ~x;
                const long long left = len - (long long)w * 64;   // bits of this word inside the series
                if (left <= 0) x = 0ull;
                else if (left < 64) x &= (1ull << left) - 1ull;
                const int pc = __popcll(x);
                int pi = pc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(PDG_FULL, pi, o); if (lane >= o) pi += v; }
                if (w < words) { s_ser[w] = x; s_cum[w] = run + (u32)(pi - pc); }
                run += (u32)__shfl_sync(PDG_FULL, pi, 31);
                carry ^= __shfl_sync(PDG_FULL, incl, 31);
            }
            if (lane == 0) { s_ser[words] = 0ull; s_ser[words + 1] = 0ull; s_cum[words] = run; }
            __syncwarp();
            const double nn = (double)len, n1 = (double)run, m = n1 / nn, c0 = m - m * m;   // sum (x - m)^2 / n of a 0/1 series
            double tau = -1.0;
            bool stop = !(c0 > 0.0);
            for (long long h0 = 0; h0 < len && !stop; h0 += 32) {
                const long long h = h0 + lane;
                double rh = 0.0;
                if (h < len) {
                    const int q = (int)(h >> 6), sh = (int)(h & 63);
                    int s11 = 0;
                    if (sh == 0) {
                        for (int w = 0; w + q < words; ++w) s11 += __popcll(s_ser[w] & s_ser[w + q]);
                    } else {
#pragma unroll 4
                        for (int w = 0; w + q < words; ++w) {
                            const u64 v = (s_ser[w + q] >> sh) | (s_ser[w + q + 1] << (64 - sh));
                            s11 += __popcll(s_ser[w] & v);
                        }
                    }
                    // ones among the first (n - h) positions and among the last (n - h)
                    const long long cutA = len - h;
                    const int qa = (int)(cutA >> 6), ra = (int)(cutA & 63);
                    const u32 A = s_cum[qa] + (ra ? (u32)__popcll(s_ser[qa] & ((1ull << ra) - 1ull)) : 0u);
                    const u32 firsth = s_cum[q] + (sh ? (u32)__popcll(s_ser[q] & ((1ull << sh) - 1ull)) : 0u);
                    const u32 B = run - firsth;
                    rh = ((double)s11 - m * ((double)A + (double)B) + (double)cutA * m * m) / nn / c0;
                }
                // pairs (2t, 2t + 1): even lanes hold the pair sum; the pair exists while 2t + 1 < n
                const double g = rh + __shfl_down_sync(PDG_FULL, rh, 1);
                const bool pair = (lane & 1) == 0;
                const bool bad = pair && (h + 1 >= len || g <= 0.0);
                const u32 badm = __ballot_sync(PDG_FULL, bad);
                const int first = badm ? __ffs((int)badm) - 1 : 32;
                double part = (pair && lane < first) ? g : 0.0;
                // sequential order of the reference loop: add the pairs lane 0, 2, 4, ...
                for (int l = 0; l < first; l += 2) tau += 2.0 * __shfl_sync(PDG_FULL, part, l);
                if (badm) stop = true;
            }
            if (lane == 0) vals[s] = nn / (tau > 1e-12 ? tau : 1e-12);
            __syncwarp();
        }
        __syncthreads();
        // ---- median of the nq values (numpy.median: mean of the two middle order statistics)
        {
            const int k1 = (nq - 1) >> 1, k2 = nq >> 1;
            __shared__ double s_med[2];
            for (int i = threadIdx.x; i < nq; i += nthr) {
                const double v = vals[i];
                int less = 0, eq = 0;
                for (int j = 0; j < nq; ++j) { const double u = vals[j]; less += (u < v) ? 1 : 0; eq += (u == v) ? 1 : 0; }
                if (less <= k1 && k1 < less + eq) s_med[0] = v;
                if (less <= k2 && k2 < less + eq) s_med[1] = v;
            }
            __syncthreads();
            if (threadIdx.x == 0) ess_out[chain] = 0.5 * (s_med[0] + s_med[1]);
        }
        __syncthreads();
    }
}
