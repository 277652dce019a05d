// Edge-inclusion tallies on the device (trajectory.py:126-152 + empirical_graph_distribution.py
// :35-41): the graph at list position j is the state after every accepted move with MCMC index
// <= j + 1, so an edge switched on at index i counts from position i - 1.  One warp replays the
// packed update records of one chain against a per-chain "present since" table and adds the
// time-weighted presence of every edge to the shared p x p tally.
#pragma once
#include "pdg_kernels.cuh"

__global__ void __launch_bounds__(32) k_edge_tallies(PdgParams P, int *since_all, long long burnin,
                                                     unsigned long long *tally)
{
    const int p = P.p, W = P.W, lane = threadIdx.x;
    const long long chain = blockIdx.x;
    int *since = since_all + (size_t)chain * p * p;
    for (int i = lane; i < p * p; i += 32) since[i] = -1;
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
            // simplex = C - Cadj - {node}: the neighbours gained / lost (parallel_moves.py:287-309)
            u64 x = rb[(size_t)r * 2 * W + w] & ~rb[(size_t)r * 2 * W + W + w];
            if ((node >> 6) == w) x &= ~(1ull << (node & 63));
            const int cnt = __popcll(x);
            for (int j = lane; j < cnt; j += 32) {
                const int y = w * 64 + select64(x, j);
                const int a = node < y ? node : y, b = node < y ? y : node;
                int *s = &since[a * p + b];
                if (h.type == 0) { if (*s < 0) *s = pos; }
                else if (*s >= 0) {
                    const long long from = *s > burnin ? *s : burnin;
                    const long long d = (long long)pos - from;
                    if (d > 0) atomicAdd(&tally[a * p + b], (unsigned long long)d);
                    *s = -1;
                }
            }
        }
        __syncwarp();
    }
    for (int i = lane; i < p * p; i += 32) {
        const int s = since[i];
        if (s >= 0) {
            const long long from = s > burnin ? s : burnin;
            const long long d = P.n_samples - from;
            if (d > 0) atomicAdd(&tally[i], (unsigned long long)d);
        }
    }
}

// graph adjacency of every chain from its cliques (junction_tree.py:201-219), one warp per chain
__global__ void __launch_bounds__(32) k_snapshot_graph(PdgParams P)
{
    const int p = P.p, W = P.W, lane = threadIdx.x;
    const long long chain = blockIdx.x;
    const u64 *Mc = P.M + (size_t)chain * p * W, *Nc = P.N + (size_t)chain * p * W;
    u64 *G = P.G0 + (size_t)chain * p * W;
    for (int i = lane; i < p * W; i += 32) {
        const int v = i / W, w = i - v * W;
        u64 acc = 0ull;
        for (int tw = 0; tw < W; ++tw) {
            u64 x = Nc[(size_t)v * W + tw];
            while (x) {
                const int b = __ffsll((long long)x) - 1;
                x &= x - 1;
                acc |= Mc[(size_t)(tw * 64 + b) * W + w];
            }
        }
        if ((v >> 6) == w) acc &= ~(1ull << (v & 63));
        G[i] = acc;
    }
}
