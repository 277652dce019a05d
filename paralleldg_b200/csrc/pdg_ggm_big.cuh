// Hyper-inverse-Wishart score of LARGE complete sets (more than 32 vertices): one CTA per set.
// Restates gaussian_graphical_model.py:26-39 + wishart.py:17-37 for the batch scorer
// (sd.log_likelihood_partial); the sampler itself meets such cliques too rarely to need it.
//
//   gather   : the |X| rows of A = D + X^T X named by the set are staged whole into shared memory
//              with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx::bytes, eight rows in
//              flight per buffer set, two buffer sets), and the |X| columns are picked out of every
//              staged row into a packed lower triangle that stays in shared memory
//   factorise: blocked right-looking LDL^T, panels of 8 columns.  The panel is factorised column by
//              column; the trailing update subtracts the eight rank-1 terms of the panel from every
//              remaining element IN COLUMN ORDER, which is the order of the unblocked elimination
//              (pdg_device.cuh ldl_logdets) -- so the pivots, and the score, carry the same bits
//              whichever path computed them.  Trailing update: a warp owns four rows at a time and its
//              lanes sweep the columns; the panel's columns are first transposed into a compact buffer
//              so that those sweeps read consecutive doubles.
//   logdet   : sum of log pivots, reduced in the order of the warp path.
#pragma once
#include "pdg_kernels.cuh"

#define PDG_BIG_NB 8            // panel width
#define PDG_BIG_STAGES 4        // rows in flight per buffer set
#define PDG_BIG_MIN 33          // smallest set the CTA kernel takes
#define PDG_BIG_LD 256          // row length of the transposed panel buffer = largest set (PdgParams.kbig <= 256)

struct BigSmem { int off_bar, off_idx, off_red, off_cp, off_stage, off_tri, row_bytes, total, kmax; };

// layout for a given budget of dynamic shared memory: everything but the packed triangle is fixed, the
// triangle takes the rest (kmax = largest set that fits)
__host__ __device__ inline BigSmem big_smem_layout(int p, int budget_bytes)
{
    BigSmem s;
    int o = 0;
    s.off_bar = o;   o += 2 * PDG_BIG_STAGES * 8;
    s.off_idx = o;   o += PDG_BIG_LD * 2;
    s.off_red = o;   o += 8 * 8;
    s.off_cp = o;    o += PDG_BIG_NB * PDG_BIG_LD * 8;             // transposed panel columns [NB][PDG_BIG_LD]
    s.row_bytes = ((p * 8 + 15) / 16) * 16 + 16;                   // one row of A, from a 16-byte-aligned address
    s.off_stage = o; o += 2 * PDG_BIG_STAGES * s.row_bytes;
    s.off_tri = o;
    const long long room = ((long long)budget_bytes - o) / 8;     // doubles left for the packed triangle
    int k = 0;
    while ((long long)(k + 1) * (k + 2) / 2 <= room && k < PDG_BIG_LD) ++k;
    s.kmax = k;
    s.total = budget_bytes;
    return s;
}
// bytes of dynamic shared memory for sets of at most k vertices
__host__ inline int big_smem_bytes_for(int p, int k)
{
    const BigSmem s = big_smem_layout(p, 0);
    return s.off_tri + (k * (k + 1) / 2) * 8 + 16;
}

__device__ __forceinline__ void big_mbar_init(unsigned int bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void big_mbar_expect_tx(unsigned int bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void big_mbar_wait(unsigned int bar, unsigned int parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared (1-D, size and both addresses multiples of 16 bytes)
__device__ __forceinline__ void big_bulk_load(unsigned int dst, const void *src, unsigned int bytes, unsigned int bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// one factorisation of mat[idx, idx]; returns sum of log pivots (same bits as ldl_logdets) on every thread
__device__ __forceinline__ bool big_logdet(const double *__restrict__ mat, int p, int k, unsigned char *smem, const BigSmem &L,
                                           unsigned int &phase_bits, double &logdet)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned short *idx = (const unsigned short *)(smem + L.off_idx);
    double *a = (double *)(smem + L.off_tri);
    double *cp = (double *)(smem + L.off_cp);
    double *red = (double *)(smem + L.off_red);
    const unsigned int bar0 = (unsigned int)__cvta_generic_to_shared(smem + L.off_bar);
    const unsigned int stage0 = (unsigned int)__cvta_generic_to_shared(smem + L.off_stage);
    const unsigned char *stage = smem + L.off_stage;
    // ---- gather through staged rows ------------------------------------------------------------
    const int ngroups = (k + PDG_BIG_STAGES - 1) / PDG_BIG_STAGES;
    auto issue = [&](int g) {
        const int set = g & 1;
        for (int s = 0; s < PDG_BIG_STAGES; ++s) {
            const int i = g * PDG_BIG_STAGES + s;
            if (i >= k) break;
            const unsigned long long addr = (unsigned long long)(mat + (size_t)idx[i] * p);
            const unsigned long long lo = addr & ~15ull, hi = (addr + (unsigned long long)p * 8ull + 15ull) & ~15ull;
            const unsigned int bytes = (unsigned int)(hi - lo);
            const unsigned int bar = bar0 + 8u * (unsigned int)(set * PDG_BIG_STAGES + s);
            big_mbar_expect_tx(bar, bytes);
            big_bulk_load(stage0 + (unsigned int)((set * PDG_BIG_STAGES + s) * L.row_bytes), (const void *)lo, bytes, bar);
        }
    };
    if (tid == 0) { issue(0); if (ngroups > 1) issue(1); }
    for (int g = 0; g < ngroups; ++g) {
        const int set = g & 1;
        for (int s = 0; s < PDG_BIG_STAGES; ++s) {
            const int i = g * PDG_BIG_STAGES + s;
            if (i >= k) break;
            const int b = set * PDG_BIG_STAGES + s;
            big_mbar_wait(bar0 + 8u * (unsigned int)b, (phase_bits >> b) & 1u);
            const unsigned long long addr = (unsigned long long)(mat + (size_t)idx[i] * p);
            const double *row = (const double *)(stage + (size_t)b * L.row_bytes + (addr & 15ull));
            double *dst = a + tri(i, 0);
            for (int l = tid; l <= i; l += 256) dst[l] = row[idx[l]];
        }
        // flip the parities of the barriers this group used
        for (int s = 0; s < PDG_BIG_STAGES; ++s) if (g * PDG_BIG_STAGES + s < k) phase_bits ^= 1u << (set * PDG_BIG_STAGES + s);
        __syncthreads();                                   // the buffer set is free again
        if (tid == 0 && g + 2 < ngroups) issue(g + 2);
    }
    // ---- blocked LDL^T -----------------------------------------------------------------------------
    for (int J = 0; J < k; J += PDG_BIG_NB) {
        const int nb = (k - J < PDG_BIG_NB) ? k - J : PDG_BIG_NB, Je = J + nb;
        for (int j = J; j < Je; ++j) {
            const double inv = 1.0 / a[tri(j, j)];
            for (int i = j + 1 + tid; i < k; i += 256) {
                double *row = a + tri(i, 0);
                const double f = row[j] * inv;
                const int le = (i < Je - 1) ? i : Je - 1;
                int cj = tri(j + 1, j);
                for (int l = j + 1; l <= le; ++l) { row[l] -= f * a[cj]; cj += l + 1; }
            }
            __syncthreads();
        }
        if (Je >= k) break;
        // panel columns, unnormalised, transposed: cp[jj][l] = a[l][J + jj] for the trailing rows l
        for (int e = tid; e < nb * (k - Je); e += 256) {
            const int jj = e / (k - Je), l = Je + e - jj * (k - Je);
            cp[jj * PDG_BIG_LD + l] = a[tri(l, J + jj)];
        }
        double inv[PDG_BIG_NB];
#pragma unroll
        for (int jj = 0; jj < PDG_BIG_NB; ++jj) inv[jj] = (jj < nb) ? 1.0 / a[tri(J + jj, J + jj)] : 0.0;
        __syncthreads();
        // trailing update: warp w takes rows Je + w, Je + w + 8, ... four at a time; lanes sweep the columns
        for (int i0 = Je + warp; i0 < k; i0 += 32) {
            double f[4][PDG_BIG_NB];
            int ri[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                ri[r] = i0 + 8 * r;
#pragma unroll
                for (int jj = 0; jj < PDG_BIG_NB; ++jj) f[r][jj] = (ri[r] < k && jj < nb) ? cp[jj * PDG_BIG_LD + ri[r]] * inv[jj] : 0.0;
            }
            const int imax = (i0 + 24 < k) ? i0 + 24 : ((i0 + 16 < k) ? i0 + 16 : ((i0 + 8 < k) ? i0 + 8 : i0));
            for (int l = Je + lane; l <= imax; l += 32) {
                double c[PDG_BIG_NB];
#pragma unroll
                for (int jj = 0; jj < PDG_BIG_NB; ++jj) c[jj] = cp[jj * PDG_BIG_LD + l];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (ri[r] < k && l <= ri[r]) {
                        double v = a[tri(ri[r], l)];
#pragma unroll
                        for (int jj = 0; jj < PDG_BIG_NB; ++jj) if (jj < nb) v -= f[r][jj] * c[jj];
                        a[tri(ri[r], l)] = v;
                    }
                }
            }
        }
        __syncthreads();
    }
    // ---- sum of log pivots in the warp path's order (lane j % 32, ascending; xor butterfly) ---------------
    bool ok = true;
    if (warp == 0) {
        double acc = 0.0;
        for (int j = lane; j < k; j += 32) {
            const double d = a[tri(j, j)];
            if (!(d > 0.0)) ok = false;
            acc += log(d);
        }
        acc = warp_sum(acc);
        ok = __all_sync(PDG_FULL, ok);
        if (lane == 0) { red[0] = acc; red[1] = ok ? 1.0 : 0.0; }
    }
    __syncthreads();
    logdet = red[0];
    ok = red[1] != 0.0;
    __syncthreads();
    return ok;
}

// list[b] = index of a set with more than 32 vertices; one CTA each
template <int WMAX>
__global__ void __launch_bounds__(256) k_score_big(PdgParams P, const u64 *bits, const int *list, int n_list, double *out,
                                                   int *err, int smem_bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BigSmem L = big_smem_layout(P.p, smem_bytes);
    unsigned short *idx = (unsigned short *)(smem + L.off_idx);
    double *red = (double *)(smem + L.off_red);
    const unsigned int bar0 = (unsigned int)__cvta_generic_to_shared(smem + L.off_bar);
    if (tid == 0) {
        for (int b = 0; b < 2 * PDG_BIG_STAGES; ++b) big_mbar_init(bar0 + 8u * (unsigned int)b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned int phase_bits = 0u;
    for (int b = blockIdx.x; b < n_list; b += gridDim.x) {
        const long long s = list[b];
        // sorted vertex list of the set: thread t looks at vertices t and t + 256
        int k = 0;
#pragma unroll
        for (int w = 0; w < WMAX; ++w) k += (w < P.W) ? __popcll(bits[(size_t)s * P.W + w]) : 0;
        for (int v = tid; v < 64 * P.W; v += 256) {
            const u64 x = bits[(size_t)s * P.W + (v >> 6)];
            if ((x >> (v & 63)) & 1ull) {
                int rank = __popcll(x & ((1ull << (v & 63)) - 1ull));
                for (int w = 0; w < (v >> 6); ++w) rank += __popcll(bits[(size_t)s * P.W + w]);
                idx[rank] = (unsigned short)v;
            }
        }
        __syncthreads();
        int e = 0;
        double ldA = 0.0, ldD = 0.0;
        if (k > L.kmax) e = ERR_BIGCLIQUE;
        else {
            if (!big_logdet(P.A, P.p, k, smem, L, phase_bits, ldA)) e = ERR_NUMERIC;
            if (P.dmode == 1) {
                if (warp == 0) {
                    double acc = 0.0;
                    for (int j = lane; j < k; j += 32) acc += P.logDdiag[idx[j]];
                    acc = warp_sum(acc);
                    if (lane == 0) red[2] = acc;
                }
                __syncthreads();
                ldD = red[2];
            } else if (P.dmode == 2) {
                if (!big_logdet(P.Dm, P.p, k, smem, L, phase_bits, ldD)) e = ERR_NUMERIC;
            }
        }
        if (tid == 0) {
            out[s] = e ? 0.0 : ggm_score_from(P, k, ldA, ldD);
            if (e) atomicOr(err, e);
        }
        __syncthreads();
    }
}
