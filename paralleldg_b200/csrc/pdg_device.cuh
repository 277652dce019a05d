// Device-side helpers shared by the parallelDG B200 kernels: Philox4x32-10 streams, bitset
// selection, warp reductions and the packed-triangular LDL^T log-determinant used by the
// hyper-inverse-Wishart clique scorer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define PDG_FULL 0xffffffffu

// Random streams (production mode).  key = (seed lo, seed hi); counter = (global chain id,
// MCMC iteration, tag, index).  One Philox block per draw keeps every draw addressable, so any
// thread can produce any draw without communication.
enum { TAG_ITER = 0, TAG_MOVE = 1, TAG_PRUF_V = 2, TAG_PRUF_W = 3, TAG_PERM = 4, TAG_PAD = 5 };

__device__ __forceinline__ void philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32 out[4])
{
    // rounds deliberately NOT unrolled: the MH kernel is instruction-fetch bound and calls this from
    // several places; the dynamic instruction count is the same
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        const u32 hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const u32 hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const u32 n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ u32 draw_int(u32 k0, u32 k1, u32 chain, u32 iter, u32 tag, u32 idx, u32 n)
{
    u32 o[4];
    philox4x32_10(chain, iter, tag, idx, k0, k1, o);
    return __umulhi(o[0], n);
}

// 53-bit uniform in [0,1): (a >> 5) * 2^26 + (b >> 6), over 2^53
__device__ __forceinline__ double u01_from(u32 a, u32 b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ double draw_unif(u32 k0, u32 k1, u32 chain, u32 iter, u32 tag, u32 idx)
{
    u32 o[4];
    philox4x32_10(chain, iter, tag, idx, k0, k1, o);
    return u01_from(o[0], o[1]);
}

// the same draw with the ten rounds unrolled: this one sits on the per-iteration path of the MH
// kernel (one uniform per proposal), where the loop counter, the key schedule and the register moves
// of the rolled form are half of its instructions
__device__ __forceinline__ double draw_unif_unrolled(u32 k0, u32 k1, u32 c0, u32 c1, u32 c2, u32 c3)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const u32 hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const u32 hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const u32 n0 = hi1 ^ c1 ^ (k0 + 0x9E3779B9u * (u32)r), n2 = hi0 ^ c3 ^ (k1 + 0xBB67AE85u * (u32)r);
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return u01_from(c0, c1);
}

// ---- bitsets ---------------------------------------------------------------------------
__device__ __forceinline__ int bit32(const u32 *w, int i) { return (w[i >> 5] >> (i & 31)) & 1u; }
__device__ __forceinline__ int bit64(const u64 *w, int i) { return (int)((w[i >> 6] >> (i & 63)) & 1ull); }

// position of the n-th (0-based) set bit of x; n < popc(x)
__device__ __forceinline__ int select64(u64 x, int n)
{
    const u32 lo = (u32)x;
    const int c = __popc(lo);
    if (n < c) return (int)__fns(lo, 0, n + 1);
    return 32 + (int)__fns((u32)(x >> 32), 0, n - c + 1);
}

__device__ __forceinline__ int select32(u32 x, int n) { return (int)__fns(x, 0, n + 1); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PDG_FULL, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PDG_FULL, v, o);
    return v;
}

// packed lower triangle: element (i, l), l <= i
__device__ __forceinline__ int tri(int i, int l) { return ((i * (i + 1)) >> 1) + l; }

// Right-looking LDL^T of the k x k SPD matrix held as a packed lower triangle in `a` (shared
// memory for k <= 32, global scratch above), rows dealt round-robin to the 32 lanes.  The index
// order is [S0 (ks) | B \ S0 (kb - ks) | v] so that ONE factorisation yields the four
// log-determinants a move needs:
//   ld(S0)   = sum_{j<ks} log d_j           ld(B)    = sum_{j<kb} log d_j
//   ld(B+v)  = ld(B) + log d_{k-1}          ld(S0+v) = ld(S0) + log(Schur complement of v on S0)
// where the Schur complement is the v-row diagonal after ks pivots.  When the set has no v
// (kb == k) callers ignore ldv / lschur.  Returns false on a non-positive pivot.
__device__ __forceinline__ bool ldl_logdets(double *a, int lane, int k, int ks, int kb,
                                            double &sumS, double &sumB, double &ldv, double &lschur)
{
    double schur = 1.0;
    const int vlane = (k - 1) & 31;
    for (int j = 0; j < k; ++j) {
        __syncwarp();
        const double d = a[tri(j, j)];
        if (j == ks && lane == vlane) schur = a[tri(k - 1, k - 1)];
        if (j < k - 1) {
            const double inv = 1.0 / d;
            int i = j + 1 + ((lane - (j + 1)) & 31);
            for (; i < k; i += 32) {
                const double f = a[tri(i, j)] * inv;
                double *row = a + tri(i, 0);
                // four elements at a time, loads before stores (row i and column j never overlap:
                // l > j), same multiply-subtract per element
                int l = j + 1, cj = tri(j + 1, j);
                for (; l + 3 <= i; l += 4) {
                    const double c0 = a[cj], c1 = a[cj + l + 1], c2 = a[cj + 2 * l + 3], c3 = a[cj + 3 * l + 6];
                    double r0 = row[l], r1 = row[l + 1], r2 = row[l + 2], r3 = row[l + 3];
                    r0 -= f * c0; r1 -= f * c1; r2 -= f * c2; r3 -= f * c3;
                    row[l] = r0; row[l + 1] = r1; row[l + 2] = r2; row[l + 3] = r3;
                    cj += 4 * l + 10;
                }
                for (; l <= i; ++l) { row[l] -= f * a[cj]; cj += l + 1; }
            }
        }
    }
    __syncwarp();
    double accS = 0.0, accB = 0.0;
    bool ok = true;
    for (int j = lane; j < k; j += 32) {
        const double d = a[tri(j, j)];
        if (!(d > 0.0)) ok = false;
        const double l = log(d);
        if (j < ks) accS += l;
        if (j < kb) accB += l;
    }
    sumS = warp_sum(accS);
    sumB = warp_sum(accB);
    const double dv = a[tri(k - 1, k - 1)];
    ldv = log(dv);
    const double sc = __shfl_sync(PDG_FULL, schur, vlane);
    if (!(sc > 0.0)) ok = false;
    lschur = log(sc);
    return __all_sync(PDG_FULL, ok);
}
