// Machine ceilings the scoring kernels are judged against, measured on the box the bench runs on
// (MEASURED_PEAKS.json carries HBM and bf16 tensor throughput only):
//   PDG_PROBE_FP64       fp64 pipe, GFLOP/s: independent DFMA chains on every SM (the ceiling of the
//                        hyper-inverse-Wishart log-det path, wishart.py:17-37)
//   PDG_PROBE_L2_READ    GB/s of 16-byte loads over a 48 MB buffer that stays L2-resident (the
//                        contingency-table scans read level-code columns that 1024 warps share)
//   PDG_PROBE_HBM_READ   GB/s of the same loads over a 4 GB buffer
//   PDG_PROBE_RED_SHARED_FREE / _RANDOM   shared-memory reductions per second (G/s), conflict-free
//                        and with the bank pattern of random cells of a 64-cell x 32-replica table
//                        (the counting step of the scans: one red.shared per data row)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pdg_b200.h"

namespace {

__global__ void __launch_bounds__(256) k_probe_dfma(double *out, int iters, double x, double y)
{
    double a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = (double)(threadIdx.x + q) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = fma(a[q], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += a[q];
    if (s == 123.456) out[0] = s;          // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) k_probe_read(const uint4 *buf, size_t n_vec, int passes, unsigned int *sink)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    unsigned int acc = 0u;
    for (int ps = 0; ps < passes; ++ps) {
        size_t i = tid;
        for (; i + 3 * nth < n_vec; i += 4 * nth) {          // four independent loads in flight per thread
            const uint4 v0 = __ldg(buf + i), v1 = __ldg(buf + i + nth), v2 = __ldg(buf + i + 2 * nth), v3 = __ldg(buf + i + 3 * nth);
            acc += v0.x ^ v1.y ^ v2.z ^ v3.w;
        }
        for (; i < n_vec; i += nth) acc += __ldg(buf + i).x;
    }
    if (acc == 0xDEADBEEFu) sink[0] = acc;
}

// 4 warps per CTA, 4096 counters per warp (the scan's per-warp table).  mode 0: lane l always hits bank l
// (what a <= 128-cell table with 32 replicas does: one wavefront per instruction); mode 1: random cells
// of a 256-cell table with 16 replicas (lanes l and l + 16 share one: the k = 8 binary scan).
__global__ void __launch_bounds__(128) k_probe_red(int iters, int mode, unsigned int *sink)
{
    extern __shared__ unsigned int tab[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4 * 4096; i += 128) tab[i] = 0u;
    __syncthreads();
    const unsigned int base = (unsigned int)__cvta_generic_to_shared(tab + warp * 4096);
    unsigned int s = 0x9E3779B9u * (unsigned int)(threadIdx.x + 1 + blockIdx.x * 128);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            s = s * 1664525u + 1013904223u;
            unsigned int cell;
            if (mode == 0) cell = ((s >> 16) & 127u) * 32u + (unsigned int)lane;
            else cell = ((s >> 16) & 255u) * 16u + (unsigned int)(lane & 15);
            asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(base + 4u * cell) : "memory");
        }
    }
    __syncthreads();
    if (tab[threadIdx.x] == 0xFFFFFFFFu) sink[0] = s;
}

}  // namespace

extern "C" int pdg_probe(int device, int which, double *out)
{
    if (!out) return PDG_E_ARG;
    *out = 0.0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return PDG_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return PDG_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PDG_E_CUDA;
    const int sms = prop.multiProcessorCount;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    int rc = PDG_OK;
    void *d = nullptr; unsigned int *sink = nullptr;
    cudaMalloc((void **)&sink, 64);
    if (which == PDG_PROBE_FP64) {
        const int iters = 1 << 16, grid = sms * 8;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            k_probe_dfma<<<grid, 256>>>((double *)sink, iters, 1.0000001, 1e-9);
            cudaEventRecord(b);
            if (cudaEventSynchronize(b) != cudaSuccess) { rc = PDG_E_CUDA; break; }
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep > 0 && ms < best) best = ms;
        }
        if (!rc) *out = 2.0 * 8.0 * (double)iters * (double)grid * 256.0 / (best * 1e-3) / 1e9;
    } else if (which == PDG_PROBE_L2_READ || which == PDG_PROBE_HBM_READ) {
        const size_t bytes = which == PDG_PROBE_L2_READ ? ((size_t)48 << 20) : ((size_t)4 << 30);
        const int passes = which == PDG_PROBE_L2_READ ? 40 : 1;
        if (cudaMalloc(&d, bytes) != cudaSuccess) { rc = PDG_E_NOMEM; }
        else {
            cudaMemset(d, 1, bytes);
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(a);
                k_probe_read<<<sms * 8, 256>>>((const uint4 *)d, bytes / 16, passes, sink);
                cudaEventRecord(b);
                if (cudaEventSynchronize(b) != cudaSuccess) { rc = PDG_E_CUDA; break; }
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (rep > 0 && ms < best) best = ms;
            }
            if (!rc) *out = (double)bytes * passes / (best * 1e-3) / 1e9;
        }
    } else if (which == PDG_PROBE_RED_SHARED_FREE || which == PDG_PROBE_RED_SHARED_RANDOM) {
        const int iters = 1 << 14, grid = sms * 3;
        cudaFuncSetAttribute(k_probe_red, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4096 * 4);
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            k_probe_red<<<grid, 128, 4 * 4096 * 4>>>(iters, which == PDG_PROBE_RED_SHARED_RANDOM ? 1 : 0, sink);
            cudaEventRecord(b);
            if (cudaEventSynchronize(b) != cudaSuccess) { rc = PDG_E_CUDA; break; }
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep > 0 && ms < best) best = ms;
        }
        if (!rc) *out = 8.0 * (double)iters * (double)grid * 128.0 / (best * 1e-3) / 1e9;
    } else rc = PDG_E_ARG;
    if (d) cudaFree(d);
    cudaFree(sink);
    cudaEventDestroy(a); cudaEventDestroy(b);
    if (rc == PDG_OK && cudaGetLastError() != cudaSuccess) rc = PDG_E_CUDA;
    return rc;
}
