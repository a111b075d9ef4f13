// In-place measurement of the FP64 roofline denominators (pk_fp64_peak): dense DMMA 8x8x4 and DFMA
// issue loops that fill every SM.  MEASURED_PEAKS.json carries only HBM and bf16 numbers.
#include "pk_internal.h"

namespace pk {

__global__ void peak_dfma_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void peak_dmma_kernel(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t measure_fp64_peak(pk_handle* h, double* dmma, double* dfma) {
    const int threads = 512, bps = 2, iters = 4000;
    const int blocks = h->sm_count * bps;
    double* out = nullptr;
    cudaError_t e = cudaMalloc((void**)&out, sizeof(double) * blocks * threads);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best_f = 1e30f, best_m = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        float ms;
        cudaEventRecord(e0, h->stream);
        peak_dfma_kernel<<<blocks, threads, 0, h->stream>>>(out, iters);
        cudaEventRecord(e1, h->stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_f) best_f = ms;
        cudaEventRecord(e0, h->stream);
        peak_dmma_kernel<<<blocks, threads, 0, h->stream>>>(out, iters);
        cudaEventRecord(e1, h->stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_m) best_m = ms;
    }
    if (dfma) *dfma = 2.0 * 8 * iters * (double)threads * blocks / best_f * 1e-9;
    if (dmma) *dmma = 2.0 * 256 * 8 * iters * (double)(threads / 32) * blocks / best_m * 1e-9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError();
}

}  // namespace pk
