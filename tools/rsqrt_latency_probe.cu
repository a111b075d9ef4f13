// Dependent latencies of the pieces of the 8 x 8 pivot chain (pk_chol_tile.cuh diag_block_factor) on B200: one warp per SM,
// clock64 around a chain of 4096 dependent operations.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rsqrt_latency_probe tools/rsqrt_latency_probe.cu && ./rsqrt_latency_probe
#include <cstdio>
#include <cuda_runtime.h>

enum { LIB_RSQRT, APPROX_RSQRT, F32_RSQRT, DFMA, DMUL, SHFL64, LDS64, LIB_SQRT, LIB_RCP, F32_SEEDED, APPROX_RCP };

__device__ __forceinline__ double approx_rsqrt(double x) {
    double y;
    asm volatile("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double approx_rcp(double x) {
    double y;
    asm volatile("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
    return y;
}

template <int WHAT>
__global__ void chain(double* out, long long* cyc, int iters, double x0) {
    __shared__ double sm[64];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    double x = x0 + threadIdx.x * 1e-12;
    unsigned idx = 0;
    const long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < iters; ++it) {
        if (WHAT == LIB_RSQRT) x = rsqrt(x);
        if (WHAT == APPROX_RSQRT) x = approx_rsqrt(x);
        if (WHAT == APPROX_RCP) x = approx_rcp(x);
        if (WHAT == F32_RSQRT) x = (double)rsqrtf((float)x);
        if (WHAT == F32_SEEDED) {  // fp32 seed + one third-order step (what a replacement of the library rsqrt would cost)
            const double y = (double)rsqrtf((float)x);
            const double e = fma(-(y * y), x, 1.0);
            x = fma(fma(e, 0.375, 0.5), y * e, y);
        }
        if (WHAT == DFMA) x = fma(x, 1.0000001, 1e-9);
        if (WHAT == DMUL) x = x * 1.0000001;
        if (WHAT == SHFL64) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
        if (WHAT == LDS64) {
            x += sm[idx];
            idx = (unsigned)__double2loint(x) & 63;
        }
        if (WHAT == LIB_SQRT) x = sqrt(x);
        if (WHAT == LIB_RCP) x = 1.0 / x;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int WHAT>
static void run(const char* name, double* out, long long* cyc, double x0) {
    const int iters = 4096;
    long long h = 0;
    chain<WHAT><<<8, 32>>>(out, cyc, iters, x0);
    chain<WHAT><<<8, 32>>>(out, cyc, iters, x0);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"op\": \"%s\", \"cycles_per_dependent_op\": %.1f}\n", name, (double)h / iters);
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * 8 * 32);
    cudaMalloc(&cyc, sizeof(long long));
    run<DFMA>("dfma", out, cyc, 1.0);
    run<DMUL>("dmul", out, cyc, 1.0);
    run<LIB_RSQRT>("rsqrt(double) (library: MUFU.RSQ64H + third-order step)", out, cyc, 1.7);
    run<APPROX_RSQRT>("rsqrt.approx.ftz.f64 (MUFU.RSQ64H)", out, cyc, 1.7);
    run<APPROX_RCP>("rcp.approx.ftz.f64 (MUFU.RCP64H)", out, cyc, 1.7);
    run<F32_RSQRT>("double -> float, rsqrtf, float -> double", out, cyc, 1.7);
    run<F32_SEEDED>("fp32 seed + third-order step", out, cyc, 1.7);
    run<LIB_SQRT>("sqrt(double)", out, cyc, 1.7);
    run<LIB_RCP>("1.0 / x", out, cyc, 1.7);
    run<SHFL64>("shfl of a double (2 x SHFL.32)", out, cyc, 1.0);
    run<LDS64>("dadd + LDS.64 + index", out, cyc, 0.0);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}
