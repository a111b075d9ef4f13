// Dependent-issue latency of the FP64 pipe on B200 (one warp per SM sub-partition, clock64 around a chain), and the DFMA rate of
// W warps per scheduler each running I independent chains: how many chains in flight saturate the pipe?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency_probe tools/fp64_latency_probe.cu && ./fp64_latency_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int I>
__global__ void chains(double* out, long long* cyc, int iters) {
    double a[I];
    for (int i = 0; i < I; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double b = 1.0000001, c = 1e-9;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < I; ++i) a[i] = fma(a[i], b, c);
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < I; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void dmma_chain(double* out, long long* cyc, int iters) {
    double c0 = 0, c1 = 0, a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int I>
static void run(int sms, double* out, long long* cyc, int warps_per_sched) {
    const int iters = 4096, threads = 128 * warps_per_sched;
    long long h = 0;
    chains<I><<<sms, threads>>>(out, cyc, iters);
    chains<I><<<sms, threads>>>(out, cyc, iters);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"warps_per_scheduler\": %d, \"chains_per_warp\": %d, \"cycles_per_dfma_per_scheduler\": %.2f, \"cycles_per_chain_step\": %.1f}\n",
           warps_per_sched, I, (double)h / ((double)iters * I * warps_per_sched), (double)h / iters);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * 2048);
    cudaMalloc(&cyc, sizeof(long long));
    for (int w : {1, 2, 4, 6, 8}) {
        run<1>(sms, out, cyc, w);
        run<2>(sms, out, cyc, w);
        run<4>(sms, out, cyc, w);
        run<8>(sms, out, cyc, w);
    }
    long long h = 0;
    dmma_chain<<<sms, 128>>>(out, cyc, 4096);
    dmma_chain<<<sms, 128>>>(out, cyc, 4096);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"dmma_dependent_latency_cycles\": %.1f}\n", (double)h / 4096);
    return 0;
}
