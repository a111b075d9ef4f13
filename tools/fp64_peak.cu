// Measures the FP64 roofline denominators MEASURED_PEAKS.json does not carry:
// DFMA (vector) and DMMA (mma.sync m8n8k4.f64, "FP64 tensor core") throughput of the whole GPU.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak tools/fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dmma_kernel(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half of the warps issue DMMA, the other half DFMA: do the two share execution resources?
__global__ void mixed_kernel(double* out, int iters) {
    const int warp = threadIdx.x >> 5;
    if (warp & 1) {
        double c[8][2];
        for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
        double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
        double s = 0;
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
        const double b = 1.0000001, c = 1e-9;
        for (int i = 0; i < iters * 8; ++i) {  // 8x the iterations: same stand-alone duration as the DMMA warps
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    }
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            float best_f = 1e30f, best_m = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                float ms;
                cudaEventRecord(e0); dfma_kernel<<<sms * bps, threads>>>(out, iters); cudaEventRecord(e1);
                cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best_f) best_f = ms;
                cudaEventRecord(e0); dmma_kernel<<<sms * bps, threads>>>(out, iters); cudaEventRecord(e1);
                cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best_m) best_m = ms;
            }
            const double fl_f = 2.0 * 8 * iters * (double)threads * sms * bps;
            const double fl_m = 2.0 * 256 * 8 * iters * (double)(threads / 32) * sms * bps;
            printf("{\"sms\": %d, \"threads\": %d, \"blocks_per_sm\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f}\n",
                   sms, threads, bps, fl_f / best_f * 1e-9, fl_m / best_m * 1e-9);
        }
    }
    {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            float ms;
            cudaEventRecord(e0); mixed_kernel<<<sms * 2, 512>>>(out, iters); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        // per block: 8 warps DMMA (8 x 256 FMA x iters each), 8 warps DFMA (8 x 32 FMA x iters each)
        const double fl_m = 2.0 * 256 * 8 * iters * 8.0 * sms * 2;
        const double fl_f = 2.0 * 8 * (8.0 * iters) * 256.0 * sms * 2;
        printf("{\"mixed_ms\": %.3f, \"mixed_dmma_tflops\": %.2f, \"mixed_dfma_tflops\": %.2f, \"note\": \"concurrent DMMA+DFMA warps; both rates are over the same interval\"}\n",
               best, fl_m / best * 1e-9, fl_f / best * 1e-9);
    }
    // sustained (2 s) DMMA
    {
        cudaEventRecord(e0);
        int reps = 0;
        float ms = 0;
        while (ms < 2000.f) {
            dmma_kernel<<<sms * 2, 512>>>(out, iters);
            ++reps;
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        const double fl = 2.0 * 256 * 8 * iters * 16.0 * sms * 2 * reps;
        printf("{\"dmma_tflops_sustained\": %.2f, \"seconds\": %.2f}\n", fl / ms * 1e-9, ms * 1e-3);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
