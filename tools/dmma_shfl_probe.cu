// Does shuffle traffic cap the DMMA rate?  The triangular solve of the INT8 Cholesky (pk_chol_i8.cu solve<>) runs its DMMAs at
// about 45 % of the FP64 tensor peak with "math pipe throttle" as its top stall, whether 8 or 16 warps are in flight; between
// the DMMAs it re-lays accumulators out as A fragments with warp shuffles (1.7 SHFL per DMMA).  Variants, 8 warps per SM:
//   0  DMMA only, operands rotate over 8 A and 8 B registers
//   1  + 2 SHFL.32 per DMMA     2  + 4 SHFL.32 per DMMA     3  + one LDS.64 per DMMA     4  + 2 FSEL-like selects per DMMA
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_shfl_probe tools/dmma_shfl_probe.cu && ./dmma_shfl_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int V>
__global__ void probe(double* out, int iters) {
    __shared__ double sm[1024];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    double c[8][2], a[8], b[8];
    for (int i = 0; i < 8; ++i) {
        c[i][0] = c[i][1] = 0.0;
        a[i] = threadIdx.x * 1e-3 + i;
        b[i] = 1.0 - threadIdx.x * 1e-3 - i;
    }
    double x = lane * 0.5, y = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dmma(c[i][0], c[i][1], a[i], b[(i + it) & 7]);
            if (V == 1 || V == 2) {
                x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);   // a double shuffle = 2 SHFL.32
                if (V == 2) y = __shfl_sync(0xffffffffu, x, lane ^ 2);
            }
            if (V == 3) y += sm[(lane * 8 + i + it) & 1023];
            if (V == 4) { x = (lane & 1) ? y : x; y = (lane & 2) ? x : y + 1.0; }
        }
    }
    double s = x + y;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
static void run(int sms, double* out, const char* what) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        float ms;
        cudaEventRecord(e0);
        probe<V><<<sms, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fl = 2.0 * 256 * 8 * iters * (double)(threads / 32) * sms;
    printf("{\"variant\": %d, \"what\": \"%s\", \"dmma_tflops\": %.2f}\n", V, what, fl / best * 1e-9);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 1024);
    run<0>(sms, out, "DMMA only, rotating operands");
    run<1>(sms, out, "+ 2 SHFL.32 per DMMA");
    run<2>(sms, out, "+ 4 SHFL.32 per DMMA");
    run<3>(sms, out, "+ 1 LDS.64 + DADD per DMMA");
    run<4>(sms, out, "+ selects / DADD per DMMA");
    return 0;
}
