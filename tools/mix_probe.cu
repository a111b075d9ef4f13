// Does a CTA of FP64-ALU work (the table-driven 2^x chains of the Psi kernel) fill the FP64 pipe that a CTA of DMMA work
// (the panel loop of the blocked Cholesky) leaves idle on the same SM?  Decides whether Psi construction is worth fusing
// into the task-mode Cholesky kernel as one more task type.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I pykriging_b200/csrc -o tools/mix_probe tools/mix_probe.cu
// Roles per CTA (256 threads, ~100 KB dynamic shared memory each so that exactly two are resident per SM):
//   D: DMMA loop fed by shared-memory fragment loads (2 x 8 accumulator tiles per warp, like cta::gemm_consume)
//   P: per thread K = 10 table-driven 2^x + one e^-S per "entry" (pk_math.cuh, the instruction mix of psi_pop_kernel)
// Modes: DD (two D CTAs per SM), PP, DP (one of each per SM).  Reports work per second of each role.
#include <cstdio>
#include <cuda_runtime.h>

#include "pk_math.cuh"

using namespace pk;

constexpr int THREADS = 256;
constexpr int K = 10;

__device__ __forceinline__ void role_dmma(double* smem, int iters, double* out, int stall) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* As = smem;                 // 128 x 16
    double* Bs = smem + 128 * 16;      // 64 x 16
    for (int i = tid; i < 192 * 16; i += THREADS) smem[i] = 1e-3 * (i % 97);
    __syncthreads();
    double acc[2][8][2];
    for (int mi = 0; mi < 2; ++mi)
        for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const int fr = lane >> 2, fc = lane & 3;
    for (int it = 0; it < iters; ++it) {
        if (stall > 0) {  // a pipeline bubble per chunk (the real kernel waits for TMA data / its partner warps here)
            const long long t = clock64();
            while (clock64() - t < stall) {}
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int off = ((i ^ (fr & 3)) << 2) + fc;
            double a[2], b[8];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) a[mi] = As[(warp * 16 + mi * 8 + fr) * 16 + off];
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) b[ni] = Bs[(ni * 8 + fr) * 16 + off];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    double s = 0;
    for (int mi = 0; mi < 2; ++mi)
        for (int ni = 0; ni < 8; ++ni) s += acc[mi][ni][0] + acc[mi][ni][1];
    out[blockIdx.x * THREADS + tid] = s;
}

__device__ __forceinline__ void role_psi(double* smem, int iters, double* out) {
    const int tid = threadIdx.x;
    const unsigned base = (unsigned)__cvta_generic_to_shared(smem);
    unsigned tab = (base + 0x7fffu) & ~0x7fffu;
    asm volatile("" : "+r"(tab));
    double* tab4k = smem + (tab - base) / sizeof(double);
    double2* thp = reinterpret_cast<double2*>(tab4k + 4096);  // 128 candidates x K (theta, p) behind the table
    fill_exp2_4096(tab4k, tid, THREADS);
    for (int i = tid; i < 128 * K; i += THREADS) thp[i] = make_double2(0.5 + 1e-3 * (i % 61), 1.0 + 0.9 * ((i * 7) % 11) / 11.0);
    __syncthreads();
    double lh[K], ll[K];
#pragma unroll
    for (int d = 0; d < K; ++d) {
        lh[d] = -0.3 - 0.01 * ((tid * (d + 3)) % 251);
        ll[d] = 1e-18 * tid;
    }
    double total = 0.0;
    for (int it = 0; it < iters; ++it) {
        const double2* tpg = thp + (it & 127) * K;
        double r[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const double2 tp = tpg[d];
            r[d] = tp.x * exp2_lean4k(tp.y, lh[d], ll[d], tab);
        }
#pragma unroll
        for (int d = 8; d < K; ++d) {
            const double2 tp = tpg[d];
            r[d & 7] = fma(tp.x, exp2_lean4k(tp.y, lh[d], ll[d], tab), r[d & 7]);
        }
        const double S = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        total += expneg_4k(S, tab);
    }
    out[blockIdx.x * THREADS + tid] = total;
}

// mode 0: all D, 1: all P, 2: one D and one P CTA on every SM (the role follows the order of arrival on the SM)
__global__ void __launch_bounds__(THREADS, 2) mix_kernel(int mode, int sms, int iters_d, int iters_p, double* out, long long* clk,
                                                         int* arrivals, int stall) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int slot_s;
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        slot_s = atomicAdd(arrivals + smid, 1);
    }
    __syncthreads();
    const int slot = slot_s;
    const bool is_p = (mode == 1) || (mode == 2 && (slot & 1));
    if (mode == 3 && (slot & 1)) return;  // mode 3: one D CTA per SM, alone
    const long long t0 = clock64();
    if (is_p) role_psi(smem, iters_p, out);
    else role_dmma(smem, iters_d, out, stall);
    if (threadIdx.x == 0) clk[blockIdx.x] = (clock64() - t0) * 2 + (is_p ? 1 : 0);  // role in the low bit
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const size_t smem = 100 * 1024;
    cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    double* out;
    long long* clk;
    cudaMalloc(&out, sizeof(double) * sms * 2 * THREADS);
    cudaMalloc(&clk, sizeof(long long) * sms * 2);
    int* arrivals;
    cudaMalloc(&arrivals, sizeof(int) * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int it_d = 40000, it_p = 30000;
    long long* h = new long long[sms * 2];
    int stall = 0;
    auto run = [&](int mode, int itd, int itp, const char* name) {
        float best = 1e30f;
        double cd = 0, cp = 0;
        for (int rep = 0; rep < 3; ++rep) {
            float ms;
            cudaMemset(arrivals, 0, sizeof(int) * 1024);
            cudaEventRecord(e0);
            mix_kernel<<<sms * 2, THREADS, smem>>>(mode, sms, itd, itp, out, clk, arrivals, stall);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) {
                best = ms;
                cudaMemcpy(h, clk, sizeof(long long) * sms * 2, cudaMemcpyDeviceToHost);
                cd = cp = 0;
                int nd = 0, np = 0;
                for (int b = 0; b < 2 * sms; ++b) {
                    if (h[b] & 1) { cp += (double)(h[b] >> 1); ++np; }
                    else { cd += (double)(h[b] >> 1); ++nd; }
                }
                cd /= (nd ? nd : 1);
                cp /= (np ? np : 1);
            }
        }
        // per CTA: D = iters * 4 ksteps * 16 DMMA * 8 warps; P = iters * 256 entries
        printf("{\"stall\": %d, \"mode\": \"%s\", \"ms\": %.3f, \"iters_d\": %d, \"iters_p\": %d, \"avg_cycles_D_cta\": %.0f, \"avg_cycles_P_cta\": %.0f}\n",
               stall, name, best, itd, itp, cd, cp);
        return best;
    };
  for (int st : {0, 300, 600, 1000}) {
    stall = st;
    run(3, 2 * it_d, 0, "D_alone");
    const float dd = run(0, it_d, 0, "DD");          // 2 x sms CTAs of D
    const float pp = run(1, 0, it_p, "PP");          // 2 x sms CTAs of P
    // the same total work, one role per resident CTA: every D CTA does twice the iterations (there are half as many), same for P
    const float dp = run(2, 2 * it_d, 2 * it_p, "DP");
    printf("{\"serial_ms\": %.3f, \"mixed_ms\": %.3f, \"gain\": %.3f}\n", dd + pp, dp, (dd + pp) / dp);
    // sweep the balance: how long does the P CTA take while a D CTA runs beside it, and vice versa
    run(2, 2 * it_d, it_p, "DP_halfP");
    run(2, it_d, 2 * it_p, "DP_halfD");
  }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        printf("CUDA error %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
