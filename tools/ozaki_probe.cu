// FP64 matrix product on the INT8 tensor cores (tcgen05.mma.kind::i8, accumulators in TMEM): stand-alone probe.
//   C(M x 64) = A(M x K) * B(64 x K)^T,  |a|, |b| < 1, FP64 in / FP64 out
// Ozaki splitting: every operand is cut into S = 8 signed 7-bit digits (x = sum_t d_t 2^-7t, exact in FP64), the 36
// digit-pair products with t + u <= 9 run as INT8 MMAs with exact INT32 accumulation, pairs of equal weight share an
// accumulator (8 accumulators x 64 columns = all 512 TMEM columns), and the epilogue adds the 8 levels in FP64.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I pykriging_b200/csrc -o tools/ozaki_probe tools/ozaki_probe.cu
// Prints max |C - C_ref| against an FP64 reference and the INT8 MMA rate.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pk_common.cuh"

using namespace pk;

constexpr int S = 8;                 // digits per operand
constexpr int BM = 128, BN = 64;     // tile: 128 rows of A (TMEM lanes) x 64 rows of B (TMEM columns per level)
constexpr int BK = 64;               // bytes (= INT8 elements) of K per pipeline stage: one 64-byte swizzle row
constexpr int STAGES = 2;
constexpr int A_SLICE = BM * BK, B_SLICE = BN * BK;
constexpr int STAGE_BYTES = S * (A_SLICE + B_SLICE);   // 96 KB
constexpr int THREADS = 192;         // warps 0-3 epilogue (TMEM lanes 32w .. 32w+31), warp 4 TMA + TMEM alloc, warp 5 MMA

__device__ __forceinline__ uint64_t smem_desc_k_sw64(uint32_t saddr) {
    // K-major tile, rows of 64 bytes, SWIZZLE_64B: groups of 8 rows are 512 bytes apart
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
ozaki_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K, int kreps,
             double* __restrict__ C) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024 - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023)) & 1023);
    __shared__ unsigned long long full[STAGES], empty[STAGES], tmem_full;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(&tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int nkb = K / BK;
    const int total = nkb * kreps;
    const int m0 = blockIdx.x * BM;
    if (warp == 4 && lane == 0) {
        // ---- TMA producer ----
        for (int i = 0; i < total; ++i) {
            const int st = i % STAGES;
            if (i >= STAGES) mbar_wait(empty + st, ((i / STAGES) - 1) & 1);
            unsigned char* sa = smem + st * STAGE_BYTES;
            unsigned char* sb = sa + S * A_SLICE;
            mbar_arrive_expect_tx(full + st, (unsigned)STAGE_BYTES);
            const int k0 = (i % nkb) * BK;
            for (int t = 0; t < S; ++t) tma_load_3d(sa + t * A_SLICE, &tmA, k0, m0, t, full + st);
            for (int t = 0; t < S; ++t) tma_load_3d(sb + t * B_SLICE, &tmB, k0, 0, t, full + st);
        }
    } else if (warp == 5 && lane == 0) {
        // ---- MMA issuer ----
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int i = 0; i < total; ++i) {
            const int st = i % STAGES;
            mbar_wait(full + st, (i / STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t sa = (unsigned)__cvta_generic_to_shared(smem + st * STAGE_BYTES);
            const uint32_t sb = sa + S * A_SLICE;
#pragma unroll
            for (int kk = 0; kk < BK / 32; ++kk) {
#pragma unroll
                for (int t = 0; t < S; ++t) {
#pragma unroll
                    for (int u = 0; u + t < S; ++u) {  // digits t+1, u+1: (t+1) + (u+1) <= S + 1
                        const int level = t + u;
                        const uint32_t first = (i == 0 && kk == 0 && (t == 0 || u == level)) ? 0u : 1u;
                        // the first pair that touches a level in the first chunk overwrites the accumulator: pair (t = 0, u = level)
                        umma_i8(tmem_base + level * BN, smem_desc_k_sw64(sa + t * A_SLICE + kk * 32),
                                smem_desc_k_sw64(sb + u * B_SLICE + kk * 32), idesc, (i == 0 && kk == 0 && t == 0) ? 0u : 1u);
                        (void)first;
                    }
                }
            }
            umma_commit(empty + st);
        }
        umma_commit(&tmem_full);
    } else if (warp < 4) {
        // ---- epilogue: 8 levels of INT32 -> FP64 ----
        mbar_wait(&tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int row = m0 + warp * 32 + lane;
        for (int half = 0; half < 2; ++half) {
            double v[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = 0.0;
            for (int level = S - 1; level >= 0; --level) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + level * BN + half * 32, r);
                const double sc = ldexp(1.0, -7 * (level + 2));
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] += (double)(int)r[c] * sc;
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) C[(size_t)row * BN + half * 32 + c] = v[c];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base));
}

static void slice(const std::vector<double>& X, int rows, int K, std::vector<int8_t>& planes) {
    planes.assign((size_t)S * rows * K, 0);
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < K; ++k) {
            double y = X[(size_t)r * K + k];
            for (int t = 0; t < S; ++t) {
                y *= 128.0;
                const double d = nearbyint(y);
                planes[((size_t)t * rows + r) * K + k] = (int8_t)d;
                y -= d;
            }
        }
}

static CUresult make_map(CUtensorMap* tm, void* base, int K, int rows, int box_rows) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    const cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)S};
    const cuuint64_t gstr[2] = {(cuuint64_t)K, (cuuint64_t)K * rows};
    const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    return ((encode_fn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const int K = (argc > 1) ? atoi(argv[1]) : 512;
    const int tiles = (argc > 2) ? atoi(argv[2]) : sms;
    const int M = tiles * BM;
    std::vector<double> A((size_t)M * K), B((size_t)BN * K);
    srand(7);
    auto rnd = [] { return (rand() / (double)RAND_MAX) * 1.98 - 0.99; };
    for (auto& v : A) v = rnd() * ((rand() & 3) ? 1.0 : 1e-4);
    for (auto& v : B) v = rnd();
    std::vector<int8_t> pa, pb;
    slice(A, M, K, pa);
    slice(B, BN, K, pb);
    int8_t *dA, *dB;
    double* dC;
    cudaMalloc(&dA, pa.size());
    cudaMalloc(&dB, pb.size());
    cudaMalloc(&dC, sizeof(double) * (size_t)M * BN);
    cudaMemcpy(dA, pa.data(), pa.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, pb.data(), pb.size(), cudaMemcpyHostToDevice);
    CUtensorMap tmA, tmB;
    if (make_map(&tmA, dA, K, M, BM) != CUDA_SUCCESS || make_map(&tmB, dB, K, BN, BN) != CUDA_SUCCESS) {
        printf("tensor map encode failed\n");
        return 1;
    }
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
    cudaFuncSetAttribute(ozaki_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ozaki_kernel<<<tiles, THREADS, smem>>>(tmA, tmB, K, 1, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
        return 1;
    }
    std::vector<double> Cg((size_t)M * BN);
    cudaMemcpy(Cg.data(), dC, sizeof(double) * Cg.size(), cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int r = 0; r < M; r += 7)
        for (int c = 0; c < BN; ++c) {
            long double s = 0;
            for (int k = 0; k < K; ++k) s += (long double)A[(size_t)r * K + k] * B[(size_t)c * K + k];
            maxerr = fmax(maxerr, fabs((double)(s - (long double)Cg[(size_t)r * BN + c])));
            maxref = fmax(maxref, fabs((double)s));
        }
    printf("{\"K\": %d, \"tiles\": %d, \"max_abs_err\": %.3e, \"max_abs_ref\": %.3e, \"fp64_eps_times_sqrtK\": %.3e}\n", K, tiles, maxerr,
           maxref, 1.1e-16 * sqrt((double)K));
    // throughput: repeat the K loop
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int kreps : {8, 64}) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            float ms;
            cudaEventRecord(e0);
            ozaki_kernel<<<tiles, THREADS, smem>>>(tmA, tmB, K, kreps, dC);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double macs = (double)tiles * BM * BN * (double)K * kreps;           // FP64-equivalent multiply-adds
        printf("{\"kreps\": %d, \"ms\": %.3f, \"fp64_equiv_tflops\": %.2f, \"int8_tops\": %.1f}\n", kreps, best,
               2.0 * macs / best * 1e-9, 2.0 * macs * 36 / best * 1e-9);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
