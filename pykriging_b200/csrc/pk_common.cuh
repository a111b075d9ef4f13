// Shared device helpers for the pykriging_b200 kernels (sm_100a, FP64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#define PK_MAX_K 64          // largest supported number of design variables
#define PK_TILE 64           // tile edge of the blocked Cholesky (panel width nb)
#define PK_AUG_ROWS 8        // augmented rows appended under Psi (row 0 = 1^T, row 1 = y^T, rest 0)
#define PK_SMALL_N_MAX 128   // n <= this: one CTA per candidate, Psi never leaves shared memory

// diag extra of the interpolating model: np.spacing(1) (matrixops.py:30)
#define PK_EPS_DIAG 2.220446049250313e-16

namespace pk {

// ------------------------------------------------------------------------------------------------
// Summation in NumPy's order.  The reference reduces theta*|d|^p over the contiguous last axis with
// np.sum (matrixops.py:28,70): plain left-to-right for k < 8, otherwise NumPy's 8-accumulator
// pairwise scheme (pairwise_sum, block <= 128).  Reproducing the order keeps the exponent of
// exp(-S) identical to the reference whenever the individual terms are.
// `term(d)` returns the d-th addend.
// ------------------------------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ double sum_numpy_order(int k, F term) {
    if (k < 8) {
        double res = 0.0;
        for (int d = 0; d < k; ++d) res += term(d);
        return res;
    }
    double r0 = term(0), r1 = term(1), r2 = term(2), r3 = term(3);
    double r4 = term(4), r5 = term(5), r6 = term(6), r7 = term(7);
    int i = 8;
    const int lim = k - (k % 8);
    for (; i < lim; i += 8) {
        r0 += term(i + 0); r1 += term(i + 1); r2 += term(i + 2); r3 += term(i + 3);
        r4 += term(i + 4); r5 += term(i + 5); r6 += term(i + 6); r7 += term(i + 7);
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < k; ++i) res += term(i);
    return res;
}

// |d|^p exactly as np.power would treat the special values the path can produce:
// pow(0, p>0) = 0, pow(d, 1) = d, pow(d, 2) = d*d (correctly rounded).
__device__ __forceinline__ double pow_abs(double d, double p) {
    if (p == 2.0) return d * d;
    if (p == 1.0) return d;
    return pow(d, p);
}

// S(x, z) = sum_d theta_d |x_d - z_d|^p_d  (matrixops.py:28 / :70), NumPy summation order.
__device__ __forceinline__ double corr_exponent(const double* __restrict__ x, const double* __restrict__ z,
                                                const double* __restrict__ theta, const double* __restrict__ p,
                                                int k) {
    return sum_numpy_order(k, [&](int d) { return theta[d] * pow_abs(fabs(x[d] - z[d]), p[d]); });
}

// Expected improvement / weighted expected improvement exactly as krige.py:207-217 / 227-233 (operation order kept):
// S <= 0 -> 0;  u = y_min - yhat;  EI = u (0.5 + 0.5 erf(u / (S sqrt 2))) + S / sqrt(2 pi) exp(-u^2 / (2 S^2)); w >= 0
// weights the two terms with w and 1 - w.
__device__ __forceinline__ double expected_improvement(double yhat, double S, double y_min, double w) {
    if (S <= 0.0) return 0.0;
    const double u = y_min - yhat;
    const double one = u * (0.5 + 0.5 * erf((1.0 / sqrt(2.0)) * (u / S)));
    const double two = (S * (1.0 / sqrt(2.0 * CUDART_PI))) * exp(-(1.0 / 2.0) * ((u * u) / (S * S)));
    if (w < 0.0) return one + two;
    return w * one + (1.0 - w) * two;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024); `scratch` holds >= 32 doubles.
// Every thread receives the result.  Deterministic (fixed tree).
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double t = (lane < nw) ? scratch[lane] : 0.0;
    t = warp_sum(t);
    return t;
}

// ------------------------------------------------------------------------------------------------
// small PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
// ---- bulk asynchronous copies (TMA engine, no tensor map) with mbarrier completion ---------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// one arrival + the number of bytes the bulk copies of this phase will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PK_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PK_MBAR_DONE_%=;\n"
        "bra PK_MBAR_WAIT_%=;\n"
        "PK_MBAR_DONE_%=:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
// TMA tile copy global -> shared through a 3-D tensor map (coordinates: innermost first)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}
// the same with an L2 eviction policy (createpolicy): tiles that are re-read soon stay, streamed ones go first
__device__ __forceinline__ void tma_load_4d_hint(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, unsigned long long* bar,
                                                 unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"((unsigned)__cvta_generic_to_shared(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, int c4, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}
// orders this thread's earlier generic-proxy accesses (global and shared) before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

// D(8x8) += A(8x4, row) * B(4x8, col) on the FP64 tensor pipe.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}


}  // namespace pk
