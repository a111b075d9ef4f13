// Digit conventions of the INT8-tensor-core contraction (pk_predict_i8.cu), shared with the psi kernels that write the
// digit planes of psi directly (pk_predict2.cu).
#pragma once
#include <stdint.h>

namespace pk {
namespace i8 {

constexpr int S = 7;                  // base-256 digits per operand
// psi lies in [0, 1]: X = round(psi 2^54) is a 55-bit unsigned integer and its 7 bytes ARE the digits (unsigned, most
// significant first: 64 for psi = 1), i.e. psi = 4 * sum_t d_t 256^-t
constexpr double PSI_TO_FIXED = 18014398509481984.0;  // 2^54
constexpr double PSI_SCALE = 0.25;                    // psi 2^54 = (psi / 4) 2^56: the planes hold psi / 4

// byte t (0 = most significant digit) of the fixed-point psi value
__device__ __forceinline__ unsigned long long psi_fixed(double psi) { return (unsigned long long)__double2ll_rn(psi * PSI_TO_FIXED); }
__device__ __forceinline__ uint8_t psi_digit(unsigned long long X, int t) { return (uint8_t)(X >> (8 * (S - 1 - t))); }

// x (|x| < 0.5 - 2^-9) -> S balanced base-256 digits of round(x 2^56), most significant first: x = sum_t d_t 256^-(t+1),
// -128 <= d_t <= 127
__device__ __forceinline__ void balanced_digits(double x, int (&d)[S]) {
    long long X = __double2ll_rn(x * 72057594037927936.0);  // 2^56
#pragma unroll
    for (int t = S - 1; t >= 1; --t) {
        const int r = (int)((X + 128) & 255) - 128;
        d[t] = r;
        X = (X - r) >> 8;
    }
    d[0] = (int)X;
}

// ---- tcgen05 wrappers (INT8 MMA with TMEM accumulators) ---------------------------------------------------------
// K-major tile, rows of 64 bytes, SWIZZLE_64B (what TMA writes with CU_TENSOR_MAP_SWIZZLE_64B): 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t smem_desc_k_sw64(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// K-major tile, rows of 32 bytes, SWIZZLE_32B (CU_TENSOR_MAP_SWIZZLE_32B): 8-row groups 256 bytes apart
__device__ __forceinline__ uint64_t smem_desc_k_sw32(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// instruction descriptor: D = S32, A / B 8-bit (signed flags), both K-major, M x N
__host__ __device__ constexpr uint32_t idesc_i8(bool a_signed, bool b_signed, int M, int N) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// 16 TMEM lanes x 16 columns in the layout of two m16n8 accumulator tiles (16x256b, repeated twice): lane (fr = lane / 4,
// q = lane % 4) receives r[4 g + 2 h + e] = element (row fr + 8 h, column 8 g + 2 q + e); taddr names the first of the 16 lanes
__device__ __forceinline__ void tmem_ld_frag16(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// 16 columns of a row of the product from its S level accumulators (level L = sum of the digit-pair products of weight
// 256^-(L+2)): smallest weights first, levels pairwise in integer arithmetic ((acc_L << 8) + acc_{L+1} has at most 39 bits)
__device__ __forceinline__ void levels_to_fp64(uint32_t lane_base, int col0, int ncols_per_level, double (&v)[16]) {
    {
        uint32_t last[16];
        tmem_ld16(lane_base + (S - 1) * ncols_per_level + col0, last);
        tmem_ld_wait();
        const double w = ldexp(1.0, -8 * (S + 1));
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (double)(int)last[j] * w;
    }
#pragma unroll
    for (int lp = (S - 1) / 2 - 1; lp >= 0; --lp) {
        uint32_t hi[16], lo[16];
        tmem_ld16(lane_base + (2 * lp) * ncols_per_level + col0, hi);
        tmem_ld16(lane_base + (2 * lp + 1) * ncols_per_level + col0, lo);
        tmem_ld_wait();
        const double w = ldexp(1.0, -8 * (2 * lp + 3));
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const long long both = ((long long)(int)hi[j] << 8) + (long long)(int)lo[j];
            v[j] = fma((double)both, w, v[j]);
        }
    }
}

}  // namespace i8
}  // namespace pk
