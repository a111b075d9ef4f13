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

}  // namespace i8
}  // namespace pk
