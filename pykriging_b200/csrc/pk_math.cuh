// Table-driven FP64 transcendentals shared by the Psi and predictor kernels: log2 split, 2^x, e^-S.
// They replace np.power / np.exp of matrixops.py:28,70 (glibc, < 1 ulp) with < 2 ulp (pow) and < 1.5 ulp
// (exp) implementations whose hot part is branch free (tests/test_gpu_parity.py::test_device_pow_and_exp_accuracy).
#pragma once
#include "pk_common.cuh"
#include "pk_exp2_table.h"
#include "pk_exp2_4096.h"

namespace pk {

// 2^(j/64) as hi + lo (mpmath-generated, tools/gen_exp2_table.py) and the log2(1+z) series
static __constant__ double c_exp2_hi[64] = {PK_EXP2_HI_LIST};
static __constant__ double c_exp2_lo[64] = {PK_EXP2_LO_LIST};
static __constant__ double c_log2p1[9] = {PK_LOG2P1_LIST};
static __constant__ double c_l2c_rc[64] = {PK_LOG2C_RC_LIST};
static __constant__ double c_l2c_hi[64] = {PK_LOG2C_HI_LIST};
static __constant__ double c_l2c_lo[64] = {PK_LOG2C_LO_LIST};

// ---- table-driven 2^x and e^-S ----------------------------------------------------------------------
// tab = shared copy of g_exp2_table replicated over the 16 eight-byte banks: tab[j * 16 + (lane & 15)]
// is conflict free for any j pattern.
constexpr double MAGIC64 = 6755399441055744.0 / 64.0;  // 1.5 * 2^52 / 64: rounds to multiples of 1/64
constexpr double MAGIC1 = 6755399441055744.0;          // 1.5 * 2^52

__device__ __forceinline__ double scale_pow2(double v, int n) {
    // v in [1, 2): returns v * 2^n with flush-to-zero below the normal range and inf above it
    if (n < -1021) return 0.0;
    if (n > 1022) return CUDART_INF;
    const long long bits = __double_as_longlong(v) + ((long long)n << 52);
    return __longlong_as_double(bits);
}

// 2^(s + e) where s is a double and e a tiny correction (|e| << 1/64)
__device__ __forceinline__ double exp2_table(double s, double e, const double* __restrict__ tab, int bank) {
    const double kd = s + MAGIC64;
    const int ki = __double2loint(kd);          // round(s * 64) in the low mantissa bits
    const double kf = kd - MAGIC64;             // round(s * 64) / 64
    const double r = (s - kf) + e;              // |r| <= 1/128 (+ e)
    // 2^r - 1 = r (c1 + r (c2 + r (c3 + r (c4 + r c5)))), |r| <= 2^-7: truncation < 2^-56
    double q = fma(r, 0x1.5d87fe78a6731p-10, 0x1.3b2ab6fba4e77p-7);  // ln2^5/120, ln2^4/24
    q = fma(r, q, 0x1.c6b08d704a0c0p-5);                              // ln2^3/6
    q = fma(r, q, 0x1.ebfbdff82c58fp-3);                              // ln2^2/2
    q = fma(r, q, 0x1.62e42fefa39efp-1);                              // ln2
    q = q * r;
    const double T = tab[(ki & 63) * 16 + bank];
    return scale_pow2(fma(T, q, T), ki >> 6);
}

// e^(-S), S >= 0 (also fine for negative S), Cody-Waite reduction with ln2/64 split in two
__device__ __forceinline__ double expneg_table(double S, const double* __restrict__ tab, int bank) {
    const double t = S * -0x1.71547652b82fep+6;  // -64 / ln 2
    const double kd = t + MAGIC1;
    const int ki = __double2loint(kd);
    const double kf = kd - MAGIC1;              // k = round(-S * 64 / ln2)
    // r = -S - k * ln2/64  (hi part has 21 trailing zero bits: k * hi is exact for |k| < 2^21)
    double r = fma(kf, -0x1.62e42fee00000p-7, -S);   // fdlibm ln2HI / 64
    r = fma(kf, -0x1.a39ef35793c76p-39, r);          // fdlibm ln2LO / 64
    // e^r - 1 = r + r^2/2 + r^3/6 + r^4/24 + r^5/120, |r| <= ln2/128
    double q = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    q = fma(r, q, 1.6666666666666666e-01);
    q = fma(r, q, 0.5);
    q = fma(r, q, 1.0);
    q = q * r;
    const double T = tab[(ki & 63) * 16 + bank];
    if (!(S < 1.0e6)) return (S != S) ? S : 0.0;  // NaN propagates, huge S underflows
    return scale_pow2(fma(T, q, T), ki >> 6);
}

// log2 d as an unevaluated sum L + res with L = fl(log2 d) and |res| <= ulp(L)/2, accurate to ~2^-59
// ABSOLUTE (not relative to L): d = 2^E m, m in [1,2); the top six mantissa bits j select the centre
// c_j = 1 + (2j+1)/128 of m's 1/64-wide interval, z = (m - c_j) / c_j (m - c_j is exact, 1/c_j is tabulated),
// log2 d = (E + hi_j) + (lo_j + log2(1 + z)) with log2 c_j = hi_j + lo_j tabulated (hi_j on a 2^-40 grid, so
// E + hi_j is exact) and a degree-9 series for log2(1 + z), |z| <= 2^-7.  No division, no libm call.
// Runs once per (pair, dimension) per CTA, never in the candidate loop.
__device__ __forceinline__ void log2_split(double d, double& L, float& res) {
    if (!(d >= 2.2250738585072014e-308)) {  // pow(0, p > 0) = 0: a hugely negative exponent underflows to 0
        L = -1100.0;
        res = 0.0f;
        return;
    }
    const int hi = __double2hiint(d);
    const int E = ((hi >> 20) & 0x7ff) - 1023;
    const int j = (hi >> 14) & 63;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(d));
    const double c = fma((double)(2 * j + 1), 0.0078125, 1.0);  // exact
    const double z = (m - c) * c_l2c_rc[j];
    double q = c_log2p1[8];
#pragma unroll
    for (int i = 7; i >= 0; --i) q = fma(q, z, c_log2p1[i]);
    const double t = fma(q, z, c_l2c_lo[j]);     // lo_j + log2(1 + z)
    const double lh = (double)E + c_l2c_hi[j];   // exact
    L = lh + t;
    const double bb = L - lh;                    // TwoSum: (lh + t) = L + r exactly
    const double r = (lh - (L - bb)) + (t - bb);
    res = (float)r;
}

// ---- 4096-entry variants for the population Psi kernel ----------------------------------------------------
// That kernel is ISSUE bound: an FP64 instruction holds a scheduler's dispatch port for two cycles, every
// other instruction for one -- 2 x 151 + 140 = 442 cycles per warp and Psi entry with the 64-entry functions
// below is exactly what ncu measured (profiles/ncu_psi_pop_r01_r.txt: sm__cycles_active / entries per
// scheduler), so what counts is the instruction count.  With 2^(j/4096) tabulated the reduced argument is
// |r| <= 2^-13 and (2^r - 1)/r needs a quadratic (c3 r^4 < 0.01 ulp): 8 FP64 instructions per 2^x instead of
// 10; the exponent clamp (two integer min/max) is gone because log2_split_clamped keeps |p log2 d| < 1022 for
// p <= 3; table index, address and scaling cost five integer instructions instead of seven.  The table is a
// plain 32 KB shared array: the loads of a half-warp collide on its banks, but the LSU has the wavefronts to
// spare.  Measured on the C3 population: 6.53 ms -> 5.74 ms; a bank-replicated 256-entry table with a cubic
// (conflict free, one FP64 instruction more) measured 5.93 ms, two candidates per thread and iteration
// (124 registers, two CTAs per SM) 5.90 ms.
static __device__ const double g_exp2_4096[4096] = {PK_EXP2_4096_LIST};
constexpr double MAGIC32K = 6755399441055744.0 / 32768.0;  // 1.5 * 2^37: ulp 2^-15
constexpr double HALF_STEP_4096 = 1.0 / 8192.0;
constexpr double MAGIC4K = 6755399441055744.0 / 4096.0;    // 1.5 * 2^40: ulp 2^-12
constexpr double L2_CLAMP = 340.0;
static __constant__ double c_e2q[3] = {0x1.62e42fefa39efp-1, 0x1.ebfbdff82c58fp-3, 0x1.c6b08d704a0c0p-5};

__device__ __forceinline__ void fill_exp2_4096(double* tab, int tid, int nthreads) {
    for (int i = tid; i < 4096; i += nthreads) tab[i] = g_exp2_4096[i];
}

// 2^(pd * (lh + ll)) for |pd * lh| < 1022.  `tab_saddr` = 32-bit shared address of the table, ALIGNED TO
// 32 KB.  x = pd * lh (+ half a table step) is rounded to a multiple of 2^-15 by the magic add, so the low word
// kl of kd holds 32768 (x + 1/8192): bits 3..14 are the table index -- as a byte offset they are OR-ed into
// the address by the same LOP3 that masks them -- bits 15.. are the binary exponent, and kd with the low
// three bits cleared is x rounded to the nearest 1/4096: kf.
// `magic` = MAGIC4K; the population Psi kernel passes it in a REGISTER (made opaque to the compiler once per thread): as an
// immediate it takes the one special-operand slot of the DFMA, and a pd that arrives in a uniform register (constant-memory
// hyper-parameters, pk_psi.cu) would be copied into vector registers first -- two moves per call.
__device__ __forceinline__ double exp2_lean4k(double pd, double lh, double ll, unsigned tab_saddr, double magic = MAGIC4K) {
#ifdef PK_EXP2_4K_MASKED
    // first version: x rounded to 2^-15, low three bits masked off (one LOP3 + one register move more per call)
    const double kd = fma(pd, lh, MAGIC32K + HALF_STEP_4096);
    const int kl = __double2loint(kd);
    const double kf = __longlong_as_double(__double_as_longlong(kd) & ~0x7LL) - MAGIC32K;
    double r = fma(pd, lh, -kf);
    r = fma(pd, ll, r);                          // |r| <= 2^-13 + 2^-16
    double q = fma(r, c_e2q[2], c_e2q[1]);
    q = fma(r, q, c_e2q[0]);
    double T;
    asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tab_saddr | (unsigned)(kl & 0x7ff8)));
    const double v = fma(T * r, q, T);
    int hi;
    asm("mad.lo.s32 %0, %1, 1048576, %2;" : "=r"(hi) : "r"(kl >> 15), "r"(__double2hiint(v)));
    return __hiloint2double(hi, __double2loint(v));
#else
    // x = pd * lh rounded to the nearest multiple of 2^-12 by the magic add: the low word of kd is 4096 x in two's
    // complement -- bits 0..11 the table index, bits 12.. the binary exponent -- and kd - MAGIC4K is the rounded x
    // itself.  8 FP64 + 4 integer instructions + 1 shared load.
    const double kd = fma(pd, lh, magic);
    const int k8 = __double2loint(kd) << 3;      // index as a byte offset in bits 3..14, exponent in bits 15..
    const double kf = kd - magic;
    double r = fma(pd, lh, -kf);
    r = fma(pd, ll, r);                          // |r| <= 2^-13
    double q = fma(r, c_e2q[2], c_e2q[1]);
    q = fma(r, q, c_e2q[0]);
    double T;
    // base + offset (the base is 32 KB aligned, so this is the OR): the uniform base rides in the load's address mode
    // (LDS [R + UR]) instead of being copied into a vector register for a LOP3 per call
    asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tab_saddr + (unsigned)(k8 & 0x7ff8)));
    const double v = fma(T * r, q, T);
    int hi;
    asm("mad.lo.s32 %0, %1, 1048576, %2;" : "=r"(hi) : "r"(k8 >> 15), "r"(__double2hiint(v)));
    return __hiloint2double(hi, __double2loint(v));
#endif
}

// e^(-S), S >= 0: Cody-Waite with ln2/4096 = hi (20 bits: k * hi is exact for |k| < 2^33) + lo, quadratic.
__device__ __forceinline__ double expneg_4k(double S, unsigned tab_saddr) {
    const double kd = fma(S, -PK_4096_OVER_LN2, MAGIC1);
    const int ki = __double2loint(kd);
    const double kf = kd - MAGIC1;              // k = round(-S * 4096 / ln2)
    double r = fma(kf, -PK_LN2_4096_HI, -S);
    r = fma(kf, -PK_LN2_4096_LO, r);            // |r| <= ln2/8192
    double q = fma(r, 1.6666666666666666e-01, 0.5);
    q = fma(r, q, 1.0);                         // (e^r - 1) / r up to r^2/6
    double T;
    asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tab_saddr + (unsigned)((ki & 4095) << 3)));
    const double v = fma(T * r, q, T);
    if (!(S < 1.0e5)) return (S != S) ? S : 0.0;  // NaN propagates, huge S underflows
    return scale_pow2(v, ki >> 12);
}

// 2^(s + e) with full range checks and the 64-entry table read from the constant bank: the rare slow path
// (candidates with an exponent that is exactly 1 or 2, or outside [0.25, 3])
__device__ __forceinline__ double exp2_const(double s, double e) {
    if (!(s > -1100.0)) return 0.0;
    if (!(s < 1100.0)) return CUDART_INF;
    const double kd = s + MAGIC64;
    const int ki = __double2loint(kd);
    const double kf = kd - MAGIC64;
    const double r = (s - kf) + e;
    double q = fma(r, 0x1.5d87fe78a6731p-10, 0x1.3b2ab6fba4e77p-7);
    q = fma(r, q, 0x1.c6b08d704a0c0p-5);
    q = fma(r, q, 0x1.ebfbdff82c58fp-3);
    q = fma(r, q, 0x1.62e42fefa39efp-1);
    q = q * r;
    const double T = c_exp2_hi[ki & 63];
    return scale_pow2(fma(T, q, T), ki >> 6);
}

// log2_split with the result clamped to [-L2_CLAMP, L2_CLAMP] (d = 0 -> -L2_CLAMP): for p <= 3 the product
// p * log2 d then stays inside the exponent range and exp2_lean4k needs no clamp of its own; 2^(-340 p) is
// < 1e-102 -- in a sum that is >= 0 and feeds e^-S it is indistinguishable from pow(0, p) = 0.
__device__ __forceinline__ void log2_split_clamped(double d, double& L, float& res) {
    log2_split(d, L, res);
    if (!(L > -L2_CLAMP)) { L = -L2_CLAMP; res = 0.0f; }
    if (L > L2_CLAMP) { L = L2_CLAMP; res = 0.0f; }
}

// ---- lean 2^x for the candidate loop ---------------------------------------------------------------
// Same arithmetic as exp2_table, written so that the hot loop carries no branches, no 64-bit immediates
// (the polynomial lives in the constant bank and is read as a DFMA operand) and no address
// re-computation: `tabp` is the byte address of this lane's bank copy of the 2^(j/64) table.  The
// result is scaled by adding the clamped binary exponent to the high word, so anything below
// 2^-1022 comes out as ~1e-308 instead of 0 (it is then multiplied by theta <= 100 and added to a
// sum that is >= 0: indistinguishable from pow(0, p) = 0 in every Psi entry) .
static __constant__ double c_e2[5] = {0x1.62e42fefa39efp-1, 0x1.ebfbdff82c58fp-3, 0x1.c6b08d704a0c0p-5,
                               0x1.3b2ab6fba4e77p-7, 0x1.5d87fe78a6731p-10};

// 2^(pd * (lh + ll)): the product never appears as a rounded double -- the reduction constant is added
// inside the FMA (kd), and the reduced argument r = pd*lh - kf + pd*ll is formed by two FMAs, each exact
// product rounded once at magnitude <= 2^-7 (absolute error < 2^-60).  10 FP64 instructions.
__device__ __forceinline__ double exp2_lean(double pd, double lh, double ll, const char* __restrict__ tabp) {
    const double kd = fma(pd, lh, MAGIC64);
    const int ki = __double2loint(kd);          // round(64 * pd * lh) in the low mantissa bits
    const double kf = kd - MAGIC64;
    double r = fma(pd, lh, -kf);
    r = fma(pd, ll, r);
    double q = fma(r, c_e2[4], c_e2[3]);
    q = fma(r, q, c_e2[2]);
    q = fma(r, q, c_e2[1]);
    q = fma(r, q, c_e2[0]);
    const double T = *reinterpret_cast<const double*>(tabp + ((ki & 63) << 7));
    const double v = fma(T * r, q, T);
    const int n = min(max(ki >> 6, -1022), 1023);
    return __hiloint2double(__double2hiint(v) + (n << 20), __double2loint(v));
}

// log2_split with the three 64-entry tables in a bank-replicated shared-memory copy (ltab = 3 x 64 x 16
// doubles filled by fill_log2_tables; bank = lane & 15): for kernels that take a log2 per (query, sample,
// dimension) -- a constant-bank lookup with a per-lane index serialises, this one does not.
__device__ __forceinline__ void log2_split_smem(double d, const double* __restrict__ ltab, int bank, double& L,
                                                float& res) {
    const bool zero = !(d >= 2.2250738585072014e-308);
    const int hi = __double2hiint(d);
    const int E = ((hi >> 20) & 0x7ff) - 1023;
    const int j = (hi >> 14) & 63;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(d));
    const double c = fma((double)(2 * j + 1), 0.0078125, 1.0);  // exact
    const double* e = ltab + j * 16 + bank;
    const double z = (m - c) * e[0];
    double q = c_log2p1[8];
#pragma unroll
    for (int i = 7; i >= 0; --i) q = fma(q, z, c_log2p1[i]);
    const double t = fma(q, z, e[2 * 1024]);
    const double lh = (double)E + e[1024];
    const double Ls = lh + t;
    const double bb = Ls - lh;
    const double r = (lh - (Ls - bb)) + (t - bb);
    L = zero ? -1100.0 : Ls;   // pow(0, p > 0) = 0: a hugely negative exponent underflows to 0
    res = zero ? 0.0f : (float)r;
}

__device__ __forceinline__ void fill_log2_tables(double* ltab, int tid, int nthreads) {
    for (int i = tid; i < 64 * 16; i += nthreads) {
        ltab[i] = c_l2c_rc[i >> 4];
        ltab[1024 + i] = c_l2c_hi[i >> 4];
        ltab[2048 + i] = c_l2c_lo[i >> 4];
    }
}

// fill the bank-replicated shared copy of the 2^(j/64) table (64 x 16 doubles); caller synchronises
__device__ __forceinline__ void fill_exp2_table(double* tab, int tid, int nthreads) {
    for (int i = tid; i < 64 * 16; i += nthreads) tab[i] = c_exp2_hi[i >> 4];
}

// |d|^p with the reference's exact special cases (pk_common.cuh pow_abs) and the table-driven general path
__device__ __forceinline__ double pow_abs_lean(double d, double p, const char* tabp, const double* ltab, int bank) {
    if (p == 2.0) return d * d;
    if (p == 1.0) return d;
    double lh;
    float ll;
    log2_split_smem(d, ltab, bank, lh, ll);
    return exp2_lean(p, lh, (double)ll, tabp);
}

}  // namespace pk
