// Batched Psi construction for a whole GA/PSO population (the K1 kernel).
//
// Replaces the Psi part of matrixops.updatePsi / regupdatePsi (matrixops.py:28-30, 38-40):
//     Psi_ij = exp(-sum_d theta_d |x_id - x_jd|^p_d)   (i != j),   Psi_ii = 1 + eps | 1 + lambda
// for C candidates at once, written as the lower-triangle 64x64 tiles the blocked Cholesky consumes.
//
// Design (FP64-ALU bound, not HBM bound: k pow + 1 exp per 8 bytes written):
//   * a CTA owns a (RB rows x 64 cols) block of sample PAIRS and loops over a group of candidates.
//     The candidate-independent part of |d|^p = 2^(p log2 d) -- log2 d as an unevaluated sum
//     lh + ll -- is computed ONCE per CTA and kept in registers, so the per-candidate cost of a
//     (pair, dimension) is one table-driven 2^x (about 15 FP64 instructions) instead of a full pow().
//   * sample coordinates of the block and the (theta, p) of the candidate group are staged in shared
//     memory; every thread writes PPT consecutive doubles per candidate (coalesced 16/32-byte stores,
//     a warp covers whole 128-byte lines).
//   * p == 2 and p == 1 (what a bounded PSO clips to, krige.py:383) take the exact d*d / d path.
//   * the sum over dimensions follows NumPy's reduction order (pk_common.cuh), the final exp(-S) is a
//     Cody-Waite reduced table-driven exponential (< 1 ulp).
#include "pk_exp2_table.h"
#include "pk_internal.h"

namespace pk {

// 2^(j/64) as hi + lo (mpmath-generated, tools/gen_exp2_table.py) and the log2(1+z) series
__constant__ double c_exp2_hi[64] = {PK_EXP2_HI_LIST};
__constant__ double c_exp2_lo[64] = {PK_EXP2_LO_LIST};
__constant__ double c_log2p1[9] = {PK_LOG2P1_LIST};
__constant__ double c_l2c_rc[64] = {PK_LOG2C_RC_LIST};
__constant__ double c_l2c_hi[64] = {PK_LOG2C_HI_LIST};
__constant__ double c_l2c_lo[64] = {PK_LOG2C_LO_LIST};

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

// ---- lean 2^x for the candidate loop ---------------------------------------------------------------
// Same arithmetic as exp2_table, written so that the hot loop carries no branches, no 64-bit immediates
// (the polynomial lives in the constant bank and is read as a DFMA operand) and no address
// re-computation: `tabp` is the byte address of this lane's bank copy of the 2^(j/64) table.  The
// result is scaled by adding the clamped binary exponent to the high word, so anything below
// 2^-1022 comes out as ~1e-308 instead of 0 (it is then multiplied by theta <= 100 and added to a
// sum that is >= 0: indistinguishable from pow(0, p) = 0 in every Psi entry) .
__constant__ double c_e2[5] = {0x1.62e42fefa39efp-1, 0x1.ebfbdff82c58fp-3, 0x1.c6b08d704a0c0p-5,
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

constexpr int PSI_THREADS = 256;
constexpr int PSI_GROUP = 64;  // candidates per CTA (amortises the log2 split of the CTA's pairs)

// PPT pairs per thread: large K keeps ONE pair per thread so that the per-pair log2 tables (3 registers per
// dimension) leave room for three resident CTAs per SM -- the per-candidate 2^x chain is latency bound and
// needs the extra warps more than it needs wider stores.
template <int K, int PPT>
__global__ void __launch_bounds__(PSI_THREADS, (PPT == 1) ? (K <= 12 ? 3 : 2) : 1)
psi_pop_kernel(int n, int np, int ld, size_t elems, int C, const double* __restrict__ X,
               const double* __restrict__ theta, const double* __restrict__ p,
               const double* __restrict__ lambda, int mode, double* __restrict__ ws) {
    constexpr int RB = PSI_THREADS * PPT / PK_TILE;  // rows of the pair block
    constexpr int SPT = PK_TILE / RB;                // row blocks per tile
    __shared__ double tab[64 * 16];
    __shared__ double Xr[RB * K], Xc[PK_TILE * K];
    __shared__ double2 thp[PSI_GROUP * K];  // (theta_d, p_d) pairs: one 16-byte broadcast load per (candidate, dimension)
    __shared__ double dg[PSI_GROUP];
    __shared__ int spec[PSI_GROUP];  // candidate has an exponent that is exactly 1 or 2 (exact d / d*d path)
    const int tid = threadIdx.x, bank = tid & 15;
    // block -> (lower tile, row block)
    const int t = blockIdx.x / SPT, sb = blockIdx.x % SPT;
    int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    const int tj = t - ti * (ti + 1) / 2;
    const int r0 = ti * PK_TILE + sb * RB, c0 = tj * PK_TILE;
    const int g0 = blockIdx.y * PSI_GROUP;
    const int gn = min(PSI_GROUP, C - g0);

    for (int i = tid; i < 64 * 16; i += PSI_THREADS) tab[i] = c_exp2_hi[i >> 4];
    for (int i = tid; i < RB * K; i += PSI_THREADS) {
        const int rr = i / K;
        Xr[i] = (r0 + rr < n) ? X[(size_t)(r0 + rr) * K + (i % K)] : 0.0;
    }
    for (int i = tid; i < PK_TILE * K; i += PSI_THREADS) {
        const int rr = i / K;
        Xc[i] = (c0 + rr < n) ? X[(size_t)(c0 + rr) * K + (i % K)] : 0.0;
    }
    for (int i = tid; i < gn * K; i += PSI_THREADS) {
        thp[i] = make_double2(theta[(size_t)g0 * K + i], p[(size_t)g0 * K + i]);
    }
    if (tid < gn) {
        dg[tid] = 1.0 + (mode == PK_MODE_REGRESSION ? lambda[g0 + tid] : PK_EPS_DIAG);
        int sp = 0;
        for (int d = 0; d < K; ++d) {
            const double pd = p[(size_t)(g0 + tid) * K + d];
            sp |= (pd == 1.0 || pd == 2.0) ? 1 : 0;
        }
        spec[tid] = sp;
    }
    __syncthreads();
    const char* tabp = reinterpret_cast<const char*>(tab + bank);

    const int lr = tid / (PK_TILE / PPT);            // local row
    const int lc = (tid % (PK_TILE / PPT)) * PPT;    // first local column
    const int gr = r0 + lr;
    double lh[PPT][K];
    float ll[PPT][K];
    bool live[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        const int gc = c0 + lc + q;
        live[q] = (gr < n) && (gc < gr);
#pragma unroll
        for (int d = 0; d < K; ++d) {
            if (live[q]) log2_split(fabs(Xr[lr * K + d] - Xc[(lc + q) * K + d]), lh[q][d], ll[q][d]);
            else { lh[q][d] = 0.0; ll[q][d] = 0.0f; }
        }
    }
    bool any_live = false;
#pragma unroll
    for (int q = 0; q < PPT; ++q) any_live |= live[q];

    for (int g = 0; g < gn; ++g) {
        double out[PPT];
        const double2* tpg = thp + g * K;
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const int gc = c0 + lc + q;
            double v;
            if (live[q] && !spec[g]) {
                // generic exponents: k table-driven 2^x, theta folded in with an FMA, NumPy's summation tree
                // term(d) = theta_d * |d|^p_d
                auto tw = [&](int d, double& th_d) {
                    const double2 tp = tpg[d];
                    th_d = tp.x;
                    return exp2_lean(tp.y, lh[q][d], (double)ll[q][d], tabp);
                };
                double S;
                if (K < 8) {
                    S = 0.0;
#pragma unroll
                    for (int d = 0; d < K; ++d) {
                        double t;
                        const double wv = tw(d, t);
                        S = fma(t, wv, S);
                    }
                } else {
                    double r[8];
#pragma unroll
                    for (int d = 0; d < 8; ++d) {
                        double t;
                        const double wv = tw(d, t);
                        r[d] = t * wv;
                    }
#pragma unroll
                    for (int d = 8; d < K - (K % 8); ++d) {
                        double t;
                        const double wv = tw(d, t);
                        r[d & 7] = fma(t, wv, r[d & 7]);
                    }
                    S = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
#pragma unroll
                    for (int d = K - (K % 8); d < K; ++d) {
                        double t;
                        const double wv = tw(d, t);
                        S = fma(t, wv, S);
                    }
                }
                v = expneg_table(S, tab, bank);
            } else if (live[q]) {
                // a candidate with an exponent that is exactly 1 or 2 (bounder-clipped PSO particles,
                // krige.py:383): np.power returns d and d*d exactly there, so do the same
                double term[K];
#pragma unroll
                for (int d = 0; d < K; ++d) {
                    const double pd = tpg[d].y;
                    double w;
                    if (pd == 2.0) {
                        const double dd = fabs(Xr[lr * K + d] - Xc[(lc + q) * K + d]);
                        w = dd * dd;
                    } else if (pd == 1.0) {
                        w = fabs(Xr[lr * K + d] - Xc[(lc + q) * K + d]);
                    } else {
                        const double s = pd * lh[q][d];
                        const double e = fma(pd, (double)ll[q][d], fma(pd, lh[q][d], -s));
                        w = exp2_table(s, e, tab, bank);
                    }
                    term[d] = tpg[d].x * w;
                }
                const double S = sum_numpy_order(K, [&](int d) { return term[d]; });
                v = expneg_table(S, tab, bank);
            } else if (gr >= n || gc >= n) {
                v = (gr == gc) ? 1.0 : 0.0;   // identity padding
            } else {
                v = (gr == gc) ? dg[g] : 0.0;  // diagonal / strictly upper part of a diagonal tile
            }
            out[q] = v;
        }
        double* dst = ws + (size_t)(g0 + g) * elems + (size_t)gr * ld + c0 + lc;
        if (PPT == 4) {
            reinterpret_cast<double2*>(dst)[0] = make_double2(out[0], out[1]);
            reinterpret_cast<double2*>(dst)[1] = make_double2(out[2 % PPT], out[3 % PPT]);
        } else if (PPT == 2) {
            reinterpret_cast<double2*>(dst)[0] = make_double2(out[0], out[1 % PPT]);
        } else {
            dst[0] = out[0];  // a warp still writes 256 contiguous bytes
        }
    }
    (void)any_live;
}

// diagnostic: the device pow / exp building blocks on arrays (tests measure their ulp error)
__global__ void debug_math_kernel(int64_t m, const double* __restrict__ d, const double* __restrict__ p,
                                  double* __restrict__ out_pow, double* __restrict__ out_expneg) {
    __shared__ double tab[64 * 16];
    for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) tab[i] = c_exp2_hi[i >> 4];
    __syncthreads();
    const int bank = threadIdx.x & 15;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        double lh;
        float ll;
        log2_split(fabs(d[i]), lh, ll);
        const double s = p[i] * lh;
        const double e = fma(p[i], (double)ll, fma(p[i], lh, -s));
        // odd elements go through the lean (hot-loop) variant, even ones through the range-checked one
        out_pow[i] = (i & 1) ? exp2_lean(p[i], lh, (double)ll, reinterpret_cast<const char*>(tab + bank))
                             : exp2_table(s, e, tab, bank);
        out_expneg[i] = expneg_table(d[i], tab, bank);
    }
}

cudaError_t launch_debug_math(pk_handle* h, int64_t m, const double* d, const double* p, double* out_pow,
                              double* out_expneg) {
    cudaError_t e;
    debug_math_kernel<<<296, 256, 0, h->stream>>>(m, d, p, out_pow, out_expneg);
    h->launches++;
    return cudaGetLastError();
}

// generic fallback (any k <= PK_MAX_K): defined in pk_tiled.cu
cudaError_t launch_psi_build_generic(pk_handle* h, const TiledShape& s, int C, const double* theta, const double* p,
                                     const double* lambda, int mode, double* ws);
cudaError_t launch_aug_init(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info);

template <int K>
static cudaError_t launch_psi_pop(pk_handle* h, const TiledShape& s, int C, const double* theta, const double* p,
                                  const double* lambda, int mode, double* ws) {
    constexpr int PPT = (K > 6) ? 1 : 4;
    constexpr int RB = PSI_THREADS * PPT / PK_TILE;
    const int nt = s.np / PK_TILE;
    dim3 grid(nt * (nt + 1) / 2 * (PK_TILE / RB), (C + PSI_GROUP - 1) / PSI_GROUP);
    psi_pop_kernel<K, PPT><<<grid, PSI_THREADS, 0, h->stream>>>(s.n, s.np, s.ld, s.elems(), C, h->dX, theta, p,
                                                                 lambda, mode, ws);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t launch_psi_build(pk_handle* h, const TiledShape& s, int C, const double* theta, const double* p,
                             const double* lambda, int mode, double* ws, int32_t* info) {
    cudaError_t e = cudaSuccess;
    switch (h->k) {
#define PK_CASE(KK) case KK: e = launch_psi_pop<KK>(h, s, C, theta, p, lambda, mode, ws); break;
        PK_CASE(1) PK_CASE(2) PK_CASE(3) PK_CASE(4) PK_CASE(5) PK_CASE(6) PK_CASE(7) PK_CASE(8)
        PK_CASE(9) PK_CASE(10) PK_CASE(11) PK_CASE(12) PK_CASE(13) PK_CASE(14) PK_CASE(15) PK_CASE(16)
#undef PK_CASE
        default: e = launch_psi_build_generic(h, s, C, theta, p, lambda, mode, ws); break;
    }
    if (e != cudaSuccess) return e;
    return launch_aug_init(h, s, C, ws, info);
}

}  // namespace pk
