// Prediction path, split variant: two kernels per block of query points.
//
// Replaces the same reference code as pk_predict.cu (matrixops.predict_normalized / predicterr_normalized,
// matrixops.py:68-95, and kriging.expimp / weightedexpimp, krige.py:201-234), for LARGE query sets:
//
//   psi_rows_kernel   psi[q, r] = exp(-sum_d theta_d |X_rd - xq_d|^p_d)  for a block of queries, written row-major
//                     (q, np) into an HBM scratch, together with yhat_q = mu + psi_q . w.  Pure FP64-ALU kernel:
//                     a warp owns a query, its lanes walk the samples (coalesced 256-byte stores), two samples
//                     per lane in flight.
//   trinorm_kernel    |L^-1 psi_q|^2 for 64 queries x 256 columns of v at a time: a TMA-fed DMMA GEMM of the psi
//                     tile against the rows of the cached L^-1 (lower triangular: chunks right of the diagonal
//                     are neither loaded nor multiplied), followed by s = sqrt|sigma2 (1 + extra - |v|^2)| and EI.
//
// Why two kernels: in the fused predictors a warp alternates between a ~100-deep dependent FP64 chain (psi) and a
// burst of DMMAs; both run on the same FP64 pipe, the chain's instructions queue behind the other warps' 16-cycle
// DMMAs and the pipe idles 42 % of the time (profiles/ncu_predict_reg_r01_z.txt).  Apart, each kernel keeps
// the pipe busy on its own; the price -- psi through HBM, 2 x 8 np bytes per query -- is 136 GB for the 16 Mi-point
// grid of config C5, 21 ms of HBM time hidden under 200 ms of arithmetic.
#include <cuda.h>

#include "pk_i8.cuh"
#include "pk_internal.h"
#include "pk_math.cuh"

namespace pk {

// ---- kernel A: psi rows + yhat -----------------------------------------------------------------------
// ISSUE bound like psi_pop_kernel (an FP64 instruction holds a scheduler's dispatch port for two cycles, anything
// else for one: profiles/ncu_psi_rows_r01_s4.txt, 2 x 79 + 132 = 290 cycles per warp and psi value at k = 2), so
// what counts is the instruction count:
//   * log2_split_rows: no int -> double conversions (the exponent and the interval centre are assembled from
//     bit patterns), no float round trip of the low part, no special case for d = 0 (E = -1023 falls out of the
//     bit fields and the clamp at -340 does the rest);
//   * MEMOISED DIMENSIONS: a warp owns a contiguous run of query points and keeps theta_d |X_rd - xq_d|^p_d of the
//     leading dimensions of its previous query in shared memory.  When the next query has the same leading
//     coordinates -- every row-major grid (plot / snapshot / the 4096 x 4096 EI grid of config C5:
//     krige.py:531-537, 699-708) -- those terms are read back instead of recomputed: bit-identical by
//     construction, and one pow per (query, sample) instead of k.
constexpr int PA_THREADS = 768;
constexpr int PA_WARPS = PA_THREADS / 32;
constexpr int PA_XT_MAX = 12288;   // staged (transposed) sample coordinates: k * np doubles <= 96 KB
constexpr int PA_SMALL_MAX = 6 * 0x8000 - 0x400;  // bytes in front of the 2^x table that still leave room for it

// bytes in front of the 32 KB-aligned 2^x table: log2 tables + hyper-parameters + per-dimension mode + X^T + memo
__host__ __device__ inline unsigned pa_small_bytes(int xt_doubles, int memo_doubles) {
    return (unsigned)(sizeof(double) * (3 * 64 * 16 + 2 * PK_MAX_K + xt_doubles + memo_doubles) + sizeof(int) * PK_MAX_K);
}

// log2 d = lh + ll (unevaluated sum, same tables and series as log2_split), lh clamped to [-340, 340]
__device__ __forceinline__ void log2_split_rows(double d, const double* __restrict__ ltab, int bank, double& lh_out,
                                                double& ll_out) {
    const int hi = __double2hiint(d);
    const int j = (hi >> 14) & 63;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(d));
    const double c = __hiloint2double(0x3ff00000 | ((2 * j + 1) << 13), 0);  // 1 + (2j+1)/128
    const double* e = ltab + j * 16 + bank;
    const double z = (m - c) * e[0];
    double q = c_log2p1[8];
#pragma unroll
    for (int i = 7; i >= 0; --i) q = fma(q, z, c_log2p1[i]);
    const double t = fma(q, z, e[2 * 1024]);
    // biased exponent as a double: 2^52 + b has b in its low word
    const double Ed = __hiloint2double(0x43300000, (hi >> 20) & 0x7ff) - 4503599627371519.0;  // - (2^52 + 1023)
    const double lh = Ed + e[1024];              // exact
    const double Ls = lh + t;
    const double bb = Ls - lh;                   // TwoSum: (lh + t) = Ls + r exactly
    ll_out = (lh - (Ls - bb)) + (t - bb);
    lh_out = fmin(fmax(Ls, -L2_CLAMP), L2_CLAMP);
}

// The product is rounded on its own (__dmul_rn, as NumPy's theta * np.power(...) is) and never contracted into the
// sum over dimensions: the generic and the grid-mode kernel add the same rounded terms in the same order.
// theta_d |d|^p_d for one dimension.  mode: 0 = table-driven 2^(p log2 d) (p in [0.25, 3], not 1 or 2),
// 1 = p == 1 (exact d), 2 = p == 2 (exact d*d, as np.power), 3 = any other exponent (range-checked path)
__device__ __forceinline__ double pa_term_lean(double dd, double th, double p, const double* ltab, int bank,
                                               unsigned tab_saddr) {
    double lh, ll;
    log2_split_rows(dd, ltab, bank, lh, ll);
    return __dmul_rn(th, exp2_lean4k(p, lh, ll, tab_saddr));
}
__device__ __forceinline__ double pa_term(double dd, double th, double p, int mode, const double* ltab, int bank,
                                          unsigned tab_saddr) {
    if (mode == 0) return pa_term_lean(dd, th, p, ltab, bank, tab_saddr);
    if (mode == 2) return __dmul_rn(th, dd * dd);
    if (mode == 1) return __dmul_rn(th, dd);
    double l0;
    float l1;
    log2_split(dd, l0, l1);
    const double s = p * l0;
    const double e = fma(p, (double)l1, fma(p, l0, -s));
    return __dmul_rn(th, exp2_const(s, e));
}

template <int KK>
__global__ void __launch_bounds__(PA_THREADS, 1)
psi_rows_kernel(int n, int np, int k, int64_t mc, const double* __restrict__ X, const double* __restrict__ Xq,
                const double* __restrict__ hp, const double* __restrict__ w, double mu, int xt_staged, int memo_dims,
                double* __restrict__ psi_out, double* __restrict__ yhat_out, const int* __restrict__ gridinfo,
                uint8_t* __restrict__ planes, size_t plane_stride) {
    if (gridinfo && gridinfo[1]) return;  // the query set is a row-major grid: psi_rows_grid_kernel does this block
    extern __shared__ __align__(16) double pa_smem[];
    double* ltab = pa_smem;                                   // 3 x 64 x 16
    double* th_s = ltab + 3 * 64 * 16;                        // PK_MAX_K
    double* p_s = th_s + PK_MAX_K;                            // PK_MAX_K
    int* mode_s = reinterpret_cast<int*>(p_s + PK_MAX_K);     // PK_MAX_K
    double* Xt = reinterpret_cast<double*>(mode_s + PK_MAX_K);  // k x np (transposed: conflict-free per-lane reads)
    double* memo = Xt + (xt_staged ? k * np : 0);             // PA_WARPS x memo_dims x np
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(pa_smem);
    const unsigned tab_saddr =
        (smem_base + pa_small_bytes(xt_staged ? k * np : 0, PA_WARPS * memo_dims * np) + 0x7fffu) & ~0x7fffu;
    double* tab4k = pa_smem + (tab_saddr - smem_base) / sizeof(double);
    fill_exp2_4096(tab4k, tid, PA_THREADS);
    fill_log2_tables(ltab, tid, PA_THREADS);
    if (tid < k) {
        const double pd = hp[PK_MAX_K + tid];
        th_s[tid] = hp[tid];
        p_s[tid] = pd;
        mode_s[tid] = (pd == 1.0) ? 1 : (pd == 2.0) ? 2 : (pd >= 0.25 && pd <= 3.0) ? 0 : 3;
    }
    if (xt_staged) {
        for (int i = tid; i < k * np; i += PA_THREADS) {
            const int d = i / np, r = i - d * np;
            Xt[i] = (r < n) ? X[(size_t)r * k + d] : 0.0;
        }
    }
    __syncthreads();
    const int bank = tid & 15;
    double th_r[KK > 0 ? KK : 1], p_r[KK > 0 ? KK : 1];
    if (KK > 0) {
#pragma unroll
        for (int d = 0; d < KK; ++d) {
            th_r[d] = th_s[d];
            p_r[d] = p_s[d];
        }
    }
    bool lean = true;  // every exponent takes the table-driven path: straight-line code, two chains interleaved
    for (int d = 0; d < k; ++d) lean &= (mode_s[d] == 0);
    const int nit = np / 32;  // even: np is a multiple of 64
    // this warp's contiguous run of queries
    const int64_t nwarps = (int64_t)gridDim.x * PA_WARPS;
    const int64_t per = (mc + nwarps - 1) / nwarps;
    const int64_t qa = ((int64_t)blockIdx.x * PA_WARPS + warp) * per;
    const int64_t qb = (qa + per < mc) ? qa + per : mc;
    double* mw = memo + (size_t)warp * memo_dims * np;
    double mx[KK > 0 ? KK : 1];  // leading coordinates the memo was built for
#pragma unroll
    for (int d = 0; d < (KK > 0 ? KK : 1); ++d) mx[d] = CUDART_NAN;
    for (int64_t q = qa; q < qb; ++q) {
        double xq_r[KK > 0 ? KK : 1];
        bool hit = (KK > 0) && lean && memo_dims > 0;
        if constexpr (KK > 0) {
#pragma unroll
            for (int d = 0; d < KK; ++d) {
                xq_r[d] = __ldg(Xq + q * KK + d);
                if (d < memo_dims) hit &= (xq_r[d] == mx[d]);
            }
        }
        const bool use_memo = (KK > 0) && lean && memo_dims > 0;
        const double* xqg = Xq + q * k;
        // a NaN coordinate must poison the whole psi row (np.power / np.exp propagate it, matrixops.py:70): the
        // table-driven pow treats a distance that is not >= tiny like 0, so the query is flagged here (warp uniform)
        bool qnan = false;
        for (int d = 0; d < k; ++d) {
            const double c = __ldg(xqg + d);
            qnan |= (c != c);
        }
        auto S_any = [&](int r) {
            return sum_numpy_order(k, [&](int d) {
                const double xr = xt_staged ? Xt[d * np + r] : __ldg(X + (size_t)r * k + d);
                return pa_term(fabs(xr - __ldg(xqg + d)), th_s[d], p_s[d], lean ? 0 : mode_s[d], ltab, bank, tab_saddr);
            });
        };
        double ys0 = 0.0, ys1 = 0.0;
        double* prow = psi_out ? psi_out + (size_t)q * np : nullptr;
        for (int it = 0; it < nit; it += 2) {
            const int r0 = it * 32 + lane, r1 = r0 + 32;
            const int c0 = min(r0, n - 1), c1 = min(r1, n - 1);  // padding rows: computed on a valid sample, then zeroed
            double S0, S1;
            if constexpr (KK > 0) {
                if (lean) {
                    // S(r) = sum_d theta_d |X_rd - xq_d|^p_d in NumPy's summation order (plain left to right, KK < 8)
                    double t0[KK], t1[KK];
                    auto term = [&](int d, int r) {
                        const double xr = xt_staged ? Xt[d * np + r] : __ldg(X + (size_t)r * KK + d);
                        return pa_term_lean(fabs(xr - xq_r[d]), th_r[d], p_r[d], ltab, bank, tab_saddr);
                    };
                    if (hit) {
#pragma unroll
                        for (int d = 0; d < KK; ++d) {
                            if (d < memo_dims) {
                                t0[d] = mw[d * np + r0];
                                t1[d] = mw[d * np + r1];
                            } else {
                                t0[d] = term(d, c0);
                                t1[d] = term(d, c1);
                            }
                        }
                    } else {
#pragma unroll
                        for (int d = 0; d < KK; ++d) {
                            t0[d] = term(d, c0);
                            t1[d] = term(d, c1);
                            if (use_memo && d < memo_dims) {
                                mw[d * np + r0] = t0[d];
                                mw[d * np + r1] = t1[d];
                            }
                        }
                    }
                    S0 = sum_numpy_order(KK, [&](int d) { return t0[d]; });
                    S1 = sum_numpy_order(KK, [&](int d) { return t1[d]; });
                } else {
                    S0 = S_any(c0);
                    S1 = S_any(c1);
                }
            } else {
                S0 = S_any(c0);
                S1 = S_any(c1);
            }
            double v0 = expneg_4k(S0, tab_saddr), v1 = expneg_4k(S1, tab_saddr);
            v0 = (r0 < n) ? v0 : 0.0;
            v1 = (r1 < n) ? v1 : 0.0;
            if (qnan) v0 = v1 = CUDART_NAN;
            ys0 = fma(v0, __ldg(w + c0), ys0);
            ys1 = fma(v1, __ldg(w + c1), ys1);
            if (prow) {
                prow[r0] = v0;
                prow[r1] = v1;
            }
            if (planes) {
                // the 7 bytes of round(psi 2^54): the digit planes the INT8 contraction reads (pk_i8.cuh); a warp writes 32
                // consecutive bytes per plane
                const unsigned long long X0 = i8::psi_fixed(v0), X1 = i8::psi_fixed(v1);
                uint8_t* pq = planes + (size_t)q * np;
#pragma unroll
                for (int t = 0; t < i8::S; ++t) {
                    pq[(size_t)t * plane_stride + r0] = i8::psi_digit(X0, t);
                    pq[(size_t)t * plane_stride + r1] = i8::psi_digit(X1, t);
                }
            }
        }
        if constexpr (KK > 0) {
            if (use_memo && !hit) {
#pragma unroll
                for (int d = 0; d < KK; ++d) mx[d] = xq_r[d];
                __syncwarp();
            }
        }
        const double ys = warp_sum(ys0 + ys1);
        if (lane == 0 && yhat_out) yhat_out[q] = mu + ys;
    }
}

// ---- kernel A, grid mode: the query set is a row-major tensor grid -----------------------------------------
// plot / snapshot / the 4096 x 4096 EI grid of config C5 (krige.py:531-537, 699-708) evaluate a meshgrid: query
// g = row * P + col has the leading coordinates of its row and the last coordinate of its column.  The exponent
// then separates, S(g, r) = lead[row][r] + last[col][r] with lead = sum of the first k - 1 terms (left to right,
// NumPy's order for k < 8) and last = theta |X_r - x|^p of the fast axis, and a CTA that owns an R x R patch of the
// grid computes 2 R tables of np entries once and is left with ONE e^-S per (query, sample): (2 R np) pow instead of
// (R^2 np k).  Same functions on the same inputs in the same order: bit-identical to psi_rows_kernel
// (test_predict_grid_memoisation_is_exact).  The structure is DETECTED on the device (grid_period_kernel: P = length
// of the first run of equal leading coordinates; grid_verify_kernel: every query equals its row's lead and its
// column's last coordinate, exactly); the verdict stays in device memory -- both psi kernels are launched and the
// one that does not apply returns at once, so the call stays asynchronous.
__global__ void grid_init_kernel(int* info) {
    info[0] = 0x7fffffff;  // P
    info[1] = 1;           // ok
}
__global__ void grid_period_kernel(int k, int64_t m, const double* __restrict__ Xq, int* __restrict__ info) {
    const int64_t lim = (m < ((int64_t)1 << 22)) ? m : ((int64_t)1 << 22);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < lim; i += (int64_t)gridDim.x * blockDim.x) {
        bool differs = false;
        for (int d = 0; d < k - 1; ++d) differs |= (Xq[i * k + d] != Xq[d]);
        if (differs) {
            atomicMin(info, (int)i);
            return;  // later indices of this thread are larger
        }
    }
}
__global__ void grid_verify_kernel(int k, int64_t m, const double* __restrict__ Xq, int* __restrict__ info, int min_period) {
    const int P = info[0];
    if (P == 0x7fffffff || P < min_period) {
        if (blockIdx.x == 0 && threadIdx.x == 0) info[1] = 0;
        return;
    }
    bool bad = false;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < m; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = g / P, col = g - row * P;
        for (int d = 0; d < k - 1; ++d) bad |= (Xq[g * k + d] != Xq[row * P * k + d]);
        bad |= (Xq[g * k + k - 1] != Xq[col * k + k - 1]);
    }
    if (bad) info[1] = 0;
}

constexpr int PG_RMAX = 16;
__global__ void __launch_bounds__(PA_THREADS, 1)
psi_rows_grid_kernel(int n, int np, int k, int64_t m, int64_t q0, int64_t mc, int R, const double* __restrict__ X,
                     const double* __restrict__ Xq, const double* __restrict__ hp, const double* __restrict__ w, double mu,
                     const int* __restrict__ gridinfo, double* __restrict__ psi_out, double* __restrict__ yhat_out,
                     uint8_t* __restrict__ planes, size_t plane_stride) {
    if (!gridinfo[1]) return;
    const int P = gridinfo[0];
    extern __shared__ __align__(16) double pa_smem[];
    double* ltab = pa_smem;                                   // 3 x 64 x 16
    double* th_s = ltab + 3 * 64 * 16;                        // PK_MAX_K
    double* p_s = th_s + PK_MAX_K;                            // PK_MAX_K
    int* mode_s = reinterpret_cast<int*>(p_s + PK_MAX_K);     // PK_MAX_K
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(pa_smem);
    unsigned tab_saddr = (smem_base + pa_small_bytes(0, 0) + 0x7fffu) & ~0x7fffu;
    double* tab4k = pa_smem + (tab_saddr - smem_base) / sizeof(double);
    double* lead_t = tab4k + 4096;                            // R x np
    double* last_t = lead_t + (size_t)R * np;                 // R x np
    fill_exp2_4096(tab4k, tid, PA_THREADS);
    fill_log2_tables(ltab, tid, PA_THREADS);
    if (tid < k) {
        const double pd = hp[PK_MAX_K + tid];
        th_s[tid] = hp[tid];
        p_s[tid] = pd;
        mode_s[tid] = (pd == 1.0) ? 1 : (pd == 2.0) ? 2 : (pd >= 0.25 && pd <= 3.0) ? 0 : 3;
    }
    const int bank = tid & 15;
    const int64_t row_lo = q0 / P, row_hi = (q0 + mc - 1) / P;
    const int nrp = (int)((row_hi - row_lo + R) / R), ncp = (P + R - 1) / R;
    const int nit = np / 32;
    for (int patch = blockIdx.x; patch < nrp * ncp; patch += gridDim.x) {
        const int64_t row0 = row_lo + (int64_t)(patch / ncp) * R;
        const int col0 = (patch % ncp) * R;
        __syncthreads();  // tables of the previous patch are no longer read (first pass: th / p / mode are in place)
        for (int e = tid; e < 2 * R * np; e += PA_THREADS) {
            const int which = e / (R * np), rem = e - which * R * np;
            const int idx = rem / np, r = rem - idx * np;
            double v = 0.0;
            if (r < n) {
                if (which == 0) {
                    const int64_t g = (row0 + idx) * P;  // first query of the row
                    if (g < m) {
                        const double* xq = Xq + g * k;
                        for (int d = 0; d < k - 1; ++d)
                            v += pa_term(fabs(__ldg(X + (size_t)r * k + d) - __ldg(xq + d)), th_s[d], p_s[d], mode_s[d], ltab, bank,
                                         tab_saddr);
                    }
                } else {
                    const int col = col0 + idx;
                    if (col < P && col < m) {
                        const int d = k - 1;
                        v = pa_term(fabs(__ldg(X + (size_t)r * k + d) - __ldg(Xq + (size_t)col * k + d)), th_s[d], p_s[d], mode_s[d], ltab,
                                    bank, tab_saddr);
                    }
                }
            }
            (which == 0 ? lead_t : last_t)[idx * np + r] = v;
        }
        __syncthreads();
        for (int qi = warp; qi < R * R; qi += PA_WARPS) {
            const int ri = qi / R, ci = qi - ri * R;
            const int col = col0 + ci;
            const int64_t g = (row0 + ri) * P + col;
            if (col >= P || g < q0 || g >= q0 + mc) continue;  // warp uniform
            const double* lt = lead_t + ri * np;
            const double* ct = last_t + ci * np;
            double ys0 = 0.0, ys1 = 0.0;
            double* prow = psi_out ? psi_out + (size_t)(g - q0) * np : nullptr;
            for (int it = 0; it < nit; it += 2) {
                const int r0 = it * 32 + lane, r1 = r0 + 32;
                const int c0 = min(r0, n - 1), c1 = min(r1, n - 1);
                double v0 = expneg_4k(lt[c0] + ct[c0], tab_saddr), v1 = expneg_4k(lt[c1] + ct[c1], tab_saddr);
                v0 = (r0 < n) ? v0 : 0.0;
                v1 = (r1 < n) ? v1 : 0.0;
                ys0 = fma(v0, __ldg(w + c0), ys0);
                ys1 = fma(v1, __ldg(w + c1), ys1);
                if (prow) {
                    prow[r0] = v0;
                    prow[r1] = v1;
                }
                if (planes) {
                    const unsigned long long X0 = i8::psi_fixed(v0), X1 = i8::psi_fixed(v1);
                    uint8_t* pq = planes + (size_t)(g - q0) * np;
#pragma unroll
                    for (int t = 0; t < i8::S; ++t) {
                        pq[(size_t)t * plane_stride + r0] = i8::psi_digit(X0, t);
                        pq[(size_t)t * plane_stride + r1] = i8::psi_digit(X1, t);
                    }
                }
            }
            const double ys = warp_sum(ys0 + ys1);
            if (lane == 0 && yhat_out) yhat_out[g - q0] = mu + ys;
        }
    }
}

// ---- kernel B: |L^-1 psi|^2 (TMA-fed DMMA), s, EI ------------------------------------------------------
namespace tn {
constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;
constexpr int KC = 16;                       // samples per pipeline chunk: one stage row = 128 bytes
constexpr int QT = 64;                       // queries per CTA tile
constexpr int CB = 256;                      // columns of v per pass
constexpr int STAGES = 4;                    // 5 stages measured 1.6 % slower
constexpr int BOX_DOUBLES = 64 * KC;         // one TMA box: 64 rows x 16 doubles = 8 KB
constexpr int A_STAGE = QT * KC;
constexpr int B_STAGE = CB * KC;
constexpr int STAGE_DOUBLES = A_STAGE + B_STAGE;   // 40 KB
constexpr int SMEM_ALIGN = 1024;
// + per-tile partial sums qsm[2][8][64] + mbarriers + release counters
constexpr int SMEM_DOUBLES = STAGES * STAGE_DOUBLES + 2 * 8 * QT + STAGES + STAGES + 2;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

// acc(32 x 32 per warp: 4 x 4 DMMA tiles, column blocks LO .. HI-1 only) += A(32 x 16) * B(32 x 16)^T for one
// stage.  Stage rows are 128 bytes with the 32-byte atoms XOR-swizzled by the row (SWIZZLE_128B_ATOM_32B):
// lane (fr, fc) reads column 4i + fc of a row congruent to fr mod 8 at soff[i] = ((i ^ (fr & 3)) << 2) + fc.
template <int LO, int HI>
__device__ __forceinline__ void consume(double (&acc)[4][4][2], const double* a_s, const double* b_even,
                                        const double* b_odd, const int (&soff)[4]) {
    // (skipping the k-steps of the first live block that lie right of the diagonal as well -- a warp-uniform branch
    // per k-step, 3 % fewer DMMAs -- measured 3.5 % SLOWER: the branch splits the DMMA stream of a k-step)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a[4], b[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = a_s[mi * 8 * KC + soff[i]];
#pragma unroll
        for (int ni = LO; ni < HI; ++ni) b[ni] = ((ni & 1) ? b_odd : b_even)[ni * 64 * KC + soff[i]];
#pragma unroll
        for (int ni = LO; ni < HI; ++ni)
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

__global__ void __launch_bounds__(THREADS, 1)
trinorm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, int np, int64_t mc,
               double sigma2, double y_min, double ei_weight, double var_extra, const double* __restrict__ yhat,
               double* __restrict__ s_out, double* __restrict__ ei_out) {
    extern __shared__ __align__(16) double tn_raw[];
    double* smem = tn_raw + ((SMEM_ALIGN - ((unsigned)__cvta_generic_to_shared(tn_raw) & (SMEM_ALIGN - 1))) &
                             (SMEM_ALIGN - 1)) / sizeof(double);
    double* pipe = smem;
    double (*qsm)[8][QT] = reinterpret_cast<double (*)[8][QT]>(pipe + STAGES * STAGE_DOUBLES);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(qsm + 2);
    int* rel = reinterpret_cast<int*>(bars + STAGES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int wq = warp & 1, cg = warp >> 1;  // query half (32 rows), column group: 8-column blocks cg, cg + 8, ...
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + s, 1);
            rel[s] = 0;
        }
        mbar_fence_init();
    }
    __syncthreads();
    int soff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) soff[i] = ((i ^ (fr & 3)) << 2) + fc;

    const int64_t ntiles = (mc + QT - 1) / QT;
    const int ncb = (np + CB - 1) / CB;
    // chunks of one tile: column pass cb needs the samples r < min(np, (cb + 1) * 256)
    // position of a chunk in the CTA's flat sequence: (tile, cb, kc)
    struct Pos {
        int64_t tile;
        int cb, kc;
    };
    auto nk_of = [&](int cb) { return min(np, (cb + 1) * CB) / KC; };
    auto advance = [&](Pos& p) {
        if (++p.kc == nk_of(p.cb)) {
            p.kc = 0;
            if (++p.cb == ncb) {
                p.cb = 0;
                p.tile += gridDim.x;
            }
        }
    };
    auto issue = [&](const Pos& p, unsigned g) {
        const int ps = (int)(g % STAGES);
        unsigned long long* bar = bars + ps;
        const int c0 = p.cb * CB;
        const int nbox = min(4, (np - c0) / 64);
        // boxes of 64 columns that lie entirely left of this chunk's samples are zeros of the triangle: not loaded
        int b0 = (p.kc * KC - c0) >> 6;
        b0 = (b0 < 0) ? 0 : b0;
        mbar_arrive_expect_tx(bar, (unsigned)((1 + nbox - b0) * BOX_DOUBLES * sizeof(double)));
        double* st = pipe + ps * STAGE_DOUBLES;
        tma_load_2d(st, &tmapA, p.kc * KC, (int)(p.tile * QT), bar);
        for (int b = b0; b < nbox; ++b) tma_load_2d(st + A_STAGE + b * BOX_DOUBLES, &tmapB, p.kc * KC, c0 + 64 * b, bar);
    };
    Pos cons{(int64_t)blockIdx.x, 0, 0};
    if (cons.tile >= ntiles) return;
    Pos prod = cons;
    if (tid == 0) {
        Pos p = prod;
        for (int c = 0; c < STAGES && p.tile < ntiles; ++c) {
            issue(p, (unsigned)c);
            advance(p);
        }
    }
    for (int c = 0; c < STAGES; ++c) advance(prod);  // every warp tracks the chunk STAGES ahead of its own

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    double qpart[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned g = 0;
    int tpar = 0;
    while (cons.tile < ntiles) {
        const int ps = (int)(g % STAGES);
        mbar_wait(bars + ps, (g / STAGES) & 1u);
        const double* st = pipe + ps * STAGE_DOUBLES;
        const double* a_s = st + (wq * 32 + fr) * KC;
        // this warp's 8-column blocks of the pass: columns c0 + 64 ni + 8 cg for even ni, c0 + 64 ni + 8 (7 - cg) for
        // odd ni -- the triangle gives a block at column c about c / 16 live chunks, and mirroring every other block
        // hands all warps the same number of DMMAs per pass (with cg everywhere, warp 7 did 20 % more than warp 0 and
        // the others spun on the next stage's mbarrier: 11 % of the stall samples)
        const double* b_even = st + A_STAGE + (cg * 8 + fr) * KC;
        const double* b_odd = st + A_STAGE + ((7 - cg) * 8 + fr) * KC;
        const int c0 = cons.cb * CB;
        // live column blocks: ni with (first column of the block) + 7 >= 16 kc (lower triangle) and < np; the block
        // columns grow with ni, so the live ones are ni = lo .. hi - 1
        const int edge = KC * cons.kc - c0 - 7;  // a block starting at relative column >= edge is live
        int lo = 0;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) lo += (64 * ni + 8 * ((ni & 1) ? 7 - cg : cg) < edge) ? 1 : 0;
        const int hi = min(4, (np - c0) >> 6);
        switch (lo * 4 + hi - 1) {
            case 0: consume<0, 1>(acc, a_s, b_even, b_odd, soff); break;
            case 1: consume<0, 2>(acc, a_s, b_even, b_odd, soff); break;
            case 2: consume<0, 3>(acc, a_s, b_even, b_odd, soff); break;
            case 3: consume<0, 4>(acc, a_s, b_even, b_odd, soff); break;
            case 5: consume<1, 2>(acc, a_s, b_even, b_odd, soff); break;
            case 6: consume<1, 3>(acc, a_s, b_even, b_odd, soff); break;
            case 7: consume<1, 4>(acc, a_s, b_even, b_odd, soff); break;
            case 10: consume<2, 3>(acc, a_s, b_even, b_odd, soff); break;
            case 11: consume<2, 4>(acc, a_s, b_even, b_odd, soff); break;
            case 15: consume<3, 4>(acc, a_s, b_even, b_odd, soff); break;
            default: break;
        }
        __syncwarp();
        if (lane == 0 && atomicAdd(rel + ps, 1) == WARPS - 1) {  // last warp to release this stage: refill it
            rel[ps] = 0;
            if (prod.tile < ntiles) issue(prod, g + STAGES);
        }
        advance(prod);
        ++g;
        const bool pass_end = (cons.kc + 1 == nk_of(cons.cb));
        const bool tile_end = pass_end && (cons.cb + 1 == ncb);
        if (pass_end) {
            // |v|^2 of this pass' columns, accumulators back to zero
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                double v = qpart[mi];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    v = fma(acc[mi][ni][0], acc[mi][ni][0], v);
                    v = fma(acc[mi][ni][1], acc[mi][ni][1], v);
                    acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                }
                qpart[mi] = v;
            }
        }
        if (tile_end) {
            const int64_t q0 = cons.tile * QT;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                double v = qpart[mi];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (fc == 0) qsm[tpar][cg][wq * 32 + mi * 8 + fr] = v;
                qpart[mi] = 0.0;
            }
            // one CTA barrier per tile (48 chunks at np = 512); the TMA queue keeps running underneath.  Letting the
            // last warp to arrive finish the tile instead (no barrier at all) measured 1.3 % slower.
            __syncthreads();
            const bool last = (warp == 0);
            if (last) {
#pragma unroll
                for (int h2 = 0; h2 < QT / 32; ++h2) {
                    const int row = h2 * 32 + lane;
                    if (q0 + row < mc) {
                        double qs = 0.0;
#pragma unroll
                        for (int z = 0; z < 8; ++z) qs += qsm[tpar][z][row];
                        const double ssqr = sigma2 * (1.0 + var_extra - qs);
                        const double S = sqrt(fabs(ssqr));
                        if (s_out) s_out[q0 + row] = S;
                        if (ei_out) ei_out[q0 + row] = expected_improvement(yhat[q0 + row], S, y_min, ei_weight);
                    }
                }
            }
            tpar ^= 1;
        }
        advance(cons);
    }
}

static cudaError_t encode_2d(CUtensorMap* tmap, const double* base, uint64_t cols, uint64_t rows) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = (encode_fn)fn;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)KC, 64};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (r == CUDA_SUCCESS) ? cudaSuccess : cudaErrorInvalidValue;
}
}  // namespace tn

template <int KK>
static cudaError_t launch_psi_rows(pk_handle* h, int64_t mc, const double* Xq, const double* hp, double mu,
                                   double* psi_out, double* yhat_out, const int* gridinfo, uint8_t* planes, size_t plane_stride) {
    const int k = h->k, np = h->fit_np;
    const int xt = (k * np <= PA_XT_MAX) ? 1 : 0;
    // memoised leading dimensions (compile-time k only): as many of the first k - 1 as fit beside the tables
    int memo_dims = 0;
    if (KK > 1) {
        const int room = PA_SMALL_MAX - (int)pa_small_bytes(xt ? k * np : 0, 0);
        const int per_dim = PA_WARPS * np * (int)sizeof(double);
        memo_dims = (room > 0) ? room / per_dim : 0;
        if (memo_dims > KK - 1) memo_dims = KK - 1;
    }
    // the dynamic window starts 1 KB into the CTA's shared memory: leave room for the 32 KB-aligned table wherever
    // the base lies (same sizing rule as psi_pop_kernel)
    const size_t smem =
        ((pa_small_bytes(xt ? k * np : 0, PA_WARPS * memo_dims * np) + 0x400 + 0x7fff) & ~(size_t)0x7fff) + 0x8000;
    static bool attr_done_dev[PK_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(psi_rows_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const int64_t want = (mc + PA_WARPS - 1) / PA_WARPS;
    const int grid = (int)((want < (int64_t)h->sm_count) ? want : (int64_t)h->sm_count);
    psi_rows_kernel<KK><<<grid, PA_THREADS, smem, h->stream>>>(h->n, np, k, mc, h->dX, Xq, hp, h->d_w, mu, xt, memo_dims,
                                                               psi_out, yhat_out, gridinfo, planes, plane_stride);
    h->launches++;
    return cudaGetLastError();
}

// yhat / s / ei may each be null; all pointers are device pointers.
cudaError_t launch_predict_split(pk_handle* h, int64_t m, const double* Xq, const double* hp, double mu, double sigma2,
                                 double y_min, double ei_weight, double var_extra, double* yhat, double* sout,
                                 double* ei) {
    const int np = h->fit_np, k = h->k;
    const bool need_s = (sout != nullptr || ei != nullptr);
    // the contraction |L^-1 psi|^2: INT8 tensor cores (pk_predict_i8.cu; the psi kernels then write digit planes, 7 bytes
    // per entry, instead of doubles) or the FP64 tensor pipe (tn::trinorm_kernel on an FP64 psi scratch)
    const bool use_i8 = need_s && h->predict_contraction != 1;
    // queries per block: psi scratch of at most 1 GiB, at least one wave of 64-query tiles
    int64_t mcap = ((int64_t)1 << 30) / ((int64_t)np * 8);
    mcap = (mcap / 128) * 128;
    if (mcap < 128 * (int64_t)h->sm_count) mcap = 128 * (int64_t)h->sm_count;
    if (mcap > m) mcap = ((m + 127) / 128) * 128;
    // FP64 scratch: psi rows (DMMA contraction only) + yhat (when the caller wants none)
    const size_t need = need_s ? (use_i8 ? 0 : (size_t)mcap * np) + (size_t)mcap : 0;
    if (need > h->cap_pscratch) {
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) return e;
        if (h->d_pscratch) cudaFree(h->d_pscratch);
        h->d_pscratch = nullptr;
        h->cap_pscratch = 0;
        e = cudaMalloc((void**)&h->d_pscratch, need * sizeof(double));
        if (e != cudaSuccess) return e;
        h->cap_pscratch = need;
    }
    double* psi = (need_s && !use_i8) ? h->d_pscratch : nullptr;
    double* ytmp = need_s ? h->d_pscratch + (use_i8 ? 0 : (size_t)mcap * np) : nullptr;
    uint8_t* planes = nullptr;
    size_t plane_stride = 0;
    if (use_i8) {
        cudaError_t e = psi_planes_i8(h, mcap, &planes, &plane_stride);
        if (e != cudaSuccess) return e;
    }
    CUtensorMap tmapB;
    const size_t smemB = sizeof(double) * (size_t)tn::SMEM_DOUBLES + tn::SMEM_ALIGN;
    if (need_s && !use_i8) {
        cudaError_t e = tn::encode_2d(&tmapB, h->d_linv, (uint64_t)np, (uint64_t)np);
        if (e != cudaSuccess) return e;
        static bool attr_done_dev[PK_MAX_DEVICES] = {};
        bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
        if (!attr_done) {
            e = cudaFuncSetAttribute(tn::trinorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB);
            if (e != cudaSuccess) return e;
            attr_done = true;
        }
    }
    // row-major tensor grid?  (2 <= k < 8: the separated exponent needs NumPy's left-to-right summation)
    int* gridinfo = nullptr;
    int R = 0;
    if (k >= 2 && k < 8 && h->predict_grid_mode) {
        R = (int)(((size_t)128 * 1024) / ((size_t)2 * np * sizeof(double)));
        if (R > PG_RMAX) R = PG_RMAX;
        if (R >= 2) {
            if (!h->d_gridinfo) {
                cudaError_t e = cudaMalloc((void**)&h->d_gridinfo, 2 * sizeof(int));
                if (e != cudaSuccess) return e;
            }
            gridinfo = h->d_gridinfo;
            grid_init_kernel<<<1, 1, 0, h->stream>>>(gridinfo);
            grid_period_kernel<<<2 * h->sm_count, 256, 0, h->stream>>>(k, m, Xq, gridinfo);
            grid_verify_kernel<<<4 * h->sm_count, 256, 0, h->stream>>>(k, m, Xq, gridinfo, 2 * R);
            h->launches += 3;
            static bool attr_done_dev[PK_MAX_DEVICES] = {};
            bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
            if (!attr_done) {
                cudaError_t e = cudaFuncSetAttribute(psi_rows_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (e != cudaSuccess) return e;
                attr_done = true;
            }
        }
    }
    for (int64_t q0 = 0; q0 < m; q0 += mcap) {
        const int64_t mc = (m - q0 < mcap) ? (m - q0) : mcap;
        double* yh = yhat ? yhat + q0 : ytmp;
        cudaError_t e;
        if (gridinfo) {
            // table + small arrays as in psi_rows_kernel (64 KB with the alignment slack) + the 2 R patch tables
            const size_t smem = ((pa_small_bytes(0, 0) + 0x400 + 0x7fff) & ~(size_t)0x7fff) + 0x8000 +
                                (size_t)2 * R * np * sizeof(double);
            psi_rows_grid_kernel<<<h->sm_count, PA_THREADS, smem, h->stream>>>(h->n, np, k, m, q0, mc, R, h->dX, Xq, hp, h->d_w, mu,
                                                                              gridinfo, psi, yh, planes, plane_stride);
            h->launches++;
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        switch (k) {
#define PK_PA(KK_) case KK_: e = launch_psi_rows<KK_>(h, mc, Xq + q0 * k, hp, mu, psi, yh, gridinfo, planes, plane_stride); break;
            PK_PA(1) PK_PA(2) PK_PA(3)
#undef PK_PA
            default: e = launch_psi_rows<0>(h, mc, Xq + q0 * k, hp, mu, psi, yh, gridinfo, planes, plane_stride); break;
        }
        if (e != cudaSuccess) return e;
        if (!need_s) continue;
        if (use_i8) {
            // the contraction on the INT8 tensor cores (pk_predict_i8.cu): exact digit products, twice the DMMA peak
            e = launch_trinorm_i8(h, mc, mcap, sigma2, y_min, ei_weight, var_extra, yh, sout ? sout + q0 : nullptr,
                                  ei ? ei + q0 : nullptr);
            if (e != cudaSuccess) return e;
            continue;
        }
        CUtensorMap tmapA;
        e = tn::encode_2d(&tmapA, psi, (uint64_t)np, (uint64_t)mc);
        if (e != cudaSuccess) return e;
        const int64_t ntiles = (mc + tn::QT - 1) / tn::QT;
        const int grid = (int)((ntiles < (int64_t)h->sm_count) ? ntiles : (int64_t)h->sm_count);
        tn::trinorm_kernel<<<grid, tn::THREADS, smemB, h->stream>>>(tmapA, tmapB, np, mc, sigma2, y_min, ei_weight, var_extra,
                                                                    yh, sout ? sout + q0 : nullptr, ei ? ei + q0 : nullptr);
        h->launches++;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace pk
