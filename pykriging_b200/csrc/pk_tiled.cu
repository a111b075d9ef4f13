// Large-n likelihood path: batched Psi construction + blocked left-looking Cholesky whose GEMM
// runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), with the forward
// substitutions L^-1 1, L^-1 y (and, for pk_fit, L^-1 itself) carried as augmented rows.
//
// Replaces matrixops.updatePsi/regupdatePsi (matrixops.py:24-42: Psi build + np.linalg.cholesky)
// and matrixops.neglikelihood/regneglikelihood (matrixops.py:45-66) for a whole population of
// candidates (the loop of krige.py:436-451).
//
// Data layout (HBM), per candidate one row-major matrix of `rows x ld` doubles (see TiledShape):
// Psi tiles (lower triangle) are written once by psi_build_kernel, then each 64-wide block column
// j is finalised in place by
//     chol_diag_kernel : T = Psi_jj - L[j,0:j] L[j,0:j]^T ; unblocked Cholesky of the 64x64 tile
//     chol_off_kernel  : T = M[i,j] - L[i,0:j] L[j,0:j]^T ; L[i,j] = T L_jj^-T   (all rows below)
// Left-looking means every tile of the matrix is read once and written once; the trailing matrix
// is never re-written (HBM traffic ~ n^3/(6*64) doubles of L re-reads per candidate).
#include "pk_internal.h"

namespace pk {

// ------------------------------------------------------------------------------------------------
// Psi construction for a population: one CTA per (lower tile, candidate).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
psi_build_kernel(int n, int k, int np, int ld, size_t elems, const double* __restrict__ X,
                 const double* __restrict__ theta, const double* __restrict__ p,
                 const double* __restrict__ lambda, int mode, double* __restrict__ ws) {
    __shared__ double Xr[PK_TILE * 17];  // row-tile coordinates (k chunked by 16, stride 17)
    __shared__ double Xc[PK_TILE * 17];
    __shared__ double th[PK_MAX_K], pp[PK_MAX_K];
    // decode lower-triangular tile index -> (ti >= tj)
    const int t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    const int tj = t - ti * (ti + 1) / 2;
    const int cand = blockIdx.y;
    const int tid = threadIdx.x;
    double* M = ws + (size_t)cand * elems;
    if (tid < k) {
        th[tid] = theta[(size_t)cand * k + tid];
        pp[tid] = p[(size_t)cand * k + tid];
    }
    const double diag = 1.0 + (mode == PK_MODE_REGRESSION ? lambda[cand] : PK_EPS_DIAG);
    const int r0 = ti * PK_TILE, c0 = tj * PK_TILE;
    // each thread: row lr = tid/4 (0..63), 16 columns lc = (tid%4)*16 .. +15
    const int lr = tid >> 2, lcb = (tid & 3) * 16;
    double S[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) S[i] = 0.0;
    const bool numpy_pairwise = (k >= 8);
    if (!numpy_pairwise) {
        // k < 8: plain left-to-right accumulation over dimensions (NumPy order), one chunk
        __syncthreads();
        for (int i = tid; i < PK_TILE * k; i += 256) {
            const int rr = i / k, d = i % k;
            Xr[rr * 17 + d] = (r0 + rr < n) ? X[(size_t)(r0 + rr) * k + d] : 0.0;
            Xc[rr * 17 + d] = (c0 + rr < n) ? X[(size_t)(c0 + rr) * k + d] : 0.0;
        }
        __syncthreads();
        for (int d = 0; d < k; ++d) {
            const double xr = Xr[lr * 17 + d], t_ = th[d], p_ = pp[d];
#pragma unroll
            for (int i = 0; i < 16; ++i) S[i] += t_ * pow_abs(fabs(xr - Xc[(lcb + i) * 17 + d]), p_);
        }
    } else {
        // k >= 8: NumPy's pairwise order needs all terms of a pair at once -> per-entry evaluation
        __syncthreads();
        for (int i = 0; i < 16; ++i) {
            const int gr = r0 + lr, gc = c0 + lcb + i;
            if (gr < n && gc < gr)
                S[i] = corr_exponent(X + (size_t)gr * k, X + (size_t)gc * k, th, pp, k);
        }
    }
    double out[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int gr = r0 + lr, gc = c0 + lcb + i;
        double v;
        if (gr >= n || gc >= n) v = (gr == gc) ? 1.0 : 0.0;     // identity padding
        else if (gr == gc) v = diag;
        else if (gc > gr) v = 0.0;                                  // strictly upper part of diagonal tiles
        else v = exp(-S[i]);
        out[i] = v;
    }
    double2* dst = reinterpret_cast<double2*>(M + (size_t)(r0 + lr) * ld + c0 + lcb);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_double2(out[2 * i], out[2 * i + 1]);
}

// augmented rows (+ identity block for pk_fit) and info reset
__global__ void aug_init_kernel(int n, int np, int ld, int rows, size_t elems, const double* __restrict__ y,
                                double* __restrict__ ws, int32_t* __restrict__ info) {
    const int cand = blockIdx.y;
    double* M = ws + (size_t)cand * elems;
    const size_t total = (size_t)(rows - np) * ld;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / ld), c = (int)(i % ld);
        double v = 0.0;
        if (r == 0) v = (c < n) ? 1.0 : 0.0;
        else if (r == 1) v = (c < n) ? y[c] : 0.0;
        else if (r >= PK_AUG_ROWS) v = (r - PK_AUG_ROWS == c) ? 1.0 : 0.0;
        M[(size_t)(np + r) * ld + c] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) info[cand] = 0;
}

// ------------------------------------------------------------------------------------------------
// GEMM core: acc(rows x 64) = A(rows x K) * B(64 x K)^T, A and B row-major with leading dimension ld
// (K contiguous), staged through shared memory with a multi-stage cp.async pipeline and consumed by
// DMMA 8x8x4 fragments.  256 threads = 4 (M) x 2 (N) warps, warp tile 32x32 (4x4 DMMA tiles,
// 32 FP64 accumulators per thread), CTA tile up to 128 x 64.
// ------------------------------------------------------------------------------------------------
constexpr int KC = 16;            // K chunk per pipeline stage
constexpr int SLD = KC + 4;       // smem row stride (doubles): 20 -> conflict-free fragment loads
constexpr int STAGES = 3;
constexpr int BM = 128;           // rows per CTA tile
constexpr int THREADS = 256;
constexpr int A_STAGE = BM * SLD;
constexpr int B_STAGE = PK_TILE * SLD;
constexpr int PIPE_DOUBLES = STAGES * (A_STAGE + B_STAGE);
constexpr int TLD = PK_TILE + 1;  // 65: odd stride -> conflict-free row-per-thread walks (diag tile)
constexpr int LLD = PK_TILE + 4;  // 68 (== 4 mod 16): L_jj rows, conflict-free DMMA B-fragment loads
constexpr int XLD = 12;           // solved 8-column block of the panel, conflict-free A-fragment loads
constexpr int TRSM_DOUBLES = PK_TILE * LLD + PK_TILE + 2 * BM * XLD;
constexpr int DIAG_DOUBLES = PK_TILE * TLD + PK_TILE * LLD + 2 * PK_TILE;
constexpr int SMEM_MAX1 = (PIPE_DOUBLES > TRSM_DOUBLES) ? PIPE_DOUBLES : TRSM_DOUBLES;
constexpr int SMEM_DOUBLES = (SMEM_MAX1 > DIAG_DOUBLES) ? SMEM_MAX1 : DIAG_DOUBLES;

// Rows of a CTA tile: local rows [0, split) map to global rows rowA + r, local rows [split, ..) to
// rowB + (r - split).  (The look-ahead CTA carries tile (j+1, j) AND the 8 augmented rows.)
struct RowMap {
    int rowA, split, rowB;
    __device__ __forceinline__ int operator()(int r) const { return r < split ? rowA + r : rowB + (r - split); }
};

// a_rows: rows of A actually loaded (multiple of 8, <= BM)
__device__ __forceinline__ void gemm_load_stage(double* As, double* Bs, const double* __restrict__ M,
                                                const RowMap& rm, int brow0, int ld, int k0, int a_rows) {
    for (int i = threadIdx.x; i < a_rows * 8; i += THREADS) {
        const int r = i >> 3, ch = i & 7;
        cp_async16(As + r * SLD + ch * 2, M + (size_t)rm(r) * ld + k0 + ch * 2);
    }
    for (int i = threadIdx.x; i < PK_TILE * 8; i += THREADS) {
        const int r = i >> 3, ch = i & 7;
        cp_async16(Bs + r * SLD + ch * 2, M + (size_t)(brow0 + r) * ld + k0 + ch * 2);
    }
}

__device__ __forceinline__ double neg(double x) {
    return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));  // sign flip on the ALU pipe
}

// acc[mi][ni][2] -= A(rows x K) * B(64 x K)^T ; warp (wm, wn) covers rows wm*32.., cols wn*32..
__device__ __forceinline__ void gemm_mainloop(double (&acc)[4][4][2], double* pipe, const double* __restrict__ M,
                                              const RowMap& rm, int brow0, int ld, int K, int a_rows, int wm, int wn,
                                              int mi_count) {
    double* As = pipe;
    double* Bs = pipe + STAGES * A_STAGE;
    const int lane = threadIdx.x & 31;
    const int nk = K / KC;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) gemm_load_stage(As + s * A_STAGE, Bs + s * B_STAGE, M, rm, brow0, ld, s * KC, a_rows);
        cp_async_commit();
    }
    const int fr = lane >> 2, fc = lane & 3;
    for (int it = 0; it < nk; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = it + STAGES - 1;
        if (nxt < nk) {
            const int s = nxt % STAGES;
            gemm_load_stage(As + s * A_STAGE, Bs + s * B_STAGE, M, rm, brow0, ld, nxt * KC, a_rows);
        }
        cp_async_commit();
        const double* a_s = As + (it % STAGES) * A_STAGE + (wm * 32 + fr) * SLD + fc;
        const double* b_s = Bs + (it % STAGES) * B_STAGE + (wn * 32 + fr) * SLD + fc;
        if (mi_count > 0) {
#pragma unroll
            for (int kk = 0; kk < KC; kk += 4) {
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = (mi < mi_count) ? neg(a_s[mi * 8 * SLD + kk]) : 0.0;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = b_s[ni * 8 * SLD + kk];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    if (mi < mi_count) {
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

// acc = M[rows, col0 .. col0+64) in DMMA accumulator layout (issued before the main loop: the loads
// overlap the pipeline prologue)
__device__ __forceinline__ void load_acc(double (&acc)[4][4][2], const double* __restrict__ M, const RowMap& rm, int ld,
                                         int col0, int wm, int wn, int mi_count) {
    const int lane = threadIdx.x & 31;
    const int fr = lane >> 2, fc2 = (lane & 3) * 2;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            if (mi < mi_count) {
                const int r = wm * 32 + mi * 8 + fr, c = wn * 32 + ni * 8 + fc2;
                const double2 v = *reinterpret_cast<const double2*>(M + (size_t)rm(r) * ld + col0 + c);
                acc[mi][ni][0] = v.x;
                acc[mi][ni][1] = v.y;
            } else {
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// In-tile unblocked Cholesky of the 64x64 tile T (smem, stride TLD), 256 threads, ONE barrier per
// column.  Off-diagonal entries are updated by sequential FMAs in k order; every pivot is formed as
// t_cc - (sum_k l_ck^2) with the sum accumulated separately (dsum) and subtracted once -- the
// dot-product form of LAPACK's dpotf2 (see pk_small.cu).  The scaled column is written to Lout
// (smem, stride LLD) while T keeps the unscaled values, so no second barrier is needed.
// Returns 0 or the (1-based, tile-local) index of the first pivot <= piv_tol * diag0[c].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int potf2_tile(double* T, double* Lout, double* dsum, const double* diag0,
                                          double piv_tol) {
    const int tid = threadIdx.x;
    const int r = tid & 63, h = tid >> 6;  // 4 threads per row
    if (tid < PK_TILE) dsum[tid] = 0.0;
    for (int c = 0; c < PK_TILE; ++c) {
        __syncthreads();
        const double piv = T[c * TLD + c] - dsum[c];
        if (!(piv > piv_tol * diag0[c])) return c + 1;  // uniform: every thread reads the same values
        // 1/sqrt(piv) straight from the reciprocal-square-root unit (+ Newton), l = piv * rinv: halves the
        // serial sqrt -> divide latency chain of the 64 dependent pivot steps (<= 2 ulp on l and rinv)
        const double rinv = rsqrt(piv);
        const double l = piv * rinv;
        if (r > c) {
            const double lrc = T[r * TLD + c] * rinv;
            int cc = c + 1 + h;
            for (; cc + 12 < r; cc += 16) {  // 4 independent updates in flight
                const double u0 = T[cc * TLD + c], u1 = T[(cc + 4) * TLD + c], u2 = T[(cc + 8) * TLD + c],
                             u3 = T[(cc + 12) * TLD + c];
                const double v0 = T[r * TLD + cc], v1 = T[r * TLD + cc + 4], v2 = T[r * TLD + cc + 8],
                             v3 = T[r * TLD + cc + 12];
                T[r * TLD + cc] = fma(-lrc, u0 * rinv, v0);
                T[r * TLD + cc + 4] = fma(-lrc, u1 * rinv, v1);
                T[r * TLD + cc + 8] = fma(-lrc, u2 * rinv, v2);
                T[r * TLD + cc + 12] = fma(-lrc, u3 * rinv, v3);
            }
            for (; cc < r; cc += 4) T[r * TLD + cc] = fma(-lrc, T[cc * TLD + c] * rinv, T[r * TLD + cc]);
            if (h == 0) {
                Lout[r * LLD + c] = lrc;
                dsum[r] = fma(lrc, lrc, dsum[r]);
            }
        } else if (h == 0) {
            Lout[r * LLD + c] = (r == c) ? l : 0.0;
        }
    }
    __syncthreads();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Panel rows of block column j:  T = M[rows, j] - L[rows, 0:j] L[j, 0:j]^T  (DMMA main loop, T stays
// in the accumulator registers), then X L_jj^T = T solved IN REGISTERS:
//   for each 8-column block cb: the warps that own those columns run the 8-step substitution inside
//   each thread quad (the four lanes of a quad hold the 8 columns of a row; the solved value is
//   broadcast with a quad shuffle), publish the solved block to shared memory, and after one barrier
//   every warp folds it into its remaining columns with two DMMA k-steps per 8x8 tile.
// 8 barriers, ~230 FP64 instructions per thread, 896 DMMAs per CTA -- instead of a 2016-FMA
// row-per-thread substitution fed from shared memory.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void panel_rows(double* smem, double* __restrict__ M, int ld, int j, const RowMap& rm,
                                           int rows_valid) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int mi_count = max(0, min(4, (rows_valid - wm * 32) / 8));
    const int fr = lane >> 2, q = lane & 3;
    double acc[4][4][2];
    load_acc(acc, M, rm, ld, j * PK_TILE, wm, wn, mi_count);
    gemm_mainloop(acc, smem, M, rm, j * PK_TILE, ld, j * PK_TILE, rows_valid, wm, wn, mi_count);

    double* Ljj = smem;                        // 64 x LLD
    double* rdiag = Ljj + PK_TILE * LLD;       // 64
    double* Xs = rdiag + PK_TILE;              // 2 x BM x XLD
    for (int i = tid; i < PK_TILE * (PK_TILE / 2); i += THREADS) {
        const int rr = i >> 5, c2 = (i & 31) * 2;
        const double2 v = *reinterpret_cast<const double2*>(M + (size_t)(j * PK_TILE + rr) * ld + j * PK_TILE + c2);
        *reinterpret_cast<double2*>(Ljj + rr * LLD + c2) = v;
        if (c2 == rr) rdiag[rr] = 1.0 / v.x;
        else if (c2 + 1 == rr) rdiag[rr] = 1.0 / v.y;
    }
    __syncthreads();

#pragma unroll 1
    for (int cb = 0; cb < 8; ++cb) {
        double* Xb = Xs + (cb & 1) * BM * XLD;
        if (wn == (cb >> 2)) {
            // move the owned 8-column block into a small register array (keeps the loop rolled: the
            // accumulator file is only ever indexed with compile-time constants)
            double t0[4], t1[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                switch (cb & 3) {
                    case 0: t0[mi] = acc[mi][0][0]; t1[mi] = acc[mi][0][1]; break;
                    case 1: t0[mi] = acc[mi][1][0]; t1[mi] = acc[mi][1][1]; break;
                    case 2: t0[mi] = acc[mi][2][0]; t1[mi] = acc[mi][2][1]; break;
                    default: t0[mi] = acc[mi][3][0]; t1[mi] = acc[mi][3][1]; break;
                }
            }
            const double* Lb = Ljj + (cb * 8) * LLD + cb * 8;   // diagonal 8x8 block of L_jj
            // 8-step substitution inside the quad; the 4 row groups (mi) are independent chains
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int qo = c >> 1;
                const double rd = rdiag[cb * 8 + c];
                const double l0 = Lb[(2 * q) * LLD + c];      // L[col 2q  ][c]
                const double l1 = Lb[(2 * q + 1) * LLD + c];  // L[col 2q+1][c]
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    double xc = ((c & 1) ? t1[mi] : t0[mi]) * rd;
                    xc = __shfl_sync(0xffffffffu, xc, (lane & ~3) | qo);
                    if (q == qo) {
                        if (c & 1) t1[mi] = xc; else t0[mi] = xc;
                    }
                    if (2 * q > c) t0[mi] = fma(-xc, l0, t0[mi]);
                    if (2 * q + 1 > c) t1[mi] = fma(-xc, l1, t1[mi]);
                }
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                switch (cb & 3) {
                    case 0: acc[mi][0][0] = t0[mi]; acc[mi][0][1] = t1[mi]; break;
                    case 1: acc[mi][1][0] = t0[mi]; acc[mi][1][1] = t1[mi]; break;
                    case 2: acc[mi][2][0] = t0[mi]; acc[mi][2][1] = t1[mi]; break;
                    default: acc[mi][3][0] = t0[mi]; acc[mi][3][1] = t1[mi]; break;
                }
                if (mi < mi_count) {
                    const int r = wm * 32 + mi * 8 + fr;
                    Xb[r * XLD + 2 * q] = t0[mi];
                    Xb[r * XLD + 2 * q + 1] = t1[mi];
                }
            }
        }
        if (cb == 7) break;
        __syncthreads();
        // fold the solved block into the remaining columns: acc[:, gb] -= X_cb * L[gb, cb]^T
        if (mi_count > 0) {
            double a[4][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                a[mi][0] = neg(Xb[(wm * 32 + mi * 8 + fr) * XLD + q]);
                a[mi][1] = neg(Xb[(wm * 32 + mi * 8 + fr) * XLD + 4 + q]);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int gb = wn * 4 + ni;
                if (gb > cb) {
                    const double b0 = Ljj[(gb * 8 + fr) * LLD + cb * 8 + q];
                    const double b1 = Ljj[(gb * 8 + fr) * LLD + cb * 8 + 4 + q];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        if (mi < mi_count) {
                            dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi][0], b0);
                            dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi][1], b1);
                        }
                    }
                }
            }
        }
    }
    // write the solved panel back (accumulator layout -> 16-byte stores, 64 B contiguous per quad)
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        if (mi < mi_count) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int r = wm * 32 + mi * 8 + fr, c = wn * 32 + ni * 8 + 2 * q;
                *reinterpret_cast<double2*>(M + (size_t)rm(r) * ld + j * PK_TILE + c) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
        }
    }
    __syncthreads();  // global writes of this CTA are visible to all of its threads from here on
}

// ------------------------------------------------------------------------------------------------
// Diagonal tile jd: T = Psi[jd,jd] - L[jd, 0:jd] L[jd, 0:jd]^T, unblocked Cholesky, write L[jd,jd].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void diag_tile(double* smem, double* __restrict__ M, int ld, int jd, double piv_tol,
                                          int32_t* __restrict__ info, int cand) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int mi_count = (wm < 2) ? 4 : 0;  // 64 rows: two of the four M-warps hold data
    const RowMap rm{jd * PK_TILE, BM, 0};
    double acc[4][4][2];
    load_acc(acc, M, rm, ld, jd * PK_TILE, wm, wn, mi_count);
    gemm_mainloop(acc, smem, M, rm, jd * PK_TILE, ld, jd * PK_TILE, PK_TILE, wm, wn, mi_count);
    double* T = smem;                       // 64 x TLD
    double* Lout = T + PK_TILE * TLD;       // 64 x LLD
    double* dsum = Lout + PK_TILE * LLD;    // 64
    double* diag0 = dsum + PK_TILE;         // 64
    {
        const int fr = lane >> 2, fc2 = (lane & 3) * 2;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                if (mi < mi_count) {
                    const int r = wm * 32 + mi * 8 + fr, c = wn * 32 + ni * 8 + fc2;
                    T[r * TLD + c] = acc[mi][ni][0];
                    T[r * TLD + c + 1] = acc[mi][ni][1];
                }
    }
    if (tid < PK_TILE) diag0[tid] = M[(size_t)(jd * PK_TILE + tid) * ld + jd * PK_TILE + tid];
    __syncthreads();
    const int bad = potf2_tile(T, Lout, dsum, diag0, piv_tol);
    if (bad) {
        if (tid == 0) info[cand] = jd * PK_TILE + bad;  // dpotrf convention: order of the failing minor
        return;
    }
    for (int i = tid; i < PK_TILE * PK_TILE; i += THREADS) {
        const int rr = i >> 6, cc = i & 63;
        M[(size_t)(jd * PK_TILE + rr) * ld + jd * PK_TILE + cc] = (cc <= rr) ? Lout[rr * LLD + cc] : 0.0;
    }
}

// first diagonal tile (j = 0): grid = candidates
__global__ void __launch_bounds__(THREADS, 2)
chol_first_kernel(int ld, size_t elems, double piv_tol, double* __restrict__ ws, int32_t* __restrict__ info) {
    extern __shared__ __align__(16) double smem[];
    const int cand = blockIdx.x;
    if (info[cand] != 0) return;
    diag_tile(smem, ws + (size_t)cand * elems, ld, 0, piv_tol, info, cand);
}

// ------------------------------------------------------------------------------------------------
// One launch per block column j.  grid = (1 + row blocks, candidates):
//   blockIdx.x == 0: LOOK-AHEAD role -- the 64 rows of tile (j+1, j) together with the 8 augmented
//       rows (1^T, y^T), then immediately the diagonal tile (j+1, j+1), so the latency-bound
//       unblocked factorisation overlaps the DMMA-bound panel CTAs of the same launch (for the last
//       block column only the augmented rows remain);
//   blockIdx.x >= 1: PANEL role -- 128-row blocks of rows [(j+2)*64, np) and, for pk_fit, of the
//       inverse rows [np+8, rows).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 2)
chol_step_kernel(int j, int nt, int rows, int ld, size_t elems, double piv_tol, double* __restrict__ ws,
                 int32_t* __restrict__ info) {
    extern __shared__ __align__(16) double smem[];
    const int cand = blockIdx.y;
    if (info[cand] != 0) return;
    double* M = ws + (size_t)cand * elems;
    const int np = nt * PK_TILE;
    const bool has_next = (j + 1 < nt);
    RowMap rm;
    int rows_valid;
    bool lookahead = false;
    if (blockIdx.x == 0) {
        if (has_next) {
            rm = RowMap{(j + 1) * PK_TILE, PK_TILE, np};
            rows_valid = PK_TILE + PK_AUG_ROWS;
            lookahead = true;
        } else {
            rm = RowMap{np, BM, 0};
            rows_valid = PK_AUG_ROWS;
        }
    } else {
        const int blk = (int)blockIdx.x - 1;
        const int first = (j + 2) * PK_TILE;                       // rows below the look-ahead tile
        const int nblk1 = (np > first) ? (np - first + BM - 1) / BM : 0;
        int row0;
        if (blk < nblk1) {
            row0 = first + blk * BM;
            rows_valid = min(BM, np - row0);
        } else {
            row0 = np + PK_AUG_ROWS + (blk - nblk1) * BM;          // inverse rows (pk_fit)
            rows_valid = min(BM, rows - row0);
        }
        if (rows_valid <= 0) return;
        rm = RowMap{row0, BM, 0};
    }
    panel_rows(smem, M, ld, j, rm, rows_valid);
    if (lookahead) diag_tile(smem, M, ld, j + 1, piv_tol, info, cand);
}

// ------------------------------------------------------------------------------------------------
// mu, sigma^2, ln|Psi|, NegLnLike from the factorised augmented matrix (matrixops.py:46-56).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nll_epilogue_kernel(int n, int np, int ld, size_t elems, const double* __restrict__ ws, double* __restrict__ out4,
                    int out_stride, const int32_t* __restrict__ info, int* __restrict__ fault) {
    __shared__ double scratch[32];
    const int cand = blockIdx.x, tid = threadIdx.x;
    if (info[cand] != 0) {
        if (tid < 4) out4[tid * out_stride + cand] = CUDART_NAN;
        if (tid == 0 && info[cand] < 0) atomicOr(fault, 1);  // not a numerical outcome: a spin-wait timed out (pk_api.cu check_fault)
        return;
    }
    const double* M = ws + (size_t)cand * elems;
    const double* a = M + (size_t)np * ld;
    const double* d = a + ld;
    double saa = 0.0, sad = 0.0, slog = 0.0;
    for (int c = tid; c < n; c += 256) {
        const double av = a[c], dv = d[c];
        saa += av * av;
        sad += av * dv;
        slog += log(M[(size_t)c * ld + c]);
    }
    saa = block_sum(saa, scratch);
    sad = block_sum(sad, scratch);
    slog = block_sum(slog, scratch);
    const double mu = sad / saa;
    double srr = 0.0;
    for (int c = tid; c < n; c += 256) {
        const double r = d[c] - mu * a[c];
        srr += r * r;
    }
    srr = block_sum(srr, scratch);
    if (tid == 0) {
        const double sigma2 = srr / (double)n;
        const double lndet = 2.0 * slog;
        out4[0 * out_stride + cand] = 0.5 * (double)n * log(sigma2) + 0.5 * lndet;
        out4[1 * out_stride + cand] = mu;
        out4[2 * out_stride + cand] = sigma2;
        out4[3 * out_stride + cand] = lndet;
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_psi_build_generic(pk_handle* h, const TiledShape& s, int C, const double* theta, const double* p,
                                     const double* lambda, int mode, double* ws) {
    const int nt = s.np / PK_TILE;
    dim3 grid(nt * (nt + 1) / 2, C);
    psi_build_kernel<<<grid, 256, 0, h->stream>>>(s.n, h->k, s.np, s.ld, s.elems(), h->dX, theta, p, lambda, mode, ws);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t launch_aug_init(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info) {
    const int aug_elems = (s.rows - s.np) * s.ld;
    dim3 g2(min(64, (aug_elems + 255) / 256), C);
    aug_init_kernel<<<g2, 256, 0, h->stream>>>(s.n, s.np, s.ld, s.rows, s.elems(), h->dy, ws, info);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t launch_cholesky_steps(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info) {
    const int nt = s.np / PK_TILE;
    const size_t smem = sizeof(double) * (size_t)SMEM_DOUBLES;
    static bool attr_done_dev[PK_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(chol_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(chol_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    chol_first_kernel<<<C, THREADS, smem, h->stream>>>(s.ld, s.elems(), h->pivot_tol, ws, info);
    h->launches++;
    for (int j = 0; j < nt; ++j) {
        const int first = (j + 2) * PK_TILE;
        const int nblk1 = (s.np > first) ? (s.np - first + BM - 1) / BM : 0;
        const int inv_rows = s.rows - (s.np + PK_AUG_ROWS);
        const int nblk2 = (inv_rows + BM - 1) / BM;
        dim3 grid(1 + nblk1 + nblk2, C);
        chol_step_kernel<<<grid, THREADS, smem, h->stream>>>(j, nt, s.rows, s.ld, s.elems(), h->pivot_tol, ws, info);
        h->launches++;
    }
    return cudaGetLastError();
}

int cholesky_pick_variant(pk_handle* h, const TiledShape& s, int C, bool have_planes) {
    const bool planes = have_planes;
    if (h->part_variant && !s.with_inverse) return h->part_variant;
    int v = h->chol_variant;
    if (v == 5 && (!planes || s.with_inverse || s.np > 16384)) v = 0;  // INT32 level sums: 7 x 2^14 x K < 2^31  // the INT8 kernel factorises populations only (pk_fit: gangs)
    int64_t Ceff = (h->hint_C > C) ? h->hint_C : (int64_t)C;  // the population this call is a slice of
    if (h->call_C > Ceff) Ceff = h->call_C;                   // ... or a chunk / pipeline part of
    // Automatic: task mode (pk_chol_cta.cu).  Its tiles are produced by the arithmetic of the one-CTA-per-candidate kernel,
    // so a candidate's result does not depend on how many candidates it is evaluated with -- a single SLSQP evaluation,
    // a stencil of 21, a shard of 128 and a population of 1024 agree bit for bit -- and its task granularity follows the
    // size of the call (profiles/task_sweep_r02_*.jsonl: faster than gangs from 1 candidate up at n = 1024, than one CTA
    // per candidate for every population that is not a whole number of waves).
    // One CTA per candidate (bit-identical) is kept where it measured faster: populations of two waves or more whose
    // last wave is full or nearly so -- C = 1024 at n = 1024: 12.3 against 12.9 ms -- and tiny matrices (n <= 256),
    // where a task is too short to pay for its ticket.
    // Populations: the INT8-tensor-core kernel (pk_chol_i8.cu), one CTA per candidate and SM -- a candidate's bits do
    // not depend on the size of the call or of its shard (the rule looks at the hinted population only).  It runs whole
    // waves of sm_count candidates at the latency of ONE factorisation, so it pays when the waves are reasonably full:
    // profiles/i8_sweep_r02.jsonl -- per candidate of a full wave it needs 0.66 (n = 1024) / 0.58 (n = 2048) of task
    // mode's time, at n = 512 the two are level (the FP64 side of a pass does not shrink with K) and task mode stays.
    if (v == 0 && planes && !s.with_inverse && s.np >= 768 && s.np <= 16384) {
        const int64_t waves = (Ceff + h->sm_count - 1) / h->sm_count;
        const double fill = (double)Ceff / (double)(waves * h->sm_count);
        if (fill >= (s.np >= 1536 ? 0.42 : 0.6)) v = 5;
    }
    if (v == 0) {
        const long long slots = 2LL * h->sm_count;
        const bool whole = (C >= 2 * slots && (C % slots == 0 || C % slots >= slots / 3)) || (s.np <= 256 && C >= h->sm_count);
        v = whole ? 2 : 4;
    }
    // pk_fit (one matrix + L^-T rows): the gang kernel, the inverse rows being further panel passes (per-block-column
    // kernels when pinned to variant 1 or when the matrix has fewer than three block columns)
    if (s.with_inverse) v = (h->chol_variant != 1 && cholesky_gang_size(h, s, Ceff, true) > 1) ? 3 : 1;
    return v;
}

cudaError_t launch_cholesky(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info, int8_t* planes) {
    const int v = cholesky_pick_variant(h, s, C, planes != nullptr);
    if (v == 5) return launch_cholesky_i8(h, s, C, ws, planes, info);
    if (v == 4) return launch_cholesky_task(h, s, C, ws, info);
    return (v >= 2) ? launch_cholesky_cta(h, s, C, ws, info) : launch_cholesky_steps(h, s, C, ws, info);
}

cudaError_t launch_nll_epilogue(pk_handle* h, const TiledShape& s, int C, const double* ws, double* out4,
                                int out_stride, int32_t* info) {
    nll_epilogue_kernel<<<C, 256, 0, h->stream>>>(s.n, s.np, s.ld, s.elems(), ws, out4, out_stride, info, h->d_fault);
    h->launches++;
    return cudaGetLastError();
}

}  // namespace pk
