// Prediction path: fused psi-vector / predictor / RMSE / expected-improvement kernel over large
// query sets, plus the small helpers that turn a factorised augmented matrix into the cached
// predictor state.
//
// Replaces the per-point Python loops over matrixops.predict_normalized (matrixops.py:68-77),
// matrixops.predicterr_normalized (matrixops.py:79-95) and kriging.expimp / weightedexpimp
// (krige.py:201-234): for every query point x
//     psi_r = exp(-sum_d theta_d |X_rd - x_d|^p_d)            (never written to HBM)
//     yhat  = mu + psi . w,           w = Psi^-1 (y - mu 1)   (query independent, cached)
//     v     = L^-1 psi  (FP64 tensor-pipe GEMM against the cached L^-1),  q = |v|^2 = psi^T Psi^-1 psi
//     s     = sqrt(|SigmaSqr (1 + extra - q)|),  EI(yhat, s, y_min)
// HBM traffic per query: 8k bytes in, 8..24 bytes out; L^-1 (n^2/2 doubles) streams from L2.
#include "pk_internal.h"
#include "pk_math.cuh"

namespace pk {

// ---- fit helpers ---------------------------------------------------------------------------------
// M = factorised augmented matrix with inverse rows (TiledShape with_inverse).
__global__ void fit_finalize_kernel(int np, int ld, const double* __restrict__ M, double* __restrict__ linv,
                                    double* __restrict__ ad) {
    // linv[c][r] = M[(np + 8 + r) * ld + c]  (transpose of the inverse rows), tiled through smem
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;  // bx: c block, by: r block
    const int tx = threadIdx.x, ty = threadIdx.y;          // 32 x 8
    for (int i = ty; i < 32; i += 8) tile[i][tx] = M[(size_t)(np + PK_AUG_ROWS + by + i) * ld + bx + tx];
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = bx + i, r = by + tx;
        linv[(size_t)c * np + r] = (r <= c) ? tile[tx][i] : 0.0;
    }
    if (blockIdx.y == 0 && ty == 0) {
        ad[bx + tx] = M[(size_t)np * ld + bx + tx];
        ad[np + bx + tx] = M[(size_t)(np + 1) * ld + bx + tx];
    }
}

// w_r = sum_c R[r][c] (d_c - mu a_c), R = inverse rows of M (R = L^-T); one warp per row.
__global__ void compute_w_kernel(int np, int ld, const double* __restrict__ M, const double* __restrict__ ad,
                                 double mu, double* __restrict__ w) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= np) return;
    const double* row = M + (size_t)(np + PK_AUG_ROWS + r) * ld;
    double acc = 0.0;
    for (int c = r - (r & 31) + lane; c < np; c += 32)  // R[r][c] == 0 for c < r
        if (c >= r) acc += row[c] * (ad[np + c] - mu * ad[c]);
    acc = warp_sum(acc);
    if (lane == 0) w[r] = acc;
}

__global__ void extract_U_kernel(int n, int ld, const double* __restrict__ M, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < n) out[(size_t)r * n + c] = (c >= r) ? M[(size_t)c * ld + r] : 0.0;  // U = L^T
}

// full symmetric n x n Psi from the lower-triangle tiles of the tiled layout (blockIdx.z = candidate)
__global__ void extract_psi_kernel(int n, int ld, size_t elems, const double* __restrict__ ws,
                                   double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    const double* M = ws + (size_t)blockIdx.z * elems;
    double* o = out + (size_t)blockIdx.z * n * n;
    if (c < n) o[(size_t)r * n + c] = (c <= r) ? M[(size_t)r * ld + c] : M[(size_t)c * ld + r];
}

// ---- incremental append of one sample to a cached factor (pk_append_point) ---------------------------
// Replaces the full rebuild kriging.addPoint triggers (krige.py:135-156 -> updateData -> updatePsi:
// O(n^2 k) pow + n^3/3 flop) by the bordered update of L, L^-1, a = L^-1 1 and d = L^-1 y:
//     psi = corr(X, x_new);  l = L^-1 psi;  t = sqrt(diag - |l|^2)   (diag = 1 + eps | 1 + lambda)
//     L <- [L 0; l^T t],  L^-1 <- [L^-1 0; -(l^T L^-1)/t  1/t],  a_n = (1 - l.a)/t,  d_n = (y_new - l.d)/t
// O(n^2) flop and bytes.  The padded order np is unchanged (the caller refits when n crosses a tile boundary).
__global__ void append_psi_kernel(int n, int k, const double* __restrict__ X, const double* __restrict__ xnew,
                                  const double* __restrict__ hp, double* __restrict__ psi) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) psi[r] = exp(-corr_exponent(X + (size_t)r * k, xnew, hp, hp + PK_MAX_K, k));
}

// l[c] = sum_{r <= c} Linv[c][r] psi[r], one warp per row c (Linv row-major, lower)
__global__ void append_solve_kernel(int n, int np, const double* __restrict__ linv, const double* __restrict__ psi,
                                    double* __restrict__ l) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= n) return;
    const double* row = linv + (size_t)c * np;
    double acc = 0.0;
    for (int r = lane; r <= c; r += 32) acc = fma(row[r], psi[r], acc);
    acc = warp_sum(acc);
    if (lane == 0) l[c] = acc;
}

// one CTA: pivot test, then row n of L, the new entries of a and d.  info[0] = n + 1 on failure (dpotrf's
// "leading minor of order n+1"), in which case nothing is modified.
__global__ void append_pivot_kernel(int n, int np, int ld, double diag, double piv_tol, double y_new,
                                    const double* __restrict__ l, double* __restrict__ M, double* __restrict__ ad,
                                    double* __restrict__ scal, int32_t* __restrict__ info) {
    __shared__ double scratch[32];
    const int tid = threadIdx.x;
    double sll = 0.0, sla = 0.0, sld = 0.0;
    for (int c = tid; c < n; c += blockDim.x) {
        const double lc = l[c];
        sll = fma(lc, lc, sll);
        sla = fma(lc, ad[c], sla);
        sld = fma(lc, ad[np + c], sld);
    }
    sll = block_sum(sll, scratch);
    sla = block_sum(sla, scratch);
    sld = block_sum(sld, scratch);
    const double piv = diag - sll;
    if (!(piv > piv_tol * diag)) {
        if (tid == 0) info[0] = n + 1;
        return;
    }
    const double t = sqrt(piv);
    for (int c = tid; c < n; c += blockDim.x) M[(size_t)n * ld + c] = l[c];
    if (tid == 0) {
        info[0] = 0;
        M[(size_t)n * ld + n] = t;
        const double an = (1.0 - sla) / t, dn = (y_new - sld) / t;
        M[(size_t)np * ld + n] = an;
        M[(size_t)(np + 1) * ld + n] = dn;
        ad[n] = an;
        ad[np + n] = dn;
        scal[0] = t;
    }
}

// new row n of L^-1: Linv[n][r] = -(1/t) sum_{c >= r} l[c] Linv[c][r], Linv[n][n] = 1/t; mirrored into the
// inverse rows of M (R = L^-T: R[r][n]).  One thread per column r (coalesced over r for every c).
__global__ void append_inverse_kernel(int n, int np, int ld, const double* __restrict__ l, double* __restrict__ linv,
                                      double* __restrict__ M, const double* __restrict__ scal,
                                      const int32_t* __restrict__ info) {
    if (info[0] != 0) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    const double t = scal[0];
    double v;
    if (r == n) {
        v = 1.0 / t;
    } else {
        double acc = 0.0;
        for (int c = r; c < n; ++c) acc = fma(l[c], linv[(size_t)c * np + r], acc);
        v = -acc / t;
    }
    linv[(size_t)n * np + r] = v;
    M[(size_t)(np + PK_AUG_ROWS + r) * ld + n] = v;
}

// Psi copy (dense, row-major) grown from n x n to (n+1) x (n+1): border = psi_new, corner = diag
__global__ void append_psi_copy_kernel(int n, double diag, const double* __restrict__ old, const double* __restrict__ psi,
                                       double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c > n) return;
    double v;
    if (r < n && c < n) v = old[(size_t)r * n + c];
    else if (r == n && c == n) v = diag;
    else v = psi[(r == n) ? c : r];
    out[(size_t)r * (n + 1) + c] = v;
}

cudaError_t launch_append_point(pk_handle* h, const TiledShape& s, const double* xnew_dev, double y_new, double diag,
                                double* scratch /* >= 2 np + 8 doubles */, int32_t* info) {
    const int n = s.n, np = s.np, k = h->k;
    double* psi = scratch;
    double* l = scratch + np;
    double* scal = scratch + 2 * np;
    append_psi_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(n, k, h->dX, xnew_dev, h->d_fit_hp, psi);
    append_solve_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(n, np, h->d_linv, psi, l);
    append_pivot_kernel<<<1, 256, 0, h->stream>>>(n, np, s.ld, diag, h->pivot_tol, y_new, l, h->d_fit, h->d_ad, scal, info);
    append_inverse_kernel<<<(n + 256) / 256, 256, 0, h->stream>>>(n, np, s.ld, l, h->d_linv, h->d_fit, scal, info);
    h->launches += 4;
    return cudaGetLastError();
}

cudaError_t launch_append_psi_copy(pk_handle* h, int n, double diag, const double* old, const double* psi, double* out) {
    dim3 grid((n + 256) / 256, n + 1);
    append_psi_copy_kernel<<<grid, 256, 0, h->stream>>>(n, diag, old, psi, out);
    h->launches++;
    return cudaGetLastError();
}

// ---- fused predictor -------------------------------------------------------------------------------
constexpr int PKC = 16;           // K chunk of the L^-1 pipeline
constexpr int PSLD = PKC + 4;     // 20
constexpr int PSTAGES = 3;
constexpr int PRED_THREADS = 256;

template <int QT>
__global__ void __launch_bounds__(PRED_THREADS)
predict_kernel(int n, int np, int k, int64_t m, const double* __restrict__ X, const double* __restrict__ Xq,
               const double* __restrict__ hp, const double* __restrict__ linv, const double* __restrict__ w,
               double mu, double sigma2, double y_min, double ei_weight, double var_extra, int need_s,
               double* __restrict__ yhat_out, double* __restrict__ s_out, double* __restrict__ ei_out) {
    extern __shared__ __align__(16) double smem[];
    const int psl = np + 4;                              // psi tile row stride (== 4 mod 16)
    double* psi = smem;                                  // QT x psl
    double* Bs = psi + QT * psl;                         // PSTAGES x 64 x PSLD
    double* xq = Bs + PSTAGES * PK_TILE * PSLD;          // QT x k
    double* th = xq + QT * k;                            // k
    double* pp = th + k;                                 // k
    double* ysum = pp + k;                               // QT
    double* qsum = ysum + QT;                            // QT x 8 (per n-warp-group partials)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * QT;
    const int qv = (int)min((int64_t)QT, m - q0);

    for (int i = tid; i < QT * k; i += PRED_THREADS) {
        const int q = i / k;
        xq[i] = (q < qv) ? Xq[q0 * k + i] : 0.0;
    }
    if (tid < k) {
        th[tid] = hp[tid];
        pp[tid] = hp[PK_MAX_K + tid];
    }
    __syncthreads();

    // ---- psi tile + yhat partials: warp `warp` handles queries warp, warp+8, ... ------------------
    for (int q = warp; q < QT; q += PRED_THREADS / 32) {
        double ys = 0.0;
        for (int r = lane; r < np; r += 32) {
            double v = 0.0;
            if (r < n && q < qv) {
                v = exp(-corr_exponent(X + (size_t)r * k, xq + q * k, th, pp, k));
                ys += v * w[r];
            }
            psi[q * psl + r] = v;
        }
        ys = warp_sum(ys);
        if (lane == 0) ysum[q] = ys;
    }
    __syncthreads();

    double qpart[2] = {0.0, 0.0};
    if (need_s) {
        // ---- v = psi * Linv^T, column tile by column tile; accumulate |v|^2 per query ---------------
        constexpr int MT = QT / 8;                       // m-subtiles (8 queries each)
        constexpr int WM = (MT >= 2) ? 2 : 1;            // m-subtiles per warp
        constexpr int MW = MT / WM;                      // warps along m
        constexpr int NW = (PRED_THREADS / 32) / MW;     // warps along n
        constexpr int WN = 8 / NW;                       // n-subtiles (8 cols) per warp
        static_assert(MW * NW == PRED_THREADS / 32 && WN >= 1, "bad predict tiling");
        const int wm = warp % MW, wn = warp / MW;
        const int fr = lane >> 2, fc = lane & 3;
        const int nct = np / PK_TILE;
        for (int ct = 0; ct < nct; ++ct) {
            double acc[WM][WN][2];
#pragma unroll
            for (int a = 0; a < WM; ++a)
#pragma unroll
                for (int b = 0; b < WN; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
            const double* Bg = linv + (size_t)(ct * PK_TILE) * np;
            const int nk = (ct + 1) * PK_TILE / PKC;     // Linv[c][r] == 0 for r > c
            auto load_stage = [&](int s, int k0) {
                for (int i = tid; i < PK_TILE * 8; i += PRED_THREADS) {
                    const int r = i >> 3, ch = i & 7;
                    cp_async16(Bs + s * PK_TILE * PSLD + r * PSLD + ch * 2, Bg + (size_t)r * np + k0 + ch * 2);
                }
            };
#pragma unroll
            for (int s = 0; s < PSTAGES - 1; ++s) {
                if (s < nk) load_stage(s, s * PKC);
                cp_async_commit();
            }
            for (int it = 0; it < nk; ++it) {
                cp_async_wait<PSTAGES - 2>();
                __syncthreads();
                const int nxt = it + PSTAGES - 1;
                if (nxt < nk) load_stage(nxt % PSTAGES, nxt * PKC);
                cp_async_commit();
                const double* a_s = psi + (wm * WM * 8 + fr) * psl + it * PKC + fc;
                const double* b_s = Bs + (it % PSTAGES) * PK_TILE * PSLD + (wn * WN * 8 + fr) * PSLD + fc;
#pragma unroll
                for (int kk = 0; kk < PKC; kk += 4) {
                    double a[WM], b[WN];
#pragma unroll
                    for (int x = 0; x < WM; ++x) a[x] = a_s[x * 8 * psl + kk];
#pragma unroll
                    for (int x = 0; x < WN; ++x) b[x] = b_s[x * 8 * PSLD + kk];
#pragma unroll
                    for (int x = 0; x < WM; ++x)
#pragma unroll
                        for (int z = 0; z < WN; ++z) dmma884(acc[x][z][0], acc[x][z][1], a[x], b[z]);
                }
            }
            cp_async_wait<0>();
            __syncthreads();
#pragma unroll
            for (int x = 0; x < WM; ++x)
#pragma unroll
                for (int z = 0; z < WN; ++z) qpart[x] += acc[x][z][0] * acc[x][z][0] + acc[x][z][1] * acc[x][z][1];
        }
        // reduce over the 4 lanes sharing a row, then over the n-warps through smem
#pragma unroll
        for (int x = 0; x < WM; ++x) {
            double v = qpart[x];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (fc == 0) qsum[(wm * WM * 8 + x * 8 + fr) * 8 + wn] = v;
        }
        static_assert(NW <= 8, "qsum holds 8 partials per query");
        __syncthreads();
        if (tid < QT) {
            double v = 0.0;
            for (int z = 0; z < NW; ++z) v += qsum[tid * 8 + z];
            qsum[tid * 8] = v;
        }
        __syncthreads();
    }

    if (tid < qv) {
        const double yhat = mu + ysum[tid];
        if (yhat_out) yhat_out[q0 + tid] = yhat;
        if (need_s) {
            const double ssqr = sigma2 * (1.0 + var_extra - qsum[tid * 8]);
            const double S = sqrt(fabs(ssqr));
            if (s_out) s_out[q0 + tid] = S;
            if (ei_out) ei_out[q0 + tid] = expected_improvement(yhat, S, y_min, ei_weight);
        }
    }
}


// ---- fused predictor, any n (np > 3072: a psi tile no longer fits in shared memory) --------------------
// Same arithmetic as predict_kernel<32>, but the psi tile of a CTA (32 queries x np) lives in a per-CTA slab of a
// global scratch buffer (L2 resident: 32 x 4100 doubles = 1 MB per CTA) and its 32 x 16 chunks ride through the
// cp.async stages together with the 64 x 16 chunks of L^-1.  Persistent: a CTA walks query tiles with stride
// gridDim.x, so the scratch is grid-sized, not m-sized.  Config C4 (regression kriging, n = 4096).
constexpr int PB_QT = 32;
__global__ void __launch_bounds__(PRED_THREADS)
predict_big_kernel(int n, int np, int k, int64_t m, const double* __restrict__ X, const double* __restrict__ Xq,
                   const double* __restrict__ hp, const double* __restrict__ linv, const double* __restrict__ w,
                   double mu, double sigma2, double y_min, double ei_weight, double var_extra, int need_s,
                   double* __restrict__ psi_scratch, double* __restrict__ yhat_out, double* __restrict__ s_out,
                   double* __restrict__ ei_out) {
    extern __shared__ __align__(16) double smem[];
    constexpr int QT = PB_QT;
    const int psl = np + 4;
    double* As = smem;                                   // PSTAGES x QT x PSLD
    double* Bs = As + PSTAGES * QT * PSLD;               // PSTAGES x 64 x PSLD
    double* xq = Bs + PSTAGES * PK_TILE * PSLD;          // QT x k
    double* th = xq + QT * k;                            // k
    double* pp = th + k;                                 // k
    double* ysum = pp + k;                               // QT
    double* qsum = ysum + QT;                            // QT x 8
    double* psi = psi_scratch + (size_t)blockIdx.x * QT * psl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < k) {
        th[tid] = hp[tid];
        pp[tid] = hp[PK_MAX_K + tid];
    }
    const int64_t ntiles = (m + QT - 1) / QT;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t q0 = tile * QT;
        const int qv = (int)min((int64_t)QT, m - q0);
        __syncthreads();  // previous tile: outputs written, xq / ysum / qsum / stages free
        for (int i = tid; i < QT * k; i += PRED_THREADS) {
            const int q = i / k;
            xq[i] = (q < qv) ? Xq[q0 * k + i] : 0.0;
        }
        __syncthreads();
        for (int q = warp; q < QT; q += PRED_THREADS / 32) {
            double ys = 0.0;
            for (int r = lane; r < np; r += 32) {
                double v = 0.0;
                if (r < n && q < qv) {
                    v = exp(-corr_exponent(X + (size_t)r * k, xq + q * k, th, pp, k));
                    ys += v * w[r];
                }
                psi[(size_t)q * psl + r] = v;
            }
            ys = warp_sum(ys);
            if (lane == 0) ysum[q] = ys;
        }
        __syncthreads();  // the slab (global memory, written by this CTA) is complete for every thread of the CTA

        if (need_s) {
            // warps: 2 along the queries (16 each) x 4 along the 64 columns of a tile (16 each)
            const int wm = warp & 1, wn = warp >> 1;
            const int fr = lane >> 2, fc = lane & 3;
            const int nct = np / PK_TILE;
            double qpart[2] = {0.0, 0.0};
            for (int ct = 0; ct < nct; ++ct) {
                double acc[2][2][2];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
                const double* Bg = linv + (size_t)(ct * PK_TILE) * np;
                const int nk = (ct + 1) * PK_TILE / PKC;  // Linv[c][r] == 0 for r > c
                auto load_stage = [&](int s, int k0) {
                    for (int i = tid; i < (PK_TILE + QT) * 8; i += PRED_THREADS) {
                        const int r = i >> 3, ch = i & 7;
                        if (r < PK_TILE)
                            cp_async16(Bs + s * PK_TILE * PSLD + r * PSLD + ch * 2, Bg + (size_t)r * np + k0 + ch * 2);
                        else
                            cp_async16(As + s * QT * PSLD + (r - PK_TILE) * PSLD + ch * 2,
                                       psi + (size_t)(r - PK_TILE) * psl + k0 + ch * 2);
                    }
                };
#pragma unroll
                for (int s = 0; s < PSTAGES - 1; ++s) {
                    if (s < nk) load_stage(s, s * PKC);
                    cp_async_commit();
                }
                for (int it = 0; it < nk; ++it) {
                    cp_async_wait<PSTAGES - 2>();
                    __syncthreads();
                    const int nxt = it + PSTAGES - 1;
                    if (nxt < nk) load_stage(nxt % PSTAGES, nxt * PKC);
                    cp_async_commit();
                    const double* a_s = As + (it % PSTAGES) * QT * PSLD + (wm * 16 + fr) * PSLD + fc;
                    const double* b_s = Bs + (it % PSTAGES) * PK_TILE * PSLD + (wn * 16 + fr) * PSLD + fc;
#pragma unroll
                    for (int kk = 0; kk < PKC; kk += 4) {
                        const double a0 = a_s[kk], a1 = a_s[8 * PSLD + kk];
                        const double b0 = b_s[kk], b1 = b_s[8 * PSLD + kk];
                        dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
                        dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
                        dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
                        dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
                    }
                }
                cp_async_wait<0>();
                __syncthreads();
#pragma unroll
                for (int x = 0; x < 2; ++x)
#pragma unroll
                    for (int z = 0; z < 2; ++z)
                        qpart[x] += acc[x][z][0] * acc[x][z][0] + acc[x][z][1] * acc[x][z][1];
            }
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                double v = qpart[x];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (fc == 0) qsum[(wm * 16 + x * 8 + fr) * 8 + wn] = v;
            }
            __syncthreads();
        }
        if (tid < qv) {
            const double yhat = mu + ysum[tid];
            if (yhat_out) yhat_out[q0 + tid] = yhat;
            if (need_s) {
                const double qs = (qsum[tid * 8] + qsum[tid * 8 + 1]) + (qsum[tid * 8 + 2] + qsum[tid * 8 + 3]);
                const double ssqr = sigma2 * (1.0 + var_extra - qs);
                const double S = sqrt(fabs(ssqr));
                if (s_out) s_out[q0 + tid] = S;
                if (ei_out) ei_out[q0 + tid] = expected_improvement(yhat, S, y_min, ei_weight);
            }
        }
    }
}

// ---- fused predictor, register-resident variant (n <= 512) ------------------------------------------
// v = L^-1 psi for a tile of 32 query points is held ENTIRELY in accumulator registers: the 16 warps of a
// CTA split the columns of v in 8-column blocks, block b -> warp b % 16 (so every warp owns a slice of every
// 128 columns and the triangular saving of L^-1 is spread evenly), 4 x NTW DMMA tiles per warp.  The K loop
// runs over the samples r in chunks of 16: the psi chunk (32 queries x 16 samples, one value per thread) is
// computed on the fly into a double-buffered shared-memory A stage -- psi never exists as a whole tile --
// and the B operand L^-1[c, r] is read straight from L2 into registers (each element is used by exactly one
// warp, so staging it through shared memory would buy nothing), software-prefetched one k-step ahead.
// Column blocks left of the chunk are skipped by dispatching on the first live block through a
// warp-uniform switch (a predicated-off DMMA would still occupy the tensor pipe).  Four warps per
// scheduler: while one warp walks the latency-bound psi chain the others issue DMMAs.
constexpr int PR_THREADS = 512;
constexpr int PR_WARPS = PR_THREADS / 32;
constexpr int PR_QT = 32;
constexpr int PR_KC = 16;
constexpr int PR_ALD = PR_KC + 4;

// One 16-sample chunk of v += psi_chunk * Linv_chunk^T for this warp's column blocks LO .. NTW-1.  The B
// operand comes from the warp's private shared-memory stage: block ni = 8 rows x 16 k (1 KB), 16-byte pieces
// XOR-swizzled with the row so that the fragment loads are conflict free without padding
// (boff[ks] = (fc & 1) + 2 * ((2 ks) ^ (fc >> 1) ^ fr), precomputed per lane).
template <int NTW, int LO>
__device__ __forceinline__ void pred_consume(double (&acc)[4][NTW][2], const double* As, const double* Bw,
                                             const int (&boff)[4], int lane) {
    const int fr = lane >> 2, fc = lane & 3;
    const double* a_s = As + fr * PR_ALD + fc;
    const double* b_s = Bw + fr * 16;
#pragma unroll
    for (int ks = 0; ks < PR_KC / 4; ++ks) {
        double a[4], b[NTW];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = a_s[mi * 8 * PR_ALD + 4 * ks];
#pragma unroll
        for (int ni = LO; ni < NTW; ++ni) b[ni] = b_s[ni * 128 + boff[ks]];
#pragma unroll
        for (int ni = LO; ni < NTW; ++ni) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
}

constexpr int PR_XS_MAX = 3072;  // sample coordinates staged in shared memory when n * k fits (24 KB)

// KK > 0: number of design variables known at compile time (theta, p, the query point in registers, a
// straight-line psi chain); KK == 0: any k <= PK_MAX_K.
template <int NTW, int KK>
__global__ void __launch_bounds__(PR_THREADS, 1)
predict_reg_kernel(int n, int np, int k, int64_t m, const double* __restrict__ X, const double* __restrict__ Xq,
                   const double* __restrict__ hp, const double* __restrict__ linv, const double* __restrict__ w,
                   double mu, double sigma2, double y_min, double ei_weight, double var_extra, int need_s,
                   double* __restrict__ yhat_out, double* __restrict__ s_out, double* __restrict__ ei_out) {
    // PERSISTENT: one CTA per SM walks query tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The tables, the staged
    // sample coordinates and the hyper-parameters are set up once per CTA, and the chunk pipeline runs straight
    // through tile boundaries: psi chunk 0 / the first L^-1 chunk of the NEXT tile are produced under the last
    // DMMA chunk of the current one, and a tile's final reduction (yhat, s, EI: 32 threads) happens after the next
    // tile's first barrier -- no prologue, epilogue or pipeline fill is exposed between tiles.
    extern __shared__ __align__(16) double pr_smem[];
    double* tab = pr_smem;                                          // 64 x 16 : 2^(j/64), bank replicated
    double* ltab = tab + 64 * 16;                                   // 3 x 64 x 16 : log2 split tables
    double (*As)[PR_QT * PR_ALD] = reinterpret_cast<double (*)[PR_QT * PR_ALD]>(ltab + 3 * 64 * 16);
    const int xst = k;                                              // row stride of a staged query point
    double* xqs0 = reinterpret_cast<double*>(As + 2);               // 2 (tile parity) x 32 x k
    double* xqs[2] = {xqs0, xqs0 + PR_QT * xst};
    double2* thp = reinterpret_cast<double2*>(xqs0 + 2 * PR_QT * ((xst + 1) & ~1));  // PK_MAX_K
    double (*ypart)[PR_WARPS][PR_QT] = reinterpret_cast<double (*)[PR_WARPS][PR_QT]>(thp + PK_MAX_K);  // per tile parity
    double (*qpart)[PR_WARPS][PR_QT + 1] = reinterpret_cast<double (*)[PR_WARPS][PR_QT + 1]>(ypart + 2);
    // B stages: 2 x (16 warps x NTW blocks x 128 doubles), every warp owns its slice
    double* Bst = reinterpret_cast<double*>(qpart + 2);
    Bst += (16 - ((size_t)(Bst - pr_smem) & 15)) & 15;              // 128-byte aligned rows
    double* Xs = Bst + 2 * PR_WARPS * NTW * 128;                    // n x k when it fits
    const bool x_staged = (n * k <= PR_XS_MAX);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int64_t ntiles = (m + PR_QT - 1) / PR_QT;

    fill_exp2_table(tab, tid, PR_THREADS);
    fill_log2_tables(ltab, tid, PR_THREADS);
    if (tid < k) thp[tid] = make_double2(hp[tid], hp[PK_MAX_K + tid]);
    if (x_staged)
        for (int i = tid; i < n * k; i += PR_THREADS) Xs[i] = X[i];
    // query points of a tile into the shared buffer of its parity (generic-k path and the final reduction's qv)
    auto stage_queries = [&](int64_t tile, int par) {
        const int64_t q0 = tile * PR_QT;
        const int qv = (int)min((int64_t)PR_QT, m - q0);
        for (int i = tid; i < PR_QT * k; i += PR_THREADS) {
            const int q = i / k;
            xqs[par][i] = (q < qv) ? Xq[q0 * k + i] : 0.0;
        }
    };
    int64_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    stage_queries(tile, 0);
    __syncthreads();
    const char* tabp = reinterpret_cast<const char*>(tab + (tid & 15));
    const double* Xsrc = x_staged ? Xs : X;
    // compile-time k: hyper-parameters and this thread's query point live in registers
    double th_r[KK > 0 ? KK : 1], p_r[KK > 0 ? KK : 1], xq_r[KK > 0 ? KK : 1], xq_n[KK > 0 ? KK : 1];
    bool special = false;  // some exponent is exactly 1 or 2 (exact d / d*d path, CTA uniform)
    if (KK > 0) {
#pragma unroll
        for (int d = 0; d < KK; ++d) {
            th_r[d] = thp[d].x;
            p_r[d] = thp[d].y;
            xq_r[d] = xqs[0][lane * xst + d];
            xq_n[d] = 0.0;
            special |= (p_r[d] == 1.0 || p_r[d] == 2.0);
        }
    }

    // psi producer: this thread owns query `lane` and sample slot `warp` of every chunk
    auto make_psi = [&](int kc, double* dst, const double (&xr_q)[KK > 0 ? KK : 1], const double* xqq, int qv,
                        double& ys) {
        const int r = kc * PR_KC + warp;
        double v = 0.0;
        if (r < n && lane < qv) {
            const double* xr = Xsrc + (size_t)r * k;
            double S;
            if (KK > 0 && !special) {
                // straight line: KK x (log2 split + 2^x), NumPy's summation order (plain left to right for k < 8)
                double term[KK > 0 ? KK : 1];
#pragma unroll
                for (int d = 0; d < KK; ++d) {
                    double lh;
                    float ll;
                    log2_split_smem(fabs(xr[d] - xr_q[d]), ltab, tid & 15, lh, ll);
                    term[d] = th_r[d] * exp2_lean(p_r[d], lh, (double)ll, tabp);
                }
                S = sum_numpy_order(KK, [&](int d) { return term[d]; });
            } else {
                S = sum_numpy_order(k, [&](int d) {
                    const double2 tp = thp[d];
                    return tp.x * pow_abs_lean(fabs(xr[d] - xqq[d]), tp.y, tabp, ltab, tid & 15);
                });
            }
            v = expneg_table(S, tab, tid & 15);
            ys = fma(v, __ldg(w + r), ys);
        }
        dst[lane * PR_ALD + warp] = v;
    };

    const int nch = np / PR_KC;
    double acc[4][NTW][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < NTW; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    // L^-1 operand: rows (columns of v) 128 ni + 8 warp .. + 7, 16 samples per chunk, streamed from L2 into this
    // warp's shared stage one whole chunk ahead (cp.async, 8 pieces of 16 bytes per lane and chunk): the
    // L2 latency hides behind a chunk of DMMAs instead of a single k-step.  Blocks left of the chunk (zero part
    // of the triangular L^-1) are neither loaded nor multiplied: the first live block is dispatched through a
    // warp-uniform switch (a predicated-off DMMA would still occupy the tensor pipe).
    unsigned valid_mask = 0;
#pragma unroll
    for (int ni = 0; ni < NTW; ++ni)
        if (8 * PR_WARPS * ni + 8 * warp < np) valid_mask |= 1u << ni;
    double* Bw = Bst + warp * NTW * 128;
    constexpr int BSTAGE = PR_WARPS * NTW * 128;
    int boff[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) boff[ks] = (fc & 1) + 2 * ((2 * ks) ^ (fc >> 1) ^ fr);
    auto first_live = [&](int kc) {
        // block ni covers columns 128 ni + 8 warp .. + 7 and is needed iff that reaches 16 kc
        const int num = PR_KC * kc - 8 * warp - 7;
        return (num <= 0) ? 0 : ((num + 8 * PR_WARPS - 1) / (8 * PR_WARPS));
    };
    auto load_B = [&](int kc, int stage) {
        const int lo = first_live(kc);
        double* dst = Bw + stage * BSTAGE;
        const double* src = linv + (size_t)(8 * warp) * np + PR_KC * kc;
#pragma unroll
        for (int i = 0; i < 2 * NTW; ++i) {
            const int idx = i * 32 + lane, ni = idx >> 6, row = (idx >> 3) & 7, piece = idx & 7;
            if (ni >= lo && ((valid_mask >> ni) & 1u))
                cp_async16(dst + ni * 128 + row * 16 + 2 * (piece ^ row),
                           src + (size_t)(8 * PR_WARPS * ni + row) * np + 2 * piece);
        }
    };
    // final reduction of a finished tile (partials of parity `par`): yhat partials over the 16 sample slots, |v|^2
    // over the column blocks, then s and EI.  Runs on the first 32 threads after a CTA barrier.
    auto finish_tile = [&](int64_t t, int par) {
        const int64_t q0 = t * PR_QT;
        const int qv = (int)min((int64_t)PR_QT, m - q0);
        if (tid < qv) {
            double ysum = 0.0, qs = 0.0;
#pragma unroll
            for (int z = 0; z < PR_WARPS; ++z) ysum += ypart[par][z][tid];
            const double yhat = mu + ysum;
            if (yhat_out) yhat_out[q0 + tid] = yhat;
            if (need_s) {
#pragma unroll
                for (int z = 0; z < PR_WARPS; ++z) qs += qpart[par][z][tid];
                const double ssqr = sigma2 * (1.0 + var_extra - qs);
                const double S = sqrt(fabs(ssqr));
                if (s_out) s_out[q0 + tid] = S;
                if (ei_out) ei_out[q0 + tid] = expected_improvement(yhat, S, y_min, ei_weight);
            }
        }
    };
    if (need_s) {
        // blocks beyond the padded order are never loaded: they must read as zeros in both stages
#pragma unroll
        for (int ni = 0; ni < NTW; ++ni)
            if (!((valid_mask >> ni) & 1u))
                for (int i = lane; i < 128; i += 32) Bw[ni * 128 + i] = Bw[BSTAGE + ni * 128 + i] = 0.0;
        __syncwarp();
        load_B(0, 0);
    }
    cp_async_commit();
    double ys = 0.0, ys_n = 0.0;
    int qv = (int)min((int64_t)PR_QT, m - tile * PR_QT);
    make_psi(0, As[0], xq_r, xqs[0] + lane * xst, qv, ys);
    unsigned g = 0;    // chunks consumed so far by this CTA: stage of chunk g = g & 1
    int tpar = 0;      // parity of the current tile (query buffer, partials)
    int64_t done_tile = -1;  // tile whose partials wait for their final reduction
    for (;;) {
        const int64_t next_tile = tile + gridDim.x;
        const bool has_next = next_tile < ntiles;
        int qv_n = 0;
        if (has_next) {
            // the next tile's query points: shared buffer of the other parity (visible after this tile's barriers)
            // and, for compile-time k, straight into registers
            stage_queries(next_tile, tpar ^ 1);
            qv_n = (int)min((int64_t)PR_QT, m - next_tile * PR_QT);
            if (KK > 0) {
#pragma unroll
                for (int d = 0; d < KK; ++d)
                    xq_n[d] = (lane < qv_n) ? __ldg(Xq + (next_tile * PR_QT + lane) * KK + d) : 0.0;
            }
        }
        for (int kc = 0; kc < nch; ++kc, ++g) {
            const bool last = (kc + 1 == nch);
            const int st = (int)(g & 1u);
            if (need_s && (!last || has_next)) load_B(last ? 0 : kc + 1, st ^ 1);  // that stage was consumed by this warp in chunk g - 1
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();  // chunk g of psi is complete (and every lane's B pieces have landed); the other A stage is free
            if (done_tile >= 0) {  // partials of the previous tile are complete as of this barrier
                finish_tile(done_tile, tpar ^ 1);
                done_tile = -1;
            }
            if (!last) make_psi(kc + 1, As[st ^ 1], xq_r, xqs[tpar] + lane * xst, qv, ys);
            else if (has_next) make_psi(0, As[st ^ 1], xq_n, xqs[tpar ^ 1] + lane * xst, qv_n, ys_n);
            if (need_s) {
                const double* as = As[st];
                const double* bw = Bw + st * BSTAGE;
                switch (first_live(kc)) {
                    case 0: pred_consume<NTW, 0>(acc, as, bw, boff, lane); break;
                    case 1:
                        if (NTW > 1) pred_consume<NTW, (NTW > 1 ? 1 : 0)>(acc, as, bw, boff, lane);
                        break;
                    case 2:
                        if (NTW > 2) pred_consume<NTW, (NTW > 2 ? 2 : 0)>(acc, as, bw, boff, lane);
                        break;
                    case 3:
                        if (NTW > 3) pred_consume<NTW, (NTW > 3 ? 3 : 0)>(acc, as, bw, boff, lane);
                        break;
                    default: break;
                }
                __syncwarp();  // all lanes are done with this stage before the next load_B overwrites it
            }
        }
        // ---- partials of this tile; the final reduction follows the next CTA barrier -------------------
        ypart[tpar][warp][lane] = ys;
        if (need_s) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                double v = 0.0;
#pragma unroll
                for (int ni = 0; ni < NTW; ++ni) {
                    v += acc[mi][ni][0] * acc[mi][ni][0] + acc[mi][ni][1] * acc[mi][ni][1];
                    acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                }
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (fc == 0) qpart[tpar][warp][mi * 8 + fr] = v;
            }
        }
        done_tile = tile;
        if (!has_next) break;
        tile = next_tile;
        tpar ^= 1;
        qv = qv_n;
        ys = ys_n;
        ys_n = 0.0;
        if (KK > 0) {
#pragma unroll
            for (int d = 0; d < KK; ++d) xq_r[d] = xq_n[d];
        }
    }
    __syncthreads();
    finish_tile(done_tile, tpar);
}

// ---- launchers -------------------------------------------------------------------------------------
cudaError_t launch_fit_finalize(pk_handle* h, const TiledShape& s) {
    dim3 grid(s.np / 32, s.np / 32), block(32, 8);
    fit_finalize_kernel<<<grid, block, 0, h->stream>>>(s.np, s.ld, h->d_fit, h->d_linv, h->d_ad);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t launch_compute_w(pk_handle* h, const TiledShape& s, double mu) {
    compute_w_kernel<<<(s.np + 7) / 8, 256, 0, h->stream>>>(s.np, s.ld, h->d_fit, h->d_ad, mu, h->d_w);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t launch_extract_U(pk_handle* h, const TiledShape& s, double* out) {
    dim3 grid((s.n + 255) / 256, s.n);
    extract_U_kernel<<<grid, 256, 0, h->stream>>>(s.n, s.ld, h->d_fit, out);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t launch_extract_psi(pk_handle* h, const TiledShape& s, const double* M, double* out, int C) {
    dim3 grid((s.n + 255) / 256, s.n, C);
    extract_psi_kernel<<<grid, 256, 0, h->stream>>>(s.n, s.ld, s.elems(), M, out);
    h->launches++;
    return cudaGetLastError();
}

template <int QT>
static cudaError_t launch_predict_t(pk_handle* h, int64_t m, const double* Xq, const double* hp, double mu,
                                    double sigma2, double y_min, double ei_weight, double var_extra,
                                    double* yhat, double* sout, double* ei) {
    const int np = h->fit_np, k = h->k;
    const size_t smem = sizeof(double) * ((size_t)QT * (np + 4) + PSTAGES * PK_TILE * PSLD + (size_t)QT * k + 2 * k +
                                          QT + QT * 8);
    auto kern = predict_kernel<QT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int need_s = (sout != nullptr || ei != nullptr) ? 1 : 0;
    const int64_t blocks = (m + QT - 1) / QT;
    kern<<<(unsigned)blocks, PRED_THREADS, smem, h->stream>>>(h->n, np, k, m, h->dX, Xq, hp, h->d_linv, h->d_w, mu,
                                                              sigma2, y_min, ei_weight, var_extra, need_s, yhat,
                                                              sout, ei);
    h->launches++;
    return cudaGetLastError();
}

static cudaError_t launch_predict_big(pk_handle* h, int64_t m, const double* Xq, const double* hp, double mu,
                                      double sigma2, double y_min, double ei_weight, double var_extra, double* yhat,
                                      double* sout, double* ei) {
    const int np = h->fit_np, k = h->k;
    const size_t smem = sizeof(double) * ((size_t)PSTAGES * (PB_QT + PK_TILE) * PSLD + (size_t)PB_QT * k + 2 * k +
                                          PB_QT + PB_QT * 8);
    const int64_t ntiles = (m + PB_QT - 1) / PB_QT;
    const int grid = (int)((ntiles < 2 * (int64_t)h->sm_count) ? ntiles : 2 * (int64_t)h->sm_count);
    const size_t need = (size_t)grid * PB_QT * (np + 4);
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
    cudaError_t e = cudaFuncSetAttribute(predict_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int need_s = (sout != nullptr || ei != nullptr) ? 1 : 0;
    predict_big_kernel<<<grid, PRED_THREADS, smem, h->stream>>>(h->n, np, k, m, h->dX, Xq, hp, h->d_linv, h->d_w, mu,
                                                                sigma2, y_min, ei_weight, var_extra, need_s,
                                                                h->d_pscratch, yhat, sout, ei);
    h->launches++;
    return cudaGetLastError();
}


template <int NTW, int KK>
static cudaError_t launch_predict_reg_k(pk_handle* h, int64_t m, const double* Xq, const double* hp, double mu,
                                        double sigma2, double y_min, double ei_weight, double var_extra, double* yhat,
                                        double* sout, double* ei) {
    const int need_s = (sout != nullptr || ei != nullptr) ? 1 : 0;
    const int64_t ntiles = (m + PR_QT - 1) / PR_QT;
    const int64_t blocks = (ntiles < (int64_t)h->sm_count) ? ntiles : (int64_t)h->sm_count;  // persistent: one CTA per SM
    const int k = h->k;
    const size_t xs = ((size_t)h->n * k <= (size_t)PR_XS_MAX) ? (size_t)h->n * k : 0;  // staged sample coordinates
    const size_t smem = sizeof(double) * (64 * 16 + 3 * 64 * 16 + 2 * PR_QT * PR_ALD + 2 * PR_QT * ((k + 1) & ~1) +
                                          2 * PK_MAX_K + 2 * PR_WARPS * PR_QT + 2 * PR_WARPS * (PR_QT + 1) + 16 +
                                          2 * PR_WARPS * NTW * 128 + xs);
    static bool attr_done_dev[PK_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(predict_reg_kernel<NTW, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    predict_reg_kernel<NTW, KK><<<(unsigned)blocks, PR_THREADS, smem, h->stream>>>(
        h->n, h->fit_np, h->k, m, h->dX, Xq, hp, h->d_linv, h->d_w, mu, sigma2, y_min, ei_weight, var_extra, need_s, yhat,
        sout, ei);
    h->launches++;
    return cudaGetLastError();
}

template <int NTW>
static cudaError_t launch_predict_reg(pk_handle* h, int64_t m, const double* Xq, const double* hp, double mu,
                                      double sigma2, double y_min, double ei_weight, double var_extra, double* yhat,
                                      double* sout, double* ei) {
    switch (h->k) {
        case 1: return launch_predict_reg_k<NTW, 1>(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
        case 2: return launch_predict_reg_k<NTW, 2>(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
        case 3: return launch_predict_reg_k<NTW, 3>(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
        case 4: return launch_predict_reg_k<NTW, 4>(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
        default: return launch_predict_reg_k<NTW, 0>(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
    }
}

cudaError_t launch_predict(pk_handle* h, int64_t m, const double* Xq, const double* hp, double mu, double sigma2,
                           double y_min, double ei_weight, double var_extra, double* yhat, double* sout,
                           double* ei) {
    const int np = h->fit_np;
    // large query sets: psi rows + yhat in one kernel, the triangular DMMA contraction + s + EI in a second one
    const int64_t meff = (h->hint_m > m) ? h->hint_m : m;  // the query set this call is a slice of
    if (h->predict_variant == 3 || (h->predict_variant == 0 && meff >= 4096))
        return launch_predict_split(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
    if (np <= 512 && h->predict_variant != 1) {
        switch ((np / PK_TILE + 1) / 2) {  // 8-column blocks per warp: np/8 blocks over 16 warps
#define PK_PR(NN) case NN: return launch_predict_reg<NN>(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
            PK_PR(1) PK_PR(2) PK_PR(3) PK_PR(4)
#undef PK_PR
        }
    }
    if (np <= 512) return launch_predict_t<32>(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
    if (np <= 1280) return launch_predict_t<16>(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
    if (np <= 2944) return launch_predict_t<8>(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
    return launch_predict_big(h, m, Xq, hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, sout, ei);
}

}  // namespace pk
