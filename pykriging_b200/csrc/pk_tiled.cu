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
// GEMM core: acc(BM x 64) = A(BM x K) * B(64 x K)^T, A and B row-major with leading dimension ld
// (K contiguous), staged through shared memory with a 3-stage cp.async pipeline and consumed by
// DMMA 8x8x4 fragments.  Warp tile 32x32 (4x4 DMMA tiles, 32 FP64 accumulators per thread).
// ------------------------------------------------------------------------------------------------
constexpr int KC = 16;            // K chunk per pipeline stage
constexpr int SLD = KC + 4;       // smem row stride (doubles): 20 -> conflict-free fragment loads
constexpr int STAGES = 3;

template <int BM, int THREADS>
struct GemmSmem {
    static constexpr int A_STAGE = BM * SLD;
    static constexpr int B_STAGE = PK_TILE * SLD;
    static constexpr int PIPE_DOUBLES = STAGES * (A_STAGE + B_STAGE);
};

template <int BM, int THREADS>
__device__ __forceinline__ void gemm_load_stage(double* As, double* Bs, const double* __restrict__ Ag,
                                                const double* __restrict__ Bg, int ld, int k0, int a_rows_valid) {
    // A: BM rows x 8 chunks of 16 B ; B: 64 rows x 8 chunks
    for (int i = threadIdx.x; i < BM * 8; i += THREADS) {
        const int r = i >> 3, ch = i & 7;
        const int rs = (r < a_rows_valid) ? r : (a_rows_valid - 1);  // clamp: never read past the matrix
        cp_async16(As + r * SLD + ch * 2, Ag + (size_t)rs * ld + k0 + ch * 2);
    }
    for (int i = threadIdx.x; i < PK_TILE * 8; i += THREADS) {
        const int r = i >> 3, ch = i & 7;
        cp_async16(Bs + r * SLD + ch * 2, Bg + (size_t)r * ld + k0 + ch * 2);
    }
}

// acc[mi][ni][2]; warp (wm, wn) covers rows wm*32.., cols wn*32..
template <int BM, int THREADS>
__device__ __forceinline__ void gemm_mainloop(double (&acc)[4][4][2], double* pipe, const double* __restrict__ Ag,
                                              const double* __restrict__ Bg, int ld, int K, int a_rows_valid,
                                              int wm, int wn, int mi_count) {
    using S = GemmSmem<BM, THREADS>;
    double* As = pipe;
    double* Bs = pipe + STAGES * S::A_STAGE;
    const int lane = threadIdx.x & 31;
    const int nk = K / KC;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) gemm_load_stage<BM, THREADS>(As + s * S::A_STAGE, Bs + s * S::B_STAGE, Ag, Bg, ld, s * KC, a_rows_valid);
        cp_async_commit();
    }
    const int fr = lane >> 2, fc = lane & 3;
    for (int it = 0; it < nk; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = it + STAGES - 1;
        if (nxt < nk) {
            const int s = nxt % STAGES;
            gemm_load_stage<BM, THREADS>(As + s * S::A_STAGE, Bs + s * S::B_STAGE, Ag, Bg, ld, nxt * KC, a_rows_valid);
        }
        cp_async_commit();
        const double* a_s = As + (it % STAGES) * S::A_STAGE + (wm * 32 + fr) * SLD + fc;
        const double* b_s = Bs + (it % STAGES) * S::B_STAGE + (wn * 32 + fr) * SLD + fc;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = (mi < mi_count) ? a_s[mi * 8 * SLD + kk] : 0.0;
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
    cp_async_wait<0>();
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Diagonal tile: T = Psi_jj - L[j,0:j] L[j,0:j]^T, then unblocked Cholesky in shared memory.
// grid = candidates, 128 threads (2x2 warps of 32x32).
// ------------------------------------------------------------------------------------------------
constexpr int DIAG_THREADS = 128;
constexpr int TLD = PK_TILE + 1;  // 65: odd stride -> conflict-free row-per-thread walks

__global__ void __launch_bounds__(DIAG_THREADS)
chol_diag_kernel(int j, int n, int ld, size_t elems, double piv_tol, double* __restrict__ ws,
                 int32_t* __restrict__ info) {
    extern __shared__ __align__(16) double smem[];
    const int cand = blockIdx.x;
    if (info[cand] != 0) return;
    double* M = ws + (size_t)cand * elems;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const double* Ag = M + (size_t)(j * PK_TILE) * ld;
    gemm_mainloop<PK_TILE, DIAG_THREADS>(acc, smem, Ag, Ag, ld, j * PK_TILE, PK_TILE, wm, wn, 4);

    double* T = smem;  // aliases the (drained) pipeline buffers
    {
        const int fr = lane >> 2, fc2 = (lane & 3) * 2;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int r = wm * 32 + mi * 8 + fr, c = wn * 32 + ni * 8 + fc2;
                const double2 v = *reinterpret_cast<const double2*>(M + (size_t)(j * PK_TILE + r) * ld + j * PK_TILE + c);
                T[r * TLD + c] = v.x - acc[mi][ni][0];
                T[r * TLD + c + 1] = v.y - acc[mi][ni][1];
            }
    }
    // In-tile unblocked Cholesky.  Off-diagonal entries are updated by sequential FMAs; every pivot is
    // formed as t_cc - (sum_k l_ck^2) with the sum kept in a register of the thread that owns row c
    // (dot-product form of dpotf2, see pk_small.cu).
    const int r = tid & 63, h = tid >> 6;
    double dsum = 0.0;                   // meaningful in threads with h == 0
    // original diagonal entry Psi_rr of this thread's row (scale of the pivot threshold)
    const double diag0 = M[(size_t)(j * PK_TILE + r) * ld + j * PK_TILE + r];
    double* piv_s = smem + PK_TILE * TLD;  // 1 double: pivot broadcast
    for (int c = 0; c < PK_TILE; ++c) {
        __syncthreads();
        if (tid == c) {
            piv_s[0] = T[c * TLD + c] - dsum;
            piv_s[1] = piv_tol * diag0;
        }
        __syncthreads();
        const double piv = piv_s[0];
        if (!(piv > piv_s[1])) {  // dpotrf convention: index of the first non-positive (or NaN) pivot
            if (tid == 0) info[cand] = j * PK_TILE + c + 1;
            return;
        }
        const double l = sqrt(piv);
        const double rinv = 1.0 / l;
        if (h == 0 && r > c) T[r * TLD + c] *= rinv;
        __syncthreads();
        if (tid == 0) T[c * TLD + c] = l;
        if (r > c) {
            const double lrc = T[r * TLD + c];
            for (int cc = c + 1 + h; cc < r; cc += 2) T[r * TLD + cc] = fma(-lrc, T[cc * TLD + c], T[r * TLD + cc]);
            if (h == 0) dsum = fma(lrc, lrc, dsum);
        }
    }
    __syncthreads();
    // write L_jj (lower) and zeros above the diagonal
    for (int i = tid; i < PK_TILE * PK_TILE; i += DIAG_THREADS) {
        const int rr = i >> 6, cc = i & 63;
        M[(size_t)(j * PK_TILE + rr) * ld + j * PK_TILE + cc] = (cc <= rr) ? T[rr * TLD + cc] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Rows below the diagonal tile of block column j: T = M[i,j] - L[i,0:j] L[j,0:j]^T, L[i,j] = T L_jj^-T.
// grid = (row blocks of 128, candidates), 256 threads (4x2 warps of 32x32).
// ------------------------------------------------------------------------------------------------
constexpr int OFF_BM = 128;
constexpr int OFF_THREADS = 256;
constexpr int LLD = PK_TILE + 2;  // 66: L_jj rows 16-byte aligned for broadcast LDS.128

__global__ void __launch_bounds__(OFF_THREADS, 2)
chol_off_kernel(int j, int rows, int ld, size_t elems, double* __restrict__ ws, const int32_t* __restrict__ info) {
    extern __shared__ __align__(16) double smem[];
    const int cand = blockIdx.y;
    if (info[cand] != 0) return;
    double* M = ws + (size_t)cand * elems;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int row0 = (j + 1) * PK_TILE + blockIdx.x * OFF_BM;
    const int rows_valid = min(OFF_BM, rows - row0);            // multiple of 8, >= 8
    const int mi_count = max(0, min(4, (rows_valid - wm * 32) / 8));

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const double* Ag = M + (size_t)row0 * ld;
    const double* Bg = M + (size_t)(j * PK_TILE) * ld;
    gemm_mainloop<OFF_BM, OFF_THREADS>(acc, smem, Ag, Bg, ld, j * PK_TILE, rows_valid, wm, wn, mi_count);

    double* T = smem;                      // OFF_BM x TLD
    double* Ljj = smem + OFF_BM * TLD;     // 64 x LLD
    double* rdiag = Ljj + PK_TILE * LLD;   // 64
    {
        const int fr = lane >> 2, fc2 = (lane & 3) * 2;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            if (mi < mi_count) {
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int r = wm * 32 + mi * 8 + fr, c = wn * 32 + ni * 8 + fc2;
                    const double2 v = *reinterpret_cast<const double2*>(M + (size_t)(row0 + r) * ld + j * PK_TILE + c);
                    T[r * TLD + c] = v.x - acc[mi][ni][0];
                    T[r * TLD + c + 1] = v.y - acc[mi][ni][1];
                }
            }
        }
    }
    for (int i = tid; i < PK_TILE * (PK_TILE / 2); i += OFF_THREADS) {
        const int rr = i >> 5, c2 = (i & 31) * 2;
        const double2 v = *reinterpret_cast<const double2*>(M + (size_t)(j * PK_TILE + rr) * ld + j * PK_TILE + c2);
        *reinterpret_cast<double2*>(Ljj + rr * LLD + c2) = v;
        if (c2 == rr) rdiag[rr] = 1.0 / v.x;
        else if (c2 + 1 == rr) rdiag[rr] = 1.0 / v.y;
    }
    __syncthreads();

    // forward substitution X L_jj^T = T, one thread per row, 8-column register blocks
    if (tid < rows_valid) {
        double* trow = T + tid * TLD;
#pragma unroll 1
        for (int cb = 0; cb < 8; ++cb) {
            double x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = trow[cb * 8 + i];
#pragma unroll 1
            for (int pb = 0; pb < cb; ++pb) {
                double xp[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) xp[q] = trow[pb * 8 + q];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double2* lrow = reinterpret_cast<const double2*>(Ljj + (cb * 8 + i) * LLD + pb * 8);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 l2 = lrow[q];
                        x[i] = fma(-xp[2 * q], l2.x, x[i]);
                        x[i] = fma(-xp[2 * q + 1], l2.y, x[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double* lrow = Ljj + (cb * 8 + i) * LLD + cb * 8;
#pragma unroll
                for (int q = 0; q < i; ++q) x[i] = fma(-x[q], lrow[q], x[i]);
                x[i] *= rdiag[cb * 8 + i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) trow[cb * 8 + i] = x[i];
        }
    }
    __syncthreads();
    for (int i = tid; i < rows_valid * PK_TILE; i += OFF_THREADS) {
        const int rr = i >> 6, cc = i & 63;
        M[(size_t)(row0 + rr) * ld + j * PK_TILE + cc] = T[rr * TLD + cc];
    }
}

// ------------------------------------------------------------------------------------------------
// mu, sigma^2, ln|Psi|, NegLnLike from the factorised augmented matrix (matrixops.py:46-56).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nll_epilogue_kernel(int n, int np, int ld, size_t elems, const double* __restrict__ ws, double* __restrict__ out4,
                    int out_stride, const int32_t* __restrict__ info) {
    __shared__ double scratch[32];
    const int cand = blockIdx.x, tid = threadIdx.x;
    if (info[cand] != 0) {
        if (tid < 4) out4[tid * out_stride + cand] = CUDART_NAN;
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

cudaError_t launch_cholesky(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info) {
    const int nt = s.np / PK_TILE;
    const size_t diag_smem = sizeof(double) * (size_t)max(GemmSmem<PK_TILE, DIAG_THREADS>::PIPE_DOUBLES, PK_TILE * TLD + 8);
    const size_t off_smem = sizeof(double) * (size_t)max(GemmSmem<OFF_BM, OFF_THREADS>::PIPE_DOUBLES,
                                                          OFF_BM * TLD + PK_TILE * LLD + PK_TILE);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(chol_off_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off_smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    for (int j = 0; j < nt; ++j) {
        chol_diag_kernel<<<C, DIAG_THREADS, diag_smem, h->stream>>>(j, s.n, s.ld, s.elems(), h->pivot_tol, ws, info);
        h->launches++;
        const int rows_below = s.rows - (j + 1) * PK_TILE;
        if (rows_below > 0) {
            dim3 grid((rows_below + OFF_BM - 1) / OFF_BM, C);
            chol_off_kernel<<<grid, OFF_THREADS, off_smem, h->stream>>>(j, s.rows, s.ld, s.elems(), ws, info);
            h->launches++;
        }
    }
    return cudaGetLastError();
}

cudaError_t launch_nll_epilogue(pk_handle* h, const TiledShape& s, int C, const double* ws, double* out4,
                                int out_stride, int32_t* info) {
    nll_epilogue_kernel<<<C, 256, 0, h->stream>>>(s.n, s.np, s.ld, s.elems(), ws, out4, out_stride, info);
    h->launches++;
    return cudaGetLastError();
}

}  // namespace pk
