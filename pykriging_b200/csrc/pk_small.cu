// Fused small-n likelihood kernel: ONE CTA per (theta, p[, lambda]) candidate.
//
// Replaces, per candidate, matrixops.updatePsi / regupdatePsi (matrixops.py:24-42) followed by
// matrixops.neglikelihood / regneglikelihood (matrixops.py:45-66) -- the body of the candidate loop in
// kriging.fittingObjective (krige.py:436-451).  For n <= PK_SMALL_N_MAX the correlation matrix is built
// straight into shared memory (it never touches HBM), factorised in place, and the two forward
// substitutions L^-1 1 and L^-1 y ride along as two augmented rows under Psi.  HBM traffic per candidate:
// 2k doubles in, 4 doubles + 1 int out (the n x k sample matrix is re-read from L2 by every CTA).
#include "pk_internal.h"

namespace pk {

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
small_nll_kernel(int n, int k, const double* __restrict__ X, const double* __restrict__ y,
                 const double* __restrict__ theta, const double* __restrict__ p,
                 const double* __restrict__ lambda, int mode, double piv_tol, double* __restrict__ out4,
                 int out_stride, int32_t* __restrict__ info) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int cand = blockIdx.x;
    const int ld = n | 1;          // odd leading dimension -> conflict-free column walks
    const int rows = n + 2;        // + augmented rows 1^T, y^T
    double* A = sm;                // rows x ld (lower triangle of Psi + 2 full augmented rows)
    double* Xs = A + rows * ld;    // n x k
    double* th = Xs + n * k;       // k
    double* pp = th + k;           // k
    double* scratch = pp + k;      // 32
    double* dsum = scratch + 32;   // n: running sum_k L[r][k]^2 of every diagonal entry (see below)

    for (int i = tid; i < n * k; i += THREADS) Xs[i] = X[i];
    if (tid < k) {
        th[tid] = theta[(size_t)cand * k + tid];
        pp[tid] = p[(size_t)cand * k + tid];
    }
    __syncthreads();

    const double diag = 1.0 + (mode == PK_MODE_REGRESSION ? lambda[cand] : PK_EPS_DIAG);
    const int tx = tid & 15, ty = tid >> 4;
    constexpr int TY = THREADS / 16;

    // ---- Psi (lower triangle) + augmented rows -------------------------------------------------
    for (int r = ty; r < n; r += TY) {
        for (int c = tx; c <= r; c += 16) {
            double v;
            if (c == r) {
                v = diag;
            } else {
                v = exp(-corr_exponent(Xs + r * k, Xs + c * k, th, pp, k));
            }
            A[r * ld + c] = v;
        }
    }
    for (int c = tid; c < n; c += THREADS) {
        A[n * ld + c] = 1.0;
        A[(n + 1) * ld + c] = y[c];
        dsum[c] = 0.0;
    }

    // ---- right-looking Cholesky on [Psi; 1^T; y^T] ----------------------------------------------
    // Off-diagonal entries accumulate a_rc - sum_k l_rk l_ck by sequential FMAs in k order; the pivot
    // is formed as a_cc - (sum_k l_ck^2) with the sum accumulated separately and subtracted once.
    // That is the operation order of the unblocked LAPACK kernel behind np.linalg.cholesky
    // (dot-product form of dpotf2), which matters only for *which* numerically singular matrices
    // trip the "pivot <= 0" test -- the failure rule itself is LAPACK's.
    for (int c = 0; c < n; ++c) {
        __syncthreads();
        const double piv = A[c * ld + c] - dsum[c];
        if (!(piv > piv_tol * A[c * ld + c])) {  // dpotrf: "ajj <= 0 or NaN" -> info = c+1 (np.linalg.LinAlgError in the reference)
            if (tid == 0) {
                info[cand] = c + 1;
                const double qnan = CUDART_NAN;
                out4[0 * out_stride + cand] = qnan;
                out4[1 * out_stride + cand] = qnan;
                out4[2 * out_stride + cand] = qnan;
                out4[3 * out_stride + cand] = qnan;
            }
            return;
        }
        const double l = sqrt(piv);
        const double rinv = 1.0 / l;
        for (int r = c + 1 + tid; r < rows; r += THREADS) A[r * ld + c] *= rinv;
        __syncthreads();
        if (tid == 0) A[c * ld + c] = l;
        for (int r = c + 1 + ty; r < rows; r += TY) {
            const double lrc = A[r * ld + c];
            const int cmax = (r < n) ? (r - 1) : (n - 1);
            for (int cc = c + 1 + tx; cc <= cmax; cc += 16) A[r * ld + cc] = fma(-lrc, A[cc * ld + c], A[r * ld + cc]);
            if (tx == 0 && r < n) dsum[r] = fma(lrc, lrc, dsum[r]);
        }
    }
    __syncthreads();

    // ---- mu, sigma^2, ln|Psi|, NegLnLike (matrixops.py:46-56 in the a = L^-1 1, d = L^-1 y form) ------
    const double* a = A + n * ld;
    const double* d = A + (n + 1) * ld;
    double saa = 0.0, sad = 0.0, slog = 0.0;
    for (int c = tid; c < n; c += THREADS) {
        saa += a[c] * a[c];
        sad += a[c] * d[c];
        slog += log(A[c * ld + c]);
    }
    saa = block_sum(saa, scratch);
    sad = block_sum(sad, scratch);
    slog = block_sum(slog, scratch);
    const double mu = sad / saa;
    double srr = 0.0;
    for (int c = tid; c < n; c += THREADS) {
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
        info[cand] = 0;
    }
}

cudaError_t launch_small_nll(pk_handle* h, int C, const double* theta, const double* p, const double* lambda,
                             int mode, double* out4, int32_t* info) {
    const int n = h->n, k = h->k;
    const int ld = n | 1;
    const size_t smem = sizeof(double) * ((size_t)(n + 2) * ld + (size_t)n * k + 2 * k + 32 + n);
    cudaError_t e;
    if (n <= 64) {
        auto kern = small_nll_kernel<128>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<C, 128, smem, h->stream>>>(n, k, h->dX, h->dy, theta, p, lambda, mode, h->pivot_tol, out4, h->cap_C, info);
    } else {
        auto kern = small_nll_kernel<256>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<C, 256, smem, h->stream>>>(n, k, h->dX, h->dy, theta, p, lambda, mode, h->pivot_tol, out4, h->cap_C, info);
    }
    h->launches++;
    return cudaGetLastError();
}

}  // namespace pk
