// Batched blocked Cholesky, population variant: ONE persistent CTA factorises a whole candidate.
//
// Replaces np.linalg.cholesky (matrixops.py:31 / :41) plus the six np.linalg.solve calls of
// matrixops.neglikelihood / regneglikelihood (matrixops.py:45-66) for a GA/PSO population when there
// are at least as many candidates as SMs: every CTA pulls candidates from a device-side queue and
// runs the complete left-looking factorisation of its matrix without any grid-wide step -- block
// column by block column
//     diag  : T = Psi_jj - L[j,0:j] L[j,0:j]^T        (DMMA main loop, 64 rows)
//             L_jj = chol(T), D_b = inv(L_jj[b,b])   (8-wide panels: one warp factors a panel in
//                                                     registers with shuffles, all warps apply the
//                                                     trailing update with DMMA 8x8 blocks)
//     panel : T = M[r,j] - L[r,0:j] L[j,0:j]^T        (DMMA main loop, 128 rows per pass)
//             X L_jj^T = T solved IN REGISTERS as four more pipeline chunks: L_jj streams through
//             the B stages, each 8-column block is multiplied by the inverted 8x8 diagonal block
//             (2 DMMAs) and folded into the columns to its right (2 DMMAs per block)
// so the forward substitutions L^-1 1, L^-1 y (augmented rows) and, for pk_fit, L^-1 itself are just
// more panel rows.  A warp owns 16 complete rows of the 128 x 64 CTA tile (2 x 8 DMMA tiles), which
// is what keeps the triangular solve free of any cross-warp traffic.
//
// Two CTAs are resident per SM (256 threads, 128 registers, ~99 KB shared memory each): while one
// sits in the latency-bound 64x64 factorisation the other keeps the FP64 tensor pipe busy.
#include "pk_internal.h"

namespace pk {

namespace cta {

constexpr int THREADS = 256;
constexpr int KC = 16;             // K chunk per pipeline stage
constexpr int SLD = KC + 4;        // 20: conflict-free fragment loads (row stride == 4 mod 16 doubles)
constexpr int STAGES = 3;
constexpr int BM = 128;            // rows per panel pass (8 warps x 16 rows)
constexpr int A_STAGE = BM * SLD;
constexpr int B_STAGE = PK_TILE * SLD;
constexpr int PIPE_DOUBLES = STAGES * (A_STAGE + B_STAGE);
constexpr int TS = PK_TILE + 4;    // 68: stride of the diagonal tile while it is factorised in smem
constexpr int DLD = 12;            // row stride of the inverted 8x8 diagonal blocks
constexpr int D_DOUBLES = 8 * 8 * DLD;
constexpr int SMEM_DOUBLES = PIPE_DOUBLES + D_DOUBLES + PK_TILE + 2;  // + diag0[64] + control words
static_assert(PK_TILE * TS <= STAGES * A_STAGE, "the diagonal tile reuses the pipeline buffers");

__device__ __forceinline__ double negd(double x) {
    return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));  // sign flip on the ALU pipe
}

// One pipeline stage: B = L[j*64 .. +64, k0 .. k0+16) always, A = M[row0 .. row0+a_rows, k0 .. k0+16)
// only while the chunk lies left of block column j (the last four chunks carry L_jj itself for the
// triangular solve and need no A operand).
__device__ __forceinline__ void load_stage(double* As, double* Bs, const double* __restrict__ M, int ld, int row0,
                                           int a_rows, int brow0, int k0, bool with_a) {
    if (with_a) {
        for (int i = threadIdx.x; i < a_rows * 8; i += THREADS) {
            const int r = i >> 3, ch = i & 7;
            cp_async16(As + r * SLD + ch * 2, M + (size_t)(row0 + r) * ld + k0 + ch * 2);
        }
    }
    for (int i = threadIdx.x; i < PK_TILE * 8; i += THREADS) {
        const int r = i >> 3, ch = i & 7;
        cp_async16(Bs + r * SLD + ch * 2, M + (size_t)(brow0 + r) * ld + k0 + ch * 2);
    }
}

// acc(8*MI x 64 per warp) -= A(8*MI x 16) * B(64 x 16)^T for one stage.  MI (rows of 8 this warp owns in
// the pass) is a template parameter, selected by a warp-uniform branch: a DMMA that is merely
// predicated off still occupies its issue slots on the tensor pipe (measured: profiles/).
template <int MI>
__device__ __forceinline__ void gemm_consume(double (&acc)[2][8][2], const double* As, const double* Bs, int warp,
                                             int lane) {
    const int fr = lane >> 2, fc = lane & 3;
    const double* a_s = As + (warp * 16 + fr) * SLD + fc;
    const double* b_s = Bs + fr * SLD + fc;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
        double a[MI], b[8];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) a[mi] = negd(a_s[mi * 8 * SLD + kk]);
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) b[ni] = b_s[ni * 8 * SLD + kk];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
}

// A-operand fragments (two k-steps) of the 8-column block held in accumulator layout by (v0, v1):
// lane (fr, q) of a quad needs columns q and 4+q of its row; columns 2s, 2s+1 live in quad lane s.
__device__ __forceinline__ void acc_to_afrag(double v0, double v1, int lane, double& a0, double& a1) {
    const int q = lane & 3;
    const int src0 = (lane & ~3) | (q >> 1);
    const int src1 = src0 | 2;
    const double p00 = __shfl_sync(0xffffffffu, v0, src0), p01 = __shfl_sync(0xffffffffu, v1, src0);
    const double p10 = __shfl_sync(0xffffffffu, v0, src1), p11 = __shfl_sync(0xffffffffu, v1, src1);
    a0 = (q & 1) ? p01 : p00;
    a1 = (q & 1) ? p11 : p10;
}

// Triangular solve, pipeline chunk T (columns 16T .. 16T+15 of L_jj are in Bs): for the two 8-column
// blocks cb of the chunk  X_cb = T_cb * D_cb^T,  then  T_gb -= X_cb * L_jj[gb, cb]^T  for gb > cb.
template <int T, int MI>
__device__ __forceinline__ void trsm_consume(double (&acc)[2][8][2], const double* Bs, const double* Dsm, int lane) {
    const int fr = lane >> 2, q = lane & 3;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int cb = 2 * T + half;
        const double d0 = Dsm[(cb * 8 + fr) * DLD + q], d1 = Dsm[(cb * 8 + fr) * DLD + 4 + q];
        double ax[MI][2];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            double a0, a1;
            acc_to_afrag(acc[mi][2 * T + half][0], acc[mi][2 * T + half][1], lane, a0, a1);
            double x0 = 0.0, x1 = 0.0;
            dmma884(x0, x1, a0, d0);
            dmma884(x0, x1, a1, d1);
            acc[mi][2 * T + half][0] = x0;
            acc[mi][2 * T + half][1] = x1;
            acc_to_afrag(x0, x1, lane, a0, a1);
            ax[mi][0] = negd(a0);
            ax[mi][1] = negd(a1);
        }
#pragma unroll
        for (int gb = 2 * T + half + 1; gb < 8; ++gb) {
            const double b0 = Bs[(gb * 8 + fr) * SLD + half * 8 + q];
            const double b1 = Bs[(gb * 8 + fr) * SLD + half * 8 + 4 + q];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                dmma884(acc[mi][gb][0], acc[mi][gb][1], ax[mi][0], b0);
                dmma884(acc[mi][gb][0], acc[mi][gb][1], ax[mi][1], b1);
            }
        }
    }
}

// acc = M[row0 + 16 warp .., j*64 .. +64) -= L[rows, 0:j*64] L[j, 0:j*64]^T, then (with_trsm) the
// in-register solve against L_jj.  rows_valid (multiple of 8) rows of the pass exist.
template <bool WITH_TRSM>
__device__ __forceinline__ void block_rows(double (&acc)[2][8][2], double* pipe, const double* Dsm,
                                           const double* __restrict__ M, int ld, int j, int row0, int rows_valid) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, q = lane & 3;
    const int mi_count = max(0, min(2, (rows_valid - warp * 16) / 8));
    double* As = pipe;
    double* Bs = pipe + STAGES * A_STAGE;
    const int brow0 = j * PK_TILE;
    const int nk_gemm = j * (PK_TILE / KC);
    const int nkt = nk_gemm + (WITH_TRSM ? PK_TILE / KC : 0);
    // start the pipeline first, then fetch the accumulator tile while the copies are in flight
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkt) load_stage(As + s * A_STAGE, Bs + s * B_STAGE, M, ld, row0, rows_valid, brow0, s * KC, s < nk_gemm);
        cp_async_commit();
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            if (mi < mi_count) {
                const double2 v = *reinterpret_cast<const double2*>(
                    M + (size_t)(row0 + warp * 16 + mi * 8 + fr) * ld + brow0 + ni * 8 + 2 * q);
                acc[mi][ni][0] = v.x;
                acc[mi][ni][1] = v.y;
            } else {
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            }
        }
    }
    for (int it = 0; it < nk_gemm; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = it + STAGES - 1;
        if (nxt < nkt) {
            const int s = nxt % STAGES;
            load_stage(As + s * A_STAGE, Bs + s * B_STAGE, M, ld, row0, rows_valid, brow0, nxt * KC, nxt < nk_gemm);
        }
        cp_async_commit();
        const int s = it % STAGES;
        if (mi_count == 2) gemm_consume<2>(acc, As + s * A_STAGE, Bs + s * B_STAGE, warp, lane);
        else if (mi_count == 1) gemm_consume<1>(acc, As + s * A_STAGE, Bs + s * B_STAGE, warp, lane);
    }
    if (WITH_TRSM) {
#define PK_TRSM_STEP(TT)                                                                                          \
        {                                                                                                          \
            const int it = nk_gemm + TT;                                                                           \
            cp_async_wait<STAGES - 2>();                                                                           \
            __syncthreads();                                                                                       \
            const int nxt = it + STAGES - 1;                                                                       \
            if (nxt < nkt) {                                                                                       \
                const int s2 = nxt % STAGES;                                                                       \
                load_stage(As + s2 * A_STAGE, Bs + s2 * B_STAGE, M, ld, row0, rows_valid, brow0, nxt * KC, false); \
            }                                                                                                      \
            cp_async_commit();                                                                                     \
            if (mi_count == 2) trsm_consume<TT, 2>(acc, Bs + (it % STAGES) * B_STAGE, Dsm, lane);                  \
            else if (mi_count == 1) trsm_consume<TT, 1>(acc, Bs + (it % STAGES) * B_STAGE, Dsm, lane);             \
        }
        PK_TRSM_STEP(0)
        PK_TRSM_STEP(1)
        PK_TRSM_STEP(2)
        PK_TRSM_STEP(3)
#undef PK_TRSM_STEP
    }
    cp_async_wait<0>();
    __syncthreads();  // every warp is done with the stages (the caller reuses them)
}

// ------------------------------------------------------------------------------------------------
// 64 x 64 Cholesky in shared memory (T, stride TS), 8-wide panels.
// ------------------------------------------------------------------------------------------------
// Panel kb, executed by ONE warp: lane l owns rows l and l+32 of the panel's 8 columns in registers;
// the eight column steps exchange the pivot and the pivot-block column entries with shuffles.
// Returns 0 or the 1-based tile-local column of the first pivot <= piv_tol * diag0 (warp uniform).
template <int SLOT>
__device__ __forceinline__ int panel_factor(double* T, const double* diag0, int kb, double piv_tol, int lane) {
    double v[2][8];
    const int c0 = kb * 8;
#pragma unroll
    for (int s = SLOT; s < 2; ++s) {
        const double2* src = reinterpret_cast<const double2*>(T + (lane + 32 * s) * TS + c0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double2 t = src[c];
            v[s][2 * c] = t.x;
            v[s][2 * c + 1] = t.y;
        }
    }
    // The failure test is recorded, not branched on: a non-positive pivot only turns the rest of the
    // (discarded) panel into NaNs, and keeping the compare off the dependent chain matters because every
    // FP64 instruction of this warp queues behind the other CTA's DMMAs.
    int bad = 0;
    double piv = __shfl_sync(0xffffffffu, v[SLOT][0], c0 & 31);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int gc = c0 + c;
        if (!(piv > piv_tol * diag0[gc]) && bad == 0) bad = gc + 1;
        const double rinv = rsqrt(piv);  // l_cc = piv * rinv (<= 2 ulp), rows below scale by rinv
        double l[2];
#pragma unroll
        for (int s = SLOT; s < 2; ++s) {
            l[s] = v[s][c] * rinv;
            v[s][c] = l[s];
        }
        if (c < 7)  // next pivot straight from its owner lane: a_cc - l^2 is the same FMA the update below performs
            piv = __shfl_sync(0xffffffffu, fma(-l[SLOT], l[SLOT], v[SLOT][c + 1]), (gc + 1) & 31);
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2) {
            const double m = __shfl_sync(0xffffffffu, l[SLOT], (c0 + c2) & 31);  // L[c0 + c2][gc]
#pragma unroll
            for (int s = SLOT; s < 2; ++s) v[s][c2] = fma(-l[s], m, v[s][c2]);
        }
    }
#pragma unroll
    for (int s = SLOT; s < 2; ++s) {
        const int r = lane + 32 * s;
        if (r >= c0) {
            double2* dst = reinterpret_cast<double2*>(T + r * TS + c0);
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[c] = make_double2(v[s][2 * c], v[s][2 * c + 1]);
        }
    }
    return bad;
}

// trailing update after panel kb: T[bi, bl] -= L[bi, kb] L[bl, kb]^T for kb < bl <= bi, 8x8 blocks
// spread over the warps, two DMMAs each.
__device__ __forceinline__ void trailing_update(double* T, int kb, int warp, int lane) {
    const int fr = lane >> 2, q = lane & 3;
    int cnt = 0;
    for (int bi = kb + 1; bi < 8; ++bi) {
        for (int bl = kb + 1; bl <= bi; ++bl, ++cnt) {
            if ((cnt & 7) != warp) continue;
            double2* cp = reinterpret_cast<double2*>(T + (bi * 8 + fr) * TS + bl * 8 + 2 * q);
            double2 c = *cp;
            const double a0 = negd(T[(bi * 8 + fr) * TS + kb * 8 + q]);
            const double a1 = negd(T[(bi * 8 + fr) * TS + kb * 8 + 4 + q]);
            const double b0 = T[(bl * 8 + fr) * TS + kb * 8 + q];
            const double b1 = T[(bl * 8 + fr) * TS + kb * 8 + 4 + q];
            dmma884(c.x, c.y, a0, b0);
            dmma884(c.x, c.y, a1, b1);
            *cp = c;
        }
    }
}

// D_w = inv(L_jj[w, w]) (8 x 8 lower triangular), one warp per block; lanes 0..7 own a row each
// (the other lanes mirror them), rows finalised top-down and broadcast with shuffles.
__device__ __forceinline__ void invert_diag_block(const double* T, double* Dsm, int w, int lane) {
    const int r = lane & 7;
    double Lr[8], s[8], Dr[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        Lr[m] = T[(w * 8 + r) * TS + w * 8 + m];
        s[m] = 0.0;
        Dr[m] = 0.0;
    }
    double rinv = 1.0;
#pragma unroll
    for (int m = 0; m < 8; ++m)
        if (r == m) rinv = 1.0 / Lr[m];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        if (r == m) {
#pragma unroll
            for (int c = 0; c < m; ++c) Dr[c] = -rinv * s[c];
            Dr[m] = rinv;
        }
#pragma unroll
        for (int c = 0; c <= m; ++c) {
            const double bm = __shfl_sync(0xffffffffu, Dr[c], (lane & ~7) | m);
            if (r > m) s[c] = fma(Lr[m], bm, s[c]);
        }
    }
    if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) Dsm[(w * 8 + r) * DLD + c] = Dr[c];
    }
}

// ------------------------------------------------------------------------------------------------
// Diagonal tile update T = Psi_jj - L[j,0:j] L[j,0:j]^T, LOWER 8x8 blocks only (36 of 64).  Warp w owns
// block row br(w) = w for w < 4 and 11 - w otherwise, so the two warps that share a scheduler carry
// 9 blocks together and the four tensor pipes stay balanced.  A and B are the same 64 rows: only the
// B stages are loaded.  NI = br + 1 is dispatched by a warp-uniform switch (no predicated-off DMMAs).
// ------------------------------------------------------------------------------------------------
template <int NI>
__device__ __forceinline__ void diag_consume(double (&acc)[8][2], const double* Bs, int br, int lane) {
    const int fr = lane >> 2, fc = lane & 3;
    const double* a_s = Bs + (br * 8 + fr) * SLD + fc;
    const double* b_s = Bs + fr * SLD + fc;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
        const double a = negd(a_s[kk]);
        double b[NI];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) b[ni] = b_s[ni * 8 * SLD + kk];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) dmma884(acc[ni][0], acc[ni][1], a, b[ni]);
    }
}

__device__ __forceinline__ void diag_gemm(double* T, double* pipe, const double* __restrict__ M, int ld, int j) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, q = lane & 3;
    const int br = (warp < 4) ? warp : 11 - warp;
    double* Bs = pipe + STAGES * A_STAGE;
    const int row0 = j * PK_TILE;
    const int nk = j * (PK_TILE / KC);
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(nullptr, Bs + s * B_STAGE, M, ld, 0, 0, row0, s * KC, false);
        cp_async_commit();
    }
    double acc[8][2];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
        if (ni <= br) {
            const double2 v =
                *reinterpret_cast<const double2*>(M + (size_t)(row0 + br * 8 + fr) * ld + row0 + ni * 8 + 2 * q);
            acc[ni][0] = v.x;
            acc[ni][1] = v.y;
        } else {
            acc[ni][0] = acc[ni][1] = 0.0;
        }
    }
    for (int it = 0; it < nk; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = it + STAGES - 1;
        if (nxt < nk) load_stage(nullptr, Bs + (nxt % STAGES) * B_STAGE, M, ld, 0, 0, row0, nxt * KC, false);
        cp_async_commit();
        const double* bs = Bs + (it % STAGES) * B_STAGE;
        switch (br) {
            case 0: diag_consume<1>(acc, bs, br, lane); break;
            case 1: diag_consume<2>(acc, bs, br, lane); break;
            case 2: diag_consume<3>(acc, bs, br, lane); break;
            case 3: diag_consume<4>(acc, bs, br, lane); break;
            case 4: diag_consume<5>(acc, bs, br, lane); break;
            case 5: diag_consume<6>(acc, bs, br, lane); break;
            case 6: diag_consume<7>(acc, bs, br, lane); break;
            default: diag_consume<8>(acc, bs, br, lane); break;
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // the stages are drained: T (aliasing the A stages) may be written
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
        if (ni <= br)
            *reinterpret_cast<double2*>(T + (br * 8 + fr) * TS + ni * 8 + 2 * q) = make_double2(acc[ni][0], acc[ni][1]);
    }
}

// Diagonal tile of block column j.  Returns 0 or the tile-local failing column (CTA uniform).
__device__ __forceinline__ int diag_tile(double* pipe, double* Dsm, double* diag0, int* ctl, double* __restrict__ M,
                                         int ld, int j, double piv_tol) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = j * PK_TILE;
    if (tid < PK_TILE) diag0[tid] = M[(size_t)(row0 + tid) * ld + row0 + tid];
    double* T = pipe;
    diag_gemm(T, pipe, M, ld, j);
    if (tid == 0) ctl[1] = 0;
    __syncthreads();
    for (int kb = 0; kb < 8; ++kb) {
        if (warp == 0) {
            const int bad = (kb < 4) ? panel_factor<0>(T, diag0, kb, piv_tol, lane)
                                     : panel_factor<1>(T, diag0, kb, piv_tol, lane);
            if (bad && lane == 0) ctl[1] = bad;
        }
        __syncthreads();
        if (ctl[1]) return ctl[1];
        if (kb < 7) {
            trailing_update(T, kb, warp, lane);
            __syncthreads();
        }
    }
    invert_diag_block(T, Dsm, warp, lane);
    for (int i = tid; i < PK_TILE * (PK_TILE / 2); i += THREADS) {
        const int rr = i >> 5, c2 = (i & 31) * 2;
        const double2 t = *reinterpret_cast<const double2*>(T + rr * TS + c2);
        *reinterpret_cast<double2*>(M + (size_t)(row0 + rr) * ld + row0 + c2) =
            make_double2((c2 <= rr) ? t.x : 0.0, (c2 + 1 <= rr) ? t.y : 0.0);
    }
    __syncthreads();  // L_jj is in global memory and D in shared memory for every thread of the CTA
    return 0;
}

__global__ void __launch_bounds__(THREADS, 2)
chol_cta_kernel(int C, int nt, int rows, int ld, size_t elems, double piv_tol, double* __restrict__ ws,
                int32_t* __restrict__ info, unsigned int* __restrict__ queue) {
    extern __shared__ __align__(16) double smem[];
    double* pipe = smem;
    double* Dsm = pipe + PIPE_DOUBLES;
    double* diag0 = Dsm + D_DOUBLES;
    int* ctl = reinterpret_cast<int*>(diag0 + PK_TILE);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, q = lane & 3;
    for (;;) {
        __syncthreads();
        if (tid == 0) ctl[0] = (int)atomicAdd(queue, 1u);
        __syncthreads();
        const int cand = ctl[0];
        if (cand >= C) break;
        if (info[cand] != 0) continue;
        double* M = ws + (size_t)cand * elems;
        for (int j = 0; j < nt; ++j) {
            const int bad = diag_tile(pipe, Dsm, diag0, ctl, M, ld, j, piv_tol);
            if (bad) {
                if (tid == 0) info[cand] = j * PK_TILE + bad;  // dpotrf convention: order of the failing minor
                break;
            }
            for (int row0 = (j + 1) * PK_TILE; row0 < rows; row0 += BM) {
                const int rows_valid = min(BM, rows - row0);
                double acc[2][8][2];
                block_rows<true>(acc, pipe, Dsm, M, ld, j, row0, rows_valid);
                const int mi_count = max(0, min(2, (rows_valid - warp * 16) / 8));
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    if (mi < mi_count) {
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni)
                            *reinterpret_cast<double2*>(M + (size_t)(row0 + warp * 16 + mi * 8 + fr) * ld +
                                                        j * PK_TILE + ni * 8 + 2 * q) =
                                make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                    }
                }
            }
            __syncthreads();  // the panel of column j is visible to the whole CTA before column j+1 reads it
        }
    }
}

}  // namespace cta

cudaError_t launch_cholesky_cta(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info) {
    const int nt = s.np / PK_TILE;
    const size_t smem = sizeof(double) * (size_t)cta::SMEM_DOUBLES;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e =
            cudaFuncSetAttribute(cta::chol_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    cudaError_t e = cudaMemsetAsync(h->d_queue, 0, sizeof(unsigned int), h->stream);
    if (e != cudaSuccess) return e;
    const int grid = (C < 2 * h->sm_count) ? C : 2 * h->sm_count;
    cta::chol_cta_kernel<<<grid, cta::THREADS, smem, h->stream>>>(C, nt, s.rows, s.ld, s.elems(), h->pivot_tol, ws,
                                                                  info, h->d_queue);
    h->launches++;
    return cudaGetLastError();
}

}  // namespace pk
