// 64 x 64 tile factorisation and fragment helpers shared by the DMMA Cholesky kernels (pk_chol_cta.cu) and the
// INT8-tensor-core Cholesky (pk_chol_i8.cu).
#pragma once
#include "pk_common.cuh"

namespace pk {
namespace cta {

constexpr int TS = PK_TILE + 4;    // 68: stride of the diagonal tile while it is factorised in smem
constexpr int DLD = 12;            // row stride of the inverted 8x8 diagonal blocks
constexpr int D_DOUBLES = 8 * 8 * DLD;

__device__ __forceinline__ double negd(double x) {
    return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));  // sign flip on the ALU pipe
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

// ------------------------------------------------------------------------------------------------
// 64 x 64 Cholesky (+ 8 augmented rows) in shared memory (T, stride TS), 8-wide panels.  Only the 8x8
// diagonal block of a panel is factorised sequentially; everything below it is DMMA work:
//   (a) one warp: L_d = chol(T[kb,kb]) and D_kb = L_d^-1 together, by carrying the identity as eight
//       extra rows through the same eight column steps ([A; I] -> [L; L^-T]);
//   (b) all warps: T[bi, kb] <- T[bi, kb] D_kb^T for the row blocks below (one 8x8 block per warp);
//   (c) all warps: trailing update T[bi, bl] -= L[bi, kb] L[bl, kb]^T.
// The D_kb are exactly the inverted diagonal blocks the panel solve needs afterwards.
// ------------------------------------------------------------------------------------------------
// (a): lanes 0..7 own one row of the 8x8 block each (the other lanes mirror them).  Returns 0 or the
// 1-based tile-local column of the first pivot <= piv_tol * diag0 (warp uniform).  The failure test is
// recorded, not branched on: a non-positive pivot only turns the rest of the (discarded) tile into NaNs,
// and the compare stays off the dependent chain -- every FP64 instruction of this warp queues behind
// the other CTA's DMMAs, so the chain is kept to rsqrt, one multiply and one FMA per column.
__device__ __forceinline__ int diag_block_factor(double* T, double* Dsm, const double* diag0, int kb, double piv_tol,
                                                 int lane) {
    const int r = lane & 7, grp = lane & ~7, c0 = kb * 8;
    double v0[8], v1[8];
    {
        const double2* src = reinterpret_cast<const double2*>(T + (c0 + r) * TS + c0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double2 t = src[c];
            v0[2 * c] = t.x;
            v0[2 * c + 1] = t.y;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) v1[c] = (c == r) ? 1.0 : 0.0;
    }
    int bad = 0;
    double piv = __shfl_sync(0xffffffffu, v0[0], grp);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (!(piv > piv_tol * diag0[c0 + c]) && bad == 0) bad = c0 + c + 1;
        const double rinv = rsqrt(piv);  // l_cc = piv * rinv (<= 2 ulp)
        const double l0 = v0[c] * rinv, l1 = v1[c] * rinv;
        v0[c] = l0;
        v1[c] = l1;
        if (c < 7)  // next pivot straight from its owner lane (the same FMA the update below performs)
            piv = __shfl_sync(0xffffffffu, fma(-l0, l0, v0[c + 1]), grp | (c + 1));
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2) {
            const double m = __shfl_sync(0xffffffffu, l0, grp | c2);  // L_d[c2][c]
            v0[c2] = fma(-l0, m, v0[c2]);
            v1[c2] = fma(-l1, m, v1[c2]);
        }
    }
    if (lane < 8) {
        double2* dst = reinterpret_cast<double2*>(T + (c0 + r) * TS + c0);
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c] = make_double2(v0[2 * c], v0[2 * c + 1]);
        // v1 = row r of L_d^-T, i.e. column r of D = L_d^-1
#pragma unroll
        for (int c = 0; c < 8; ++c) Dsm[(c0 + c) * DLD + r] = v1[c];
    }
    return bad;
}

// (b): row block bi = kb + 1 + warp (block 8 = the augmented rows) times D_kb^T, in place.
// NBR = block rows of the tile: 9 with the augmented rows as block row 8 (DMMA kernels), 8 without (INT8 kernel)
template <int NBR = 9>
__device__ __forceinline__ void panel_scale_blocks(double* T, const double* Dsm, int kb, int warp, int lane) {
    const int bi = kb + 1 + warp;
    if (bi > NBR - 1) return;
    const int fr = lane >> 2, q = lane & 3;
    double* row = T + (bi * 8 + fr) * TS + kb * 8;
    const double a0 = row[q], a1 = row[4 + q];
    const double d0 = Dsm[(kb * 8 + fr) * DLD + q], d1 = Dsm[(kb * 8 + fr) * DLD + 4 + q];
    double x0 = 0.0, x1 = 0.0;
    dmma884(x0, x1, a0, d0);
    dmma884(x0, x1, a1, d1);
    *reinterpret_cast<double2*>(row + 2 * q) = make_double2(x0, x1);
}

// one 8x8 block of the trailing update after panel kb: T[bi, bl] -= L[bi, kb] L[bl, kb]^T (two DMMAs)
__device__ __forceinline__ void trailing_block(double* T, int kb, int bi, int bl, int lane) {
    const int fr = lane >> 2, q = lane & 3;
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

// trailing update after panel kb, WITHOUT the next diagonal block (kb+1, kb+1): that one belongs to warp 0, which
// updates and factorises it while warps 1..7 work through the rest (block row 8 = the augmented rows)
template <int NBR = 9>
__device__ __forceinline__ void trailing_update_rest(double* T, int kb, int warp, int lane) {
    int cnt = 0;
    for (int bi = kb + 1; bi < NBR; ++bi) {
        for (int bl = kb + 1; bl <= min(bi, 7); ++bl) {
            if (bi == kb + 1) continue;  // (kb+1, kb+1)
            if (cnt++ % 7 + 1 != warp) continue;
            trailing_block(T, kb, bi, bl, lane);
        }
    }
}

// The same for a CTA whose warps 0 and 4 share a scheduler (eight FP64 warps, one CTA per SM: the INT8 kernel): warp 4 stays
// out, so that the sequential pivot chain of warp 0 -- the critical path of the tile -- has its scheduler's FP64 pipe to
// itself instead of queueing behind warp 4's DMMAs (16 cycles each); the six other warps share the blocks.
template <int NBR = 8>
__device__ __forceinline__ void trailing_update_rest_quiet(double* T, int kb, int warp, int lane) {
    if (warp == 4) return;
    const int slot = (warp < 4) ? warp - 1 : warp - 2;  // warps 1 2 3 5 6 7 -> 0 .. 5
    int cnt = 0;
    for (int bi = kb + 1; bi < NBR; ++bi) {
        for (int bl = kb + 1; bl <= min(bi, 7); ++bl) {
            if (bi == kb + 1) continue;  // (kb+1, kb+1)
            if (cnt++ % 6 != slot) continue;
            trailing_block(T, kb, bi, bl, lane);
        }
    }
}

}  // namespace cta
}  // namespace pk
