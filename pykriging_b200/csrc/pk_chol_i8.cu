// Batched Cholesky on the INT8 tensor cores: the product sums of the left-looking factorisation
//     T[r, j] = Psi[r, j] - L[r, 0:j] L[j, 0:j]^T
// (np.linalg.cholesky of matrixops.py:31 / :41 and the forward substitutions behind the six np.linalg.solve calls of
// matrixops.py:45-66, for a whole population) run as tcgen05.mma.kind::i8 with accumulators in TMEM.
//
// The 5th-generation tensor cores have no FP64 path (DMMA runs on the legacy pipe, 37 TFLOP/s), but their INT8 path adds
// exact INT32 sums at petaop rates.  Ozaki splitting: every finished entry of L is stored as S = 7 balanced base-256
// digits of a fixed-point number,
//     L_ik = 2^e (d_1 2^-8 + ... + d_7 2^-56),   -128 <= d_t <= 127,   2^e > 2 max|L|  (|L_ik| <= sqrt(Psi_ii)),
// in 7 byte planes per candidate; the digit-pair products with t + u <= 8 (28 of 49) are INT8 MMAs whose INT32 sums are
// exact and order independent, pairs of equal weight share one TMEM accumulator (7 levels x 64 columns), and the levels
// are recombined in integer arithmetic before ONE rounding to FP64: a dot product with no rounding inside the sum and
// operands good to 2^-56 of the scale -- the accuracy class of the DMMA path (which rounds every partial sum).
//
// One persistent CTA per SM factorises a whole candidate (candidates from a device-side queue), block column j by block
// column, in passes of 128 rows that start AT the diagonal tile (pass 0 = diagonal tile + the 64 rows below it):
//   warp 8    TMA producer: per 64-byte K chunk the 7 planes of the pass's 128 rows and of row block j (two 4-D tile
//             copies, 84 KB per stage, 2 stages); chunk j-1 -- written during block column j-1 -- waits for that column
//   warp 9    MMA issuer: per chunk and A plane t ONE MMA against the B planes 0 .. 6-t, which lie back to back in shared
//             memory and whose levels t .. 6 lie back to back in TMEM (N up to 256: 10 MMAs per K step instead of 28)
//   warps 0-7 FP64: read the accumulators straight into DMMA fragment layout (tcgen05.ld 16x256b: a warp owns 16 rows
//             of the pass), T = Psi - 2^2e sum; pass 0: 64 x 64 factorisation in shared memory (8-wide panels, DMMA);
//             all passes: X L_jj^T = T solved in registers against L_jj and its inverted 8 x 8 diagonal blocks, X cut
//             into digits and stored (per-warp staging: 16 rows x 64 bytes per plane and store).
// The MMAs of pass p+1 run while pass p is solved (the accumulators are free as soon as they are in registers).
// The two right-hand sides (augmented rows 1^T, y^T -> a = L^-1 1, d = L^-1 y) are ordinary rows of the last pass with
// their own, growing scale: their FP64 values stay in the workspace and all their digits are cut again whenever the
// exponent of a row grows.  FP64 L itself is never stored (the diagonal is, for ln|Psi|).
#include <cuda.h>

#include <cstdio>
#include <vector>

#include "pk_chol_tile.cuh"
#include "pk_i8.cuh"
#include "pk_internal.h"

namespace pk {
namespace ci8 {

using namespace i8;
using cta::D_DOUBLES;
using cta::DLD;
using cta::TS;

constexpr int BM = 128;              // rows per pass (TMEM lanes)
constexpr int BN = PK_TILE;          // block column width (TMEM columns per level)
constexpr int BK = 64;               // bytes of K per pipeline chunk: one 64-byte swizzle row = one block column
constexpr int STAGES = 2;
constexpr int A_SLICE = BM * BK, B_SLICE = BN * BK;
constexpr int STAGE_BYTES = S * (A_SLICE + B_SLICE);  // 84 KB
constexpr int WTHREADS = 256;        // warps 0-7: FP64 workers
constexpr int THREADS = 384;         // + warp 8 (TMA) + warp 9 (MMA, TMEM allocation) + two idle warps: the third warpgroup hands
                                     // its registers to the FP64 warpgroups (setmaxnreg)
constexpr int STG_LD = 80;           // bytes per row of a warp's digit staging (16 rows x 64 bytes; 80: conflict-free 2-byte stores)
constexpr int STG_BYTES = 16 * STG_LD;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + (PK_TILE * TS + D_DOUBLES + PK_TILE) * 8 + 8 * STG_BYTES;
enum : int { AUG_NONE = -4000 };     // exponent of an augmented row that has no digits yet

// PERM: the digit planes are stored with the rows of every group of 8 permuted -- storage row 8 g + p holds logical row
// 8 g + sigma(p), sigma(2 q + e) = q + 4 e -- so that the TMEM columns of a pass (= the rows of row block j as TMA brought
// them) reach the FP64 warps with lane (fr, q) holding the LOGICAL columns q and q + 4 of its row: exactly the two A
// fragments (k-steps) of the next DMMA.  The triangular solve then chains DMMAs without re-laying accumulators out as A
// fragments (1.7 warp shuffles per DMMA before; shuffles issue at one per clock and SM and capped the solve at about half
// the DMMA rate, tools/dmma_shfl_probe.cu).  Rows: TMEM lane fr of a group is logical row sigma(fr), and is written back to
// storage row fr -- slice_store does not change; bytes of a row: position 2 q + e of a group of 8 holds logical column
// q + 4 e, the same convention for every row, so the K sums of the INT8 MMAs are unchanged.
__device__ __forceinline__ int sig(int p) { return (p >> 1) + 4 * (p & 1); }
__device__ __forceinline__ int sig_inv(int l) { return 2 * (l & 3) + (l >> 2); }

__device__ __forceinline__ void wbar() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

// digits of x (already scaled by 2^(56 - e)): the bytes of Y = (X + B) ^ B, B = 0x80 in each of the 7 low bytes, are the
// balanced digits (byte 0 = least significant digit 7, byte 6 = digit 1)
__device__ __forceinline__ unsigned long long digit_bytes(double xs) {
    const double lim = 0.49 * 72057594037927936.0;
    if (!(fabs(xs) < lim)) xs = 0.0;  // NaN / out of range: the candidate is failing anyway
    const unsigned long long B = 0x0080808080808080ull;
    return ((unsigned long long)__double2ll_rn(xs) + B) ^ B;
}

// 64 x 64 factorisation of T (shared memory, stride TS) by the eight FP64 warps; see cta::diag_tile (pk_chol_cta.cu).
// TM (PK_CTA_TIMING builds), ft != nullptr in thread 0 only: cycles of {first diagonal block + barrier, panel scale + barrier,
// block (kb+1, kb+1), pivot chains, barrier after the chain} as warp 0 sees them (the warp is re-converged around every
// reading: a diverged warp takes the slow path of the shuffles)
template <bool QUIET, bool TM>
__device__ __forceinline__ int factor_tile(double* T, double* Dsm, const double* diag0, volatile int* ctl, double piv_tol,
                                           int warp, int lane, long long* ft = nullptr) {
    long long tp = (TM && ft) ? clock64() : 0;
#define PK_FT(i)                             \
    if (TM) {                                \
        __syncwarp();                        \
        if (ft) {                            \
            const long long tn_ = clock64(); \
            ft[i] += tn_ - tp;               \
            tp = tn_;                        \
        }                                    \
        __syncwarp();                        \
    }
    if (warp == 0) {
        const int bad = cta::diag_block_factor(T, Dsm, diag0, 0, piv_tol, lane);
        if (bad && lane == 0) ctl[1] = bad;
    }
    wbar();
    PK_FT(0)
    if (ctl[1]) return ctl[1];
    for (int kb = 0; kb < 8; ++kb) {
        cta::panel_scale_blocks<8>(T, Dsm, kb, warp, lane);
        wbar();
        PK_FT(1)
        if (kb < 7) {
            if (warp == 0) {
                cta::trailing_block(T, kb, kb + 1, kb + 1, lane);
                __syncwarp();
                PK_FT(2)
                const int bad = cta::diag_block_factor(T, Dsm, diag0, kb + 1, piv_tol, lane);
                if (bad && lane == 0) ctl[1] = bad;
                PK_FT(3)
            } else {
                if (QUIET) cta::trailing_update_rest_quiet<8>(T, kb, warp, lane);
                else cta::trailing_update_rest<8>(T, kb, warp, lane);
            }
            wbar();
            PK_FT(4)
            if (ctl[1]) return ctl[1];
        }
    }
#undef PK_FT
    return 0;
}

// accumulators of the pass (7 levels) -> acc = acc - 2^(eA + eB) * sum, in DMMA fragment layout.  sneg[mi] = -2^(eA + eB)
// of the lane's row in row tile mi.
template <int NMI>
__device__ __forceinline__ void drain(double (&acc)[2][8][2], uint32_t tlane, double sneg0, double sneg1) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {  // 16 columns at a time
        long long hi[8], lo[8];
        {
            uint32_t r0[8], r1[8], r2[8], r3[8];
            tmem_ld_frag16(tlane + 0 * BN + g * 16, r0);
            tmem_ld_frag16(tlane + 1 * BN + g * 16, r1);
            tmem_ld_frag16(tlane + 2 * BN + g * 16, r2);
            tmem_ld_frag16(tlane + 3 * BN + g * 16, r3);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i)
                hi[i] = ((long long)(int)r0[i] << 24) + ((long long)(int)r1[i] << 16) + ((long long)(int)r2[i] << 8) +
                        (long long)(int)r3[i];
        }
        {
            uint32_t r4[8], r5[8], r6[8];
            tmem_ld_frag16(tlane + 4 * BN + g * 16, r4);
            tmem_ld_frag16(tlane + 5 * BN + g * 16, r5);
            tmem_ld_frag16(tlane + 6 * BN + g * 16, r6);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) lo[i] = ((long long)(int)r4[i] << 16) + ((long long)(int)r5[i] << 8) + (long long)(int)r6[i];
        }
        // level L has weight 256^-(L+2): hi = sum_{L<4} acc_L 256^(3-L) counts 2^-40, lo = sum_{L>=4} acc_L 256^(6-L) counts 2^-64
#pragma unroll
        for (int nn = 0; nn < 2; ++nn)
#pragma unroll
            for (int mi = 0; mi < NMI; ++mi)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = nn * 4 + mi * 2 + e;
                    const double t = fma((double)lo[i], 5.421010862427522e-20, (double)hi[i] * 9.094947017729282e-13);
                    acc[mi][2 * g + nn][e] = fma(mi ? sneg1 : sneg0, t, acc[mi][2 * g + nn][e]);
                }
    }
}

// X L_jj^T = T in registers: 8-column block cb is multiplied by the inverted diagonal block D_cb (2 DMMAs per row tile)
// and folded into the blocks to its right (2 DMMAs each); L_jj from the factorised tile in shared memory.
// PERM: accumulators hold logical columns (q, q + 4): they ARE the A fragments; the B fragments take row sigma(fr).
template <int NMI, bool PERM>
__device__ __forceinline__ void solve(double (&acc)[2][8][2], const double* Tl, const double* Dsm, int lane) {
    const int fr = lane >> 2, q = lane & 3;
    const int br = PERM ? sig(fr) : fr;  // row of the B operand this lane loads (= column of the product it addresses)
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) {
        const double d0 = Dsm[(cb * 8 + br) * DLD + q], d1 = Dsm[(cb * 8 + br) * DLD + 4 + q];
        double ax[NMI][2];
#pragma unroll
        for (int mi = 0; mi < NMI; ++mi) {
            double a0, a1;
            if (PERM) {
                a0 = acc[mi][cb][0];
                a1 = acc[mi][cb][1];
            } else {
                cta::acc_to_afrag(acc[mi][cb][0], acc[mi][cb][1], lane, a0, a1);
            }
            double x0 = 0.0, x1 = 0.0;
            dmma884(x0, x1, a0, d0);
            dmma884(x0, x1, a1, d1);
            acc[mi][cb][0] = x0;
            acc[mi][cb][1] = x1;
            if (PERM) {
                a0 = x0;
                a1 = x1;
            } else {
                cta::acc_to_afrag(x0, x1, lane, a0, a1);
            }
            ax[mi][0] = cta::negd(a0);
            ax[mi][1] = cta::negd(a1);
        }
#pragma unroll
        for (int gb = cb + 1; gb < 8; ++gb) {
            const double b0 = Tl[(gb * 8 + br) * TS + cb * 8 + q];
            const double b1 = Tl[(gb * 8 + br) * TS + cb * 8 + 4 + q];
#pragma unroll
            for (int mi = 0; mi < NMI; ++mi) {
                dmma884(acc[mi][gb][0], acc[mi][gb][1], ax[mi][0], b0);
                dmma884(acc[mi][gb][0], acc[mi][gb][1], ax[mi][1], b1);
            }
        }
    }
}

// the solved 16 x 64 rows of a warp -> digit planes P[t][row][col0 ..]: per plane the 2-byte pieces of the fragment layout go
// through the warp's staging rows and leave as 16-byte stores (a warp store covers 16 rows x 64 bytes)
__device__ __forceinline__ void slice_store(double (&acc)[2][8][2], double down, unsigned char* stg, int8_t* __restrict__ P,
                                            size_t plane_stride, int np, int row_first, int col0, int lane) {
    const int fr = lane >> 2, q = lane & 3;
    uint32_t lo[2][8][2], hi[2][8][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const unsigned long long Y = digit_bytes(acc[mi][ni][e] * down);
                lo[mi][ni][e] = (uint32_t)Y;
                hi[mi][ni][e] = (uint32_t)(Y >> 32);
            }
    const int srow = lane >> 1, shalf = lane & 1;
    int8_t* gdst = P + (size_t)(row_first + srow) * np + col0 + shalf * 32;
    // The staging round trip of plane t (2-byte stores, warp barrier, 16-byte loads) is latency bound -- the INT8 MMAs of the
    // next pass keep the shared memory busy meanwhile -- so the global stores of plane t wait until the pieces of plane t+1
    // are on their way: the loads have a whole plane's PRMT / STS work to come back.
    uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
#pragma unroll
    for (int t = 0; t < S; ++t) {
        // digit t+1 = byte 6 - t of Y: bytes 4..6 in the high word
        const uint32_t sel = (t < 3) ? (uint32_t)(0x40 + 0x11 * (2 - t)) : (uint32_t)(0x40 + 0x11 * (6 - t));
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                const uint32_t w0 = (t < 3) ? hi[mi][ni][0] : lo[mi][ni][0];
                const uint32_t w1 = (t < 3) ? hi[mi][ni][1] : lo[mi][ni][1];
                const uint32_t pr = __byte_perm(w0, w1, sel);  // byte 0 = digit of column 2q, byte 1 = of column 2q + 1
                *reinterpret_cast<unsigned short*>(stg + (mi * 8 + fr) * STG_LD + ni * 8 + 2 * q) = (unsigned short)pr;
            }
        __syncwarp();
        if (t > 0) {
            *reinterpret_cast<uint4*>(gdst + (size_t)(t - 1) * plane_stride) = v0;
            *reinterpret_cast<uint4*>(gdst + (size_t)(t - 1) * plane_stride + 16) = v1;
        }
        v0 = *reinterpret_cast<const uint4*>(stg + srow * STG_LD + shalf * 32);
        v1 = *reinterpret_cast<const uint4*>(stg + srow * STG_LD + shalf * 32 + 16);
        __syncwarp();
    }
    *reinterpret_cast<uint4*>(gdst + (size_t)(S - 1) * plane_stride) = v0;
    *reinterpret_cast<uint4*>(gdst + (size_t)(S - 1) * plane_stride + 16) = v1;
}

// Augmented rows (1^T -> a, y^T -> d) after their block of column j has been solved (values in acc[0][..], logical rows 0, 1
// of the warp's first row tile): FP64 values to the workspace, exponent of each row = max(previous, what the new block
// needs), digits of the new block -- or of the whole row so far when the exponent grew.
template <bool PERM>
__device__ __forceinline__ void aug_store(double (&acc)[2][8][2], double* __restrict__ M, int8_t* __restrict__ P,
                                          size_t plane_stride, int np, int ld, int col0, volatile int* ctl, int lane) {
    const int fr = lane >> 2, q = lane & 3;
    const int lr = PERM ? sig(fr) : fr;  // logical row of this lane
    double mx = 0.0;
    if (lr < 2) {
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            if (PERM) {
                M[(size_t)(np + lr) * ld + col0 + ni * 8 + q] = acc[0][ni][0];
                M[(size_t)(np + lr) * ld + col0 + ni * 8 + 4 + q] = acc[0][ni][1];
            } else {
                *reinterpret_cast<double2*>(M + (size_t)(np + lr) * ld + col0 + ni * 8 + 2 * q) = make_double2(acc[0][ni][0], acc[0][ni][1]);
            }
            mx = fmax(mx, fmax(fabs(acc[0][ni][0]), fabs(acc[0][ni][1])));
        }
    }
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int srow = PERM ? sig_inv(i) : i;  // storage row (= lane group) of logical row i
        const double m = __shfl_sync(0xffffffffu, mx, 4 * srow);
        const int old = ctl[3 + i];
        const int need = (m > 0.0 && m < CUDART_INF) ? ilogb(m) + 3 : AUG_NONE;  // |x| 2^-need < 0.25
        const int cur = (need > old) ? need : old;
        const double down = ldexp(1.0, 56 - cur);
        const double* src = M + (size_t)(np + i) * ld;
        int8_t* dst = P + (size_t)(np + srow) * np;
        for (int c = ((cur != old) ? 0 : col0) + lane; c < col0 + PK_TILE; c += 32) {
            const unsigned long long Y = digit_bytes(__ldcg(src + c) * down);
            const int pos = PERM ? ((c & ~7) + sig_inv(c & 7)) : c;  // byte of the row that holds logical column c
#pragma unroll
            for (int t = 0; t < S; ++t) dst[(size_t)t * plane_stride + pos] = (int8_t)(Y >> (8 * (S - 1 - t)));
        }
        __syncwarp();
        if (lane == 0) ctl[3 + i] = cur;
    }
}

// TIMING: lane 0 of FP64 warp 3 (rows of both kinds of pass) accumulates clock64() per phase into dbg[8 * blockIdx.x ..]:
// {prefetch + wait for the MMAs, drain, tile factorisation, solve, digits, end of column, candidates}
template <bool TIMING, bool PERM>
__global__ void __launch_bounds__(THREADS, 1)
chol_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int C, int nt, int ld,
               size_t elems, size_t plane_stride, size_t cand_stride, double piv_tol, double* __restrict__ ws,
               int8_t* __restrict__ planes, int32_t* __restrict__ info, unsigned int* __restrict__ queue, long long* __restrict__ dbg, int l2_mode, unsigned int* __restrict__ started) {
    long long tmv[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;
    const bool tmw = TIMING && threadIdx.x == 96;
    long long ftv[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // TIMING: thread 0's view of the tile factorisation (factor_tile)
    const bool tm0 = TIMING && threadIdx.x == 0;
    long long tjv[24] = {};                       // TIMING: tile factorisation per block column (thread 96)
#define PK_TM(i)                          \
    if (tmw) {                            \
        const long long tn_ = clock64();  \
        tmv[i] += tn_ - tprev;            \
        tprev = tn_;                      \
    }
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024 - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023)) & 1023);
    double* T = reinterpret_cast<double*>(smem + STAGES * STAGE_BYTES);
    double* Dsm = T + PK_TILE * TS;
    double* diag0 = Dsm + D_DOUBLES;
    unsigned char* stg_all = reinterpret_cast<unsigned char*>(diag0 + PK_TILE);
    __shared__ unsigned long long full[STAGES], empty[STAGES], tmem_full, tmem_empty, col_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int ctl_s[8];  // 0 candidate, 1 failing column of the tile, 2 abort, 3 / 4 exponents of the augmented rows
    volatile int* ctl = ctl_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int np = nt * PK_TILE;
    const bool quiet = (l2_mode & 4) != 0;  // warp 4 leaves the scheduler to warp 0's pivot chain (factor_tile)
    if (tid == 0) {
        // ragged-first pipeline (pk_nll_batch): the Psi kernel of the whole waves is released by a stream wait on this
        // counter, i.e. once every CTA of this short wave holds its SM
        if (started) atomicAdd(started, 1u);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(&tmem_full, 1);
        mbar_init(&tmem_empty, WTHREADS);
        mbar_init(&col_done, WTHREADS);
        mbar_fence_init();
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    unsigned it = 0;    // K chunks issued / consumed so far (TMA and MMA threads)
    unsigned mp = 0;    // passes with MMAs so far (MMA thread and FP64 warps)
    unsigned ccol = 0;  // block columns finished before this candidate (TMA thread)
    // next candidate of this CTA (every thread of the CTA calls it: two CTA barriers), -1 = the queue is empty
    auto next_candidate = [&]() -> int {
        for (;;) {
            __syncthreads();
            if (tid == 0) {
                ctl_s[0] = (int)atomicAdd(queue, 1u);
                ctl_s[2] = 0;
                ctl_s[3] = ctl_s[4] = AUG_NONE;
            }
            __syncthreads();
            const int c = ctl_s[0];
            if (c >= C) return -1;
            if (info[c] == 0) return c;  // candidates that failed before (NaN hyper-parameters) are skipped
        }
    };
    // 12 warps at 168 registers fill the register file, and the FP64 warps spill at 168: the producer warpgroup (TMA, MMA and
    // two idle warps) hands its registers over.  Each side runs its own candidate loop behind its setmaxnreg.
    if (warp >= 8) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
      for (;;) {
        const int cand = next_candidate();
        if (cand < 0) break;
        if (warp == 8) {
            // ---------------- TMA producer ----------------
            if (lane == 0) {
                // L2 policy of the digit tiles: row block j (B) is read again by every pass of the block column and should
                // stay, the rows of a pass (A) come back only a whole block column later and should go first
                const unsigned long long polA = l2_policy_evict_first(), polB = l2_policy_evict_last();
                for (int j = 1; j < nt; ++j) {
                    const int npass = (np + PK_AUG_ROWS - j * PK_TILE + BM - 1) / BM;
                    bool waited = false;
                    for (int p = 0; p < npass; ++p) {
                        const int r0 = j * PK_TILE + p * BM;
                        for (int c = 0; c < j; ++c, ++it) {
                            // the digits of block column j-1 are complete.  The augmented rows wait for it with ALL their
                            // chunks: when a row's exponent grows, block column j-1 cuts the digits of its older columns again.
                            if (!waited && (c == j - 1 || r0 + BM > np)) {
                                mbar_wait(&col_done, (ccol + (unsigned)(j - 1)) & 1u);
                                fence_proxy_async();
                                waited = true;
                            }
                            const int st = (int)(it % STAGES);
                            if (it >= STAGES) mbar_wait(empty + st, ((it / STAGES) - 1) & 1);
                            if (ctl[2]) {  // the candidate failed: keep the handshakes, skip the copies
                                mbar_arrive(full + st);
                                continue;
                            }
                            unsigned char* sa = smem + st * STAGE_BYTES;
                            mbar_arrive_expect_tx(full + st, (unsigned)STAGE_BYTES);
                            if (l2_mode & 1) tma_load_4d_hint(sa, &tmA, c * BK, r0, 0, cand, full + st, polA);
                            else tma_load_4d(sa, &tmA, c * BK, r0, 0, cand, full + st);
                            if (l2_mode & 2) tma_load_4d_hint(sa + S * A_SLICE, &tmB, c * BK, j * PK_TILE, 0, cand, full + st, polB);
                            else tma_load_4d(sa + S * A_SLICE, &tmB, c * BK, j * PK_TILE, 0, cand, full + st);
                        }
                    }
                }
                ccol += (unsigned)nt;
            }
        } else if (warp == 9) {
            // ---------------- MMA issuer ----------------
            if (lane == 0) {
                for (int j = 1; j < nt; ++j) {
                    const int npass = (np + PK_AUG_ROWS - j * PK_TILE + BM - 1) / BM;
                    for (int p = 0; p < npass; ++p, ++mp) {
                        if (mp > 0) {  // the FP64 warps hold the previous pass's sums in registers
                            mbar_wait(&tmem_empty, (mp - 1) & 1);
                            tc_fence_after();
                        }
                        for (int c = 0; c < j; ++c, ++it) {
                            const int st = (int)(it % STAGES);
                            mbar_wait(full + st, (it / STAGES) & 1);
                            tc_fence_after();
                            if (!ctl[2]) {
                                const uint32_t sa = (unsigned)__cvta_generic_to_shared(smem + st * STAGE_BYTES);
                                const uint32_t sb = sa + S * A_SLICE;
#pragma unroll
                                for (int kk = 0; kk < BK / 32; ++kk) {
#pragma unroll
                                    for (int t = 0; t < S; ++t) {
                                        // digit t+1 of the pass's rows against digits 1 .. 7-t of row block j: levels t .. 6
                                        const uint64_t ad = smem_desc_k_sw64(sa + t * A_SLICE + kk * 32);
                                        const uint32_t first = (c == 0 && kk == 0 && t == 0) ? 0u : 1u;
                                        const int cnt = S - t, n1 = (cnt < 4) ? cnt : 4;
                                        umma_i8(tmem_base + t * BN, ad, smem_desc_k_sw64(sb + kk * 32), idesc_i8(true, true, BM, BN * n1), first);
                                        if (cnt > 4)
                                            umma_i8(tmem_base + (t + 4) * BN, ad, smem_desc_k_sw64(sb + 4 * B_SLICE + kk * 32),
                                                    idesc_i8(true, true, BM, BN * (cnt - 4)), first);
                                    }
                                }
                            }
                            umma_commit(empty + st);
                        }
                        umma_commit(&tmem_full);
                    }
                }
            }
        }
      }
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
      for (;;) {
        const int cand = next_candidate();
        if (cand < 0) break;
        {
            // ---------------- FP64 warps ----------------
            const int quarter = warp & 3, half = warp >> 2;
            const int wrow = quarter * 32 + half * 16;  // this warp's 16 rows of a pass = its 16 TMEM lanes
            const int fr = lane >> 2, q = lane & 3;
            const uint32_t tlane = tmem_base + ((uint32_t)wrow << 16);
            unsigned char* stg = stg_all + warp * STG_BYTES;
            double* M = ws + (size_t)cand * elems;
            int8_t* P = planes + (size_t)cand * cand_stride;
            // digit scale of the candidate: |L_ik| <= sqrt(Psi_ii) = sqrt(Psi_00) for every row
            const double dmax = M[0];
            const int e = (dmax > 0.0 && dmax < CUDART_INF) ? ilogb(sqrt(dmax) * 1.0078125) + 2 : 2;
            const double down = ldexp(1.0, 56 - e);
            const double sc2 = -ldexp(1.0, 2 * e);
            bool dead = false;  // the candidate failed: only the handshakes are kept up
            if (tmw) {
                tprev = clock64();
                tmv[6] += 1;
            }
            // this lane's 16 values of one row of a Psi tile (8 column blocks x 2)
            auto load_tile_rows = [&](double (&dst)[8][2], const double* src) {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) {
                    if (PERM) {
                        dst[ni][0] = src[ni * 8 + q];
                        dst[ni][1] = src[ni * 8 + 4 + q];
                    } else {
                        const double2 v = *reinterpret_cast<const double2*>(src + ni * 8 + 2 * q);
                        dst[ni][0] = v.x;
                        dst[ni][1] = v.y;
                    }
                }
            };
            for (int j = 0; j < nt; ++j) {
                const int row0 = j * PK_TILE;
                const int npass = (np + PK_AUG_ROWS - row0 + BM - 1) / BM;
                if (!dead && tid < PK_TILE) diag0[tid] = M[(size_t)(row0 + tid) * ld + row0 + tid];
                for (int p = 0; p < npass; ++p) {
                    const int rbase = row0 + p * BM + wrow;
                    const bool is_aug = (rbase == np);
                    const int nmi = dead ? 0 : ((rbase >= np + PK_AUG_ROWS) ? 0 : (is_aug ? 1 : 2));
                    const int lrow = PERM ? sig(fr) : fr;  // logical row of this lane within its row tile
                    double acc[2][8][2];
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) {
                        if (mi < nmi) load_tile_rows(acc[mi], M + (size_t)(rbase + mi * 8 + lrow) * ld + row0);
                        else {
#pragma unroll
                            for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                        }
                    }
                    if (j > 0) {
                        mbar_wait(&tmem_full, mp & 1);
                        tc_fence_after();
                        PK_TM(0)
                        if (nmi == 2) {
                            drain<2>(acc, tlane, sc2, sc2);
                        } else if (nmi == 1) {
                            // augmented rows: row fr carries its own exponent (rows 2 .. 7 are zero)
                            const int ea = (lrow < 2) ? ctl[3 + lrow] : AUG_NONE;
                            drain<1>(acc, tlane, (ea == AUG_NONE) ? 0.0 : -ldexp(1.0, ea + e), 0.0);
                        }
                        tc_fence_before();
                        mbar_arrive(&tmem_empty);
                        ++mp;
                        PK_TM(1)
                    }
                    if (dead) continue;
                    const bool in_tile = (p == 0 && wrow < PK_TILE);
                    if (p == 0) {
                        if (in_tile) {
#pragma unroll
                            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                                for (int ni = 0; ni < 8; ++ni) {
                                    double* dstT = T + (wrow + mi * 8 + lrow) * TS + ni * 8;
                                    if (PERM) {
                                        dstT[q] = acc[mi][ni][0];
                                        dstT[4 + q] = acc[mi][ni][1];
                                    } else {
                                        *reinterpret_cast<double2*>(dstT + 2 * q) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                                    }
                                }
                        }
                        if (tid == 0) ctl[1] = 0;
                        wbar();
                        const int bad = quiet ? factor_tile<true, TIMING>(T, Dsm, diag0, ctl, piv_tol, warp, lane, tm0 ? ftv : nullptr)
                                              : factor_tile<false, TIMING>(T, Dsm, diag0, ctl, piv_tol, warp, lane, tm0 ? ftv : nullptr);
                        if (bad) {
                            if (tid == 0) {
                                info[cand] = row0 + bad;  // dpotrf convention: order of the failing minor
                                ctl[2] = 1;
                            }
                            dead = true;
                            continue;
                        }
                        if (tid < PK_TILE) M[(size_t)(row0 + tid) * ld + row0 + tid] = T[tid * TS + tid];  // L_ii: ln|Psi| (nll_epilogue_kernel)
                        if (tmw && j < 24) tjv[j] += clock64() - tprev;
                        PK_TM(2)
                    }
                    if (in_tile || nmi == 0) continue;
                    if (nmi == 2) {
                        solve<2, PERM>(acc, T, Dsm, lane);
                        PK_TM(3)
                        slice_store(acc, down, stg, P, plane_stride, np, rbase, row0, lane);
                        PK_TM(4)
                    } else {
                        solve<1, PERM>(acc, T, Dsm, lane);
                        aug_store<PERM>(acc, M, P, plane_stride, np, ld, row0, ctl, lane);
                    }
                }
                // the digits of block column j are in global memory: order them before the TMA reads, then let the producer go
                fence_proxy_async();
                mbar_arrive(&col_done);
                wbar();  // L_jj, D and the exponents of the augmented rows are free / visible for the next column
                PK_TM(5)
            }
        }
      }
    }
#undef PK_TM
    if (tmw) {
        for (int i = 0; i < 8; ++i) dbg[40 * blockIdx.x + i] = tmv[i];
        for (int i = 0; i < 24; ++i) dbg[40 * blockIdx.x + 16 + i] = tjv[i];
    }
    if (tm0) {
        for (int i = 0; i < 8; ++i) dbg[40 * blockIdx.x + 8 + i] = ftv[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base));
}

static cudaError_t encode_planes4(CUtensorMap* tm, const void* base, uint64_t np, uint64_t rows, uint64_t C, uint32_t box_rows) {
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
    const cuuint64_t gdim[4] = {(cuuint64_t)np, (cuuint64_t)rows, (cuuint64_t)S, (cuuint64_t)C};
    const cuuint64_t gstride[3] = {(cuuint64_t)np, (cuuint64_t)np * rows, (cuuint64_t)np * rows * S};
    const cuuint32_t box[4] = {(cuuint32_t)BK, box_rows, (cuuint32_t)S, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (r == CUDA_SUCCESS) ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace ci8

// bytes of digit planes one candidate needs (7 planes of (np + 8) rows x np bytes)
size_t cholesky_i8_plane_bytes(const TiledShape& s) { return (size_t)i8::S * (size_t)(s.np + PK_AUG_ROWS) * (size_t)s.np; }

cudaError_t launch_cholesky_i8(pk_handle* h, const TiledShape& s, int C, double* ws, int8_t* planes, int32_t* info) {
    if (s.with_inverse || !planes) return cudaErrorInvalidValue;
    const int nt = s.np / PK_TILE;
    const size_t rowsP = (size_t)s.np + PK_AUG_ROWS;
    const size_t plane_stride = rowsP * s.np, cand_stride = plane_stride * i8::S;
    const size_t smem = (size_t)ci8::SMEM_BYTES + 1024;
    static bool attr_done_dev[PK_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
    cudaError_t e;
    if (!attr_done) {
        const void* kernels[4] = {(const void*)ci8::chol_i8_kernel<false, false>, (const void*)ci8::chol_i8_kernel<false, true>,
                                  (const void*)ci8::chol_i8_kernel<true, false>, (const void*)ci8::chol_i8_kernel<true, true>};
        for (const void* kfn : kernels) {
            e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        attr_done = true;
    }
    CUtensorMap tmA, tmB;
    e = ci8::encode_planes4(&tmA, planes, (uint64_t)s.np, rowsP, (uint64_t)C, ci8::BM);
    if (e != cudaSuccess) return e;
    e = ci8::encode_planes4(&tmB, planes, (uint64_t)s.np, rowsP, (uint64_t)C, ci8::BN);
    if (e != cudaSuccess) return e;
    unsigned int* queue = h->d_queue + h->queue_slot;
    e = cudaMemsetAsync(queue, 0, sizeof(unsigned int), h->stream);
    if (e != cudaSuccess) return e;
    const int grid = (C < h->sm_count) ? C : h->sm_count;
    unsigned int* started = h->i8_started;
    h->i8_started = nullptr;
    const bool perm = h->i8_perm != 0;  // permuted digit-plane layout (shuffle-free solve); 0 = the first layout (PK_I8_PERM)
    if (h->env_cta_timing) {  // diagnostic (PK_CTA_TIMING): per-phase cycle counts of one FP64 warp, printed to stderr
        long long* dbg = nullptr;
        cudaMalloc((void**)&dbg, sizeof(long long) * 40 * grid);
        cudaMemset(dbg, 0, sizeof(long long) * 40 * grid);
        auto kfn = perm ? ci8::chol_i8_kernel<true, true> : ci8::chol_i8_kernel<true, false>;
        kfn<<<grid, ci8::THREADS, smem, h->stream>>>(tmA, tmB, C, nt, s.ld, s.elems(), plane_stride, cand_stride, h->pivot_tol, ws, planes,
                                                     info, queue, dbg, h->i8_l2_mode | (h->i8_quiet ? 4 : 0), nullptr);
        cudaStreamSynchronize(h->stream);
        std::vector<long long> host(40 * (size_t)grid);
        cudaMemcpy(host.data(), dbg, sizeof(long long) * host.size(), cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        double sum[40] = {};
        for (int b = 0; b < grid; ++b)
            for (int i = 0; i < 40; ++i) sum[i] += (double)host[40 * b + i];
        const double nc = sum[6] > 0 ? sum[6] : 1;
        fprintf(stderr,
                "[pk i8 timing] perm %d, grid %d, candidates %.0f; cycles per candidate (warp 3): wait for MMAs %.0f, drain %.0f, tile "
                "factorisation %.0f, solve %.0f, digits %.0f, end of column %.0f\n",
                (int)perm, grid, sum[6], sum[0] / nc, sum[1] / nc, sum[2] / nc, sum[3] / nc, sum[4] / nc, sum[5] / nc);
        fprintf(stderr,
                "[pk i8 timing] tile factorisation seen by warp 0: first diagonal block %.0f, panel scale + barrier %.0f, block (kb+1, kb+1) "
                "%.0f, pivot chains %.0f, barrier after the chain %.0f\n",
                sum[8] / nc, sum[9] / nc, sum[10] / nc, sum[11] / nc, sum[12] / nc);
        fprintf(stderr, "[pk i8 timing] tile factorisation per block column (its MMAs of pass 1 run meanwhile):");
        for (int j = 0; j < nt && j < 24; ++j) fprintf(stderr, " %.0f", sum[16 + j] / nc);
        fprintf(stderr, "\n");
    } else {
        auto kfn = perm ? ci8::chol_i8_kernel<false, true> : ci8::chol_i8_kernel<false, false>;
        kfn<<<grid, ci8::THREADS, smem, h->stream>>>(tmA, tmB, C, nt, s.ld, s.elems(), plane_stride, cand_stride, h->pivot_tol, ws, planes,
                                                     info, queue, nullptr, h->i8_l2_mode | (h->i8_quiet ? 4 : 0), started);
    }
    h->launches++;
    return cudaGetLastError();
}

}  // namespace pk
