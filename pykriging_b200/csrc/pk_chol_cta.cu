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
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pk_chol_tile.cuh"
#include "pk_internal.h"

namespace pk {

namespace cta {

constexpr int THREADS = 256;
constexpr int KC = 16;             // K chunk per pipeline stage: one row of a stage = 16 doubles = 128 bytes
constexpr int STAGES = 4;
constexpr int BM = 128;            // rows per panel pass (8 warps x 16 rows)
// A stage = two TMA boxes of 64 rows x 128 bytes, B stage = one; dense rows, 32-byte atoms XOR-swizzled with the
// row (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): element (r, c) of a box lives at r * 16 + (((c >> 2) ^ (r & 3)) << 2)
// + (c & 3).  A DMMA fragment load reads, per half-warp, 4 consecutive k of 4 consecutive rows: with the 32-byte
// atom of row r rotated by r & 3 those are 16 different 8-byte bank pairs (the 16-byte swizzle of plain
// SWIZZLE_128B maps rows r and r ^ 1 onto the same 32 bytes: measured 1.1e9 bank conflicts per launch).
constexpr int BOX_ROWS = PK_TILE;
constexpr int BOX_DOUBLES = BOX_ROWS * KC;  // 8 KB
constexpr int A_STAGE = BM * KC;
constexpr int B_STAGE = PK_TILE * KC;
constexpr int PIPE_DOUBLES = STAGES * (A_STAGE + B_STAGE);
constexpr int SMEM_DOUBLES = PIPE_DOUBLES + D_DOUBLES + PK_TILE + 2 + STAGES + STAGES / 2;  // + diag0[64] + control words + mbarriers + release counters
constexpr int SMEM_ALIGN = 1024;   // the swizzle pattern is a function of the shared address: stages on 1 KB boundaries
static_assert((PK_TILE + 32) * TS <= STAGES * A_STAGE, "the diagonal tile reuses the pipeline buffers");


// ---- gangs: several CTAs share ONE candidate (few candidates, large n) ---------------------------------------
// With fewer candidates than SMs one CTA per candidate leaves most of the GPU idle (n = 1024, 21 candidates of an
// SLSQP stencil: 2.2 ms whatever the count).  A gang of G CTAs (consecutive block indices, all resident: the grid
// never exceeds 2 CTAs per SM) then works on the same matrix: the K range of the diagonal-tile product is split in
// segments of one block column, dealt round-robin over the ranks; the 72 x 64 partial sums are exchanged through a
// global scratch and every rank adds them in SEGMENT order, so all ranks hold bit-identical tiles -- whatever G is --
// and factorise them redundantly (no broadcast of L_jj or of the inverted diagonal blocks is needed); the panel's
// 64-row passes are dealt round-robin as well.  Two gang barriers per block column (a monotone counter in global
// memory: release fence + atomic add, acquire spin).  A candidate's result is therefore the same in a gang of any
// size: SLSQP's single-candidate evaluations and its stencil populations agree bit for bit.
struct Gang {
    int G, rank;
    unsigned* ctr;       // this gang's barrier counter (zeroed by the launcher)
    double* part;        // this gang's partial-sum scratch: (nt - 1) slots of GANG_PART doubles (one per K segment)
    unsigned epoch;      // arrivals expected at the next barrier
};
constexpr int GANG_PART = (PK_TILE + PK_AUG_ROWS) * PK_TILE;  // 72 x 64

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// returns false if the other ranks did not arrive within ~4e9 cycles (never on a healthy launch; keeps a broken
// one from hanging the GPU)
__device__ __forceinline__ bool gang_barrier(Gang& g, int* ctl) {
    g.epoch += (unsigned)g.G;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(g.ctr, 1u);
        const long long t0 = clock64();
        int ok = 1;
        while (ld_acquire_gpu(g.ctr) < g.epoch) {
            if (clock64() - t0 > (1LL << 32)) {
                ok = 0;
                break;
            }
        }
        __threadfence();
        ctl[2] = ok;
    }
    __syncthreads();
    fence_proxy_async();  // the other ranks' stores are read through TMA from here on
    return ctl[2] != 0;
}

// ---- task mode: dependencies that are only needed LATE in a task ------------------------------------------------
// A task of cta::chol_task_kernel starts as soon as the inputs of its FIRST pipeline chunks exist and lets the thread
// that issues the TMA copies wait for the rest: the product chunks of a panel pass only read finished block columns,
// the pass needs L_jj and the inverted diagonal blocks (diagonal tile j) only for its four solve chunks; a diagonal
// tile accumulates block columns 0 .. j-2 while the pass that produces column j-1 of its rows is still running.  This
// takes the DMMA main loops off the critical path D(j) -> P(j, first) -> D(j+1) -> ... of a candidate, which is what
// bounds the run time when there are fewer candidates than CTA slots.
__device__ __forceinline__ int ld_acquire_gpu_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_s32(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
struct LateDep {
    const int* f0;   // first counter (never null)
    int need0;
    const int* f1;   // second counter or null
    int need1;
    const int* dead; // candidate's info word: non-zero = the candidate failed, nothing will ever arrive
    int* fault;      // sticky launch-wide flag: a wait timed out somewhere, the launch is being abandoned
    int from;        // first K chunk (within the task) that needs the counters
};
// Spin (one thread) until the counters have arrived; gives up when the candidate is dead or the launch abandoned --
// the task then runs on stale data, which nobody will read.  Ends with the proxy fence that orders this thread's TMA
// copies behind the producers' generic-proxy stores.
__device__ __forceinline__ void late_wait(const LateDep& d) {
    const long long t0 = clock64();
    for (;;) {
        if (ld_acquire_gpu_s32(d.f0) >= d.need0 && (d.f1 == nullptr || ld_acquire_gpu_s32(d.f1) >= d.need1)) break;
        if (ld_acquire_gpu_s32(d.dead) != 0 || ld_acquire_gpu_s32(d.fault) != 0) break;
        if (clock64() - t0 > (1LL << 32)) {
            atomicOr(d.fault, 1);
            break;
        }
        __nanosleep(32);
    }
    fence_proxy_async();
}

// acc(8*MI x 64 per warp) += A(8*MI x 16) * B(64 x 16)^T for one stage (the products are accumulated with a
// PLUS sign and subtracted from the original tile in one go when it arrives with the solve chunks: no sign
// flips in the hot loop -- every non-DMMA instruction costs the scheduler's dispatch port a cycle the tensor
// pipe does not get).  MI (rows of 8 this warp owns in
// the pass) is a template parameter, selected by a warp-uniform branch: a DMMA that is merely
// predicated off still occupies its issue slots on the tensor pipe (measured: profiles/).
// Per-lane swizzled column offsets of the four k-steps of a chunk: lane (fr, fc) reads column 4i + fc of a row
// whose index is fr mod 8:  soff[i] = ((i ^ (fr & 3)) << 2) + fc.
struct Frag {
    int soff[4];
};
__device__ __forceinline__ Frag make_frag(int lane) {
    const int fr = lane >> 2, fc = lane & 3;
    Frag f;
#pragma unroll
    for (int i = 0; i < 4; ++i) f.soff[i] = ((i ^ (fr & 3)) << 2) + fc;
    return f;
}

template <int MI, int HALF>
__device__ __forceinline__ void gemm_consume(double (&acc)[2][8][2], const double* As, const double* Bs, int wrow,
                                             int lane, const Frag& fg) {
    const int fr = lane >> 2;
    const double* a_s = As + (wrow + fr) * KC;
    const double* b_s = Bs + fr * KC;
#pragma unroll
    for (int i = 2 * HALF; i < 2 * HALF + 2; ++i) {
        double a[MI], b[8];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) a[mi] = a_s[mi * 8 * KC + fg.soff[i]];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) b[ni] = b_s[ni * 8 * KC + fg.soff[i]];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
}

// Triangular solve, pipeline chunk T (columns 16T .. 16T+15 of L_jj are in Bs): for the two 8-column
// blocks cb of the chunk  X_cb = T_cb * D_cb^T,  then  P_gb += X_cb * L_jj[gb, cb]^T  for gb > cb (P_gb = the
// still positive product sum of block gb; it becomes T_gb = tile - P_gb when its own solve chunk arrives).
template <int T, int MI, int half>
__device__ __forceinline__ void trsm_consume(double (&acc)[2][8][2], const double* Bs, const double* Dsm, int lane,
                                             const Frag& fg) {
    const int fr = lane >> 2, q = lane & 3;
    constexpr int cb = 2 * T + half;
    const double d0 = Dsm[(cb * 8 + fr) * DLD + q], d1 = Dsm[(cb * 8 + fr) * DLD + 4 + q];
    double ax[MI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        double a0, a1;
        acc_to_afrag(acc[mi][cb][0], acc[mi][cb][1], lane, a0, a1);
        double x0 = 0.0, x1 = 0.0;
        dmma884(x0, x1, a0, d0);
        dmma884(x0, x1, a1, d1);
        acc[mi][cb][0] = x0;
        acc[mi][cb][1] = x1;
        acc_to_afrag(x0, x1, lane, ax[mi][0], ax[mi][1]);
    }
#pragma unroll
    for (int gb = cb + 1; gb < 8; ++gb) {
        const double b0 = Bs[(gb * 8 + fr) * KC + fg.soff[2 * half]];
        const double b1 = Bs[(gb * 8 + fr) * KC + fg.soff[2 * half + 1]];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            dmma884(acc[mi][gb][0], acc[mi][gb][1], ax[mi][0], b0);
            dmma884(acc[mi][gb][0], acc[mi][gb][1], ax[mi][1], b1);
        }
    }
}

// The original tile enters through the pipeline as well: the A stage of solve chunk T holds columns
// 16T .. 16T+15 of M[rows, j] (still Psi / identity / right-hand sides), which are added to the two
// accumulator blocks that chunk is about to solve (tile - accumulated products).  No accumulator pre-load from
// global memory.
template <int T, int MI, int half>
__device__ __forceinline__ void add_tile_chunk(double (&acc)[2][8][2], const double* As, int wrow, int lane) {
    const int fr = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        const double2 v =
            *reinterpret_cast<const double2*>(As + (wrow + mi * 8 + fr) * KC + (((half * 2 + (q >> 1)) ^ (fr & 3)) << 2) +
                                              2 * (q & 1));
        acc[mi][2 * T + half][0] = v.x - acc[mi][2 * T + half][0];
        acc[mi][2 * T + half][1] = v.y - acc[mi][2 * T + half][1];
    }
}

// one half (8 columns) of solve chunk T
template <int T, int half>
__device__ __forceinline__ void solve_half(double (&acc)[2][8][2], const double* As, const double* Bs,
                                           const double* Dsm, int wrow, int lane, int mi_count, const Frag& fg) {
    if (mi_count == 2) {
        add_tile_chunk<T, 2, half>(acc, As, wrow, lane);
        trsm_consume<T, 2, half>(acc, Bs, Dsm, lane, fg);
    } else if (mi_count == 1) {
        add_tile_chunk<T, 1, half>(acc, As, wrow, lane);
        trsm_consume<T, 1, half>(acc, Bs, Dsm, lane, fg);
    }
}

// half a pipeline chunk of a pass: product chunk (ck < nk_gemm) or solve chunk ck - nk_gemm
template <int HALF>
__device__ __forceinline__ void consume_half(double (&acc)[2][8][2], const double* as, const double* bs,
                                             const double* Dsm, int ck, int nk_gemm, int wrow, int lane, int mi_count,
                                             const Frag& fg) {
    if (ck < nk_gemm) {
        if (mi_count == 2) gemm_consume<2, HALF>(acc, as, bs, wrow, lane, fg);
        else if (mi_count == 1) gemm_consume<1, HALF>(acc, as, bs, wrow, lane, fg);
    } else {
        switch (ck - nk_gemm) {
            case 0: solve_half<0, HALF>(acc, as, bs, Dsm, wrow, lane, mi_count, fg); break;
            case 1: solve_half<1, HALF>(acc, as, bs, Dsm, wrow, lane, mi_count, fg); break;
            case 2: solve_half<2, HALF>(acc, as, bs, Dsm, wrow, lane, mi_count, fg); break;
            default: solve_half<3, HALF>(acc, as, bs, Dsm, wrow, lane, mi_count, fg); break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Panel of block column j: every row below the diagonal tile (rows [(j+1)*64, rows)), 128 rows per
// pass, as ONE continuous stream of pipeline chunks.  A pass consists of 4j product chunks
//     acc -= M[rows, 16c..] * L[j, 16c..]^T
// followed by four solve chunks (A stage = original tile columns, B stage = L_jj columns).  The
// cp.async pipeline never drains between passes: while a pass is being solved and stored, the first
// chunks of the next pass are already in flight.
// ------------------------------------------------------------------------------------------------
// tl != nullptr (diagnostic build only): lanes 0 of warps 0 and 4 accumulate clock64() deltas of the loop's
// sections {barrier wait, early produce, first half, late produce, second half, store} into tl[0..5] / tl[8..13].
template <bool GANG>
__device__ __forceinline__ void panel_column(double* pipe, double* Dsm_w, double* __restrict__ M, int ld, int j,
                                             int rows, const CUtensorMap* tmap, int cidx, unsigned long long* bars,
                                             int* rel, unsigned& seq, const Frag& fg, int G_, int rank_, int inv_base_,
                                             long long* tl = nullptr, int first_ = -1, const LateDep* late = nullptr,
                                             const double* __restrict__ dinv_g = nullptr) {
    const double* Dsm = Dsm_w;
    // GANG == false: one CTA per candidate -- the gang arithmetic below folds away at compile time
    const int G = GANG ? G_ : 1, rank = GANG ? rank_ : 0, inv_base = GANG ? inv_base_ : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, q = lane & 3;
    double* As = pipe;
    double* Bs = pipe + STAGES * A_STAGE;
    const int brow0 = j * PK_TILE;
    // first row of the panel: everything below the diagonal tile, or (task mode) the first row of ONE pass whose
    // last row is rows - 1
    const int first = (first_ >= 0) ? first_ : (j + 1) * PK_TILE;
    const int nk_gemm = j * (PK_TILE / KC);
    const int nkt = nk_gemm + PK_TILE / KC;
    // passes: 128 rows each for a CTA on its own; in a gang 64-row passes dealt round-robin over the ranks
    const int prow = (G > 1) ? PK_TILE : BM;
    // pk_fit (gang mode only): the identity block that becomes L^-T starts at row inv_base = np + 8; its tiles 0 .. j
    // are further passes of block column j (tiles beyond j are still zero in every column <= j)
    const int nbelow = (rows - first + prow - 1) / prow;
    const int npass = nbelow + ((inv_base > 0) ? j + 1 : 0);
    const int nblk = (rank < npass) ? (npass - rank + G - 1) / G : 0;
    // first row of this rank's i-th pass
    auto pass_row = [&](int i) {
        const int P = rank + i * G;
        return (P < nbelow) ? first + P * prow : inv_base + (P - nbelow) * PK_TILE;
    };
    const int total = nblk * nkt;
    // Producer: ONE thread per chunk.  A chunk is three TMA tile copies (two 64-row boxes of A, one of B, 16
    // columns each) described by a 3-D tensor map over (column, row, candidate); completion is counted in bytes on
    // the stage's mbarrier.  Chunk c of the column goes to stage (seq + c) % STAGES; the first STAGES chunks are
    // issued here, chunk c + STAGES by whichever warp is the LAST to release chunk c (a counter per stage):
    // nobody ever waits for a stage to become free.
    bool late_ok = (late == nullptr);
    auto issue = [&](int c, int k_idx, int r0) {
        if (!late_ok && k_idx >= late->from) {  // task mode: the solve chunks need diagonal tile j
            late_wait(*late);
            late_ok = true;
        }
        const int ps = (int)((seq + (unsigned)c) % STAGES);
        unsigned long long* bar = bars + ps;
        const int nbox = (G == 1 && rows - r0 > BOX_ROWS) ? 2 : 1;
        // task mode: the inverted diagonal blocks diagonal tile j's task left in the global scratch travel with the FIRST
        // solve chunk of the task (one more bulk copy counted on the same mbarrier): whoever has waited for that chunk
        // has them -- no extra barrier, no exposed L2 latency
        const bool with_dinv = (dinv_g != nullptr) && (c == nk_gemm);
        mbar_arrive_expect_tx(bar, (unsigned)((nbox + 1) * BOX_DOUBLES * sizeof(double)) + (with_dinv ? (unsigned)(D_DOUBLES * sizeof(double)) : 0u));
        if (with_dinv) bulk_g2s(Dsm_w, dinv_g, (unsigned)(D_DOUBLES * sizeof(double)), bar);
        double* as = As + ps * A_STAGE;
        tma_load_3d(as, tmap, k_idx * KC, r0, cidx, bar);
        if (nbox == 2) tma_load_3d(as + BOX_DOUBLES, tmap, k_idx * KC, r0 + BOX_ROWS, cidx, bar);
        tma_load_3d(Bs + ps * B_STAGE, tmap, k_idx * KC, brow0, cidx, bar);
    };
    if (tid == 0) {
        for (int c = 0; c < STAGES && c < total; ++c) issue(c, c % nkt, pass_row(c / nkt));
    }
    // position of the chunk STAGES ahead of the one being consumed (tracked by every warp: any of them may issue it)
    int pk = STAGES % nkt, p_pass = STAGES / nkt, p_row0 = pass_row(STAGES / nkt);
    double acc[2][8][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    // A 128-row pass gives every warp 16 rows (two DMMA row tiles); a 64-row pass (odd number of tiles left)
    // gives every warp 8 rows, so all four schedulers keep two warps either way.
    auto pass_shape = [&](int r0, int& wrow, int& mic) {
        const bool full = (G == 1) && (rows - r0) >= BM;
        wrow = warp * (full ? 16 : 8);
        mic = full ? 2 : 1;
    };
    int cb = 0, ck = 0, cs = (int)(seq % STAGES);
    unsigned cpar = (seq / STAGES) & 1u;
    int row0 = pass_row(0), wrow, mi_count;
    pass_shape(row0, wrow, mi_count);
    long long tprev = 0, tacc[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool tlw = tl != nullptr && lane == 0 && (warp & 3) == 0;
    if (tlw) tprev = clock64();
#define PK_TL(i)                                  \
    if (tlw) {                                    \
        const long long tn_ = clock64();          \
        tacc[i] += tn_ - tprev;                   \
        tprev = tn_;                              \
    }
    // No CTA barrier in this loop: a warp waits for its data (full[stage]) and releases the stage when its
    // fragment loads are done (one count per warp on the stage's release counter); the last warp to release a
    // stage refills it.  Warps on different schedulers may be up to STAGES - 1 chunks apart.
    for (int g = 0; g < total; ++g) {
        mbar_wait(bars + cs, cpar);  // chunk g has landed
        PK_TL(0)
        PK_TL(1)
        const double* as = As + cs * A_STAGE;
        const double* bs = Bs + cs * B_STAGE;
        int* rcnt = rel + cs;
        if (++cs == STAGES) {
            cs = 0;
            cpar ^= 1u;
        }
        consume_half<0>(acc, as, bs, Dsm, ck, nk_gemm, wrow, lane, mi_count, fg);
        PK_TL(2)
        PK_TL(3)
        consume_half<1>(acc, as, bs, Dsm, ck, nk_gemm, wrow, lane, mi_count, fg);
        __syncwarp();
        if (lane == 0 && atomicAdd(rcnt, 1) == THREADS / 32 - 1) {  // last warp to release this stage: refill it
            *rcnt = 0;
            if (g + STAGES < total) issue(g + STAGES, pk, p_row0);
        }
        if (++pk == nkt) {
            pk = 0;
            p_row0 = pass_row(++p_pass);
        }
        PK_TL(4)
        if (tlw) tacc[6] += 1;
        if (++ck == nkt) {
            // pass complete: store the solved rows, reset the accumulators, move to the next pass
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                if (mi < mi_count) {
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni)
                        *reinterpret_cast<double2*>(M + (size_t)(row0 + wrow + mi * 8 + fr) * ld + brow0 + ni * 8 +
                                                    2 * q) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                }
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            }
            ck = 0;
            ++cb;
            row0 = pass_row(cb);
            pass_shape(row0, wrow, mi_count);
            PK_TL(5)
        }
    }
#undef PK_TL
    if (tlw) {
#pragma unroll
        for (int i = 0; i < 7; ++i) tl[(warp >> 2) * 8 + i] += tacc[i];
    }
    seq += (unsigned)total;
    fence_proxy_async();  // the stores above are read by bulk copies (async proxy) in the columns to come
    __syncthreads();      // stages drained; the stores of this column are visible to the whole CTA
}

// ------------------------------------------------------------------------------------------------
// Diagonal tile update T = Psi_jj - L[j,0:j] L[j,0:j]^T, LOWER 8x8 blocks only (36 of 64).  Warp w owns
// block row br(w) = w for w < 4 and 11 - w otherwise, so the two warps that share a scheduler carry
// 9 blocks together and the four tensor pipes stay balanced.  A and B are the same 64 rows: only the
// B stages are loaded.  NI = br + 1 is dispatched by a warp-uniform switch (no predicated-off DMMAs).
// ------------------------------------------------------------------------------------------------
template <int NI, bool AUG, int HALF>
__device__ __forceinline__ void diag_consume(double (&acc)[8][2], double (&accA)[2][2], const double* Bs,
                                             const double* Aaug, int br, int lane, const Frag& fg) {
    const int fr = lane >> 2;
    const double* a_s = Bs + (br * 8 + fr) * KC;
    const double* b_s = Bs + fr * KC;
#pragma unroll
    for (int i = 2 * HALF; i < 2 * HALF + 2; ++i) {
        const int kk = fg.soff[i];
        const double a = a_s[kk];
        double b[NI];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) b[ni] = b_s[ni * 8 * KC + kk];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) dmma884(acc[ni][0], acc[ni][1], a, b[ni]);
        if (AUG) {
            // the 8 augmented rows (1^T, y^T, ...) x column blocks 2*br, 2*br+1 (br == warp < 4 here)
            const double aa = Aaug[fr * KC + kk];
            const double b0 = b_s[(2 * br) * 8 * KC + kk], b1 = b_s[(2 * br + 1) * 8 * KC + kk];
            dmma884(accA[0][0], accA[0][1], aa, b0);
            dmma884(accA[1][0], accA[1][1], aa, b1);
        }
    }
}

template <bool GANG>
__device__ __forceinline__ bool diag_gemm(double* T, double* pipe, const double* __restrict__ M, int ld, int j,
                                          int np, const CUtensorMap* tmap, int cidx, unsigned long long* bars,
                                          int* rel, unsigned& seq, const Frag& fg, Gang* gang, int* ctl,
                                          const LateDep* late = nullptr) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, q = lane & 3;
    const int br = (warp < 4) ? warp : 11 - warp;
    double* Bs = pipe + STAGES * A_STAGE;
    const int row0 = j * PK_TILE;
    const int nk_all = j * (PK_TILE / KC);
    // A gang splits the K range in SEGMENTS of one block column (4 chunks): segment c < j belongs to rank c % G, its
    // partial sum goes to slot c of the gang's scratch, and every rank adds the slots in segment order -- so the tile
    // (and with it every bit of the factor) does not depend on the gang size.
    constexpr int SEG = PK_TILE / KC;
    const int G = GANG ? gang->G : 1, rank = GANG ? gang->rank : 0;
    const int nseg = (G > 1) ? ((rank < j) ? (j - rank + G - 1) / G : 0) : 0;  // segments of this rank
    const int nk = (G > 1) ? nseg * SEG : nk_all;
    // chunk i of this rank -> K chunk index
    auto kchunk = [&](int i) { return (G > 1) ? (rank + (i / SEG) * G) * SEG + (i % SEG) : i; };
    // two TMA boxes per chunk (L[j, :] and the box that starts at the 8 augmented rows; its 56 rows beyond the
    // matrix are zero-filled by the TMA unit); issued like the panel's chunks (see panel_column)
    bool late_ok = (late == nullptr);
    auto issue = [&](int c) {
        if (!late_ok && c >= late->from) {  // task mode: the last block column of this row block is still being produced
            late_wait(*late);
            late_ok = true;
        }
        const int ps = (int)((seq + (unsigned)c) % STAGES);
        unsigned long long* bar = bars + ps;
        mbar_arrive_expect_tx(bar, (unsigned)(2 * BOX_DOUBLES * sizeof(double)));
        const int kc = kchunk(c);
        tma_load_3d(Bs + ps * B_STAGE, tmap, kc * KC, row0, cidx, bar);
        tma_load_3d(pipe + ps * A_STAGE, tmap, kc * KC, np, cidx, bar);
    };
    if (tid == 0) {
        for (int c = 0; c < STAGES && c < nk; ++c) issue(c);
    }
    int cs = (int)(seq % STAGES);
    unsigned cpar = (seq / STAGES) & 1u;
    double acc[8][2];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) acc[ni][0] = acc[ni][1] = 0.0;
    double accA[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    // slot of the gang's scratch <- this thread's accumulators (minus the ORIGINAL tile for slot 0), accumulators
    // back to zero.  Nobody reads the original after the barrier: a faster rank may already be overwriting it with L_jj.
    auto flush_slot = [&](int slot) {
        double* dst = gang->part + (size_t)slot * GANG_PART;
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            if (ni <= br) {
                double2 v = make_double2(acc[ni][0], acc[ni][1]);
                if (slot == 0) {
                    const double2 o = *reinterpret_cast<const double2*>(M + (size_t)(row0 + br * 8 + fr) * ld + row0 + ni * 8 + 2 * q);
                    v.x -= o.x;
                    v.y -= o.y;
                }
                *reinterpret_cast<double2*>(dst + (br * 8 + fr) * PK_TILE + ni * 8 + 2 * q) = v;
            }
            acc[ni][0] = acc[ni][1] = 0.0;
        }
        if (warp < 4) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double2 v = make_double2(accA[i][0], accA[i][1]);
                if (slot == 0) {
                    const double2 o = *reinterpret_cast<const double2*>(M + (size_t)(np + fr) * ld + row0 + (2 * warp + i) * 8 + 2 * q);
                    v.x -= o.x;
                    v.y -= o.y;
                }
                *reinterpret_cast<double2*>(dst + (PK_TILE + fr) * PK_TILE + (2 * warp + i) * 8 + 2 * q) = v;
                accA[i][0] = accA[i][1] = 0.0;
            }
        }
    };
    for (int it = 0; it < nk; ++it) {
        mbar_wait(bars + cs, cpar);
        const double* bs = Bs + cs * B_STAGE;
        const double* aa = pipe + cs * A_STAGE;
        int* rcnt = rel + cs;
        if (++cs == STAGES) {
            cs = 0;
            cpar ^= 1u;
        }
#define PK_DIAG_HALF(HH)                                                                        \
        switch (warp) {                                                                         \
            case 0: diag_consume<1, true, HH>(acc, accA, bs, aa, br, lane, fg); break;              \
            case 1: diag_consume<2, true, HH>(acc, accA, bs, aa, br, lane, fg); break;              \
            case 2: diag_consume<3, true, HH>(acc, accA, bs, aa, br, lane, fg); break;              \
            case 3: diag_consume<4, true, HH>(acc, accA, bs, aa, br, lane, fg); break;              \
            case 4: diag_consume<8, false, HH>(acc, accA, bs, aa, br, lane, fg); break;             \
            case 5: diag_consume<7, false, HH>(acc, accA, bs, aa, br, lane, fg); break;             \
            case 6: diag_consume<6, false, HH>(acc, accA, bs, aa, br, lane, fg); break;             \
            default: diag_consume<5, false, HH>(acc, accA, bs, aa, br, lane, fg); break;            \
        }
        PK_DIAG_HALF(0)
        PK_DIAG_HALF(1)
        __syncwarp();
        if (lane == 0 && atomicAdd(rcnt, 1) == THREADS / 32 - 1) {
            *rcnt = 0;
            if (it + STAGES < nk) issue(it + STAGES);
        }
        if (G > 1 && (it % SEG) == SEG - 1) flush_slot(rank + (it / SEG) * G);  // segment complete
    }
#undef PK_DIAG_HALF
    seq += (unsigned)nk;
    if (G > 1) {
        if (j == 0 && rank == 0) flush_slot(0);  // no products yet: slot 0 = -original
        if (!gang_barrier(*gang, ctl)) return false;
        const int nslot = (j > 0) ? j : 1;
        for (int r = 0; r < nslot; ++r) {
            const double* pr = gang->part + (size_t)r * GANG_PART;
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                if (ni <= br) {
                    const double2 v = __ldcg(reinterpret_cast<const double2*>(pr + (br * 8 + fr) * PK_TILE + ni * 8 + 2 * q));
                    acc[ni][0] += v.x;
                    acc[ni][1] += v.y;
                }
            }
            if (warp < 4) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double2 v = __ldcg(reinterpret_cast<const double2*>(pr + (PK_TILE + fr) * PK_TILE + (2 * warp + i) * 8 + 2 * q));
                    accA[i][0] += v.x;
                    accA[i][1] += v.y;
                }
            }
        }
        // T = -(sum of the slots) = original - products
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
            if (ni <= br)
                *reinterpret_cast<double2*>(T + (br * 8 + fr) * TS + ni * 8 + 2 * q) = make_double2(-acc[ni][0], -acc[ni][1]);
        if (warp < 4) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
                *reinterpret_cast<double2*>(T + (PK_TILE + fr) * TS + (2 * warp + i) * 8 + 2 * q) =
                    make_double2(-accA[i][0], -accA[i][1]);
        }
        return true;
    }
    __syncthreads();  // the stages are drained: T (aliasing the A stages) may be written
    // T = original tile - accumulated products
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
        if (ni <= br) {
            const double2 v =
                *reinterpret_cast<const double2*>(M + (size_t)(row0 + br * 8 + fr) * ld + row0 + ni * 8 + 2 * q);
            *reinterpret_cast<double2*>(T + (br * 8 + fr) * TS + ni * 8 + 2 * q) =
                make_double2(v.x - acc[ni][0], v.y - acc[ni][1]);
        }
    }
    if (warp < 4) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double2 v =
                *reinterpret_cast<const double2*>(M + (size_t)(np + fr) * ld + row0 + (2 * warp + i) * 8 + 2 * q);
            *reinterpret_cast<double2*>(T + (PK_TILE + fr) * TS + (2 * warp + i) * 8 + 2 * q) =
                make_double2(v.x - accA[i][0], v.y - accA[i][1]);
        }
    }
    return true;
}

// Diagonal tile of block column j.  Returns 0 or the tile-local failing column (CTA uniform).
template <bool GANG>
__device__ __forceinline__ int diag_tile(double* pipe, double* Dsm, double* diag0, int* ctl, double* __restrict__ M,
                                         int ld, int j, int np, double piv_tol, const CUtensorMap* tmap, int cidx,
                                         unsigned long long* bars, int* rel, unsigned& seq, const Frag& fg,
                                         long long* tm, Gang* gang, const LateDep* late = nullptr) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = j * PK_TILE;
    if (tid < PK_TILE) diag0[tid] = M[(size_t)(row0 + tid) * ld + row0 + tid];
    double* T = pipe;
    long long t0 = 0;
    if (tm) t0 = clock64();
    if (!diag_gemm<GANG>(T, pipe, M, ld, j, np, tmap, cidx, bars, rel, seq, fg, gang, ctl, late)) return -1;  // gang barrier timed out
    if (tid == 0) ctl[1] = 0;
    __syncthreads();
    if (tm) {
        const long long t1 = clock64();
        tm[0] += t1 - t0;
        t0 = t1;
    }
    // Panel kb: (a) its 8x8 diagonal block is factorised by warp 0 -- a chain of 8 dependent pivots, latency
    // bound; (b) all warps scale the blocks below; (c) trailing update.  (a) of panel kb+1 runs DURING (c) of
    // panel kb: warp 0 updates block (kb+1, kb+1) first and walks its chain while warps 1..7 do the rest.
    if (warp == 0) {
        const int bad = diag_block_factor(T, Dsm, diag0, 0, piv_tol, lane);
        if (bad && lane == 0) ctl[1] = bad;
    }
    __syncthreads();
    if (ctl[1]) return ctl[1];
    for (int kb = 0; kb < 8; ++kb) {
        long long ta = 0;
        if (tm) ta = clock64();
        panel_scale_blocks(T, Dsm, kb, warp, lane);
        __syncthreads();
        if (tm) {
            const long long tb = clock64();
            tm[6] += tb - ta;
            ta = tb;
        }
        if (kb < 7) {
            if (warp == 0) {
                trailing_block(T, kb, kb + 1, kb + 1, lane);
                __syncwarp();
                const int bad = diag_block_factor(T, Dsm, diag0, kb + 1, piv_tol, lane);
                if (bad && lane == 0) ctl[1] = bad;
            } else {
                trailing_update_rest(T, kb, warp, lane);
            }
            __syncthreads();
            if (ctl[1]) return ctl[1];
        }
        if (tm) tm[5] += clock64() - ta;
    }
    if (tm) {
        const long long t1 = clock64();
        tm[1] += t1 - t0;
        t0 = t1;
    }
    for (int i = tid; i < PK_TILE * (PK_TILE / 2); i += THREADS) {
        const int rr = i >> 5, c2 = (i & 31) * 2;
        const double2 t = *reinterpret_cast<const double2*>(T + rr * TS + c2);
        *reinterpret_cast<double2*>(M + (size_t)(row0 + rr) * ld + row0 + c2) =
            make_double2((c2 <= rr) ? t.x : 0.0, (c2 + 1 <= rr) ? t.y : 0.0);
    }
    {   // the solved augmented rows: a = L^-1 1, d = L^-1 y for this block column
        const int rr = tid >> 5, c2 = (tid & 31) * 2;
        *reinterpret_cast<double2*>(M + (size_t)(np + rr) * ld + row0 + c2) =
            *reinterpret_cast<const double2*>(T + (PK_TILE + rr) * TS + c2);
    }
    // T (written through the generic proxy) is about to be overwritten by bulk copies, and L_jj / the solved
    // augmented rows are about to be READ by them: order both before the async proxy
    fence_proxy_async();
    __syncthreads();  // L_jj is in global memory and D in shared memory for every thread of the CTA
    if (tm) tm[2] += clock64() - t0;
    return 0;
}

// TIMING: thread 0 accumulates clock64() per phase {diag product, 64x64 factorisation, inverse + copy-out,
// panel column, candidates} into dbg[8 * blockIdx.x ..] (diagnostic build of the same kernel).
template <bool TIMING, bool GANG>
__global__ void __launch_bounds__(THREADS, 2)
chol_cta_kernel(const __grid_constant__ CUtensorMap tmap, int C, int nt, int rows, int ld, size_t elems, double piv_tol,
                double* __restrict__ ws, int32_t* __restrict__ info, unsigned int* __restrict__ queue,
                long long* __restrict__ dbg, int alias, int G, unsigned* __restrict__ gang_ctr,
                double* __restrict__ gang_part, int inv_base) {
    long long tmv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long* tm = TIMING ? tmv : nullptr;
    extern __shared__ __align__(16) double smem_raw[];
    // stages on a 1 KB boundary (the launcher allocates the slack): the TMA swizzle is a function of the address
    double* smem = smem_raw + ((SMEM_ALIGN - ((unsigned)__cvta_generic_to_shared(smem_raw) & (SMEM_ALIGN - 1))) &
                               (SMEM_ALIGN - 1)) / sizeof(double);
    double* pipe = smem;
    double* Dsm = pipe + PIPE_DOUBLES;
    double* diag0 = Dsm + D_DOUBLES;
    int* ctl = reinterpret_cast<int*>(diag0 + PK_TILE);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(diag0 + PK_TILE + 2);  // one mbarrier per stage
    int* rel = reinterpret_cast<int*>(bars + STAGES);                                      // one release counter per stage
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + s, 1);  // the issuing thread's arrive.expect_tx + the bytes of the chunk
            rel[s] = 0;
        }
        mbar_fence_init();
    }
    const Frag fg = make_frag(tid & 31);
    unsigned seq = 0;  // pipeline chunks issued == consumed so far (stage = seq % STAGES, parity = seq / STAGES & 1)
    if constexpr (GANG) {
        // ---- gang mode: G consecutive CTAs factorise candidate blockIdx.x / G together (see struct Gang) ----
        const int cand = blockIdx.x / G;
        if (cand >= C) return;
        Gang gang{G, (int)(blockIdx.x % G), gang_ctr + cand, gang_part + (size_t)cand * (nt - 1) * GANG_PART, 0u};
        fence_proxy_async();
        __syncthreads();
        if (info[cand] != 0) return;  // uniform over the gang
        double* M = ws + (size_t)cand * elems;
        for (int j = 0; j < nt; ++j) {
            const int bad = diag_tile<true>(pipe, Dsm, diag0, ctl, M, ld, j, nt * PK_TILE, piv_tol, &tmap, cand, bars, rel, seq, fg,
                                            nullptr, &gang);
            if (bad) {  // every rank holds the same tile and sees the same failure (or the same timed-out barrier)
                if (tid == 0) info[cand] = (bad > 0) ? j * PK_TILE + bad : -1;
                break;
            }
            if (j + 1 < nt || inv_base > 0) {
                panel_column<true>(pipe, Dsm, M, ld, j, nt * PK_TILE, &tmap, cand, bars, rel, seq, fg, G, gang.rank, inv_base);
                // rows of block column j written by the other ranks are read from the next column on
                if (!gang_barrier(gang, ctl)) {
                    if (tid == 0) info[cand] = -1;
                    break;
                }
            }
        }
        return;
    }
    for (;;) {
        fence_proxy_async();  // a failed factorisation leaves generic-proxy writes in the stage buffers
        __syncthreads();
        if (tid == 0) ctl[0] = (int)atomicAdd(queue, 1u);
        __syncthreads();
        const int cand = ctl[0];
        if (cand >= C) break;
        if (info[cand] != 0) continue;
        const int cidx = alias ? (cand & 7) : cand;  // alias: timing experiment only (results are garbage)
        double* M = ws + (size_t)cidx * elems;
        for (int j = 0; j < nt; ++j) {
            const int bad = diag_tile<false>(pipe, Dsm, diag0, ctl, M, ld, j, nt * PK_TILE, piv_tol, &tmap, cidx, bars, rel, seq, fg, tm,
                                             nullptr);
            if (bad && !alias) {
                if (tid == 0) info[cand] = j * PK_TILE + bad;  // dpotrf convention: order of the failing minor
                break;
            }
            long long t0 = 0;
            if (TIMING) t0 = clock64();
            if (j + 1 < nt) panel_column<false>(pipe, Dsm, M, ld, j, nt * PK_TILE, &tmap, cidx, bars, rel, seq, fg, 1, 0, 0, (TIMING && blockIdx.x == 0) ? dbg + 8 * gridDim.x : nullptr);
            if (TIMING) tmv[3] += clock64() - t0;
        }
        if (TIMING) tmv[4] += 1;
    }
    if (TIMING && tid == 0) {
        for (int i = 0; i < 8; ++i) dbg[8 * blockIdx.x + i] = tmv[i];
    }
}


// ------------------------------------------------------------------------------------------------
// TASK mode: the factorisation of every candidate as a stream of tile tasks.
//
// One CTA per candidate leaves SMs idle whenever the population is not a multiple of the resident CTA slots
// (128 candidates per GPU when a population of 1024 is sharded over 8 GPUs: 128 of 296 slots; the reference's
// default population of 300; the 150-particle regression swarm at n = 4096) and ties a candidate's latency to ONE
// CTA.  Here the persistent CTAs pull TASKS from a global ticket counter instead:
//     D(j)        diagonal tile of block column j: product, 64 x 64 factorisation, the 8 augmented rows
//     P(j, r, b)  one panel pass of block column j: the b (1 or 2) 64-row blocks starting at row block r
// Ticket t belongs to candidate (t % S) of its cohort and to entry t / S of the per-shape task list built by the
// launcher (task_schedule below): every task depends only on tasks of the SAME candidate that come EARLIER in the
// list, so a ticket's dependencies have always been claimed by a running CTA (or are done) and waiting for them can
// never deadlock, resident or not.  Completion is published per candidate in global memory -- ddone = number of
// finished diagonal tiles, rdone[rb] = number of finished block columns of row block rb (release store after a
// device-scope fence; the waiter acquires, then orders its TMA reads with fence.proxy.async).  The inverted 8 x 8
// diagonal blocks a panel pass needs travel through a global scratch (4 KB per diagonal tile).
// Every tile is produced by exactly the arithmetic of the one-CTA-per-candidate kernel above (same K order, same
// accumulators), so the results are bit-identical to it for any population size, cohort size or pass partition.
// ------------------------------------------------------------------------------------------------
struct Task {
    int type, j, rb, nb;  // type 0 = D(j), 1 = P(j, rb, nb)
};
constexpr int DINV_DOUBLES = D_DOUBLES;  // the eight inverted diagonal blocks of a tile, as they lie in shared memory
constexpr int TASK_MAX_BLOCKS = 30;      // row blocks of one panel task (one lane of warp 0 watches each)


// warp 0: lane 0 watches *a0 >= need0 (a0 may be null), lanes 1..nb watch arows[lane - 1] >= need_r.
// Returns 1 = ready, 0 = the candidate failed meanwhile (skip), -1 = timed out (never on a healthy launch).
__device__ __forceinline__ int task_wait(const int* a0, int need0, const int* arows, int nb, int need_r,
                                         const int32_t* info_c, const int* fault, int lane) {
    const int* addr = (lane == 0) ? a0 : arows + (lane - 1);
    const int need = (lane == 0) ? need0 : need_r;
    const bool watch = (lane == 0) ? (a0 != nullptr) : (lane <= nb);
    const long long t0 = clock64();
    for (;;) {
        const int v = watch ? ld_acquire_gpu_s32(addr) : need;
        if (__all_sync(0xffffffffu, v >= need)) return 1;
        int bad = 0, dead = 0;
        if (lane == 0) {
            bad = ld_acquire_gpu_s32(reinterpret_cast<const int*>(info_c));
            dead = ld_acquire_gpu_s32(fault);
        }
        bad = __shfl_sync(0xffffffffu, bad, 0);
        dead = __shfl_sync(0xffffffffu, dead, 0);
        if (bad != 0) return 0;
        if (dead != 0 || clock64() - t0 > (1LL << 32)) return -1;  // somebody timed out already: the launch is being abandoned
        __nanosleep(64);
    }
}

__global__ void __launch_bounds__(THREADS, 2)
chol_task_kernel(const __grid_constant__ CUtensorMap tmap, int C, int nt, int ld, size_t elems, double piv_tol,
                 double* __restrict__ ws, int32_t* __restrict__ info, unsigned int* __restrict__ queue,
                 const Task* __restrict__ tasks, int ntasks, int S, int skew, int* __restrict__ flags, int fstride,
                 double* __restrict__ dinv, int* __restrict__ fault) {
    extern __shared__ __align__(16) double smem_raw[];
    double* smem = smem_raw + ((SMEM_ALIGN - ((unsigned)__cvta_generic_to_shared(smem_raw) & (SMEM_ALIGN - 1))) &
                               (SMEM_ALIGN - 1)) / sizeof(double);
    double* pipe = smem;
    double* Dsm = pipe + PIPE_DOUBLES;
    double* diag0 = Dsm + D_DOUBLES;
    int* ctl = reinterpret_cast<int*>(diag0 + PK_TILE);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(diag0 + PK_TILE + 2);
    int* rel = reinterpret_cast<int*>(bars + STAGES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + s, 1);
            rel[s] = 0;
        }
        mbar_fence_init();
    }
    const Frag fg = make_frag(lane);
    unsigned seq = 0;
    const int np = nt * PK_TILE;
    // tickets: cohorts of S consecutive candidates walk the task list together (ticket = cohort base + step * size
    // + member), one cohort after the other.  Member m is `m % skew` entries BEHIND the step, so that any window of
    // consecutive tickets -- what the resident CTAs are working on at a given moment -- mixes latency-bound diagonal
    // tiles with DMMA-bound panel tasks instead of holding one kind (steps before a member's first / after its last
    // entry are null tickets).  Dependencies still point to earlier tickets: entry i' < i of a member lies in an earlier step.
    const int nsteps = ntasks + skew - 1;
    const long long total = (long long)C * nsteps;
    for (;;) {
        fence_proxy_async();  // a failed or skipped task leaves generic-proxy writes in the stage buffers
        __syncthreads();
        if (tid == 0) {
            const unsigned t = atomicAdd(queue, 1u);
            ctl[0] = (int)t;
            ctl[3] = ld_acquire_gpu_s32(fault);
        }
        __syncthreads();
        const long long t = (unsigned)ctl[0];
        if (t >= total || ctl[3] != 0) break;  // fault: a dependency wait timed out somewhere -- abandon the launch (the host fails loudly)
        const int cohort = (int)(t / ((long long)S * nsteps));
        const int csize = min(S, C - cohort * S);
        const int within = (int)(t - (long long)cohort * S * nsteps);
        const int member = within % csize;
        const int entry = within / csize - member % skew;
        if (entry < 0 || entry >= ntasks) continue;  // null ticket
        const int cand = cohort * S + member;
        const Task tk = tasks[entry];
        int* fl = flags + (size_t)cand * fstride;  // [0] = finished diagonal tiles, [1 + rb] = finished block columns of row block rb
        const int j = tk.j;
        // what the FIRST chunks of the task read (everything else: LateDep, waited for by the thread that issues the copies)
        if (warp == 0) {
            int ok = 1;
            if (tk.type == 0) {
                // block columns 0 .. j-2 of row block j and of the augmented rows
                if (j >= 2) ok = task_wait(fl, j - 1, fl + 1 + j, 1, j - 1, info + cand, fault, lane);
            } else if (j >= 1) {
                // block columns 0 .. j-1 of the pass's row blocks and of row block j (the B operand)
                ok = task_wait(fl + 1 + j, j, fl + 1 + tk.rb, tk.nb, j, info + cand, fault, lane);
            }
            if (lane == 0) ctl[2] = ok;
        }
        __syncthreads();
        const int ok = ctl[2];
        if (ok <= 0) {
            if (ok < 0 && tid == 0) {
                atomicCAS(reinterpret_cast<int*>(info + cand), 0, -1);
                atomicOr(fault, 1);
                __threadfence();
            }
            continue;
        }
        fence_proxy_async();  // the producers' (generic-proxy) stores are read through TMA from here on
        double* M = ws + (size_t)cand * elems;
        double* dv = dinv + ((size_t)cand * nt + j) * DINV_DOUBLES;
        if (tk.type == 0) {
            // the last block column (K chunks 4(j-1) ..) waits for the pass that produces it and for diagonal tile j-1
            const LateDep late{fl, j, fl + 1 + j, j, reinterpret_cast<const int*>(info + cand), fault, 4 * (j - 1)};
            const int bad = diag_tile<false>(pipe, Dsm, diag0, ctl, M, ld, j, np, piv_tol, &tmap, cand, bars, rel, seq, fg,
                                             nullptr, nullptr, (j >= 1) ? &late : nullptr);
            if (bad) {
                if (tid == 0) {
                    atomicCAS(reinterpret_cast<int*>(info + cand), 0, j * PK_TILE + bad);  // dpotrf convention; the first failure stays
                    __threadfence();
                }
                continue;
            }
            for (int i = tid; i < DINV_DOUBLES; i += THREADS) dv[i] = Dsm[i];  // as it lies in shared memory (padded rows)
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                st_release_gpu_s32(fl, j + 1);
            }
        } else {
            // the four solve chunks need L_jj and the inverted diagonal blocks: diagonal tile j
            const LateDep late{fl, j + 1, nullptr, 0, reinterpret_cast<const int*>(info + cand), fault, 4 * j};
            const int r0 = tk.rb * PK_TILE;
            panel_column<false>(pipe, Dsm, M, ld, j, r0 + tk.nb * PK_TILE, &tmap, cand, bars, rel, seq, fg, 1, 0, 0, nullptr, r0,
                                &late, dv);
            // A task publishes its row blocks when it ends.  (Publishing every pass of a multi-pass task on its own --
            // a device-scope fence per warp and pass -- was measured: nothing at 300 candidates, 4 % slower at 1024.)
            if (tid < tk.nb) {
                __threadfence();
                st_release_gpu_s32(fl + 1 + tk.rb + tid, j + 1);
            }
        }
    }
}

}  // namespace cta

// 3-D tensor map over the candidate workspace: (column, row, candidate), boxes of 16 columns x 64 rows,
// 128-byte swizzle span with 32-byte atoms.  cuTensorMapEncodeTiled comes from the driver through the runtime's entry-point query (no
// link against libcuda).
static cudaError_t make_ws_tensor_map(CUtensorMap* tmap, double* ws, const TiledShape& s, int C) {
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
    const cuuint64_t gdim[3] = {(cuuint64_t)s.ld, (cuuint64_t)s.rows, (cuuint64_t)C};
    const cuuint64_t gstride[2] = {(cuuint64_t)s.ld * sizeof(double), (cuuint64_t)s.elems() * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)cta::KC, (cuuint32_t)cta::BOX_ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, ws, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (r == CUDA_SUCCESS) ? cudaSuccess : cudaErrorInvalidValue;
}

// Gang size for a population of Ceff candidates (the whole population when this call is a shard): as many CTAs per
// candidate as 2 per SM allow, at most one per 64-row pass of the first panel.  Automatic mode wants at least 3
// ranks: with 2 the exchange and the barriers cost more than the second CTA brings (profiles/variant_sweep_small_r01.jsonl:
// n = 1024, 128 candidates: 2.30 ms in pairs, 2.20 ms alone; 96 candidates in gangs of 3: 1.76 against 2.19).
int cholesky_gang_size(const pk_handle* h, const TiledShape& s, int64_t Ceff, bool forced) {
    if (Ceff > h->sm_count) return 1;
    const int nt = s.np / PK_TILE;
    int64_t G = (2 * (int64_t)h->sm_count) / (Ceff > 0 ? Ceff : 1);
    if (G > nt - 1) G = nt - 1;
    return (G < (forced ? 2 : 3)) ? 1 : (int)G;
}

cudaError_t launch_cholesky_cta(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info) {
    const int nt = s.np / PK_TILE;
    const size_t smem = sizeof(double) * (size_t)cta::SMEM_DOUBLES + cta::SMEM_ALIGN;
    CUtensorMap tmap;
    {
        cudaError_t e = make_ws_tensor_map(&tmap, ws, s, C);
        if (e != cudaSuccess) return e;
    }
    static bool attr_done_dev[PK_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
    if (!attr_done) {
        cudaError_t e =
            cudaFuncSetAttribute(cta::chol_cta_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(cta::chol_cta_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(cta::chol_cta_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    unsigned int* queue = h->d_queue + h->queue_slot;
    cudaError_t e = cudaMemsetAsync(queue, 0, sizeof(unsigned int), h->stream);
    if (e != cudaSuccess) return e;
    // few candidates: gangs of G CTAs per candidate (chol_variant 3 forces gangs, 2 forbids them)
    int G = 1;
    {
        const int64_t Ceff = (h->hint_C > C) ? h->hint_C : (int64_t)C;
        if (h->chol_variant != 2 || s.with_inverse) G = cholesky_gang_size(h, s, Ceff, h->chol_variant == 3 || s.with_inverse);
        if (h->env_gang > 0) {  // experiment knob (PK_CHOL_GANG)
            G = h->env_gang;
            if (G > nt - 1) G = (nt - 1 > 0) ? nt - 1 : 1;
        }
        if ((int64_t)C * G > 2 * (int64_t)h->sm_count) G = 1;  // all CTAs of a gang must be resident
    }
    if (s.with_inverse && G == 1) return cudaErrorInvalidValue;  // the L^-T rows are passes of the gang kernel only
    if (G > 1) {
        const size_t ctr_bytes = (sizeof(unsigned) * (size_t)h->sm_count + 255) & ~(size_t)255;
        // one slot per K segment and candidate; sized for the largest population that still runs in gangs, so that a
        // varying candidate count (GA generation, then SLSQP stencils, then single evaluations) never reallocates
        const size_t part_bytes = sizeof(double) * (size_t)h->sm_count * (nt - 1) * cta::GANG_PART;
        if (ctr_bytes + part_bytes > h->gang_bytes) {
            e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) return e;
            if (h->d_gang) cudaFree(h->d_gang);
            h->d_gang = nullptr;
            h->gang_bytes = 0;
            e = cudaMalloc((void**)&h->d_gang, ctr_bytes + part_bytes);
            if (e != cudaSuccess) return e;
            h->gang_bytes = ctr_bytes + part_bytes;
        }
        e = cudaMemsetAsync(h->d_gang, 0, ctr_bytes, h->stream);
        if (e != cudaSuccess) return e;
        unsigned* ctr = reinterpret_cast<unsigned*>(h->d_gang);
        double* part = reinterpret_cast<double*>(reinterpret_cast<char*>(h->d_gang) + ctr_bytes);
        cta::chol_cta_kernel<false, true><<<C * G, cta::THREADS, smem, h->stream>>>(tmap, C, nt, s.rows, s.ld, s.elems(), h->pivot_tol,
                                                                              ws, info, queue, nullptr, 0, G, ctr, part,
                                                                              s.with_inverse ? s.np + PK_AUG_ROWS : 0);
        h->launches++;
        return cudaGetLastError();
    }
    int per_sm = 2;
    per_sm = h->env_ctas_per_sm;  // experiment knob (PK_CHOL_CTAS_PER_SM)
    const int grid = (C < per_sm * h->sm_count) ? C : per_sm * h->sm_count;
    if (h->env_cta_timing) {  // diagnostic: per-phase cycle counts of the persistent kernel, printed to stderr
        long long* dbg = nullptr;
        cudaMalloc((void**)&dbg, sizeof(long long) * (8 * grid + 16));
        cudaMemset(dbg, 0, sizeof(long long) * (8 * grid + 16));
        cta::chol_cta_kernel<true, false><<<grid, cta::THREADS, smem, h->stream>>>(tmap, C, nt, s.rows, s.ld, s.elems(), h->pivot_tol,
                                                                            ws, info, queue, dbg, h->env_cta_alias ? 1 : 0,
                                                                            1, nullptr, nullptr, 0);
        cudaStreamSynchronize(h->stream);
        std::vector<long long> host(8 * (size_t)grid + 16);
        cudaMemcpy(host.data(), dbg, sizeof(long long) * host.size(), cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        double sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int b2 = 0; b2 < grid; ++b2)
            for (int i = 0; i < 8; ++i) sum[i] += (double)host[8 * b2 + i];
        for (int w = 0; w < 2; ++w) {
            const long long* t = host.data() + 8 * (size_t)grid + 8 * w;
            const double it = (double)(t[6] > 0 ? t[6] : 1);
            fprintf(stderr,
                    "[pk cta timeline] CTA 0 warp %d: %.0f chunks; cycles per chunk: barrier %.0f, early produce %.0f, half0 %.0f, "
                    "late produce %.0f, half1 %.0f, pass store (per chunk) %.0f\n",
                    4 * w, it, t[0] / it, t[1] / it, t[2] / it, t[3] / it, t[4] / it, t[5] / it);
        }
        fprintf(stderr,
                "[pk cta timing] grid %d, candidates %.0f; cycles per candidate: diag product %.0f, factorise 64x64 %.0f, "
                "copy-out %.0f, panel column %.0f; factorise = 8x8 block %.0f + scale %.0f + trailing %.0f\n",
                grid, sum[4], sum[0] / sum[4], sum[1] / sum[4], sum[2] / sum[4], sum[3] / sum[4], sum[5] / sum[4],
                sum[6] / sum[4], sum[7] / sum[4]);
    } else {
        cta::chol_cta_kernel<false, false><<<grid, cta::THREADS, smem, h->stream>>>(tmap, C, nt, s.rows, s.ld, s.elems(),
                                                                             h->pivot_tol, ws, info, queue, nullptr, 0, 1,
                                                                             nullptr, nullptr, 0);
    }
    h->launches++;
    return cudaGetLastError();
}


// Task list of one candidate for nt block columns (see cta::chol_task_kernel).  A panel task takes up to
// `blocks_per_task` row blocks from the top (2 = one 128-row pass, two DMMA row tiles per warp; more = several passes
// streamed through the pipeline without draining it: fewer, longer tasks for large populations, where nobody waits for
// anybody); a column with at most `single_below` row blocks left is cut into single blocks instead -- the late
// columns are a chain of short dependent tasks and want the parallelism more than the operand reuse.  Order, with look-ahead: D(0), P(0,first), D(1), P(0,rest), P(1,first),
// D(2), P(1,rest), ... -- the diagonal tile of column j+1 is claimed as soon as the pass that feeds it is.
void task_schedule(int nt, int single_below, int blocks_per_task, std::vector<cta::Task>& out) {
    out.clear();
    if (blocks_per_task < 1) blocks_per_task = 1;
    if (blocks_per_task > cta::TASK_MAX_BLOCKS) blocks_per_task = cta::TASK_MAX_BLOCKS;
    auto passes = [&](int j, std::vector<cta::Task>& ps) {
        ps.clear();
        const int left = nt - 1 - j;
        int b = j + 1;
        while (b < nt) {
            int nb = (left <= single_below) ? 1 : blocks_per_task;
            if (b + nb > nt) nb = nt - b;
            ps.push_back(cta::Task{1, j, b, nb});
            b += nb;
        }
    };
    std::vector<cta::Task> cur, nxt;
    out.push_back(cta::Task{0, 0, 0, 0});
    passes(0, cur);
    if (!cur.empty()) out.push_back(cur[0]);
    for (int j = 0; j < nt; ++j) {
        passes(j, cur);
        if (j + 1 < nt) out.push_back(cta::Task{0, j + 1, 0, 0});
        for (size_t i = 1; i < cur.size(); ++i) out.push_back(cur[i]);
        if (j + 1 < nt) {
            passes(j + 1, nxt);
            if (!nxt.empty()) out.push_back(nxt[0]);
        }
    }
}

// true when every task's inputs are produced by tasks earlier in the list (what makes ticket order deadlock free)
bool task_schedule_valid(int nt, const std::vector<cta::Task>& tl) {
    std::vector<int> ddone(1, 0), rdone(nt, 0);
    std::vector<char> seen_d(nt, 0);
    std::vector<std::vector<char>> seen_p(nt, std::vector<char>(nt, 0));  // [j][rb]
    for (const cta::Task& t : tl) {
        if (t.j < 0 || t.j >= nt) return false;
        if (t.type == 0) {
            if (ddone[0] != t.j) return false;                  // D tasks in order
            if (t.j > 0 && rdone[t.j] != t.j) return false;     // row block j complete up to column j - 1
            ddone[0] = t.j + 1;
            seen_d[t.j] = 1;
        } else {
            if (t.nb < 1 || t.nb > cta::TASK_MAX_BLOCKS || t.rb <= t.j || t.rb + t.nb > nt) return false;
            if (ddone[0] < t.j + 1) return false;
            for (int b = t.rb; b < t.rb + t.nb; ++b) {
                if (rdone[b] != t.j || seen_p[t.j][b]) return false;
                rdone[b] = t.j + 1;
                seen_p[t.j][b] = 1;
            }
        }
    }
    for (int j = 0; j < nt; ++j) {
        if (!seen_d[j]) return false;
        for (int b = j + 1; b < nt; ++b)
            if (!seen_p[j][b]) return false;
    }
    return true;
}

bool task_schedule_flat(int nt, int single_below, int blocks_per_task, std::vector<int>& out4) {
    std::vector<cta::Task> tl;
    task_schedule(nt, single_below, blocks_per_task, tl);
    out4.clear();
    for (const cta::Task& t : tl) {
        out4.push_back(t.type);
        out4.push_back(t.j);
        out4.push_back(t.rb);
        out4.push_back(t.nb);
    }
    return task_schedule_valid(nt, tl);
}

cudaError_t launch_cholesky_task(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info) {
    if (s.with_inverse) return cudaErrorInvalidValue;  // pk_fit's L^-T rows are passes of the gang kernel
    const int nt = s.np / PK_TILE;
    const size_t smem = sizeof(double) * (size_t)cta::SMEM_DOUBLES + cta::SMEM_ALIGN;
    CUtensorMap tmap;
    cudaError_t e = make_ws_tensor_map(&tmap, ws, s, C);
    if (e != cudaSuccess) return e;
    static bool attr_done_dev[PK_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
    if (!attr_done) {
        e = cudaFuncSetAttribute(cta::chol_task_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    // Granularity (never changes a bit of the result, so it may follow the size of THIS call): few candidates want many
    // short tasks -- every CTA slot busy, short dependency chains -- a population of several waves wants few long ones.
    int single_below = h->task_single_below, bpt = h->task_blocks;
    if (single_below < 0 || bpt <= 0) {
        const long long slots = 2LL * h->sm_count;
        int a_sb, a_bpt;
        if ((long long)C * 5 <= slots) { a_sb = 0; a_bpt = 1; }            // <= 59 candidates: 64-row tasks throughout
        else if ((long long)C * 4 <= slots * 3) { a_sb = 4; a_bpt = 2; }   // <= 222: 128-row tasks, 64-row ones in the late columns
        else {                                                              // a wave or more: long tasks, about 8 passes each
            a_sb = 0;
            a_bpt = 128 / nt;
            if (a_bpt < 2) a_bpt = 2;
            if (a_bpt > cta::TASK_MAX_BLOCKS) a_bpt = cta::TASK_MAX_BLOCKS;
        }
        if (single_below < 0) single_below = a_sb;
        if (bpt <= 0) bpt = a_bpt;
    }
    // per-shape task list, cached on the device
    if (h->task_nt != nt || h->task_single != single_below || h->task_bpt != bpt || !h->d_tasks) {
        std::vector<cta::Task> tl;
        task_schedule(nt, single_below, bpt, tl);
        if (!task_schedule_valid(nt, tl)) return cudaErrorInvalidValue;
        e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) return e;
        if (h->d_tasks) cudaFree(h->d_tasks);
        h->d_tasks = nullptr;
        e = cudaMalloc((void**)&h->d_tasks, sizeof(cta::Task) * tl.size());
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(h->d_tasks, tl.data(), sizeof(cta::Task) * tl.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
        h->task_nt = nt;
        h->task_single = single_below;
        h->task_bpt = bpt;
        h->task_count = (int)tl.size();
    }
    // completion flags (zeroed per launch) + the inverted diagonal blocks
    const int fstride = ((nt + 1 + 31) / 32) * 32;
    const size_t flag_bytes = ((sizeof(int) * (size_t)C * fstride) + 255) & ~(size_t)255;
    const size_t dinv_bytes = sizeof(double) * (size_t)C * nt * cta::DINV_DOUBLES;
    if (flag_bytes + dinv_bytes > h->task_bytes) {
        e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) return e;
        if (h->d_taskws) cudaFree(h->d_taskws);
        h->d_taskws = nullptr;
        h->task_bytes = 0;
        e = cudaMalloc((void**)&h->d_taskws, flag_bytes + dinv_bytes);
        if (e != cudaSuccess) return e;
        h->task_bytes = flag_bytes + dinv_bytes;
    }
    int* flags = reinterpret_cast<int*>(h->d_taskws);
    double* dinv = reinterpret_cast<double*>(reinterpret_cast<char*>(h->d_taskws) + flag_bytes);
    e = cudaMemsetAsync(flags, 0, flag_bytes, h->stream);
    if (e != cudaSuccess) return e;
    unsigned int* queue = h->d_queue + h->queue_slot;
    e = cudaMemsetAsync(queue, 0, sizeof(unsigned int), h->stream);
    if (e != cudaSuccess) return e;
    int skew = h->task_skew;
    if (skew <= 0) skew = 1;  // measured without effect (profiles/task_sweep_r02_d.jsonl): the CTAs desynchronise on their own
    const long long total = (long long)C * (h->task_count + skew - 1);
    if (total >= (1LL << 32)) return cudaErrorInvalidValue;  // 32-bit ticket counter
    int S = h->task_cohort > 0 ? h->task_cohort : C;
    if (S > C) S = C;
    const long long slots = 2LL * h->sm_count;
    const int grid = (int)(total < slots ? total : slots);
    cta::chol_task_kernel<<<grid, cta::THREADS, smem, h->stream>>>(tmap, C, nt, s.ld, s.elems(), h->pivot_tol, ws, info, queue,
                                                               reinterpret_cast<const cta::Task*>(h->d_tasks), h->task_count, S, skew,
                                                               flags, fstride, dinv, h->d_fault);
    h->launches++;
    return cudaGetLastError();
}

}  // namespace pk
