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
#include "pk_internal.h"
#include <mutex>

#include "pk_math.cuh"

namespace pk {

constexpr int PSI_THREADS = 256;
constexpr int PSI_GROUP = 128;  // candidates per CTA (amortises the log2 split of the CTA's pairs and the table copy)

// PPT pairs per thread: large K keeps ONE pair per thread so that the per-pair log2 tables (3 registers per
// dimension) leave room for three resident CTAs per SM -- the per-candidate 2^x chain is latency bound and
// needs the extra warps more than it needs wider stores.
// (theta_d, p_d) of one candidate group in constant memory, two buffers (pk_handle::psi_const_min_k = 5 <= k <= PSI_CONST_MAX_K; k <= 4: 5 - 10 % slower this way, profiles/psi_const_r02.jsonl): the candidate loop reads them
// through the uniform datapath (LDCU: one register copy per warp) instead of a 16-byte shared-memory broadcast per lane,
// which costs four cycles of the shared-memory pipe per (candidate, dimension) -- 37 of the kernel's 85 wavefronts per warp
// and entry (profiles/ncu_psi_pop_r02_u.txt: that pipe is 92 % busy)
constexpr int PSI_CONST_MAX_K = 10;
__constant__ double2 c_thp[2][PSI_GROUP * PSI_CONST_MAX_K];

// bytes of the per-CTA arrays in front of the 2^x table
template <int K, int PPT>
__host__ __device__ constexpr unsigned psi_small_bytes() {
    return (unsigned)(sizeof(double) * (2 * PSI_GROUP * K + (PSI_THREADS * PPT / PK_TILE + PK_TILE) * K + PSI_GROUP) +
                      sizeof(int) * PSI_GROUP);
}

// CB >= 0: the launch covers ONE candidate group (blockIdx.y = 0, candidates g0 ...), whose (theta, p) lie in c_thp[CB]
template <int K, int PPT, int CB = -1>
__global__ void __launch_bounds__(PSI_THREADS, (PPT == 1) ? (K <= 12 ? 3 : 2) : 1)
psi_pop_kernel(int n, int np, int ld, size_t elems, int C, const double* __restrict__ X,
               const double* __restrict__ theta, const double* __restrict__ p,
               const double* __restrict__ lambda, int mode, double* __restrict__ ws, int group0 = 0) {
    constexpr int RB = PSI_THREADS * PPT / PK_TILE;  // rows of the pair block
    constexpr int SPT = PK_TILE / RB;                // row blocks per tile
    extern __shared__ __align__(16) double psi_smem[];
    double2* thp = reinterpret_cast<double2*>(psi_smem);             // (theta_d, p_d): one 16-byte broadcast load per (candidate, dimension)
    double* Xr = reinterpret_cast<double*>(thp + PSI_GROUP * K);     // RB x K
    double* Xc = Xr + RB * K;                                        // 64 x K
    double* dg = Xc + PK_TILE * K;                                   // PSI_GROUP diagonal values
    int* spec = reinterpret_cast<int*>(dg + PSI_GROUP);              // candidate takes the exact / range-checked path
    const int tid = threadIdx.x;
    // 2^(j/4096) (32 KB) at the first 32 KB-aligned shared address behind the arrays above (the launcher sizes the
    // dynamic allocation for it): exp2_lean4k ORs the entry's byte offset into the address
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(psi_smem);
    unsigned tab_saddr = (smem_base + psi_small_bytes<K, PPT>() + 0x7fffu) & ~0x7fffu;
    asm volatile("" : "+r"(tab_saddr));  // opaque: keeps the address in a register instead of re-deriving it (5 instructions) per candidate
    double* tab4k = psi_smem + (tab_saddr - smem_base) / sizeof(double);
    // block -> (lower tile, row block)
    const int t = blockIdx.x / SPT, sb = blockIdx.x % SPT;
    int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    const int tj = t - ti * (ti + 1) / 2;
    const int r0 = ti * PK_TILE + sb * RB, c0 = tj * PK_TILE;
    const int g0 = (group0 + blockIdx.y) * PSI_GROUP;
    const int gn = min(PSI_GROUP, C - g0);

    fill_exp2_4096(tab4k, tid, PSI_THREADS);
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
            sp |= (pd == 1.0 || pd == 2.0 || !(pd >= 0.25 && pd <= 3.0)) ? 1 : 0;
        }
        spec[tid] = sp;
    }
    __syncthreads();

    const int lr = tid / (PK_TILE / PPT);            // local row
    const int lc = (tid % (PK_TILE / PPT)) * PPT;    // first local column
    const int gr = r0 + lr;
    double lh[PPT][K];
    double ll[PPT][K];  // kept as doubles: a float would be re-converted (F2F.F64.F32, FP64-rate) per candidate
    bool live[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        const int gc = c0 + lc + q;
        live[q] = (gr < n) && (gc < gr);
#pragma unroll
        for (int d = 0; d < K; ++d) {
            if (live[q]) {
                float lf;
                log2_split_clamped(fabs(Xr[lr * K + d] - Xc[(lc + q) * K + d]), lh[q][d], lf);
                ll[q][d] = (double)lf;
            } else {
                lh[q][d] = 0.0;
                ll[q][d] = 0.0;
            }
        }
    }
    bool any_live = false;
#pragma unroll
    for (int q = 0; q < PPT; ++q) any_live |= live[q];

    // MAGIC4K = 1.5 * 2^40 in a REGISTER (n >> 30 is zero, which neither the compiler nor ptxas can know): as an immediate it
    // takes the DFMA's one special-operand slot, and the uniform p_d would be copied into vector registers for every 2^x
    const double magic4k = (CB >= 0) ? __hiloint2double(0x42780000 | (n >> 30), 0) : MAGIC4K;
    auto one_candidate = [&](int g) {
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
                    const double2 tp = (CB >= 0 && K <= PSI_CONST_MAX_K) ? c_thp[CB < 0 ? 0 : CB][g * K + d] : tpg[d];
                    th_d = tp.x;
                    return exp2_lean4k(tp.y, lh[q][d], ll[q][d], tab_saddr, magic4k);
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
                v = expneg_4k(S, tab_saddr);
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
                        const double dd = fabs(Xr[lr * K + d] - Xc[(lc + q) * K + d]);
                        double l0;
                        float l1;
                        log2_split(dd, l0, l1);  // unclamped: exp2_const checks the range itself
                        const double s = pd * l0;
                        const double e = fma(pd, (double)l1, fma(pd, l0, -s));
                        w = exp2_const(s, e);
                    }
                    term[d] = tpg[d].x * w;
                }
                const double S = sum_numpy_order(K, [&](int d) { return term[d]; });
                v = expneg_4k(S, tab_saddr);
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
    };

    for (int g = 0; g < gn; ++g) one_candidate(g);
    (void)any_live;
}

// diagnostic: the device pow / exp building blocks on arrays (tests measure their ulp error)
__global__ void debug_math_kernel(int64_t m, const double* __restrict__ d, const double* __restrict__ p,
                                  double* __restrict__ out_pow, double* __restrict__ out_expneg) {
    __shared__ double tab[64 * 16];
    extern __shared__ __align__(16) double dbg_dyn[];
    const unsigned dbg_base = (unsigned)__cvta_generic_to_shared(dbg_dyn);
    const unsigned dbg_tab = (dbg_base + 0x7fffu) & ~0x7fffu;
    double* dbg_tab4k = dbg_dyn + (dbg_tab - dbg_base) / sizeof(double);
    for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) tab[i] = c_exp2_hi[i >> 4];
    fill_exp2_4096(dbg_tab4k, threadIdx.x, blockDim.x);
    __syncthreads();
    const int bank = threadIdx.x & 15;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        double lh;
        float ll;
        log2_split(fabs(d[i]), lh, ll);
        const double s = p[i] * lh;
        const double e = fma(p[i], (double)ll, fma(p[i], lh, -s));
        // element i goes through variant i % 4: range-checked 64-entry table | lean 64-entry (predictor hot
        // loop) | lean 4096-entry (population Psi hot loop, clamped log2) | range-checked constant-bank path
        double r;
        switch ((int)(i & 3)) {
            case 0: r = exp2_table(s, e, tab, bank); break;
            case 1: r = exp2_lean(p[i], lh, (double)ll, reinterpret_cast<const char*>(tab + bank)); break;
            case 2: {
                double lc;
                float lr;
                log2_split_clamped(fabs(d[i]), lc, lr);
                r = exp2_lean4k(p[i], lc, (double)lr, dbg_tab);
                break;
            }
            default: r = exp2_const(s, e); break;
        }
        out_pow[i] = r;
        out_expneg[i] = (i & 1) ? expneg_4k(d[i], dbg_tab) : expneg_table(d[i], tab, bank);
    }
}

cudaError_t launch_debug_math(pk_handle* h, int64_t m, const double* d, const double* p, double* out_pow,
                              double* out_expneg) {
    static bool attr_done_dev[PK_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(debug_math_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 0x10000);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    debug_math_kernel<<<296, 256, 0x10000, h->stream>>>(m, d, p, out_pow, out_expneg);
    h->launches++;
    return cudaGetLastError();
}

// generic fallback (any k <= PK_MAX_K): defined in pk_tiled.cu
cudaError_t launch_psi_build_generic(pk_handle* h, const TiledShape& s, int C, const double* theta, const double* p,
                                     const double* lambda, int mode, double* ws);
cudaError_t launch_aug_init(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info);

// interleaves theta and p of a population into (theta_d, p_d) pairs: the source of the constant-memory copies
__global__ void pack_thp_kernel(size_t m, const double* __restrict__ theta, const double* __restrict__ p, double2* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_double2(theta[i], p[i]);
}

template <int K, int PPT>
static cudaError_t launch_psi_pop_ppt(pk_handle* h, const TiledShape& s, int C, const double* theta, const double* p,
                                      const double* lambda, int mode, double* ws) {
    constexpr int RB = PSI_THREADS * PPT / PK_TILE;
    const int nt = s.np / PK_TILE;
    const int groups = (C + PSI_GROUP - 1) / PSI_GROUP;
    dim3 grid(nt * (nt + 1) / 2 * (PK_TILE / RB), groups);
    // dynamic shared memory starts 1 KB into the CTA's window (the first KB is reserved by the system): room for
    // the table at the next 32 KB boundary wherever the base actually lies
    constexpr size_t smem = ((psi_small_bytes<K, PPT>() + 0x400 + 0x7fff) & ~(size_t)0x7fff) + 0x8000;
    static bool attr_done_dev[PK_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[h->device & (PK_MAX_DEVICES - 1)];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(psi_pop_kernel<K, PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if constexpr (PPT == 1 && K <= PSI_CONST_MAX_K) {
            if (e == cudaSuccess) e = cudaFuncSetAttribute(psi_pop_kernel<K, PPT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(psi_pop_kernel<K, PPT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    if constexpr (PPT == 1 && K <= PSI_CONST_MAX_K) if (h->psi_const && K >= h->psi_const_min_k) {
        // One launch per candidate group, (theta, p) of the group in constant memory.  The two buffers belong to the DEVICE
        // (every handle of the process shares them): g_thp_used[b] is recorded behind the last kernel that read buffer b and
        // waited for by the next copy into it, under a lock that keeps wait / copy / launch / record of one group together.
        // Several groups: they alternate between two side streams forked from the caller's stream, so that the next group's
        // CTAs fill the tail of the previous launch (one stream: 5.66 instead of 5.26 ms per 1024 candidates at n = 1024, k = 10;
        // shared-memory loads: 5.71 ms).
        const size_t m = (size_t)C * K;
        if (h->thp_cap < m) {
            if (h->d_thp) cudaFree(h->d_thp);
            h->d_thp = nullptr;
            h->thp_cap = 0;
            cudaError_t e = cudaMalloc((void**)&h->d_thp, sizeof(double2) * m);
            if (e != cudaSuccess) return e;
            h->thp_cap = m;
        }
        static std::mutex g_thp_lock;
        static cudaEvent_t g_thp_used[PK_MAX_DEVICES][2] = {};
        cudaEvent_t* used = g_thp_used[h->device & (PK_MAX_DEVICES - 1)];
        cudaError_t e = cudaSuccess;
        const bool fan = groups > 1;
        if (fan && !h->psi_stream[0]) {
            for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&h->psi_stream[i], cudaStreamNonBlocking);
            for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->psi_event[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        double2* thp = reinterpret_cast<double2*>(h->d_thp);
        const size_t pb = (m + 255) / 256;
        pack_thp_kernel<<<(unsigned)(pb < 296 ? pb : 296), 256, 0, h->stream>>>(m, theta, p, thp);
        h->launches++;
        if (fan) {
            e = cudaEventRecord(h->psi_event[2], h->stream);
            for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(h->psi_stream[i], h->psi_event[2], 0);
            if (e != cudaSuccess) return e;
        }
        grid.y = 1;
        {
            std::lock_guard<std::mutex> lock(g_thp_lock);
            for (int b = 0; b < 2 && e == cudaSuccess; ++b)
                if (!used[b]) e = cudaEventCreateWithFlags(&used[b], cudaEventDisableTiming);
            for (int g = 0; g < groups && e == cudaSuccess; ++g) {
                const int gn = (C - g * PSI_GROUP < PSI_GROUP) ? C - g * PSI_GROUP : PSI_GROUP;
                const int b = g & 1;
                cudaStream_t st = fan ? h->psi_stream[b] : h->stream;
                e = cudaStreamWaitEvent(st, used[b], 0);  // never recorded: no wait
                if (e == cudaSuccess)
                    e = cudaMemcpyToSymbolAsync(c_thp, thp + (size_t)g * PSI_GROUP * K, sizeof(double2) * (size_t)gn * K,
                                                sizeof(double2) * (size_t)b * PSI_GROUP * PSI_CONST_MAX_K, cudaMemcpyDeviceToDevice, st);
                if (e != cudaSuccess) break;
                if (b == 0)
                    psi_pop_kernel<K, PPT, 0><<<grid, PSI_THREADS, smem, st>>>(s.n, s.np, s.ld, s.elems(), C, h->dX, theta, p, lambda, mode, ws, g);
                else
                    psi_pop_kernel<K, PPT, 1><<<grid, PSI_THREADS, smem, st>>>(s.n, s.np, s.ld, s.elems(), C, h->dX, theta, p, lambda, mode, ws, g);
                h->launches++;
                e = cudaEventRecord(used[b], st);
            }
        }
        if (fan) {
            for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
                e = cudaEventRecord(h->psi_event[i], h->psi_stream[i]);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(h->stream, h->psi_event[i], 0);
            }
        }
        return (e != cudaSuccess) ? e : cudaGetLastError();
    }
    psi_pop_kernel<K, PPT><<<grid, PSI_THREADS, smem, h->stream>>>(s.n, s.np, s.ld, s.elems(), C, h->dX, theta, p,
                                                                    lambda, mode, ws);
    h->launches++;
    return cudaGetLastError();
}

// sample pairs per thread: one; up to k = 6 round 1 took four (175 registers, one CTA per SM), which measures 7 - 38 % slower
// (profiles/psi_const_r02.jsonl) and is kept behind PK_PSI_PPT_SMALL=4 (pk_handle::psi_ppt_small); same bits
template <int K>
static cudaError_t launch_psi_pop(pk_handle* h, const TiledShape& s, int C, const double* theta, const double* p,
                                  const double* lambda, int mode, double* ws) {
    if constexpr (K <= 6) {
        if (h->psi_ppt_small == 4) return launch_psi_pop_ppt<K, 4>(h, s, C, theta, p, lambda, mode, ws);
    }
    return launch_psi_pop_ppt<K, 1>(h, s, C, theta, p, lambda, mode, ws);
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
