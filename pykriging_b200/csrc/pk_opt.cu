// Device-resident GA / PSO population (the candidate-population loop of kriging.train, krige.py:373-399, driven in the
// reference by inspyred.ec.GA / inspyred.swarm.PSO through fittingObjective, krige.py:429-452).
//
// The population never leaves the GPU between generations: the host keeps what seed parity pins -- the draws of
// Python's Mersenne Twister and the comparisons of fitness values (sort orders, ring-topology bests), all on the
// per-candidate SCALARS the likelihood kernels return -- and sends index / uniform arrays; the kernels here move and
// recombine the candidate vectors themselves:
//     ga:  generational replacement (gather of the survivors from [offspring | old population]), then rank-selected
//          parents recombined by one-point crossover                         (inspyred 1.0.x semantics)
//     pso: personal-best archive update, then x' = x + w (x - x_prev) + c1 r1 (pbest - x) + c2 r2 (nbest - x), clipped to
//          the bounds -- evaluated in the expression order of the host implementation (no FMA contraction), so a
//          device-resident swarm follows the host swarm bit for bit.
#include "pk_internal.h"

namespace pk {

// dst[i, :] = (keep[i] < nc) ? child[keep[i], :] : pop[keep[i] - nc, :]
__global__ void opt_gather_kernel(double* __restrict__ dst, const double* __restrict__ child, const double* __restrict__ pop,
                                  const int32_t* __restrict__ keep, int nc, int n, int dims) {
    const int i = blockIdx.x * blockDim.y + threadIdx.y;
    if (i >= n) return;
    const int src = keep[i];
    const double* row = (src < nc) ? child + (size_t)src * dims : pop + (size_t)(src - nc) * dims;
    for (int d = threadIdx.x; d < dims; d += blockDim.x) dst[(size_t)i * dims + d] = row[d];
}

// pair i: mom = pop[parents[2i]], dad = pop[parents[2i+1]]; cut[i] == 0: children (mom, dad) unchanged, otherwise
// bro = dad[:cut] + mom[cut:], sis = mom[:cut] + dad[cut:]  (n_point_crossover with one cut point)
__global__ void ga_cross_kernel(double* __restrict__ child, const double* __restrict__ pop, const int32_t* __restrict__ parents,
                                const int32_t* __restrict__ cut, int npairs, int dims) {
    const int i = blockIdx.x * blockDim.y + threadIdx.y;
    if (i >= npairs) return;
    const double* mom = pop + (size_t)parents[2 * i] * dims;
    const double* dad = pop + (size_t)parents[2 * i + 1] * dims;
    const int c = cut[i];
    for (int d = threadIdx.x; d < dims; d += blockDim.x) {
        const double m = mom[d], f = dad[d];
        const bool swapped = (c > 0) && (d >= c);
        child[(size_t)(2 * i) * dims + d] = (c == 0) ? m : (swapped ? m : f);
        child[(size_t)(2 * i + 1) * dims + d] = (c == 0) ? f : (swapped ? f : m);
    }
}

// personal bests: arch[i, :] = x[i, :] where the new position is not strictly worse (mask computed on the host from the
// fitness scalars, inspyred's swarm archiver)
__global__ void pso_archive_kernel(double* __restrict__ arch, const double* __restrict__ x, const int32_t* __restrict__ improved,
                                   int n, int dims) {
    const int i = blockIdx.x * blockDim.y + threadIdx.y;
    if (i >= n || !improved[i]) return;
    for (int d = threadIdx.x; d < dims; d += blockDim.x) arch[(size_t)i * dims + d] = x[(size_t)i * dims + d];
}

// One element per thread.  value = ((x + w (x - prev)) + (c1 r1) (arch - x)) + (c2 r2) (nbest - x): the grouping NumPy /
// Python give `x + inertia * (x - prev_x) + cognitive * r1 * (arch_x - x) + social * r2 * (nb - x)`, every product and
// sum rounded on its own.
__global__ void pso_move_kernel(double* __restrict__ x, double* __restrict__ prev, const double* __restrict__ arch,
                                const double* __restrict__ r, const int32_t* __restrict__ nbest, int first, double inertia,
                                double cognitive, double social, const double* __restrict__ lo, const double* __restrict__ hi,
                                int n, int dims) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * dims) return;
    const int i = e / dims, d = e % dims;
    const double xv = x[e];
    const double pv = first ? xv : prev[e];
    const double av = arch[e];
    const double nv = arch[(size_t)nbest[i] * dims + d];
    const double r1 = r[2 * (size_t)e], r2 = r[2 * (size_t)e + 1];
    double v = __dadd_rn(xv, __dmul_rn(inertia, __dsub_rn(xv, pv)));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(cognitive, r1), __dsub_rn(av, xv)));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(social, r2), __dsub_rn(nv, xv)));
    v = fmax(fmin(v, hi[d]), lo[d]);  // Bounder: np.maximum(np.minimum(v, hi), lo)
    prev[e] = xv;
    x[e] = v;
}

// candidate rows [theta_1..k, p_1..k(, lambda)] -> the contiguous theta / p / lambda arrays pk_nll_batch takes
__global__ void opt_split_kernel(const double* __restrict__ cand, int n, int k, int dims, double* __restrict__ theta,
                                 double* __restrict__ p, double* __restrict__ lambda) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * dims) return;
    const int i = e / dims, d = e % dims;
    const double v = cand[e];
    if (d < k) theta[(size_t)i * k + d] = v;
    else if (d < 2 * k) p[(size_t)i * k + (d - k)] = v;
    else if (lambda) lambda[i] = v;
}

static dim3 row_block(int dims) { return dim3(dims <= 32 ? 32 : 64, dims <= 32 ? 8 : 4); }

cudaError_t launch_opt_gather(pk_handle* h, double* dst, const double* child, const double* pop, const int32_t* keep, int nc,
                              int n, int dims) {
    const dim3 b = row_block(dims);
    opt_gather_kernel<<<(n + b.y - 1) / b.y, b, 0, h->stream>>>(dst, child, pop, keep, nc, n, dims);
    h->launches++;
    return cudaGetLastError();
}
cudaError_t launch_ga_cross(pk_handle* h, double* child, const double* pop, const int32_t* parents, const int32_t* cut,
                            int npairs, int dims) {
    const dim3 b = row_block(dims);
    ga_cross_kernel<<<(npairs + b.y - 1) / b.y, b, 0, h->stream>>>(child, pop, parents, cut, npairs, dims);
    h->launches++;
    return cudaGetLastError();
}
cudaError_t launch_pso_step(pk_handle* h, double* x, double* prev, double* arch, const double* r, const int32_t* nbest,
                            const int32_t* improved, int first, double inertia, double cognitive, double social,
                            const double* lo, const double* hi, int n, int dims) {
    if (improved) {
        const dim3 b = row_block(dims);
        pso_archive_kernel<<<(n + b.y - 1) / b.y, b, 0, h->stream>>>(arch, x, improved, n, dims);
        h->launches++;
    }
    pso_move_kernel<<<(n * dims + 255) / 256, 256, 0, h->stream>>>(x, prev, arch, r, nbest, first, inertia, cognitive, social, lo,
                                                                  hi, n, dims);
    h->launches++;
    return cudaGetLastError();
}
cudaError_t launch_opt_split(pk_handle* h, const double* cand, int n, int k, int dims, double* theta, double* p, double* lambda) {
    opt_split_kernel<<<(n * dims + 255) / 256, 256, 0, h->stream>>>(cand, n, k, dims, theta, p, lambda);
    h->launches++;
    return cudaGetLastError();
}

}  // namespace pk
