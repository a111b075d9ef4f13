// C ABI of libpykriging_b200.so (see include/pykriging_b200.h for the contract of every entry point).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "pk_internal.h"

namespace pk {
cudaError_t launch_debug_math(pk_handle* h, int64_t m, const double* d, const double* p, double* out_pow,
                              double* out_expneg);
}

static std::string g_create_error;

#define PK_FAIL(h, ...)                                  \
    do {                                                 \
        char buf_[512];                                  \
        snprintf(buf_, sizeof(buf_), __VA_ARGS__);       \
        (h)->err = buf_;                                 \
        return 1;                                        \
    } while (0)

#define PK_CUDA(h, call)                                                                             \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) PK_FAIL(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

namespace pk {
void phase_begin(pk_handle* h, int phase) {
    if (!h->profiling) return;
    if (h->ev_used + 2 > h->ev_pool.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        h->ev_pool.push_back(a);
        h->ev_pool.push_back(b);
        h->ev_phase.push_back(phase);
    }
    h->ev_phase[h->ev_used / 2] = phase;
    cudaEventRecord(h->ev_pool[h->ev_used], h->stream);
}
void phase_end(pk_handle* h) {
    if (!h->profiling) return;
    cudaEventRecord(h->ev_pool[h->ev_used + 1], h->stream);
    h->ev_used += 2;
}
}  // namespace pk

static bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

template <typename T>
static cudaError_t regrow(T** ptr, size_t count) {
    if (*ptr) {
        cudaError_t e = cudaFree(*ptr);
        *ptr = nullptr;
        if (e != cudaSuccess) return e;
    }
    return cudaMalloc((void**)ptr, count * sizeof(T));
}

static int ensure_candidates(pk_handle* h, int C) {
    if (C <= h->cap_C && h->k <= h->cap_k) return 0;
    const int cap = C > h->cap_C ? C : h->cap_C;
    const int ck = h->k > h->cap_k ? h->k : h->cap_k;
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    PK_CUDA(h, regrow(&h->d_theta, (size_t)cap * ck));
    PK_CUDA(h, regrow(&h->d_p, (size_t)cap * ck));
    PK_CUDA(h, regrow(&h->d_lambda, (size_t)cap));
    PK_CUDA(h, regrow(&h->d_out, (size_t)cap * 4));
    PK_CUDA(h, regrow(&h->d_info, (size_t)cap));
    h->cap_C = cap;
    h->cap_k = ck;
    return 0;
}

// Return a device pointer holding `count` doubles of `src` (copying through `stage` when src is a host pointer).
static int to_device(pk_handle* h, const double* src, double* stage, size_t count, const double** out) {
    if (!src) {
        *out = nullptr;
        return 0;
    }
    if (is_device_ptr(src)) {
        *out = src;
        return 0;
    }
    PK_CUDA(h, cudaMemcpyAsync(stage, src, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    *out = stage;
    return 0;
}

// info < 0 is not a numerical outcome: a spin-wait of the persistent Cholesky (gang barrier / task dependency) timed out
// -- its CTAs were starved by another kernel holding the SMs for seconds.  The likelihood epilogue raises a sticky
// device flag for such candidates; it is read back wherever the stream has just been synchronised (host outputs,
// pk_synchronize), so the failure is loud wherever the outputs live.
static int check_fault(pk_handle* h, const char* where) {
    int f = 0;
    PK_CUDA(h, cudaMemcpy(&f, h->d_fault, sizeof(int), cudaMemcpyDeviceToHost));
    if (f != 0) {
        PK_CUDA(h, cudaMemset(h->d_fault, 0, sizeof(int)));
        PK_FAIL(h, "%s: a spin-wait of the persistent Cholesky timed out for at least one candidate (GPU shared with a "
                   "long-running kernel?); its results are invalid", where);
    }
    return 0;
}

extern "C" {

int pk_version(void) { return PK_VERSION; }

const char* pk_last_error(const pk_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int pk_create(pk_handle** out, int device) {
    if (!out) return 1;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no usable CUDA device: ") + cudaGetErrorString(e) +
                         " (pykriging_b200 has no CPU fallback)";
        cudaGetLastError();
        return 1;
    }
    if (device < 0 || device >= count) {
        g_create_error = "device index out of range";
        return 1;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return 1;
    }
    pk_handle* h = new (std::nothrow) pk_handle();
    if (!h) return 1;
    h->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
        delete h;
        return 1;
    }
    e = cudaMalloc((void**)&h->d_hp, sizeof(double) * 2 * PK_MAX_K);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_fit_hp, sizeof(double) * 2 * PK_MAX_K);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_try_hp, sizeof(double) * 2 * PK_MAX_K);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_queue, sizeof(unsigned int) * PK_QUEUE_SLOTS);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_fault, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->d_fault, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
    if (const char* v = getenv("PK_SPLIT")) h->env_split = v;
    if (const char* v = getenv("PK_CHOL_GANG")) h->env_gang = atoi(v) > 0 ? atoi(v) : 0;
    if (const char* v = getenv("PK_CHOL_CTAS_PER_SM")) h->env_ctas_per_sm = (atoi(v) == 1) ? 1 : 2;
    h->env_cta_timing = getenv("PK_CTA_TIMING") != nullptr;
    h->env_cta_alias = getenv("PK_CTA_ALIAS") != nullptr;
    if (const char* v = getenv("PK_RAGGED_FIRST")) h->ragged_first = atoi(v) != 0;
    if (const char* v = getenv("PK_I8_L2")) h->i8_l2_mode = atoi(v) & 3;
    if (const char* v = getenv("PK_STRAGGLERS")) h->stragglers = atoi(v) != 0;
    if (const char* v = getenv("PK_I8_PERM")) h->i8_perm = atoi(v) != 0;
    if (const char* v = getenv("PK_I8_QUIET")) h->i8_quiet = atoi(v) != 0;
    if (const char* v = getenv("PK_PSI_CONST")) h->psi_const = atoi(v) != 0;
    if (const char* v = getenv("PK_PSI_PPT_SMALL")) h->psi_ppt_small = (atoi(v) == 4) ? 4 : 1;
    if (const char* v = getenv("PK_PSI_CONST_MINK")) h->psi_const_min_k = atoi(v);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        cudaStreamDestroy(h->stream);
        delete h;
        return 1;
    }
    *out = h;
    return 0;
}

int pk_destroy(pk_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    void* ptrs[] = {h->dX, h->dy, h->d_theta, h->d_p, h->d_lambda, h->d_out, h->d_info, h->d_ws, h->d_fit, h->d_try,
                    h->d_psi, h->d_linv, h->d_ad, h->d_w, h->d_q, h->d_qout, h->d_hp, h->d_queue, h->d_fit_hp, h->d_try_hp, h->d_pscratch, h->d_gang, h->d_gridinfo, h->d_tasks, h->d_taskws, h->d_fault, h->d_opt[0], h->d_opt[1], h->d_opt[2],
                    h->d_opt[3], h->d_opt[4], h->d_opt_idx, h->d_opt_r, h->d_linv_i8, h->d_linv_scale, h->d_psi_i8, h->d_thp};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_pipe) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_pred) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i)
        if (h->psi_stream[i]) {
            cudaStreamSynchronize(h->psi_stream[i]);
            cudaStreamDestroy(h->psi_stream[i]);
        }
    for (int i = 0; i < 3; ++i)
        if (h->psi_event[i]) cudaEventDestroy(h->psi_event[i]);
    if (h->stream2) {
        cudaStreamSynchronize(h->stream2);
        cudaStreamDestroy(h->stream2);
    }
    cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int pk_synchronize(pk_handle* h) {
    if (!h) return 1;
    PK_CUDA(h, cudaSetDevice(h->device));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    return check_fault(h, "pk_synchronize");
}

void* pk_stream(pk_handle* h) { return h ? (void*)h->stream : nullptr; }

int64_t pk_launch_count(const pk_handle* h) { return h ? h->launches : 0; }

int pk_set_pivot_tolerance(pk_handle* h, double tol) {
    if (!h || !(tol >= 0.0)) return 1;
    h->pivot_tol = tol;
    return 0;
}

int pk_set_cholesky_variant(pk_handle* h, int variant) {
    if (!h || variant < 0 || variant > 5) return 1;
    h->chol_variant = variant;
    return 0;
}

int pk_set_task_options(pk_handle* h, int single_below, int blocks_per_task, int cohort, int skew) {
    if (!h || single_below < -1 || blocks_per_task < 0 || cohort < 0 || skew < 0) return 1;
    h->task_single_below = single_below;
    h->task_blocks = blocks_per_task;
    h->task_cohort = cohort;
    h->task_skew = skew;
    return 0;
}

int pk_task_schedule(int nt, int single_below, int blocks_per_task, int32_t* out4, int cap) {
    if (nt < 1 || single_below < 0 || blocks_per_task < 1) return -1;
    std::vector<int> flat;
    if (!pk::task_schedule_flat(nt, single_below, blocks_per_task, flat)) return -1;
    const int count = (int)(flat.size() / 4);
    if (out4)
        for (int i = 0; i < count && i < cap; ++i)
            for (int c = 0; c < 4; ++c) out4[4 * i + c] = flat[4 * i + c];
    return count;
}

int pk_set_shard_hint(pk_handle* h, int64_t population, int64_t query_points) {
    if (!h || population < 0 || query_points < 0) return 1;
    h->hint_C = population;
    h->hint_m = query_points;
    h->hint_first = 0;
    return 0;
}

int pk_set_shard_offset(pk_handle* h, int64_t first_candidate) {
    if (!h || first_candidate < 0) return 1;
    h->hint_first = first_candidate;
    return 0;
}

int pk_set_overlap(pk_handle* h, int enabled) {
    if (!h) return 1;
    h->overlap = enabled ? 1 : 0;
    return 0;
}

int pk_set_predict_variant(pk_handle* h, int variant) {
    if (!h || variant < 0 || variant > 4) return 1;
    h->predict_grid_mode = (variant == 4) ? 0 : 1;  // 4 = split pipeline with the grid detection switched off
    h->predict_variant = (variant == 4) ? 3 : variant;
    return 0;
}

int pk_set_predict_contraction(pk_handle* h, int mode) {
    if (!h || mode < 0 || mode > 2) return 1;
    h->predict_contraction = mode;
    return 0;
}

int pk_debug_math(pk_handle* h, int64_t m, const double* d, const double* p, double* out_pow, double* out_expneg) {
    if (!h || m < 0 || !d || !p || !out_pow || !out_expneg) return 1;
    if (m == 0) return 0;
    PK_CUDA(h, cudaSetDevice(h->device));
    double* buf = nullptr;
    PK_CUDA(h, cudaMalloc((void**)&buf, sizeof(double) * 4 * (size_t)m));
    PK_CUDA(h, cudaMemcpyAsync(buf, d, sizeof(double) * m, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, cudaMemcpyAsync(buf + m, p, sizeof(double) * m, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, pk::launch_debug_math(h, m, buf, buf + m, buf + 2 * m, buf + 3 * m));
    PK_CUDA(h, cudaMemcpyAsync(out_pow, buf + 2 * m, sizeof(double) * m, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, cudaMemcpyAsync(out_expneg, buf + 3 * m, sizeof(double) * m, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    PK_CUDA(h, cudaFree(buf));
    return 0;
}

int pk_set_profiling(pk_handle* h, int enabled) {
    if (!h) return 1;
    h->profiling = enabled != 0;
    return 0;
}

int pk_phase_times(pk_handle* h, double* ms4) {
    if (!h || !ms4) return 1;
    PK_CUDA(h, cudaSetDevice(h->device));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i + 1 < h->ev_used + 1 && i < h->ev_used; i += 2) {
        float ms = 0.f;
        PK_CUDA(h, cudaEventElapsedTime(&ms, h->ev_pool[i], h->ev_pool[i + 1]));
        const int ph = h->ev_phase[i / 2];
        if (ph >= 0 && ph < 4) ms4[ph] += (double)ms;
    }
    h->ev_used = 0;
    return 0;
}

int pk_fp64_peak(pk_handle* h, double* dmma_tflops, double* dfma_tflops) {
    if (!h) return 1;
    PK_CUDA(h, cudaSetDevice(h->device));
    PK_CUDA(h, pk::measure_fp64_peak(h, dmma_tflops, dfma_tflops));
    return 0;
}

int pk_set_workspace_limit(pk_handle* h, int64_t bytes) {
    if (!h || bytes < 0) return 1;
    h->ws_limit = (size_t)bytes;
    return 0;
}

int pk_set_data(pk_handle* h, int n, int k, const double* X, const double* y) {
    if (!h) return 1;
    if (n < 2 || k < 1 || k > PK_MAX_K || !X || !y) PK_FAIL(h, "pk_set_data: need n>=2, 1<=k<=%d, X, y", PK_MAX_K);
    PK_CUDA(h, cudaSetDevice(h->device));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    // capacity up to the next tile boundary, so that pk_append_point can add samples in place
    const int cap_rows = (n / PK_TILE + 1) * PK_TILE;
    PK_CUDA(h, regrow(&h->dX, (size_t)cap_rows * k));
    PK_CUDA(h, regrow(&h->dy, (size_t)cap_rows));
    h->cap_rows = cap_rows;
    PK_CUDA(h, cudaMemcpyAsync(h->dX, X, sizeof(double) * n * k, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, cudaMemcpyAsync(h->dy, y, sizeof(double) * n, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    h->n = n;
    h->k = k;
    h->fit_valid = false;  // a cached factor belongs to the previous data set
    h->w_valid = false;
    h->linv_i8_valid = false;
    h->psi_n = 0;          // and so does the Psi copy
    h->psi_matches_fit = false;
    return 0;
}

static int ensure_workspace(pk_handle* h, size_t per_matrix_bytes, int C, int* chunk_out) {
    // NB: cudaMemGetInfo serialises with in-flight work (tens of ms while kernels run), so the automatic
    // limit is sampled once and cached, never on the per-generation path.
    size_t limit = h->ws_limit;
    if (limit == 0) {
        if (h->ws_limit_auto == 0) {
            size_t free_b = 0, total_b = 0;
            PK_CUDA(h, cudaMemGetInfo(&free_b, &total_b));
            size_t lim = (size_t)((double)(free_b + h->ws_bytes) * 0.6);
            const size_t cap = (size_t)64 << 30;
            h->ws_limit_auto = lim > cap ? cap : lim;
        }
        limit = h->ws_limit_auto;
    }
    size_t chunk = limit / per_matrix_bytes;
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)C) chunk = (size_t)C;
    if (chunk > 65535) chunk = 65535;  // some launchers put the candidate index on grid.y / grid.z
    const size_t need = chunk * per_matrix_bytes;
    if (need > h->ws_bytes) {
        PK_CUDA(h, cudaStreamSynchronize(h->stream));
        if (h->d_ws) {
            PK_CUDA(h, cudaFree(h->d_ws));
            h->d_ws = nullptr;
            h->ws_bytes = 0;
        }
        PK_CUDA(h, cudaMalloc((void**)&h->d_ws, need));
        h->ws_bytes = need;
    }
    *chunk_out = (int)chunk;
    return 0;
}

// Stragglers.  The INT8 kernel runs whole waves of sm_count candidates at the latency of ONE factorisation, so the few
// candidates of a short last wave cost a whole wave (150 particles at n = 4096: 2 x 43 ms for 148 + 2).  When the matrices
// are small relative to the population the ragged-first pipeline below hides that wave under the Psi kernel; when they are
// large (the Psi kernel of the whole waves is shorter than a wave) the stragglers are factorised by the DMMA task kernel,
// which spreads a handful of candidates over every SM.  Which candidates are stragglers is a property of the POPULATION
// (hinted size, global index), never of the call: a shard, a chunk and the whole population agree on every candidate's kernel
// and therefore on its bits.  Returns the global index of the first straggler (= Ceff: none).
static int64_t first_straggler(pk_handle* h, const pk::TiledShape& s, int64_t Ceff) {
    if (!h->stragglers || h->chol_variant != 0) return Ceff;
    const int64_t slots = h->sm_count, rem = Ceff % slots, whole = Ceff - rem;
    if (whole == 0 || rem == 0 || rem * 16 > slots * 7) return Ceff;  // <= 64 of 148: beyond that the short wave is worth its latency
    if (whole * (h->k + 1) * 2 >= (int64_t)3 * s.n) return Ceff;      // Psi of the whole waves outlasts a wave: ragged first hides it
    return whole;
}

// Split a resident chunk of `cc` candidates into pipeline parts: Psi of part i+1 is built on stream2 while part i is
// factorised on the main stream.
//  * INT8-tensor-core Cholesky (one CTA per candidate and SM, a wave of sm_count candidates costs the latency of ONE
//    factorisation): a short ragged wave goes FIRST -- its few CTAs factorise while the Psi kernel of the whole waves
//    has the other SMs, so the population pays for its whole waves only (default population 300 = 4 + 2 x 148).  A
//    candidate's bits do not depend on the part it is in (same kernel for every part: h->call_C).
//  * DMMA kernels: the two-part pipeline of round 1 (pk_set_overlap; measured slower -- both kernels are bound by the
//    FP64 pipe -- and off).
static void plan_parts(pk_handle* h, const pk::TiledShape& s, int cc, std::vector<int>& parts) {
    parts.clear();
    if (!h->env_split.empty()) {  // experiment knob (PK_SPLIT): comma-separated part sizes
        int left = cc;
        const char* q = h->env_split.c_str();
        while (*q && left > 0) {
            int v = atoi(q);
            if (v <= 0) break;
            if (v > left) v = left;
            parts.push_back(v);
            left -= v;
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        if (left > 0) parts.push_back(left);
        return;
    }
    if (h->profiling || h->chol_variant == 1) {
        parts.push_back(cc);
        return;
    }
    const int variant = pk::cholesky_pick_variant(h, s, cc, true);
    if (variant == 5) {
        const int slots = h->sm_count;
        const int rem = cc % slots;
        // the short wave's SM time comes out of the Psi kernel's, so what is won is wave latency x (1 - rem / slots): take it
        // up to half a wave (profiles/i8_knobs_r02_b.jsonl: 300 = 4 + 296: 5.30 -> 4.40 ms, 333 = 37 + 296: 5.52 -> 4.81)
        if (h->ragged_first && cc > slots && rem > 0 && rem * 2 <= slots) {
            parts.push_back(rem);
            parts.push_back(cc - rem);
        } else {
            parts.push_back(cc);
        }
        return;
    }
    const int slots = 2 * h->sm_count;
    if (!h->overlap || cc <= slots) {
        parts.push_back(cc);
        return;
    }
    const int rem = cc % slots;
    if (rem > 0) parts.push_back(rem);
    parts.push_back(cc - rem);
}

// stream memory operation: work queued on `stream` after this call waits until *addr >= value (cuStreamWaitValue32 through
// the runtime's driver entry point).  stream == nullptr only asks whether the operation is available.
static cudaError_t stream_wait_geq(cudaStream_t stream, unsigned int* addr, unsigned value) {
    typedef CUresult (*wait_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static wait_fn fn = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (wait_fn)f;
    }
    if (!fn) return cudaErrorNotSupported;
    if (!stream) return cudaSuccess;
    return fn((CUstream)stream, (CUdeviceptr)addr, value, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS ? cudaSuccess : cudaErrorUnknown;
}

int pk_nll_batch(pk_handle* h, int C, const double* theta, const double* p, const double* lambda, int mode,
                 double* nll, double* mu, double* sigma2, double* lndet, int32_t* info) {
    if (!h) return 1;
    if (h->n == 0) PK_FAIL(h, "pk_nll_batch: pk_set_data has not been called");
    if (C < 0 || !theta || !p) PK_FAIL(h, "pk_nll_batch: bad arguments");
    if (mode == PK_MODE_REGRESSION && !lambda) PK_FAIL(h, "pk_nll_batch: regression mode needs lambda");
    if (mode != PK_MODE_REGRESSION && mode != PK_MODE_INTERP) PK_FAIL(h, "pk_nll_batch: unknown mode %d", mode);
    if (C == 0) return 0;
    PK_CUDA(h, cudaSetDevice(h->device));
    if (ensure_candidates(h, C)) return 1;
    const int k = h->k;
    const double *d_theta, *d_p, *d_lambda;
    if (to_device(h, theta, h->d_theta, (size_t)C * k, &d_theta)) return 1;
    if (to_device(h, p, h->d_p, (size_t)C * k, &d_p)) return 1;
    if (to_device(h, mode == PK_MODE_REGRESSION ? lambda : nullptr, h->d_lambda, (size_t)C, &d_lambda)) return 1;

    if (h->n <= PK_SMALL_N_MAX) {
        pk::phase_begin(h, 1);
        PK_CUDA(h, pk::launch_small_nll(h, C, d_theta, d_p, d_lambda, mode, h->d_out, h->d_info));
        pk::phase_end(h);
    } else {
        const pk::TiledShape s = pk::tiled_shape(h->n, false);
        int chunk = 0;
        const size_t plane_bytes = pk::cholesky_i8_plane_bytes(s);
        if (ensure_workspace(h, s.elems() * sizeof(double) + plane_bytes, C, &chunk)) return 1;
        int8_t* planes0 = reinterpret_cast<int8_t*>(h->d_ws + (size_t)chunk * s.elems());  // behind the chunk's FP64 matrices
        struct CallScope {  // every chunk and part of this call picks the kernel the whole population picks
            pk_handle* h;
            CallScope(pk_handle* h_, int C_) : h(h_) { h->call_C = C_; }
            ~CallScope() { h->call_C = 0; }
        } call_scope(h, C);
        for (int c0 = 0; c0 < C; c0 += chunk) {
            const int cc = (C - c0 < chunk) ? (C - c0) : chunk;
            {
                // stragglers of the population in this chunk: one Psi launch, the whole waves on the INT8 tensor cores, the
                // stragglers in task mode
                int64_t Ceff = (h->hint_C > C) ? h->hint_C : (int64_t)C;
                const int64_t g0 = (h->hint_C > 0 ? h->hint_first : 0) + c0;
                const int64_t B = (pk::cholesky_pick_variant(h, s, cc, true) == 5) ? first_straggler(h, s, Ceff) : Ceff;
                if (B < g0 + cc) {
                    const int n_main = (B > g0) ? (int)(B - g0) : 0, n_strag = cc - n_main;
                    pk::phase_begin(h, 0);
                    cudaError_t pe = pk::launch_psi_build(h, s, cc, d_theta + (size_t)c0 * k, d_p + (size_t)c0 * k,
                                                          d_lambda ? d_lambda + c0 : nullptr, mode, h->d_ws, h->d_info + c0);
                    pk::phase_end(h);
                    PK_CUDA(h, pe);
                    pk::phase_begin(h, 1);
                    h->queue_slot = 0;
                    h->part_variant = 4;
                    pe = pk::launch_cholesky(h, s, n_strag, h->d_ws + (size_t)n_main * s.elems(), h->d_info + c0 + n_main, nullptr);
                    if (pe == cudaSuccess && n_main > 0) {
                        h->queue_slot = 1;
                        h->part_variant = 5;
                        pe = pk::launch_cholesky(h, s, n_main, h->d_ws, h->d_info + c0, planes0);
                    }
                    h->part_variant = 0;
                    h->queue_slot = 0;
                    pk::phase_end(h);
                    PK_CUDA(h, pe);
                    pk::phase_begin(h, 2);
                    PK_CUDA(h, pk::launch_nll_epilogue(h, s, cc, h->d_ws, h->d_out + c0, h->cap_C, h->d_info + c0));
                    pk::phase_end(h);
                    continue;
                }
            }
            // Parts of the resident chunk: Psi of part i+1 (stream2) runs while part i is factorised (stream).
            std::vector<int> parts;
            plan_parts(h, s, cc, parts);
            if (parts.size() > 1) {
                while (h->ev_pipe.size() < parts.size() + 1) {
                    cudaEvent_t ev;
                    PK_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                    h->ev_pipe.push_back(ev);
                }
                // stream2 starts after everything already queued on the main stream (uploads, the previous
                // chunk's factorisations that still read the workspace)
                PK_CUDA(h, cudaEventRecord(h->ev_pipe[0], h->stream));
                PK_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_pipe[0], 0));
            }
            // [short last wave, whole waves] of the INT8 kernel (plan_parts)
            const bool ragged = parts.size() == 2 && h->env_split.empty() && pk::cholesky_pick_variant(h, s, cc, true) == 5;
            unsigned int* started = h->d_queue + (PK_QUEUE_SLOTS - 1);
            if (ragged) PK_CUDA(h, cudaMemsetAsync(started, 0, sizeof(unsigned int), h->stream2));
            int o = 0;
            for (size_t pi = 0; pi < parts.size(); ++pi) {
                const int pc = parts[pi], a0 = c0 + o;
                double* wsp = h->d_ws + (size_t)o * s.elems();
                cudaStream_t main_stream = h->stream;
                if (parts.size() > 1) h->stream = h->stream2;
                pk::phase_begin(h, 0);
                cudaError_t pe = pk::launch_psi_build(h, s, pc, d_theta + (size_t)a0 * k, d_p + (size_t)a0 * k,
                                                      d_lambda ? d_lambda + a0 : nullptr, mode, wsp, h->d_info + a0);
                pk::phase_end(h);
                h->stream = main_stream;
                PK_CUDA(h, pe);
                if (parts.size() > 1) {
                    PK_CUDA(h, cudaEventRecord(h->ev_pipe[pi + 1], h->stream2));
                    PK_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_pipe[pi + 1], 0));
                }
                h->queue_slot = (int)(pi % (PK_QUEUE_SLOTS - 1));
                // ragged first: the few CTAs of the short wave need an SM each to themselves (219 KB of shared memory).  Were
                // the Psi kernel of the whole waves released right away, its CTAs would take every SM first and keep refilling
                // them, and the short wave would start when that grid runs dry -- so the Psi stream waits (device side, stream
                // memory operation) until the short wave's CTAs have counted themselves in.
                const bool gate = ragged && pi == 0 && stream_wait_geq(nullptr, nullptr, 0) == cudaSuccess;
                if (gate) h->i8_started = started;
                pk::phase_begin(h, 1);
                PK_CUDA(h, pk::launch_cholesky(h, s, pc, wsp, h->d_info + a0, planes0 + (size_t)o * plane_bytes));
                pk::phase_end(h);
                h->i8_started = nullptr;
                if (gate) PK_CUDA(h, stream_wait_geq(h->stream2, started, (unsigned)((pc < h->sm_count) ? pc : h->sm_count)));
                pk::phase_begin(h, 2);
                PK_CUDA(h, pk::launch_nll_epilogue(h, s, pc, wsp, h->d_out + a0, h->cap_C, h->d_info + a0));
                pk::phase_end(h);
                o += pc;
            }
            h->queue_slot = 0;
        }
    }
    double* outs[4] = {nll, mu, sigma2, lndet};
    bool host_out = false;
    for (int i = 0; i < 4; ++i) {
        if (!outs[i]) continue;
        PK_CUDA(h, cudaMemcpyAsync(outs[i], h->d_out + (size_t)i * h->cap_C, sizeof(double) * C, cudaMemcpyDefault,
                                   h->stream));
        host_out |= !is_device_ptr(outs[i]);
    }
    if (info) {
        PK_CUDA(h, cudaMemcpyAsync(info, h->d_info, sizeof(int32_t) * C, cudaMemcpyDefault, h->stream));
        host_out |= !is_device_ptr(info);
    }
    if (host_out) {
        PK_CUDA(h, cudaStreamSynchronize(h->stream));
        if (check_fault(h, "pk_nll_batch")) return 1;  // device outputs: reported by pk_synchronize
    }
    return 0;
}

static int ensure_fit_buffers(pk_handle* h, const pk::TiledShape& s) {
    const size_t bytes = s.elems() * sizeof(double);
    if (h->fit_bytes >= bytes && h->fit_np == s.np) return 0;
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    PK_CUDA(h, regrow(&h->d_fit, s.elems()));
    PK_CUDA(h, regrow(&h->d_try, s.elems()));
    PK_CUDA(h, regrow(&h->d_psi, (size_t)s.np * s.np));
    PK_CUDA(h, regrow(&h->d_linv, (size_t)s.np * s.np));
    PK_CUDA(h, regrow(&h->d_ad, (size_t)2 * s.np));
    PK_CUDA(h, regrow(&h->d_w, (size_t)s.np));
    h->fit_bytes = bytes;
    h->fit_np = s.np;
    h->fit_valid = false;
    h->w_valid = false;
    h->linv_i8_valid = false;
    return 0;
}

int pk_fit(pk_handle* h, const double* theta, const double* p, double lambda, int mode, double* out4,
           int32_t* info) {
    if (!h) return 1;
    if (h->n == 0) PK_FAIL(h, "pk_fit: pk_set_data has not been called");
    if (!theta || !p) PK_FAIL(h, "pk_fit: bad arguments");
    if (mode != PK_MODE_REGRESSION && mode != PK_MODE_INTERP) PK_FAIL(h, "pk_fit: unknown mode %d", mode);
    PK_CUDA(h, cudaSetDevice(h->device));
    if (ensure_candidates(h, 1)) return 1;
    const pk::TiledShape s = pk::tiled_shape(h->n, true);
    if (ensure_fit_buffers(h, s)) return 1;
    const int k = h->k;
    const double *d_theta, *d_p;
    if (to_device(h, theta, h->d_theta, (size_t)k, &d_theta)) return 1;
    if (to_device(h, p, h->d_p, (size_t)k, &d_p)) return 1;
    PK_CUDA(h, cudaMemcpyAsync(h->d_lambda, &lambda, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    PK_CUDA(h, cudaMemcpyAsync(h->d_try_hp, d_theta, sizeof(double) * k, cudaMemcpyDeviceToDevice, h->stream));
    PK_CUDA(h, cudaMemcpyAsync(h->d_try_hp + PK_MAX_K, d_p, sizeof(double) * k, cudaMemcpyDeviceToDevice, h->stream));
    PK_CUDA(h, pk::launch_psi_build(h, s, 1, d_theta, d_p, h->d_lambda, mode, h->d_try, h->d_info));
    PK_CUDA(h, pk::launch_extract_psi(h, s, h->d_try, h->d_psi, 1));
    h->psi_n = h->n;
    h->psi_matches_fit = false;  // until the factorisation below is known to have succeeded
    PK_CUDA(h, pk::launch_cholesky(h, s, 1, h->d_try, h->d_info));
    PK_CUDA(h, pk::launch_nll_epilogue(h, s, 1, h->d_try, h->d_out, h->cap_C, h->d_info));
    double host4[4];
    int32_t host_info = 0;
    for (int i = 0; i < 4; ++i)
        PK_CUDA(h, cudaMemcpyAsync(&host4[i], h->d_out + (size_t)i * h->cap_C, sizeof(double), cudaMemcpyDeviceToHost,
                                   h->stream));
    PK_CUDA(h, cudaMemcpyAsync(&host_info, h->d_info, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (host_info < 0) PK_FAIL(h, "pk_fit: gang barrier timed out (GPU shared with a long-running kernel?)");
    if (host_info == 0) {
        double* t = h->d_fit;
        h->d_fit = h->d_try;
        h->d_try = t;
        t = h->d_fit_hp;
        h->d_fit_hp = h->d_try_hp;
        h->d_try_hp = t;
        h->fit_lambda = lambda;
        h->fit_mode = mode;
        PK_CUDA(h, pk::launch_fit_finalize(h, s));
        h->fit_valid = true;
        h->fit_n = h->n;
        h->w_valid = false;
    h->linv_i8_valid = false;
        h->psi_matches_fit = true;
    }
    if (out4) PK_CUDA(h, cudaMemcpy(out4, host4, sizeof(host4), cudaMemcpyDefault));
    if (info) PK_CUDA(h, cudaMemcpy(info, &host_info, sizeof(int32_t), cudaMemcpyDefault));
    return 0;
}

int pk_append_point(pk_handle* h, const double* x_new, double y_new, double* out4, int32_t* info, int32_t* appended) {
    if (!h) return 1;
    if (!x_new || !appended) PK_FAIL(h, "pk_append_point: bad arguments");
    *appended = 0;
    if (h->n == 0) PK_FAIL(h, "pk_append_point: pk_set_data has not been called");
    // in place only with a cached factor of the current data and room inside the padded order
    if (!h->fit_valid || h->fit_n != h->n) return 0;
    // ... and with the Psi copy of THAT factor (a failed pk_fit since has replaced it: take the rebuild route)
    if (!h->psi_matches_fit || h->psi_n != h->n) return 0;
    const int n = h->n, k = h->k;
    const pk::TiledShape s = pk::tiled_shape(n, true);
    if (n + 1 > s.np || n + 1 > h->cap_rows || s.np != h->fit_np) return 0;
    PK_CUDA(h, cudaSetDevice(h->device));
    if (ensure_candidates(h, 1)) return 1;
    // d_try is scratch between fits: [psi | l | t] at its head, the grown Psi copy behind them
    double* scratch = h->d_try;
    double* psi_new = h->d_try + 2 * s.np + 8;
    PK_CUDA(h, cudaMemcpyAsync(h->dX + (size_t)n * k, x_new, sizeof(double) * k, cudaMemcpyDefault, h->stream));
    const double diag = 1.0 + (h->fit_mode == PK_MODE_REGRESSION ? h->fit_lambda : PK_EPS_DIAG);
    PK_CUDA(h, pk::launch_append_point(h, s, h->dX + (size_t)n * k, y_new, diag, scratch, h->d_info));
    int32_t host_info = 0;
    PK_CUDA(h, cudaMemcpyAsync(&host_info, h->d_info, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (info) PK_CUDA(h, cudaMemcpy(info, &host_info, sizeof(int32_t), cudaMemcpyDefault));
    if (host_info != 0) return 0;  // not positive definite: nothing was modified (the sample row is beyond n)
    PK_CUDA(h, cudaMemcpyAsync(h->dy + n, &y_new, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    PK_CUDA(h, pk::launch_append_psi_copy(h, n, diag, h->d_psi, scratch, psi_new));
    PK_CUDA(h, cudaMemcpyAsync(h->d_psi, psi_new, sizeof(double) * (size_t)(n + 1) * (n + 1), cudaMemcpyDeviceToDevice,
                               h->stream));
    h->n = n + 1;
    h->fit_n = n + 1;
    h->psi_n = n + 1;
    h->w_valid = false;
    h->linv_i8_valid = false;
    const pk::TiledShape s1 = pk::tiled_shape(n + 1, true);
    PK_CUDA(h, pk::launch_nll_epilogue(h, s1, 1, h->d_fit, h->d_out, h->cap_C, h->d_info));
    double host4[4];
    for (int i = 0; i < 4; ++i)
        PK_CUDA(h, cudaMemcpyAsync(&host4[i], h->d_out + (size_t)i * h->cap_C, sizeof(double), cudaMemcpyDeviceToHost,
                                   h->stream));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (out4) PK_CUDA(h, cudaMemcpy(out4, host4, sizeof(host4), cudaMemcpyDefault));
    *appended = 1;
    return 0;
}

int pk_predict_batch(pk_handle* h, int64_t m, const double* Xq, const double* theta, const double* p, double mu,
                     double sigma2, double y_min, double ei_weight, double var_extra, double* yhat, double* s,
                     double* ei) {
    if (!h) return 1;
    if (!h->fit_valid || h->fit_n != h->n) PK_FAIL(h, "pk_predict_batch: no cached factor (call pk_fit first)");
    if (m < 0 || !Xq || !theta || !p) PK_FAIL(h, "pk_predict_batch: bad arguments");
    if (m == 0) return 0;
    PK_CUDA(h, cudaSetDevice(h->device));
    const int k = h->k;
    const pk::TiledShape sh = pk::tiled_shape(h->n, true);
    PK_CUDA(h, cudaMemcpyAsync(h->d_hp, theta, sizeof(double) * k, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, cudaMemcpyAsync(h->d_hp + PK_MAX_K, p, sizeof(double) * k, cudaMemcpyDefault, h->stream));
    if (!h->w_valid || !(h->w_mu == mu)) {
        PK_CUDA(h, pk::launch_compute_w(h, sh, mu));
        h->w_mu = mu;
        h->w_valid = true;
    }
    const bool q_dev = is_device_ptr(Xq);
    const bool out_dev = (!yhat || is_device_ptr(yhat)) && (!s || is_device_ptr(s)) && (!ei || is_device_ptr(ei));
    if (q_dev && out_dev) {
        pk::phase_begin(h, 3);
        PK_CUDA(h, pk::launch_predict(h, m, Xq, h->d_hp, mu, sigma2, y_min, ei_weight, var_extra, yhat, s, ei));
        pk::phase_end(h);
        return 0;
    }
    // Host path: the queries stream through device staging buffers in chunks of 1 Mi points, double buffered -- the
    // upload of chunk c+1 and the download of chunk c-1 (copy stream) run under the kernels of chunk c (main stream).
    const int64_t chunk = m < ((int64_t)1 << 20) ? m : ((int64_t)1 << 20);
    const int nbuf = (m > chunk) ? 2 : 1;
    if (!q_dev && nbuf * chunk * k > h->cap_q) {
        PK_CUDA(h, cudaStreamSynchronize(h->stream));
        PK_CUDA(h, regrow(&h->d_q, (size_t)nbuf * chunk * k));
        h->cap_q = nbuf * chunk * k;
    }
    if (nbuf * chunk > h->cap_qout) {  // sized on its own: k may have changed since the query buffer last grew
        PK_CUDA(h, cudaStreamSynchronize(h->stream));
        PK_CUDA(h, regrow(&h->d_qout, (size_t)nbuf * chunk * 3));
        h->cap_qout = nbuf * chunk;
    }
    while (h->ev_pred.size() < 7) {
        cudaEvent_t ev;
        PK_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        h->ev_pred.push_back(ev);
    }
    cudaEvent_t* ev_in = h->ev_pred.data();        // [2] chunk uploaded
    cudaEvent_t* ev_done = h->ev_pred.data() + 2;  // [2] chunk computed
    cudaEvent_t* ev_free = h->ev_pred.data() + 4;  // [2] chunk downloaded: its output buffer may be overwritten
    // the copy stream starts behind everything already queued on the main stream (hyper-parameters, w)
    PK_CUDA(h, cudaEventRecord(h->ev_pred[6], h->stream));
    PK_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_pred[6], 0));
    const int64_t nchunks = (m + chunk - 1) / chunk;
    auto upload = [&](int64_t c) -> cudaError_t {
        const int b = (int)(c % nbuf);
        const int64_t q0 = c * chunk, mc = (m - q0 < chunk) ? (m - q0) : chunk;
        if (!q_dev) {
            cudaError_t e = cudaMemcpyAsync(h->d_q + (size_t)b * chunk * k, Xq + q0 * k, sizeof(double) * mc * k,
                                            cudaMemcpyDefault, h->stream2);
            if (e != cudaSuccess) return e;
        }
        return cudaEventRecord(ev_in[b], h->stream2);
    };
    PK_CUDA(h, upload(0));
    for (int64_t c = 0; c < nchunks; ++c) {
        const int b = (int)(c % nbuf);
        const int64_t q0 = c * chunk, mc = (m - q0 < chunk) ? (m - q0) : chunk;
        // next upload first: on the copy stream it sits in front of this chunk's download, which waits for the kernels
        if (c + 1 < nchunks) PK_CUDA(h, upload(c + 1));
        const double* xq_d = q_dev ? Xq + q0 * k : h->d_q + (size_t)b * chunk * k;
        double* ob = h->d_qout + (size_t)b * 3 * chunk;
        double* o0 = yhat ? ob : nullptr;
        double* o1 = s ? ob + chunk : nullptr;
        double* o2 = ei ? ob + 2 * chunk : nullptr;
        PK_CUDA(h, cudaStreamWaitEvent(h->stream, ev_in[b], 0));
        if (c >= nbuf) PK_CUDA(h, cudaStreamWaitEvent(h->stream, ev_free[b], 0));
        pk::phase_begin(h, 3);
        PK_CUDA(h, pk::launch_predict(h, mc, xq_d, h->d_hp, mu, sigma2, y_min, ei_weight, var_extra, o0, o1, o2));
        pk::phase_end(h);
        PK_CUDA(h, cudaEventRecord(ev_done[b], h->stream));
        PK_CUDA(h, cudaStreamWaitEvent(h->stream2, ev_done[b], 0));
        if (yhat) PK_CUDA(h, cudaMemcpyAsync(yhat + q0, o0, sizeof(double) * mc, cudaMemcpyDefault, h->stream2));
        if (s) PK_CUDA(h, cudaMemcpyAsync(s + q0, o1, sizeof(double) * mc, cudaMemcpyDefault, h->stream2));
        if (ei) PK_CUDA(h, cudaMemcpyAsync(ei + q0, o2, sizeof(double) * mc, cudaMemcpyDefault, h->stream2));
        PK_CUDA(h, cudaEventRecord(ev_free[b], h->stream2));
    }
    PK_CUDA(h, cudaStreamSynchronize(h->stream2));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int pk_get_psi(pk_handle* h, double* out) {
    if (!h || !out) return 1;
    if (!h->d_psi || h->fit_np == 0 || h->psi_n != h->n)
        PK_FAIL(h, "pk_get_psi: no Psi for the current data (pk_fit has not been called since pk_set_data)");
    PK_CUDA(h, cudaSetDevice(h->device));
    PK_CUDA(h, cudaMemcpyAsync(out, h->d_psi, sizeof(double) * h->n * h->n, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int pk_get_U(pk_handle* h, double* out) {
    if (!h || !out) return 1;
    if (!h->fit_valid || h->fit_n != h->n) PK_FAIL(h, "pk_get_U: no successful pk_fit yet");
    PK_CUDA(h, cudaSetDevice(h->device));
    const pk::TiledShape s = pk::tiled_shape(h->n, true);
    if (is_device_ptr(out)) {
        PK_CUDA(h, pk::launch_extract_U(h, s, out));
        return 0;
    }
    // d_try is scratch between fits: borrow its head for the n x n result
    PK_CUDA(h, pk::launch_extract_U(h, s, h->d_try));
    PK_CUDA(h, cudaMemcpyAsync(out, h->d_try, sizeof(double) * h->n * h->n, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int pk_psi_batch(pk_handle* h, int C, const double* theta, const double* p, const double* lambda, int mode,
                 double* out) {
    if (!h) return 1;
    if (h->n == 0) PK_FAIL(h, "pk_psi_batch: pk_set_data has not been called");
    if (C < 0 || !theta || !p || !out) PK_FAIL(h, "pk_psi_batch: bad arguments");
    if (mode == PK_MODE_REGRESSION && !lambda) PK_FAIL(h, "pk_psi_batch: regression mode needs lambda");
    if (C == 0) return 0;
    PK_CUDA(h, cudaSetDevice(h->device));
    if (ensure_candidates(h, C)) return 1;
    const int k = h->k;
    const double *d_theta, *d_p, *d_lambda;
    if (to_device(h, theta, h->d_theta, (size_t)C * k, &d_theta)) return 1;
    if (to_device(h, p, h->d_p, (size_t)C * k, &d_p)) return 1;
    if (to_device(h, mode == PK_MODE_REGRESSION ? lambda : nullptr, h->d_lambda, (size_t)C, &d_lambda)) return 1;
    // Psi is built by the same population kernel the likelihood path uses (lower-triangle tiles in the
    // workspace), then mirrored into full symmetric matrices.
    const size_t per = (size_t)h->n * h->n;
    const pk::TiledShape s = pk::tiled_shape(h->n, false);
    const bool out_dev = is_device_ptr(out);
    int chunk = 0;
    if (ensure_workspace(h, (s.elems() + (out_dev ? 0 : per)) * sizeof(double), C, &chunk)) return 1;
    double* stage = h->d_ws + (size_t)chunk * s.elems();
    for (int c0 = 0; c0 < C; c0 += chunk) {
        const int cc = (C - c0 < chunk) ? (C - c0) : chunk;
        pk::phase_begin(h, 0);
        PK_CUDA(h, pk::launch_psi_build(h, s, cc, d_theta + (size_t)c0 * k, d_p + (size_t)c0 * k,
                                        d_lambda ? d_lambda + c0 : nullptr, mode, h->d_ws, h->d_info + c0));
        pk::phase_end(h);
        if (out_dev) {
            PK_CUDA(h, pk::launch_extract_psi(h, s, h->d_ws, out + (size_t)c0 * per, cc));
        } else {
            PK_CUDA(h, pk::launch_extract_psi(h, s, h->d_ws, stage, cc));
            PK_CUDA(h, cudaMemcpyAsync(out + (size_t)c0 * per, stage, sizeof(double) * per * cc, cudaMemcpyDefault,
                                       h->stream));
        }
    }
    if (!out_dev) PK_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}


// ---- device-resident optimiser population ---------------------------------------------------------------------------
static int opt_buffer_ok(pk_handle* h, int which) { return h->opt_n > 0 && which >= 0 && which <= 3 && h->d_opt[which]; }

int pk_opt_set(pk_handle* h, int which, int n, int dims, const double* rows) {
    if (!h) return 1;
    if (which < 0 || which > 3 || n < 1 || dims < 1 || !rows) PK_FAIL(h, "pk_opt_set: bad arguments");
    PK_CUDA(h, cudaSetDevice(h->device));
    if (which == 0) {
        const int need = n * dims;
        if (need > h->opt_cap) {
            PK_CUDA(h, cudaStreamSynchronize(h->stream));
            for (int b = 0; b < 5; ++b) PK_CUDA(h, regrow(&h->d_opt[b], (size_t)need));
            h->opt_cap = need;
        }
        h->opt_n = n;
        h->opt_dims = dims;
        h->opt_pso_started = false;
    } else if (n != h->opt_n || dims != h->opt_dims) {
        PK_FAIL(h, "pk_opt_set: buffer %d must have the population's shape (%d x %d)", which, h->opt_n, h->opt_dims);
    }
    PK_CUDA(h, cudaMemcpyAsync(h->d_opt[which], rows, sizeof(double) * (size_t)n * dims, cudaMemcpyDefault, h->stream));
    if (!is_device_ptr(rows)) PK_CUDA(h, cudaStreamSynchronize(h->stream));  // the caller may reuse its buffer
    return 0;
}

int pk_opt_get(pk_handle* h, int which, int row0, int nrows, double* out) {
    if (!h) return 1;
    if (!opt_buffer_ok(h, which) || row0 < 0 || nrows < 0 || row0 + nrows > h->opt_n || !out) PK_FAIL(h, "pk_opt_get: bad arguments");
    if (nrows == 0) return 0;
    PK_CUDA(h, cudaSetDevice(h->device));
    PK_CUDA(h, cudaMemcpyAsync(out, h->d_opt[which] + (size_t)row0 * h->opt_dims, sizeof(double) * (size_t)nrows * h->opt_dims,
                               cudaMemcpyDefault, h->stream));
    if (!is_device_ptr(out)) PK_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int pk_opt_evaluate(pk_handle* h, int which, int row0, int count, int mode, double* nll, double* mu, double* sigma2,
                    double* lndet, int32_t* info) {
    if (!h) return 1;
    if (!opt_buffer_ok(h, which) || row0 < 0 || count < 0 || row0 + count > h->opt_n) PK_FAIL(h, "pk_opt_evaluate: bad arguments");
    if (h->n == 0) PK_FAIL(h, "pk_opt_evaluate: pk_set_data has not been called");
    const int k = h->k;
    if (h->opt_dims != 2 * k + (mode == PK_MODE_REGRESSION ? 1 : 0))
        PK_FAIL(h, "pk_opt_evaluate: candidates of %d numbers do not fit k = %d in mode %d", h->opt_dims, k, mode);
    if (count == 0) return 0;
    PK_CUDA(h, cudaSetDevice(h->device));
    if (ensure_candidates(h, count)) return 1;
    PK_CUDA(h, pk::launch_opt_split(h, h->d_opt[which] + (size_t)row0 * h->opt_dims, count, k, h->opt_dims, h->d_theta, h->d_p,
                                    mode == PK_MODE_REGRESSION ? h->d_lambda : nullptr));
    return pk_nll_batch(h, count, h->d_theta, h->d_p, mode == PK_MODE_REGRESSION ? h->d_lambda : nullptr, mode, nll, mu, sigma2,
                        lndet, info);
}

static int opt_upload_idx(pk_handle* h, const int32_t* const* src, const int* count, int parts, int32_t** dst) {
    size_t total = 0;
    for (int i = 0; i < parts; ++i) total += (size_t)(src[i] ? count[i] : 0);
    if (total > h->opt_idx_cap) {
        PK_CUDA(h, cudaStreamSynchronize(h->stream));
        PK_CUDA(h, regrow(&h->d_opt_idx, total));
        h->opt_idx_cap = total;
    }
    size_t off = 0;
    bool host = false;
    for (int i = 0; i < parts; ++i) {
        dst[i] = nullptr;
        if (!src[i] || count[i] == 0) continue;
        dst[i] = h->d_opt_idx + off;
        PK_CUDA(h, cudaMemcpyAsync(dst[i], src[i], sizeof(int32_t) * (size_t)count[i], cudaMemcpyDefault, h->stream));
        host |= !is_device_ptr(src[i]);
        off += (size_t)count[i];
    }
    if (host) PK_CUDA(h, cudaStreamSynchronize(h->stream));  // pageable sources may be reused by the caller
    return 0;
}

int pk_ga_step(pk_handle* h, int nc, const int32_t* keep, int nsel, const int32_t* parents, const int32_t* cut) {
    if (!h) return 1;
    if (!opt_buffer_ok(h, 0) || nc < 0 || nc > h->opt_n || nsel < 0 || nsel > h->opt_n || (nsel & 1) || (nsel > 0 && (!parents || !cut)))
        PK_FAIL(h, "pk_ga_step: bad arguments");
    PK_CUDA(h, cudaSetDevice(h->device));
    const int n = h->opt_n, dims = h->opt_dims;
    const int32_t* src[3] = {keep, parents, cut};
    const int cnt[3] = {n, nsel, nsel / 2};
    int32_t* dev[3];
    if (opt_upload_idx(h, src, cnt, 3, dev)) return 1;
    if (keep) {
        // generational replacement: the survivors of [offspring | old population], in their new (best-first) order
        PK_CUDA(h, pk::launch_opt_gather(h, h->d_opt[4], h->d_opt[1], h->d_opt[0], dev[0], nc, n, dims));
        double* t = h->d_opt[0];
        h->d_opt[0] = h->d_opt[4];
        h->d_opt[4] = t;
    }
    if (nsel > 0) PK_CUDA(h, pk::launch_ga_cross(h, h->d_opt[1], h->d_opt[0], dev[1], dev[2], nsel / 2, dims));
    return 0;
}

int pk_pso_step(pk_handle* h, const double* r, const int32_t* nbest, const int32_t* improved, double inertia, double cognitive,
                double social, const double* lo, const double* hi) {
    if (!h) return 1;
    if (!opt_buffer_ok(h, 0) || !r || !nbest || !lo || !hi) PK_FAIL(h, "pk_pso_step: bad arguments");
    PK_CUDA(h, cudaSetDevice(h->device));
    const int n = h->opt_n, dims = h->opt_dims;
    const size_t nr = (size_t)n * dims * 2, need = nr + 2 * (size_t)dims;
    if (need > h->opt_r_cap) {
        PK_CUDA(h, cudaStreamSynchronize(h->stream));
        PK_CUDA(h, regrow(&h->d_opt_r, need));
        h->opt_r_cap = need;
    }
    PK_CUDA(h, cudaMemcpyAsync(h->d_opt_r, r, sizeof(double) * nr, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, cudaMemcpyAsync(h->d_opt_r + nr, lo, sizeof(double) * dims, cudaMemcpyDefault, h->stream));
    PK_CUDA(h, cudaMemcpyAsync(h->d_opt_r + nr + dims, hi, sizeof(double) * dims, cudaMemcpyDefault, h->stream));
    const int32_t* src[2] = {nbest, improved};
    const int cnt[2] = {n, n};
    int32_t* dev[2];
    if (opt_upload_idx(h, src, cnt, 2, dev)) return 1;
    PK_CUDA(h, pk::launch_pso_step(h, h->d_opt[0], h->d_opt[2], h->d_opt[3], h->d_opt_r, dev[0], dev[1], h->opt_pso_started ? 0 : 1,
                                   inertia, cognitive, social, h->d_opt_r + nr, h->d_opt_r + nr + dims, n, dims));
    h->opt_pso_started = true;
    return 0;
}

}  // extern "C"
