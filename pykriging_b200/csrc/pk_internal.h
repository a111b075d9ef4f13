// Host-side internals shared by the translation units of libpykriging_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pykriging_b200.h"
#include "pk_common.cuh"

constexpr int PK_QUEUE_SLOTS = 64;  // the last counter is the "CTAs resident" word of the ragged-first pipeline (pk_nll_batch)
constexpr int PK_MAX_DEVICES = 64;  // per-device "function attributes set" flags (cudaFuncSetAttribute is per device)

struct pk_handle {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;   // recorded events, 2 per phase instance
    std::vector<int> ev_phase;          // phase id of every event pair
    size_t ev_used = 0;
    double pivot_tol = 2.0 * PK_EPS_DIAG;  // relative pivot threshold, 2 ulps (see pk_set_pivot_tolerance)
    int chol_variant = 0;                  // 0 auto, 1 one launch per block column, 2 persistent CTA per candidate, 3 gangs, 4 task mode, 5 INT8 tensor cores
    int predict_variant = 0;               // 0 auto, 1 shared-memory psi tile kernel, 2 register-resident (n <= 512), 3 split (psi rows kernel + TMA-fed DMMA kernel)
    // this handle's calls are slices of a population / query set of these sizes (pk_set_shard_hint; 0 = no hint):
    // automatic kernel selection uses them, so a shard computes bit for bit what the unsharded call would
    int64_t hint_C = 0, hint_m = 0;
    int64_t hint_first = 0;                // ... and where the slice starts in that population (pk_set_shard_offset)
    int part_variant = 0;                  // transient: the Cholesky kernel of the next launch_cholesky (pk_nll_batch runs the stragglers of a population in task mode)
    int stragglers = 1;                    // PK_STRAGGLERS=0: never peel the candidates of a short last wave off the INT8 kernel
    int64_t call_C = 0;                    // population of the pk_nll_batch call in flight: its chunks and pipeline parts pick the kernel the whole call picks
    size_t gang_bytes = 0;
    void* d_gang = nullptr;                // gang barrier counters + partial-sum scratch of the persistent Cholesky (gang mode)
    unsigned int* d_queue = nullptr;       // candidate queues of the persistent Cholesky kernel (PK_QUEUE_SLOTS counters)
    int queue_slot = 0;                    // counter the next launch_cholesky_cta uses
    // task mode of the persistent Cholesky (cta::chol_task_kernel): per-shape task list + per-launch completion flags
    // and inverted diagonal blocks
    void* d_tasks = nullptr;
    int task_nt = 0, task_single = -1, task_bpt = 0, task_count = 0;  // what the cached list was built for
    int task_single_below = -1;            // block columns with at most this many row blocks left use 64-row tasks (-1 = automatic)
    int task_blocks = 0;                   // row blocks per panel task (0 = automatic)
    int task_cohort = 0;                   // candidates that walk the task list together (0 = the whole call)
    int task_skew = 0;                     // member m of a cohort runs m % skew entries behind the step (0 = automatic)
    void* d_taskws = nullptr;
    size_t task_bytes = 0;
    // Psi construction of the NEXT part of a population overlaps the Cholesky of the current one: the Psi
    // kernels are issued on stream2, the factorisations on stream, ordered by events (pk_nll_batch)
    cudaStream_t stream2 = nullptr;
    std::vector<cudaEvent_t> ev_pipe;
    std::vector<cudaEvent_t> ev_pred;       // upload / compute / download events of the chunked host path of pk_predict_batch
    int overlap = 0;                       // 1: two-part pipeline on two streams (pk_set_overlap; measured slower, off)
    int ragged_first = 1;                  // INT8 Cholesky: a short last wave is factorised first, under the Psi kernel of the whole waves (PK_RAGGED_FIRST=0: off)
    // experiment knobs, read from the environment ONCE in pk_create (never on the per-generation path)
    std::string env_split;                 // PK_SPLIT: comma-separated part sizes of a resident chunk
    int env_gang = 0;                      // PK_CHOL_GANG: forced gang size (0 = automatic)
    int env_ctas_per_sm = 2;               // PK_CHOL_CTAS_PER_SM
    bool env_cta_timing = false;           // PK_CTA_TIMING: per-phase cycle counts of the persistent kernel on stderr
    bool env_cta_alias = false;            // PK_CTA_ALIAS: timing experiment (results are garbage)
    unsigned int* i8_started = nullptr;    // when set, every CTA of the next INT8 Cholesky launch counts itself in here as it starts (and the pointer is cleared)
    int psi_ppt_small = 1;                 // Psi kernel, k <= 6: sample pairs per thread; 4 was round 1's choice (175 registers, one CTA per SM), 1 is 7 - 38 % faster (PK_PSI_PPT_SMALL=4 pins the old one; same bits)
    int psi_const_min_k = 5;               // smallest k that takes the constant-memory path (PK_PSI_CONST_MINK; k <= 4 measured 5 - 10 % slower that way)
    int psi_const = 1;                     // Psi kernel, psi_const_min_k <= k <= 10: (theta, p) through constant memory, one launch per candidate group (PK_PSI_CONST=0: shared memory)
    double* d_thp = nullptr;               // (theta_d, p_d) pairs of the population (source of the constant-memory copies)
    size_t thp_cap = 0;
    cudaStream_t psi_stream[2] = {nullptr, nullptr};  // side streams of the per-group Psi launches (forked from / joined to the caller's stream)
    cudaEvent_t psi_event[3] = {nullptr, nullptr, nullptr};
    int i8_quiet = 1;                      // INT8 Cholesky, tile factorisation: warp 4 stays out of the trailing update (PK_I8_QUIET=0: all seven warps)
    int i8_perm = 1;                       // INT8 Cholesky: permuted digit-plane layout, shuffle-free triangular solve (PK_I8_PERM=0: first layout)
    int i8_l2_mode = 0;                    // INT8 Cholesky, L2 policy of the digit tiles: bit 0 rows of a pass evict-first, bit 1 row block j evict-last (PK_I8_L2)
    int* d_fault = nullptr;                // sticky device flag: some candidate came back with info < 0 (a spin-wait timed out)

    // model data (pk_set_data)
    int n = 0, k = 0;
    double* dX = nullptr;  // n x k, model units
    double* dy = nullptr;  // n

    // hyper-parameter / result staging for host callers
    int cap_C = 0;
    double* d_theta = nullptr;   // cap_C x k
    double* d_p = nullptr;       // cap_C x k
    double* d_lambda = nullptr;  // cap_C
    double* d_out = nullptr;     // 4 x cap_C  (nll | mu | sigma2 | lndet)
    int32_t* d_info = nullptr;   // cap_C
    int cap_k = 0;

    // workspace of the tiled path: chunk_C matrices of (rows x ld) doubles
    double* d_ws = nullptr;
    size_t ws_bytes = 0;
    size_t ws_limit = 0;
    size_t ws_limit_auto = 0;  // cached default (60% of free memory at first use)

    // cached fit (pk_fit) for the predictors
    int fit_n = 0;            // n the cache was built for (0 = none)
    int fit_np = 0;           // padded order (multiple of PK_TILE)
    double* d_fit = nullptr;  // factorised augmented matrix of the last successful fit
    double* d_try = nullptr;  // scratch for the attempt (swapped in on success)
    double* d_psi = nullptr;  // n x n Psi of the last attempt (row-major, full symmetric)
    int psi_n = 0;            // order of the matrix in d_psi (0 = none: pk_set_data drops it)
    bool psi_matches_fit = false;  // d_psi is the Psi of the CACHED factor (false after a failed attempt)
    double* d_linv = nullptr; // fit_np x fit_np row-major L^-1 (lower)
    double* d_ad = nullptr;   // 2 x fit_np : a = L^-1 1, d = L^-1 y
    double* d_w = nullptr;    // fit_np : w = L^-T (d - mu a)
    double* d_fit_hp = nullptr;  // 2 x PK_MAX_K: (theta | p) of the cached factor (pk_append_point)
    double* d_try_hp = nullptr;  // ... of the attempt
    double fit_lambda = 0.0;
    int fit_mode = 0;
    int cap_rows = 0;            // sample capacity of dX / dy (rows)
    double w_mu = 0.0;
    bool w_valid = false;
    bool fit_valid = false;
    size_t fit_bytes = 0;

    // device-resident optimiser population (pk_opt_*, pk_ga_step, pk_pso_step): rows of `opt_dims` doubles
    int opt_n = 0, opt_dims = 0, opt_cap = 0;   // cap in doubles per buffer
    double* d_opt[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // 0 population / x, 1 offspring, 2 previous x, 3 archive (pbest), 4 scratch
    int32_t* d_opt_idx = nullptr;               // index / mask uploads of a step
    double* d_opt_r = nullptr;                  // uniforms + bounds of a PSO step
    size_t opt_idx_cap = 0, opt_r_cap = 0;
    bool opt_pso_started = false;

    // staging for predict
    double* d_q = nullptr;    // query chunk
    double* d_qout = nullptr; // 3 x chunk
    int64_t cap_q = 0;      // doubles in d_q
    int64_t cap_qout = 0;   // query points d_qout has room for (3 outputs each)
    double* d_hp = nullptr;   // 2 x PK_MAX_K (theta | p) for predict
    int* d_gridinfo = nullptr;     // {period P, ok}: verdict of the device-side grid detection (pk_predict2.cu)
    int predict_grid_mode = 1;     // 0: never use the grid-mode psi kernel (tests compare the two)
    // INT8-tensor-core contraction of the split predictor (pk_predict_i8.cu): digit planes of L^-1 (+ row scales) and of psi
    int8_t* d_linv_i8 = nullptr;
    double* d_linv_scale = nullptr;
    size_t cap_linv_i8 = 0;
    bool linv_i8_valid = false;
    int8_t* d_psi_i8 = nullptr;
    size_t cap_psi_i8 = 0;
    int predict_contraction = 0;   // 0 automatic (INT8 tensor cores), 1 FP64 DMMA (tn::trinorm_kernel), 2 INT8
    double* d_pscratch = nullptr;  // psi slabs of predict_big_kernel (np > 2944): grid x 32 x (np + 4)
    size_t cap_pscratch = 0;
};

namespace pk {

// Layout of one augmented matrix of the tiled path (row-major, leading dimension ld = np):
//   rows [0, n)            Psi (lower triangle used; the factor L overwrites it)
//   rows [n, np)           identity padding up to a multiple of PK_TILE
//   rows [np, np+8)        augmented rows: 1^T, y^T, 0...   -> become a^T = (L^-1 1)^T, d^T = (L^-1 y)^T
//   rows [np+8, np+8+np)   (fit only) identity               -> becomes L^-T, i.e. (L^-1) transposed
struct TiledShape {
    int n, np, ld, rows;  // rows = total row count (multiple of 8)
    bool with_inverse;
    size_t elems() const { return (size_t)rows * (size_t)ld; }
};
inline TiledShape tiled_shape(int n, bool with_inverse) {
    TiledShape s;
    s.n = n;
    s.np = ((n + PK_TILE - 1) / PK_TILE) * PK_TILE;
    s.ld = s.np;
    s.with_inverse = with_inverse;
    s.rows = s.np + PK_AUG_ROWS + (with_inverse ? s.np : 0);
    return s;
}

// ---- launchers (each returns cudaError_t of the launch) -------------------------------------------
cudaError_t launch_small_nll(pk_handle* h, int C, const double* theta, const double* p, const double* lambda,
                             int mode, double* out4, int32_t* info);

cudaError_t launch_psi_build(pk_handle* h, const TiledShape& s, int C, const double* theta, const double* p,
                             const double* lambda, int mode, double* ws, int32_t* info);
// planes: digit planes of the INT8-tensor-core Cholesky (cholesky_i8_plane_bytes per candidate) or null (DMMA kernels only)
cudaError_t launch_cholesky(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info, int8_t* planes = nullptr);
// the Cholesky kernel launch_cholesky runs for a call of C candidates (1 per-block-column launches, 2 one CTA per candidate,
// 3 gangs, 4 task mode, 5 INT8 tensor cores); looks at the hinted / whole-call population, never at the size of a part
int cholesky_pick_variant(pk_handle* h, const TiledShape& s, int C, bool have_planes);
size_t cholesky_i8_plane_bytes(const TiledShape& s);
cudaError_t launch_cholesky_i8(pk_handle* h, const TiledShape& s, int C, double* ws, int8_t* planes, int32_t* info);
cudaError_t launch_cholesky_steps(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info);
cudaError_t launch_cholesky_cta(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info);
cudaError_t launch_cholesky_task(pk_handle* h, const TiledShape& s, int C, double* ws, int32_t* info);
// task list of one candidate as (type, j, row block, blocks) quadruples; returns false if it is not a valid order
bool task_schedule_flat(int nt, int single_below, int blocks_per_task, std::vector<int>& out4);
int cholesky_gang_size(const pk_handle* h, const TiledShape& s, int64_t Ceff, bool forced);
cudaError_t launch_nll_epilogue(pk_handle* h, const TiledShape& s, int C, const double* ws, double* out4,
                                int out_stride, int32_t* info);
cudaError_t launch_extract_psi(pk_handle* h, const TiledShape& s, const double* M, double* out, int C);
cudaError_t launch_fit_finalize(pk_handle* h, const TiledShape& s);
cudaError_t launch_compute_w(pk_handle* h, const TiledShape& s, double mu);
cudaError_t launch_predict(pk_handle* h, int64_t m, const double* Xq, const double* hp, double mu, double sigma2,
                           double y_min, double ei_weight, double var_extra, double* yhat, double* sout,
                           double* ei);
cudaError_t launch_predict_split(pk_handle* h, int64_t m, const double* Xq, const double* hp, double mu, double sigma2,
                                 double y_min, double ei_weight, double var_extra, double* yhat, double* sout,
                                 double* ei);
cudaError_t psi_planes_i8(pk_handle* h, int64_t mcap, uint8_t** planes, size_t* plane_stride);
cudaError_t launch_trinorm_i8(pk_handle* h, int64_t mc, int64_t mcap, double sigma2, double y_min, double ei_weight,
                              double var_extra, const double* yhat, double* sout, double* ei);
cudaError_t launch_extract_U(pk_handle* h, const TiledShape& s, double* out);
cudaError_t launch_append_point(pk_handle* h, const TiledShape& s, const double* xnew_dev, double y_new, double diag,
                                double* scratch, int32_t* info);
cudaError_t launch_append_psi_copy(pk_handle* h, int n, double diag, const double* old, const double* psi, double* out);
cudaError_t measure_fp64_peak(pk_handle* h, double* dmma, double* dfma);
cudaError_t launch_opt_gather(pk_handle* h, double* dst, const double* child, const double* pop, const int32_t* keep, int nc,
                              int n, int dims);
cudaError_t launch_ga_cross(pk_handle* h, double* child, const double* pop, const int32_t* parents, const int32_t* cut,
                            int npairs, int dims);
cudaError_t launch_pso_step(pk_handle* h, double* x, double* prev, double* arch, const double* r, const int32_t* nbest,
                            const int32_t* improved, int first, double inertia, double cognitive, double social,
                            const double* lo, const double* hi, int n, int dims);
cudaError_t launch_opt_split(pk_handle* h, const double* cand, int n, int k, int dims, double* theta, double* p, double* lambda);

// phase timing helpers (no-ops unless h->profiling)
void phase_begin(pk_handle* h, int phase);
void phase_end(pk_handle* h);

}  // namespace pk
