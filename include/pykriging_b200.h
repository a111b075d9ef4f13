/*
 * pykriging_b200 -- C ABI of the B200-native pyKriging fit/predict hot path.
 *
 * This is the drop-in boundary for the reference's `matrixops` mixin
 * (capaulson/pyKriging, pyKriging/matrixops.py) and for the per-candidate / per-point
 * Python loops that drive it (pyKriging/krige.py, pyKriging/regressionkrige.py).
 * Every entry point cites the reference interface it replaces.  The reference is pure
 * Python, so the binding a maintainer adds is a ctypes stub.
 *
 * Conventions
 *   - All arithmetic is FP64.  Matrices are row-major (C order), like the reference's
 *     NumPy arrays.  Coordinates are in MODEL units (already min-max normalised, as
 *     self.X / self.y are after normalizeData(), krige.py:117-133).
 *   - Every data pointer may be a HOST pointer or a DEVICE pointer of the handle's GPU;
 *     the library detects which.  Host buffers are copied in/out inside the call (the
 *     call returns after outputs are valid); device buffers are used in place and the
 *     call is stream-ordered (use pk_synchronize()).
 *   - The caller owns every buffer it passes.  Nothing is retained past a call except
 *     the copies made by pk_set_data().
 *   - Return value: 0 = OK; non-zero = argument / CUDA error (fatal; text from
 *     pk_last_error()).  NUMERICAL failure (Psi not positive definite) is NOT an error
 *     status: it is reported per candidate in info[] with LAPACK's dpotrf convention
 *     (0 ok, j>0 = order of the first non-positive pivot), and the Python glue maps it to
 *     the reference's penalty 10000 (krige.py:447-450) or to its stale-U rule
 *     (regressionkrige.py:167-172).
 *   - One handle per (process, GPU).  A handle is not thread-safe.  There is no CPU
 *     fallback: without a CUDA device pk_create() fails.
 */
#ifndef PYKRIGING_B200_H
#define PYKRIGING_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pk_handle pk_handle;

#define PK_MODE_INTERP 0     /* matrixops.updatePsi:    diag = 1 + np.spacing(1)  (matrixops.py:30) */
#define PK_MODE_REGRESSION 1 /* matrixops.regupdatePsi: diag = 1 + Lambda         (matrixops.py:40) */

#define PK_VERSION 100

/* ABI version (PK_VERSION of the built library). */
int pk_version(void);

/* Last error text of a handle (or of the last failed pk_create when h == NULL). */
const char *pk_last_error(const pk_handle *h);

/* Replaces matrixops.__init__ (matrixops.py:7-16): creates the per-model device state on GPU
 * `device` (workspaces are grown lazily).  Fails (non-zero) when no CUDA device is usable. */
int pk_create(pk_handle **out, int device);
int pk_destroy(pk_handle *h);

/* Replaces matrixops.updateData (matrixops.py:18-22).  Uploads the n x k sample matrix X (model
 * units, row-major) and the n observations y.  The reference's (n,n,k) `distance` tensor is never
 * materialised; only O(n k) data is kept.  Also used after addPoint() (krige.py:135-156). */
int pk_set_data(pk_handle *h, int n, int k, const double *X, const double *y);

/* Replaces the candidate loop of kriging.fittingObjective (krige.py:429-452) and
 * regression_kriging.fittingObjective (regressionkrige.py:428-452), i.e. per candidate
 * updatePsi/regupdatePsi (matrixops.py:24-42) + neglikelihood/regneglikelihood (matrixops.py:45-66),
 * for a whole GA/PSO population in one call.
 *   theta, p : C x k (row-major);  lambda : C (mode REGRESSION) or NULL
 *   outputs (each may be NULL): nll, mu, sigma2, lndet : C doubles; info : C int32.
 * For info[c] != 0 the four doubles of candidate c are NaN. */
int pk_nll_batch(pk_handle *h, int C, const double *theta, const double *p, const double *lambda, int mode,
                 double *nll, double *mu, double *sigma2, double *lndet, int32_t *info);

/* Replaces update()/updateModel() + neglikelihood() for ONE hyper-parameter set
 * (krige.py:158-178, regressionkrige.py:151-172, matrixops.py:24-66) and caches what the predictors
 * need: the Cholesky factor L (U = L^T), L^-1, a = L^-1 1, d = L^-1 y.
 *   out4 (may be NULL) receives {NegLnLike, mu, SigmaSqr, LnDetPsi}; *info as in pk_nll_batch.
 * When *info != 0 the previously cached factor is kept (the reference keeps its old self.U when
 * np.linalg.cholesky raises, matrixops.py:31). */
int pk_fit(pk_handle *h, const double *theta, const double *p, double lambda, int mode, double *out4,
           int32_t *info);

/* Replaces the per-point loops over predict_normalized (matrixops.py:68-77),
 * predicterr_normalized (matrixops.py:79-95) and expimp / weightedexpimp (krige.py:201-234),
 * i.e. infill_objective_mse / infill_objective_ei (krige.py:236-258) and the plot/snapshot grids.
 *   Xq     : m x k query points in MODEL units (row-major)
 *   theta,p: k values used for the psi vector (the reference uses the CURRENT self.theta/self.pl)
 *   mu, sigma2 : the values the reference would have on self (they can be stale)
 *   y_min  : min(self.y) (krige.py:208);  ei_weight < 0 -> expimp, in [0,1] -> weightedexpimp(w)
 *   var_extra : 0 for predicterr_normalized; Lambda for regression_predicterr_normalized
 *               (matrixops.py:97-110)
 *   outputs (each may be NULL): yhat, s (= RMSE, sqrt|SigmaSqr*(1+extra-psi^T Psi^-1 psi)|), ei : m doubles.
 * Uses the factor cached by the last successful pk_fit. */
int pk_predict_batch(pk_handle *h, int64_t m, const double *Xq, const double *theta, const double *p, double mu,
                     double sigma2, double y_min, double ei_weight, double var_extra, double *yhat, double *s,
                     double *ei);

/* self.Psi / self.U attribute access (matrixops.py:29-32): n x n row-major copies of the correlation
 * matrix of the last pk_fit attempt and of the upper factor U = L^T of the last successful one. */
/* Incremental form of kriging.addPoint (krige.py:135-156: append the sample, updateData, updatePsi -- a
 * full O(n^2 k) Psi rebuild and O(n^3) factorisation in the reference).  With a cached factor (pk_fit) of
 * the CURRENT data and hyper-parameters, appends sample (x_new[k], y_new), both in model units, by the
 * bordered O(n^2) update of L, L^-1, L^-1 1 and L^-1 y, and returns the new (NegLnLike, mu, SigmaSqr,
 * LnDetPsi) in out4.  *appended = 1 on success.  *appended = 0 and nothing modified when there is no
 * cached factor for the current data, when n + 1 would cross a 64-row tile boundary, or when the bordered
 * pivot is not positive (*info = n + 1, dpotrf convention): the caller then takes the reference's route,
 * pk_set_data + pk_fit.  x_new / out4 / info: host or device pointers. */
int pk_append_point(pk_handle *h, const double *x_new, double y_new, double *out4, int32_t *info, int32_t *appended);

int pk_get_psi(pk_handle *h, double *out);
int pk_get_U(pk_handle *h, double *out);

/* Stand-alone batched Psi construction (matrixops.py:28-30 for a whole population): writes the full
 * symmetric n x n matrices of C candidates to `out` (C*n*n doubles).  Diagnostic / benchmark entry
 * point for the Psi kernel; pk_nll_batch does not need it. */
int pk_psi_batch(pk_handle *h, int C, const double *theta, const double *p, const double *lambda, int mode,
                 double *out);

/* ---- Device-resident GA / PSO population: the candidate-population loop of kriging.train (krige.py:373-399), which
 * the reference hands to inspyred.ec.GA / inspyred.swarm.PSO with fittingObjective (krige.py:429-452) as evaluator.
 * The candidate vectors [theta_1..k, p_1..k(, lambda)] stay in device buffers between generations; per generation only
 * the per-candidate scalars come back (pk_opt_evaluate) and only index / uniform arrays go up: the host keeps the draws
 * of Python's Mersenne Twister and the comparisons of fitness values that fix the trajectory under a seed (SURVEY App.
 * B), the device moves and recombines the vectors.  Buffers (`which`): 0 population (GA) / positions x (PSO),
 * 1 offspring, 2 previous positions, 3 personal-best archive.  Index arrays are int32, host or device. */

/* Uploads `n` rows of `dims` doubles into buffer `which`; which == 0 (re)defines the population's shape. */
int pk_opt_set(pk_handle *h, int which, int n, int dims, const double *rows);
/* Reads rows [row0, row0 + nrows) of buffer `which` (final population, the per-generation post-state rows). */
int pk_opt_get(pk_handle *h, int which, int row0, int nrows, double *out);
/* pk_nll_batch on rows [row0, row0 + count) of buffer `which`: the evaluator call of a generation (a rank's slice of
 * it when the population is sharded over GPUs), no candidate upload. */
int pk_opt_evaluate(pk_handle *h, int which, int row0, int count, int mode, double *nll, double *mu, double *sigma2,
                    double *lndet, int32_t *info);
/* One GA generation (inspyred.ec.GA: rank_selection, n_point_crossover with one cut, generational_replacement):
 *   keep (n entries or NULL): row i of the NEW population = row keep[i] of the pool [offspring(nc) | old population];
 *   then, for nsel (even) selected parents: pair i = rows parents[2i], parents[2i+1] of the new population, children
 *   2i, 2i+1 of the offspring buffer = the parents with everything from position cut[i] on exchanged (cut[i] = 0: copied). */
int pk_ga_step(pk_handle *h, int nc, const int32_t *keep, int nsel, const int32_t *parents, const int32_t *cut);
/* One PSO generation (inspyred.swarm.PSO, ring topology): archive[i] = x[i] where improved[i] (NULL: skip), then
 *   x' = clip(x + inertia (x - x_prev) + cognitive r1 (archive - x) + social r2 (archive[nbest] - x), lo, hi),
 * r = n x dims x 2 uniforms (r1, r2 interleaved, the order the host draws them), every operation rounded on its own. */
int pk_pso_step(pk_handle *h, const double *r, const int32_t *nbest, const int32_t *improved, double inertia, double cognitive,
                double social, const double *lo, const double *hi);

/* Block until all work queued on the handle's stream has finished. */
int pk_synchronize(pk_handle *h);

/* The handle's CUDA stream (a cudaStream_t) so callers can time with events on the launching stream.  The library may fan
 * part of a call out to side streams of the handle (the per-group Psi launches of a population, the ragged-first pipeline);
 * they are forked from and joined back to this stream inside the call, so events recorded on it still bracket the work. */
void *pk_stream(pk_handle *h);

/* Number of kernel launches issued by the handle since creation (bench.py's gpu_launches). */
int64_t pk_launch_count(const pk_handle *h);

/* The factorisation declares "not positive definite" (info > 0) when a pivot is <= tol * (its diagonal
 * entry of Psi).  tol = 0 is LAPACK's literal rule `ajj <= 0`; the default is a few ulps, calibrated so
 * that numerically singular matrices (cond(Psi) > 1/eps), where the sign of the computed pivot is pure
 * rounding noise and differs between BLAS builds, fail as often as they do under np.linalg.cholesky
 * (DESIGN.md "failure parity"). */
int pk_set_pivot_tolerance(pk_handle *h, double tol);

/* Tuning / test knob for the blocked Cholesky behind pk_nll_batch and pk_fit (np.linalg.cholesky,
 * matrixops.py:31,41): 0 = automatic (default), 1 = one launch per 64-wide block column with the row
 * blocks of every candidate spread over the grid, 2 = one persistent CTA per candidate (populations), 3 = the
 * persistent kernel with a gang of CTAs per candidate (few candidates: SLSQP stencils, a handful of n = 4096
 * matrices), 4 = TASK mode: the persistent CTAs pull tile tasks (diagonal tile / one 64- or 128-row panel pass of one
 * candidate) from a ticket counter, dependencies through per-candidate completion counters in global memory -- every
 * SM stays busy whatever the population size (128 candidates per GPU of a sharded population, the default 300, a
 * 150-particle swarm at n = 4096) and the results are bit-identical to variant 2; 5 = the INT8-TENSOR-CORE kernel
 * (pk_nll_batch only): the product sums of the factorisation as tcgen05.mma.kind::i8 over 7 base-256 digit planes of L
 * (exact INT32 sums, accumulators in TMEM), tile factorisation and triangular solves in FP64; one persistent CTA per
 * candidate, so a candidate's bits do not depend on the size or order of the call.
 * Automatic: the INT8 kernel for populations that fill their waves of one candidate per SM to 60 % or more (42 % from
 * n = 1473 on; counted on the whole population when the call is a shard, pk_set_shard_hint) at n > 704, task mode
 * otherwise, gangs for pk_fit.  A short last wave of the INT8 kernel (population mod SM count) is either factorised
 * first, under the Psi kernel of the whole waves (small matrices: default population 300 = 4 + 2 x 148), or -- large
 * matrices, at most 64 candidates -- handed to the task kernel ("stragglers": 150 particles at n = 4096); the choice
 * follows from the population size and the candidate's index in it (pk_set_shard_offset), never from the call. */
int pk_set_cholesky_variant(pk_handle *h, int variant);

/* Tuning knobs of task mode (variant 4); none of them changes a bit of the result.  A panel task takes up to
 * `blocks_per_task` 64-row blocks of a block column (0 = automatic: 1-2 for few candidates, the whole column for
 * populations of several waves); block columns with at most `single_below` row blocks left are cut into single blocks
 * (-1 = automatic); `cohort` = number of candidates that walk the task list together (0 = all candidates of the call);
 * candidate m of a cohort runs `m % skew` list entries behind the others, so that the CTAs resident at any moment hold
 * a mix of latency-bound diagonal tiles and DMMA-bound panel tasks (0 = automatic, 1 = none). */
int pk_set_task_options(pk_handle *h, int single_below, int blocks_per_task, int cohort, int skew);

/* Diagnostic (no device needed): the task list of ONE candidate with `nt` block columns as (type, j, row block,
 * blocks) quadruples -- type 0 = diagonal tile j, 1 = panel task of block column j over `blocks` row blocks starting
 * at `row block`.  Returns the number of tasks (writes at most `cap` of them), or -1 if the list is not a valid
 * dependency order (tests/test_library_cpu.py checks the invariants the device-side waits rely on). */
int pk_task_schedule(int nt, int single_below, int blocks_per_task, int32_t *out4, int cap);

/* Tuning / test knob for the fused predictor behind pk_predict_batch (predict_normalized /
 * predicterr_normalized / expimp loops, matrixops.py:68-95, krige.py:201-258): 0 = automatic,
 * 1 = psi tile in shared memory + L^-1 streamed through shared memory (n up to 2944; beyond that the psi tile
 *     lives in a per-CTA slab of a global scratch),
 * 2 = v = L^-1 psi held in registers, psi chunks generated on the fly (n <= 512; falls back otherwise),
 * 3 = split pipeline: psi rows + yhat in one FP64-ALU kernel (through an HBM scratch), |L^-1 psi|^2, s and EI in a
 *     TMA-fed DMMA kernel (any n; what 0 selects for 4096 query points or more).  Row-major tensor grids (plot,
 *     snapshot: krige.py:531-537, 699-708) are detected on the device and take a psi kernel that evaluates the
 *     separated exponent from per-row / per-column tables (bit-identical results),
 * 4 = 3 with that detection switched off. */
int pk_set_predict_variant(pk_handle *h, int variant);

/* The dense contraction |L^-1 psi|^2 of the split predictor (variant 3 / 4; predicterr_normalized, matrixops.py:79-95):
 * 0 = automatic, 1 = FP64 tensor pipe (mma.sync DMMA), 2 = INT8 tensor cores (tcgen05.mma.kind::i8 with TMEM accumulators):
 * psi and L^-1 are split into 8 signed 7-bit digits of a row-scaled fixed-point number, the 36 digit-pair products are
 * exact INT32 sums, the levels are added in FP64 -- an FP64 dot product without rounding inside the sum, at twice the
 * DMMA rate.  Automatic = 2. */
int pk_set_predict_contraction(pk_handle *h, int mode);

/* Multi-GPU sharding (the reference has one Python loop over the whole population,
 * krige.py:436-451, and one over the query points, krige.py:243-257): tells the handle that its next
 * pk_nll_batch / pk_predict_batch calls evaluate a SLICE of a population of `population` candidates / of a query
 * set of `query_points` points.  Automatic kernel selection then uses these sizes instead of the slice's, so every
 * shard computes bit for bit what the unsharded call computes.  0 = no hint. */
int pk_set_shard_hint(pk_handle *h, int64_t population, int64_t query_points);

/* ... and where the slice starts: the next pk_nll_batch calls evaluate candidates `first_candidate`, `first_candidate` + 1,
 * ... of the hinted population (reset to 0 by pk_set_shard_hint).  The kernel a candidate is factorised by can depend on
 * its position in the population -- the few candidates of a short last wave of the INT8 kernel ("stragglers") go through
 * the DMMA task kernel when the matrices are large -- and with the offset a shard makes the same choice per candidate as
 * the unsharded call (reference loop: krige.py:436-451, evaluated in order). */
int pk_set_shard_offset(pk_handle *h, int64_t first_candidate);

/* pk_nll_batch pipelining experiment (the candidate loop of krige.py:429-452 has no such notion: it is
 * strictly sequential).  0 (default): Psi -> Cholesky -> epilogue per resident chunk on one stream.
 * 1: a population larger than the number of resident Cholesky CTAs is split in two parts and the Psi
 * construction of the second part runs on a second stream while the first part is factorised.  Both
 * kernels are bound by the same FP64 pipe, so this only re-orders work: measured 46.0k vs 50.5k evals/s
 * at C3 (profiles/overlap_split_r01.txt) -- kept as a knob, off.  Results are identical either way. */
int pk_set_overlap(pk_handle *h, int enabled);

/* Phase timing for roofline accounting (bench.py): when enabled, CUDA events are recorded on the
 * handle's stream around the phases of every pk_nll_batch / pk_predict_batch call.
 * pk_phase_times() synchronises, ADDS the elapsed milliseconds since the last call into
 * ms[0..3] = {Psi construction, Cholesky (diag + panel kernels), likelihood epilogue, predictor} and
 * resets the accumulation. */
int pk_set_profiling(pk_handle *h, int enabled);
int pk_phase_times(pk_handle *h, double *ms4);

/* Measures this GPU's FP64 roofline denominators in place (MEASURED_PEAKS.json carries none):
 * dense DMMA (mma.sync.m8n8k4.f64) and DFMA throughput in TFLOP/s, each the best of a few ~10 ms
 * launches that fill every SM. */
int pk_fp64_peak(pk_handle *h, double *dmma_tflops, double *dfma_tflops);

/* Diagnostic: evaluates the device building blocks of the Psi kernel on arrays of m doubles --
 * out_pow[i] = |d[i]|^p[i] through the log2-split / table-driven 2^x path that replaces np.power
 * (matrixops.py:28), out_expneg[i] = e^(-d[i]) through the table-driven exponential that replaces
 * np.exp.  Tests use it to bound their error in ulps against a correctly rounded reference. */
int pk_debug_math(pk_handle *h, int64_t m, const double *d, const double *p, double *out_pow, double *out_expneg);

/* Tuning knob: limit the device workspace (bytes) used for candidate chunks (0 = default). */
int pk_set_workspace_limit(pk_handle *h, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* PYKRIGING_B200_H */
