/* libeconpdes_cuda.so -- C ABI of the B200-native finite-scheme hot path of EconPDEs.jl.
 *
 * The reference (EconPDEs.jl v3.0.0) has no FFI: its seams are Julia functions.  This
 * header declares the flat C entry points a Julia `ccall` shim (julia/EconPDEsCUDA.jl,
 * INTEGRATION.md) or the Python ctypes host (econpdes.jl_b200/) binds.  Every entry point
 * cites the reference interface it replaces (file:line relative to the reference root).
 *
 * Conventions
 *  - All arrays are FP64, column-major, laid out (S_1..S_N, F) exactly like the reference's
 *    `guess_array` (src/pdesolve.jl:111); `vec` of it is the unknown vector y.
 *  - Pointers in `econpdes_problem`/`econpdes_result` are HOST pointers owned by the caller;
 *    the library copies host->device at entry and device->host at exit and keeps every field
 *    device-resident in between.  Functions suffixed `_device` take/return device pointers.
 *  - Return value 0 = success, < 0 = error (ECONPDES_E_*); the message is available from
 *    econpdes_last_error() (thread-local).  Non-convergence is NOT an error: it is
 *    `converged = 0` plus an `exit_reason`, like the reference's @warn paths
 *    (src/finiteschemesolve.jl:452-463).
 *  - There is no CPU fallback: every compute entry point fails with ECONPDES_E_NO_DEVICE when
 *    no CUDA device is usable.  econpdes_stencil_colors/pattern/nnz, econpdes_model_* and
 *    econpdes_default_options are host-only and work without a GPU.
 */
#ifndef ECONPDES_CUDA_H
#define ECONPDES_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECONPDES_ABI_VERSION 4
#define ECONPDES_MAX_DIMS 4
#define ECONPDES_MAX_FUNCS 8
#define ECONPDES_MAX_PARAMS 16

/* error codes */
#define ECONPDES_OK 0
#define ECONPDES_E_INVALID (-1)      /* bad argument (the reference's ArgumentError paths) */
#define ECONPDES_E_NAN_GUESS (-2)    /* "G! returns NaN with the initial value", finiteschemesolve.jl:372 */
#define ECONPDES_E_NO_DEVICE (-3)    /* no usable CUDA device: there is no CPU fallback */
#define ECONPDES_E_CUDA (-4)         /* CUDA runtime error */
#define ECONPDES_E_UNSUPPORTED (-5)  /* valid in the reference, not implemented on the device path yet */
#define ECONPDES_E_NCCL (-6)

/* exit_reason (src/finiteschemesolve.jl:373-389, 452-463) */
#define ECONPDES_EXIT_CONVERGED 0
#define ECONPDES_EXIT_ALREADY_ZERO 1      /* :373-375 */
#define ECONPDES_EXIT_MAXITERS 2          /* :461 */
#define ECONPDES_EXIT_BELOW_MIN_DELTA 3   /* :456-459 */
#define ECONPDES_EXIT_INF_STEP_FAILED 4   /* :387 */

/* per-step status (src/finiteschemesolve.jl:418-450) */
#define ECONPDES_STEP_ACCEPTED 0
#define ECONPDES_STEP_REJECTED_GROWTH 1   /* :425 */
#define ECONPDES_STEP_REJECTED_INNER 2    /* :439 */
#define ECONPDES_STEP_REJECTED_NAN 3      /* :443 "(model returned NaN)" */
#define ECONPDES_STEP_REJECTED_SINGULAR 4 /* :444 "(singular Jacobian)" */

/* eps_policy */
#define ECONPDES_EPS_FINITEDIFF_COLOR_NORM 0 /* FiniteDiff.jl: one step per color, sqrt(eps)*max(1,sqrt(||x.*mask||)) */
#define ECONPDES_EPS_PER_COLUMN 1            /* sqrt(eps)*max(1,|x_j|): robust option, not reference parity */

/* linear_solver */
#define ECONPDES_LIN_AUTO 0      /* 1-D: (block-)tridiagonal cyclic reduction; 2-D/3-D: Krylov + multigrid (one unknown) or
                                    block line relaxation (several); 4-D: Krylov + line relaxation */
#define ECONPDES_LIN_GMRES_MG 1
#define ECONPDES_LIN_BICGSTAB_MG 2

typedef struct econpdes_problem {
  int32_t model_id;                 /* econpdes_model_id(name) */
  int32_t ndims;                    /* N: number of state variables (StateGrid, src/utils/stategrid.jl:8-29) */
  int32_t nfuncs;                   /* F: number of unknown functions */
  int32_t nparams;
  int64_t size[ECONPDES_MAX_DIMS];  /* S_d >= 2 (stategrid.jl:16) */
  const double* grid[ECONPDES_MAX_DIMS]; /* coordinate vector per dimension, may be non-uniform */
  const double* params;             /* model struct fields, order given by econpdes_model_info */
  const double* guess;              /* (S.., F) column-major                         src/pdesolve.jl:111 */
  const double* bc;                 /* (S.., F) or NULL = zeros (reflecting)         src/pdesolve.jl:312-335 */
  const uint8_t* is_algebraic;      /* (S.., F) or NULL = all false                  src/pdesolve.jl:109,112 */
  const double* lower_bound;        /* (S.., F) or NULL = -Inf                       src/pdesolve.jl:114 */
  const double* upper_bound;        /* (S.., F) or NULL = +Inf                       src/pdesolve.jl:115 */
  const double* delta_vec;          /* (S.., F) signed per-equation steps or NULL = scalar options.delta
                                       src/pdesolve.jl:376-394, finiteschemesolve.jl:393-399 */
} econpdes_problem;

/* keyword arguments of finiteschemesolve (src/finiteschemesolve.jl:343; docs/src/solver.md:286-298) */
typedef struct econpdes_options {
  double abstol;               /* sqrt(eps) */
  double inner_abstol;         /* sqrt(eps) */
  double delta;                /* 1.0; +-Inf = one-shot solve without continuation */
  double scale;                /* 10.0 */
  double min_delta;            /* 1e-9 */
  double max_delta;            /* Inf */
  double max_residual_growth;  /* 10.0 */
  int32_t maxiters;            /* 100 */
  int32_t inner_maxiters;      /* 10 */
  int32_t eps_policy;          /* ECONPDES_EPS_* */
  int32_t linear_solver;       /* ECONPDES_LIN_* */
  double lin_rtol;             /* relative residual target of the Krylov solve (1e-12) */
  int32_t lin_maxiters;        /* Krylov iteration cap per solve (2000) */
  int32_t lin_restart;         /* GMRES restart length (40) */
  int32_t check_monotonicity;  /* 0; `check_monotonicity` of pdesolve (src/pdesolve.jl:82-89): also count the
                                  same-unknown spatial off-diagonal entries of the first Jacobian above the tolerance */
  double monotonicity_tol;     /* 1e-6 (`monotonicity_tol`) */
} econpdes_options;

typedef struct econpdes_log_row {
  int32_t iter;
  int32_t status;        /* ECONPDES_STEP_* */
  int32_t newton_iters;
  int32_t linear_iters;  /* Krylov iterations summed over the Newton iterations of this step */
  double delta;          /* stepper.Δ printed in the TimeStep column (finiteschemesolve.jl:421) */
  double residual;       /* ||G||_2/n at the accepted point.  Rejected steps: NaN, except that the generic engine
                            stores residual/old_residual on ECONPDES_STEP_REJECTED_GROWTH rows -- the growth factor
                            the reference prints there (finiteschemesolve.jl:433-435; +Inf prints as "∞") */
} econpdes_log_row;

typedef struct econpdes_result {
  double* solution;        /* out: (S.., F), caller-allocated */
  double* saved;           /* out: (S.., NSAVED) or NULL     (_fill_saved!, src/pdesolve.jl:424-433) */
  econpdes_log_row* log;   /* out: caller-allocated, log_capacity rows (or NULL) */
  int32_t log_capacity;
  int32_t log_len;
  double residual_norm;    /* ||G(y)||_2 / n */
  int32_t converged;       /* residual_norm <= abstol (src/EconPDEs.jl:44-48) */
  int32_t exit_reason;     /* ECONPDES_EXIT_* */
  int32_t n_outer;
  int32_t n_newton;
  int64_t n_linear_iters;
  int64_t kernel_launches; /* kernels of this library launched during the call */
  double device_ms;        /* CUDA-event time of the solve proper (between the H2D and D2H copies) */
  /* Jacobian diagnostics of src/monotonicity.jl, reductions over the first assembled stencil-format Jacobian:
   * sign_flags bit k (k < F): the residual-Jacobian diagonal of unknown k is opposite in sign to its
   * pseudo-time step at EVERY checked grid point (_check_sign_convention!, :57-98; runs by default);
   * bit ECONPDES_SIGN_WHOLE_SYSTEM: uniform positive step and every unknown flagged (the single
   * "returns RHS - rho v" diagnosis).  nonmonotone_entries / nonmonotone_max: number and largest value of
   * same-unknown spatial off-diagonal entries above monotonicity_tol (_check_monotonicity!, :100-133; only
   * with options.check_monotonicity). */
  int32_t sign_flags;
  int32_t nonmonotone_entries;
  double nonmonotone_max;
  /* Quality of the linear solves behind the Newton iterations (the reference's `J \ R` is a sparse LU, exact to
   * round-off; the 2-D / 3-D path is a Krylov solve driven to options.lin_rtol, which accepts a stagnated solve
   * at relative residual <= 1e-8): number of linear solves, how many of them ended above 1e-11, and the
   * largest final relative residual.  1-D (block-)tridiagonal solves report 0 / 0. */
  int32_t n_linear_solves;
  int32_t n_linear_loose;
  double lin_max_relres;
} econpdes_result;
#define ECONPDES_SIGN_WHOLE_SYSTEM (1 << 30)

int econpdes_abi_version(void);
const char* econpdes_last_error(void);

/* --- model registry: replaces dispatch on typeof(pde) (callable structs, e.g.
 *     examples/asset_pricing/bansal_yaron.jl:22-32) --- */
int econpdes_model_count(void);
int econpdes_model_id(const char* name);            /* < 0 if unknown */
const char* econpdes_model_name(int model_id);
int econpdes_model_info(int model_id, int32_t* ndims, int32_t* nfuncs, int32_t* nparams, int32_t* nsaved,
                        int32_t* box_stencil);

/* --- sparsity pattern + coloring, host-only, bit-exact with the reference --- */
/* stencil_colors(S, F), src/utils/coloring.jl:52-70; colors_out[l*F], 1-based */
int econpdes_stencil_colors(int ndims, const int64_t* size, int nfuncs, int64_t* colors_out);
/* nnz of local_stencil_jacobian(S, F) = F^2 prod(3 S_d - 2), src/utils/coloring.jl:18-35 */
int64_t econpdes_stencil_nnz(int ndims, const int64_t* size, int nfuncs);
/* CSC pattern of local_stencil_jacobian: colptr[l*F+1], rowval[nnz], 1-based, rows sorted */
int econpdes_stencil_pattern(int ndims, const int64_t* size, int nfuncs, int64_t* colptr, int64_t* rowval);

void econpdes_default_options(econpdes_options* opts);

/* --- one-shot stationary solve: replaces finiteschemesolve behind pdesolve
 *     (src/pdesolve.jl:136-154 -> src/finiteschemesolve.jl:343-465) --- */
int econpdes_solve_stationary(const econpdes_problem* problem, const econpdes_options* options,
                              econpdes_result* result);

/* --- time-dependent solve: replaces _solve_backward behind pdesolve(pde, grid, guess, taus)
 *     (src/pdesolve.jl:158-224).  The guess is the terminal value at taus[ntaus-1]; going backward,
 *     every earlier time is one implicit_timestep (Newton solve, no step controller) with
 *     delta = taus[i] - taus[i-1], options->inner_maxiters Newton iterations at most and
 *     options->inner_abstol as its ||R||_inf tolerance.  taus must be finite and strictly increasing
 *     (_check_time_grid, :241-251).  Outputs, caller-allocated: solution (S.., F, ntaus) and, if not
 *     NULL, saved (S.., NSAVED, ntaus), column-major with time slowest; residual_norms[ntaus]
 *     (entry 0 stays 0, entry i is the Newton residual of the step that produced time i-1, as in the
 *     reference); newton_iters[ntaus] (may be NULL).  Functors of the reference's 3-argument form
 *     pde(state, u, t) (src/pdesolve.jl:253-260) read the current time from the parameter slot the model
 *     declares (econpdes_model_info: time_slot); the driver stores taus[i] there before the step from
 *     time i to i-1, as pde_onestep does (src/pdesolve.jl:195).  Two-argument functors ignore it. --- */
int econpdes_solve_backward(const econpdes_problem* problem, const econpdes_options* options, const double* taus,
                            int32_t ntaus, double* solution, double* saved, double* residual_norms,
                            int32_t* newton_iters, double* device_ms, int64_t* kernel_launches);

/* --- handle API for repeated residual / Jacobian evaluations --- */
typedef struct econpdes_handle econpdes_handle;
int econpdes_create(const econpdes_problem* problem, econpdes_handle** out);
void econpdes_destroy(econpdes_handle* h);
/* G!(ydot, y): pde!, src/pdesolve.jl:436-462 */
int econpdes_residual(econpdes_handle* h, const double* y, double* ydot);
/* second NamedTuple of the model on the whole grid: _fill_saved!, src/pdesolve.jl:424-433 */
int econpdes_saved(econpdes_handle* h, const double* y, double* saved);
/* finite_difference_jacobian!(J, G_helper!, y, fdcache) as called at finiteschemesolve.jl:184:
 * nzval of the CSC pattern above for R(y) = G(y) - (!alg)(ypost - y)/delta  (ypost NULL: R = G).
 * eps_out (may be NULL) receives the per-color steps, length F*prod(min(3,S_d)). */
int econpdes_jacobian_csc(econpdes_handle* h, const double* y, const double* ypost, double delta,
                          int32_t eps_policy, double* nzval, double* eps_out);
/* one linear solve J x = b with the solver the Newton loop uses (J from the last
 * econpdes_jacobian_csc call); replaces Julia's `J \ b` (UMFPACK) reached through NLsolve */
int econpdes_linear_solve(econpdes_handle* h, const econpdes_options* options, const double* b, double* x,
                          int32_t* iters_out, double* relres_out);
/* implicit_timestep(G!, ypost, delta; ...), src/finiteschemesolve.jl:153-230: one Newton solve of
 * the implicit-Euler step.  y_out (n) receives the Newton result, nl_residual ||R||_2/n. */
int econpdes_implicit_timestep(econpdes_handle* h, const econpdes_options* options, const double* ypost,
                               double delta, int32_t maxiters, double abstol, double* y_out,
                               double* nl_residual, int32_t* newton_iters);
/* Measurement hook for bench.py: average device time (ms per launch, CUDA events on the launching
 * stream, L2 flushed between launches) of the grid-wide kernels at y on the handle's problem:
 * ms_out[0] residual sweep G(y), [1] implicit residual R(y) + norms, [2] colored-FD Jacobian
 * (color norms + stencil planes), [3] stencil SpMV, [4] one multigrid V-cycle, [5] multigrid set-up
 * ([3..5] are 0 for 1-D problems). */
int econpdes_bench_kernels(econpdes_handle* h, const double* y, int reps, double* ms_out);
/* finiteschemesolve on the handle's problem starting from its guess */
int econpdes_solve(econpdes_handle* h, const econpdes_options* options, econpdes_result* result);

/* --- batched sweep of independent 1-D, single-unknown problems (config C4): every problem
 *     is its own pdesolve call in the reference; here one thread block owns one problem and
 *     runs its whole pseudo-transient loop on chip. --- */
typedef struct econpdes_batch econpdes_batch;
/* params[B][nparams], grids[B][S], guess[B][S] (host, row-major over the batch) */
int econpdes_batch_create(int model_id, int64_t nproblems, int64_t npoints, econpdes_batch** out);
void econpdes_batch_destroy(econpdes_batch* b);
int econpdes_batch_upload(econpdes_batch* b, const double* params, const double* grids, const double* guess);
/* keep up to log_capacity iteration-log rows per problem (0 = no log, the default) */
int econpdes_batch_set_log_capacity(econpdes_batch* b, int32_t log_capacity);
int econpdes_batch_solve(econpdes_batch* b, const econpdes_options* options, double* device_ms);
/* solution[B][S], residual_norm[B], exit_reason[B], n_outer[B], n_newton[B]; any may be NULL.
 * log rows (optional): log[B][log_capacity], log_len[B]. */
int econpdes_batch_download(econpdes_batch* b, double* solution, double* residual_norm, int32_t* exit_reason,
                            int32_t* n_outer, int32_t* n_newton, econpdes_log_row* log, int32_t log_capacity,
                            int32_t* log_len);
int64_t econpdes_batch_kernel_launches(const econpdes_batch* b);
/* upload + solve + download in one call (the end-to-end path a pdesolve loop replaces).  The batch is cut
 * into slices whose host->device copies and solves are pipelined over three streams.
 * `solution` page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory): ZERO-COPY results -- every
 * thread block stores its solution straight into the caller's buffer over PCIe as it finishes and the
 * per-problem scalars into the library's own page-locked arrays, so there is no device->host copy at all.
 * `solution` pageable (or ECONPDES_DIRECT_OUTPUT=0): device buffers + one device->host copy per slice, which
 * overlap the solves only for page-locked buffers.  Same results either way.
 * econpdes_last_batched_output_mode(): 1 if this host thread's last one-call solve was zero-copy, 0 if staged,
 * -1 if there was none. */
int econpdes_solve_batched(int model_id, int64_t nproblems, int64_t npoints, const double* params,
                           const double* grids, const double* guess, const econpdes_options* options,
                           double* solution, double* residual_norm, int32_t* exit_reason, int32_t* n_outer,
                           int32_t* n_newton);

/* Compact inputs of a sweep.  A calibration loop builds every grid from a few numbers
 * (initialize_stategrid, examples/asset_pricing/campbell_cochrane.jl:40-48:
 * vcat(range(-300, shigh[2], length = sn - div(sn, 10))[1:end-1], shigh[2:end]) and starts every unknown
 * at one value (guess = ones): shipping B x S grids and guesses over PCIe costs more than solving them.
 *   grids != NULL:  explicit grids[B][S] (n_range, range_lo, points ignored)
 *   grids == NULL:  the first n_range points of problem b are
 *                       range(range_lo[b], points[b][0], length = n_range + 1)[1:n_range]
 *                   evaluated like Julia's `range` (correctly rounded a + i (b - a)/(n - 1)), built on the device,
 *                   followed by the explicit points[b][0 .. S - n_range); 1 <= n_range < S
 *   guess != NULL:  explicit guess[B][S];  guess == NULL: every unknown starts at guess_fill */
typedef struct econpdes_batch_inputs {
  const double* params;    /* [B][nparams] */
  const double* grids;     /* [B][S] or NULL */
  int32_t n_range;
  const double* range_lo;  /* [B] */
  const double* points;    /* [B][S - n_range] */
  const double* guess;     /* [B][S] or NULL */
  double guess_fill;
} econpdes_batch_inputs;
int econpdes_last_batched_output_mode(void);

/* econpdes_batch_upload / econpdes_solve_batched with compact inputs */
int econpdes_batch_upload_inputs(econpdes_batch* b, const econpdes_batch_inputs* inputs);
int econpdes_solve_batched_inputs(int model_id, int64_t nproblems, int64_t npoints, const econpdes_batch_inputs* inputs,
                                  const econpdes_options* options, double* solution, double* residual_norm,
                                  int32_t* exit_reason, int32_t* n_outer, int32_t* n_newton);

/* econpdes_solve_batched keeps its device workspace (per host thread) between calls of the same
 * shape; this frees it. */
void econpdes_release_workspace(void);

/* Tuning knob of the batched engine: grid points owned by one thread (1, 2, 3 or 4; 0 = choose from
 * the problem size, the default).  Results agree to round-off of the tridiagonal solve for every
 * setting (the reference's sparse LU is replaced either way); exposed so tests can cover each variant.
 * Returns the previous setting. */
int econpdes_set_batch_points_per_thread(int points);

/* Device self-test of the branch-free reciprocal and quotient the kernels use in place of IEEE division
 * (stencil.cuh: ep_rcp, mdiv).  Draws `n` operand pairs (a, b) as uniformly random 64-bit patterns (every
 * exponent, subnormals, Inf and NaN, extreme mantissas) from `seed` and counts results whose bits differ
 * from the compiler's 1.0 / b and a / b.  All three counts must be 0:
 *   mismatch[0]  reciprocals, for |b| < 2^1022 and b = 0, Inf, NaN
 *   mismatch[1]  quotients a / b through RN(1/b), for |b| in [2^-1000, 2^1000] and a = 0, Inf, NaN or |a| and
 *                |a / b| at least 2^-900 (nearer the underflow threshold a subnormal result may round twice)
 *   mismatch[2]  reciprocals with |b| >= 2^1022 (subnormal results) off by more than one unit in the last place */
int econpdes_selftest_reciprocal(uint64_t seed, int64_t n, int64_t mismatch[3]);

/* --- device / multi-GPU plumbing --- */
int econpdes_device_count(void);
int econpdes_set_device(int device);
int64_t econpdes_total_kernel_launches(void);
/* Debug facility (stands in for a memcheck tool): when the process runs with ECONPDES_GUARD=1 every device buffer
 * of the library carries a 256-byte canary region on either side; returns the number of corrupted canary bytes
 * found so far (live buffers are checked now, released ones were checked at release), i.e. 0 unless some kernel
 * wrote out of bounds.  -1 when the facility is off. */
int64_t econpdes_guard_violations(void);

/* --- slab-decomposed solve of ONE large grid across the GPUs of a node (config C5).
 *     One process per GPU.  The slowest dimension (N-1; slabs are contiguous in the column-major
 *     layout of src/pdesolve.jl:111) is cut into contiguous ranges; each rank owns the planes
 *     [slab_first, slab_first + slab_count) and exchanges one ghost plane per side over
 *     NVLink (NCCL send/recv) before every residual / Jacobian / SpMV sweep; norms, Krylov dot
 *     products and FiniteDiff's per-color sums are all-gathered and folded in rank order, so every
 *     rank takes the same accept/reject decisions as the single-device solve.
 *     The reference has no counterpart (it is single-threaded): this replaces the same
 *     finiteschemesolve call as econpdes_solve_stationary. --- */
#define ECONPDES_UNIQUE_ID_BYTES 128
typedef struct econpdes_dist econpdes_dist;
/* rank 0 creates the id; the host broadcasts it to the other ranks (torch.distributed, MPI, a pipe) */
int econpdes_dist_unique_id(uint8_t* id /* [ECONPDES_UNIQUE_ID_BYTES] */);
/* collective over all ranks; binds to the calling thread's current CUDA device */
int econpdes_dist_create(int rank, int world, const uint8_t* id, econpdes_dist** out);
void econpdes_dist_destroy(econpdes_dist* d);
int64_t econpdes_dist_collectives(const econpdes_dist* d); /* NCCL calls issued so far */
/* halo exchanges done by the library's own NVLink kernel (k_halo_p2p: peer-memory stores + system-scope
 * flags through CUDA IPC mailboxes) instead of NCCL send/recv; 0 if IPC is unavailable or
 * ECONPDES_P2P_HALO=0 */
int64_t econpdes_dist_p2p_exchanges(const econpdes_dist* d);
/* balanced partition of dimension N-1 in units of two planes (slabs start on even indices so the
 * multigrid hierarchy coarsens inside a slab); host-only */
int econpdes_slab_range(int64_t size_last, int rank, int world, int64_t* first, int64_t* count);
/* problem->size / problem->grid describe the GLOBAL grid; guess, bc, is_algebraic and the bounds are
 * this rank's slab (S_1..S_{N-1}, slab_count), and so is result->solution.  Scalars of the result
 * (residual_norm, converged, log, counts) are identical on every rank. */
int econpdes_solve_stationary_slab(econpdes_dist* d, const econpdes_problem* problem, int64_t slab_first,
                                   int64_t slab_count, const econpdes_options* options, econpdes_result* result);

#ifdef __cplusplus
}
#endif
#endif /* ECONPDES_CUDA_H */
