// Fused batched engine for sweeps of independent 1-D, single-unknown problems (config C4:
// 16 384 Campbell-Cochrane calibrations of 998 points each).
//
// In the reference every problem is its own `pdesolve` call, i.e. its own pseudo-transient
// loop (src/finiteschemesolve.jl:406-451), Newton iterations (NLsolve), colored FD
// Jacobians (FiniteDiff) and sparse LU.  Here ONE THREAD BLOCK OWNS ONE PROBLEM and runs
// that whole loop on chip: the unknown lives one value per thread in registers, neighbours
// are exchanged through shared memory, the three-color Jacobian is built row-wise by
// `jacobian_point`, the tridiagonal Newton system is solved by shared-memory PCR
// (tridiag.cuh) and the time-step controller (finiteschemesolve.jl:236-275) is evaluated
// redundantly by every thread from block-reduced norms, so control flow is block-uniform.
// HBM traffic is the grid + guess in and the solution out; nothing else leaves the SM.
#pragma once
#include <type_traits>

#include "../../include/econpdes_cuda.h"
#include "kernels.cuh"
#include "sweep.cuh"
#include "tridiag.cuh"

namespace econpdes {

struct BatchArgs {
  long long nproblems;     // one past the last problem of this launch
  long long problem0 = 0;  // first problem of this launch (block b solves problem problem0 + b)
  int S;
  int nparams;
  const double* params;  // [B][nparams]
  const double* grids;   // [B][S]
  const double* guess;   // [B][S]
  double* solution;      // [B][S]
  double* residual_norm; // [B]
  int* exit_reason;      // [B]
  int* n_outer;          // [B]
  int* n_newton;         // [B]
  int* error;            // [B] 0 or ECONPDES_E_NAN_GUESS
  econpdes_log_row* log; // [B][log_capacity] or nullptr
  int log_capacity;
  int* log_len;          // [B]
  // compact inputs (econpdes_batch_inputs), expanded by the block itself while it stages its problem: a null `grids`
  // means the first n_range points of problem b are range(range_lo[b], points[b][0], length = n_range + 1)[1:n_range]
  // (range_point: Julia's twice-precision `range`) followed by the explicit points[b][0 .. S - n_range); a null
  // `guess` means every unknown starts at guess_fill
  const double* range_lo = nullptr;   // [B]
  const double* points = nullptr;     // [B][S - n_range]
  int n_range = 0;
  double guess_fill = 0.0;
  int* flags = nullptr;  // [B] or nullptr: bit 0 = some Newton iteration met a singular Jacobian (a vanishing pivot)
  int* flagged_count = nullptr;   // page-locked host counter of such problems (system-scope atomic; rare) or nullptr
  econpdes_options opt;
};

// (sum_0..sum_{NS-1}) and (max_0..max_{NM-1}) over the block, fixed order, result in all threads.
template <int NS, int NM>
__device__ __forceinline__ void block_reduce_mixed(double* sums, double* maxs, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) sums[k] = warp_sum(sums[k]);
#pragma unroll
  for (int k = 0; k < NM; ++k) maxs[k] = warp_max(maxs[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) scratch[k * 32 + warp] = sums[k];
#pragma unroll
    for (int k = 0; k < NM; ++k) scratch[(NS + k) * 32 + warp] = maxs[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NS; ++k) sums[k] = warp_sum(lane < nw ? scratch[k * 32 + lane] : 0.0);
#pragma unroll
  for (int k = 0; k < NM; ++k) maxs[k] = warp_max(lane < nw ? scratch[(NS + k) * 32 + lane] : 0.0);
}

// SERTimeStepController, src/finiteschemesolve.jl:236-275
struct SerController {
  double delta, scale, max_delta, acceleration, delta_cap;
  __host__ __device__ void init(double d, double sc, double mx) {
    delta = d; scale = sc; max_delta = mx; acceleration = 1.0; delta_cap = mx;
  }
  __host__ __device__ void accept(double old_r, double new_r) {
    if (new_r <= old_r) acceleration *= scale; else acceleration = 1.0;
    const double proposed = delta * acceleration * old_r / new_r;
    const double limit = delta_cap < max_delta ? delta_cap : max_delta;
    if (proposed > limit) { delta = limit; acceleration = 1.0; } else { delta = proposed; }
    const double c = delta_cap * scale;
    delta_cap = c < max_delta ? c : max_delta;
  }
  __host__ __device__ void reject() {
    acceleration = 1.0;
    delta_cap = delta / 2;
    delta /= 10;
  }
};

// Jacobian diagnostics of src/monotonicity.jl for ONE 1-D single-unknown problem (the single-`pdesolve` route
// through the fused kernel): the colored-FD Jacobian of G at y, one thread per row, and the two reductions of
// k_jacobian_diagnostics (engine.cu): counts[0] rows whose diagonal dG_i/dy_i exceeds tol in magnitude, counts[1]
// those of them whose sign is opposite to the step's, counts[2] off-diagonal entries above tol, maxbits the largest.
template <class M>
__global__ void __launch_bounds__(1024) k_diag_1d(int S, ParamPack prmpack, const double* __restrict__ x,
                                                  const double* __restrict__ yin, double delta, double tol, int signs,
                                                  int stencils, int* __restrict__ counts,
                                                  unsigned long long* __restrict__ maxbits) {
  static_assert(M::N == 1 && M::F == 1, "1-D single-unknown models");
  __shared__ double red[8 * 32];
  __shared__ double ysh[1024 + 2];
  const int t = threadIdx.x;
  const bool act = t < S;
  if (t == 0) { ysh[0] = 0.0; ysh[S + 1] = 0.0; }
  if (act) ysh[t + 1] = yin[t];
  __syncthreads();
  double s3[3] = {0.0, 0.0, 0.0}, m0[1];
  const double q = act ? ysh[t + 1] * ysh[t + 1] : 0.0;
  s3[t % 3] = q;
  block_reduce_mixed<3, 0>(s3, m0, red);
  double eps_table[3], reps_table[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    eps_table[c] = color_epsilon_from_sumsq(s3[c]);
    reps_table[c] = 1.0 / eps_table[c];
  }
  if (!act) return;
  GridView<1> gv;
  gv.x[0] = x; gv.S[0] = S; gv.stride[0] = 1; gv.l = S;
  long long idx[1] = {t};
  PointGeom<1, 1> g;
  double st[1];
  load_geom(gv, idx, g, st);
  typename M::Pre pre;
  M::prepare(prmpack.v, st, pre);
  const double bl[1][1] = {{0.0}}, bh[1][1] = {{0.0}};
  double w[1][3] = {{ysh[t], ysh[t + 1], ysh[t + 2]}};
  RowTerms rt;
  rt.implicit = false; rt.has_bounds = false; rt.ypost = 0.0; rt.delta = delta; rt.lb = 0.0; rt.ub = 0.0; rt.is_alg = false;
  double r0[1], va = 0.0, vb = 0.0, vc = 0.0;
  jacobian_point<M, false, false, 0>(pre, g, idx, gv, w, bl, bh, &rt, eps_table, 0, r0,
                                     [&](int k, int, int, double v) {
                                       if (k == 0) vb = v; else if (k == 1) va = v; else vc = v;
                                     }, reps_table);
  if (signs && isfinite(delta) && fabs(vb) > tol) {
    atomicAdd(&counts[0], 1);
    const bool same = (vb > 0 && delta > 0) || (vb < 0 && delta < 0);
    if (!same) atomicAdd(&counts[1], 1);
  }
  if (stencils) {
    if (va > tol) { atomicAdd(&counts[2], 1); atomicMax(maxbits, static_cast<unsigned long long>(__double_as_longlong(va))); }
    if (vc > tol) { atomicAdd(&counts[2], 1); atomicMax(maxbits, static_cast<unsigned long long>(__double_as_longlong(vc))); }
  }
}

// 2^-200 <= |v| < 2^200 (and neither Inf nor NaN): the "tame geometry" rule of sweep.cuh / engine.cu, on the
// exponent field
__device__ __forceinline__ bool moderate_magnitude(double v) {
  const int ex = (__double2hiint(v) >> 20) & 0x7ff;
  return ex >= 1023 - 200 && ex < 1023 + 200;
}

// The kernel is written as a set of lambdas over its register state; whether the compiler inlines the large ones
// depends on the size of the whole translation unit (an outlined Newton lambda costs a 1.7 KB stack frame and half the
// speed -- seen when a second instantiation was added), so inlining is not left to the heuristic.
#define EP_LAMBDA_INLINE __attribute__((always_inline))

// threads of a block that covers the engine's maximum of 1024 points with P points each (whole warps)
constexpr int batched_max_threads(int P) { return ((1024 + P - 1) / P + 31) / 32 * 32; }

// P = grid points per thread (consecutive).  With P > 1 the per-thread bookkeeping (block
// reductions, barriers, controller, color steps) is shared by P points, the P evaluations are
// independent instruction streams, and the tridiagonal solve only runs PCR on the T interface rows
// (block_chunk_solve, tridiag.cuh).
// PER_COL: eps_policy == ECONPDES_EPS_PER_COLUMN, fixed at compile time so that the Newton iteration has no
// data-independent branches: together with the branch-free reciprocals (ep_rcp, mdiv) the evaluations of a
// thread's P points form one basic block the scheduler can interleave.
// COMPACT: the block expands compact inputs (BatchArgs::range_lo / points / guess_fill) itself; a separate
// instantiation so that the explicit-input kernel keeps its register allocation (the extra prologue cost it 1.5 %)
template <class M, int P, bool PER_COL, bool COMPACT = false>
__global__ void __launch_bounds__(batched_max_threads(P), 1) k_batched_pdesolve_1d(BatchArgs args) {
  static_assert(M::N == 1 && M::F == 1, "the fused batched engine handles 1-D single-unknown models");
  static_assert(M::NPARAMS + M::NDERIVED <= kMaxBatchParams, "parameter staging area of the batched kernel too small");
  extern __shared__ double smem[];
  const int T = blockDim.x;
  const int NP = T * P;
  double* ysh = smem;                 // [NP + 2]: y with one halo slot each side
  double* red = ysh + (NP + 2);       // [8 * 32] reduction scratch
  double* prm = red + 8 * 32;         // [16]
  double* pcr = prm + 16;             // [8 * T] PCR buffers; holds the grid coordinates during set-up
  double* xch = pcr + 8 * T;          // [4 * T] chunk interface exchange
  const long long b = args.problem0 + blockIdx.x;
  const int t = threadIdx.x;
  const int S = args.S;
  const int i0 = t * P;               // first point of this thread
  const econpdes_options& opt = args.opt;

  if (t < args.nparams) prm[t] = args.params[b * args.nparams + t];
  if (t == 0) { ysh[0] = 0.0; ysh[NP + 1] = 0.0; }
  // coalesced loads of the grid (staged in the PCR area, NP <= 8 T) and the guess
  double* xs = pcr;
  if (!COMPACT || args.grids != nullptr) {
    for (int k = t; k < NP; k += T) xs[k] = k < S ? args.grids[b * S + k] : 0.0;
  } else {
    const double* pts = args.points + b * (S - args.n_range);
    const double lo = args.range_lo[b], hi = pts[0];
    for (int k = t; k < NP; k += T)
      xs[k] = k < S ? (k < args.n_range ? range_point(lo, hi, args.n_range + 1, k) : pts[k - args.n_range]) : 0.0;
  }
  if (!COMPACT || args.guess != nullptr) {
    for (int k = t; k < NP; k += T) ysh[k + 1] = k < S ? args.guess[b * S + k] : 0.0;
  } else {
    for (int k = t; k < NP; k += T) ysh[k + 1] = k < S ? args.guess_fill : 0.0;
  }
  __syncthreads();
  if (M::NDERIVED > 0) {   // parameter-only constants of the model, once per problem
    if (t == 0) M::derive(prm);
    __syncthreads();
  }
  GridView<1> gv;
  gv.x[0] = xs; gv.S[0] = S; gv.stride[0] = 1; gv.l = S;
  bool act[P];
  long long idx[P][1];
  PointGeom<1, 1> g[P];   // reciprocals of the spacings once per problem, reused by every stencil evaluation
  typename M::Pre pre[P]; // state-only part of the model, once per problem
  double ypost[P];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    act[j] = i0 + j < S;
    idx[j][0] = act[j] ? i0 + j : 0;
    double st[1];
    load_geom(gv, idx[j], g[j], st);
    M::prepare(prm, st, pre[j]);
    ypost[j] = ysh[i0 + j + 1];
  }
  const double bl[1][1] = {{0.0}}, bh[1][1] = {{0.0}};
  // "Tame" problems take the stencil quotients without their NaN-select (mdiv_tame, stencil.cuh; a quarter of the
  // kernel's instructions were those selects): the spacings and their reciprocals must be finite and of moderate
  // magnitude -- checked here, once per problem -- and every unknown below 2^199 -- checked with the other
  // predicates of each Newton iterate (kUntame below).  The boundary derivative values of this engine are zero, so
  // the tame form is TAME = 2.  Same bits either way; a block that meets an untame iterate re-evaluates it with the
  // general form and stays on it while the iterate is wild.
  bool geom_ok = true;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const bool ok = moderate_magnitude(g[j].dm[0]) && moderate_magnitude(g[j].dp[0]) && moderate_magnitude(g[j].d[0]) &&
                    moderate_magnitude(g[j].rdp[0]) && moderate_magnitude(g[j].rdm[0]) &&
                    moderate_magnitude(g[j].rdpd[0]) && moderate_magnitude(g[j].rdmd[0]) &&
                    moderate_magnitude(g[j].rdpm[0]);
    if (act[j] && !ok) geom_ok = false;
  }
  const bool geom_tame = __syncthreads_and(geom_ok) != 0;   // also: the grid staging area is free again

  // residual G at this thread's points from the y currently published in ysh
  auto evalG = [&](double* G, auto tame_tag) EP_LAMBDA_INLINE {
    constexpr int TAME = decltype(tame_tag)::value;
    double yv[P + 2];
#pragma unroll
    for (int k = 0; k < P + 2; ++k) yv[k] = ysh[i0 + k];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const double w[1][3] = {{yv[j], yv[j + 1], yv[j + 2]}};
      double o[1];
      eval_point<M, false, TAME>(pre[j], g[j], w, bl, bh, o, nullptr);
      G[j] = act[j] ? o[0] : 0.0;
    }
  };
  using General = std::integral_constant<int, 0>;
  // tame level of an iteration: 2 = stencil quotients; 3 = also the model's own quotients by the unknown (M::eval_tame)
  // and the implicit-Euler quotient by the step, for models that offer it -- then an unknown is "tame" only if it is
  // also positive and at least 2^-200, and the step must be of moderate magnitude (not Inf)
  constexpr int kTameLevel = M::HAS_TAME_EVAL ? 3 : 2;
  using Tame = std::integral_constant<int, kTameLevel>;
  auto tame_value = [](double v) EP_LAMBDA_INLINE -> bool {
    if (kTameLevel < 3) return is_tame(v);
    const int hi = __double2hiint(v);   // sign bit clear, exponent in [1023 - 200, 1023 + 199)
    return hi >= ((1023 - 200) << 20) && hi < ((1023 + 199) << 20);
  };
  auto publish = [&](const double* v) EP_LAMBDA_INLINE {
    __syncthreads();  // everyone finished reading the previous y
#pragma unroll
    for (int j = 0; j < P; ++j) ysh[i0 + j + 1] = v[j];
    __syncthreads();
  };

  double Gpost[P];
  evalG(Gpost, General{});   // once per problem: the guess has not been vetted yet
  const double nd = static_cast<double>(S);

  // ||G||_2 / n with NaN -> NaN, as norm(ydot) / length(ydot) (:371, :415)
  auto normG = [&](const double* Gv) EP_LAMBDA_INLINE -> double {
    double s1[1] = {0.0}, m1[1] = {0.0};
#pragma unroll
    for (int j = 0; j < P; ++j) {
      if (act[j] && Gv[j] == Gv[j]) s1[0] += Gv[j] * Gv[j];
      if (act[j] && Gv[j] != Gv[j]) m1[0] = 1.0;
    }
    block_reduce_mixed<1, 1>(s1, m1, red);
    return m1[0] > 0.0 ? NAN : sqrt(s1[0]) / nd;
  };
  double r = normG(Gpost);

  int exit_reason = ECONPDES_EXIT_CONVERGED, n_outer = 0, n_newton = 0, log_len = 0, error = 0;

  // One Newton solve of the implicit step from ypost with time step `delta`
  // (implicit_timestep + _newtonstep + NLsolve newton_, finiteschemesolve.jl:125-230).
  // On return y/Gy hold the Newton result and G there; nl = ||R||_2/n (Inf: singular).
  double y[P], Gy[P], nl = 0.0;
#pragma unroll
  for (int j = 0; j < P; ++j) { y[j] = ypost[j]; Gy[j] = Gpost[j]; }
  int newton_it = 0;
  bool any_singular = false;
  // Block-wide OR of a small bit mask in one barrier: warps fold their lanes with REDUX and
  // post the result with one shared atomic.  Three slots rotate: call k uses slot k%3 and,
  // after its barrier, clears slot (k+2)%3 -- last read before this barrier, next written
  // after the following call's barrier -- so no extra synchronisation is needed.
  unsigned* flagw = reinterpret_cast<unsigned*>(red + 7 * 32);
  int flag_phase = 0;
  if (t < 3) flagw[t] = 0u;
  __syncthreads();
  auto block_or_bits = [&](unsigned bits) EP_LAMBDA_INLINE -> unsigned {
    const unsigned w = __reduce_or_sync(0xffffffffu, bits);
    if ((t & 31) == 0 && w != 0u) atomicOr(&flagw[flag_phase], w);
    __syncthreads();
    const unsigned all = flagw[flag_phase];
    if (t == 0) flagw[(flag_phase + 2) % 3] = 0u;
    flag_phase = (flag_phase + 1) % 3;
    return all;
  };

  auto newton = [&](double delta, double ftol, int maxit) EP_LAMBDA_INLINE {
    RowTerms rt[P];
    const double rdelta = 1.0 / delta;   // once per Newton solve; 1/Inf = 0 makes mdiv return the IEEE quotient 0
    double R[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      rt[j].implicit = true; rt[j].has_bounds = false; rt[j].ypost = ypost[j]; rt[j].delta = delta;
      rt[j].is_alg = false; rt[j].lb = 0.0; rt[j].ub = 0.0; rt[j].rdelta = rdelta; rt[j].rdelta_always = true;
      y[j] = ypost[j];
      Gy[j] = Gpost[j];
      R[j] = act[j] ? apply_row_terms(Gy[j], y[j], rt[j]) : 0.0;
    }
    newton_it = 0;
    // NLsolve's loop only tests predicates of the residual (||R||_inf <= ftol, any NaN,
    // x == xold) -- folded block-wide as bits -- plus, for the Jacobian, FiniteDiff's three
    // masked sums of squares of the current iterate.  ||R||_2 is only needed once, at the end.
    enum : unsigned { kNaN = 1u, kBig = 2u, kMoved = 4u, kBad = 8u, kUntame = 16u };
    auto state_bits = [&](unsigned extra) EP_LAMBDA_INLINE -> unsigned {
      unsigned bb = extra;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        if (act[j]) {
          if (R[j] != R[j] || y[j] != y[j]) bb |= kNaN;
          if (!(fabs(R[j]) <= ftol)) bb |= kBig;
          if (!tame_value(y[j])) bb |= kUntame;
        }
      }
      return block_or_bits(bb);
    };
    unsigned bits = state_bits(kMoved);
    bool anynan = (bits & kNaN) != 0u;
    bool fconv = !anynan && (bits & kBig) == 0u;
    bool xconv = false;
    bool singular = false;
    const int c0 = i0 % 3;   // color of point j is (c0 + j) % 3
    // y is already published in ysh by the caller when y == ypost
    // level 3 divides by the step with the select-free quotient: the step must be finite and of moderate magnitude
    const bool step_ok = kTameLevel < 3 || moderate_magnitude(delta);
    bool tame = geom_tame && step_ok && (bits & kUntame) == 0u;
    // one Newton iteration; returns false when the linear solve broke down
    auto iterate = [&](auto tame_tag) EP_LAMBDA_INLINE -> bool {
      constexpr int TAME = decltype(tame_tag)::value;
      constexpr int ROWKIND = TAME >= 3 ? 3 : 2;
      ++newton_it;
      double s3[3] = {0.0, 0.0, 0.0}, m0[1];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const double q = act[j] ? y[j] * y[j] : 0.0;
        const int cj = (c0 + j) % 3;
        s3[0] += cj == 0 ? q : 0.0;
        s3[1] += cj == 1 ? q : 0.0;
        s3[2] += cj == 2 ? q : 0.0;
      }
      block_reduce_mixed<3, 0>(s3, m0, red);
      // FiniteDiff: sqrt(eps)*max(1, sqrt(||x .* mask_c||_2)).  Lanes 0..2 of every warp turn one
      // color each into its step and the step's reciprocal (two square roots and a division,
      // executed once per warp instead of three times), then broadcast.
      double eps_table[3], reps_table[3];
      {
        const int lane = t & 31;
        const double mine = lane == 0 ? s3[0] : (lane == 1 ? s3[1] : s3[2]);
        const double e = color_epsilon_from_sumsq(mine);
        // e >= sqrt(eps) and, for tame unknowns, far below 2^200: the select-free reciprocal has the bits of 1.0 / e
        const double re = TAME >= 1 ? ep_rcp_tame(e) : 1.0 / e;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          eps_table[c] = __shfl_sync(0xffffffffu, e, c);
          reps_table[c] = __shfl_sync(0xffffffffu, re, c);
        }
      }
      double ja[P], jb[P], jc[P], rhs[P];
      {
        double yv[P + 2];
#pragma unroll
        for (int k = 0; k < P + 2; ++k) yv[k] = ysh[i0 + k];
#pragma unroll
        for (int j = 0; j < P; ++j) {
          // inactive points (beyond the grid, last thread only) are evaluated too and replaced by
          // identity rows: a select instead of a branch keeps the P evaluations in one basic block
          double w[1][3] = {{yv[j], yv[j + 1], yv[j + 2]}};
          double r0[1] = {R[j]};
          double va = 0.0, vb = 1.0, vc = 0.0;
          jacobian_point<M, false, true, PER_COL ? 1 : 0, TAME, ROWKIND>(pre[j], g[j], idx[j], gv, w, bl, bh, &rt[j], eps_table, PER_COL,
                                                          r0, [&](int k, int, int, double v) {
                                                            if (k == 0) vb = v; else if (k == 1) va = v; else vc = v;
                                                          }, reps_table);
          ja[j] = act[j] ? va : 0.0; jb[j] = act[j] ? vb : 1.0; jc[j] = act[j] ? vc : 0.0; rhs[j] = R[j];
        }
      }
      double sol[P];
      block_chunk_solve<P>(ja, jb, jc, rhs, t, (S + P - 1) / P, pcr, xch, sol);
      unsigned extra = 0u;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        if (act[j] && !isfinite(sol[j])) extra |= kBad;
        const double xold = y[j];
        y[j] = xold + (-sol[j]);
        if (act[j] && y[j] != xold) extra |= kMoved;
      }
      publish(y);
      evalG(Gy, tame_tag);
#pragma unroll
      for (int j = 0; j < P; ++j) R[j] = act[j] ? apply_row_terms_k<ROWKIND>(Gy[j], y[j], rt[j]) : 0.0;
      bits = state_bits(extra);
      if (TAME != 0 && (bits & kUntame) != 0u) {     // the new iterate is wild: its residual again, general form
        evalG(Gy, General{});
#pragma unroll
        for (int j = 0; j < P; ++j) R[j] = act[j] ? apply_row_terms(Gy[j], y[j], rt[j]) : 0.0;
        bits = state_bits(extra);
      }
      tame = geom_tame && step_ok && (bits & kUntame) == 0u;
      if (bits & kBad) return false;                 // a vanishing pivot: SingularException
      anynan = (bits & kNaN) != 0u;
      fconv = !anynan && (bits & kBig) == 0u;
      xconv = !anynan && (bits & kMoved) == 0u;      // chebyshev(x, xold) <= xtol = 0
      return true;
    };
    while (!anynan && !fconv && !xconv && newton_it < maxit) {
      const bool ok = tame ? iterate(Tame{}) : iterate(General{});
      if (!ok) { singular = true; break; }
    }
    if (singular) {   // -> (y0, Inf), finiteschemesolve.jl:142-146
      any_singular = true;
#pragma unroll
      for (int j = 0; j < P; ++j) { y[j] = ypost[j]; Gy[j] = Gpost[j]; }
      nl = INFINITY;
      publish(y);
    } else if (anynan) {
      nl = NAN;
    } else {
      double s1[1] = {0.0}, m1[1];
#pragma unroll
      for (int j = 0; j < P; ++j) s1[0] += act[j] ? R[j] * R[j] : 0.0;
      block_reduce_mixed<1, 0>(s1, m1, red);
      nl = sqrt(s1[0]) / nd;
    }
  };

  if (r != r) {
    error = ECONPDES_E_NAN_GUESS;                    // :372
  } else if (r <= opt.abstol) {
    exit_reason = ECONPDES_EXIT_ALREADY_ZERO;        // :373-376
  } else if (isinf(opt.delta)) {                     // :377-389
    newton(INFINITY, opt.abstol, opt.maxiters);
    n_newton = newton_it;
#pragma unroll
    for (int j = 0; j < P; ++j) ypost[j] = y[j];
    r = nl;
    exit_reason = r <= opt.abstol ? ECONPDES_EXIT_CONVERGED : ECONPDES_EXIT_INF_STEP_FAILED;
  } else {
    SerController ctl;
    ctl.init(fabs(opt.delta), opt.scale, opt.max_delta);
    const double drel = opt.delta < 0 ? -1.0 : 1.0;   // sign(Δ), :397-398
    int iter = 0;
    while (iter < opt.maxiters && ctl.delta >= opt.min_delta && r > opt.abstol) {  // :406
      ++iter;
      newton(ctl.delta * drel, opt.inner_abstol, opt.inner_maxiters);
      n_newton += newton_it;
      const double old_r = r;
      double rn = normG(Gy);                                   // :415
      if (rn != rn) rn = INFINITY;                             // :417
      int status;
      if (nl <= opt.inner_abstol && rn <= opt.max_residual_growth * old_r) {   // :418
        status = ECONPDES_STEP_ACCEPTED;
        if (t == 0 && args.log != nullptr && log_len < args.log_capacity) {
          econpdes_log_row& row = args.log[b * args.log_capacity + log_len];
          row.iter = iter; row.status = status; row.newton_iters = newton_it; row.linear_iters = 0;
          row.delta = ctl.delta; row.residual = rn;
        }
        ctl.accept(old_r, rn);
#pragma unroll
        for (int j = 0; j < P; ++j) { ypost[j] = y[j]; Gpost[j] = Gy[j]; }
        r = rn;
      } else {
        if (nl <= opt.inner_abstol) status = ECONPDES_STEP_REJECTED_GROWTH;          // :425
        else if (nl != nl) status = ECONPDES_STEP_REJECTED_NAN;                       // :443
        else if (isinf(nl)) status = ECONPDES_STEP_REJECTED_SINGULAR;                 // :444
        else status = ECONPDES_STEP_REJECTED_INNER;                                   // :439
        if (t == 0 && args.log != nullptr && log_len < args.log_capacity) {
          econpdes_log_row& row = args.log[b * args.log_capacity + log_len];
          row.iter = iter; row.status = status; row.newton_iters = newton_it; row.linear_iters = 0;
          row.delta = ctl.delta;
          // growth rejections carry the factor the reference prints (finiteschemesolve.jl:433-435), like the generic engine
          row.residual = status == ECONPDES_STEP_REJECTED_GROWTH ? rn / old_r : NAN;
        }
        ctl.reject();
        publish(ypost);   // the next Newton solve starts from the reverted point
      }
      ++log_len;
    }
    n_outer = iter;
    if (r <= opt.abstol) exit_reason = ECONPDES_EXIT_CONVERGED;
    else if (ctl.delta < opt.min_delta) exit_reason = ECONPDES_EXIT_BELOW_MIN_DELTA;
    else exit_reason = ECONPDES_EXIT_MAXITERS;
  }

  publish(ypost);   // coalesced store of the solution through shared memory
  for (int k = t; k < S; k += T) args.solution[b * S + k] = ysh[k + 1];
  if (t == 0) {
    args.residual_norm[b] = r;
    args.exit_reason[b] = exit_reason;
    args.n_outer[b] = n_outer;
    args.n_newton[b] = n_newton;
    args.error[b] = error;
    if (args.log_len != nullptr) args.log_len[b] = log_len < args.log_capacity ? log_len : args.log_capacity;
    if (args.flags != nullptr) args.flags[b] = any_singular ? 1 : 0;
    if (any_singular && args.flagged_count != nullptr) atomicAdd_system(args.flagged_count, 1);
  }
}

// T threads of P points each
inline size_t batched_smem_bytes(int T, int P) { return sizeof(double) * ((T * P + 2) + 8 * 32 + 16 + 8 * T + 4 * T); }

}  // namespace econpdes
