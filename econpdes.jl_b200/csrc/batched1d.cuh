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
#include "../../include/econpdes_cuda.h"
#include "kernels.cuh"
#include "tridiag.cuh"

namespace econpdes {

struct BatchArgs {
  long long nproblems;
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

template <class M>
__global__ void __launch_bounds__(1024, 1) k_batched_pdesolve_1d(BatchArgs args) {
  static_assert(M::N == 1 && M::F == 1, "the fused batched engine handles 1-D single-unknown models");
  extern __shared__ double smem[];
  const int T = blockDim.x;
  double* ysh = smem;                 // [T + 2]: y with one halo slot each side
  double* red = ysh + (T + 2);        // [8 * 32] reduction scratch
  double* prm = red + 8 * 32;         // [16]
  double* pcr = prm + 16;             // [8 * T]
  const long long b = blockIdx.x;
  const int i = threadIdx.x;
  const int S = args.S;
  const bool act = i < S;
  const econpdes_options& opt = args.opt;

  if (i < args.nparams) prm[i] = args.params[b * args.nparams + i];
  if (i == 0) { ysh[0] = 0.0; ysh[T + 1] = 0.0; }
  const double* xg = args.grids + b * S;
  GridView<1> gv;
  gv.x[0] = xg; gv.S[0] = S; gv.stride[0] = 1; gv.l = S;
  long long idx[1] = {act ? i : 0};
  PointGeom<1, 1> g;   // reciprocals of the spacings once per problem, reused by every stencil evaluation
  double st[1];
  load_geom(gv, idx, g, st);
  double bl[1][1] = {{0.0}}, bh[1][1] = {{0.0}};
  __syncthreads();
  typename M::Pre pre;   // state-only part of the model, once per problem
  M::prepare(prm, st, pre);

  // residual G at this thread's point from the y currently published in ysh
  auto evalG = [&]() -> double {
    double w[1][3] = {{ysh[i], ysh[i + 1], ysh[i + 2]}};
    double o[1];
    eval_point<M, false>(pre, g, w, bl, bh, o, nullptr);
    return o[0];
  };
  auto publish = [&](double v) {
    __syncthreads();  // everyone finished reading the previous y
    ysh[i + 1] = v;
    __syncthreads();
  };

  double ypost = act ? args.guess[b * S + i] : 0.0;
  publish(ypost);
  double Gpost = act ? evalG() : 0.0;
  const double nd = static_cast<double>(S);

  double sums[3], maxs[2];
  sums[0] = (act && Gpost == Gpost) ? Gpost * Gpost : 0.0;
  maxs[0] = (act && Gpost != Gpost) ? 1.0 : 0.0;
  block_reduce_mixed<1, 1>(sums, maxs, red);
  double r = maxs[0] > 0.0 ? NAN : sqrt(sums[0]) / nd;  // norm(ydot)/length(ydot), :371

  int exit_reason = ECONPDES_EXIT_CONVERGED, n_outer = 0, n_newton = 0, log_len = 0, error = 0;

  // One Newton solve of the implicit step from ypost with time step `delta`
  // (implicit_timestep + _newtonstep + NLsolve newton_, finiteschemesolve.jl:125-230).
  // On return y/Gy hold the Newton result and G there; nl = ||R||_2/n (Inf: singular).
  double y = ypost, Gy = Gpost, nl = 0.0;
  int newton_it = 0;
  // Block-wide OR of a small bit mask in one barrier: warps fold their lanes with REDUX and
  // post the result with one shared atomic.  Three slots rotate: call k uses slot k%3 and,
  // after its barrier, clears slot (k+2)%3 -- last read before this barrier, next written
  // after the following call's barrier -- so no extra synchronisation is needed.
  unsigned* flagw = reinterpret_cast<unsigned*>(red + 7 * 32);
  int flag_phase = 0;
  if (i < 3) flagw[i] = 0u;
  __syncthreads();
  auto block_or_bits = [&](unsigned bits) -> unsigned {
    const unsigned w = __reduce_or_sync(0xffffffffu, bits);
    if ((i & 31) == 0 && w != 0u) atomicOr(&flagw[flag_phase], w);
    __syncthreads();
    const unsigned all = flagw[flag_phase];
    if (i == 0) flagw[(flag_phase + 2) % 3] = 0u;
    flag_phase = (flag_phase + 1) % 3;
    return all;
  };

  auto newton = [&](double delta, double ftol, int maxit) {
    RowTerms rt[1];
    rt[0].implicit = true; rt[0].has_bounds = false; rt[0].ypost = ypost; rt[0].delta = delta;
    rt[0].is_alg = false; rt[0].lb = 0.0; rt[0].ub = 0.0;
    rt[0].rdelta = isfinite(delta) ? 1.0 / delta : 0.0;   // once per Newton solve
    y = ypost;
    Gy = Gpost;
    double R = act ? apply_row_terms(Gy, y, rt[0]) : 0.0;
    newton_it = 0;
    const int ci = i % 3;
    // NLsolve's loop only tests predicates of the residual (||R||_inf <= ftol, any NaN,
    // x == xold) -- folded block-wide as bits -- plus, for the Jacobian, FiniteDiff's three
    // masked sums of squares of the current iterate.  ||R||_2 is only needed once, at the end.
    enum : unsigned { kNaN = 1u, kBig = 2u, kMoved = 4u, kBad = 8u };
    auto state_bits = [&](bool moved, bool bad) -> unsigned {
      unsigned b = 0u;
      if (act) {
        if (R != R || y != y) b |= kNaN;
        if (!(fabs(R) <= ftol)) b |= kBig;
        if (moved) b |= kMoved;
        if (bad) b |= kBad;
      }
      return block_or_bits(b);
    };
    unsigned bits = state_bits(true, false);
    bool anynan = (bits & kNaN) != 0u;
    bool fconv = !anynan && (bits & kBig) == 0u;
    bool xconv = false;
    bool singular = false;
    // y is already published in ysh by the caller when y == ypost
    while (!anynan && !fconv && !xconv && newton_it < maxit) {
      ++newton_it;
      double s3[3], m0[1];
      const double q = act ? y * y : 0.0;
      s3[0] = ci == 0 ? q : 0.0;
      s3[1] = ci == 1 ? q : 0.0;
      s3[2] = ci == 2 ? q : 0.0;
      block_reduce_mixed<3, 0>(s3, m0, red);
      // FiniteDiff: sqrt(eps)*max(1, sqrt(||x .* mask_c||_2)).  Lanes 0..2 of every warp turn one
      // color each into its step and the step's reciprocal (two square roots and a division,
      // executed once per warp instead of three times), then broadcast.
      double eps_table[3], reps_table[3];
      {
        const int lane = i & 31;
        const double mine = lane == 0 ? s3[0] : (lane == 1 ? s3[1] : s3[2]);
        const double e = color_epsilon_from_sumsq(mine);
        const double re = 1.0 / e;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          eps_table[c] = __shfl_sync(0xffffffffu, e, c);
          reps_table[c] = __shfl_sync(0xffffffffu, re, c);
        }
      }
      double ja = 0.0, jb = 1.0, jc = 0.0;
      if (act) {
        double w[1][3] = {{ysh[i], ysh[i + 1], ysh[i + 2]}};
        double r0[1] = {R};
        jacobian_point<M, false, true>(pre, g, idx, gv, w, bl, bh, rt, eps_table,
                                       opt.eps_policy == ECONPDES_EPS_PER_COLUMN, r0,
                                       [&](int k, int, int, double v) {
                                         if (k == 0) jb = v; else if (k == 1) ja = v; else jc = v;
                                       }, reps_table);
      }
      const double sol = block_pcr_solve(ja, jb, jc, act ? R : 0.0, i, S, pcr);
      const bool bad = act && !isfinite(sol);
      const double xold = y;
      y = xold + (-sol);
      publish(y);
      Gy = act ? evalG() : 0.0;
      R = act ? apply_row_terms(Gy, y, rt[0]) : 0.0;
      bits = state_bits(y != xold, bad);
      if (bits & kBad) { singular = true; break; }   // a vanishing pivot: SingularException
      anynan = (bits & kNaN) != 0u;
      fconv = !anynan && (bits & kBig) == 0u;
      xconv = !anynan && (bits & kMoved) == 0u;      // chebyshev(x, xold) <= xtol = 0
    }
    if (singular) {   // -> (y0, Inf), finiteschemesolve.jl:142-146
      y = ypost;
      Gy = Gpost;
      nl = INFINITY;
      publish(y);
    } else if (anynan) {
      nl = NAN;
    } else {
      double s1[1], m1[1];
      s1[0] = act ? R * R : 0.0;
      block_reduce_mixed<1, 0>(s1, m1, red);
      nl = sqrt(s1[0]) / nd;
    }
  };

  auto normG = [&](double Gv) -> double {
    double s1[1], m1[1];
    s1[0] = (act && Gv == Gv) ? Gv * Gv : 0.0;
    m1[0] = (act && Gv != Gv) ? 1.0 : 0.0;
    block_reduce_mixed<1, 1>(s1, m1, red);
    return m1[0] > 0.0 ? NAN : sqrt(s1[0]) / nd;
  };

  if (r != r) {
    error = ECONPDES_E_NAN_GUESS;                    // :372
  } else if (r <= opt.abstol) {
    exit_reason = ECONPDES_EXIT_ALREADY_ZERO;        // :373-376
  } else if (isinf(opt.delta)) {                     // :377-389
    newton(INFINITY, opt.abstol, opt.maxiters);
    n_newton = newton_it;
    ypost = y;
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
        if (i == 0 && args.log != nullptr && log_len < args.log_capacity) {
          econpdes_log_row& row = args.log[b * args.log_capacity + log_len];
          row.iter = iter; row.status = status; row.newton_iters = newton_it; row.linear_iters = 0;
          row.delta = ctl.delta; row.residual = rn;
        }
        ctl.accept(old_r, rn);
        ypost = y;
        Gpost = Gy;
        r = rn;
      } else {
        if (nl <= opt.inner_abstol) status = ECONPDES_STEP_REJECTED_GROWTH;          // :425
        else if (nl != nl) status = ECONPDES_STEP_REJECTED_NAN;                       // :443
        else if (isinf(nl)) status = ECONPDES_STEP_REJECTED_SINGULAR;                 // :444
        else status = ECONPDES_STEP_REJECTED_INNER;                                   // :439
        if (i == 0 && args.log != nullptr && log_len < args.log_capacity) {
          econpdes_log_row& row = args.log[b * args.log_capacity + log_len];
          row.iter = iter; row.status = status; row.newton_iters = newton_it; row.linear_iters = 0;
          row.delta = ctl.delta; row.residual = NAN;
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

  if (act) args.solution[b * S + i] = ypost;
  if (i == 0) {
    args.residual_norm[b] = r;
    args.exit_reason[b] = exit_reason;
    args.n_outer[b] = n_outer;
    args.n_newton[b] = n_newton;
    args.error[b] = error;
    if (args.log_len != nullptr) args.log_len[b] = log_len < args.log_capacity ? log_len : args.log_capacity;
  }
}

inline size_t batched_smem_bytes(int T) { return sizeof(double) * ((T + 2) + 8 * 32 + 16 + 8 * T); }

}  // namespace econpdes
