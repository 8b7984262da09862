// Grid-wide kernels of the generic (single large grid) engine: one thread per grid point,
// fields in struct-of-arrays (S.., F) column-major layout so that consecutive threads read
// consecutive addresses of every unknown's plane.
//
//   k_residual      G!(ydot, y) (src/pdesolve.jl:436-462) fused with the implicit-Euler /
//                   MCP wrapper (src/finiteschemesolve.jl:158-162, 197-203) and with the
//                   three reductions the drivers need (||.||_2^2, ||.||_inf, any NaN)
//   k_color_sumsq   ||x .* mask_c||^2 for all colors in one pass -> FiniteDiff's per-color step
//   k_jacobian      fused colored finite-difference Jacobian in stencil-coefficient planes
//   k_saved         _fill_saved! (src/pdesolve.jl:424-433)
//   k_newton_update x <- x + p, p = -(J \ R) with NLsolve's x-convergence / NaN checks
#pragma once
#include "common.cuh"
#include "pointops.cuh"

namespace econpdes {

struct ParamPack {
  double v[kMaxParams];
};

// Device-side description of R(y) = G(y) - (!alg) (ypost - y) / (delta * delta_rel), then MCP.
struct RowSpec {
  const double* ypost;      // nullptr: R = G
  double delta;             // stepper.Δ (already signed when delta_rel == nullptr)
  double rdelta;            // RN(1/delta) for a finite scalar delta (set by the host), else 0
  const double* delta_rel;  // nullptr or per-row signed profile Δrel (finiteschemesolve.jl:393-399)
  const unsigned char* is_alg;
  const double* lb;
  const double* ub;
  int has_bounds;
};

__device__ __forceinline__ RowTerms make_row_terms(const RowSpec& rs, long long j) {
  RowTerms t;
  t.implicit = rs.ypost != nullptr;
  t.has_bounds = rs.has_bounds != 0;
  t.ypost = t.implicit ? rs.ypost[j] : 0.0;
  t.delta = rs.delta_rel != nullptr ? rs.delta * rs.delta_rel[j] : rs.delta;
  t.rdelta = rs.delta_rel != nullptr ? 0.0 : rs.rdelta;
  t.is_alg = rs.is_alg != nullptr ? rs.is_alg[j] != 0 : false;
  t.lb = (t.has_bounds && rs.lb != nullptr) ? rs.lb[j] : -INFINITY;
  t.ub = (t.has_bounds && rs.ub != nullptr) ? rs.ub[j] : INFINITY;
  return t;
}

// RECIP: see PointGeom (stencil.cuh)
// single-evaluation sweeps: tabulated reciprocals pay off from three dimensions on (15+ stencil
// quotients per point); below that the extra registers cost more than the divisions (measured)
template <int N>
constexpr int kSweepRecip = N >= 2 ? 1 : 0;

template <class M, int RECIP>
struct PointLoad {
  static constexpr int N = M::N, F = M::F, K = ipow3(M::N);
  long long idx[N];
  PointGeom<N, RECIP> g;
  double st[N];
  double w[F][K], bl[F][N], bh[F][N];
  typename M::Pre pre;
  __device__ __forceinline__ void load(const GridView<N>& gv, long long pt, const double* __restrict__ y,
                                       const double* __restrict__ bc, const double* prm) {
    unravel<N>(pt, gv, idx);
    load_geom(gv, idx, g, st);
    M::prepare(prm, st, pre);
#pragma unroll
    for (int f = 0; f < F; ++f) {
      load_box<N>(y + f * gv.l, pt, gv, g, w[f]);
      load_bc<N>(bc != nullptr ? bc + f * gv.l : nullptr, pt, idx, gv, g, bl[f], bh[f]);
    }
  }
};

template <class M, int BLOCK = kBlock, int MINB = 2>
__global__ void __launch_bounds__(BLOCK, MINB) k_residual(GridView<M::N> gv, ParamPack prm, const double* __restrict__ y,
                                                     const double* __restrict__ bc, RowSpec rs,
                                                     double* __restrict__ out, NormWork nw, int want_norm) {
  constexpr int F = M::F;
  __shared__ double scratch[96];
  double s = 0.0, m = 0.0;
  int nanf = 0;
  // grid-stride: a bounded number of blocks (sweep_blocks) walks the points, so the reduction and the
  // ticket hand-shake that end a norm-producing sweep are paid once per ~50 tiles instead of holding
  // every 256-point block resident for a fence + atomic round trip (that tail cost 0.44 ms of a 1.0 ms
  // sweep at 256^3)
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long pt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pt < gv.l; pt += stride) {
    PointLoad<M, kSweepRecip<M::N>> P;
    P.load(gv, pt, y, bc, prm.v);
    double o[F];
    eval_point<M, false>(P.pre, P.g, P.w, P.bl, P.bh, o, nullptr);
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const long long j = f * gv.l + pt;
      const RowTerms t = make_row_terms(rs, j);
      const double r = apply_row_terms(o[f], P.w[f][(ipow3(M::N) - 1) / 2], t);
      out[j] = r;
      if (r != r) {
        nanf = 1;
      } else {
        s += r * r;
        const double a = fabs(r);
        m = a > m ? a : m;
      }
    }
  }
  if (want_norm) {
    // a NaN anywhere makes the 2-norm NaN, as in norm(ydot) (finiteschemesolve.jl:371,415)
    grid_norm_finish(s, m, nanf, nw, scratch);
  }
}

// Masked sums of squares for every color in one pass.  Lanes of a warp that share a color
// are folded with shuffles (3 colors per warp in the common case of a warp inside one grid
// row), warps fold into per-block bins, blocks into `part`, and the last block finishes in
// block order and converts the sums into FiniteDiff steps.  ncolors <= kMaxColors.
constexpr int kMaxColors = 27 * 8;   // 3^3 grid colors x ECONPDES_MAX_FUNCS unknowns

struct ColorWork {
  double* part;       // [max_blocks][ncolors]
  unsigned* counter;
  double* eps;        // [2 * ncolors] out: sqrt(eps)*max(1, sqrt(||x.*mask_c||_2)), then the reciprocals
  double* sumsq;      // [ncolors] out
};

template <int N>
__global__ void __launch_bounds__(kBlock) k_color_sumsq(GridView<N> gv, int F, const double* __restrict__ y,
                                                        ColorWork cw) {
  __shared__ double bins[kBlock / 32][kMaxColors];
  __shared__ double fin[kBlock / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ncg = ncolors_grid<N>(gv);
  const int nc = ncg * F;
  for (int c = lane; c < nc; c += 32) bins[warp][c] = 0.0;
  __syncwarp();
  // persistent grid: a bounded number of blocks stride over the points, so the final fold
  // reads a few hundred partials per color instead of one per 256 points
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long npass = (gv.l + stride - 1) / stride;
  for (long long pass = 0; pass < npass; ++pass) {
    const long long pt = pass * stride + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool active = pt < gv.l;
    int color0 = -1;
    if (active) {
      long long idx[N];
      unravel<N>(pt, gv, idx);
      color0 = point_color0<N>(idx, gv);
    }
    for (int f = 0; f < F; ++f) {
      const double v = active ? y[f * gv.l + pt] : 0.0;
      const double q = v * v;
      unsigned remaining = __ballot_sync(0xffffffffu, active);
      while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int c = __shfl_sync(0xffffffffu, color0, leader);
        const bool mine = active && color0 == c;
        const double sum = warp_sum(mine ? q : 0.0);
        if (lane == leader) bins[warp][c + f * ncg] += sum;
        remaining &= ~__ballot_sync(0xffffffffu, mine);
      }
    }
  }
  __syncthreads();
  const int nwarps = blockDim.x >> 5;
  for (int c = threadIdx.x; c < nc; c += blockDim.x) {
    double t = 0.0;
    for (int w = 0; w < nwarps; ++w) t += bins[w][c];
    cw.part[static_cast<long long>(blockIdx.x) * nc + c] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(cw.counter, 1u);
    is_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // fixed-order fold: every color is summed by the whole block (strided partials, then
    // warp butterflies, then warps in order), so the result does not depend on scheduling
    for (int c = 0; c < nc; ++c) {
      double t = 0.0;
      for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += cw.part[static_cast<long long>(b) * nc + c];
      t = warp_sum(t);
      if (lane == 0) fin[warp] = t;
      __syncthreads();
      if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < nwarps; ++w) tot += fin[w];
        cw.sumsq[c] = tot;
        const double e = color_epsilon_from_sumsq(tot);
        cw.eps[c] = e;
        cw.eps[nc + c] = 1.0 / e;   // RN(1/eps): the Jacobian kernel divides through it (mdiv)
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) *cw.counter = 0u;
  }
}

// Fused colored FD Jacobian: planes[((k*F + fin)*F + fout)*l + pt] = dR_fout(pt)/dy_fin(pt+off_k)
// for the stencil shape <N,BOX>.  r0 (may be nullptr) receives R(y).
// POLICY: 0 = per-color steps (eps_table holds the steps, then their reciprocals), 1 = per-column steps
// (see jacobian_point); MINB = resident blocks per SM asked of the register allocator.
template <class M, bool BOX, int POLICY, int BLOCK = kBlock, int MINB = 2>
__global__ void __launch_bounds__(BLOCK, MINB) k_jacobian(GridView<M::N> gv, ParamPack prm, const double* __restrict__ y,
                                                           const double* __restrict__ bc, RowSpec rs,
                                                           const double* __restrict__ eps_table, int eps_per_column,
                                                           double* __restrict__ planes, double* __restrict__ r0out) {
  constexpr int F = M::F;
  const long long pt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pt >= gv.l) return;
  PointLoad<M, M::N >= 2 ? (M::USES_CROSS ? 2 : 1) : 0> P;
  P.load(gv, pt, y, bc, prm.v);
  RowTerms rt[F];
#pragma unroll
  for (int f = 0; f < F; ++f) rt[f] = make_row_terms(rs, f * gv.l + pt);
  double r0[F];
  const long long l = gv.l;
  const double* reps_table = POLICY == 0 ? eps_table + ncolors_grid<M::N>(gv) * F : nullptr;
  jacobian_point<M, BOX, false, POLICY>(P.pre, P.g, P.idx, gv, P.w, P.bl, P.bh, rt, eps_table, eps_per_column, r0,
                                        [&](int k, int fin, int fout, double v) {
                                          planes[(static_cast<long long>(k * F + fin) * F + fout) * l + pt] = v;
                                        }, reps_table);
  if (r0out != nullptr) {
#pragma unroll
    for (int f = 0; f < F; ++f) r0out[f * l + pt] = r0[f];
  }
}

template <class M>
__global__ void __launch_bounds__(kBlock) k_saved(GridView<M::N> gv, ParamPack prm, const double* __restrict__ y,
                                                  const double* __restrict__ bc, double* __restrict__ saved) {
  const long long pt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pt >= gv.l) return;
  PointLoad<M, kSweepRecip<M::N>> P;
  P.load(gv, pt, y, bc, prm.v);
  double o[M::F], sv[M::NSAVED > 0 ? M::NSAVED : 1];
  eval_point<M, true>(P.pre, P.g, P.w, P.bl, P.bh, o, sv);
#pragma unroll
  for (int s = 0; s < M::NSAVED; ++s) saved[s * gv.l + pt] = sv[s];
}

// x <- xold + p with p = -s.  out[1] = max |x - xold| (NLsolve's x_converged test is
// chebyshev(x, xold) <= xtol = 0), out[2] = any NaN in x.
static __global__ void __launch_bounds__(kBlock) k_newton_update(long long n, const double* __restrict__ xold,
                                                          const double* __restrict__ s, double* __restrict__ x,
                                                          NormWork nw) {
  __shared__ double scratch[96];
  double sum = 0.0, m = 0.0;
  int nanf = 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += stride) {
    const double p = -s[j];
    const double xo = xold[j];
    const double xn = xo + p;
    x[j] = xn;
    if (xn != xn) nanf = 1;
    const double d = fabs(xn - xo);
    if (d == d) m = d > m ? d : m;
  }
  grid_norm_finish(sum, m, nanf, nw, scratch);
}

// Gather stencil planes into the CSC nzval of local_stencil_jacobian via a precomputed map
// (map[e] < 0: structural zero of a star stencil).
static __global__ void __launch_bounds__(kBlock) k_gather_csc(long long nnz, const long long* __restrict__ map,
                                                       const double* __restrict__ planes, double* __restrict__ nzval) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const long long src = map[e];
  nzval[e] = src >= 0 ? planes[src] : 0.0;
}

}  // namespace econpdes
