// Per-grid-point building blocks shared by every kernel of the hot path:
//   * gather the +-1 box of an unknown around a point            (pde!, pdesolve.jl:443-447)
//   * evaluate the model residual there                          (pdesolve.jl:448-454)
//   * the implicit-Euler / MCP residual                          (finiteschemesolve.jl:158-162, 197-203)
//   * one block-row of the colored finite-difference Jacobian    (FiniteDiff forward mode,
//     called at finiteschemesolve.jl:184) in the fused row-wise form: because the coloring
//     of coloring.jl:52-70 is distance-2, a row sees at most one perturbed column per
//     color, so perturbing the box slots one at a time with that column's color epsilon
//     reproduces the C global sweeps entry for entry.
// Everything is __host__ __device__ so tests/host_emu can run the identical code on the
// CPU of the build container (which has no GPU); the shipped library has no CPU path.
#pragma once
#include "models.cuh"

namespace econpdes {

constexpr int kMaxDims = 4;
constexpr double kSqrtEps = 1.4901161193847656e-8;

// Tensor grid: per-dimension coordinate vectors, sizes and column-major strides.
// A slab-decomposed grid (one rank of a multi-GPU solve) keeps the GLOBAL sizes in S -- they
// decide boundaries, spacings and colors -- and owns `l` points: all of dimensions 0..N-2 and
// a contiguous range of the slowest dimension starting at global index `off`.  Fields of a
// slab carry one ghost plane on either side, so neighbour addresses pt +- stride stay valid.
// Unsigned 32-bit division by a run-time invariant (Granlund-Montgomery, round-up multiplier):
// q = (t + ((n - t) >> s1)) >> s2 with t = umulhi(m, n); exact for every n < 2^32 and d >= 1.
struct FastDiv {
  unsigned m = 0, s1 = 0, s2 = 0, d = 1;
  EP_HD void set(unsigned dd) {
    d = dd;
    unsigned l = 0;
    while ((1ull << l) < dd) ++l;                                    // ceil(log2 d)
    m = static_cast<unsigned>(((1ull << 32) * ((1ull << l) - dd)) / dd + 1);
    s1 = l < 1 ? l : 1;
    s2 = l > 0 ? l - 1 : 0;
  }
  EP_HD unsigned div(unsigned n) const {
#if defined(__CUDA_ARCH__)
    const unsigned t = __umulhi(m, n);
#else
    const unsigned t = static_cast<unsigned>((static_cast<unsigned long long>(m) * n) >> 32);
#endif
    return (t + ((n - t) >> s1)) >> s2;
  }
};

template <int N>
struct GridView {
  const double* x[N];
  long long S[N];
  long long stride[N];
  long long l;        // number of grid points owned (= prod S on a single device)
  long long off = 0;  // global index of the first owned entry of dimension N-1
  // optional per-dimension geometry table [S_d][kGeomRows]: dm, dp, d, 1/dp, 1/dm, 1/(dp d),
  // 1/(dm d), 1/(dp dm) -- built once per grid (make_geom_table), so a sweep reads one 64-byte
  // record per dimension (four 16-byte loads) instead of re-deriving it with five divisions
  const double* tab[N] = {};
  FastDiv div[N];     // division by S[d] (set_sizes)
  // sizes, column-major strides and the divisors unravel uses
  EP_HD void set_sizes(const long long* sizes) {
    long long st = 1;
    for (int d = 0; d < N; ++d) {
      S[d] = sizes[d];
      stride[d] = st;
      st *= sizes[d];
      div[d].set(static_cast<unsigned>(sizes[d]));
    }
    l = st;
  }
};
constexpr int kGeomRows = 8;
struct alignas(16) Pair { double a, b; };

// Stencil shapes: the full clipped +-1 box (coloring.jl:18-35) or, for models that never
// read a cross derivative, its star subset (centre + 2N axial neighbours).
template <int N, bool BOX>
struct Stencil {
  static constexpr int NK = BOX ? ipow3(N) : 1 + 2 * N;
  EP_HD static constexpr int off(int k, int d) {
    if (BOX) return (k / p3(d)) % 3 - 1;
    if (k == 0) return 0;
    return ((k - 1) / 2 == d) ? (((k - 1) % 2) ? 1 : -1) : 0;
  }
  EP_HD static constexpr int boxslot(int k) {
    if (BOX) return k;
    int s = (ipow3(N) - 1) / 2;
    if (k == 0) return s;
    return s + (((k - 1) % 2) ? 1 : -1) * p3((k - 1) / 2);
  }
};

template <int N>
EP_HD void unravel(long long pt, const GridView<N>& gv, long long* idx) {
  // a device holds fewer than 2^31 points (checked when the problem is created): 32-bit division
  unsigned p = static_cast<unsigned>(pt);
#pragma unroll
  for (int d = 0; d + 1 < N; ++d) {
    const unsigned q = gv.div[d].div(p);
    idx[d] = p - q * static_cast<unsigned>(gv.S[d]);
    p = q;
  }
  idx[N - 1] = p + gv.off;   // global index along the slab dimension
}

// spacings (and their reciprocals) of grid index i along dimension d, plus the state coordinate
template <int N, int RECIP>
EP_HD void load_axis(const GridView<N>& gv, int d, long long i, PointGeom<N, RECIP>& g, double* state) {
  g.lo[d] = i == 0;
  g.hi[d] = i == gv.S[d] - 1;
  state[d] = gv.x[d][i];
  if (RECIP >= 1 && gv.tab[d] != nullptr) {
    // one 64-byte record per grid index, 16-byte aligned: four vector loads
    const Pair* t = reinterpret_cast<const Pair*>(gv.tab[d] + i * kGeomRows);
    const Pair t0 = t[0], t1 = t[1], t2 = t[2], t3 = t[3];
    g.dm[d] = t0.a; g.dp[d] = t0.b; g.d[d] = t1.a;
    g.rdp[d] = t1.b; g.rdm[d] = t2.a; g.rdpd[d] = t2.b; g.rdmd[d] = t3.a; g.rdpm[d] = t3.b;
    g.dpd[d] = g.dp[d] * g.d[d];
    g.dmd[d] = g.dm[d] * g.d[d];
    g.dpm[d] = g.dp[d] * g.dm[d];
  } else {
    spacing(gv.x[d], i, gv.S[d], g.dm[d], g.dp[d], g.d[d]);
    finish_axis(g, d);
  }
}

template <int N, int RECIP>
EP_HD void load_geom(const GridView<N>& gv, const long long* idx, PointGeom<N, RECIP>& g, double* state) {
#pragma unroll
  for (int d = 0; d < N; ++d) load_axis(gv, d, idx[d], g, state);
  finish_geom(g);
}

// Host side of the table (plain IEEE arithmetic, identical to the device's).
inline void make_geom_table(const double* x, long long S, double* tab /* [S][kGeomRows] */) {
  for (long long i = 0; i < S; ++i) {
    PointGeom<1, 1> g;
    spacing(x, i, S, g.dm[0], g.dp[0], g.d[0]);
    finish_axis(g, 0);
    double* t = tab + i * kGeomRows;
    t[0] = g.dm[0]; t[1] = g.dp[0]; t[2] = g.d[0];
    t[3] = g.rdp[0]; t[4] = g.rdm[0]; t[5] = g.rdpd[0];
    t[6] = g.rdmd[0]; t[7] = g.rdpm[0];
  }
}

// Gather the 3^N box of unknown f around point pt (slots outside the grid are never read
// by `differentiate`; they are filled with the centre value).
template <int N, class G>
EP_HD void load_box(const double* __restrict__ yf, long long pt, const GridView<N>& gv, const G& g, double* w) {
  constexpr int K = ipow3(N);
  int str[N];   // 32-bit offsets (see unravel)
#pragma unroll
  for (int d = 0; d < N; ++d) str[d] = static_cast<int>(gv.stride[d]);
  const double* __restrict__ yc = yf + pt;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int q = 0;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      const int o = (k / p3(d)) % 3 - 1;
      if ((o < 0 && g.lo[d]) || (o > 0 && g.hi[d])) ok = false;
      q += o * str[d];
    }
    w[k] = yc[ok ? q : 0];
  }
}

// bc[i1.., (first|last) along d, ..iN, f] keeping the other indices (derivatives.jl:27-32).
template <int N, class G>
EP_HD void load_bc(const double* __restrict__ bcf, long long pt, const long long* idx, const GridView<N>& gv,
                   const G& g, double* bc_lo, double* bc_hi) {
#pragma unroll
  for (int d = 0; d < N; ++d) {
    bc_lo[d] = 0.0;
    bc_hi[d] = 0.0;
    if (bcf != nullptr) {
      if (g.lo[d]) bc_lo[d] = bcf[pt];  // the point itself sits on the low face
      if (g.hi[d]) bc_hi[d] = bcf[pt];
    }
  }
  (void)idx;
  (void)gv;
}

// The row-local description of the implicit-Euler / MCP wrapper around G
// (finiteschemesolve.jl:158-162, 197-203).  `delta` already includes the signed relative
// profile (stepper.Δ .* Δrel).
struct RowTerms {
  bool implicit;   // subtract (ypost - y)/delta on non-algebraic rows
  bool has_bounds; // min(max(R, y-ub), y-lb)
  double ypost, delta, lb, ub;
  bool is_alg;
  double rdelta = 0.0;  // RN(1/delta) when the caller formed it once for many rows (finite delta only), else 0
  // the caller guarantees rdelta = RN(1/delta) for EVERY delta (1/Inf = 0 included: mdiv then returns
  // the IEEE quotient +-0), so the quotient never needs the division -- a compile-time constant in the
  // batched kernel, which removes the branch
  bool rdelta_always = false;
};

EP_HD double apply_row_terms(double g, double y, const RowTerms& t) {
  double r = g;
  // `.!is_algebraic .* (ypost .- y) ./ Δ`: Bool true is an exact 1, false a strong zero
  if (t.implicit && !t.is_alg)
    r = r - ((t.rdelta_always || t.rdelta != 0.0) ? mdiv(t.ypost - y, t.delta, t.rdelta) : (t.ypost - y) / t.delta);
  if (t.has_bounds) {
    const double a = y - t.ub, b = y - t.lb;
    double m = r > a ? r : a;      // max(r, y-ub)   (NaN in r propagates like Julia's max)
    if (r != r) m = r;
    r = m < b ? m : b;             // min(., y-lb)
    if (m != m) r = m;
  }
  return r;
}

// Compile-time forms of the row terms (ROWKIND): 0 = R = G; 1 = G - (ypost - y)/delta with a finite delta
// whose reciprocal the caller formed (t.rdelta), no algebraic rows, no bounds; 2 = general.
template <int ROWKIND>
EP_HD double apply_row_terms_k(double g, double y, const RowTerms& t) {
  if (ROWKIND == 0) return g;
  if (ROWKIND == 1) return g - mdiv(t.ypost - y, t.delta, t.rdelta);
  if (ROWKIND == 3) return g - mdiv_tame(t.ypost - y, t.delta, t.rdelta);   // 1 with a step of moderate magnitude and tame unknowns
  return apply_row_terms(g, y, t);
}

// Residual of all F unknowns at one point from the boxes w[f][3^N].
template <class M, bool SAVE, int TAME = 0, class G>
EP_HD void eval_point(const typename M::Pre& pre, const G& g,
                      const double (*w)[ipow3(M::N)], const double (*bc_lo)[M::N], const double (*bc_hi)[M::N],
                      double* out, double* sv) {
  Deriv<M::N> u[M::F];
#pragma unroll
  for (int f = 0; f < M::F; ++f) differentiate<TAME>(w[f], g, bc_lo[f], bc_hi[f], u[f]);
  if constexpr (TAME >= 3 && M::HAS_TAME_EVAL) M::template eval_tame<SAVE>(pre, u, out, sv);
  else M::template eval<SAVE>(pre, u, out, sv);
}

// FiniteDiff's step: sqrt(eps) * max(1, sqrt(||x .* mask_c||_2)); the kernel receives the
// per-color table.  With eps_policy = per-column the step is sqrt(eps)*max(1,|x_j|).
EP_HD double color_epsilon_from_sumsq(double sumsq) {
  const double nrm = sqrt(sumsq);
  const double e = kSqrtEps * fabs(sqrt(nrm));
  return e > kSqrtEps ? e : kSqrtEps;
}

template <int N>
EP_HD int point_color0(const long long* idx, const GridView<N>& gv) {  // 0-based grid color
  int base = 0, stride = 1;
#pragma unroll
  for (int d = 0; d < N; ++d) {
    base += static_cast<int>(idx[d] % 3) * stride;
    stride *= gv.S[d] < 3 ? static_cast<int>(gv.S[d]) : 3;
  }
  return base;
}

template <int N>
EP_HD int ncolors_grid(const GridView<N>& gv) {
  int stride = 1;
#pragma unroll
  for (int d = 0; d < N; ++d) stride *= gv.S[d] < 3 ? static_cast<int>(gv.S[d]) : 3;
  return stride;
}

// One block-row of the Jacobian of R = rowterms(G).  `emit(k, fin, fout, value)` receives
// d R_fout(pt) / d y_fin(pt + off_k) for every stencil entry k of shape <N,BOX> (zero for
// entries outside the grid).  r0 returns R(pt) itself (the unperturbed residual).
// With HAVE_R0 the caller already holds R(pt) in r0 (FiniteDiff re-evaluates f(x); the
// value is identical, so kernels that kept it skip that evaluation).
// POLICY: -1 = `eps_per_column` decides at run time; 0 = per-color steps from eps_table, with their
// reciprocals in reps_table (required); 1 = per-column steps.  The compile-time forms are branch-free.
// SKIP_UNUSED: entries of box slots the model never reads (M::slot_used) are not emitted at all (the caller's
// planes were zeroed once and nothing else writes them).
template <class M, bool BOX, bool HAVE_R0 = false, int POLICY = -1, int TAME = 0, int ROWKIND = 2,
          bool SKIP_UNUSED = false, class G, class Emit>
EP_HD void jacobian_point(const typename M::Pre& pre, const G& g, const long long* idx,
                          const GridView<M::N>& gv, double (*w)[ipow3(M::N)], const double (*bc_lo)[M::N],
                          const double (*bc_hi)[M::N], const RowTerms* rt, const double* eps_table,
                          int eps_per_column, double* r0, Emit emit, const double* reps_table = nullptr) {
  constexpr int N = M::N, F = M::F;
  constexpr int C = (ipow3(N) - 1) / 2;
  using St = Stencil<N, BOX>;
  if (!HAVE_R0) {
    double g0[F];
    eval_point<M, false, TAME>(pre, g, w, bc_lo, bc_hi, g0, nullptr);
#pragma unroll
    for (int f = 0; f < F; ++f) r0[f] = apply_row_terms_k<ROWKIND>(g0[f], w[f][C], rt[f]);
  }
  const int ncg = ncolors_grid<N>(gv);
  int i3[N];
#pragma unroll
  for (int d = 0; d < N; ++d) i3[d] = static_cast<int>(idx[d] % 3);
  // wide systems (the 8-unknown 3-D long-run-risk model: 216 columns per row) keep these two loops rolled: the
  // body is the whole model evaluation, and unrolling it a hundred times buys nothing but compile time
  constexpr bool WIDE = F * St::NK > 48;
  constexpr int UNR_FIN = WIDE ? 1 : F, UNR_K = WIDE ? 1 : St::NK;
#pragma unroll UNR_FIN
  for (int fin = 0; fin < F; ++fin) {
#pragma unroll UNR_K
    for (int k = 0; k < St::NK; ++k) {
      bool ok = true;
      int color = 0, stride = 1;
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const int o = St::off(k, d);
        if ((o < 0 && g.lo[d]) || (o > 0 && g.hi[d])) ok = false;
        color += ((i3[d] + o + 3) % 3) * stride;
        stride *= gv.S[d] < 3 ? static_cast<int>(gv.S[d]) : 3;
      }
      const int slot = St::boxslot(k);
      // entries outside the grid: skipped, or (branch-free forms) evaluated anyway and zeroed below
      if (SKIP_UNUSED && !M::slot_used(slot)) continue;
      if ((POLICY < 0 && !ok) || !M::slot_used(slot)) {
#pragma unroll
        for (int fout = 0; fout < F; ++fout) emit(k, fin, fout, 0.0);
        continue;
      }
      const double saved = w[fin][slot];
      double eps, reps = 0.0;   // reps: RN(1/eps) when the caller tabulated it with the color steps
      if (POLICY == 1 || (POLICY < 0 && eps_per_column)) {
        const double a = fabs(saved);
        eps = kSqrtEps * (a > 1.0 ? a : 1.0);
        if (POLICY == 1) reps = ep_rcp(eps);
      } else {
        eps = eps_table[color + fin * ncg];
        if (POLICY == 0 || reps_table != nullptr) reps = reps_table[color + fin * ncg];
      }
      w[fin][slot] = saved + eps;
      double g1[F];
      eval_point<M, false, TAME>(pre, g, w, bc_lo, bc_hi, g1, nullptr);
#pragma unroll
      for (int fout = 0; fout < F; ++fout) {
        const double r1 = apply_row_terms_k<ROWKIND>(g1[fout], w[fout][C], rt[fout]);
        // tame level 3 (finite residuals of tame unknowns, a color step of moderate magnitude): select-free quotient
        const double q = (TAME >= 3 && POLICY >= 0)
                             ? mdiv_tame(r1 - r0[fout], eps, reps)
                             : ((POLICY >= 0 || reps != 0.0) ? mdiv(r1 - r0[fout], eps, reps) : (r1 - r0[fout]) / eps);
        emit(k, fin, fout, ok ? q : 0.0);
      }
      w[fin][slot] = saved;
    }
  }
}

}  // namespace econpdes
