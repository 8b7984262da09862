// TEST INFRASTRUCTURE: runs the product's __host__ __device__ per-point code
// (csrc/stencil.cuh, models.cuh, pointops.cuh) in plain CPU loops so that the stencil,
// functor and fused-Jacobian arithmetic can be checked against the oracle in the build
// container, which has no GPU.  Not linked into libeconpdes_cuda.so; not a fallback.
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>

#include "pointops.cuh"

using namespace econpdes;

namespace {

template <class M>
GridView<M::N> make_grid(const long long* S, const double* const* grids) {
  GridView<M::N> gv;
  gv.set_sizes(S);
  for (int d = 0; d < M::N; ++d) gv.x[d] = grids[d];
  return gv;
}

// the caller's parameters followed by the model's derived constants (as the engine builds them)
template <class M>
struct full_params {
  double v[kMaxParams] = {};
  explicit full_params(const double* prm) {
    for (int i = 0; i < M::NPARAMS; ++i) v[i] = prm[i];
    M::derive(v);
  }
};

int g_recip = 0;   // PointGeom RECIP mode under test (all modes must give the same bits)

template <class M, int RECIP>
int residual(const long long* S, const double* const* grids, const double* prm, const double* y, const double* bc,
             double* out, double* saved) {
  constexpr int N = M::N, F = M::F, K = ipow3(N);
  auto gv = make_grid<M>(S, grids);
  for (long long pt = 0; pt < gv.l; ++pt) {
    long long idx[N];
    unravel<N>(pt, gv, idx);
    PointGeom<N, RECIP> g;
    double st[N];
    load_geom(gv, idx, g, st);
    double w[F][K], bl[F][N], bh[F][N];
    for (int f = 0; f < F; ++f) {
      load_box<N>(y + f * gv.l, pt, gv, g, w[f]);
      load_bc<N>(bc ? bc + f * gv.l : nullptr, pt, idx, gv, g, bl[f], bh[f]);
    }
    double o[F], sv[M::NSAVED > 0 ? M::NSAVED : 1];
    typename M::Pre pre;
    M::prepare(full_params<M>(prm).v, st, pre);
    if (saved) eval_point<M, true>(pre, g, w, bl, bh, o, sv);
    else eval_point<M, false>(pre, g, w, bl, bh, o, sv);
    for (int f = 0; f < F; ++f) out[f * gv.l + pt] = o[f];
    if (saved) for (int s = 0; s < M::NSAVED; ++s) saved[s * gv.l + pt] = sv[s];
  }
  return 0;
}

template <class M, bool BOX, int RECIP>
int jacobian(const long long* S, const double* const* grids, const double* prm, const double* y, const double* bc,
             const double* ypost, double delta, const double* eps_table, int eps_per_column, double* planes,
             double* r0out) {
  constexpr int N = M::N, F = M::F, K = ipow3(N);
  auto gv = make_grid<M>(S, grids);
  for (long long pt = 0; pt < gv.l; ++pt) {
    long long idx[N];
    unravel<N>(pt, gv, idx);
    PointGeom<N, RECIP> g;
    double st[N];
    load_geom(gv, idx, g, st);
    double w[F][K], bl[F][N], bh[F][N];
    RowTerms rt[F];
    for (int f = 0; f < F; ++f) {
      load_box<N>(y + f * gv.l, pt, gv, g, w[f]);
      load_bc<N>(bc ? bc + f * gv.l : nullptr, pt, idx, gv, g, bl[f], bh[f]);
      rt[f].implicit = ypost != nullptr;
      rt[f].has_bounds = false;
      rt[f].ypost = ypost ? ypost[f * gv.l + pt] : 0.0;
      rt[f].delta = delta;
      rt[f].is_alg = false;
      rt[f].lb = rt[f].ub = 0.0;
    }
    double r0[F];
    const long long l = gv.l;
    typename M::Pre pre;
    M::prepare(full_params<M>(prm).v, st, pre);
    jacobian_point<M, BOX>(pre, g, idx, gv, w, bl, bh, rt, eps_table, eps_per_column, r0,
                           [&](int k, int fin, int fout, double v) { planes[((long long)(k * F + fin) * F + fout) * l + pt] = v; });
    for (int f = 0; f < F; ++f) r0out[f * l + pt] = r0[f];
  }
  return 0;
}

}  // namespace

extern "C" {

void emu_set_recip(int mode) { g_recip = mode; }

// the device grid generator (range_point, stencil.cuh): out[r * n + i] = range(a[r], b[r], length = n)[i]
void emu_range_rows(const double* a, const double* b, int rows, int n, double* out) {
  for (int r = 0; r < rows; ++r)
    for (int i = 0; i < n; ++i) out[static_cast<long long>(r) * n + i] = range_point(a[r], b[r], n, i);
}

// FastDiv (division by a grid size in unravel) against `/` for n = 0..count-1 and near 2^31 / 2^32
long long emu_fastdiv_mismatches(unsigned d, unsigned count) {
  FastDiv f;
  f.set(d);
  long long bad = 0;
  for (unsigned n = 0; n < count; ++n) bad += f.div(n) != n / d;
  for (unsigned k = 0; k < 4096; ++k) {
    const unsigned a = 0x7fffffffu - k, b = 0xffffffffu - k, c = 0x80000000u + k;
    bad += (f.div(a) != a / d) + (f.div(b) != b / d) + (f.div(c) != c / d);
  }
  return bad;
}

int emu_model_count() {
  int n = 0;
#define X(M) ++n;
  ECONPDES_FOR_EACH_MODEL(X)
#undef X
  return n;
}

int emu_model_id(const char* name) {
  int id = 0;
#define X(M) if (std::strcmp(name, M::name()) == 0) return id; ++id;
  ECONPDES_FOR_EACH_MODEL(X)
#undef X
  return -1;
}

int emu_residual(int model_id, const long long* S, const double* const* grids, const double* prm, const double* y,
                 const double* bc, double* out, double* saved) {
  int id = 0;
#define X(M)                                                                        \
  if (id++ == model_id)                                                             \
    return g_recip == 0 ? residual<M, 0>(S, grids, prm, y, bc, out, saved)          \
         : g_recip == 1 ? residual<M, 1>(S, grids, prm, y, bc, out, saved)          \
                        : residual<M, 2>(S, grids, prm, y, bc, out, saved);
  ECONPDES_FOR_EACH_MODEL(X)
#undef X
  return -1;
}

// apply_row_terms (G_helper! and the min-max MCP residual, finiteschemesolve.jl:158-162, 197-203) row by row.
// mode 0: IEEE division by delta; 1: reciprocal formed per row when delta is finite (rdelta), else division;
// 2: reciprocal always (1/Inf = 0 included), the batched kernel's branch-free form.
void emu_row_terms(long long n, const double* g, const double* y, const double* ypost, const double* delta,
                   const unsigned char* is_alg, const double* lb, const double* ub, int has_bounds, int mode, double* out) {
  for (long long i = 0; i < n; ++i) {
    RowTerms t;
    t.implicit = ypost != nullptr;
    t.has_bounds = has_bounds != 0;
    t.ypost = ypost ? ypost[i] : 0.0;
    t.delta = delta[i];
    t.is_alg = is_alg && is_alg[i];
    t.lb = lb ? lb[i] : 0.0;
    t.ub = ub ? ub[i] : 0.0;
    if (mode == 1 && std::isfinite(delta[i])) t.rdelta = 1.0 / delta[i];
    if (mode == 2) { t.rdelta = 1.0 / delta[i]; t.rdelta_always = true; }
    out[i] = apply_row_terms(g[i], y[i], t);
  }
}

// box != 0: full 3^N stencil planes; box == 0: star planes (1 + 2N)
int emu_jacobian(int model_id, int box, const long long* S, const double* const* grids, const double* prm,
                 const double* y, const double* bc, const double* ypost, double delta, const double* eps_table,
                 int eps_per_column, double* planes, double* r0) {
  int id = 0;
#define X(M)                                                                                                  \
  if (id++ == model_id)                                                                                       \
    return box ? (g_recip == 0 ? jacobian<M, true, 0>(S, grids, prm, y, bc, ypost, delta, eps_table, eps_per_column, planes, r0) \
                               : jacobian<M, true, 2>(S, grids, prm, y, bc, ypost, delta, eps_table, eps_per_column, planes, r0)) \
               : (g_recip == 0 ? jacobian<M, false, 0>(S, grids, prm, y, bc, ypost, delta, eps_table, eps_per_column, planes, r0) \
                               : jacobian<M, false, 2>(S, grids, prm, y, bc, ypost, delta, eps_table, eps_per_column, planes, r0));
  ECONPDES_FOR_EACH_MODEL(X)
#undef X
  return -1;
}
}
