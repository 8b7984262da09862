// Per-model kernel launchers: one ModelOpsT<M> per registered functor (models.cuh), holding the
// instantiations of the residual / Jacobian / saved-output sweeps and of the fused batched 1-D solve.
// This file is compiled kModelOpsParts times with -DEP_PART=k; slice k instantiates the functors whose
// registry index is k modulo kModelOpsParts (the build runs the slices in parallel).
#include <cstdlib>

#include "batched1d.cuh"
#include "engine.cuh"
#include "sweep.cuh"

#ifndef EP_PART
#error "compile with -DEP_PART=<slice>"
#endif

namespace econpdes {
namespace {

template <class M, bool ONE_D = (M::N == 1 && M::F == 1)>
struct BatchedLaunch {
  static bool run(const BatchArgs&, cudaStream_t) { return false; }
};
template <class M>
struct BatchedLaunch<M, true> {
  template <int P, bool PER_COL, bool COMPACT>
  static void launch(const BatchArgs& a, cudaStream_t s) {
    const int T = (((a.S + P - 1) / P + 31) / 32) * 32;
    const size_t smem = batched_smem_bytes(T, P);
    // per device (the attribute lives in the device's context): set once for every device this process uses
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    EP_CUDA(cudaGetDevice(&dev));
    if (!((configured.load() >> (dev & 63)) & 1ull)) {
      EP_CUDA(cudaFuncSetAttribute((k_batched_pdesolve_1d<M, P, PER_COL, COMPACT>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(batched_smem_bytes(batched_max_threads(P), P))));
      configured.fetch_or(1ull << (dev & 63));
    }
    EP_LAUNCH((k_batched_pdesolve_1d<M, P, PER_COL, COMPACT>), static_cast<unsigned>(a.nproblems - a.problem0), T, smem, s, a);
  }
  template <int P>
  static void launch(const BatchArgs& a, cudaStream_t s) {
    const bool compact = a.grids == nullptr || a.guess == nullptr;
    if (a.opt.eps_policy == ECONPDES_EPS_PER_COLUMN) {
      if (compact) launch<P, true, true>(a, s); else launch<P, true, false>(a, s);
    } else {
      if (compact) launch<P, false, true>(a, s); else launch<P, false, false>(a, s);
    }
  }
  static bool run(const BatchArgs& a, cudaStream_t s) {
    if (a.S > 1024) return false;
    // grid points per thread (batched1d.cuh): econpdes_set_batch_points_per_thread, else the
    // ECONPDES_BATCH_P environment variable, else by problem size
    static const int env = [] {
      const char* e = std::getenv("ECONPDES_BATCH_P");
      return e ? std::atoi(e) : 0;
    }();
    const int knob = g_batch_points_per_thread.load();
    const int forced = knob ? knob : env;
    // measured on B200 (tools/bench_batched_sizes.py, 16 384 Campbell-Cochrane problems; ms for P = 1 / 2 / 3 / 4):
    // 58 points 3.8 / 3.6 / 4.0 / 4.9, 98 points 3.1 / 2.6 / 3.1 / 3.5, 198 points 4.2 / 3.7 / 3.6 / 3.5,
    // 298 points 5.4 / 4.3 / 4.1 / 5.0, 498 points 8.0 / 5.6 / 5.6 / 5.2, 998 points 13.4 / 8.4 / 7.5 / 6.6
    int P = forced ? forced : (a.S > 384 ? 4 : (a.S > 128 ? 3 : 2));
    while (P > 1 && (a.S + P - 1) / P > batched_max_threads(P)) --P;
    if (P >= 4) launch<4>(a, s);
    else if (P == 3) launch<3>(a, s);
    else if (P >= 2) launch<2>(a, s);
    else launch<1>(a, s);
    return true;
  }
};

// Slab path: every rank holds every rank's per-color sums of squares (rank-major); fold them in
// rank order and turn the totals into FiniteDiff's steps.
__global__ void k_color_eps_fold(int world, int nc, const double* __restrict__ all, double* __restrict__ sumsq,
                                 double* __restrict__ eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  double t = 0.0;
  for (int r = 0; r < world; ++r) t += all[r * nc + c];
  sumsq[c] = t;
  const double e = color_epsilon_from_sumsq(t);
  eps[c] = e;
  eps[nc + c] = 1.0 / e;
}

template <class M>
struct ModelOpsT final : ModelOps {
  const char* name() const override { return M::name(); }
  int N() const override { return M::N; }
  int F() const override { return M::F; }
  int nparams() const override { return M::NPARAMS; }
  void derive(double* prm) const override {
    static_assert(M::NPARAMS + M::NDERIVED + (M::TIME_SLOT >= 0 ? 1 : 0) <= kMaxParams, "parameter pack too small");
    M::derive(prm);
  }
  int time_slot() const override { return M::TIME_SLOT; }
  int nsaved() const override { return M::NSAVED; }
  bool box() const override { return M::USES_CROSS && M::N > 1; }
  unsigned plane_mask() const override {
    unsigned m = 0;
    for (int k = 0; k < ipow3(M::N); ++k) {
      bool axial = true;   // centre and axial neighbours are always kept (the line smoothers solve along them)
      int nz = 0;
      for (int d = 0; d < M::N; ++d) nz += ((k / p3(d)) % 3 != 1) ? 1 : 0;
      axial = nz <= 1;
      if (axial || M::slot_used(k)) m |= 1u << k;
    }
    return m;
  }
  // ---- marching sweeps (sweep.cuh) for 2-D / 3-D grids; ECONPDES_MARCH=0 restores the one-thread-per-point kernels
  static int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
  }
  // the marching sweeps keep three planes of every unknown's box in registers: not for wide systems
  // (8 unknowns on a 3-D box = 216 values), which take the one-thread-per-point kernels
  static constexpr bool kMarchable = M::N >= 2 && M::N <= 3 && M::F * ipow3(M::N) <= 108;
  static bool march_on() {
    static const bool on = env_int("ECONPDES_MARCH", 1) != 0;
    return on && kMarchable;
  }
  template <int ROWKIND, int TAME>
  void residual_march(const Ctx& c, const double* y, const RowSpec& rs, double* out, bool want_norm) const {
    if constexpr (kMarchable) {
      constexpr int RB = 128, MINB = 4, UNR = M::N >= 3 ? 1 : 3;
      static const int force_L = env_int("ECONPDES_MARCH_L", 0);
      const MarchPlan mp = make_march_plan(c.l / c.slab_count, c.slab_count, RB, 148 * MINB, force_L);
      const int ntiles = mp.ncolblocks * mp.nseg;
      const int blocks = ntiles < 148 * MINB ? ntiles : 148 * MINB;
      EP_LAUNCH((k_residual_march<M, ROWKIND, TAME, UNR, RB, MINB>), blocks, RB, 0, c.stream, c.grid_view<M::N>(), c.prm,
                y, c.bc.p, rs, out, c.nw(), want_norm ? 1 : 0, mp);
    }
  }
  void residual(const Ctx& c, const double* y, const RowSpec& rs, double* out, bool want_norm,
                bool tame) const override {
    if (march_on()) {
      const bool simple = rs.ypost != nullptr && rs.delta_rel == nullptr && rs.is_alg == nullptr && !rs.has_bounds &&
                          rs.rdelta != 0.0;
      // tame forms: with boundary derivative values (1) or the default all-zero ones (2)
      if (!(tame && c.geom_tame)) residual_march<2, 0>(c, y, rs, out, want_norm);
      else if (c.bc.p != nullptr) residual_march<2, 1>(c, y, rs, out, want_norm);
      else if (rs.ypost == nullptr) residual_march<0, 2>(c, y, rs, out, want_norm);
      else if (simple) residual_march<1, 2>(c, y, rs, out, want_norm);
      else residual_march<2, 2>(c, y, rs, out, want_norm);
      return;
    }
    // 128-thread blocks, four per SM (128 registers): finer scheduling granularity than 256 x 2 at the
    // same occupancy (measured 3-8 % faster on the 2048^2 and 256^3 sweeps)
    constexpr int RB = 128;
    EP_LAUNCH((k_residual<M, RB, 4>), sweep_blocks(c.l, RB), RB, 0, c.stream, c.grid_view<M::N>(), c.prm, y, c.bc.p, rs,
              out, c.nw(), want_norm ? 1 : 0);
  }
  template <bool BOX, int JMINB>
  void jacobian_march(const Ctx& c, const double* y, const RowSpec& rs, int per_col, double* planes, double* r0,
                      bool tame) const {
    if constexpr (kMarchable) {
      constexpr int JB = 128;
      static const int force_L = env_int("ECONPDES_MARCH_LJ", 0);
      const MarchPlan mp = make_march_plan(c.l / c.slab_count, c.slab_count, JB, 148 * JMINB, force_L);
      const int ntiles = mp.ncolblocks * mp.nseg;
      const int blocks = ntiles < 148 * JMINB ? ntiles : 148 * JMINB;
      auto gom = [&](auto kernel) {
        EP_LAUNCH(kernel, blocks, JB, 0, c.stream, c.grid_view<M::N>(), c.prm, y, c.bc.p, rs, c.color_eps.p, planes, r0,
                  mp);
      };
      const bool simple = rs.ypost != nullptr && rs.delta_rel == nullptr && rs.is_alg == nullptr && !rs.has_bounds &&
                          rs.rdelta != 0.0;
      if (per_col) gom(k_jacobian_march<M, BOX, 1, 0, 2, JB, JMINB>);
      else if (!(tame && c.geom_tame)) gom(k_jacobian_march<M, BOX, 0, 0, 2, JB, JMINB>);
      else if (c.bc.p != nullptr) gom(k_jacobian_march<M, BOX, 0, 1, 2, JB, JMINB>);
      else if (rs.ypost == nullptr) gom(k_jacobian_march<M, BOX, 0, 2, 0, JB, JMINB>);
      else if (simple) gom(k_jacobian_march<M, BOX, 0, 2, 1, JB, JMINB>);
      else gom(k_jacobian_march<M, BOX, 0, 2, 2, JB, JMINB>);
    }
  }
  void jacobian(const Ctx& c, const double* y, const RowSpec& rs, int eps_policy, double* planes,
                double* r0, bool tame) const override {
    const int per_col = eps_policy == ECONPDES_EPS_PER_COLUMN;
    if (!per_col) {
      const int cblocks = blocks_for(c.l) < 148 * 4 ? blocks_for(c.l) : 148 * 4;   // persistent: 4 blocks per SM
      EP_LAUNCH(k_color_sumsq<M::N>, cblocks, kBlock, 0, c.stream, c.grid_view<M::N>(), M::F, y, c.cw());
      if (c.dist != nullptr) {   // one all-gather of the C masked norms per Jacobian
        c.dist->all_gather(c.color_sumsq.p, c.gather.p, c.ncolors, c.stream);
        EP_LAUNCH(k_color_eps_fold, 1, 128, 0, c.stream, c.dist->world(), c.ncolors, c.gather.p, c.color_sumsq.p,
                  c.color_eps.p);
      }
    }
    constexpr bool BOX = M::USES_CROSS && M::N > 1;
    if constexpr (kMarchable) {
      if (march_on()) {
        // 128-thread blocks, three per SM (168 registers): 2048^2 0.258 ms against 0.271 with four blocks of
        // 128 registers (spills); 256^3 1.59 ms against 2.19
        jacobian_march<BOX, 3>(c, y, rs, per_col, planes, r0, tame);
        return;
      }
    }
    // step policy fixed at compile time (branch-free row evaluation), two resident blocks per SM:
    // measured 0.53 -> 0.35 ms at 2048^2 and 2.94 -> 2.47 ms at 256^3 against the run-time-policy kernel
    // 3-D rows (10+ evaluations of a 27-point box) get 170 registers: 128-thread blocks, three per SM
    constexpr int JB = M::N >= 3 ? 128 : kBlock, JMINB = M::N >= 3 ? 3 : 2;
    auto go = [&](auto kernel) {
      EP_LAUNCH(kernel, blocks_for(c.l, JB), JB, 0, c.stream, c.grid_view<M::N>(), c.prm, y, c.bc.p, rs, c.color_eps.p,
                per_col, planes, r0);
    };
    if (per_col) go(k_jacobian<M, BOX, 1, JB, JMINB>);
    else go(k_jacobian<M, BOX, 0, JB, JMINB>);
  }
  void saved(const Ctx& c, const double* y, double* out) const override {
    if (M::NSAVED == 0) return;
    EP_LAUNCH(k_saved<M>, blocks_for(c.l), kBlock, 0, c.stream, c.grid_view<M::N>(), c.prm, y, c.bc.p, out);
  }
  bool batched(const BatchArgs& a, cudaStream_t s) const override { return BatchedLaunch<M>::run(a, s); }
  static ParamPack pack_params(const double* prm) {
    ParamPack pk{};
    for (int i = 0; i < M::NPARAMS; ++i) pk.v[i] = prm[i];
    M::derive(pk.v);
    return pk;
  }
  bool diag_1d(int S, const double* prm, const double* x, const double* y, double delta, double tol, int signs,
               int stencils, int* counts, unsigned long long* maxbits, cudaStream_t s) const override {
    if constexpr (M::N == 1 && M::F == 1) {
      const int T = ((S + 31) / 32) * 32;
      EP_LAUNCH(k_diag_1d<M>, 1, T, 0, s, S, pack_params(prm), x, y, delta, tol, signs, stencils, counts, maxbits);
      return true;
    } else {
      return false;
    }
  }
  bool saved_1d(int S, const double* prm, const double* x, const double* y, double* out, cudaStream_t s) const override {
    if constexpr (M::N == 1 && M::F == 1) {
      if (M::NSAVED == 0) return true;
      GridView<1> gv;
      const long long sizes[1] = {S};
      gv.set_sizes(sizes);
      gv.x[0] = x;
      EP_LAUNCH(k_saved<M>, blocks_for(S), kBlock, 0, s, gv, pack_params(prm), y, static_cast<const double*>(nullptr), out);
      return true;
    } else {
      return false;
    }
  }
};

template <int I>
const ModelOps* pick(int) { return nullptr; }
template <int I, class M, class... Rest>
const ModelOps* pick(int id) {
  if constexpr (I % kModelOpsParts == EP_PART) {
    if (id == I) {
      static const ModelOpsT<M> ops;
      return &ops;
    }
  }
  return pick<I + 1, Rest...>(id);
}
struct Dummy_ {};
template <class D, class... Ms>
const ModelOps* pick_all(int id) { return pick<0, Ms...>(id); }

}  // namespace

#define X(M) , M
#define EP_CAT2(a, b) a##b
#define EP_CAT(a, b) EP_CAT2(a, b)
const ModelOps* EP_CAT(model_ops_part, EP_PART)(int model_id) {
  return pick_all<Dummy_ ECONPDES_FOR_EACH_MODEL(X)>(model_id);
}
#undef X

}  // namespace econpdes
