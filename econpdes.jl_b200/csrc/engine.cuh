// Host-side structures of the generic engine: device-resident problem, the per-model
// kernel dispatch table, the linear-solver interface and the pseudo-transient driver.
#pragma once
#include <memory>
#include <vector>

#include "../../include/econpdes_cuda.h"
#include "common.cuh"
#include "dist.cuh"
#include "kernels.cuh"
#include "sweep.cuh"

namespace econpdes {

// Debug facility standing in for compute-sanitizer's memcheck (the tool is not available on every pool): with
// ECONPDES_GUARD=1 every device buffer of the library is allocated with a 256-byte canary region on either side
// (0xA5 bytes); the regions are verified when the buffer is released and on econpdes_guard_violations(), which
// returns the number of corrupted canary bytes seen so far -- an out-of-bounds WRITE next to any buffer shows up
// there.  Off by default (no cost).
struct GuardRegistry {
  static constexpr size_t kBytes = 256;
  static bool enabled();
  static void add(void* base, size_t total_bytes);
  static void remove(void* base);            // verifies the buffer's canaries, then forgets it
  static long long violations(bool check_live);
};

template <class T>
struct DevBuf {
  T* p = nullptr;  // first logical element (pad elements of ghost storage precede and follow it)
  size_t n = 0;    // logical element count
  size_t cap = 0;  // allocated logical element count (grow-only: work buffers are re-used across solves)
  size_t pad = 0;  // ghost elements on either side (slab-decomposed fields: one plane); set before alloc
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      T* base = p - pad_alloc_ - guard_alloc_;
      if (guard_alloc_) GuardRegistry::remove(base);
      cudaFree(base);
    }
    p = nullptr;
    n = 0;
    cap = 0;
  }
  void alloc(size_t count) {
    if (count <= cap && p && pad == pad_alloc_) {
      n = count;
      return;
    }
    release();
    if (count == 0) return;
    const size_t guard = GuardRegistry::enabled() ? (GuardRegistry::kBytes + sizeof(T) - 1) / sizeof(T) : 0;
    T* base = nullptr;
    const size_t total = count + 2 * pad + 2 * guard;
    EP_CUDA(cudaMalloc(&base, total * sizeof(T)));
    if (pad) {   // ghosts at the global ends stay 0
      EP_CUDA(cudaMemset(base + guard, 0, (count + 2 * pad) * sizeof(T)));
      // the memset runs on the legacy stream, which the library's non-blocking streams do not wait for
      EP_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    }
    if (guard) {
      EP_CUDA(cudaMemset(base, 0xA5, guard * sizeof(T)));
      EP_CUDA(cudaMemset(base + guard + count + 2 * pad, 0xA5, guard * sizeof(T)));
      EP_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
      GuardRegistry::add(base, total * sizeof(T));
    }
    p = base + guard + pad;
    pad_alloc_ = pad;
    guard_alloc_ = guard;
    n = count;
    cap = count;
  }
  void zero(cudaStream_t s) { if (p) EP_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count, cudaStream_t s) {
    alloc(count);
    EP_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void download(T* h, size_t count, cudaStream_t s) const {
    EP_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }

 private:
  size_t pad_alloc_ = 0, guard_alloc_ = 0;
};

extern std::atomic<int> g_batch_points_per_thread;   // 0 = automatic (api: econpdes_set_batch_points_per_thread)

// Problem resident on one device.
struct Ctx {
  int model_id = -1, N = 0, F = 0, nparams = 0, nsaved = 0;
  bool box = false;        // stencil shape of the Jacobian planes
  unsigned plane_mask = 0xffffffffu;   // box stencils of one unknown: planes the model's residual can depend on
  int nk = 0;              // planes per (fin, fout) pair
  long long S[kMaxDims] = {1, 1, 1, 1};  // GLOBAL grid sizes
  long long l = 0, n = 0;             // points / unknowns owned by this device
  long long n_glob = 0;               // unknowns of the whole grid (norms divide by it)
  int ncolors = 0;
  // slab decomposition along dimension N-1 (dist == nullptr: the whole grid lives here)
  Dist* dist = nullptr;
  long long slab_first = 0, slab_count = 0;   // owned range of dimension N-1
  long long halo = 0;                          // ghost plane size (prod S_0..S_{N-2}) when dist != nullptr
  DevBuf<double> gather;                       // [world][kGatherMax] landing area of small all-gathers
  double* h_gather = nullptr;                  // pinned copy
  ParamPack prm{};
  DevBuf<double> grid[kMaxDims], geomtab[kMaxDims], guess, bc, lb, ub, delta_rel;
  DevBuf<unsigned char> is_alg;
  bool has_bounds = false, has_delta_vec = false, all_delta_inf = false;
  bool geom_tame = false;   // grid reciprocals and boundary values finite and of moderate magnitude (sweep.cuh)
  double delta_base = 1.0;  // min finite |Δ_i| of a vector Δ (finiteschemesolve.jl:394)
  cudaStream_t stream = nullptr;
  // reductions
  DevBuf<double> norm_part, norm_out, color_part, color_eps, color_sumsq;
  DevBuf<unsigned> counters;
  double* h_norm = nullptr;  // pinned [4]
  int max_blocks = 0;
  NormWork nw() const { return NormWork{norm_part.p, counters.p, norm_out.p, max_blocks}; }
  ColorWork cw() const { return ColorWork{color_part.p, counters.p + 1, color_eps.p, color_sumsq.p}; }
  template <int NN>
  GridView<NN> grid_view() const {
    GridView<NN> gv;
    gv.set_sizes(S);
    for (int d = 0; d < NN; ++d) {
      gv.x[d] = grid[d].p;
      gv.tab[d] = geomtab[d].p;
    }
    gv.l = l;
    gv.off = slab_first;
    return gv;
  }
  void exchange_halo(const double* field) const {
    if (dist != nullptr) dist->exchange_halo(const_cast<double*>(field), halo, slab_count, stream);
  }
  // Per-rank partials (device, `count` doubles) -> every rank's copy on the host, rank-major.
  // Returns the pinned landing area; valid until the next call.
  const double* gather_to_host(const double* partial_dev, int count);
  ~Ctx();
};

struct Norms {
  double sumsq, maxabs;
  bool any_nan;
  bool untame = false;   // some unknown of the sweep's input is not tame (sweep.cuh)
  double norm2_over_n(long long n) const;  // NaN if any NaN, like norm(v)/length(v)
};

// Per-model kernel launchers (instantiated from the functor registry).
struct ModelOps {
  virtual ~ModelOps() = default;
  virtual const char* name() const = 0;
  virtual int N() const = 0;
  virtual int F() const = 0;
  virtual int nparams() const = 0;
  virtual void derive(double* prm) const = 0;
  virtual int time_slot() const = 0;            // parameter slot holding the time of pde(state, u, t), or -1   // appends the model's parameter-only constants (models.cuh)
  virtual int nsaved() const = 0;
  virtual bool box() const = 0;
  virtual unsigned plane_mask() const = 0;      // bit k: the residual reads box slot k (M::slot_used)
  // tame: the caller vouches that every unknown the sweep reads is tame (sweep.cuh); the residual sweep
  // itself reports untame unknowns in bit 1 of its flag word, after which its output must be recomputed
  // with tame = false
  virtual void residual(const Ctx& c, const double* y, const RowSpec& rs, double* out, bool want_norm,
                        bool tame = false) const = 0;
  virtual void jacobian(const Ctx& c, const double* y, const RowSpec& rs, int eps_policy, double* planes,
                        double* r0, bool tame = false) const = 0;
  virtual void saved(const Ctx& c, const double* y, double* out) const = 0;
  // fused batched engine; returns false if the model is not 1-D single-unknown
  virtual bool batched(const struct BatchArgs& a, cudaStream_t s) const = 0;
  // single 1-D single-unknown problem (device pointers): the Jacobian diagnostics of k_diag_1d and the saved
  // outputs, without a Ctx; return false for other models
  virtual bool diag_1d(int S, const double* prm, const double* x, const double* y, double delta, double tol, int signs,
                       int stencils, int* counts, unsigned long long* maxbits, cudaStream_t s) const = 0;
  virtual bool saved_1d(int S, const double* prm, const double* x, const double* y, double* out, cudaStream_t s) const = 0;
};
const ModelOps* model_ops(int model_id);
int model_count();
// slices of the registry (modelops.cu, compiled once per slice)
constexpr int kModelOpsParts = 8;
const ModelOps* model_ops_part0(int), *model_ops_part1(int), *model_ops_part2(int), *model_ops_part3(int),
    *model_ops_part4(int), *model_ops_part5(int), *model_ops_part6(int), *model_ops_part7(int);

// Linear solver behind each Newton iteration: J x = b with J in stencil planes.
struct LinearSolver {
  virtual ~LinearSolver() = default;
  // returns 0 on success, 1 if the solve failed (singular / non-finite / not converged)
  virtual int solve(const double* planes, const double* b, double* x, const econpdes_options& opt, int* iters,
                    double* relres) = 0;
  // a new Newton solve starts (new time step, new ypost): solvers that keep state across calls (the
  // multigrid hierarchy) refresh it at the next solve
  virtual void new_step() {}
  // device ms per launch of the solver's building blocks (bench.py): [0] stencil SpMV,
  // [1] one preconditioner application (multigrid V-cycle), [2] hierarchy set-up
  virtual void bench(const double* planes, const double* v, int reps, double* flush, size_t flush_bytes,
                     double* ms_out) {
    (void)planes; (void)v; (void)reps; (void)flush; (void)flush_bytes;
    ms_out[0] = ms_out[1] = ms_out[2] = 0.0;
  }
};
std::unique_ptr<LinearSolver> make_tridiag_solver(Ctx& c);
std::unique_ptr<LinearSolver> make_krylov_solver(Ctx& c);  // krylov.cu

// work vectors of the regularised normal-equations solve (normaleq.cuh), allocated on first use
struct NormalEqWork {
  DevBuf<double> b, r, p, q, w, partial, scalar;
  DevBuf<unsigned long long> bits;
};

struct NewtonOutcome {
  double nl;        // ||R(zero)||_2/n, Inf if singular, NaN if NaN
  int iters;
  long long lin_iters;
  bool singular;
  int singular_retries = 0;   // iterations that took NLsolve's regularised-normal-equations direction
  int lin_solves = 0, lin_loose = 0;   // linear solves, and those that ended with a relative residual above 1e-11
  double lin_max_relres = 0.0;
};

class Solver {
 public:
  explicit Solver(std::unique_ptr<Ctx> ctx);
  Ctx& ctx() { return *c_; }
  Norms eval_G(const double* y, double* out);  // G!(ydot, y) + norms
  Norms eval_R(const double* y, const double* ypost, double delta, double* out);
  void jacobian(const double* y, const double* ypost, double delta, int eps_policy, bool tame = false);
  bool check_tame(const double* y);
  double* planes() { return planes_.p; }
  LinearSolver& linear() { return *lin_; }
  NewtonOutcome newton(const double* ypost, double delta, double ftol, int maxit, const econpdes_options& opt,
                       double* y_out);
  int solve(const econpdes_options& opt, econpdes_result* res);
  // _solve_backward (src/pdesolve.jl:158-224); host output arrays as in econpdes_solve_backward
  int solve_backward(const econpdes_options& opt, const double* taus, int ntaus, double* solution, double* saved,
                     double* residual_norms, int32_t* newton_iters, double* device_ms, int64_t* launches);
  void csc_gather(double* nzval_dev);
  DevBuf<double> ypost_, y_, ynew_, R_, s_, G_, planes_;
  long long nnz_ = 0;
  // src/monotonicity.jl diagnostics of the first assembled Jacobian (see econpdes_result)
  int sign_flags_ = 0, nonmonotone_entries_ = 0;
  double nonmonotone_max_ = 0.0;
  bool sign_checked_ = false, stencils_checked_ = false;

 private:
  RowSpec row_spec(const double* ypost, double delta) const;
  Norms read_norms();
  Norms residual_with_norms(const double* y, const RowSpec& rs, double* out);
  bool last_tame_ = false;
  // NLsolve's regularised-normal-equations direction for a singular Jacobian (normaleq.cuh): s_ from planes_ and R_
  int regularized_step();
  NormalEqWork neq_;
  double masked_norm(const double* ydot, const double* y);
  void jacobian_diagnostics(double delta, const econpdes_options& opt);
  std::unique_ptr<Ctx> c_;
  const ModelOps* ops_;
  std::unique_ptr<LinearSolver> lin_;
  DevBuf<long long> csc_map_;
};

constexpr int kGatherMax = 128;  // doubles per rank in one small all-gather

// validates + uploads; throws.  With `dist`, the problem arrays are this rank's slab
// (S_0..S_{N-2}, slab_count) while problem->size and problem->grid describe the GLOBAL grid.
std::unique_ptr<Ctx> make_ctx(const econpdes_problem* p, Dist* dist = nullptr, long long slab_first = 0,
                              long long slab_count = 0);
void ensure_device();

}  // namespace econpdes
