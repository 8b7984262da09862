// extern "C" entry points of libeconpdes_cuda.so (include/econpdes_cuda.h).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

#include "batched1d.cuh"
#include "engine.cuh"

using namespace econpdes;

namespace {

template <class Fn>
int guarded(Fn&& fn) {
  try {
    return fn();
  } catch (const std::invalid_argument& e) {
    set_error(e.what());
    return ECONPDES_E_INVALID;
  } catch (const NcclError& e) {
    set_error(e.what());
    return ECONPDES_E_NCCL;
  } catch (const CudaError& e) {
    set_error(e.what());
    const std::string m = e.what();
    return m.find("no usable CUDA device") != std::string::npos ? ECONPDES_E_NO_DEVICE : ECONPDES_E_CUDA;
  } catch (const std::exception& e) {
    set_error(e.what());
    return ECONPDES_E_UNSUPPORTED;
  }
}

}  // namespace

namespace {
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long& st) {
  unsigned long long z = (st += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ bool same_bits(double x, double y) {
  return (x != x && y != y) || __double_as_longlong(x) == __double_as_longlong(y);
}
__global__ void k_selftest_reciprocal(unsigned long long seed, long long n, unsigned long long* mismatch) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  unsigned long long m0 = 0, m1 = 0, m2 = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned long long st = seed + 0x632BE59BD9B4E019ull * static_cast<unsigned long long>(i);
    unsigned long long bb = splitmix64(st);
    const unsigned long long ab = splitmix64(st);
    // every 8th draw: a mantissa of all ones / all zeros / one bit (the hard cases of the correction step)
    if ((i & 7) == 7) bb = (bb & 0xFFF0000000000000ull) | ((i & 8) ? 0x000FFFFFFFFFFFFFull : ((i & 16) ? 1ull : 0ull));
    const double b = __longlong_as_double(static_cast<long long>(bb));
    const double a = __longlong_as_double(static_cast<long long>(ab));
    const double want = 1.0 / b, got = ep_rcp(b);
    const double fb = fabs(b);
    if (fb >= 0x1p1022 && fb < INFINITY) {
      const long long d = __double_as_longlong(got) - __double_as_longlong(want);
      if (d > 1 || d < -1) ++m2;
    } else if (!same_bits(got, want)) {
      ++m0;
    }
    // quotients: a = 0 / Inf / NaN, or a and a / b at least 2^-900 (the correction step a - q0 b must not
    // underflow; below that the subnormal result may be rounded twice)
    const double fa = fabs(a), q = a / b;
    const bool a_ok = !(fa > 0.0) || !(fa < INFINITY) || (fa >= 0x1p-900 && fabs(q) >= 0x1p-900);
    if (fb >= 0x1p-1000 && fb <= 0x1p1000 && a_ok && !same_bits(mdiv(a, b, got), q)) ++m1;
    // degenerate divisors (0, +-Inf, NaN) with every kind of dividend
    {
      const int sel = static_cast<int>((ab >> 7) & 3);
      const double sgn = (ab >> 9) & 1 ? -1.0 : 1.0;
      const double bs = sel == 0 ? sgn * 0.0 : (sel == 1 ? sgn * INFINITY : (sel == 2 ? NAN : b));
      const int asel = static_cast<int>((ab >> 10) & 7);
      const double as = asel == 0 ? 0.0 : (asel == 1 ? -0.0 : (asel == 2 ? INFINITY : (asel == 3 ? -INFINITY : (asel == 4 ? NAN : a))));
      if (sel != 3 && !same_bits(mdiv(as, bs, ep_rcp(bs)), as / bs)) ++m1;
    }
  }
  if (m0) atomicAdd(&mismatch[0], m0);
  if (m1) atomicAdd(&mismatch[1], m1);
  if (m2) atomicAdd(&mismatch[2], m2);
}
}  // namespace

struct econpdes_handle {
  std::unique_ptr<Solver> solver;
  DevBuf<double> a, b, c;
  bool have_jacobian = false;
};

struct econpdes_dist {
  std::unique_ptr<Dist> dist;
};

struct econpdes_batch {
  int model_id = -1;
  long long B = 0;
  int S = 0, nparams = 0;
  cudaStream_t stream = nullptr;
  static constexpr int kPipe = 3;
  cudaStream_t pipe[kPipe] = {};   // copy/compute pipeline of the one-call path (econpdes_solve_batched)
  // pinned staging for the per-problem scalars (error, exit reason, counts, residual norm): a pageable
  // target would make every slice's read-back block the host and serialise the pipeline
  int* nflag_host = nullptr;       // page-locked: problems of the last launches that met a singular Jacobian
  int sms = 0;                     // multiprocessors of the batch's device (slice sizes are whole waves)
  bool last_direct = false;        // the last one-call solve stored its results straight into pinned host memory
  int* err_host = nullptr;         // [4][B]: error, exit_reason, n_outer, n_newton
  double* rn_host = nullptr;       // [B]
  DevBuf<double> params, grids, guess, solution, residual_norm;
  // compact inputs (econpdes_batch_inputs): grids built on the device, broadcast guess
  DevBuf<double> range_lo, points;
  int n_range = 0;
  bool explicit_grids = true, explicit_guess = true;
  double guess_fill = 0.0;
  // per-problem integers in ONE allocation, [4][B] = error, exit_reason, n_outer, n_newton, so a slice's four
  // ranges come back with a single pitched copy (a small copy costs ~8 us of copy-engine time each way)
  DevBuf<int> ints;
  struct IntView {
    int* p = nullptr;
    void download(int* h, size_t count, cudaStream_t s) const {
      EP_CUDA(cudaMemcpyAsync(h, p, count * sizeof(int), cudaMemcpyDeviceToHost, s));
    }
  } exit_reason, n_outer, n_newton, error;
  DevBuf<int> log_len, flags;
  DevBuf<econpdes_log_row> log;
  int log_capacity = 0;
  int64_t launches = 0;
  int device = -1;
  ~econpdes_batch() {
    if (stream) cudaStreamDestroy(stream);
    for (cudaStream_t p : pipe)
      if (p) cudaStreamDestroy(p);
    if (nflag_host) cudaFreeHost(nflag_host);
    if (err_host) cudaFreeHost(err_host);
    if (rn_host) cudaFreeHost(rn_host);
  }
};

extern "C" {

int econpdes_model_count(void) { return model_count(); }

int econpdes_model_id(const char* name) {
  if (name == nullptr) return ECONPDES_E_INVALID;
  for (int i = 0; i < model_count(); ++i)
    if (std::strcmp(model_ops(i)->name(), name) == 0) return i;
  set_error(std::string("unknown model `") + name + "`");
  return ECONPDES_E_INVALID;
}

const char* econpdes_model_name(int model_id) {
  const ModelOps* o = model_ops(model_id);
  return o ? o->name() : nullptr;
}

int econpdes_model_info(int model_id, int32_t* ndims, int32_t* nfuncs, int32_t* nparams, int32_t* nsaved,
                        int32_t* box_stencil) {
  const ModelOps* o = model_ops(model_id);
  if (!o) {
    set_error("unknown model id");
    return ECONPDES_E_INVALID;
  }
  if (ndims) *ndims = o->N();
  if (nfuncs) *nfuncs = o->F();
  if (nparams) *nparams = o->nparams();
  if (nsaved) *nsaved = o->nsaved();
  if (box_stencil) *box_stencil = o->box() ? 1 : 0;
  return ECONPDES_OK;
}

void econpdes_default_options(econpdes_options* o) {
  if (!o) return;
  const double sqrt_eps = 1.4901161193847656e-8;
  o->abstol = sqrt_eps;
  o->inner_abstol = sqrt_eps;
  o->delta = 1.0;
  o->scale = 10.0;
  o->min_delta = 1e-9;
  o->max_delta = std::numeric_limits<double>::infinity();
  o->max_residual_growth = 10.0;
  o->maxiters = 100;
  o->inner_maxiters = 10;
  o->eps_policy = ECONPDES_EPS_FINITEDIFF_COLOR_NORM;
  o->linear_solver = ECONPDES_LIN_AUTO;
  o->lin_rtol = 1e-12;
  o->lin_maxiters = 2000;
  o->lin_restart = 40;
  o->check_monotonicity = 0;
  o->monotonicity_tol = 1e-6;
}

int econpdes_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int econpdes_set_device(int device) {
  return guarded([&] {
    ensure_device();
    EP_CUDA(cudaSetDevice(device));
    return ECONPDES_OK;
  });
}

int64_t econpdes_total_kernel_launches(void) { return g_kernel_launches.load(); }

int64_t econpdes_guard_violations(void) { return GuardRegistry::enabled() ? GuardRegistry::violations(true) : -1; }

int econpdes_create(const econpdes_problem* problem, econpdes_handle** out) {
  return guarded([&] {
    if (!out) throw std::invalid_argument("null handle pointer");
    auto h = std::make_unique<econpdes_handle>();
    h->solver = std::make_unique<Solver>(make_ctx(problem));
    *out = h.release();
    return ECONPDES_OK;
  });
}

void econpdes_destroy(econpdes_handle* h) { delete h; }

int econpdes_residual(econpdes_handle* h, const double* y, double* ydot) {
  return guarded([&] {
    if (!h || !y || !ydot) throw std::invalid_argument("null argument");
    Ctx& c = h->solver->ctx();
    h->a.upload(y, static_cast<size_t>(c.n), c.stream);
    h->b.alloc(static_cast<size_t>(c.n));
    h->solver->eval_G(h->a.p, h->b.p);
    h->b.download(ydot, static_cast<size_t>(c.n), c.stream);
    EP_CUDA(cudaStreamSynchronize(c.stream));
    return ECONPDES_OK;
  });
}

int econpdes_saved(econpdes_handle* h, const double* y, double* saved) {
  return guarded([&] {
    if (!h || !y || !saved) throw std::invalid_argument("null argument");
    Ctx& c = h->solver->ctx();
    if (c.nsaved == 0) return ECONPDES_OK;
    h->a.upload(y, static_cast<size_t>(c.n), c.stream);
    h->b.alloc(static_cast<size_t>(c.nsaved) * c.l);
    model_ops(c.model_id)->saved(c, h->a.p, h->b.p);
    h->b.download(saved, static_cast<size_t>(c.nsaved) * c.l, c.stream);
    EP_CUDA(cudaStreamSynchronize(c.stream));
    return ECONPDES_OK;
  });
}

int econpdes_jacobian_csc(econpdes_handle* h, const double* y, const double* ypost, double delta, int32_t eps_policy,
                          double* nzval, double* eps_out) {
  return guarded([&] {
    if (!h || !y || !nzval) throw std::invalid_argument("null argument");
    Ctx& c = h->solver->ctx();
    h->a.upload(y, static_cast<size_t>(c.n), c.stream);
    if (ypost) h->b.upload(ypost, static_cast<size_t>(c.n), c.stream);
    h->solver->jacobian(h->a.p, ypost ? h->b.p : nullptr, delta, eps_policy, h->solver->check_tame(h->a.p));
    h->have_jacobian = true;
    int64_t size[kMaxDims];
    for (int d = 0; d < c.N; ++d) size[d] = c.S[d];
    const int64_t nnz = econpdes_stencil_nnz(c.N, size, c.F);
    h->c.alloc(static_cast<size_t>(nnz));
    h->solver->csc_gather(h->c.p);
    h->c.download(nzval, static_cast<size_t>(nnz), c.stream);
    if (eps_out && eps_policy == ECONPDES_EPS_FINITEDIFF_COLOR_NORM)
      c.color_eps.download(eps_out, static_cast<size_t>(c.ncolors), c.stream);
    EP_CUDA(cudaStreamSynchronize(c.stream));
    return ECONPDES_OK;
  });
}

int econpdes_linear_solve(econpdes_handle* h, const econpdes_options* options, const double* b, double* x,
                          int32_t* iters_out, double* relres_out) {
  return guarded([&] {
    if (!h || !b || !x || !options) throw std::invalid_argument("null argument");
    if (!h->have_jacobian) throw std::invalid_argument("call econpdes_jacobian_csc before econpdes_linear_solve");
    Ctx& c = h->solver->ctx();
    h->a.upload(b, static_cast<size_t>(c.n), c.stream);
    h->b.alloc(static_cast<size_t>(c.n));
    int it = 0;
    double rr = 0.0;
    const int fail = h->solver->linear().solve(h->solver->planes(), h->a.p, h->b.p, *options, &it, &rr);
    h->b.download(x, static_cast<size_t>(c.n), c.stream);
    EP_CUDA(cudaStreamSynchronize(c.stream));
    if (iters_out) *iters_out = it;
    if (relres_out) *relres_out = rr;
    return fail ? 1 : ECONPDES_OK;
  });
}

int econpdes_implicit_timestep(econpdes_handle* h, const econpdes_options* options, const double* ypost, double delta,
                               int32_t maxiters, double abstol, double* y_out, double* nl_residual,
                               int32_t* newton_iters) {
  return guarded([&] {
    if (!h || !options || !ypost || !y_out) throw std::invalid_argument("null argument");
    Ctx& c = h->solver->ctx();
    h->a.upload(ypost, static_cast<size_t>(c.n), c.stream);
    h->b.alloc(static_cast<size_t>(c.n));
    NewtonOutcome o = h->solver->newton(h->a.p, delta, abstol, maxiters, *options, h->b.p);
    h->b.download(y_out, static_cast<size_t>(c.n), c.stream);
    EP_CUDA(cudaStreamSynchronize(c.stream));
    if (nl_residual) *nl_residual = o.nl;
    if (newton_iters) *newton_iters = o.iters;
    return ECONPDES_OK;
  });
}

int econpdes_bench_kernels(econpdes_handle* h, const double* y, int reps, double* ms_out) {
  return guarded([&] {
    if (!h || !y || !ms_out || reps < 1) throw std::invalid_argument("bad argument");
    Solver& sv = *h->solver;
    Ctx& c = sv.ctx();
    cudaStream_t s = c.stream;
    h->a.upload(y, static_cast<size_t>(c.n), s);
    h->b.alloc(static_cast<size_t>(c.n));
    DevBuf<double> flush;
    const size_t flush_bytes = 512ull << 20;
    flush.alloc(flush_bytes / sizeof(double));
    cudaEvent_t e0, e1;
    EP_CUDA(cudaEventCreate(&e0));
    EP_CUDA(cudaEventCreate(&e1));
    auto timed = [&](auto&& fn) {
      double tot = 0.0;
      for (int r = 0; r < reps + 1; ++r) {
        EP_CUDA(cudaMemsetAsync(flush.p, 0, flush_bytes, s));   // evict L2 between timed launches
        EP_CUDA(cudaEventRecord(e0, s));
        fn();
        EP_CUDA(cudaEventRecord(e1, s));
        EP_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        EP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) tot += ms;
      }
      return tot / reps;
    };
    const ModelOps* ops = model_ops(c.model_id);
    RowSpec g{};   // R = G
    g.ypost = nullptr; g.delta = INFINITY; g.rdelta = 0.0; g.delta_rel = nullptr; g.is_alg = nullptr; g.lb = nullptr; g.ub = nullptr;
    g.has_bounds = 0;
    RowSpec r = g;
    r.ypost = c.guess.p;
    r.delta = 1.0;
    r.rdelta = 1.0;
    const bool tame = sv.check_tame(h->a.p);   // what the solver's own residual sweep establishes
    ms_out[0] = timed([&] { ops->residual(c, h->a.p, g, h->b.p, false, tame); });
    ms_out[1] = timed([&] { ops->residual(c, h->a.p, r, h->b.p, true, tame); });
    ms_out[2] = timed([&] { ops->jacobian(c, h->a.p, r, ECONPDES_EPS_FINITEDIFF_COLOR_NORM, sv.planes(), nullptr, tame); });
    h->have_jacobian = true;
    double lin[3] = {0, 0, 0};
    if (c.N >= 2) sv.linear().bench(sv.planes(), h->a.p, reps, flush.p, flush_bytes, lin);
    ms_out[3] = lin[0];
    ms_out[4] = lin[1];
    ms_out[5] = lin[2];
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ECONPDES_OK;
  });
}

int econpdes_solve(econpdes_handle* h, const econpdes_options* options, econpdes_result* result) {
  return guarded([&] {
    if (!h || !options || !result) throw std::invalid_argument("null argument");
    return h->solver->solve(*options, result);
  });
}

// A single 1-D single-unknown problem with the default boundary rows and no algebraic rows, bounds or per-row
// steps is what the fused batched kernel solves (batched1d.cuh) -- one thread block, the whole pseudo-transient
// loop on chip, ONE launch instead of the ~75 of the generic engine (which is launch-latency bound at this size:
// the README quickstart took 4.5 ms for a few microseconds of arithmetic).  The Jacobian diagnostics of
// src/monotonicity.jl and the saved outputs follow as one small launch each.  ECONPDES_FUSED_1D=0 keeps such
// problems on the generic engine.
static thread_local econpdes_batch* t_single = nullptr;
// Where a launch writes its results when that is not the batch's own device buffers: device-visible aliases of
// pinned HOST memory (the one-call path's zero-copy mode).  Indexed by problem like the device buffers.
struct BatchOut {
  double* solution = nullptr;
  double* residual_norm = nullptr;
  int* ints = nullptr;   // [4][B]: error, exit_reason, n_outer, n_newton
};

static int batch_launch(econpdes_batch* b, const econpdes_options* options, int log_capacity, long long first,
                        long long last, cudaStream_t s, const BatchOut* out = nullptr);

static bool fused_singular_falls_back() {
  static const bool on = [] { const char* e = std::getenv("ECONPDES_SINGULAR_RETRY"); return !(e && e[0] == '0'); }();
  return on;
}

static bool solve_fused_1d(const econpdes_problem* p, const econpdes_options* opt, econpdes_result* res, int* rc_out) {
  static const bool enabled = [] { const char* e = std::getenv("ECONPDES_FUSED_1D"); return !(e && e[0] == '0'); }();
  if (!enabled) return false;
  const ModelOps* ops = model_ops(p->model_id);
  if (ops == nullptr || ops->N() != 1 || ops->F() != 1 || ops->time_slot() >= 0) return false;
  if (p->ndims != 1 || p->nfuncs != 1 || p->nparams != ops->nparams()) return false;   // make_ctx reports these
  const long long S = p->size[0];
  if (S < 2 || S > 1024 || p->grid[0] == nullptr || p->guess == nullptr || p->delta_vec != nullptr) return false;
  if (std::isnan(opt->delta) || opt->delta == 0.0) return false;                        // Solver::solve reports it
  if (p->bc != nullptr)
    for (long long j = 0; j < S; ++j)
      if (p->bc[j] != 0.0) return false;
  if (p->is_algebraic != nullptr)
    for (long long j = 0; j < S; ++j)
      if (p->is_algebraic[j] != 0) return false;
  if (p->lower_bound != nullptr)
    for (long long j = 0; j < S; ++j)
      if (p->lower_bound[j] != -INFINITY) return false;
  if (p->upper_bound != nullptr)
    for (long long j = 0; j < S; ++j)
      if (p->upper_bound[j] != INFINITY) return false;
  ensure_device();
  int device = -1;
  cudaGetDevice(&device);
  econpdes_batch* b = t_single;
  if (b == nullptr || b->model_id != p->model_id || b->S != S || b->device != device) {
    delete t_single;
    t_single = nullptr;
    if (econpdes_batch_create(p->model_id, 1, S, &b) != ECONPDES_OK) throw std::runtime_error(econpdes_last_error());
    b->device = device;
    t_single = b;
  }
  cudaStream_t s = b->stream;
  const int64_t l0 = g_kernel_launches.load();
  if (b->nparams > 0) b->params.upload(p->params, static_cast<size_t>(b->nparams), s);
  b->grids.upload(p->grid[0], static_cast<size_t>(S), s);
  b->guess.upload(p->guess, static_cast<size_t>(S), s);
  b->explicit_grids = b->explicit_guess = true;
  const int cap = res->log != nullptr && res->log_capacity > 0 ? res->log_capacity : 0;
  b->log_capacity = cap;
  if (cap > 0) b->log.alloc(static_cast<size_t>(cap));
  cudaEvent_t e0, e1;
  EP_CUDA(cudaEventCreate(&e0));
  EP_CUDA(cudaEventCreate(&e1));
  EP_CUDA(cudaEventRecord(e0, s));
  batch_launch(b, opt, cap, 0, -1, nullptr);
  // results of the solve, needed to decide on the diagnostics
  int h_int[5] = {0, 0, 0, 0, 0};   // error, exit_reason, n_outer, n_newton, log_len
  int h_flags = 0;
  double h_rn = 0.0;
  b->flags.download(&h_flags, 1, s);
  b->error.download(&h_int[0], 1, s);
  b->exit_reason.download(&h_int[1], 1, s);
  b->n_outer.download(&h_int[2], 1, s);
  b->n_newton.download(&h_int[3], 1, s);
  b->log_len.download(&h_int[4], 1, s);
  b->residual_norm.download(&h_rn, 1, s);
  EP_CUDA(cudaStreamSynchronize(s));
  if ((h_flags & 1) != 0 && fused_singular_falls_back()) {
    // a Newton iteration met a singular Jacobian: the fused kernel gives such a step up at once, the generic engine
    // first takes NLsolve's regularised-normal-equations direction (normaleq.cuh) -- solve the problem there
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return false;
  }
  res->sign_flags = 0;
  res->nonmonotone_entries = 0;
  res->nonmonotone_max = 0.0;
  // src/monotonicity.jl: the sign convention is checked on the first Jacobian assembled with a finite step, stencil
  // monotonicity (opt-in) on the first Jacobian whatever the step -- both at the guess, where the first Newton
  // iteration of the solve linearised
  const bool do_sign = std::isfinite(opt->delta), do_stencils = opt->check_monotonicity != 0;
  if (h_int[0] == 0 && h_int[3] > 0 && (do_sign || do_stencils)) {
    DevBuf<int> counts;
    DevBuf<unsigned long long> mx;
    counts.alloc(3);
    mx.alloc(1);
    counts.zero(s);
    mx.zero(s);
    ops->diag_1d(static_cast<int>(S), p->params, b->grids.p, b->guess.p, opt->delta, opt->monotonicity_tol, do_sign ? 1 : 0,
                 do_stencils ? 1 : 0, counts.p, mx.p, s);
    int h[3];
    unsigned long long hm = 0;
    counts.download(h, 3, s);
    mx.download(&hm, 1, s);
    EP_CUDA(cudaStreamSynchronize(s));
    if (do_sign && h[0] > 0 && h[1] == h[0]) {
      res->sign_flags |= 1;
      if (opt->delta > 0) res->sign_flags |= ECONPDES_SIGN_WHOLE_SYSTEM;
    }
    if (do_stencils) {
      res->nonmonotone_entries = h[2];
      std::memcpy(&res->nonmonotone_max, &hm, sizeof(double));
    }
  }
  DevBuf<double> saved;
  if (h_int[0] == 0 && res->saved != nullptr && ops->nsaved() > 0) {
    saved.alloc(static_cast<size_t>(ops->nsaved()) * S);
    ops->saved_1d(static_cast<int>(S), p->params, b->grids.p, b->solution.p, saved.p, s);
  }
  EP_CUDA(cudaEventRecord(e1, s));
  if (res->solution != nullptr) b->solution.download(res->solution, static_cast<size_t>(S), s);
  if (saved.p != nullptr) saved.download(res->saved, saved.n, s);
  const int nrows = h_int[4] < cap ? h_int[4] : cap;
  if (nrows > 0) b->log.download(res->log, static_cast<size_t>(nrows), s);
  EP_CUDA(cudaEventSynchronize(e1));
  EP_CUDA(cudaStreamSynchronize(s));
  float ms = 0.f;
  EP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  res->log_len = nrows;
  res->residual_norm = h_rn;
  res->converged = h_rn <= opt->abstol ? 1 : 0;
  res->exit_reason = h_int[1];
  res->n_outer = h_int[2];
  res->n_newton = h_int[3];
  res->n_linear_iters = 0;
  res->n_linear_solves = 0;
  res->n_linear_loose = 0;
  res->lin_max_relres = 0.0;
  res->device_ms = ms;
  res->kernel_launches = g_kernel_launches.load() - l0;
  *rc_out = ECONPDES_OK;
  if (h_int[0] != 0) {
    set_error("G! returns NaN with the initial value");   // finiteschemesolve.jl:372
    *rc_out = ECONPDES_E_NAN_GUESS;
  }
  return true;
}

int econpdes_solve_stationary(const econpdes_problem* problem, const econpdes_options* options,
                              econpdes_result* result) {
  return guarded([&] {
    if (!problem || !options || !result) throw std::invalid_argument("null argument");
    int rc = ECONPDES_OK;
    if (solve_fused_1d(problem, options, result, &rc)) return rc;
    Solver solver(make_ctx(problem));
    return solver.solve(*options, result);
  });
}

int econpdes_solve_backward(const econpdes_problem* problem, const econpdes_options* options, const double* taus,
                            int32_t ntaus, double* solution, double* saved, double* residual_norms,
                            int32_t* newton_iters, double* device_ms, int64_t* kernel_launches) {
  return guarded([&] {
    if (!problem || !options) throw std::invalid_argument("null argument");
    Solver solver(make_ctx(problem));
    return solver.solve_backward(*options, taus, ntaus, solution, saved, residual_norms, newton_iters, device_ms,
                                 kernel_launches);
  });
}

// ---- slab-decomposed solve ----
int econpdes_dist_unique_id(uint8_t* id) {
  return guarded([&] {
    if (!id) throw std::invalid_argument("null id");
    Dist::unique_id(id);
    return ECONPDES_OK;
  });
}

int econpdes_dist_create(int rank, int world, const uint8_t* id, econpdes_dist** out) {
  return guarded([&] {
    if (!id || !out) throw std::invalid_argument("null argument");
    ensure_device();
    auto d = std::make_unique<econpdes_dist>();
    d->dist = std::make_unique<Dist>(rank, world, id);
    *out = d.release();
    return ECONPDES_OK;
  });
}

void econpdes_dist_destroy(econpdes_dist* d) { delete d; }

int64_t econpdes_dist_collectives(const econpdes_dist* d) { return d && d->dist ? d->dist->collectives() : 0; }

int64_t econpdes_dist_p2p_exchanges(const econpdes_dist* d) { return d && d->dist ? d->dist->p2p_exchanges() : 0; }

int econpdes_slab_range(int64_t size_last, int rank, int world, int64_t* first, int64_t* count) {
  return guarded([&] {
    if (!first || !count || world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("bad rank/world");
    const int64_t units = size_last / 2;   // pairs of planes
    if (units < world) throw std::invalid_argument("every rank needs at least 2 planes of the slowest dimension");
    const int64_t base = units / world, rem = units % world;
    const int64_t ufirst = rank * base + (rank < rem ? rank : rem);
    const int64_t ucount = base + (rank < rem ? 1 : 0);
    *first = 2 * ufirst;
    *count = 2 * ucount + ((rank == world - 1) ? size_last % 2 : 0);
    return ECONPDES_OK;
  });
}

int econpdes_solve_stationary_slab(econpdes_dist* d, const econpdes_problem* problem, int64_t slab_first,
                                   int64_t slab_count, const econpdes_options* options, econpdes_result* result) {
  return guarded([&] {
    if (!d || !d->dist || !problem || !options || !result) throw std::invalid_argument("null argument");
    Solver solver(make_ctx(problem, d->dist.get(), slab_first, slab_count));
    return solver.solve(*options, result);
  });
}

// ---- batched sweep ----
int econpdes_batch_create(int model_id, int64_t nproblems, int64_t npoints, econpdes_batch** out) {
  return guarded([&] {
    const ModelOps* ops = model_ops(model_id);
    if (!ops) throw std::invalid_argument("unknown model id");
    if (ops->N() != 1 || ops->F() != 1)
      throw std::invalid_argument("the batched engine takes 1-D single-unknown models");
    if (ops->time_slot() >= 0)
      throw std::invalid_argument("the batched engine solves stationary problems: the PDE function takes a time argument");
    if (nproblems < 1 || npoints < 2) throw std::invalid_argument("need >= 1 problem of >= 2 points");
    if (npoints > 1024) throw std::runtime_error("batched engine: at most 1024 points per problem");
    if (!out) throw std::invalid_argument("null batch pointer");
    ensure_device();
    auto b = std::make_unique<econpdes_batch>();
    b->model_id = model_id;
    b->B = nproblems;
    b->S = static_cast<int>(npoints);
    b->nparams = ops->nparams();
    EP_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    const size_t B = static_cast<size_t>(nproblems), S = static_cast<size_t>(npoints);
    b->params.alloc(B * (b->nparams > 0 ? b->nparams : 1));
    b->solution.alloc(B * S);   // grids / guess: allocated by the upload that brings them
    b->residual_norm.alloc(B);
    b->ints.alloc(4 * B);
    b->error.p = b->ints.p;
    b->exit_reason.p = b->ints.p + B;
    b->n_outer.p = b->ints.p + 2 * B;
    b->n_newton.p = b->ints.p + 3 * B;
    b->log_len.alloc(B);
    b->flags.alloc(B);
    *out = b.release();
    return ECONPDES_OK;
  });
}

void econpdes_batch_destroy(econpdes_batch* b) { delete b; }

int econpdes_batch_upload(econpdes_batch* b, const double* params, const double* grids, const double* guess) {
  return guarded([&] {
    if (!b || !grids || !guess) throw std::invalid_argument("null argument");
    const size_t B = static_cast<size_t>(b->B), S = static_cast<size_t>(b->S);
    if (b->nparams > 0) {
      if (!params) throw std::invalid_argument("null params");
      b->params.upload(params, B * b->nparams, b->stream);
    }
    b->grids.upload(grids, B * S, b->stream);
    b->guess.upload(guess, B * S, b->stream);
    b->explicit_grids = b->explicit_guess = true;
    EP_CUDA(cudaStreamSynchronize(b->stream));
    return ECONPDES_OK;
  });
}

// validates a compact input description against the batch shape and sizes the device buffers it needs
static void batch_accept_inputs(econpdes_batch* b, const econpdes_batch_inputs* in) {
  if (!in) throw std::invalid_argument("null inputs");
  const size_t B = static_cast<size_t>(b->B), S = static_cast<size_t>(b->S);
  if (b->nparams > 0 && !in->params) throw std::invalid_argument("null params");
  if (in->grids != nullptr) {
    b->explicit_grids = true;
    b->grids.alloc(B * S);
  } else {
    if (in->n_range < 1 || in->n_range >= b->S) throw std::invalid_argument("n_range must lie in [1, npoints)");
    if (!in->range_lo || !in->points) throw std::invalid_argument("null range_lo / points");
    b->explicit_grids = false;
    b->n_range = in->n_range;
    b->range_lo.alloc(B);
    b->points.alloc(B * (S - static_cast<size_t>(in->n_range)));
  }
  if (in->guess != nullptr) {
    b->explicit_guess = true;
    b->guess.alloc(B * S);
  } else {
    b->explicit_guess = false;
    b->guess_fill = in->guess_fill;
  }
}

// queues the host->device copies of problems [first, last) on stream s
static void batch_upload_slice(econpdes_batch* b, const econpdes_batch_inputs* in, long long first, long long last,
                               cudaStream_t s) {
  const size_t n = static_cast<size_t>(last - first), o = static_cast<size_t>(first);
  const size_t S = static_cast<size_t>(b->S), P = static_cast<size_t>(b->nparams);
  if (P > 0)
    EP_CUDA(cudaMemcpyAsync(b->params.p + o * P, in->params + o * P, n * P * sizeof(double), cudaMemcpyHostToDevice, s));
  if (b->explicit_grids) {
    EP_CUDA(cudaMemcpyAsync(b->grids.p + o * S, in->grids + o * S, n * S * sizeof(double), cudaMemcpyHostToDevice, s));
  } else {
    const size_t np = S - static_cast<size_t>(b->n_range);
    EP_CUDA(cudaMemcpyAsync(b->range_lo.p + o, in->range_lo + o, n * sizeof(double), cudaMemcpyHostToDevice, s));
    EP_CUDA(cudaMemcpyAsync(b->points.p + o * np, in->points + o * np, n * np * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  if (b->explicit_guess)
    EP_CUDA(cudaMemcpyAsync(b->guess.p + o * S, in->guess + o * S, n * S * sizeof(double), cudaMemcpyHostToDevice, s));
}

int econpdes_batch_upload_inputs(econpdes_batch* b, const econpdes_batch_inputs* in) {
  return guarded([&] {
    if (!b) throw std::invalid_argument("null batch");
    batch_accept_inputs(b, in);
    batch_upload_slice(b, in, 0, b->B, b->stream);
    EP_CUDA(cudaStreamSynchronize(b->stream));
    return ECONPDES_OK;
  });
}

// launches the problems [first, last) on stream s (the whole batch on the batch's own stream by default)
static int batch_launch(econpdes_batch* b, const econpdes_options* options, int log_capacity, long long first,
                        long long last, cudaStream_t s, const BatchOut* out) {
  if (std::isnan(options->delta) || options->delta == 0.0)
    throw std::invalid_argument("`Δ` entries must be nonzero and not NaN (use `Inf` for no damping, a negative entry for forward relaxation)");
  BatchArgs a;
  a.problem0 = first;
  a.nproblems = last < 0 ? b->B : last;
  a.S = b->S;
  a.nparams = b->nparams;
  a.params = b->params.p;
  a.grids = b->grids.p;
  a.guess = b->guess.p;
  if ((b->explicit_grids && !a.grids) || (b->explicit_guess && !a.guess))
    throw std::invalid_argument("the batch has no grids / guess yet: upload them first");
  a.solution = b->solution.p;
  a.residual_norm = b->residual_norm.p;
  a.exit_reason = b->exit_reason.p;
  a.n_outer = b->n_outer.p;
  a.n_newton = b->n_newton.p;
  a.error = b->error.p;
  if (log_capacity > 0) {
    b->log.alloc(static_cast<size_t>(b->B) * log_capacity);
    b->log_capacity = log_capacity;
  }
  a.log = b->log_capacity > 0 ? b->log.p : nullptr;
  a.log_capacity = b->log_capacity;
  a.log_len = b->log_len.p;
  a.flags = b->flags.p;
  if (!b->nflag_host) {
    EP_CUDA(cudaMallocHost(&b->nflag_host, sizeof(int)));
    *b->nflag_host = 0;
  }
  {
    void* dp = nullptr;
    EP_CUDA(cudaHostGetDevicePointer(&dp, b->nflag_host, 0));
    a.flagged_count = static_cast<int*>(dp);
  }
  a.opt = *options;
  if (out != nullptr) {
    const size_t nB = static_cast<size_t>(b->B);
    if (out->solution) a.solution = out->solution;
    if (out->residual_norm) a.residual_norm = out->residual_norm;
    if (out->ints) {
      a.error = out->ints;
      a.exit_reason = out->ints + nB;
      a.n_outer = out->ints + 2 * nB;
      a.n_newton = out->ints + 3 * nB;
    }
  }
  const int64_t l0 = g_kernel_launches.load();
  // compact inputs: every block builds its own grid / guess while it stages the problem (no expansion pass, no
  // B x S arrays in HBM)
  if (!b->explicit_grids) {
    a.grids = nullptr;
    a.range_lo = b->range_lo.p;
    a.points = b->points.p;
    a.n_range = b->n_range;
  }
  if (!b->explicit_guess) {
    a.guess = nullptr;
    a.guess_fill = b->guess_fill;
  }
  if (!model_ops(b->model_id)->batched(a, s ? s : b->stream)) throw std::runtime_error("model has no batched kernel");
  b->launches += g_kernel_launches.load() - l0;
  return ECONPDES_OK;
}

// Problems of the batch whose Newton iteration met a singular Jacobian (flags[b] & 1) are solved again by the generic
// engine, which takes NLsolve's regularised-normal-equations direction there instead of giving the step up
// (normaleq.cuh).  Rare by construction (*nflag_host counts them; no problem of the C4 sweep gets here), so one
// problem at a time.  `host`: where the one-call path's results live (host pointers, indexed by problem) -- null for
// the handle path, whose results stay in the batch's device buffers.
struct BatchHostOut {
  double* solution = nullptr;
  double* residual_norm = nullptr;
  int* ints = nullptr;   // [4][B]: error, exit_reason, n_outer, n_newton
};
static void batch_resolve_singular(econpdes_batch* b, const econpdes_options* opt, const BatchHostOut* host) {
  static const bool on = [] { const char* e = std::getenv("ECONPDES_SINGULAR_RETRY"); return !(e && e[0] == '0'); }();
  if (!b->nflag_host || *b->nflag_host == 0) return;
  *b->nflag_host = 0;
  if (!on) return;
  const size_t B = static_cast<size_t>(b->B), S = static_cast<size_t>(b->S), P = static_cast<size_t>(b->nparams);
  cudaStream_t s = b->stream;
  std::vector<int> flags(B);
  b->flags.download(flags.data(), B, s);
  EP_CUDA(cudaStreamSynchronize(s));
  std::vector<double> grid(S), guess(S), prm(P > 0 ? P : 1), sol(S);
  const int cap = b->log_capacity;
  std::vector<econpdes_log_row> log(cap > 0 ? cap : 1);
  for (size_t i = 0; i < B; ++i) {
    if ((flags[i] & 1) == 0) continue;
    if (b->explicit_grids) {
      EP_CUDA(cudaMemcpyAsync(grid.data(), b->grids.p + i * S, S * sizeof(double), cudaMemcpyDeviceToHost, s));
    } else {   // the same range_point the kernel evaluates (plain IEEE arithmetic on either side)
      const size_t nr = static_cast<size_t>(b->n_range), np = S - nr;
      double lo = 0.0;
      EP_CUDA(cudaMemcpyAsync(&lo, b->range_lo.p + i, sizeof(double), cudaMemcpyDeviceToHost, s));
      EP_CUDA(cudaMemcpyAsync(grid.data() + nr, b->points.p + i * np, np * sizeof(double), cudaMemcpyDeviceToHost, s));
      EP_CUDA(cudaStreamSynchronize(s));
      for (size_t k = 0; k < nr; ++k) grid[k] = range_point(lo, grid[nr], static_cast<int>(nr) + 1, static_cast<int>(k));
    }
    if (b->explicit_guess)
      EP_CUDA(cudaMemcpyAsync(guess.data(), b->guess.p + i * S, S * sizeof(double), cudaMemcpyDeviceToHost, s));
    else
      std::fill(guess.begin(), guess.end(), b->guess_fill);
    if (P > 0) EP_CUDA(cudaMemcpyAsync(prm.data(), b->params.p + i * P, P * sizeof(double), cudaMemcpyDeviceToHost, s));
    EP_CUDA(cudaStreamSynchronize(s));
    econpdes_problem p{};
    p.model_id = b->model_id;
    p.ndims = 1;
    p.nfuncs = 1;
    p.nparams = b->nparams;
    p.size[0] = b->S;
    p.grid[0] = grid.data();
    p.params = P > 0 ? prm.data() : nullptr;
    p.guess = guess.data();
    econpdes_result r{};
    r.solution = sol.data();
    r.log = cap > 0 ? log.data() : nullptr;
    r.log_capacity = cap;
    Solver solver(make_ctx(&p));
    const int rc = solver.solve(*opt, &r);
    const int ints[4] = {rc == ECONPDES_OK ? 0 : rc, r.exit_reason, r.n_outer, r.n_newton};
    if (host != nullptr) {
      if (host->solution) std::memcpy(host->solution + i * S, sol.data(), S * sizeof(double));
      if (host->residual_norm) host->residual_norm[i] = r.residual_norm;
      if (host->ints)
        for (int q = 0; q < 4; ++q) host->ints[q * B + i] = ints[q];
    } else {
      EP_CUDA(cudaMemcpyAsync(b->solution.p + i * S, sol.data(), S * sizeof(double), cudaMemcpyHostToDevice, s));
      EP_CUDA(cudaMemcpyAsync(b->residual_norm.p + i, &r.residual_norm, sizeof(double), cudaMemcpyHostToDevice, s));
      for (int q = 0; q < 4; ++q)
        EP_CUDA(cudaMemcpyAsync(b->ints.p + q * B + i, &ints[q], sizeof(int), cudaMemcpyHostToDevice, s));
      if (cap > 0 && b->log.p != nullptr) {
        const int len = r.log_len < cap ? r.log_len : cap;
        EP_CUDA(cudaMemcpyAsync(b->log.p + i * cap, log.data(), static_cast<size_t>(len) * sizeof(econpdes_log_row),
                                cudaMemcpyHostToDevice, s));
        EP_CUDA(cudaMemcpyAsync(b->log_len.p + i, &len, sizeof(int), cudaMemcpyHostToDevice, s));
      }
      EP_CUDA(cudaStreamSynchronize(s));   // the sources are locals of this iteration
    }
  }
}

int econpdes_batch_solve(econpdes_batch* b, const econpdes_options* options, double* device_ms) {
  return guarded([&] {
    if (!b || !options) throw std::invalid_argument("null argument");
    cudaEvent_t e0, e1;
    EP_CUDA(cudaEventCreate(&e0));
    EP_CUDA(cudaEventCreate(&e1));
    EP_CUDA(cudaEventRecord(e0, b->stream));
    batch_launch(b, options, b->log_capacity, 0, -1, nullptr);
    EP_CUDA(cudaEventRecord(e1, b->stream));
    EP_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    EP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (device_ms) *device_ms = ms;
    batch_resolve_singular(b, options, nullptr);
    return ECONPDES_OK;
  });
}

int econpdes_batch_set_log_capacity(econpdes_batch* b, int32_t log_capacity) {
  return guarded([&] {
    if (!b || log_capacity < 0) throw std::invalid_argument("bad log capacity");
    b->log_capacity = log_capacity;
    if (log_capacity > 0) b->log.alloc(static_cast<size_t>(b->B) * log_capacity);
    return ECONPDES_OK;
  });
}

int econpdes_batch_download(econpdes_batch* b, double* solution, double* residual_norm, int32_t* exit_reason,
                            int32_t* n_outer, int32_t* n_newton, econpdes_log_row* log, int32_t log_capacity,
                            int32_t* log_len) {
  return guarded([&] {
    if (!b) throw std::invalid_argument("null batch");
    const size_t B = static_cast<size_t>(b->B), S = static_cast<size_t>(b->S);
    std::vector<int> err(B);
    b->error.download(err.data(), B, b->stream);
    if (solution) b->solution.download(solution, B * S, b->stream);
    if (residual_norm) b->residual_norm.download(residual_norm, B, b->stream);
    if (exit_reason) b->exit_reason.download(exit_reason, B, b->stream);
    if (n_outer) b->n_outer.download(n_outer, B, b->stream);
    if (n_newton) b->n_newton.download(n_newton, B, b->stream);
    if (log && log_len) {
      if (log_capacity != b->log_capacity || b->log.p == nullptr)
        throw std::invalid_argument("log capacity differs from the one the batch was solved with");
      b->log.download(log, B * static_cast<size_t>(log_capacity), b->stream);
      b->log_len.download(log_len, B, b->stream);
    }
    EP_CUDA(cudaStreamSynchronize(b->stream));
    for (size_t i = 0; i < B; ++i)
      if (err[i] != 0) {
        set_error("G! returns NaN with the initial value (problem " + std::to_string(i) + ")");
        return ECONPDES_E_NAN_GUESS;
      }
    return ECONPDES_OK;
  });
}

int64_t econpdes_batch_kernel_launches(const econpdes_batch* b) { return b ? b->launches : 0; }

// The one-call path keeps its device workspace between calls of the same shape (a
// calibration loop calls it thousands of times): cudaMalloc/cudaFree of ~0.5 GB per call
// cost more than the solve.  econpdes_release_workspace() frees it.
static thread_local econpdes_batch* t_workspace = nullptr;

int econpdes_selftest_reciprocal(uint64_t seed, int64_t n, int64_t mismatch[3]) {
  return guarded([&] {
    if (!mismatch || n < 0) throw std::invalid_argument("bad argument");
    ensure_device();
    DevBuf<unsigned long long> d;
    d.alloc(3);
    EP_CUDA(cudaMemset(d.p, 0, 3 * sizeof(unsigned long long)));
    EP_LAUNCH(k_selftest_reciprocal, 148 * 8, 256, 0, nullptr, seed, n, d.p);
    unsigned long long h[3];
    EP_CUDA(cudaMemcpy(h, d.p, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 3; ++i) mismatch[i] = static_cast<int64_t>(h[i]);
    return static_cast<int>(ECONPDES_OK);
  });
}

int econpdes_set_batch_points_per_thread(int points) {
  if (points < 0 || points > 4) {
    set_error("points per thread must be 0 (automatic), 1, 2, 3 or 4");
    return ECONPDES_E_INVALID;
  }
  return g_batch_points_per_thread.exchange(points);
}

int econpdes_last_batched_output_mode(void) { return t_workspace ? (t_workspace->last_direct ? 1 : 0) : -1; }

void econpdes_release_workspace(void) {
  delete t_workspace;
  t_workspace = nullptr;
  delete t_single;
  t_single = nullptr;
}

int econpdes_solve_batched(int model_id, int64_t nproblems, int64_t npoints, const double* params, const double* grids,
                           const double* guess, const econpdes_options* options, double* solution,
                           double* residual_norm, int32_t* exit_reason, int32_t* n_outer, int32_t* n_newton) {
  if (!grids || !guess) {
    set_error("null argument");
    return ECONPDES_E_INVALID;
  }
  econpdes_batch_inputs in{};
  in.params = params;
  in.grids = grids;
  in.guess = guess;
  return econpdes_solve_batched_inputs(model_id, nproblems, npoints, &in, options, solution, residual_norm, exit_reason,
                                       n_outer, n_newton);
}

int econpdes_solve_batched_inputs(int model_id, int64_t nproblems, int64_t npoints, const econpdes_batch_inputs* in,
                                  const econpdes_options* options, double* solution, double* residual_norm,
                                  int32_t* exit_reason, int32_t* n_outer, int32_t* n_newton) {
  int device = -1;
  cudaGetDevice(&device);
  econpdes_batch* b = t_workspace;
  if (b == nullptr || b->model_id != model_id || b->B != nproblems || b->S != npoints || b->device != device) {
    econpdes_release_workspace();
    const int rc = econpdes_batch_create(model_id, nproblems, npoints, &b);
    if (rc != ECONPDES_OK) return rc;
    b->device = device;
    t_workspace = b;
  }
  // Copy/compute pipeline: the batch is cut into slices; slice c's upload, solve and download are queued in
  // order on stream c % 3, so the host->device copy of one slice, the kernel of another and the device->host
  // copy of a third run at the same time (separate copy engines), and kernels of neighbouring slices fill each
  // other's tail waves.  Problems are independent, so the result does not depend on the slicing.
  return guarded([&] {
    if (!options) throw std::invalid_argument("null argument");
    batch_accept_inputs(b, in);
    const long long B = b->B;
    const size_t S = static_cast<size_t>(b->S);
    if (b->sms == 0) {
      b->sms = 148;
      cudaDeviceGetAttribute(&b->sms, cudaDevAttrMultiProcessorCount, device);
    }
    const int sms = b->sms;
    // Zero-copy results: when the caller's solution buffer is pinned (page-locked and mapped into the device's
    // address space, which cudaHostAlloc / cudaHostRegister memory is under unified addressing), every block stores
    // its solution straight into it over PCIe as it finishes, and the per-problem scalars go to the library's own
    // pinned arrays the same way: the writes of one block overlap the arithmetic of the others, there is no
    // device->host copy to queue or to drain, and a slice costs two or three small uploads and two launches.
    // Pageable buffers (and ECONPDES_DIRECT_OUTPUT=0) take the staged path: device buffers + copies per slice.
    static const bool direct_allowed = [] {
      const char* e = std::getenv("ECONPDES_DIRECT_OUTPUT");
      return !(e && e[0] == '0');
    }();
    BatchOut direct;
    bool use_direct = false;
    if (direct_allowed && solution != nullptr) {
      cudaPointerAttributes at{};
      if (cudaPointerGetAttributes(&at, solution) == cudaSuccess && at.type == cudaMemoryTypeHost &&
          at.devicePointer != nullptr) {
        direct.solution = static_cast<double*>(at.devicePointer);
        use_direct = true;
      } else {
        cudaGetLastError();
      }
    }
    b->last_direct = use_direct;
    // Slice sizes in whole waves (one block per SM): short slices at both ends so the pipeline fills and
    // drains quickly (the first kernel waits for one small upload, the last download is small), long ones
    // in the middle (at most a sixth of the batch) where copies hide behind the kernels.
    std::vector<long long> bounds{0};
    // bytes a problem uploads; with zero-copy results and light inputs (compact grids: ~0.85 KB per problem, 30x
    // faster to upload than to solve) three or four slices growing by 4x hide every upload but the first
    size_t up_bytes = static_cast<size_t>(b->nparams) * sizeof(double);
    up_bytes += b->explicit_grids ? S * sizeof(double) : (S - static_cast<size_t>(b->n_range) + 1) * sizeof(double);
    if (b->explicit_guess) up_bytes += S * sizeof(double);
    if (use_direct && up_bytes <= 2048) {
      const long long waves = (B + sms - 1) / sms;
      long long left = waves, w = 2;
      while (left > 0) {
        const long long take = (w < left && left - w >= 2) ? w : left;
        const long long nxt = bounds.back() + take * sms;
        bounds.push_back(nxt < B ? nxt : B);
        left -= take;
        w *= 4;
      }
    } else {
      const long long waves = (B + sms - 1) / sms;
      const long long cap = waves / 6 > 2 ? waves / 6 : 2;
      std::vector<long long> sizes, back;
      long long left = waves, w = 2;
      while (left > 0) {                 // 2, 4, 8, ... waves from both ends towards the middle, capped
        long long take = w < left ? w : left;
        sizes.push_back(take);
        left -= take;
        if (left <= 0) break;
        take = w < left ? w : left;
        back.push_back(take);
        left -= take;
        w = 2 * w < cap ? 2 * w : cap;
      }
      sizes.insert(sizes.end(), back.rbegin(), back.rend());
      for (long long wv : sizes) {
        const long long nxt = bounds.back() + wv * sms;
        bounds.push_back(nxt < B ? nxt : B);
        if (bounds.back() == B) break;
      }
      if (bounds.back() != B) bounds.push_back(B);
    }
    for (cudaStream_t& p : b->pipe)
      if (!p) EP_CUDA(cudaStreamCreateWithFlags(&p, cudaStreamNonBlocking));
    if (!b->err_host) EP_CUDA(cudaMallocHost(&b->err_host, 4 * static_cast<size_t>(B) * sizeof(int)));
    if (!b->rn_host) EP_CUDA(cudaMallocHost(&b->rn_host, static_cast<size_t>(B) * sizeof(double)));
    const size_t nB = static_cast<size_t>(B);
    b->log_capacity = 0;
    int* err = b->err_host;
    if (use_direct) {
      void* dp = nullptr;
      EP_CUDA(cudaHostGetDevicePointer(&dp, b->err_host, 0));
      direct.ints = static_cast<int*>(dp);
      EP_CUDA(cudaHostGetDevicePointer(&dp, b->rn_host, 0));
      direct.residual_norm = static_cast<double*>(dp);
    }
    // if queuing fails half-way, copies into the caller's buffers may still be in flight: drain the
    // streams before the error leaves this call
    struct Drain {
      econpdes_batch* b;
      bool armed = true;
      ~Drain() {
        if (armed)
          for (cudaStream_t p : b->pipe)
            if (p) cudaStreamSynchronize(p);
      }
    } drain{b};
    // ECONPDES_BATCH_TIMELINE=1: where one call's time goes (host clock around queueing and waiting, device events
    // around each slice's upload and solve), printed to stderr
    static const bool timeline = [] { const char* e = std::getenv("ECONPDES_BATCH_TIMELINE"); return e && e[0] == '1'; }();
    std::vector<cudaEvent_t> tev;
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_us = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count(); };
    auto mark = [&](cudaStream_t s) {
      if (!timeline) return;
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, s);
      tev.push_back(e);
    };
    for (size_t c = 0; c + 1 < bounds.size(); ++c) {
      const long long first = bounds[c], last = bounds[c + 1];
      const size_t n = static_cast<size_t>(last - first), o = static_cast<size_t>(first);
      cudaStream_t s = b->pipe[c % econpdes_batch::kPipe];
      mark(s);
      batch_upload_slice(b, in, first, last, s);
      mark(s);
      batch_launch(b, options, 0, first, last, s, use_direct ? &direct : nullptr);
      mark(s);
      if (use_direct) continue;   // the kernel has the host addresses: nothing to copy back
      // error, exit_reason, n_outer, n_newton of the slice: four rows of one pitched copy
      EP_CUDA(cudaMemcpy2DAsync(err + o, nB * sizeof(int), b->ints.p + o, nB * sizeof(int), n * sizeof(int), 4,
                                cudaMemcpyDeviceToHost, s));
      if (solution)
        EP_CUDA(cudaMemcpyAsync(solution + o * S, b->solution.p + o * S, n * S * sizeof(double), cudaMemcpyDeviceToHost, s));
      if (residual_norm)
        EP_CUDA(cudaMemcpyAsync(b->rn_host + o, b->residual_norm.p + o, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    const double t_queued = host_us();
    for (cudaStream_t p : b->pipe) EP_CUDA(cudaStreamSynchronize(p));
    const double t_synced = host_us();
    drain.armed = false;
    if (timeline) {
      std::fprintf(stderr, "[batch timeline] %zu slices, direct %d: queued %.0f us, synced %.0f us;", bounds.size() - 1,
                   use_direct ? 1 : 0, t_queued, t_synced);
      for (size_t c = 0; 3 * c + 2 < tev.size(); ++c) {
        float up0 = 0.f, k0 = 0.f, k1 = 0.f;
        cudaEventElapsedTime(&up0, tev[0], tev[3 * c]);
        cudaEventElapsedTime(&k0, tev[0], tev[3 * c + 1]);
        cudaEventElapsedTime(&k1, tev[0], tev[3 * c + 2]);
        std::fprintf(stderr, " slice %zu (%lld problems): upload from %.0f, solve %.0f..%.0f us;", c,
                     bounds[c + 1] - bounds[c], up0 * 1e3, k0 * 1e3, k1 * 1e3);
      }
      std::fprintf(stderr, "\n");
      for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    {
      BatchHostOut ho;
      ho.solution = solution;
      ho.residual_norm = b->rn_host;
      ho.ints = err;
      batch_resolve_singular(b, options, &ho);
    }
    if (residual_norm) std::memcpy(residual_norm, b->rn_host, nB * sizeof(double));
    if (exit_reason) std::memcpy(exit_reason, err + nB, nB * sizeof(int32_t));
    if (n_outer) std::memcpy(n_outer, err + 2 * nB, nB * sizeof(int32_t));
    if (n_newton) std::memcpy(n_newton, err + 3 * nB, nB * sizeof(int32_t));
    for (size_t i = 0; i < static_cast<size_t>(B); ++i)
      if (err[i] != 0) {
        set_error("G! returns NaN with the initial value (problem " + std::to_string(i) + ")");
        return static_cast<int>(ECONPDES_E_NAN_GUESS);
      }
    return static_cast<int>(ECONPDES_OK);
  });
}

}  // extern "C"
