// Generic engine: per-model kernel dispatch, the Newton solve of one implicit step and
// the pseudo-transient continuation driver -- the device-side counterpart of
// `finiteschemesolve` (src/finiteschemesolve.jl:343-465), `implicit_timestep` (:153-230),
// `_newtonstep` (:125-147) and NLsolve's `newton_` (restated in oracle/econpdes_oracle.py).
// Scalars of the control loop (norms, the SER controller) live on the host; every field
// stays in HBM for the whole solve.
#include "engine.cuh"
#include "normaleq.cuh"

#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include <cstdlib>

#include "batched1d.cuh"
#include "tridiag.cuh"

namespace econpdes {

std::atomic<int64_t> g_kernel_launches{0};

// ---- canary regions around device buffers (ECONPDES_GUARD=1; see GuardRegistry in engine.cuh) ----
namespace {
std::mutex g_guard_mutex;
std::map<void*, size_t> g_guard_live;        // base -> total bytes (canaries included)
long long g_guard_violations = 0;
long long count_bad_canaries(void* base, size_t total) {
  unsigned char h[2 * GuardRegistry::kBytes];
  // guard lengths are kBytes rounded up to whole elements: checking the outermost kBytes of either side is exact
  // for every element size that divides 256 (all of the library's: 1, 4, 8, 16, 24-byte rows are checked in part)
  if (cudaMemcpy(h, base, GuardRegistry::kBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  if (cudaMemcpy(h + GuardRegistry::kBytes, static_cast<unsigned char*>(base) + total - GuardRegistry::kBytes,
                 GuardRegistry::kBytes, cudaMemcpyDeviceToHost) != cudaSuccess)
    return 0;
  long long bad = 0;
  for (unsigned char c : h) bad += c != 0xA5;
  return bad;
}
}  // namespace
bool GuardRegistry::enabled() {
  static const bool on = [] { const char* e = std::getenv("ECONPDES_GUARD"); return e && e[0] == '1'; }();
  return on;
}
void GuardRegistry::add(void* base, size_t total_bytes) {
  std::lock_guard<std::mutex> lk(g_guard_mutex);
  g_guard_live[base] = total_bytes;
}
void GuardRegistry::remove(void* base) {
  std::lock_guard<std::mutex> lk(g_guard_mutex);
  auto it = g_guard_live.find(base);
  if (it == g_guard_live.end()) return;
  cudaDeviceSynchronize();
  g_guard_violations += count_bad_canaries(base, it->second);
  g_guard_live.erase(it);
}
long long GuardRegistry::violations(bool check_live) {
  std::lock_guard<std::mutex> lk(g_guard_mutex);
  long long v = g_guard_violations;
  if (check_live) {
    cudaDeviceSynchronize();
    for (auto& kv : g_guard_live) v += count_bad_canaries(kv.first, kv.second);
  }
  return v;
}

void ensure_device() {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    throw CudaError(std::string("no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") +
                    "); libeconpdes_cuda has no CPU fallback");
  }
}

const double* Ctx::gather_to_host(const double* partial_dev, int count) {
  if (count > kGatherMax) throw std::invalid_argument("gather_to_host: too many values");
  const int world = dist != nullptr ? dist->world() : 1;
  const double* src = partial_dev;
  if (dist != nullptr) {
    dist->all_gather(partial_dev, gather.p, count, stream);
    src = gather.p;
  }
  EP_CUDA(cudaMemcpyAsync(h_gather, src, static_cast<size_t>(world) * count * sizeof(double), cudaMemcpyDeviceToHost,
                          stream));
  EP_CUDA(cudaStreamSynchronize(stream));
  return h_gather;
}

Ctx::~Ctx() {
  if (h_gather) cudaFreeHost(h_gather);
  if (h_norm) cudaFreeHost(h_norm);
  if (stream) cudaStreamDestroy(stream);
}

double Norms::norm2_over_n(long long n) const {
  if (any_nan) return std::numeric_limits<double>::quiet_NaN();
  return std::sqrt(sumsq) / static_cast<double>(n);
}

// ---------------------------------------------------------------------------------------
// model registry: the per-model kernel launchers are compiled in modelops.cu, once per slice of the
// functor list (modelops.cuh), so the build runs in parallel
// ---------------------------------------------------------------------------------------
std::atomic<int> g_batch_points_per_thread{0};

namespace {
#define X(M) +1
constexpr int kModelCount = 0 ECONPDES_FOR_EACH_MODEL(X);
#undef X
}  // namespace
int model_count() { return kModelCount; }

const ModelOps* model_ops(int model_id) {
  if (model_id < 0 || model_id >= model_count()) return nullptr;
  switch (model_id % kModelOpsParts) {
    case 0: return model_ops_part0(model_id);
    case 1: return model_ops_part1(model_id);
    case 2: return model_ops_part2(model_id);
    case 3: return model_ops_part3(model_id);
    case 4: return model_ops_part4(model_id);
    case 5: return model_ops_part5(model_id);
    case 6: return model_ops_part6(model_id);
    default: return model_ops_part7(model_id);
  }
}

// ---------------------------------------------------------------------------------------
// problem upload
// ---------------------------------------------------------------------------------------
std::unique_ptr<Ctx> make_ctx(const econpdes_problem* p, Dist* dist, long long slab_first, long long slab_count) {
  if (p == nullptr) throw std::invalid_argument("null problem");
  const ModelOps* ops = model_ops(p->model_id);
  if (ops == nullptr) throw std::invalid_argument("unknown model id " + std::to_string(p->model_id));
  if (p->ndims != ops->N() || p->nfuncs != ops->F())
    throw std::invalid_argument(std::string("model `") + ops->name() + "` expects " + std::to_string(ops->N()) +
                                " state variable(s) and " + std::to_string(ops->F()) + " unknown(s)");
  if (p->nparams != ops->nparams())
    throw std::invalid_argument(std::string("model `") + ops->name() + "` expects " +
                                std::to_string(ops->nparams()) + " parameters");
  if (p->guess == nullptr) throw std::invalid_argument("null guess");
  ensure_device();
  auto c = std::make_unique<Ctx>();
  c->model_id = p->model_id;
  c->N = ops->N();
  c->F = ops->F();
  c->nparams = ops->nparams();
  c->nsaved = ops->nsaved();
  c->box = ops->box();
  if (c->box && ops->F() == 1) c->plane_mask = ops->plane_mask();
  c->nk = c->box ? ipow3(c->N) : 1 + 2 * c->N;
  c->l = 1;
  c->ncolors = c->F;
  for (int d = 0; d < c->N; ++d) {
    if (p->size[d] < 2) throw std::invalid_argument("state grid entries must have at least 2 points");  // stategrid.jl:16
    if (p->grid[d] == nullptr) throw std::invalid_argument("null grid vector");
    c->S[d] = p->size[d];
    c->l *= p->size[d];
    c->ncolors *= p->size[d] < 3 ? static_cast<int>(p->size[d]) : 3;
  }
  c->n_glob = c->l * c->F;
  c->slab_count = c->S[c->N - 1];
  if (dist != nullptr) {
    if (c->N < 2 || c->F != 1)
      throw std::invalid_argument("the slab-decomposed solve takes 2-D/3-D grids with one unknown");
    if (p->delta_vec != nullptr) throw std::invalid_argument("the slab-decomposed solve takes a scalar Δ");
    if (slab_first < 0 || slab_count < 2 || slab_first + slab_count > c->S[c->N - 1])
      throw std::invalid_argument("slab range out of bounds (each rank needs >= 2 planes)");
    if (c->ncolors > kGatherMax) throw std::invalid_argument("too many colors for the slab path");
    c->dist = dist;
    c->slab_first = slab_first;
    c->slab_count = slab_count;
    c->halo = c->l / c->S[c->N - 1];
    c->l = c->halo * slab_count;
  }
  c->n = c->l * c->F;
  if (c->n + 2 * c->halo >= (1LL << 31))
    throw std::invalid_argument("grids of 2^31 unknowns or more per device are not supported (32-bit point indexing)");
  for (int i = 0; i < c->nparams; ++i) c->prm.v[i] = p->params[i];
  ops->derive(c->prm.v);
  EP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  cudaStream_t s = c->stream;
  for (int d = 0; d < c->N; ++d) {
    c->grid[d].upload(p->grid[d], static_cast<size_t>(c->S[d]), s);
    std::vector<double> tab(static_cast<size_t>(kGeomRows) * c->S[d]);
    make_geom_table(p->grid[d], c->S[d], tab.data());
    c->geomtab[d].upload(tab.data(), tab.size(), s);
    EP_CUDA(cudaStreamSynchronize(s));
  }
  c->guess.upload(p->guess, static_cast<size_t>(c->n), s);
  if (p->bc != nullptr) {
    bool any = false;
    for (long long j = 0; j < c->n && !any; ++j) any = p->bc[j] != 0.0;
    if (any) c->bc.upload(p->bc, static_cast<size_t>(c->n), s);
  }
  {
    // "tame" geometry (sweep.cuh): every spacing and tabulated reciprocal finite and within 2^+-200, every
    // boundary value below 2^199 -- then no stencil quotient of tame unknowns can overflow or hit 0 * Inf
    auto moderate = [](double v) { const double a = std::fabs(v); return a >= 0x1p-200 && a <= 0x1p200; };
    bool ok = true;
    for (int d = 0; d < c->N && ok; ++d) {
      std::vector<double> tab(static_cast<size_t>(kGeomRows) * c->S[d]);
      make_geom_table(p->grid[d], c->S[d], tab.data());
      for (double v : tab) ok = ok && moderate(v);
    }
    if (ok && c->bc.p != nullptr)
      for (long long j = 0; j < c->n && ok; ++j) ok = std::fabs(p->bc[j]) < 0x1p199;
    c->geom_tame = ok;
  }
  if (p->is_algebraic != nullptr) {
    bool any = false;
    for (long long j = 0; j < c->n && !any; ++j) any = p->is_algebraic[j] != 0;
    if (any) c->is_alg.upload(p->is_algebraic, static_cast<size_t>(c->n), s);
  }
  // Any non-default bound turns the solve into an MCP (finiteschemesolve.jl:156, 355)
  bool lbany = false, ubany = false;
  if (p->lower_bound != nullptr)
    for (long long j = 0; j < c->n && !lbany; ++j) lbany = p->lower_bound[j] != -INFINITY;
  if (p->upper_bound != nullptr)
    for (long long j = 0; j < c->n && !ubany; ++j) ubany = p->upper_bound[j] != INFINITY;
  c->has_bounds = lbany || ubany;
  if (dist != nullptr) {
    // Slab path: a bound that is finite on one rank only must still switch EVERY rank to the MCP residual and
    // the masked norm -- the masked norm adds collectives, which all ranks have to call in the same order
    const size_t world = static_cast<size_t>(dist->world());
    c->gather.alloc(world * kGatherMax);
    EP_CUDA(cudaMallocHost(&c->h_gather, world * kGatherMax * sizeof(double)));
    DevBuf<double> flag;
    const double mine = c->has_bounds ? 1.0 : 0.0;
    flag.upload(&mine, 1, s);
    const double* all = c->gather_to_host(flag.p, 1);
    for (size_t r = 0; r < world; ++r) c->has_bounds = c->has_bounds || all[r] != 0.0;
  }
  if (c->has_bounds) {
    std::vector<double> tmp(static_cast<size_t>(c->n));
    for (long long j = 0; j < c->n; ++j) tmp[j] = p->lower_bound ? p->lower_bound[j] : -INFINITY;
    c->lb.upload(tmp.data(), tmp.size(), s);
    EP_CUDA(cudaStreamSynchronize(s));
    for (long long j = 0; j < c->n; ++j) tmp[j] = p->upper_bound ? p->upper_bound[j] : INFINITY;
    c->ub.upload(tmp.data(), tmp.size(), s);
    EP_CUDA(cudaStreamSynchronize(s));
  }
  if (p->delta_vec != nullptr) {  // finiteschemesolve.jl:349-353, 393-396
    double base = INFINITY;
    bool allinf = true;
    for (long long j = 0; j < c->n; ++j) {
      const double v = p->delta_vec[j];
      if (std::isnan(v) || v == 0.0)
        throw std::invalid_argument("`Δ` entries must be nonzero and not NaN (use `Inf` for no damping, a negative entry for forward relaxation)");
      if (std::isfinite(v)) {
        allinf = false;
        base = std::fabs(v) < base ? std::fabs(v) : base;
      }
    }
    c->has_delta_vec = true;
    c->all_delta_inf = allinf;
    if (!allinf) {
      c->delta_base = base;
      std::vector<double> rel(static_cast<size_t>(c->n));
      for (long long j = 0; j < c->n; ++j) rel[j] = p->delta_vec[j] / base;
      c->delta_rel.upload(rel.data(), rel.size(), s);
      EP_CUDA(cudaStreamSynchronize(s));
    }
  }
  // the smallest block size any norm-producing sweep uses; the marching sweeps (sweep.cuh) launch up to
  // 148 x 6 blocks whatever the grid size (a tile can be a few points wide on a narrow grid)
  c->max_blocks = blocks_for(c->n, 128) + 1;
  if (c->max_blocks < 148 * 8) c->max_blocks = 148 * 8;
  c->norm_part.alloc(3 * static_cast<size_t>(c->max_blocks));
  c->norm_out.alloc(4);
  c->color_part.alloc(static_cast<size_t>(blocks_for(c->l)) * c->ncolors);
  c->color_eps.alloc(2 * static_cast<size_t>(c->ncolors));   // steps, then their reciprocals
  c->color_sumsq.alloc(static_cast<size_t>(c->ncolors));
  c->counters.alloc(4);
  c->counters.zero(s);
  EP_CUDA(cudaMallocHost(&c->h_norm, 4 * sizeof(double)));
  if (dist == nullptr) {
    c->gather.alloc(kGatherMax);
    EP_CUDA(cudaMallocHost(&c->h_gather, kGatherMax * sizeof(double)));
  }
  if (dist != nullptr) dist->enable_p2p(c->halo, s);   // collective: every rank builds its Ctx at the same point
  EP_CUDA(cudaStreamSynchronize(s));
  return c;
}

// ---------------------------------------------------------------------------------------
// 1-D linear solver: (block-)tridiagonal PCR
// ---------------------------------------------------------------------------------------
namespace {

template <int F>
int pcr_solve(Ctx& c, const double* planes, const double* b, double* x, DevBuf<double>* work) {
  const long long n = c.l;
  const size_t per = static_cast<size_t>((3 * F * F + F) * n);
  work[0].alloc(per);
  work[1].alloc(per);
  auto view = [&](double* base) { return PcrSystemOut{base, base + F * F * n, base + 2 * F * F * n, base + 3 * F * F * n}; };
  PcrSystem in{planes + 1 * F * F * n, planes, planes + 2 * F * F * n, b};
  int cur = 0;
  for (long long s = 1; s < n; s <<= 1) {
    PcrSystemOut out = view(work[cur].p);
    EP_LAUNCH(k_pcr_step<F>, blocks_for(n), kBlock, 0, c.stream, n, s, in, out);
    in = PcrSystem{out.A, out.B, out.C, out.d};
    cur ^= 1;
  }
  EP_LAUNCH(k_pcr_final<F>, blocks_for(n), kBlock, 0, c.stream, n, in, x, c.nw());
  EP_CUDA(cudaMemcpyAsync(c.h_norm, c.norm_out.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  EP_CUDA(cudaStreamSynchronize(c.stream));
  return c.h_norm[2] != 0.0 ? 1 : 0;
}

struct TridiagSolver final : LinearSolver {
  explicit TridiagSolver(Ctx& c) : c_(c) {}
  int solve(const double* planes, const double* b, double* x, const econpdes_options&, int* iters,
            double* relres) override {
    if (iters) *iters = 0;
    if (relres) *relres = 0.0;
    switch (c_.F) {
      case 1: return pcr_solve<1>(c_, planes, b, x, work_);
      case 2: return pcr_solve<2>(c_, planes, b, x, work_);
      case 3: return pcr_solve<3>(c_, planes, b, x, work_);
      case 4: return pcr_solve<4>(c_, planes, b, x, work_);
      case 5: return pcr_solve<5>(c_, planes, b, x, work_);
      case 6: return pcr_solve<6>(c_, planes, b, x, work_);
      case 7: return pcr_solve<7>(c_, planes, b, x, work_);
      case 8: return pcr_solve<8>(c_, planes, b, x, work_);
      default: throw std::invalid_argument("block-tridiagonal solver supports up to 8 unknowns");
    }
  }
  Ctx& c_;
  DevBuf<double> work_[2];
};

}  // namespace

std::unique_ptr<LinearSolver> make_tridiag_solver(Ctx& c) { return std::make_unique<TridiagSolver>(c); }

// ---------------------------------------------------------------------------------------
// Solver
// ---------------------------------------------------------------------------------------
namespace {

// ||ydot .* mask|| / sum(mask), mask = lb + eps <= y <= ub - eps  (finiteschemesolve.jl:410-413)
__global__ void __launch_bounds__(kBlock) k_masked_norm(long long n, const double* __restrict__ ydot,
                                                        const double* __restrict__ y, const double* __restrict__ lb,
                                                        const double* __restrict__ ub, NormWork nw) {
  __shared__ double scratch[96];
  const long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  double s = 0.0, cnt = 0.0;
  int nanf = 0;
  if (j < n) {
    const double e = 2.220446049250313e-16;
    const bool in = (lb[j] + e <= y[j]) && (y[j] <= ub[j] - e);
    if (in) {
      cnt = 1.0;
      const double r = ydot[j];
      if (r != r) nanf = 1; else s = r * r;
    }
  }
  // reuse the (sum, max, flag) reduction: the count is summed in a second pass below
  grid_norm_finish(s, cnt, nanf, nw, scratch);
}

__global__ void __launch_bounds__(kBlock) k_count_mask(long long n, const double* __restrict__ y,
                                                       const double* __restrict__ lb, const double* __restrict__ ub,
                                                       NormWork nw) {
  __shared__ double scratch[96];
  const long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  double cnt = 0.0;
  if (j < n) {
    const double e = 2.220446049250313e-16;
    if ((lb[j] + e <= y[j]) && (y[j] <= ub[j] - e)) cnt = 1.0;
  }
  grid_norm_finish(cnt, 0.0, 0, nw, scratch);
}

}  // namespace

Solver::Solver(std::unique_ptr<Ctx> ctx) : c_(std::move(ctx)), ops_(model_ops(c_->model_id)) {
  const size_t n = static_cast<size_t>(c_->n);
  ypost_.pad = y_.pad = ynew_.pad = static_cast<size_t>(c_->halo);   // fields a stencil reads carry ghost planes
  ypost_.alloc(n); y_.alloc(n); ynew_.alloc(n); R_.alloc(n); s_.alloc(n); G_.alloc(n);
  planes_.alloc(static_cast<size_t>(c_->nk) * c_->F * c_->F * c_->l);
  // the marching Jacobian sweep never writes the planes of box slots the model does not read (exact zeros)
  EP_CUDA(cudaMemset(planes_.p, 0, planes_.n * sizeof(double)));
  EP_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
  lin_ = c_->N == 1 ? make_tridiag_solver(*c_) : make_krylov_solver(*c_);
}

RowSpec Solver::row_spec(const double* ypost, double delta) const {
  RowSpec rs;
  rs.ypost = ypost;
  rs.delta = delta;
  rs.rdelta = (ypost != nullptr && std::isfinite(delta)) ? 1.0 / delta : 0.0;
  rs.delta_rel = (ypost != nullptr && c_->has_delta_vec && !c_->all_delta_inf && std::isfinite(delta)) ? c_->delta_rel.p : nullptr;
  rs.is_alg = c_->is_alg.p;
  rs.lb = c_->lb.p;
  rs.ub = c_->ub.p;
  rs.has_bounds = (ypost != nullptr && c_->has_bounds) ? 1 : 0;
  return rs;
}

Norms Solver::read_norms() {
  // (sum of squares, max, NaN flag) of every rank, folded in rank order (one rank: itself)
  const double* g = c_->gather_to_host(c_->norm_out.p, 3);
  const int world = c_->dist != nullptr ? c_->dist->world() : 1;
  Norms out{0.0, 0.0, false, false};
  for (int r = 0; r < world; ++r) {
    out.sumsq += g[3 * r];
    out.maxabs = g[3 * r + 1] > out.maxabs ? g[3 * r + 1] : out.maxabs;
    const int flags = static_cast<int>(g[3 * r + 2]);   // kFlagNaN | kFlagUntame (sweep.cuh)
    out.any_nan = out.any_nan || (flags & kFlagNaN) != 0;
    out.untame = out.untame || (flags & kFlagUntame) != 0;
  }
  return out;
}

// The sweep first runs in its tame form (stencil quotients without the NaN-select, sweep.cuh), which also
// reports whether every unknown it read was tame; if not -- a wild Newton iterate -- the output is discarded
// and recomputed by the general form.  `last_tame_` tells the Jacobian of the same iterate which form to take.
Norms Solver::residual_with_norms(const double* y, const RowSpec& rs, double* out) {
  const bool try_tame = c_->geom_tame;
  ops_->residual(*c_, y, rs, out, true, try_tame);
  Norms n = read_norms();
  if (try_tame && n.untame) {
    ops_->residual(*c_, y, rs, out, true, false);
    n = read_norms();
    n.untame = true;
  }
  last_tame_ = try_tame && !n.untame;
  return n;
}

Norms Solver::eval_G(const double* y, double* out) {
  ProfScope ps("residual G", c_->stream);
  c_->exchange_halo(y);
  return residual_with_norms(y, row_spec(nullptr, INFINITY), out);
}

Norms Solver::eval_R(const double* y, const double* ypost, double delta, double* out) {
  ProfScope ps("residual R", c_->stream);
  c_->exchange_halo(y);
  return residual_with_norms(y, row_spec(ypost, delta), out);
}

void Solver::jacobian(const double* y, const double* ypost, double delta, int eps_policy, bool tame) {
  ProfScope ps("jacobian", c_->stream);
  c_->exchange_halo(y);
  ops_->jacobian(*c_, y, row_spec(ypost, delta), eps_policy, planes_.p, nullptr, tame && c_->geom_tame);
}

namespace {
__global__ void __launch_bounds__(kBlock) k_tame_check(long long n, const double* __restrict__ y, NormWork nw) {
  __shared__ double scratch[96];
  int flag = 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += stride)
    if (!is_tame(y[j])) flag = kFlagUntame;
  grid_norm_finish(0.0, 0.0, flag, nw, scratch);
}
}  // namespace

// Is every unknown of y tame (sweep.cuh)?  For callers that assemble a Jacobian without the residual sweep
// of the same iterate in front of it (the handle API).
bool Solver::check_tame(const double* y) {
  if (!c_->geom_tame) return false;
  EP_LAUNCH(k_tame_check, sweep_blocks(c_->n), kBlock, 0, c_->stream, c_->n, y, c_->nw());
  return !read_norms().untame;
}

double Solver::masked_norm(const double* ydot, const double* y) {
  EP_LAUNCH(k_masked_norm, blocks_for(c_->n), kBlock, 0, c_->stream, c_->n, ydot, y, c_->lb.p, c_->ub.p, c_->nw());
  Norms a = read_norms();
  EP_LAUNCH(k_count_mask, blocks_for(c_->n), kBlock, 0, c_->stream, c_->n, y, c_->lb.p, c_->ub.p, c_->nw());
  Norms b = read_norms();
  if (a.any_nan) return std::numeric_limits<double>::quiet_NaN();
  return std::sqrt(a.sumsq) / b.sumsq;
}

// Diagnostics of src/monotonicity.jl as reductions over the stencil-format Jacobian.
// counts: [0..F) rows checked per unknown, [F..2F) rows whose residual diagonal J_ii - 1/Δ_i is opposite
// in sign to Δ_i (_check_sign_convention!, :57-98), [2F] same-unknown spatial off-diagonal entries above
// tol (_check_monotonicity!, :100-133), maxbits: the largest such entry (positive doubles order like their
// bit patterns).  Integer atomics: the result does not depend on scheduling.
static __global__ void __launch_bounds__(kBlock) k_jacobian_diagnostics(long long l, int F, int nk, int kcenter,
                                                                  const double* __restrict__ planes, double delta,
                                                                  const double* __restrict__ delta_rel,
                                                                  const unsigned char* __restrict__ is_alg, double tol,
                                                                  int signs, int stencils, int* __restrict__ counts,
                                                                  unsigned long long* __restrict__ maxbits) {
  const long long pt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pt >= l) return;
  for (int k = 0; k < F; ++k) {
    const long long row = k * l + pt;
    const bool alg = is_alg != nullptr && is_alg[row] != 0;
    const double di = delta_rel != nullptr ? delta * delta_rel[row] : delta;
    if (signs && !alg && isfinite(di)) {
      const double d = planes[(static_cast<long long>(kcenter * F + k) * F + k) * l + pt] - 1 / di;
      if (fabs(d) > tol) {
        atomicAdd(&counts[k], 1);
        const bool same = (d > 0 && di > 0) || (d < 0 && di < 0);
        if (!same) atomicAdd(&counts[F + k], 1);
      }
    }
    if (stencils) {
      for (int ks = 0; ks < nk; ++ks) {
        if (ks == kcenter) continue;
        const double v = planes[(static_cast<long long>(ks * F + k) * F + k) * l + pt];
        if (v > tol) {
          atomicAdd(&counts[2 * F], 1);
          atomicMax(maxbits, static_cast<unsigned long long>(__double_as_longlong(v)));
        }
      }
    }
  }
}

void Solver::jacobian_diagnostics(double delta, const econpdes_options& opt) {
  Ctx& c = *c_;
  // Sign convention: checked once, on the first Jacobian assembled with a finite step -- a one-shot solve
  // (infinite Δ) is sign-agnostic and leaves the check pending (:70-72).  Stencil monotonicity (opt-in): on
  // the first assembled Jacobian whatever the step (the reference repeats it on later Jacobians until
  // `monotonicity_max_warnings` distinct entries were reported; the device reports the first one's count).
  if (c.dist != nullptr) return;
  const bool do_sign = !sign_checked_ && std::isfinite(delta);
  const bool do_stencils = opt.check_monotonicity != 0 && !stencils_checked_;
  if (!do_sign && !do_stencils) return;
  if (do_sign) sign_checked_ = true;
  if (do_stencils) stencils_checked_ = true;
  DevBuf<int> counts;
  DevBuf<unsigned long long> mx;
  counts.alloc(static_cast<size_t>(2 * c.F + 1));
  mx.alloc(1);
  counts.zero(c.stream);
  mx.zero(c.stream);
  const double* drel = (c.has_delta_vec && !c.all_delta_inf) ? c.delta_rel.p : nullptr;
  const int kcenter = c.box ? (ipow3(c.N) - 1) / 2 : 0;
  EP_LAUNCH(k_jacobian_diagnostics, blocks_for(c.l), kBlock, 0, c.stream, c.l, c.F, c.nk, kcenter, planes_.p, delta, drel,
            c.is_alg.p, opt.monotonicity_tol, do_sign ? 1 : 0, do_stencils ? 1 : 0, counts.p, mx.p);
  std::vector<int> h(static_cast<size_t>(2 * c.F + 1));
  unsigned long long hm = 0;
  counts.download(h.data(), h.size(), c.stream);
  mx.download(&hm, 1, c.stream);
  EP_CUDA(cudaStreamSynchronize(c.stream));
  if (do_sign) {
    int total_checked = 0;
    for (int k = 0; k < c.F; ++k) {
      total_checked += h[k];
      if (h[k] > 0 && h[c.F + k] == h[k]) sign_flags_ |= 1 << k;
    }
    bool all_wrong = total_checked > 0;
    for (int k = 0; k < c.F; ++k) all_wrong = all_wrong && h[c.F + k] == h[k];
    if (drel == nullptr && delta > 0 && all_wrong) sign_flags_ |= ECONPDES_SIGN_WHOLE_SYSTEM;
  }
  if (do_stencils) {
    nonmonotone_entries_ = h[2 * c.F];
    std::memcpy(&nonmonotone_max_, &hm, sizeof(double));
  }
}

// NLsolve newton_ with the Static line search, wrapped like _newtonstep (:125-147).
static bool opt_singular_retry() {
  static const bool on = [] { const char* e = std::getenv("ECONPDES_SINGULAR_RETRY"); return !(e && e[0] == '0'); }();
  return on;
}

int Solver::regularized_step() {
  ProfScope ps("singular_retry", c_->stream);
  const int F = c_->F;
  double lambda = 0.0;
  int iters = 0;
  auto run = [&](auto ntag) {
    constexpr int NN = decltype(ntag)::value;
    const GridView<NN> gv = c_->grid_view<NN>();
    return c_->box ? regularized_normal_solve<NN, true>(c_->stream, gv, F, planes_.p, R_.p, s_.p, neq_, &lambda, &iters)
                   : regularized_normal_solve<NN, false>(c_->stream, gv, F, planes_.p, R_.p, s_.p, neq_, &lambda, &iters);
  };
  switch (c_->N) {
    case 1: return run(std::integral_constant<int, 1>{});
    case 2: return run(std::integral_constant<int, 2>{});
    case 3: return run(std::integral_constant<int, 3>{});
    case 4: return run(std::integral_constant<int, 4>{});
    default: return 1;
  }
}

NewtonOutcome Solver::newton(const double* ypost, double delta, double ftol, int maxit, const econpdes_options& opt,
                             double* y_out) {
  const long long n = c_->n;
  NewtonOutcome out{0.0, 0, 0, false};
  double* x = y_.p;
  double* xn = ynew_.p;
  EP_CUDA(cudaMemcpyAsync(x, ypost, n * sizeof(double), cudaMemcpyDeviceToDevice, c_->stream));
  lin_->new_step();
  Norms nr = eval_R(x, ypost, delta, R_.p);
  bool stopped = nr.any_nan;
  bool fconv = !nr.any_nan && nr.maxabs <= ftol;
  bool xconv = false;
  while (!stopped && !fconv && !xconv && out.iters < maxit) {
    ++out.iters;
    if (c_->has_bounds && c_->dist == nullptr &&
        ((!sign_checked_ && std::isfinite(delta)) || (opt.check_monotonicity != 0 && !stencils_checked_))) {
      // fd_mcp_jac (finiteschemesolve.jl:205-214) hands the checker the Jacobian of the UNCLAMPED G_helper!
      // residual: rows at an active bound would otherwise show diagonal 1 and no off-diagonal entries
      ProfScope ps("jacobian", c_->stream);
      RowSpec rs = row_spec(ypost, delta);
      rs.has_bounds = 0;
      ops_->jacobian(*c_, x, rs, opt.eps_policy, planes_.p, nullptr, last_tame_ && c_->geom_tame);
      jacobian_diagnostics(delta, opt);
    }
    jacobian(x, ypost, delta, opt.eps_policy, last_tame_);   // eval_R(x) has just vouched for x
    jacobian_diagnostics(delta, opt);
    int li = 0;
    double rr = 0.0;
    const int fail = lin_->solve(planes_.p, R_.p, s_.p, opt, &li, &rr);
    out.lin_iters += li;
    ++out.lin_solves;
    if (rr > 1e-11) ++out.lin_loose;
    if (rr > out.lin_max_relres) out.lin_max_relres = rr;
    // SingularException: NLsolve first retries with the regularised normal equations (newton.jl; restated in
    // oracle/econpdes_oracle.py:nlsolve_newton) and only a failure of THAT solve leaves the Newton iteration
    bool give_up = fail != 0;
    if (fail && c_->dist == nullptr && opt_singular_retry() && regularized_step() == 0) {
      ++out.singular_retries;
      give_up = false;
    }
    if (give_up) {  // -> (y0, Inf), finiteschemesolve.jl:142-146
      out.singular = true;
      out.nl = INFINITY;
      EP_CUDA(cudaMemcpyAsync(y_out, ypost, n * sizeof(double), cudaMemcpyDeviceToDevice, c_->stream));
      return out;
    }
    EP_LAUNCH(k_newton_update, sweep_blocks(n), kBlock, 0, c_->stream, n, x, s_.p, xn, c_->nw());
    Norms nx = read_norms();
    std::swap(x, xn);
    nr = eval_R(x, ypost, delta, R_.p);
    xconv = !nx.any_nan && nx.maxabs <= 0.0;
    fconv = !nr.any_nan && nr.maxabs <= ftol;
    stopped = nx.any_nan || nr.any_nan;
  }
  out.nl = nr.norm2_over_n(c_->n_glob);
  if (x != y_out) EP_CUDA(cudaMemcpyAsync(y_out, x, n * sizeof(double), cudaMemcpyDeviceToDevice, c_->stream));
  return out;
}

int Solver::solve(const econpdes_options& opt, econpdes_result* res) {
  Ctx& c = *c_;
  if (ops_->time_slot() >= 0)
    throw std::invalid_argument("the PDE function takes a time argument `(state, u, t)`: pass a time grid `τs`");
  const long long n = c.n;
  cudaStream_t s = c.stream;
  const int64_t launches0 = g_kernel_launches.load();
  cudaEvent_t e0, e1;
  EP_CUDA(cudaEventCreate(&e0));
  EP_CUDA(cudaEventCreate(&e1));
  EP_CUDA(cudaEventRecord(e0, s));
  res->log_len = 0;
  res->n_outer = 0;
  res->n_newton = 0;
  res->n_linear_iters = 0;
  res->n_linear_solves = 0;
  res->n_linear_loose = 0;
  res->lin_max_relres = 0.0;
  auto count_linear = [&](const NewtonOutcome& o) {
    res->n_linear_solves += o.lin_solves;
    res->n_linear_loose += o.lin_loose;
    if (o.lin_max_relres > res->lin_max_relres) res->lin_max_relres = o.lin_max_relres;
  };
  res->exit_reason = ECONPDES_EXIT_CONVERGED;

  if (!c.has_delta_vec && (std::isnan(opt.delta) || opt.delta == 0.0))
    throw std::invalid_argument("`Δ` entries must be nonzero and not NaN (use `Inf` for no damping, a negative entry for forward relaxation)");

  const long long ng = c.n_glob;   // norms are over the whole grid; copies move this device's part
  DevBuf<double> ytrial;
  ytrial.pad = static_cast<size_t>(c.halo);
  ytrial.alloc(static_cast<size_t>(n));
  EP_CUDA(cudaMemcpyAsync(ypost_.p, c.guess.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  double r = eval_G(ypost_.p, G_.p).norm2_over_n(ng);
  int rc = ECONPDES_OK;
  auto push_log = [&](int iter, int status, int nit, long long lit, double delta, double resid) {
    if (res->log != nullptr && res->log_len < res->log_capacity) {
      econpdes_log_row& row = res->log[res->log_len];
      row.iter = iter; row.status = status; row.newton_iters = nit; row.linear_iters = static_cast<int32_t>(lit);
      row.delta = delta; row.residual = resid;
    }
    ++res->log_len;
  };
  const bool delta_inf = c.has_delta_vec ? c.all_delta_inf : std::isinf(opt.delta);
  if (std::isnan(r)) {
    set_error("G! returns NaN with the initial value");   // finiteschemesolve.jl:372
    rc = ECONPDES_E_NAN_GUESS;
  } else if (r <= opt.abstol) {
    res->exit_reason = ECONPDES_EXIT_ALREADY_ZERO;        // :373-376
  } else if (delta_inf) {                                  // :377-389
    NewtonOutcome o = newton(ypost_.p, INFINITY, opt.abstol, opt.maxiters, opt, ytrial.p);
    res->n_newton = o.iters;
    res->n_linear_iters = o.lin_iters;
    count_linear(o);
    EP_CUDA(cudaMemcpyAsync(ypost_.p, ytrial.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    r = o.nl;
    res->exit_reason = r <= opt.abstol ? ECONPDES_EXIT_CONVERGED : ECONPDES_EXIT_INF_STEP_FAILED;
  } else {
    SerController ctl;
    double drel = 1.0;
    if (c.has_delta_vec) {
      ctl.init(c.delta_base, opt.scale, opt.max_delta);
    } else {
      ctl.init(std::fabs(opt.delta), opt.scale, opt.max_delta);
      drel = opt.delta < 0 ? -1.0 : 1.0;
    }
    int iter = 0;
    while (iter < opt.maxiters && ctl.delta >= opt.min_delta && r > opt.abstol) {   // :406
      ++iter;
      NewtonOutcome o = newton(ypost_.p, ctl.delta * drel, opt.inner_abstol, opt.inner_maxiters, opt, ytrial.p);
      res->n_newton += o.iters;
      res->n_linear_iters += o.lin_iters;
      count_linear(o);
      const double old_r = r;
      Norms ngn = eval_G(ytrial.p, G_.p);
      double rn = c.has_bounds ? masked_norm(G_.p, ytrial.p) : ngn.norm2_over_n(ng);
      if (std::isnan(rn)) rn = INFINITY;                                              // :417
      if (o.nl <= opt.inner_abstol && rn <= opt.max_residual_growth * old_r) {        // :418
        push_log(iter, ECONPDES_STEP_ACCEPTED, o.iters, o.lin_iters, ctl.delta, rn);
        ctl.accept(old_r, rn);
        EP_CUDA(cudaMemcpyAsync(ypost_.p, ytrial.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
        r = rn;
      } else {
        int status;
        if (o.nl <= opt.inner_abstol) status = ECONPDES_STEP_REJECTED_GROWTH;         // :425
        else if (std::isnan(o.nl)) status = ECONPDES_STEP_REJECTED_NAN;
        else if (std::isinf(o.nl)) status = ECONPDES_STEP_REJECTED_SINGULAR;
        else status = ECONPDES_STEP_REJECTED_INNER;                                   // :439
        // growth rejections carry the factor the reference prints ("residual grew 12.3x", :433-435); others NaN
        push_log(iter, status, o.iters, o.lin_iters, ctl.delta,
                 status == ECONPDES_STEP_REJECTED_GROWTH ? rn / old_r : std::numeric_limits<double>::quiet_NaN());
        ctl.reject();
      }
    }
    res->n_outer = iter;
    if (r <= opt.abstol) res->exit_reason = ECONPDES_EXIT_CONVERGED;
    else if (ctl.delta < opt.min_delta) res->exit_reason = ECONPDES_EXIT_BELOW_MIN_DELTA;
    else res->exit_reason = ECONPDES_EXIT_MAXITERS;
  }
  DevBuf<double> saved;
  if (rc == ECONPDES_OK && res->saved != nullptr && c.nsaved > 0) {
    saved.alloc(static_cast<size_t>(c.nsaved) * c.l);
    ops_->saved(c, ypost_.p, saved.p);
  }
  EP_CUDA(cudaEventRecord(e1, s));
  EP_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  EP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  res->device_ms = ms;
  res->sign_flags = sign_flags_;
  res->nonmonotone_entries = nonmonotone_entries_;
  res->nonmonotone_max = nonmonotone_max_;
  res->residual_norm = r;
  res->converged = (r <= opt.abstol) ? 1 : 0;
  if (res->log_len > res->log_capacity) res->log_len = res->log_capacity;
  if (res->solution != nullptr) ypost_.download(res->solution, static_cast<size_t>(n), s);
  if (saved.p != nullptr) saved.download(res->saved, saved.n, s);
  EP_CUDA(cudaStreamSynchronize(s));
  res->kernel_launches = g_kernel_launches.load() - launches0;
  if (c.dist != nullptr) c.dist->check_p2p_error(s);
  Prof::report(c.dist != nullptr ? c.dist->rank() : 0);
  return rc;
}

// Backward induction over a time grid (src/pdesolve.jl:191-211): the state stays on the device, each
// time slice is copied out as it is produced.
int Solver::solve_backward(const econpdes_options& opt, const double* taus, int ntaus, double* solution,
                           double* saved_out, double* residual_norms, int32_t* newton_iters, double* device_ms,
                           int64_t* launches) {
  Ctx& c = *c_;
  if (c.dist != nullptr) throw std::invalid_argument("the time-dependent solve is single-device");
  if (taus == nullptr || ntaus < 1) throw std::invalid_argument("the time grid `τs` must contain at least one time");
  for (int i = 0; i < ntaus; ++i) {
    if (!std::isfinite(taus[i])) throw std::invalid_argument("the time grid `τs` must contain finite real numbers");
    if (i > 0 && !(taus[i] > taus[i - 1])) throw std::invalid_argument("the time grid `τs` must be strictly increasing");
  }
  if (solution == nullptr || residual_norms == nullptr) throw std::invalid_argument("null output");
  const long long n = c.n;
  cudaStream_t s = c.stream;
  const int64_t launches0 = g_kernel_launches.load();
  cudaEvent_t e0, e1;
  EP_CUDA(cudaEventCreate(&e0));
  EP_CUDA(cudaEventCreate(&e1));
  EP_CUDA(cudaEventRecord(e0, s));
  DevBuf<double> ynext, saved;
  ynext.alloc(static_cast<size_t>(n));
  if (saved_out != nullptr && c.nsaved > 0) saved.alloc(static_cast<size_t>(c.nsaved) * c.l);
  EP_CUDA(cudaMemcpyAsync(ypost_.p, c.guess.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  for (int i = 0; i < ntaus; ++i) residual_norms[i] = 0.0;
  const int tslot = ops_->time_slot();
  for (int it = ntaus - 1; it >= 0; --it) {
    // pde_onestep = (state, u) -> pde(state, u, τs[iτ]) (src/pdesolve.jl:195): the saved outputs of this
    // time and the step towards the previous one are both evaluated at τs[iτ]
    if (tslot >= 0) c.prm.v[tslot] = taus[it];
    ypost_.download(solution + static_cast<size_t>(it) * n, static_cast<size_t>(n), s);
    if (saved.p != nullptr) {
      ops_->saved(c, ypost_.p, saved.p);
      saved.download(saved_out + static_cast<size_t>(it) * saved.n, saved.n, s);
    }
    if (it > 0) {
      NewtonOutcome o = newton(ypost_.p, taus[it] - taus[it - 1], opt.inner_abstol, opt.inner_maxiters, opt, ynext.p);
      residual_norms[it] = o.nl;
      if (newton_iters != nullptr) newton_iters[it] = o.iters;
      EP_CUDA(cudaMemcpyAsync(ypost_.p, ynext.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    } else if (newton_iters != nullptr) {
      newton_iters[0] = 0;
    }
  }
  EP_CUDA(cudaEventRecord(e1, s));
  EP_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  EP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (device_ms != nullptr) *device_ms = ms;
  if (launches != nullptr) *launches = g_kernel_launches.load() - launches0;
  return ECONPDES_OK;
}

// Map from CSC nonzeros of local_stencil_jacobian to stencil planes (host, once).
void Solver::csc_gather(double* nzval_dev) {
  Ctx& c = *c_;
  if (csc_map_.p == nullptr) {
    int64_t size[kMaxDims];
    for (int d = 0; d < c.N; ++d) size[d] = c.S[d];
    nnz_ = econpdes_stencil_nnz(c.N, size, c.F);
    std::vector<int64_t> colptr(static_cast<size_t>(c.n) + 1), rowval(static_cast<size_t>(nnz_));
    econpdes_stencil_pattern(c.N, size, c.F, colptr.data(), rowval.data());
    std::vector<long long> map(static_cast<size_t>(nnz_));
    for (long long j = 0; j < c.n; ++j) {
      const long long cpt = j % c.l, fin = j / c.l;
      for (int64_t e = colptr[j] - 1; e < colptr[j + 1] - 1; ++e) {
        const long long row = rowval[e] - 1, rpt = row % c.l, fout = row / c.l;
        long long a = cpt, b = rpt;
        int kbox = 0, nz = 0, kstar = 0, p3 = 1;
        for (int d = 0; d < c.N; ++d) {
          const int off = static_cast<int>(a % c.S[d] - b % c.S[d]);
          a /= c.S[d];
          b /= c.S[d];
          kbox += (off + 1) * p3;
          p3 *= 3;
          if (off != 0) { ++nz; kstar = 1 + 2 * d + (off > 0 ? 1 : 0); }
        }
        long long k;
        if (c.box) k = kbox; else k = nz == 0 ? 0 : (nz == 1 ? kstar : -1);
        map[e] = k < 0 ? -1 : ((k * c.F + fin) * c.F + fout) * c.l + rpt;
      }
    }
    csc_map_.upload(map.data(), map.size(), c.stream);
    EP_CUDA(cudaStreamSynchronize(c.stream));
  }
  EP_LAUNCH(k_gather_csc, blocks_for(nnz_), kBlock, 0, c.stream, nnz_, csc_map_.p, planes_.p, nzval_dev);
}

}  // namespace econpdes
