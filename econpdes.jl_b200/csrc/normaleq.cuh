// NLsolve's fallback for a singular Jacobian (NLsolve.jl newton.jl, reached from src/finiteschemesolve.jl:135-146;
// Nocedal & Wright p. 289): when `J \ f` throws, the Newton direction becomes
//
//     p = -(J'J + lambda I)^{-1} J'f,      lambda = 1e6 * sqrt(n * eps) * ||J'J||_1 .
//
// Here J is the stencil-format Jacobian (planes[((k F + fin) F + fout) l + pt], entries that point outside the
// grid hold 0).  lambda is so large against ||J'J|| (cond(J'J + lambda I) <= 1 + 1/(1e6 sqrt(n eps)): about 2 for
// 4 000 unknowns, 50 for 2) that plain conjugate gradients on the regularised normal equations converges to
// round-off in a few dozen iterations; the operator is applied matrix-free as J'(J v) + lambda v.
// A cold path: kernels are written for any N <= 4, box or star stencil and any F, not for speed.
#pragma once
#include <cmath>
#include <cstring>

#include "engine.cuh"
#include "kernels.cuh"

namespace econpdes {

template <int N, bool BOX>
__device__ __forceinline__ int stencil_slot_of(const int* o) {
  if (BOX) {
    int k = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      if (o[d] < -1 || o[d] > 1) return -1;
      k += (o[d] + 1) * p3(d);
    }
    return k;
  }
  int k = 0, nz = 0;
#pragma unroll
  for (int d = 0; d < N; ++d) {
    if (o[d] == 0) continue;
    if (o[d] < -1 || o[d] > 1) return -1;
    ++nz;
    k = 1 + 2 * d + (o[d] > 0 ? 1 : 0);
  }
  return nz > 1 ? -1 : k;
}

// y = J x (transpose = 0) or y = J' x (transpose = 1); one thread per (point, unknown)
template <int N, bool BOX>
__global__ void __launch_bounds__(kBlock) k_stencil_matvec(GridView<N> gv, int F, const double* __restrict__ A,
                                                           const double* __restrict__ x, double* __restrict__ y,
                                                           int transpose) {
  using St = Stencil<N, BOX>;
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= gv.l * F) return;
  const long long pt = e % gv.l;
  const int f = static_cast<int>(e / gv.l);
  long long idx[N];
  unravel<N>(pt, gv, idx);
  double acc = 0.0;
  for (int k = 0; k < St::NK; ++k) {
    // forward: neighbour q = pt + off_k supplies x; transpose: row r = pt - off_k supplies both the entry and x
    long long q = pt;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      const int o = transpose ? -St::off(k, d) : St::off(k, d);
      const long long i = idx[d] + o;
      if (i < 0 || i >= gv.S[d]) ok = false;
      q += o * gv.stride[d];
    }
    if (!ok) continue;
    for (int g = 0; g < F; ++g) {
      // forward: row (pt, fout = f), column (q, fin = g);  transpose: row (q, fout = g), column (pt, fin = f)
      const double a = transpose ? A[(static_cast<long long>(k * F + f) * F + g) * gv.l + q]
                                 : A[(static_cast<long long>(k * F + g) * F + f) * gv.l + pt];
      acc += a * x[g * gv.l + q];
    }
  }
  y[e] = acc;
}

// ||J'J||_1: the largest column sum of |J'J|.  One thread per column (pt, fin): every entry (pt + delta, fi) of
// that column, delta in {-2..2}^N, is the sum over the rows r = pt - off_kj that see both columns.
// out[0]: bits of the maximum (non-negative doubles order like their bit patterns), out[1]: non-zero if a sum is NaN.
template <int N, bool BOX>
__global__ void __launch_bounds__(kBlock) k_normal_onenorm(GridView<N> gv, int F, const double* __restrict__ A,
                                                           unsigned long long* __restrict__ out) {
  using St = Stencil<N, BOX>;
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= gv.l * F) return;
  const long long pt = e % gv.l;
  const int fin = static_cast<int>(e / gv.l);
  long long idx[N];
  unravel<N>(pt, gv, idx);
  int n5 = 1;
#pragma unroll
  for (int d = 0; d < N; ++d) n5 *= 5;
  double total = 0.0;
  for (int m = 0; m < n5; ++m) {
    int delta[N], mm = m;
    bool in_grid = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      delta[d] = mm % 5 - 2;
      mm /= 5;
      const long long i = idx[d] + delta[d];
      if (i < 0 || i >= gv.S[d]) in_grid = false;
    }
    if (!in_grid) continue;
    for (int fi = 0; fi < F; ++fi) {
      double acc = 0.0;
      for (int kj = 0; kj < St::NK; ++kj) {
        long long r = pt;
        bool ok = true;
        int oi[N];
#pragma unroll
        for (int d = 0; d < N; ++d) {
          const int oj = St::off(kj, d);
          const long long i = idx[d] - oj;
          if (i < 0 || i >= gv.S[d]) ok = false;
          r -= oj * gv.stride[d];
          oi[d] = oj + delta[d];
        }
        if (!ok) continue;
        const int ki = stencil_slot_of<N, BOX>(oi);
        if (ki < 0) continue;
        for (int fout = 0; fout < F; ++fout)
          acc += A[(static_cast<long long>(ki * F + fi) * F + fout) * gv.l + r] *
                 A[(static_cast<long long>(kj * F + fin) * F + fout) * gv.l + r];
      }
      total += fabs(acc);
    }
  }
  if (total != total) atomicOr(&out[1], 1ull);
  else atomicMax(&out[0], static_cast<unsigned long long>(__double_as_longlong(total)));
}

// partial[b] = sum over block b's strided share of a[i] * b[i]; fixed grid, so the sum order does not depend on n's
// divisibility by anything but the launch shape
constexpr int kDotBlocks = 296;
__global__ void __launch_bounds__(kBlock) k_dot_partial(long long n, const double* __restrict__ a,
                                                        const double* __restrict__ b, double* __restrict__ partial) {
  __shared__ double sm[kBlock / 32];
  double s = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    s += a[i] * b[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kBlock / 32; ++w) t += sm[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void k_dot_finish(int nb, const double* __restrict__ partial, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < nb; ++b) t += partial[b];
    out[0] = t;
  }
}
// y = a x + b y
__global__ void __launch_bounds__(kBlock) k_axpby(long long n, double a, const double* __restrict__ x, double b,
                                                  double* __restrict__ y) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i] + b * y[i];
}

// s = (J'J + lambda I)^{-1} J'f.  Returns 0, or 1 when the Jacobian or the residual is not finite (the reference
// rethrows then) or CG fails to converge.  *lambda_out / *iters_out for the log.
template <int N, bool BOX>
int regularized_normal_solve(cudaStream_t st, const GridView<N>& gv, int F, const double* planes, const double* f,
                             double* s, NormalEqWork& wk, double* lambda_out, int* iters_out) {
  const long long n = gv.l * F;
  const int blocks = static_cast<int>((n + kBlock - 1) / kBlock);
  wk.b.alloc(n); wk.r.alloc(n); wk.p.alloc(n); wk.q.alloc(n); wk.w.alloc(n);
  wk.partial.alloc(kDotBlocks); wk.scalar.alloc(1); wk.bits.alloc(2);
  auto dot = [&](const double* a, const double* b) {
    EP_LAUNCH(k_dot_partial, kDotBlocks, kBlock, 0, st, n, a, b, wk.partial.p);
    EP_LAUNCH(k_dot_finish, 1, 32, 0, st, kDotBlocks, wk.partial.p, wk.scalar.p);
    double h = 0.0;
    EP_CUDA(cudaMemcpyAsync(&h, wk.scalar.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    EP_CUDA(cudaStreamSynchronize(st));
    return h;
  };
  EP_CUDA(cudaMemsetAsync(wk.bits.p, 0, 2 * sizeof(unsigned long long), st));
  EP_LAUNCH((k_normal_onenorm<N, BOX>), blocks, kBlock, 0, st, gv, F, planes, wk.bits.p);
  unsigned long long hb[2];
  EP_CUDA(cudaMemcpyAsync(hb, wk.bits.p, sizeof(hb), cudaMemcpyDeviceToHost, st));
  EP_CUDA(cudaStreamSynchronize(st));
  double onenorm;
  std::memcpy(&onenorm, &hb[0], sizeof(double));
  if (hb[1] != 0ull || !std::isfinite(onenorm)) return 1;
  const double lambda = 1e6 * std::sqrt(static_cast<double>(n) * 2.220446049250313e-16) * onenorm;
  if (lambda_out) *lambda_out = lambda;
  if (!(lambda > 0.0)) return 1;   // J = 0: J'J + lambda I is the zero matrix, singular again
  EP_LAUNCH((k_stencil_matvec<N, BOX>), blocks, kBlock, 0, st, gv, F, planes, f, wk.b.p, 1);   // b = J'f
  EP_CUDA(cudaMemsetAsync(s, 0, n * sizeof(double), st));
  EP_CUDA(cudaMemcpyAsync(wk.r.p, wk.b.p, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  EP_CUDA(cudaMemcpyAsync(wk.p.p, wk.b.p, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  double rs = dot(wk.r.p, wk.r.p);
  const double b2 = rs;
  if (!std::isfinite(b2)) return 1;
  int it = 0;
  if (b2 > 0.0) {
    for (; it < 500; ++it) {
      const double rs_old = rs;
      EP_LAUNCH((k_stencil_matvec<N, BOX>), blocks, kBlock, 0, st, gv, F, planes, wk.p.p, wk.q.p, 0);   // q = J p
      EP_LAUNCH((k_stencil_matvec<N, BOX>), blocks, kBlock, 0, st, gv, F, planes, wk.q.p, wk.w.p, 1);   // w = J'q
      EP_LAUNCH(k_axpby, blocks, kBlock, 0, st, n, lambda, wk.p.p, 1.0, wk.w.p);                        // w += lambda p
      const double pw = dot(wk.p.p, wk.w.p);
      if (!(pw > 0.0) || !std::isfinite(pw)) return 1;
      const double alpha = rs / pw;
      EP_LAUNCH(k_axpby, blocks, kBlock, 0, st, n, alpha, wk.p.p, 1.0, s);
      EP_LAUNCH(k_axpby, blocks, kBlock, 0, st, n, -alpha, wk.w.p, 1.0, wk.r.p);
      const double rs_new = dot(wk.r.p, wk.r.p);
      if (!std::isfinite(rs_new)) return 1;
      rs = rs_new;
      if (rs_new <= 1e-28 * b2) { ++it; break; }   // relative residual 1e-14
      EP_LAUNCH(k_axpby, blocks, kBlock, 0, st, n, 1.0, wk.r.p, rs_new / rs_old, wk.p.p);               // p = r + beta p
    }
    if (it >= 500 && !(rs <= 1e-20 * b2)) return 1;   // stagnated above 1e-10: give the step up
  }
  if (iters_out) *iters_out = it;
  return 0;
}

}  // namespace econpdes
