// (Block-)tridiagonal solves by parallel cyclic reduction (PCR) for 1-D grids: replaces the
// sparse LU (`J \ R`, UMFPACK) NLsolve runs every Newton iteration
// (src/finiteschemesolve.jl:135-137) when N = 1.  The Jacobian of a 1-D problem with F
// unknowns is block tridiagonal with F x F blocks, read in place from the star-stencil
// planes the Jacobian kernel writes:
//      A_i = planes[k=1] (coupling to i-1), B_i = planes[k=0], C_i = planes[k=2],
// block element (fout, fin) at plane (fin*F + fout).
// PCR step with stride s:  alpha = -A_i B_{i-s}^{-1}, gamma = -C_i B_{i+s}^{-1}
//      A_i' = alpha A_{i-s},  C_i' = gamma C_{i+s},
//      B_i' = B_i + alpha C_{i-s} + gamma A_{i+s},  d_i' = d_i + alpha d_{i-s} + gamma d_{i+s}
// after ceil(log2 n) steps x_i = B_i^{-1} d_i.  No pivoting across rows: the pseudo-transient
// Jacobians are (block) diagonally dominant M-matrices for monotone schemes; a vanishing
// pivot surfaces as a non-finite solution, which the Newton driver maps to the reference's
// SingularException outcome (src/finiteschemesolve.jl:142-146).
#pragma once
#include "common.cuh"

namespace econpdes {

template <int F>
struct SmallMat {
  double a[F * F];  // column-major: a[c*F + r]
  __host__ __device__ __forceinline__ double& operator()(int r, int c) { return a[c * F + r]; }
  __host__ __device__ __forceinline__ double operator()(int r, int c) const { return a[c * F + r]; }
};

template <int F>
__host__ __device__ __forceinline__ SmallMat<F> mat_mul(const SmallMat<F>& x, const SmallMat<F>& y) {
  SmallMat<F> z;
#pragma unroll
  for (int r = 0; r < F; ++r)
#pragma unroll
    for (int c = 0; c < F; ++c) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < F; ++k) t += x(r, k) * y(k, c);
      z(r, c) = t;
    }
  return z;
}

// Gauss-Jordan inverse with partial pivoting inside the F x F block.
template <int F>
__host__ __device__ __forceinline__ SmallMat<F> mat_inv(SmallMat<F> m) {
  SmallMat<F> inv;
#pragma unroll
  for (int r = 0; r < F; ++r)
#pragma unroll
    for (int c = 0; c < F; ++c) inv(r, c) = r == c ? 1.0 : 0.0;
#pragma unroll
  for (int col = 0; col < F; ++col) {
    int piv = col;
    double best = fabs(m(col, col));
#pragma unroll
    for (int r = col + 1; r < F; ++r) {
      const double v = fabs(m(r, col));
      if (v > best) { best = v; piv = r; }
    }
    if (piv != col) {
#pragma unroll
      for (int c = 0; c < F; ++c) {
        double t = m(col, c); m(col, c) = m(piv, c); m(piv, c) = t;
        t = inv(col, c); inv(col, c) = inv(piv, c); inv(piv, c) = t;
      }
    }
    const double d = 1.0 / m(col, col);
#pragma unroll
    for (int c = 0; c < F; ++c) { m(col, c) *= d; inv(col, c) *= d; }
#pragma unroll
    for (int r = 0; r < F; ++r) {
      if (r == col) continue;
      const double f = m(r, col);
#pragma unroll
      for (int c = 0; c < F; ++c) { m(r, c) -= f * m(col, c); inv(r, c) -= f * inv(col, c); }
    }
  }
  return inv;
}

struct PcrSystem {
  const double* A;  // [F*F][n]
  const double* B;
  const double* C;
  const double* d;  // [F][n]
};
struct PcrSystemOut {
  double* A;
  double* B;
  double* C;
  double* d;
};

template <int F>
__device__ __forceinline__ SmallMat<F> load_mat(const double* __restrict__ p, long long n, long long i) {
  SmallMat<F> m;
#pragma unroll
  for (int e = 0; e < F * F; ++e) m.a[e] = p[e * n + i];
  return m;
}
template <int F>
__device__ __forceinline__ void store_mat(double* __restrict__ p, long long n, long long i, const SmallMat<F>& m) {
#pragma unroll
  for (int e = 0; e < F * F; ++e) p[e * n + i] = m.a[e];
}

template <int F>
__global__ void __launch_bounds__(kBlock) k_pcr_step(long long n, long long s, PcrSystem in, PcrSystemOut out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SmallMat<F> B = load_mat<F>(in.B, n, i);
  double d[F];
#pragma unroll
  for (int f = 0; f < F; ++f) d[f] = in.d[f * n + i];
  SmallMat<F> An, Cn;
#pragma unroll
  for (int e = 0; e < F * F; ++e) An.a[e] = Cn.a[e] = 0.0;
  const long long im = i - s, ip = i + s;
  if (im >= 0) {
    const SmallMat<F> A = load_mat<F>(in.A, n, i);
    const SmallMat<F> alpha = mat_mul<F>(A, mat_inv<F>(load_mat<F>(in.B, n, im)));
    const SmallMat<F> t1 = mat_mul<F>(alpha, load_mat<F>(in.A, n, im));
    const SmallMat<F> t2 = mat_mul<F>(alpha, load_mat<F>(in.C, n, im));
#pragma unroll
    for (int e = 0; e < F * F; ++e) { An.a[e] = -t1.a[e]; B.a[e] -= t2.a[e]; }
#pragma unroll
    for (int r = 0; r < F; ++r) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < F; ++k) t += alpha(r, k) * in.d[k * n + im];
      d[r] -= t;
    }
  }
  if (ip < n) {
    const SmallMat<F> C = load_mat<F>(in.C, n, i);
    const SmallMat<F> gamma = mat_mul<F>(C, mat_inv<F>(load_mat<F>(in.B, n, ip)));
    const SmallMat<F> t1 = mat_mul<F>(gamma, load_mat<F>(in.C, n, ip));
    const SmallMat<F> t2 = mat_mul<F>(gamma, load_mat<F>(in.A, n, ip));
#pragma unroll
    for (int e = 0; e < F * F; ++e) { Cn.a[e] = -t1.a[e]; B.a[e] -= t2.a[e]; }
#pragma unroll
    for (int r = 0; r < F; ++r) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < F; ++k) t += gamma(r, k) * in.d[k * n + ip];
      d[r] -= t;
    }
  }
  store_mat<F>(out.A, n, i, An);
  store_mat<F>(out.B, n, i, B);
  store_mat<F>(out.C, n, i, Cn);
#pragma unroll
  for (int f = 0; f < F; ++f) out.d[f * n + i] = d[f];
}

// x_i = B_i^{-1} d_i; out[1] = max|x| over finite entries, out[2] = 1 if any non-finite
template <int F>
__global__ void __launch_bounds__(kBlock) k_pcr_final(long long n, PcrSystem in, double* __restrict__ x, NormWork nw) {
  __shared__ double scratch[96];
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  double sum = 0.0, m = 0.0;
  int bad = 0;
  if (i < n) {
    const SmallMat<F> Binv = mat_inv<F>(load_mat<F>(in.B, n, i));
#pragma unroll
    for (int r = 0; r < F; ++r) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < F; ++k) t += Binv(r, k) * in.d[k * n + i];
      x[r * n + i] = t;
      if (!isfinite(t)) bad = 1;
      else m = fabs(t) > m ? fabs(t) : m;
    }
  }
  grid_norm_finish(sum, m, bad, nw, scratch);
}

// ---- scalar PCR of one system held by one thread block (one row per thread) ----
// Used by the fused batched kernel; a, b, c, d are this thread's row, sh is shared scratch
// of 8 * blockDim.x doubles (double-buffered a, b, c, d).  Rows >= n behave as identity
// rows.  Returns x_i.  All threads of the block must call it.
// 1/b to ~1 ulp without the special-case handling of an IEEE division: hardware seed (>= 20
// bits) + two Newton steps.  b = 0 or non-finite gives a non-finite result, which the caller
// reports as a singular Jacobian either way.  (The tridiagonal solve replaces the reference's
// sparse LU; its rounding is not on the parity path.)
__device__ __forceinline__ double fast_rcp(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  return fma(r, e, r);
}

__device__ __forceinline__ double block_pcr_solve(double a, double b, double c, double d, int i, int n, double* sh) {
  const int T = blockDim.x;
  int buf = 0;
  if (i >= n) { a = 0.0; b = 1.0; c = 0.0; d = 0.0; }
  for (int s = 1; s < n; s <<= 1) {
    double* wa = sh + buf * 4 * T;   // a, 1/b, c, d of every row for this step
    double* wi = wa + T;
    double* wc = wi + T;
    double* wd = wc + T;
    wa[i] = a; wi[i] = fast_rcp(b); wc[i] = c; wd[i] = d;   // each pivot is inverted once, by its owner
    __syncthreads();
    const int im = i - s, ip = i + s;
    double na = 0.0, nc = 0.0;
    if (im >= 0 && i < n) {
      const double alpha = -a * wi[im];
      na = alpha * wa[im];
      b += alpha * wc[im];
      d += alpha * wd[im];
    }
    if (ip < n && i < n) {
      const double gamma = -c * wi[ip];
      nc = gamma * wc[ip];
      b += gamma * wa[ip];
      d += gamma * wd[ip];
    }
    a = na;
    c = nc;
    buf ^= 1;  // the next step writes the other buffer, so one barrier per step suffices
  }
  return d * fast_rcp(b);
}

// ---- P consecutive rows per thread: Schur complement onto one interface row per thread ----
// Thread t owns rows P t .. P t + P - 1.  Its first P - 1 rows (the chunk interior, a small
// tridiagonal block T_t) are eliminated in registers by the Thomas algorithm with three right-hand
// sides, which expresses them through the interface unknowns u_{t-1}, u_t (u_t = x_{P t + P - 1}):
//      x_int = p - q u_{t-1} - r u_t,   p = T_t^{-1} d_int,  q = T_t^{-1} (a_0 e_0),  r = T_t^{-1} (c_{P-2} e_{P-2})
// Substituting the last interior row of chunk t and the first of chunk t+1 into interface row t
// leaves a tridiagonal system in u with one row per thread, solved by block_pcr_solve; the
// interiors follow by back-substitution.  For P = 2 this is one odd-even reduction step.  Work per
// row falls from ~log2(n) PCR updates to ~log2(n/P)/P + a few Thomas updates.
// a, b, c, d: the thread's rows (rows beyond the system must be identity rows: a = c = d = 0, b = 1);
// nred = number of interface rows that matter (ceil(n / P)); sh: 8 * blockDim.x doubles;
// xch: 4 * blockDim.x doubles.  All threads of the block must call it.
template <int P>
__device__ __forceinline__ void block_chunk_solve(const double* a, const double* b, const double* c, const double* d,
                                                  int t, int nred, double* sh, double* xch, double* x) {
  if constexpr (P == 1) {
    x[0] = block_pcr_solve(a[0], b[0], c[0], d[0], t, nred, sh);
  } else {
    constexpr int m = P - 1;
    const int T = blockDim.x;
    double cp[m], p[m], q[m], r[m];
    {
      double w = fast_rcp(b[0]);
      cp[0] = c[0] * w; p[0] = d[0] * w; q[0] = a[0] * w;
#pragma unroll
      for (int j = 1; j < m; ++j) {
        w = fast_rcp(b[j] - a[j] * cp[j - 1]);
        cp[j] = c[j] * w;
        p[j] = (d[j] - a[j] * p[j - 1]) * w;
        q[j] = -(a[j] * q[j - 1]) * w;
      }
      r[m - 1] = cp[m - 1];
#pragma unroll
      for (int j = m - 2; j >= 0; --j) {
        p[j] -= cp[j] * p[j + 1];
        q[j] -= cp[j] * q[j + 1];
        r[j] = -cp[j] * r[j + 1];
      }
    }
    // first interior row of every chunk, for the interface row of the chunk before it
    xch[t] = p[0]; xch[T + t] = q[0]; xch[2 * T + t] = r[0];
    __syncthreads();
    const bool has_next = t + 1 < T;
    const double pn = has_next ? xch[t + 1] : 0.0, qn = has_next ? xch[T + t + 1] : 0.0,
                 rn = has_next ? xch[2 * T + t + 1] : 0.0;
    const double aI = a[m], cI = c[m];
    const double A = -aI * q[m - 1];
    const double B = b[m] - aI * r[m - 1] - cI * qn;
    const double C = -cI * rn;
    const double D = d[m] - aI * p[m - 1] - cI * pn;
    const double u = block_pcr_solve(A, B, C, D, t, nred, sh);
    xch[3 * T + t] = u;
    __syncthreads();
    const double um = t > 0 ? xch[3 * T + t - 1] : 0.0;
#pragma unroll
    for (int j = 0; j < m; ++j) x[j] = p[j] - q[j] * um - r[j] * u;
    x[m] = u;
  }
}

}  // namespace econpdes
