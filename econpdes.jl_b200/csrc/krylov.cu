// Linear solver for 2-D / 3-D grids: flexible GMRES preconditioned by one geometric
// multigrid V(1,1)-cycle whose smoother is the alternating-direction zebra line solve
// (line-tridiagonal, ADI-style), all in stencil-coefficient (matrix-free) form.
//
// Replaces the sparse LU (`J \ R`, UMFPACK through NLsolve's default linsolve) that the
// reference runs every Newton iteration (src/finiteschemesolve.jl:135-137):
// a direct factorisation of the 2048^2 or 256^3 Jacobian is infeasible, a one-level line
// preconditioner needs O(10^4) iterations (SURVEY.md hard part 2).  The Krylov solve is
// driven to ~1e-12 relative residual so that Newton iterates are indistinguishable from the
// LU path's (parity is path parity).
//
// Operator storage: planes[k][pt], k over the +-1 box (3^N) or, on the finest level of a
// model without cross derivatives, its star subset (1+2N).  Coarse operators are Galerkin
// (P^T A P) with a per-level, per-dimension transfer: linear interpolation where the
// coupling along that dimension is diffusion-like (symmetric), piecewise-constant
// aggregation where it is upwind/advective (one-sided) -- linear interpolation is unstable
// for pure upwind advection (the asset dimension of the consumption-saving model), while
// aggregation alone degrades with the number of levels.
//
// Lines along dimension d > 0 are strided in memory; each level keeps a transposed copy of
// its operator per line direction so that every line solve streams contiguous memory, and
// vectors are transposed in and out around those sweeps.
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include <functional>

#include "engine.cuh"
#include "tridiag.cuh"

namespace econpdes {
namespace {

constexpr int kMaxLevels = 16;
constexpr int kCoarseMax2D = 8;    // stop coarsening when every dimension has <= this many points
constexpr int kCoarseMax3D = 4;    // (the coarsest operator is factorised densely by one thread block)
constexpr double kLinearThreshold = 0.3;

// Storage type of the multigrid hierarchy's operators.  The V-cycle is a preconditioner inside
// flexible GMRES, whose own operator (the Jacobian planes) and vectors stay FP64, so the hierarchy
// can hold its stencil coefficients in FP32: the line smoother streams 27 planes per point and
// runs at ~60 % of HBM bandwidth on a 256^3 grid, so halving those bytes is what shortens the
// cycle.  All arithmetic (sweeps, Galerkin products, residuals) is still FP64.
// Only the 3-D hierarchies do so: with the 5-plane star stencils of the 2-D models the sweeps are
// not bandwidth-bound and the conversions cost more than the bytes save (measured).
#ifdef ECONPDES_MG_DOUBLE
template <int N> using mgreal_t = double;
#else
template <int N> using mgreal_t = std::conditional_t<(N >= 3), float, double>;
#endif

template <int N>
struct LevelGeom {
  long long S[N];
  long long stride[N];
  long long l;
  // slab of a distributed level: the first / last plane of dimension N-1 has a neighbour plane
  // in the ghost storage around the vector (filled by a halo exchange), not a grid boundary
  bool open_lo = false, open_hi = false;
  // planes that can hold a non-zero coefficient (bit k = plane k): the finest level of a model that declares
  // which box slots its residual reads (M::slot_used) skips the others -- they are exact zeros, in the
  // reference's Jacobian too -- in the operator applications and the smoother; Galerkin levels use every plane
  unsigned mask = 0xffffffffu;
  FastDiv div[N];   // division by S[d]: a level holds fewer than 2^31 points (checked when the problem is created)
};

template <int N>
LevelGeom<N> make_geom(const long long* S) {
  LevelGeom<N> g;
  long long st = 1;
  for (int d = 0; d < N; ++d) {
    g.S[d] = S[d];
    g.stride[d] = st;
    st *= S[d];
    g.div[d].set(static_cast<unsigned>(S[d]));
  }
  g.l = st;
  return g;
}

template <int N>
__device__ __forceinline__ void unravel_l(long long pt, const LevelGeom<N>& g, long long* idx) {
  unsigned p = static_cast<unsigned>(pt);   // multiply-high instead of 64-bit divisions by run-time sizes
#pragma unroll
  for (int d = 0; d + 1 < N; ++d) {
    const unsigned q = g.div[d].div(p);
    idx[d] = p - q * static_cast<unsigned>(g.S[d]);
    p = q;
  }
  idx[N - 1] = p;
}

// y = A x (mode 0) or y = b - A x (mode 1); entries pointing outside the grid hold 0.
// MASKED: planes outside g.mask are not read
template <int N, bool BOX, class TA, bool MASKED = false>
__global__ void __launch_bounds__(kBlock) k_spmv(LevelGeom<N> g, const TA* __restrict__ A,
                                                 const double* __restrict__ x, const double* __restrict__ b,
                                                 double* __restrict__ y, int mode) {
  using St = Stencil<N, BOX>;
  const long long pt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pt >= g.l) return;
  long long idx[N];
  unravel_l<N>(pt, g, idx);
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < St::NK; ++k) {
    // block-uniform skip (measured: the branch beats predicated loads here, 0.37 against 0.53 ms at 256^3)
    if (MASKED && !((g.mask >> k) & 1u)) continue;
    long long q = pt;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      const int o = St::off(k, d);
      const bool lo_edge = idx[d] == 0 && !(d == N - 1 && g.open_lo);
      const bool hi_edge = idx[d] == g.S[d] - 1 && !(d == N - 1 && g.open_hi);
      if ((o < 0 && lo_edge) || (o > 0 && hi_edge)) ok = false;
      q += o * g.stride[d];
    }
    // entries that point outside the grid hold 0: load unconditionally (clamped address) so all
    // loads of a point are in flight together instead of one round trip per guarded entry
    acc += A[k * g.l + pt] * x[ok ? q : pt];
  }
  y[pt] = mode ? b[pt] - acc : acc;
}

// ---- transposes: layout d puts dimension d fastest, the others keep their order --------
// All of them are (batched) 2-D transposes of an R x C matrix (R fastest) into C x R.
template <class T>
__global__ void k_transpose(const T* __restrict__ src, T* __restrict__ dst, long long R, long long C,
                            long long batch_stride) {
  __shared__ T tile[32][33];
  const long long base = static_cast<long long>(blockIdx.z) * batch_stride;
  long long r = static_cast<long long>(blockIdx.x) * 32 + threadIdx.x;
  long long c = static_cast<long long>(blockIdx.y) * 32 + threadIdx.y;
#pragma unroll
  for (int j = 0; j < 32; j += 8)
    if (r < R && c + j < C) tile[threadIdx.y + j][threadIdx.x] = src[base + (c + j) * R + r];
  __syncthreads();
  c = static_cast<long long>(blockIdx.y) * 32 + threadIdx.x;
  r = static_cast<long long>(blockIdx.x) * 32 + threadIdx.y;
#pragma unroll
  for (int j = 0; j < 32; j += 8)
    if (c < C && r + j < R) dst[base + (r + j) * C + c] = tile[threadIdx.x][threadIdx.y + j];
}

// two transposes of the same shape in one launch (the iterate and the right-hand side of a line sweep): grid.z holds
// the batches of the first pair, then those of the second -- the levels below 128^3 / 512^2 pay per launch, not per byte
template <class T>
__global__ void k_transpose2(const T* __restrict__ src0, T* __restrict__ dst0, const T* __restrict__ src1,
                             T* __restrict__ dst1, long long R, long long C, long long batch_stride, unsigned batches) {
  __shared__ T tile[32][33];
  const bool second = blockIdx.z >= batches;
  const T* __restrict__ src = second ? src1 : src0;
  T* __restrict__ dst = second ? dst1 : dst0;
  const long long base = static_cast<long long>(second ? blockIdx.z - batches : blockIdx.z) * batch_stride;
  long long r = static_cast<long long>(blockIdx.x) * 32 + threadIdx.x;
  long long c = static_cast<long long>(blockIdx.y) * 32 + threadIdx.y;
#pragma unroll
  for (int j = 0; j < 32; j += 8)
    if (r < R && c + j < C) tile[threadIdx.y + j][threadIdx.x] = src[base + (c + j) * R + r];
  __syncthreads();
  c = static_cast<long long>(blockIdx.y) * 32 + threadIdx.x;
  r = static_cast<long long>(blockIdx.x) * 32 + threadIdx.y;
#pragma unroll
  for (int j = 0; j < 32; j += 8)
    if (c < C && r + j < R) dst[base + (r + j) * C + c] = tile[threadIdx.x][threadIdx.y + j];
}

struct TransposeShape {
  long long R, C, batches, batch_stride;
};
// forward: layout 0 -> layout d.  N == 2: d=1 transposes S0 x S1.  N == 3: d=1 transposes S0 x S1
// for each i2; d=2 transposes (S0*S1) x S2.
template <int N>
TransposeShape transpose_shape(const long long* S, int d) {
  if (N == 2) return {S[0], S[1], 1, 0};
  if (d == 1) return {S[0], S[1], S[2], S[0] * S[1]};
  return {S[0] * S[1], S[2], 1, 0};
}
template <class T>
void launch_transpose(cudaStream_t s, const T* src, T* dst, TransposeShape t, bool inverse) {
  const long long R = inverse ? t.C : t.R, C = inverse ? t.R : t.C;
  dim3 grid(static_cast<unsigned>((R + 31) / 32), static_cast<unsigned>((C + 31) / 32),
            static_cast<unsigned>(t.batches));
  EP_LAUNCH(k_transpose<T>, grid, dim3(32, 8), 0, s, src, dst, R, C, t.batch_stride);
}
template <class T>
void launch_transpose2(cudaStream_t s, const T* src0, T* dst0, const T* src1, T* dst1, TransposeShape t) {
  dim3 grid(static_cast<unsigned>((t.R + 31) / 32), static_cast<unsigned>((t.C + 31) / 32),
            static_cast<unsigned>(2 * t.batches));
  EP_LAUNCH(k_transpose2<T>, grid, dim3(32, 8), 0, s, src0, dst0, src1, dst1, t.R, t.C, t.batch_stride,
            static_cast<unsigned>(t.batches));
}
// sizes of layout d, and where each original dimension lands
template <int N>
void layout_dims(const long long* S, int d, long long* Sd, int* perm /* perm[e] = original dim at slot e */) {
  perm[0] = d;
  int e = 1;
  for (int q = 0; q < N; ++q)
    if (q != d) perm[e++] = q;
  for (int q = 0; q < N; ++q) Sd[q] = S[perm[q]];
}

// ---- zebra line smoother: one thread block per line (lines are contiguous) -------------
// Solves, for every line of the given color, the tridiagonal system made of the in-line
// couplings with right-hand side b - (off-line couplings) x, by PCR in shared memory.
template <int N, bool BOX, int RPT, bool MASKED = false>
__global__ void __launch_bounds__(1024) k_line_zebra(LevelGeom<N> g, const mgreal_t<N>* __restrict__ A,
                                                     const double* __restrict__ b, double* __restrict__ x,
                                                     int color) {
  using St = Stencil<N, BOX>;
  extern __shared__ double sh[];
  const int n = static_cast<int>(g.S[0]);
  double* sa = sh;
  double* si = sh + n;
  double* sc = sh + 2 * n;
  double* sd = sh + 3 * n;
  // which line: other coordinates j_e = c_e + 2 q_e
  long long other[N > 1 ? N - 1 : 1];
  {
    long long q = blockIdx.x;
#pragma unroll
    for (int e = 1; e < N; ++e) {
      const int ce = (color >> (e - 1)) & 1;
      const long long cnt = (g.S[e] - ce + 1) / 2;
      other[e - 1] = ce + 2 * (q % cnt);
      q /= cnt;
    }
  }
  long long base = 0;
#pragma unroll
  for (int e = 1; e < N; ++e) base += other[e - 1] * g.stride[e];
  // centre / in-line plane ids
  int kc = -1, km = -1, kp = -1;
#pragma unroll
  for (int k = 0; k < St::NK; ++k) {
    bool inl = true;
#pragma unroll
    for (int e = 1; e < N; ++e)
      if (St::off(k, e) != 0) inl = false;
    if (inl) {
      if (St::off(k, 0) == 0) kc = k;
      if (St::off(k, 0) < 0) km = k;
      if (St::off(k, 0) > 0) kp = k;
    }
  }
  double ra[RPT], rb[RPT], rc[RPT], rd[RPT];
#pragma unroll
  for (int m = 0; m < RPT; ++m) {
    const int i = threadIdx.x + m * blockDim.x;
    ra[m] = 0.0; rb[m] = 1.0; rc[m] = 0.0; rd[m] = 0.0;
    if (i < n) {
      const long long pt = base + i;
      double rhs = b[pt];
#pragma unroll
      for (int k = 0; k < St::NK; ++k) {
        const bool used = !MASKED || ((g.mask >> k) & 1u);   // block-uniform
        bool inl = true, ok = true;
        long long q = pt;
#pragma unroll
        for (int e = 0; e < N; ++e) {
          const int o = St::off(k, e);
          if (e > 0 && o != 0) inl = false;
          const long long ie = e == 0 ? i : other[e - 1];
          const bool lo_edge = ie == 0 && !(e == N - 1 && g.open_lo);
          const bool hi_edge = ie == g.S[e] - 1 && !(e == N - 1 && g.open_hi);
          if ((o < 0 && lo_edge) || (o > 0 && hi_edge)) ok = false;
          q += o * g.stride[e];
        }
        if (!inl) {   // zero coefficient outside the grid
          const double a = used ? static_cast<double>(A[k * g.l + pt]) : 0.0;
          const double xv = used ? x[ok ? q : pt] : 0.0;
          rhs -= a * xv;
        }
      }
      ra[m] = i > 0 ? A[km * g.l + pt] : 0.0;
      rb[m] = A[kc * g.l + pt];
      rc[m] = i < n - 1 ? A[kp * g.l + pt] : 0.0;
      rd[m] = rhs;
    }
  }
  for (int s = 1; s < n; s <<= 1) {
#pragma unroll
    for (int m = 0; m < RPT; ++m) {
      const int i = threadIdx.x + m * blockDim.x;
      if (i < n) { sa[i] = ra[m]; si[i] = fast_rcp(rb[m]); sc[i] = rc[m]; sd[i] = rd[m]; }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < RPT; ++m) {
      const int i = threadIdx.x + m * blockDim.x;
      if (i < n) {
        const int im = i - s, ip = i + s;
        double na = 0.0, nc = 0.0;
        if (im >= 0) {
          const double al = -ra[m] * si[im];
          na = al * sa[im];
          rb[m] += al * sc[im];
          rd[m] += al * sd[im];
        }
        if (ip < n) {
          const double ga = -rc[m] * si[ip];
          nc = ga * sc[ip];
          rb[m] += ga * sa[ip];
          rd[m] += ga * sd[ip];
        }
        ra[m] = na;
        rc[m] = nc;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < RPT; ++m) {
    const int i = threadIdx.x + m * blockDim.x;
    if (i < n) x[base + i] = rd[m] / rb[m];
  }
}

// ---- zebra line smoother, lines of up to 256 points: one WARP per line -------------------
// The block-wide PCR above spends ~80 shared-memory operations and log2(n) barriers per point.
// Here a warp owns a line:
//   1. coalesced assembly (lane l handles points l, l+32, ..): right-hand side b - (off-line
//      couplings) x and the three in-line coefficients, parked in shared memory;
//   2. lane l takes the PPT consecutive rows l*PPT .. l*PPT+PPT-1 into registers and eliminates
//      its interior rows (Thomas with three right-hand sides), which expresses them through the
//      last unknown of the previous lane and its own last unknown;
//   3. the 32 "last unknowns" form a tridiagonal system across the lanes, solved by PCR with
//      warp shuffles (5 steps, no barrier);
//   4. interior unknowns follow by substitution; the line is written back coalesced.
// ~10 shared-memory operations per point instead of ~80 and no __syncthreads.

template <int PPT>
__device__ __forceinline__ int line_pos(int i) { return PPT > 1 ? i + i / PPT : i; }   // padded: conflict-free strided reads

// ---- zebra line smoother, long lines of star-stencil (2-D) levels: P consecutive rows per thread ----
// One block per line like k_line_zebra, but the block-wide PCR (log2(n) steps over n rows, ~300
// instructions per point at n = 2048, which is what bounds the 5-plane 2-D sweeps) is replaced by the
// chunked Schur complement of the batched engine (tridiag.cuh, block_chunk_solve): each thread assembles
// its P consecutive rows straight from global memory (P consecutive doubles per plane: whole sectors),
// eliminates the chunk interior in registers, PCR runs on one interface row per thread, and the thread
// writes its P unknowns back.  Shared memory holds only the PCR scratch.
template <int N, bool BOX, int P, bool MASKED = false>
__global__ void __launch_bounds__(1024) k_line_zebra_chunk(LevelGeom<N> g, const mgreal_t<N>* __restrict__ A,
                                                           const double* __restrict__ b, double* __restrict__ x,
                                                           int color) {
  using St = Stencil<N, BOX>;
  extern __shared__ double sh[];
  const int n = static_cast<int>(g.S[0]);
  const int T = blockDim.x, t = threadIdx.x;
  double* pcr = sh;                      // [8 T]
  double* xch = pcr + 8 * T;             // [4 T]
  long long other[N > 1 ? N - 1 : 1];
  {
    long long q = blockIdx.x;
#pragma unroll
    for (int e = 1; e < N; ++e) {
      const int ce = (color >> (e - 1)) & 1;
      const long long cnt = (g.S[e] - ce + 1) / 2;
      other[e - 1] = ce + 2 * (q % cnt);
      q /= cnt;
    }
  }
  long long base = 0;
#pragma unroll
  for (int e = 1; e < N; ++e) base += other[e - 1] * g.stride[e];
  int kc = -1, km = -1, kp = -1;
#pragma unroll
  for (int k = 0; k < St::NK; ++k) {
    bool inl = true;
#pragma unroll
    for (int e = 1; e < N; ++e)
      if (St::off(k, e) != 0) inl = false;
    if (inl) {
      if (St::off(k, 0) == 0) kc = k;
      if (St::off(k, 0) < 0) km = k;
      if (St::off(k, 0) > 0) kp = k;
    }
  }
  double a[P], bb[P], c[P], d[P], xs[P];
  // Right-hand sides b - (off-line couplings) x of this thread's P consecutive rows.  The plane loop is the outer
  // one: whether a plane is used (block-uniform branch: a predicated load still costs its address arithmetic -- the
  // 256^3 level of the synthetic model reads 6 of its 24 off-line planes) and whether its neighbour line exists
  // depend on the line only, so they are decided once per plane instead of once per plane and row.
  // Skipped planes used to subtract 0 * 0: same bits.
  const int i0 = P * t;
  const long long pt0 = base + i0;
  bool inr[P];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    inr[j] = i0 + j < n;
    d[j] = inr[j] ? b[pt0 + j] : 0.0;
  }
  // the in-line coefficients first: their loads are in flight while the off-line planes are gathered
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int i = i0 + j;
    a[j] = 0.0; bb[j] = 1.0; c[j] = 0.0;
    if (inr[j]) {
      const long long pt = pt0 + j;
      a[j] = i > 0 ? A[km * g.l + pt] : 0.0;
      bb[j] = A[kc * g.l + pt];
      c[j] = i < n - 1 ? A[kp * g.l + pt] : 0.0;
    }
  }
#pragma unroll
  for (int k = 0; k < St::NK; ++k) {
    bool inl = true;   // folds at compile time (unrolled k)
#pragma unroll
    for (int e = 1; e < N; ++e)
      if (St::off(k, e) != 0) inl = false;
    if (inl) continue;
    if (MASKED && !((g.mask >> k) & 1u)) continue;   // block-uniform
    bool okk = true;
    long long dq = St::off(k, 0);
#pragma unroll
    for (int e = 1; e < N; ++e) {
      const int o = St::off(k, e);
      const long long ie = other[e - 1];
      const bool lo_edge = ie == 0 && !(e == N - 1 && g.open_lo);
      const bool hi_edge = ie == g.S[e] - 1 && !(e == N - 1 && g.open_hi);
      if ((o < 0 && lo_edge) || (o > 0 && hi_edge)) okk = false;
      dq += o * g.stride[e];
    }
    const int o0 = St::off(k, 0);
    const mgreal_t<N>* __restrict__ Ak = A + k * g.l + pt0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int i = i0 + j;
      const bool ok = okk && !((o0 < 0 && i == 0) || (o0 > 0 && i == n - 1));   // zero coefficient outside the grid
      if (inr[j]) d[j] -= static_cast<double>(Ak[j]) * x[ok ? pt0 + j + dq : pt0 + j];
    }
  }
  block_chunk_solve<P>(a, bb, c, d, t, (n + P - 1) / P, pcr, xch, xs);
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int i = P * t + j;
    if (i < n) x[base + i] = xs[j];
  }
}

template <int N, bool BOX, int PPT, bool MASKED = false>
__global__ void __launch_bounds__(256) k_line_zebra_warp(LevelGeom<N> g, const mgreal_t<N>* __restrict__ A,
                                                         const double* __restrict__ b, double* __restrict__ x,
                                                         int color, long long nlines) {
  using St = Stencil<N, BOX>;
  constexpr int NP = 32 * PPT;            // line capacity
  constexpr int PITCH = NP + 32;          // padded array length
  extern __shared__ double sh[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long line = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (line >= nlines) return;             // warp-uniform; no block-level barrier below
  double* sa = sh + warp * 4 * PITCH;
  double* sb = sa + PITCH;
  double* sc = sb + PITCH;
  double* sd = sc + PITCH;
  const int n = static_cast<int>(g.S[0]);
  long long other[N > 1 ? N - 1 : 1];
  {
    long long q = line;
#pragma unroll
    for (int e = 1; e < N; ++e) {
      const int ce = (color >> (e - 1)) & 1;
      const long long cnt = (g.S[e] - ce + 1) / 2;
      other[e - 1] = ce + 2 * (q % cnt);
      q /= cnt;
    }
  }
  long long base = 0;
#pragma unroll
  for (int e = 1; e < N; ++e) base += other[e - 1] * g.stride[e];
  int kc = -1, km = -1, kp = -1;
#pragma unroll
  for (int k = 0; k < St::NK; ++k) {
    bool inl = true;
#pragma unroll
    for (int e = 1; e < N; ++e)
      if (St::off(k, e) != 0) inl = false;
    if (inl) {
      if (St::off(k, 0) == 0) kc = k;
      if (St::off(k, 0) < 0) km = k;
      if (St::off(k, 0) > 0) kp = k;
    }
  }
  // ---- 1. coalesced assembly ----
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i = lane + 32 * j;
    double va = 0.0, vb = 1.0, vc = 0.0, vd = 0.0;   // rows beyond the line are identity rows
    if (i < n) {
      const long long pt = base + i;
      double rhs = b[pt];
#pragma unroll
      for (int k = 0; k < St::NK; ++k) {
        const bool used = !MASKED || ((g.mask >> k) & 1u);   // block-uniform
        bool inl = true, ok = true;
        long long q = pt;
#pragma unroll
        for (int e = 0; e < N; ++e) {
          const int o = St::off(k, e);
          if (e > 0 && o != 0) inl = false;
          const long long ie = e == 0 ? i : other[e - 1];
          const bool lo_edge = ie == 0 && !(e == N - 1 && g.open_lo);
          const bool hi_edge = ie == g.S[e] - 1 && !(e == N - 1 && g.open_hi);
          if ((o < 0 && lo_edge) || (o > 0 && hi_edge)) ok = false;
          q += o * g.stride[e];
        }
        if (!inl) {   // zero coefficient outside the grid
          const double a = used ? static_cast<double>(A[k * g.l + pt]) : 0.0;
          const double xv = used ? x[ok ? q : pt] : 0.0;
          rhs -= a * xv;
        }
      }
      va = i > 0 ? A[km * g.l + pt] : 0.0;
      vb = A[kc * g.l + pt];
      vc = i < n - 1 ? A[kp * g.l + pt] : 0.0;
      vd = rhs;
    }
    const int p = line_pos<PPT>(i);
    sa[p] = va; sb[p] = vb; sc[p] = vc; sd[p] = vd;
  }
  __syncwarp();
  // ---- 2. interior rows of this lane: x_j = u_j - v_j x_prev - w_j x_last ----
  constexpr int M1 = PPT > 1 ? PPT - 1 : 1;
  double cp[M1], u[M1], v[M1];
  const int r0 = lane * PPT;
  if (PPT > 1) {
    {
      const int p = line_pos<PPT>(r0);
      const double inv = fast_rcp(sb[p]);
      cp[0] = sc[p] * inv;
      u[0] = sd[p] * inv;
      v[0] = sa[p] * inv;
    }
#pragma unroll
    for (int j = 1; j < PPT - 1; ++j) {
      const int p = line_pos<PPT>(r0 + j);
      const double aj = sa[p];
      const double inv = fast_rcp(sb[p] - aj * cp[j - 1]);
      cp[j] = sc[p] * inv;
      u[j] = (sd[p] - aj * u[j - 1]) * inv;
      v[j] = (-aj * v[j - 1]) * inv;
    }
  }
  double w[M1];
  if (PPT > 1) {
    w[PPT - 2] = cp[PPT - 2];
#pragma unroll
    for (int j = PPT - 3; j >= 0; --j) {
      u[j] -= cp[j] * u[j + 1];
      v[j] -= cp[j] * v[j + 1];
      w[j] = -cp[j] * w[j + 1];
    }
  }
  // ---- 3. reduced system in the lanes' last unknowns ----
  double ra, rb, rc, rd;
  {
    const int p = line_pos<PPT>(r0 + PPT - 1);
    const double al = sa[p], bl = sb[p], cl = sc[p], dl = sd[p];
    if (PPT > 1) {
      // first interior unknown of the next lane: u'_0 - v'_0 x_last(me) - w'_0 x_last(next)
      const double un = __shfl_down_sync(0xffffffffu, u[0], 1);
      const double vn = __shfl_down_sync(0xffffffffu, v[0], 1);
      const double wn = __shfl_down_sync(0xffffffffu, w[0], 1);
      ra = -al * v[PPT - 2];
      rb = bl - al * w[PPT - 2] - cl * vn;
      rc = -cl * wn;
      rd = dl - al * u[PPT - 2] - cl * un;
    } else {
      ra = al; rb = bl; rc = cl; rd = dl;
    }
  }
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const double inv = fast_rcp(rb);
    const double am = __shfl_up_sync(0xffffffffu, ra, s), im = __shfl_up_sync(0xffffffffu, inv, s);
    const double cm = __shfl_up_sync(0xffffffffu, rc, s), dm = __shfl_up_sync(0xffffffffu, rd, s);
    const double ap = __shfl_down_sync(0xffffffffu, ra, s), ip = __shfl_down_sync(0xffffffffu, inv, s);
    const double cq = __shfl_down_sync(0xffffffffu, rc, s), dq = __shfl_down_sync(0xffffffffu, rd, s);
    double na = 0.0, nc = 0.0;
    if (lane >= s) {
      const double alpha = -ra * im;
      na = alpha * am;
      rb += alpha * cm;
      rd += alpha * dm;
    }
    if (lane + s < 32) {
      const double gamma = -rc * ip;
      nc = gamma * cq;
      rb += gamma * ap;
      rd += gamma * dq;
    }
    ra = na;
    rc = nc;
  }
  const double xl = rd / rb;
  const double xprev = __shfl_up_sync(0xffffffffu, xl, 1);   // lane 0: multiplied by v = 0
  // ---- 4. substitution, coalesced write-back ----
  __syncwarp();
  sd[line_pos<PPT>(r0 + PPT - 1)] = xl;
  if (PPT > 1) {
#pragma unroll
    for (int j = 0; j < PPT - 1; ++j) sd[line_pos<PPT>(r0 + j)] = u[j] - v[j] * xprev - w[j] * xl;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i = lane + 32 * j;
    if (i < n) x[base + i] = sd[line_pos<PPT>(i)];
  }
}

// ---- transfer operators -----------------------------------------------------------------
// 1-D weights between fine index i and coarse index I (coarse point I sits at fine 2I):
//   linear:      i = 2I -> 1 ; i = 2I+-1 -> 1/2 (an odd last point without right neighbour -> 1)
//   aggregation: i in {2I, 2I+1} -> 1
//   none (S<=2 or not coarsened): i == I -> 1
struct Transfer {
  int type[kMaxDims];      // 1 linear, 0 aggregation, -1 identity
  long long Sf[kMaxDims], Sc[kMaxDims];
};
// On a slab the indices of dimension N-1 are local; slabs start on even global indices, so
// coarse local I still sits at fine local 2I.  A neighbour rank shows up as fine index -1 /
// coarse index Sc (one past the owned range): `open_lo` admits the former, and the weights are
// evaluated with the virtual coarse extent Sc + 1 when `open_hi`.
__device__ __forceinline__ double tw(int type, long long i, long long I, long long Sc) {
  if (type < 0) return i == I ? 1.0 : 0.0;
  if (type == 0) return (i >> 1) == I ? 1.0 : 0.0;
  if ((i & 1) == 0) return (i >> 1) == I ? 1.0 : 0.0;
  const long long l = i >> 1, r = l + 1;
  if (r >= Sc) return l == I ? 1.0 : 0.0;
  return (I == l || I == r) ? 0.5 : 0.0;
}
// range of fine indices with non-zero weight for coarse I
__device__ __forceinline__ void fine_range(int type, long long I, long long Sf, long long& lo, long long& hi,
                                           bool open_lo = false) {
  if (type < 0) { lo = hi = I; return; }
  if (type == 0) { lo = 2 * I; hi = 2 * I + 1 < Sf ? 2 * I + 1 : Sf - 1; return; }
  const long long first = open_lo ? -1 : 0;
  lo = 2 * I - 1 > first ? 2 * I - 1 : first;
  hi = 2 * I + 1 < Sf ? 2 * I + 1 : Sf - 1;
}

// 32-bit forms of tw / fine_range for the transfer kernels (a level holds fewer than 2^31 points): the 64-bit index
// arithmetic of the general forms made k_prolong_add and k_restrict instruction-bound (ncu / launch list at 256^3:
// 0.72 and 0.45 ms against bandwidth bounds of 0.05 and 0.03 ms).  Same values.
__device__ __forceinline__ double tw32(int type, int i, int I, int Sc) {
  if (type < 0) return i == I ? 1.0 : 0.0;
  if (type == 0 || (i & 1) == 0) return (i >> 1) == I ? 1.0 : 0.0;
  const int l = i >> 1, r = l + 1;   // i = -1 (slab ghost): l = -1, r = 0
  if (r >= Sc) return l == I ? 1.0 : 0.0;
  return (I == l || I == r) ? 0.5 : 0.0;
}

// The per-dimension weights and index ranges are formed once per point; the 3^N (2^N) combinations are then
// compile-time loops.  The weights are 0, 1/2 or 1, so their products are exact in any association: the partial
// products of the outer dimensions are hoisted out of the inner loop; the terms are summed in the same order as ever.
template <int N>
__device__ __forceinline__ double restrict_point(const Transfer& t, const LevelGeom<N>& gf, const LevelGeom<N>& gc,
                                                 const double* rf, long long pc) {
  long long Il[N];
  unravel_l<N>(pc, gc, Il);
  double wt[N][3];
  int off[N][3];   // fine offset (index * stride); entries beyond the range carry weight 0
#pragma unroll
  for (int d = 0; d < N; ++d) {
    const int type = t.type[d], I = static_cast<int>(Il[d]), Sf = static_cast<int>(gf.S[d]);
    const int st = static_cast<int>(gf.stride[d]);
    int lo, hi;
    if (type < 0) {
      lo = hi = I;
    } else if (type == 0) {
      lo = 2 * I;
      hi = 2 * I + 1 < Sf ? 2 * I + 1 : Sf - 1;
    } else {
      const int first = (d == N - 1 && gf.open_lo) ? -1 : 0;
      lo = 2 * I - 1 > first ? 2 * I - 1 : first;
      hi = 2 * I + 1 < Sf ? 2 * I + 1 : Sf - 1;
    }
    const int Sce = static_cast<int>(gc.S[d]) + ((d == N - 1 && gf.open_hi) ? 1 : 0);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int i = lo + q <= hi ? lo + q : lo;
      wt[d][q] = lo + q <= hi ? tw32(type, i, I, Sce) : 0.0;
      off[d][q] = i * st;   // i == -1 on a slab: the lower ghost plane of rf
    }
  }
  double acc = 0.0;
  // m = e_0 + 3 e_1 + 9 e_2 ascending = e_{N-1} outermost
  if (N == 1) {
#pragma unroll
    for (int e0 = 0; e0 < 3; ++e0) {
      const double w = wt[0][e0];
      if (w != 0.0) acc += w * rf[off[0][e0]];
    }
  } else if (N == 2) {
#pragma unroll
    for (int e1 = 0; e1 < 3; ++e1) {
#pragma unroll
      for (int e0 = 0; e0 < 3; ++e0) {
        const double w = wt[0][e0] * wt[N > 1 ? 1 : 0][e1];
        if (w != 0.0) acc += w * rf[off[0][e0] + off[N > 1 ? 1 : 0][e1]];
      }
    }
  } else if (N == 3) {
#pragma unroll
    for (int e2 = 0; e2 < 3; ++e2) {
#pragma unroll
      for (int e1 = 0; e1 < 3; ++e1) {
        const double w12 = wt[N > 1 ? 1 : 0][e1] * wt[N > 2 ? 2 : 0][e2];
        const int q12 = off[N > 1 ? 1 : 0][e1] + off[N > 2 ? 2 : 0][e2];
#pragma unroll
        for (int e0 = 0; e0 < 3; ++e0) {
          const double w = wt[0][e0] * w12;
          if (w != 0.0) acc += w * rf[off[0][e0] + q12];
        }
      }
    }
  } else {
#pragma unroll
    for (int m = 0; m < ipow3(N); ++m) {
      double w = 1.0;
      int q = 0;
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const int e = (m / p3(d)) % 3;
        w *= wt[d][e];
        q += off[d][e];
      }
      if (w != 0.0) acc += w * rf[q];
    }
  }
  return acc;
}

template <int N>
__global__ void __launch_bounds__(kBlock) k_restrict(const Transfer* __restrict__ tp, LevelGeom<N> gf,
                                                     LevelGeom<N> gc, const double* __restrict__ rf,
                                                     double* __restrict__ rc) {
  const Transfer t = *tp;   // lives in device memory so a captured CUDA graph sees every set-up's choice
  const long long pc = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pc >= gc.l) return;
  rc[pc] = restrict_point<N>(t, gf, gc, rf, pc);
}

// x_f += P e_c
template <int N>
__device__ __forceinline__ double prolong_point(const Transfer& t, const LevelGeom<N>& gf, const LevelGeom<N>& gc,
                                                const double* ec, long long pf) {
  long long il[N];
  unravel_l<N>(pf, gf, il);
  // every fine point interpolates from at most 2 coarse points per dimension
  double wt[N][2];
  int off[N][2];
#pragma unroll
  for (int d = 0; d < N; ++d) {
    const int type = t.type[d], i = static_cast<int>(il[d]), st = static_cast<int>(gc.stride[d]);
    const int Sce = static_cast<int>(gc.S[d]) + ((d == N - 1 && gf.open_hi) ? 1 : 0);   // slab: I == Sc is the upper ghost of ec
#pragma unroll
    for (int bit = 0; bit < 2; ++bit) {
      const int I = type < 0 ? i : (i >> 1) + bit;
      // second candidate only for odd linear points
      const bool skip = (bit && (type != 1 || (i & 1) == 0)) || I >= Sce;
      wt[d][bit] = skip ? 0.0 : tw32(type, i, I, Sce);
      off[d][bit] = skip ? 0 : I * st;
    }
  }
  double acc = 0.0;
  // m = b_0 + 2 b_1 + 4 b_2 ascending; exact weights (0, 1/2, 1): outer partial products hoisted
  if (N == 3) {
#pragma unroll
    for (int b2 = 0; b2 < 2; ++b2) {
#pragma unroll
      for (int b1 = 0; b1 < 2; ++b1) {
        const double w12 = wt[N > 1 ? 1 : 0][b1] * wt[N > 2 ? 2 : 0][b2];
        const int q12 = off[N > 1 ? 1 : 0][b1] + off[N > 2 ? 2 : 0][b2];
#pragma unroll
        for (int b0 = 0; b0 < 2; ++b0) {
          const double w = wt[0][b0] * w12;
          if (w != 0.0) acc += w * ec[off[0][b0] + q12];
        }
      }
    }
  } else {
#pragma unroll
    for (int m = 0; m < (1 << N); ++m) {
      double w = 1.0;
      int q = 0;
#pragma unroll
      for (int d = 0; d < N; ++d) {
        w *= wt[d][(m >> d) & 1];
        q += off[d][(m >> d) & 1];
      }
      if (w != 0.0) acc += w * ec[q];
    }
  }
  return acc;
}

template <int N>
__global__ void __launch_bounds__(kBlock) k_prolong_add(const Transfer* __restrict__ tp, LevelGeom<N> gf,
                                                        LevelGeom<N> gc, const double* __restrict__ ec,
                                                        double* __restrict__ xf) {
  const Transfer t = *tp;
  const long long pf = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pf >= gf.l) return;
  xf[pf] += prolong_point<N>(t, gf, gc, ec, pf);
}

// ---- coarse levels of the V-cycle in ONE kernel ------------------------------------------------
// Below a few ten thousand points a level's ~20 (2-D) / ~40 (3-D) kernels are pure launch latency: the 2-D
// hierarchy of a 2048^2 grid spends two thirds of its cycle in levels of 256^2 points and fewer, and the three
// replicated levels of the slab-decomposed 3-D cycle cost 0.5-0.8 ms that no rank count shrinks.  Here ONE
// thread-block cluster (8 blocks of 512 threads on 8 SMs) runs the whole sub-cycle from the first "small" level
// down to the dense coarsest solve and back up: zebra line relaxation in every direction, residual, restriction,
// interpolation -- the same sweeps in the same order as the per-kernel path -- separated by cluster barriers
// (hardware barrier with release / acquire semantics at cluster scope: a few hundred nanoseconds, against several
// microseconds per dependent kernel of a graph).  All vectors stay in layout 0 (no transposed copies: the lines of
// direction d are read with stride g.stride[d]; the levels are L2-resident); the lines of a colour are dealt to the
// blocks, each block solving its share in batches by parallel cyclic reduction in shared memory.
constexpr int kCoarseMaxLevels = 8;
constexpr int kCoarseThreads = 512;
constexpr int kCoarseBatchPts = 2048;   // points of the lines one block relaxes at a time (4 x 16 KB of shared memory)
constexpr int kCoarseRpt = kCoarseBatchPts / kCoarseThreads;

template <int N>
struct CoarseLevelArg {
  LevelGeom<N> g;
  const mgreal_t<N>* A;   // box planes, layout 0
  double *x, *b, *r;
};
template <int N>
struct CoarseCycleArgs {
  int nlev;                          // levels lev[0] .. lev[nlev-1]; the last one is the dense coarsest level
  CoarseLevelArg<N> lev[kCoarseMaxLevels];
  const Transfer* transfers;         // transfers[i]: lev[i] -> lev[i+1]
  int dense_n;
  const double* dense_inv;           // column-major inverse of the coarsest operator
};

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

// y(pt) = b(pt) - (A x)(pt) on a replicated box level (every plane, closed boundaries)
template <int N>
__device__ __forceinline__ double coarse_residual_point(const LevelGeom<N>& g, const mgreal_t<N>* __restrict__ A,
                                                        const double* x, const double* b, long long pt) {
  using St = Stencil<N, true>;
  long long idx[N];
  unravel_l<N>(pt, g, idx);
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < St::NK; ++k) {
    long long q = pt;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      const int o = St::off(k, d);
      if ((o < 0 && idx[d] == 0) || (o > 0 && idx[d] == g.S[d] - 1)) ok = false;
      q += o * g.stride[d];
    }
    acc += static_cast<double>(A[k * g.l + pt]) * x[ok ? q : pt];
  }
  return b[pt] - acc;
}

// one colour of the zebra line relaxation along dimension D of level L, shared among the blocks of the cluster
template <int N, int D>
__device__ void coarse_line_sweep(const CoarseLevelArg<N>& L, int color, double* sh, unsigned crank, unsigned csize) {
  using St = Stencil<N, true>;
  const LevelGeom<N>& g = L.g;
  const int n = static_cast<int>(g.S[D]);
  double* sa = sh;
  double* si = sh + kCoarseBatchPts;
  double* sc = sh + 2 * kCoarseBatchPts;
  double* sd = sh + 3 * kCoarseBatchPts;
  // the other dimensions in ascending order; bit j of the colour is the parity of the j-th of them
  int od[N > 1 ? N - 1 : 1];
  long long cnt[N > 1 ? N - 1 : 1];
  long long nlines = 1;
  {
    int j = 0;
#pragma unroll
    for (int e = 0; e < N; ++e) {
      if (e == D) continue;
      od[j] = e;
      const int ce = (color >> j) & 1;
      cnt[j] = (g.S[e] - ce + 1) / 2;
      nlines *= cnt[j];
      ++j;
    }
  }
  if (nlines <= 0) return;
  // in-line plane ids
  constexpr int kc = (ipow3(N) - 1) / 2, km = kc - p3(D), kp = kc + p3(D);
  const int lines_per_batch = kCoarseBatchPts / n > 0 ? kCoarseBatchPts / n : 1;
  // block `crank` takes the line batches crank, crank + csize, ...
  const long long nbatches = (nlines + lines_per_batch - 1) / lines_per_batch;
  for (long long batch = crank; batch < nbatches; batch += csize) {
    const long long line0 = batch * lines_per_batch;
    const int nl = static_cast<int>(nlines - line0 < lines_per_batch ? nlines - line0 : lines_per_batch);
    const int npts = nl * n;
    double ra[kCoarseRpt], rb[kCoarseRpt], rc[kCoarseRpt], rd[kCoarseRpt];
    long long ptm[kCoarseRpt];
    int im_[kCoarseRpt];
#pragma unroll
    for (int m = 0; m < kCoarseRpt; ++m) {
      const int p = threadIdx.x + m * kCoarseThreads;
      ra[m] = 0.0; rb[m] = 1.0; rc[m] = 0.0; rd[m] = 0.0; ptm[m] = -1; im_[m] = 0;
      if (p < npts) {
        const int i = p % n;
        long long q = line0 + p / n;
        long long idx[N];
        idx[D] = i;
        long long pt = static_cast<long long>(i) * g.stride[D];
#pragma unroll
        for (int j = 0; j < N - 1; ++j) {
          const int ce = (color >> j) & 1;
          const long long v = ce + 2 * (q % cnt[j]);
          q /= cnt[j];
          idx[od[j]] = v;
          pt += v * g.stride[od[j]];
        }
        double rhs = L.b[pt];
#pragma unroll
        for (int k = 0; k < St::NK; ++k) {
          if (k == kc || k == km || k == kp) continue;
          long long qq = pt;
          bool ok = true;
#pragma unroll
          for (int e = 0; e < N; ++e) {
            const int o = St::off(k, e);
            if ((o < 0 && idx[e] == 0) || (o > 0 && idx[e] == g.S[e] - 1)) ok = false;
            qq += o * g.stride[e];
          }
          rhs -= static_cast<double>(L.A[k * g.l + pt]) * L.x[ok ? qq : pt];
        }
        ra[m] = i > 0 ? static_cast<double>(L.A[km * g.l + pt]) : 0.0;
        rb[m] = static_cast<double>(L.A[kc * g.l + pt]);
        rc[m] = i < n - 1 ? static_cast<double>(L.A[kp * g.l + pt]) : 0.0;
        rd[m] = rhs;
        ptm[m] = pt;
        im_[m] = i;
      }
    }
    for (int s = 1; s < n; s <<= 1) {
#pragma unroll
      for (int m = 0; m < kCoarseRpt; ++m) {
        const int p = threadIdx.x + m * kCoarseThreads;
        if (p < npts) { sa[p] = ra[m]; si[p] = fast_rcp(rb[m]); sc[p] = rc[m]; sd[p] = rd[m]; }
      }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < kCoarseRpt; ++m) {
        const int p = threadIdx.x + m * kCoarseThreads;
        if (p < npts) {
          const int i = im_[m];
          double na = 0.0, nc = 0.0;
          if (i - s >= 0) {
            const double al = -ra[m] * si[p - s];
            na = al * sa[p - s];
            rb[m] += al * sc[p - s];
            rd[m] += al * sd[p - s];
          }
          if (i + s < n) {
            const double ga = -rc[m] * si[p + s];
            nc = ga * sc[p + s];
            rb[m] += ga * sa[p + s];
            rd[m] += ga * sd[p + s];
          }
          ra[m] = na;
          rc[m] = nc;
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < kCoarseRpt; ++m)
      if (ptm[m] >= 0) L.x[ptm[m]] = rd[m] / rb[m];
    __syncthreads();   // the next batch re-uses the shared arrays
  }
}

template <int N, int D>
__device__ void coarse_smooth_dim(const CoarseLevelArg<N>& L, double* sh, unsigned crank, unsigned csize) {
  if (L.g.S[D] < 2) return;
  for (int color = 0; color < (1 << (N - 1)); ++color) {
    coarse_line_sweep<N, D>(L, color, sh, crank, csize);
    cluster_barrier();
  }
}

template <int N>
__device__ void coarse_smooth(const CoarseLevelArg<N>& L, bool reverse, double* sh, unsigned crank, unsigned csize) {
  if (!reverse) {
    coarse_smooth_dim<N, 0>(L, sh, crank, csize);
    if (N >= 2) coarse_smooth_dim<N, (N >= 2 ? 1 : 0)>(L, sh, crank, csize);
    if (N >= 3) coarse_smooth_dim<N, (N >= 3 ? 2 : 0)>(L, sh, crank, csize);
  } else {
    if (N >= 3) coarse_smooth_dim<N, (N >= 3 ? 2 : 0)>(L, sh, crank, csize);
    if (N >= 2) coarse_smooth_dim<N, (N >= 2 ? 1 : 0)>(L, sh, crank, csize);
    coarse_smooth_dim<N, 0>(L, sh, crank, csize);
  }
}

// lev[0].x = V-cycle(lev[0].b) over lev[0..nlev-1]
template <int N>
__global__ void __launch_bounds__(kCoarseThreads, 1) k_coarse_cycle(const __grid_constant__ CoarseCycleArgs<N> a) {
  extern __shared__ double sh[];
  const unsigned crank = cluster_ctarank(), csize = cluster_nctarank();
  const long long gtid = static_cast<long long>(crank) * kCoarseThreads + threadIdx.x;
  const long long gthreads = static_cast<long long>(csize) * kCoarseThreads;
  for (int li = 0; li + 1 < a.nlev; ++li) {
    const CoarseLevelArg<N>& L = a.lev[li];
    const CoarseLevelArg<N>& Lc = a.lev[li + 1];
    for (long long pt = gtid; pt < L.g.l; pt += gthreads) L.x[pt] = 0.0;
    cluster_barrier();
    coarse_smooth<N>(L, false, sh, crank, csize);
    for (long long pt = gtid; pt < L.g.l; pt += gthreads) L.r[pt] = coarse_residual_point<N>(L.g, L.A, L.x, L.b, pt);
    cluster_barrier();
    const Transfer t = a.transfers[li];
    for (long long pc = gtid; pc < Lc.g.l; pc += gthreads) Lc.b[pc] = restrict_point<N>(t, L.g, Lc.g, L.r, pc);
    cluster_barrier();
  }
  {   // coarsest level: x = Minv b, one warp per row
    const CoarseLevelArg<N>& L = a.lev[a.nlev - 1];
    const int n = a.dense_n;
    const int lane = threadIdx.x & 31;
    for (long long row = gtid >> 5; row < n; row += gthreads >> 5) {
      double t = 0.0;
      for (int c = lane; c < n; c += 32) t += a.dense_inv[static_cast<long long>(c) * n + row] * L.b[c];
      t = warp_sum(t);
      if (lane == 0) L.x[row] = t;
    }
    cluster_barrier();
  }
  for (int li = a.nlev - 2; li >= 0; --li) {
    const CoarseLevelArg<N>& L = a.lev[li];
    const CoarseLevelArg<N>& Lc = a.lev[li + 1];
    const Transfer t = a.transfers[li];
    for (long long pf = gtid; pf < L.g.l; pf += gthreads) L.x[pf] += prolong_point<N>(t, L.g, Lc.g, Lc.x, pf);
    cluster_barrier();
    coarse_smooth<N>(L, true, sh, crank, csize);
  }
}

// Separable Galerkin product: coarsen ONE dimension d of the operator (the full P^T A P with
// tensor-product transfers is the composition of these passes, each O(9) work per entry).
// Input planes live on grid gin (shape BOXIN), output box planes on gout (= gin with
// dimension d coarsened).  One thread per (output row, output offset).
template <int N, bool BOX>
__device__ __forceinline__ int plane_of(const int* o) {
  if (BOX) {
    int k = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) k += (o[d] + 1) * p3(d);
    return k;
  }
  int nz = 0, k = 0;
#pragma unroll
  for (int d = 0; d < N; ++d)
    if (o[d] != 0) { ++nz; k = 1 + 2 * d + (o[d] > 0 ? 1 : 0); }
  return nz == 0 ? 0 : (nz == 1 ? k : -1);
}

// On a slab, coarsening dimension N-1 needs the operator rows of the lower neighbour's last plane
// (`Aghost[k][plane]`, fine index -1), and keeps the couplings to the neighbours' coarse planes
// (J == -1 / J == Sc).
// D (the coarsened dimension) is a template parameter: with a run-time d the offset arrays below were indexed
// dynamically and lived in local memory (ncu at 256^3: 19 M local loads, 2.9e9 instructions for the first pass of a
// set-up, 4.1 ms), and the index arithmetic is 32-bit (a level holds fewer than 2^31 points).  Same values.
template <int N, bool BOXIN, int D>
__global__ void __launch_bounds__(kBlock) k_rap_1d(int type, LevelGeom<N> gin, LevelGeom<N> gout,
                                                   const mgreal_t<N>* __restrict__ Ain,
                                                   mgreal_t<N>* __restrict__ Aout,
                                                   const mgreal_t<N>* __restrict__ Aghost, FastDiv divl) {
  // one thread per (output row, offsets in the other dimensions): it reads the <= 3 x 3 fine
  // entries (fine row i, in-dimension offset od) once and scatters them into the three output
  // offsets O[D] = -1, 0, +1
  constexpr int KO = ipow3(N) / 3;   // offset combinations of the other dimensions
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= gout.l * KO) return;
  const int ko = static_cast<int>(divl.div(static_cast<unsigned>(tid)));   // KO * l < 2^32
  const long long pout = tid - static_cast<long long>(ko) * gout.l;
  long long idx[N];
  unravel_l<N>(pout, gout, idx);
  int O[N];
  {
    int q = ko;
#pragma unroll
    for (int e = 0; e < N; ++e) {
      if (e == D) { O[e] = 0; continue; }
      O[e] = q % 3 - 1;
      q /= 3;
    }
  }
  const int I = static_cast<int>(idx[D]);
  constexpr bool slab = D == N - 1;
  const bool olo = slab && gin.open_lo, ohi = slab && gin.open_hi;
  const int Sf = static_cast<int>(gin.S[D]);
  const int Sce = static_cast<int>(gout.S[D]) + (ohi ? 1 : 0);          // virtual coarse extent for the weights
  const int jlo = olo ? -2 : 0, jhi = Sf - 1 + (ohi ? 1 : 0);           // admissible fine columns [jlo, jhi]
  const int Jlo = olo ? -1 : 0;
  double acc[3] = {0.0, 0.0, 0.0};
  int lo, hi;
  if (type < 0) {
    lo = hi = I;
  } else if (type == 0) {
    lo = 2 * I;
    hi = 2 * I + 1 < Sf ? 2 * I + 1 : Sf - 1;
  } else {
    const int first = olo ? -1 : 0;
    lo = 2 * I - 1 > first ? 2 * I - 1 : first;
    hi = 2 * I + 1 < Sf ? 2 * I + 1 : Sf - 1;
  }
  long long pbase = 0;
#pragma unroll
  for (int e = 0; e < N; ++e)
    if (e != D) pbase += idx[e] * gin.stride[e];
  // input plane of the offsets (od in dimension D, O elsewhere), for od = -1, 0, +1
  int kin[3];
#pragma unroll
  for (int od = -1; od <= 1; ++od) {
    int o[N];
#pragma unroll
    for (int e = 0; e < N; ++e) o[e] = e == D ? od : O[e];
    kin[od + 1] = plane_of<N, BOXIN>(o);
  }
  const long long strideD = gin.stride[D];
#pragma unroll
  for (int qi = 0; qi < 3; ++qi) {   // at most three fine rows carry weight
    const int i = lo + qi;
    if (i > hi) break;
    const double wi = tw32(type, i, I, Sce);
    if (wi == 0.0) continue;
    const long long pin = pbase + i * strideD;
#pragma unroll
    for (int od = -1; od <= 1; ++od) {
      const int j = i + od;
      if (j < jlo || j > jhi) continue;
      const int k = kin[od + 1];
      if (k < 0) continue;
      const double a = wi * (i < 0 ? Aghost[k * strideD + pbase] : Ain[k * gin.l + pin]);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int J = I + q - 1;
        if (J < Jlo || J >= Sce) continue;
        acc[q] += a * tw32(type, j, J, Sce);
      }
    }
  }
  int kout0 = 0;   // plane of (O with O[D] = -1); the three outputs are p3(D) apart
#pragma unroll
  for (int e = 0; e < N; ++e) kout0 += (e == D ? 0 : O[e] + 1) * p3(e);
#pragma unroll
  for (int q = 0; q < 3; ++q) Aout[(kout0 + q * p3(D)) * gout.l + pout] = static_cast<mgreal_t<N>>(acc[q]);
}

// sum over the grid of min(|a+|,|a-|) and max(|a+|,|a-|) of the two couplings across each
// edge along every dimension: out[2d], out[2d+1].  Symmetric (diffusion-like) -> ratio near 1.
template <int N, bool BOX>
__global__ void __launch_bounds__(kBlock) k_edge_symmetry(LevelGeom<N> g, const mgreal_t<N>* __restrict__ A,
                                                          double* __restrict__ out) {
  using St = Stencil<N, BOX>;
  __shared__ double sm[2 * kMaxDims][kBlock / 32];
  const long long pt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  double vals[2 * N];
#pragma unroll
  for (int q = 0; q < 2 * N; ++q) vals[q] = 0.0;
  if (pt < g.l) {
    long long idx[N];
    unravel_l<N>(pt, g, idx);
#pragma unroll
    for (int d = 0; d < N; ++d) {
      if (idx[d] < g.S[d] - 1) {
        int kp = 0, km = 0;
#pragma unroll
        for (int k = 0; k < St::NK; ++k) {
          bool axial = true;
#pragma unroll
          for (int e = 0; e < N; ++e)
            if (e != d && St::off(k, e) != 0) axial = false;
          if (axial && St::off(k, d) > 0) kp = k;
          if (axial && St::off(k, d) < 0) km = k;
        }
        const double up = fabs(A[kp * g.l + pt]);
        const double dn = fabs(A[km * g.l + pt + g.stride[d]]);
        vals[2 * d] = up < dn ? up : dn;
        vals[2 * d + 1] = up < dn ? dn : up;
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 2 * N; ++q) {
    const double s = warp_sum(vals[q]);
    if (lane == 0) sm[q][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < 2 * N) {
    double s = 0.0;
    for (int w = 0; w < kBlock / 32; ++w) s += sm[threadIdx.x][w];
    atomicAdd(&out[threadIdx.x], s);  // only steers a heuristic: order-insensitive enough
  }
}

// ---- coarsest level: dense LU with partial pivoting by one thread block ----------------
template <int N, bool BOX>
__global__ void __launch_bounds__(kBlock) k_dense_build(LevelGeom<N> g, const mgreal_t<N>* __restrict__ A,
                                                        double* __restrict__ M) {
  using St = Stencil<N, BOX>;
  const long long pt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pt >= g.l) return;
  long long idx[N];
  unravel_l<N>(pt, g, idx);
  const long long n = g.l;
  for (int k = 0; k < St::NK; ++k) {
    long long q = pt;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      const int o = St::off(k, d);
      if ((o < 0 && idx[d] == 0) || (o > 0 && idx[d] == g.S[d] - 1)) ok = false;
      q += o * g.stride[d];
    }
    if (ok) M[q * n + pt] = A[k * g.l + pt];  // column-major: M(row pt, col q)
  }
}

__global__ void __launch_bounds__(256) k_dense_lu(int n, double* __restrict__ M, int* __restrict__ piv) {
  __shared__ int s_p;
  for (int c = 0; c < n; ++c) {
    if (threadIdx.x == 0) {
      int p = c;
      double best = fabs(M[c * n + c]);
      for (int r = c + 1; r < n; ++r) {
        const double v = fabs(M[c * n + r]);
        if (v > best) { best = v; p = r; }
      }
      s_p = p;
      piv[c] = p;
    }
    __syncthreads();
    const int p = s_p;
    if (p != c)
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double t = M[j * n + c];
        M[j * n + c] = M[j * n + p];
        M[j * n + p] = t;
      }
    __syncthreads();
    const double d = 1.0 / M[c * n + c];
    for (int r = c + 1 + threadIdx.x; r < n; r += blockDim.x) M[c * n + r] *= d;
    __syncthreads();
    // trailing update, column by column
    for (long long e = threadIdx.x; e < static_cast<long long>(n - c - 1) * (n - c - 1); e += blockDim.x) {
      const int r = c + 1 + static_cast<int>(e % (n - c - 1));
      const int j = c + 1 + static_cast<int>(e / (n - c - 1));
      M[j * n + r] -= M[c * n + r] * M[j * n + c];
    }
    __syncthreads();
  }
}

// Column j of the inverse from the LU factors (one thread per column, run once per set-up);
// the V-cycle then applies the coarsest solve as one dense mat-vec instead of 2n dependent steps.
__global__ void __launch_bounds__(64) k_dense_invert(int n, const double* __restrict__ M, const int* __restrict__ piv,
                                                     double* __restrict__ Minv, double* __restrict__ work) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double* y = work + static_cast<long long>(j) * n;
  for (int i = 0; i < n; ++i) y[i] = i == j ? 1.0 : 0.0;
  for (int c = 0; c < n; ++c) {
    const int p = piv[c];
    if (p != c) { const double t = y[c]; y[c] = y[p]; y[p] = t; }
    const double yc = y[c];
    if (yc != 0.0)
      for (int r = c + 1; r < n; ++r) y[r] -= M[c * n + r] * yc;
  }
  for (int c = n - 1; c >= 0; --c) {
    y[c] /= M[c * n + c];
    const double yc = y[c];
    for (int r = 0; r < c; ++r) y[r] -= M[c * n + r] * yc;
  }
  for (int i = 0; i < n; ++i) Minv[static_cast<long long>(j) * n + i] = y[i];   // column-major
}

// Both steps for n <= 64 (the usual coarsest grids: 8 x 8, 4 x 4 x 4) by ONE block in shared memory: the two kernels
// above walk global memory one dependent access at a time (0.14 + 0.69 ms per set-up at n = 64, a third of a 2-D
// set-up); here the factors and the n right-hand sides live in shared memory.  Same operations in the same order,
// hence the same inverse.  Y is stored row-major (thread j owns column j: consecutive threads, consecutive banks).
constexpr int kDenseSmall = 64;
__global__ void __launch_bounds__(256) k_dense_invert_small(int n, const double* __restrict__ Min,
                                                            double* __restrict__ Minv) {
  extern __shared__ double dsm[];
  double* M = dsm;            // [n][n] column-major: M(r, c) = M[c * n + r]
  double* Y = dsm + n * n;    // Y(i, j) = Y[i * n + j]
  __shared__ int piv[kDenseSmall];
  __shared__ int s_p;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) M[e] = Min[e];
  __syncthreads();
  for (int c = 0; c < n; ++c) {
    if (threadIdx.x == 0) {
      int p = c;
      double best = fabs(M[c * n + c]);
      for (int r = c + 1; r < n; ++r) {
        const double v = fabs(M[c * n + r]);
        if (v > best) { best = v; p = r; }
      }
      s_p = p;
      piv[c] = p;
    }
    __syncthreads();
    const int p = s_p;
    if (p != c)
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double t = M[j * n + c];
        M[j * n + c] = M[j * n + p];
        M[j * n + p] = t;
      }
    __syncthreads();
    const double d = 1.0 / M[c * n + c];
    __syncthreads();   // everyone has read the pivot before the column is scaled
    for (int r = c + 1 + threadIdx.x; r < n; r += blockDim.x) M[c * n + r] *= d;
    __syncthreads();
    for (int e = threadIdx.x; e < (n - c - 1) * (n - c - 1); e += blockDim.x) {
      const int r = c + 1 + e % (n - c - 1);
      const int j = c + 1 + e / (n - c - 1);
      M[j * n + r] -= M[c * n + r] * M[j * n + c];
    }
    __syncthreads();
  }
  const int j = threadIdx.x;
  if (j < n) {
    for (int i = 0; i < n; ++i) Y[i * n + j] = i == j ? 1.0 : 0.0;
    for (int c = 0; c < n; ++c) {
      const int p = piv[c];
      if (p != c) { const double t = Y[c * n + j]; Y[c * n + j] = Y[p * n + j]; Y[p * n + j] = t; }
      const double yc = Y[c * n + j];
      if (yc != 0.0)
        for (int r = c + 1; r < n; ++r) Y[r * n + j] -= M[c * n + r] * yc;
    }
    for (int c = n - 1; c >= 0; --c) {
      Y[c * n + j] /= M[c * n + c];
      const double yc = Y[c * n + j];
      for (int r = 0; r < c; ++r) Y[r * n + j] -= M[c * n + r] * yc;
    }
    for (int i = 0; i < n; ++i) Minv[static_cast<long long>(j) * n + i] = Y[i * n + j];   // column-major
  }
}

// x = Minv b, one warp per row
__global__ void __launch_bounds__(256) k_dense_apply(int n, const double* __restrict__ Minv,
                                                     const double* __restrict__ b, double* __restrict__ x) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  double t = 0.0;
  for (int c = threadIdx.x & 31; c < n; c += 32) t += Minv[static_cast<long long>(c) * n + r] * b[c];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) x[r] = t;
}

// ---- GMRES vector kernels ---------------------------------------------------------------
// out[q] = <V_q, w> for q < j (V: j vectors of length n, contiguous) ; out[j] = <w, w>
__global__ void __launch_bounds__(kBlock) k_multi_dot(long long n, int j, const double* __restrict__ V,
                                                      const double* __restrict__ w, double* __restrict__ part) {
  // grid.y = q (0..j), grid.x blocks stride over n; partial sums -> part[q][blockIdx.x]
  __shared__ double sm[kBlock / 32];
  const int q = blockIdx.y;
  const double* v = q < j ? V + static_cast<long long>(q) * n : w;
  double s = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    s += v[i] * w[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < kBlock / 32; ++k) t += sm[k];
    part[static_cast<long long>(q) * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void k_finish_dots(int nq, int nblocks, const double* __restrict__ part, double* __restrict__ out) {
  // one warp per dot product; fixed summation order (strided partials, then butterfly)
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  double t = 0.0;
  for (int b = threadIdx.x & 31; b < nblocks; b += 32) t += part[static_cast<long long>(q) * nblocks + b];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) out[q] = t;
}
// w -= sum_q h[q] V_q
__global__ void __launch_bounds__(kBlock) k_multi_axpy(long long n, int j, const double* __restrict__ V,
                                                       const double* __restrict__ h, double* __restrict__ w,
                                                       double sign) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = w[i];
  for (int q = 0; q < j; ++q) acc += sign * h[q] * V[static_cast<long long>(q) * n + i];
  w[i] = acc;
}
__global__ void __launch_bounds__(kBlock) k_scale_copy(long long n, double a, const double* __restrict__ src,
                                                       double* __restrict__ dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = a * src[i];
}
// dst = src / sqrt(*norm2) (nothing when the norm vanishes or is not finite): the host's 1.0 / sqrt(h), same bits
__global__ void __launch_bounds__(kBlock) k_scale_copy_rsqrt(long long n, const double* __restrict__ norm2,
                                                             const double* __restrict__ src, double* __restrict__ dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const double h = sqrt(norm2[0]);
  if (!(h > 0.0) || !isfinite(h)) return;
  const double a = 1.0 / h;
  if (i < n) dst[i] = a * src[i];
}

// Flexible GMRES(m), right-preconditioned: apply(x, b, y, mode) is y = A x (mode 0) or y = b - A x
// (mode 1), precond(v, z) is z ~ A^{-1} v, dots(n, j, V, w, out) returns <V_q, w> (q < j) and <w, w>
// on the host.  Shared by the multigrid solver and the block line-relaxation solver.
struct GmresWork {
  DevBuf<double> V, Z, w, hdev;
  double* h = nullptr;   // pinned host scratch, >= 3 (restart + 2) doubles
  // Optional: dots_dev(n, j, V, w, out) leaves <V_q, w> (q < j) and <w, w> in DEVICE memory without synchronising.
  // With it the Gram-Schmidt coefficients never visit the host between their reduction and the update that uses
  // them: one host synchronisation per Krylov iteration (for the convergence test) instead of five.
  std::function<void(long long, int, const double*, const double*, double*)> dots_dev;
  int restart_alloc = 0;
  int layout = 0;        // 0: FGMRES(restart_alloc) buffers, 1: BiCGSTAB buffers
};
template <class Apply, class Precond, class Dots>
int fgmres_run(cudaStream_t s, long long n, GmresWork& gw, const double* b, double* x, const econpdes_options& opt,
               Apply&& spmv_top, Precond&& precondition, Dots&& dots, int* iters_out, double* relres_out) {
  {
    DevBuf<double>&V_ = gw.V, &Z_ = gw.Z, &w_ = gw.w, &hdev_ = gw.hdev;
    int& restart_alloc_ = gw.restart_alloc;
    double* h_ = gw.h;
    const int m = opt.lin_restart > 0 && opt.lin_restart <= 60 ? opt.lin_restart : 40;
    if (m != restart_alloc_ || gw.layout != 0) {
      gw.layout = 0;
      restart_alloc_ = m;
      V_.alloc(static_cast<size_t>(m + 1) * n);
      Z_.alloc(static_cast<size_t>(m) * n);
      w_.alloc(static_cast<size_t>(n));
      hdev_.alloc(3 * static_cast<size_t>(m + 2));
    }
    const bool dev_gs = static_cast<bool>(gw.dots_dev);
    std::vector<double> H((m + 1) * m, 0.0), cs(m), sn(m), gvec(m + 1), hcol(m + 2), yv(m);
    double* hh = h_;  // pinned scratch (>= m + 2 doubles)
    EP_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), s));
    // ||b||
    dots(n, 0, V_.p, b, hh);
    const double bnorm = std::sqrt(hh[0]);
    int total = 0;
    double relres = 0.0;
    if (!(bnorm > 0.0) || !std::isfinite(bnorm)) {
      if (iters_out) *iters_out = 0;
      if (relres_out) *relres_out = 0.0;
      return std::isfinite(bnorm) ? 0 : 1;
    }
    double prev_true = INFINITY;
    int status = 1;
    for (int cycle = 0; cycle < 50 && total < opt.lin_maxiters; ++cycle) {
      // r = b - A x
      spmv_top(x, b, w_.p, 1);
      dots(n, 0, V_.p, w_.p, hh);
      const double beta = std::sqrt(hh[0]);
      relres = beta / bnorm;
      if (!std::isfinite(beta)) { status = 1; break; }
      if (relres <= opt.lin_rtol || (cycle > 0 && relres <= 10 * opt.lin_rtol)) { status = 0; break; }
      // stagnation at the attainable accuracy of the system: accept what LU would also deliver
      if (cycle > 0 && beta > 0.5 * prev_true) { status = relres <= 1e-8 ? 0 : 1; break; }
      prev_true = beta;
      EP_LAUNCH(k_scale_copy, blocks_for(n), kBlock, 0, s, n, 1.0 / beta, w_.p, V_.p);
      std::fill(gvec.begin(), gvec.end(), 0.0);
      gvec[0] = beta;
      int j = 0;
      for (; j < m && total < opt.lin_maxiters; ++j, ++total) {
        double* zj = Z_.p + static_cast<size_t>(j) * n;
        precondition(V_.p + static_cast<size_t>(j) * n, zj);
        spmv_top(zj, nullptr, w_.p, 0);
        // classical Gram-Schmidt, twice
        std::fill(hcol.begin(), hcol.end(), 0.0);
        double hnext;
        if (dev_gs) {
          // both passes, the norm of what is left and the next basis vector without leaving the device
          double* hd = hdev_.p;
          const int stride = m + 2;
          for (int pass = 0; pass < 2; ++pass) {
            gw.dots_dev(n, j + 1, V_.p, w_.p, hd + pass * stride);
            EP_LAUNCH(k_multi_axpy, blocks_for(n), kBlock, 0, s, n, j + 1, V_.p, hd + pass * stride, w_.p, -1.0);
          }
          gw.dots_dev(n, 0, V_.p, w_.p, hd + 2 * stride);
          EP_LAUNCH(k_scale_copy_rsqrt, blocks_for(n), kBlock, 0, s, n, hd + 2 * stride, w_.p,
                    V_.p + static_cast<size_t>(j + 1) * n);
          EP_CUDA(cudaMemcpyAsync(hh, hd, (2 * stride + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
          EP_CUDA(cudaStreamSynchronize(s));
          for (int q = 0; q <= j; ++q) hcol[q] = hh[q] + hh[stride + q];
          hnext = std::sqrt(hh[2 * stride]);
          hcol[j + 1] = hnext;
          if (!std::isfinite(hnext)) { status = 1; j = -1; break; }
        } else {
          for (int pass = 0; pass < 2; ++pass) {
            dots(n, j + 1, V_.p, w_.p, hh);
            EP_CUDA(cudaMemcpyAsync(hdev_.p, hh, (j + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
            EP_LAUNCH(k_multi_axpy, blocks_for(n), kBlock, 0, s, n, j + 1, V_.p, hdev_.p, w_.p, -1.0);
            for (int q = 0; q <= j; ++q) hcol[q] += hh[q];
            EP_CUDA(cudaStreamSynchronize(s));  // hh is reused by the next dots()
          }
          dots(n, 0, V_.p, w_.p, hh);
          hnext = std::sqrt(hh[0]);
          hcol[j + 1] = hnext;
          if (!std::isfinite(hnext)) { status = 1; j = -1; break; }
          if (hnext > 0.0)
            EP_LAUNCH(k_scale_copy, blocks_for(n), kBlock, 0, s, n, 1.0 / hnext, w_.p, V_.p + static_cast<size_t>(j + 1) * n);
        }
        // Givens rotations
        for (int q = 0; q < j; ++q) {
          const double t = cs[q] * hcol[q] + sn[q] * hcol[q + 1];
          hcol[q + 1] = -sn[q] * hcol[q] + cs[q] * hcol[q + 1];
          hcol[q] = t;
        }
        const double denom = std::hypot(hcol[j], hcol[j + 1]);
        cs[j] = denom > 0 ? hcol[j] / denom : 1.0;
        sn[j] = denom > 0 ? hcol[j + 1] / denom : 0.0;
        hcol[j] = denom;
        gvec[j + 1] = -sn[j] * gvec[j];
        gvec[j] = cs[j] * gvec[j];
        for (int q = 0; q <= j; ++q) H[q * m + j] = hcol[q];
        relres = std::fabs(gvec[j + 1]) / bnorm;
        if (relres <= opt.lin_rtol || hnext == 0.0) { ++j; ++total; break; }
      }
      if (j < 0) break;
      // x += Z y with H y = g
      for (int q = j - 1; q >= 0; --q) {
        double t = gvec[q];
        for (int r2 = q + 1; r2 < j; ++r2) t -= H[q * m + r2] * yv[r2];
        yv[q] = t / H[q * m + q];
      }
      for (int q = 0; q < j; ++q) hh[q] = yv[q];
      EP_CUDA(cudaMemcpyAsync(hdev_.p, hh, j * sizeof(double), cudaMemcpyHostToDevice, s));
      EP_LAUNCH(k_multi_axpy, blocks_for(n), kBlock, 0, s, n, j, Z_.p, hdev_.p, x, 1.0);
      EP_CUDA(cudaStreamSynchronize(s));
    }
    if (status != 0) {
      // final check of the true residual (the loop may have ended on the iteration cap)
      spmv_top(x, b, w_.p, 1);
      dots(n, 0, V_.p, w_.p, hh);
      relres = std::sqrt(hh[0]) / bnorm;
      status = (std::isfinite(relres) && relres <= 1e-8) ? 0 : 1;
    }
    if (iters_out) *iters_out = total;
    if (relres_out) *relres_out = relres;
    return status;
  }
}

// out = a x + b y + c z (null operands are skipped)
__global__ void __launch_bounds__(kBlock) k_lincomb(long long n, double a, const double* __restrict__ x, double b,
                                                    const double* __restrict__ y, double c,
                                                    const double* __restrict__ z, double* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  if (x != nullptr) acc += a * x[i];
  if (y != nullptr) acc += b * y[i];
  if (z != nullptr) acc += c * z[i];
  out[i] = acc;
}

// Right-preconditioned BiCGSTAB (options.linear_solver = ECONPDES_LIN_BICGSTAB_MG): two preconditioner
// applications and two operator applications per iteration, seven work vectors instead of the 2m + 1 of
// FGMRES(m); same operator / preconditioner / dot-product callbacks and the same acceptance rule as fgmres_run
// (relative residual <= lin_rtol, or stagnation at the attainable accuracy <= 1e-8).  `iters_out` counts
// preconditioner applications, so the figure is comparable with FGMRES iterations.
template <class Apply, class Precond, class Dots>
int bicgstab_run(cudaStream_t s, long long n, GmresWork& gw, const double* b, double* x, const econpdes_options& opt,
                 Apply&& spmv_top, Precond&& precondition, Dots&& dots, int* iters_out, double* relres_out) {
  // V_ holds the vectors the dot products run against: [rhat, t]; Z_ the rest: r, p, v, phat, shat (s shares r)
  if (gw.layout != 1) {
    gw.layout = 1;
    gw.restart_alloc = 2;   // the dot-product scratch of the callers is sized restart_alloc + 2
    gw.V.alloc(static_cast<size_t>(2) * n);
    gw.Z.alloc(static_cast<size_t>(5) * n);
    gw.w.alloc(static_cast<size_t>(n));
    gw.hdev.alloc(8);
  }
  double* rhat = gw.V.p;
  double* t = gw.V.p + n;
  double* r = gw.Z.p;
  double* p = gw.Z.p + n;
  double* v = gw.Z.p + 2 * n;
  double* phat = gw.Z.p + 3 * n;
  double* shat = gw.Z.p + 4 * n;
  double* hh = gw.h;
  const int nb = blocks_for(n);
  EP_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), s));
  dots(n, 0, rhat, b, hh);
  const double bnorm = std::sqrt(hh[0]);
  if (iters_out) *iters_out = 0;
  if (relres_out) *relres_out = 0.0;
  if (!(bnorm > 0.0) || !std::isfinite(bnorm)) return std::isfinite(bnorm) ? 0 : 1;
  EP_CUDA(cudaMemcpyAsync(r, b, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  EP_CUDA(cudaMemcpyAsync(rhat, b, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  EP_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double), s));
  EP_CUDA(cudaMemsetAsync(v, 0, n * sizeof(double), s));
  double rho = 1.0, alpha = 1.0, omega = 1.0, relres = 1.0, best = 1.0;
  int applied = 0, status = 1, since_best = 0;
  while (applied + 2 <= opt.lin_maxiters) {
    dots(n, 1, rhat, r, hh);                  // hh[0] = <rhat, r>, hh[1] = <r, r>
    const double rho_new = hh[0];
    relres = std::sqrt(hh[1]) / bnorm;
    if (!std::isfinite(relres)) break;
    if (relres <= opt.lin_rtol) { status = 0; break; }
    if (relres < 0.5 * best) { best = relres; since_best = 0; }
    else if (++since_best >= 8) { status = relres <= 1e-8 ? 0 : 1; break; }   // stagnation
    if (rho_new == 0.0 || omega == 0.0) break;   // breakdown
    const double beta = (rho_new / rho) * (alpha / omega);
    EP_LAUNCH(k_lincomb, nb, kBlock, 0, s, n, 1.0, r, beta, p, -beta * omega, v, p);   // p = r + beta (p - omega v)
    precondition(p, phat);
    spmv_top(phat, nullptr, v, 0);
    applied += 1;
    dots(n, 1, rhat, v, hh);                  // hh[0] = <rhat, v>
    if (hh[0] == 0.0 || !std::isfinite(hh[0])) break;
    alpha = rho_new / hh[0];
    EP_LAUNCH(k_lincomb, nb, kBlock, 0, s, n, 1.0, r, -alpha, v, 0.0, static_cast<const double*>(nullptr), r);   // s (in r)
    precondition(r, shat);
    spmv_top(shat, nullptr, t, 0);
    applied += 1;
    dots(n, 2, rhat, r, hh);                  // hh[1] = <t, s>
    const double ts = hh[1];
    dots(n, 0, rhat, t, hh);                  // hh[0] = <t, t>
    const double tt = hh[0];
    if (!(tt > 0.0) || !std::isfinite(tt)) {  // s is already (numerically) zero: x += alpha phat finishes
      EP_LAUNCH(k_lincomb, nb, kBlock, 0, s, n, 1.0, x, alpha, phat, 0.0, static_cast<const double*>(nullptr), x);
      omega = 0.0;
      continue;
    }
    omega = ts / tt;
    EP_LAUNCH(k_lincomb, nb, kBlock, 0, s, n, 1.0, x, alpha, phat, omega, shat, x);
    EP_LAUNCH(k_lincomb, nb, kBlock, 0, s, n, 1.0, r, -omega, t, 0.0, static_cast<const double*>(nullptr), r);
    rho = rho_new;
  }
  // the recurrence residual drifts from the true one: judge the solve on b - A x
  spmv_top(x, b, gw.w.p, 1);
  dots(n, 0, rhat, gw.w.p, hh);
  relres = std::sqrt(hh[0]) / bnorm;
  if (status != 0 || !(relres <= 10 * opt.lin_rtol)) status = (std::isfinite(relres) && relres <= 1e-8) ? 0 : 1;
  if (iters_out) *iters_out = applied;
  if (relres_out) *relres_out = relres;
  return status;
}

// ---- slab plumbing of the distributed levels -----------------------------------------------
constexpr int kMaxRanks = 16;
struct SlabMap {           // every rank's owned range of dimension N-1 at one level
  int world;
  long long first[kMaxRanks], count[kMaxRanks];
};
// last owned plane of every operator plane, packed [k][plane] for the neighbour above
template <class T>
__global__ void __launch_bounds__(kBlock) k_pack_last_plane(int nk, long long l, long long plane,
                                                            const T* __restrict__ A, T* __restrict__ out) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= nk * plane) return;
  const long long k = t / plane, p = t % plane;
  out[t] = A[k * l + (l - plane) + p];
}
// local [nk][l_loc] -> send buffer [nk][maxl]
template <class T>
__global__ void __launch_bounds__(kBlock) k_pack_slab(int nk, long long lloc, long long maxl,
                                                      const T* __restrict__ src, T* __restrict__ dst) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= nk * lloc) return;
  const long long k = t / lloc, p = t % lloc;
  dst[k * maxl + p] = src[k * lloc + p];
}
// all-gathered [rank][nk][maxl] -> full [nk][l_glob] (slabs are contiguous along dimension N-1)
template <class T>
__global__ void __launch_bounds__(kBlock) k_unpack_slabs(SlabMap m, int nk, long long plane, long long maxl,
                                                         long long lglob, const T* __restrict__ all,
                                                         T* __restrict__ full) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(nk) * lglob) return;
  const long long k = t / lglob, pt = t % lglob;
  const long long i = pt / plane;
  int r = 0;
  while (r + 1 < m.world && i >= m.first[r + 1]) ++r;
  full[t] = all[(static_cast<long long>(r) * nk + k) * maxl + (pt - m.first[r] * plane)];
}

// FP64 Jacobian planes -> storage type of the hierarchy (level 0 of the preconditioner)
template <class T>
__global__ void __launch_bounds__(kBlock) k_to_mgreal(long long n, const double* __restrict__ src,
                                                      T* __restrict__ dst) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = static_cast<T>(src[i]);
}

// ---- the solver ---------------------------------------------------------------------------
template <int N>
struct Level {
  long long S[N];                       // sizes held on this device (slab: local count of dimension N-1)
  long long l = 0;
  bool box = true;
  int nk = 0;
  // distributed (slab) level: vectors x, b, r, tx carry one ghost plane per side
  bool dist = false, open_lo = false, open_hi = false;
  long long plane = 0;                  // prod S[0..N-2]
  long long Sg_last = 0;                // global size of dimension N-1
  SlabMap map{};                        // owned ranges of all ranks at this level
  using mgreal = mgreal_t<N>;
  const mgreal* A = nullptr;           // layout 0
  DevBuf<mgreal> Aown;                  // owned storage for coarse levels
  DevBuf<mgreal> Aperm[N];              // operator in layout d (d >= 1)
  long long Sd[N][N];                   // sizes of layout d
  unsigned mask = 0xffffffffu;          // LevelGeom::mask of layout 0
  unsigned mask_d[N];                   // ... of layout d
  Transfer to_coarse;                   // valid if not the coarsest level
  DevBuf<double> x, b, r, tx, tb;       // work vectors (tx/tb: transposed copies)
};

template <int N>
class KrylovSolver final : public LinearSolver {
 public:
  using mgreal = mgreal_t<N>;
  explicit KrylovSolver(Ctx& c) : c_(c) {
    EP_CUDA(cudaMallocHost(&h_, 256 * sizeof(double)));
    gw_.h = h_;
    sym_.alloc(2 * kMaxDims);
    if (const char* wl = std::getenv("ECONPDES_WARP_LINES_MAX")) {
      warp_lines_max_ = std::atoi(wl);
      if (warp_lines_max_ > 256) warp_lines_max_ = 256;
    }
    const char* e = std::getenv("ECONPDES_NO_GRAPH");
    use_graph_ = !(e && e[0] == '1');
    // the slab cycle interleaves NCCL calls with its kernels: NCCL operations are capturable, every
    // rank captures the same sequence (ECONPDES_DIST_GRAPH=0 replays eagerly instead)
    if (c.dist != nullptr) {
      const char* dg = std::getenv("ECONPDES_DIST_GRAPH");
      if (dg && dg[0] == '0') use_graph_ = false;
    }
    if (const char* rb = std::getenv("ECONPDES_REPLICATE_BELOW")) replicate_below_ = std::atoll(rb);
  }
  ~KrylovSolver() override {
    if (h_) cudaFreeHost(h_);
    if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
    for (cudaEvent_t e : ev_)
      if (e) cudaEventDestroy(e);
  }

  void new_step() override { need_setup_ = true; }

  // The hierarchy is a preconditioner inside FLEXIBLE GMRES, whose operator is always the current
  // Jacobian: within one Newton solve (same time step, same ypost) the Jacobian often moves little
  // between iterations, so the hierarchy of the first iteration is kept while that pays: once the lagged
  // solves need, on average, more extra Krylov iterations than a set-up costs (both timed on the device),
  // lagging is switched off for the rest of this problem (upwind switching in the consumption-saving
  // model changes the Jacobian too much; Bansal-Yaron at 1024^2 saves 10 %).  Newton iterates are
  // unchanged: every solve is driven to the same relative residual.  ECONPDES_LAG_PRECOND=0 rebuilds
  // every time.
  int solve(const double* planes, const double* b, double* x, const econpdes_options& opt, int* iters,
            double* relres) override {
    static const bool lag_env = [] { const char* e = std::getenv("ECONPDES_LAG_PRECOND"); return !(e && e[0] == '0'); }();
    cudaStream_t s = c_.stream;
    if (ev_[0] == nullptr)
      for (cudaEvent_t& e : ev_) EP_CUDA(cudaEventCreate(&e));
    bool rebuild = !lag_env || lag_off_ || need_setup_ || lev_.empty();
    // lagging pays while the extra Krylov iterations of a lagged solve cost less than the set-up it
    // saves, on average over the lagged solves of this problem (single solves are noisy: their
    // right-hand sides differ)
    // (slab-decomposed solves must take the same decision on every rank -- the set-up exchanges operator
    // planes -- so there the rule uses iteration counts only, which are identical everywhere)
    if (!rebuild && n_stale_ >= 3) {
      const double avg_extra = extra_sum_ / n_stale_;
      // a set-up costs about 1.5 Krylov iterations on one device (measured: 13.5 ms against 9 ms at 256^3,
      // 2.1 ms against 1.3 ms at 2048^2).  The rule uses iteration counts only, so the decision -- and with it
      // every Krylov iterate -- is reproducible from run to run.
      const bool costly = c_.dist != nullptr ? avg_extra > 0.5 * base_iters_ + 1.0 : avg_extra > 1.5;
      if (costly) {
        rebuild = true;
        lag_off_ = true;
      }
    }
    need_setup_ = false;
    auto run = [&](bool with_setup, int* it) {
      if (with_setup) {
        EP_CUDA(cudaEventRecord(ev_[0], s));
        setup(planes);
        EP_CUDA(cudaEventRecord(ev_[1], s));
      } else {
        top_A_ = planes;
      }
      int rc;
      {
        ProfScope ps("lin fgmres total", s);
        EP_CUDA(cudaEventRecord(ev_[2], s));
        rc = fgmres(b, x, opt, it, relres);
        EP_CUDA(cudaEventRecord(ev_[3], s));
      }
      EP_CUDA(cudaEventSynchronize(ev_[3]));
      float ms = 0.f;
      if (with_setup) {
        EP_CUDA(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
        setup_ms_ = ms;
      }
      EP_CUDA(cudaEventElapsedTime(&ms, ev_[2], ev_[3]));
      if (*it > 0) iter_ms_ = ms / *it;
      return rc;
    };
    int it = 0;
    int rc = run(rebuild, &it);
    int total = it;
    if (rebuild) {
      base_iters_ = it;
      stale_iters_ = 0;
    } else if (rc != 0) {   // do not let a stale preconditioner turn into a rejected time step
      rc = run(true, &it);
      total += it;
      base_iters_ = it;
      stale_iters_ = 0;
      lag_off_ = true;
    } else {
      stale_iters_ = it;
      extra_sum_ += it - base_iters_;
      ++n_stale_;
    }
    if (iters) *iters = total;
    return rc;
  }

  void bench(const double* planes, const double* v, int reps, double* flush, size_t flush_bytes,
             double* ms_out) override {
    cudaStream_t s = c_.stream;
    cudaEvent_t e0, e1;
    EP_CUDA(cudaEventCreate(&e0));
    EP_CUDA(cudaEventCreate(&e1));
    auto timed = [&](auto&& fn) {
      double tot = 0.0;
      for (int r = 0; r < reps + 1; ++r) {
        EP_CUDA(cudaMemsetAsync(flush, 0, flush_bytes, s));   // evict L2 between timed launches
        EP_CUDA(cudaEventRecord(e0, s));
        fn();
        EP_CUDA(cudaEventRecord(e1, s));
        EP_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        EP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) tot += ms;   // first one is warm-up
      }
      return tot / reps;
    };
    ms_out[2] = timed([&] { setup(planes); });
    Level<N>& L0 = *lev_[0];
    w_.alloc(static_cast<size_t>(L0.l));
    DevBuf<double> z;
    z.alloc(static_cast<size_t>(L0.l));
    ms_out[0] = timed([&] { spmv_top(v, nullptr, w_.p, 0); });
    ms_out[1] = timed([&] { precondition(v, z.p); });
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }

 private:
  Ctx& c_;
  std::vector<std::unique_ptr<Level<N>>> lev_;
  DevBuf<double> dense_, sym_, w_, dots_, part_, dense_inv_, dense_work_;
  GmresWork gw_;
  DevBuf<mgreal> A0_, rap_tmp_[2], rap_send_, rap_ghost_;   // level-0 copy of the Jacobian planes; Galerkin scratch
  DevBuf<int> piv_;
  DevBuf<Transfer> transfers_dev_;
  // slab path: ghosted copy of the SpMV operand, block-local copy of the operator for the
  // preconditioner (couplings that leave the slab dropped), all-gather landing area
  DevBuf<double> xh_, bloc_;
  DevBuf<unsigned char> gsend_, grecv_;   // all-gather staging (operator planes or vectors)
  long long replicate_below_ = 48LL * 48 * 48;   // levels with at most this many points run replicated
  cudaGraphExec_t graph_exec_ = nullptr;
  const mgreal* graph_A_ = nullptr;
  int graph_nlev_ = 0;
  unsigned long long graph_sig_ = 0;
  int64_t graph_kernels_ = 0;
  bool use_graph_ = true;
  bool need_setup_ = true;
  int base_iters_ = 0, stale_iters_ = 0;   // Krylov iterations right after a set-up / with the lagged hierarchy
  bool lag_off_ = false;
  double extra_sum_ = 0.0;   // sum over the lagged solves of (their iterations - the fresh solve's)
  int n_stale_ = 0;
  double setup_ms_ = 0.0, iter_ms_ = 0.0;   // device time of the last set-up / per Krylov iteration
  cudaEvent_t ev_[4] = {};
  // lines of up to this many points take the warp-per-line kernel (max 256).  Default 0: on B200 the
  // block-wide PCR hides the global-load latency of the 27 operator planes better (256^3: 260 us
  // vs 302 us per color sweep, both ~55-60 % of HBM peak; coarse levels tie), so the warp kernel
  // stays an option (ECONPDES_WARP_LINES_MAX) rather than the default.
  int warp_lines_max_ = 0;
  double* h_ = nullptr;
  int dense_n_ = 0;

  template <bool BOX>
  void spmv(const Level<N>& L, const double* x, const double* b, double* y, int mode) {
    LevelGeom<N> g = make_geom<N>(L.S);
    g.mask = L.mask;
    if (BOX && masked(L)) EP_LAUNCH((k_spmv<N, BOX, mgreal, true>), blocks_for(L.l), kBlock, 0, c_.stream, g, L.A, x, b, y, mode);
    else EP_LAUNCH((k_spmv<N, BOX, mgreal>), blocks_for(L.l), kBlock, 0, c_.stream, g, L.A, x, b, y, mode);
  }
  void spmv(const Level<N>& L, const double* x, const double* b, double* y, int mode) {
    if (L.box) spmv<true>(L, x, b, y, mode); else spmv<false>(L, x, b, y, mode);
  }
  static bool masked(const Level<N>& L) { return L.box && (L.mask & ((1u << ipow3(N)) - 1u)) != ((1u << ipow3(N)) - 1u); }
  // The operator of the Krylov iteration: on a slab, A couples to the neighbour ranks' edge
  // planes, so the operand is copied into ghosted storage and its halos are exchanged first.
  void spmv_top(const double* x, const double* b, double* y, int mode) {
    ProfScope ps("lin spmv", c_.stream);
    Level<N>& L0 = *lev_[0];
    if (c_.dist == nullptr) {   // the Krylov operator is the FP64 Jacobian, not the hierarchy's copy
      LevelGeom<N> g0 = make_geom<N>(L0.S);
      g0.mask = L0.mask;
      if (c_.box && masked(L0)) EP_LAUNCH((k_spmv<N, true, double, true>), blocks_for(L0.l), kBlock, 0, c_.stream, g0, top_A_, x, b, y, mode);
      else if (c_.box) EP_LAUNCH((k_spmv<N, true, double>), blocks_for(L0.l), kBlock, 0, c_.stream, g0, top_A_, x, b, y, mode);
      else EP_LAUNCH((k_spmv<N, false, double>), blocks_for(L0.l), kBlock, 0, c_.stream, g0, top_A_, x, b, y, mode);
      return;
    }
    cudaStream_t s = c_.stream;
    xh_.pad = static_cast<size_t>(c_.halo);
    xh_.alloc(static_cast<size_t>(L0.l));
    EP_CUDA(cudaMemcpyAsync(xh_.p, x, L0.l * sizeof(double), cudaMemcpyDeviceToDevice, s));
    c_.exchange_halo(xh_.p);
    LevelGeom<N> g = make_geom<N>(L0.S);
    g.mask = L0.mask;
    g.open_lo = c_.dist->has_lo();
    g.open_hi = c_.dist->has_hi();
    if (c_.box && masked(L0)) EP_LAUNCH((k_spmv<N, true, double, true>), blocks_for(L0.l), kBlock, 0, s, g, top_A_, xh_.p, b, y, mode);
    else if (c_.box) EP_LAUNCH((k_spmv<N, true, double>), blocks_for(L0.l), kBlock, 0, s, g, top_A_, xh_.p, b, y, mode);
    else EP_LAUNCH((k_spmv<N, false, double>), blocks_for(L0.l), kBlock, 0, s, g, top_A_, xh_.p, b, y, mode);
  }
  const double* top_A_ = nullptr;

  // geometry of a level (layout 0, or layout d with dimension d fastest); on a slab the last
  // layout dimension is still N-1 for every line direction d < N-1
  LevelGeom<N> geom(const Level<N>& L) const {
    LevelGeom<N> g = make_geom<N>(L.S);
    g.open_lo = L.dist && L.open_lo;
    g.open_hi = L.dist && L.open_hi;
    g.mask = L.mask;
    return g;
  }
  LevelGeom<N> geom_layout(const Level<N>& L, int d) const {
    LevelGeom<N> g = make_geom<N>(L.Sd[d]);
    g.open_lo = L.dist && L.open_lo;
    g.open_hi = L.dist && L.open_hi;
    g.mask = d == 0 ? L.mask : L.mask_d[d];
    return g;
  }
  void halo(const Level<N>& L, double* v) { c_.dist->exchange_halo(v, L.plane, L.S[N - 1], c_.stream); }

  // [nk][l_loc] slabs of every rank -> [nk][l_glob] on every rank
  template <class T>
  void gather_slabs(const SlabMap& m, int nk, long long plane, const T* local, T* full) {
    cudaStream_t s = c_.stream;
    const int me = c_.dist->rank();
    long long maxcount = 0, total = 0;
    for (int r = 0; r < m.world; ++r) { maxcount = m.count[r] > maxcount ? m.count[r] : maxcount; total += m.count[r]; }
    const long long maxl = maxcount * plane, lloc = m.count[me] * plane, lglob = total * plane;
    const size_t per_rank = static_cast<size_t>(nk) * maxl * sizeof(T);
    gsend_.alloc(per_rank);
    grecv_.alloc(static_cast<size_t>(m.world) * per_rank);
    T* snd = reinterpret_cast<T*>(gsend_.p);
    T* rcv = reinterpret_cast<T*>(grecv_.p);
    EP_LAUNCH(k_pack_slab<T>, blocks_for(nk * lloc), kBlock, 0, s, nk, lloc, maxl, local, snd);
    c_.dist->all_gather_raw(snd, rcv, per_rank, s);
    EP_LAUNCH(k_unpack_slabs<T>, blocks_for(nk * lglob), kBlock, 0, s, m, nk, plane, maxl, lglob, rcv, full);
  }

  // Build the hierarchy for the operator in `planes` (called once per Jacobian).
  // Slab path: the fine levels stay distributed like the grid (ghost planes + halo exchanges);
  // once a level is small (<= replicate_below_ points) its operator is all-gathered and every
  // rank runs the rest of the cycle redundantly -- those levels are launch-latency bound, so
  // replicating them costs less than their halo exchanges would.
  void setup(const double* planes) {
    cudaStream_t s = c_.stream;
    ProfScope ps("lin setup", s);
    Dist* dd = c_.dist;
    if (lev_.empty()) {
      auto L0 = std::make_unique<Level<N>>();
      for (int d = 0; d < N; ++d) L0->S[d] = c_.S[d];
      L0->S[N - 1] = c_.slab_count;   // == S[N-1] on a single device
      L0->l = c_.l;
      L0->box = c_.box;
      L0->nk = c_.nk;
      L0->Sg_last = c_.S[N - 1];
      if (c_.box) L0->mask = c_.plane_mask;
      if (dd != nullptr) {
        if (dd->world() > kMaxRanks) throw std::invalid_argument("slab path: at most 16 ranks");
        L0->dist = true;
        L0->open_lo = dd->has_lo();
        L0->open_hi = dd->has_hi();
        L0->plane = c_.halo;
        L0->map.world = dd->world();
        // every rank derives every rank's range (econpdes_slab_range arithmetic)
        const long long units = c_.S[N - 1] / 2, base = units / dd->world(), rem = units % dd->world();
        for (int r = 0; r < dd->world(); ++r) {
          const long long uf = r * base + (r < rem ? r : rem), uc = base + (r < rem ? 1 : 0);
          L0->map.first[r] = 2 * uf;
          L0->map.count[r] = 2 * uc + (r == dd->world() - 1 ? c_.S[N - 1] % 2 : 0);
        }
        if (L0->map.first[dd->rank()] != c_.slab_first || L0->map.count[dd->rank()] != c_.slab_count)
          throw std::invalid_argument("slab path: ranges must come from econpdes_slab_range");
      }
      lev_.push_back(std::move(L0));
    }
    top_A_ = planes;
    if constexpr (std::is_same_v<mgreal, double>) {
      lev_[0]->A = planes;
    } else {   // level 0 of the preconditioner: the Jacobian planes in the hierarchy's storage type
      const long long cnt = static_cast<long long>(c_.nk) * c_.l;
      if (A0_.cap < static_cast<size_t>(cnt)) {   // planes outside the mask are never written: zero them once
        A0_.alloc(static_cast<size_t>(cnt));
        A0_.zero(s);
      }
      A0_.alloc(static_cast<size_t>(cnt));
      for (int k = 0; k < c_.nk; ++k)
        if ((lev_[0]->mask >> k) & 1u)
          EP_LAUNCH(k_to_mgreal<mgreal>, 148 * 4, kBlock, 0, s, c_.l, planes + static_cast<size_t>(k) * c_.l,
                    A0_.p + static_cast<size_t>(k) * c_.l);
      lev_[0]->A = A0_.p;
    }
    size_t li = 0;
    while (true) {
      Level<N>& L = *lev_[li];
      prepare_level(L);
      bool coarsest = !L.dist;
      for (int d = 0; d < N; ++d)
        if (L.S[d] > (N == 3 ? kCoarseMax3D : kCoarseMax2D)) coarsest = false;
      if (coarsest || li + 1 >= kMaxLevels) break;
      // transfer type per dimension from the symmetry of the couplings
      sym_.zero(s);
      if (L.box) EP_LAUNCH((k_edge_symmetry<N, true>), blocks_for(L.l), kBlock, 0, s, make_geom<N>(L.S), L.A, sym_.p);
      else EP_LAUNCH((k_edge_symmetry<N, false>), blocks_for(L.l), kBlock, 0, s, make_geom<N>(L.S), L.A, sym_.p);
      double symh[2 * kMaxDims] = {};
      if (L.dist) {   // every rank must take the same decision: fold all ranks' sums in rank order
        const double* g = c_.gather_to_host(sym_.p, 2 * N);
        for (int r = 0; r < dd->world(); ++r)
          for (int q = 0; q < 2 * N; ++q) symh[q] += g[r * 2 * N + q];
      } else {
        EP_CUDA(cudaMemcpyAsync(h_, sym_.p, 2 * N * sizeof(double), cudaMemcpyDeviceToHost, s));
        EP_CUDA(cudaStreamSynchronize(s));
        for (int q = 0; q < 2 * N; ++q) symh[q] = h_[q];
      }
      Transfer t{};
      long long Sc[N];   // sizes of the coarse level as held on this device if it stays distributed
      for (int d = 0; d < N; ++d) {
        const long long Sfull = (L.dist && d == N - 1) ? L.Sg_last : L.S[d];
        t.Sf[d] = L.S[d];
        if (Sfull <= 2) {
          t.type[d] = -1;
          Sc[d] = L.S[d];
        } else {
          const double ratio = symh[2 * d + 1] > 0 ? symh[2 * d] / symh[2 * d + 1] : 1.0;
          t.type[d] = ratio > kLinearThreshold ? 1 : 0;
          Sc[d] = (L.S[d] + 1) / 2;
        }
        t.Sc[d] = Sc[d];
      }
      // does the coarse level stay distributed?
      SlabMap cmap{};
      bool coarse_dist = false;
      long long Scg_last = Sc[N - 1], coarse_glob = 1;
      if (L.dist) {
        if (t.type[N - 1] < 0) throw std::runtime_error("slab path: slab dimension too short to coarsen");
        Scg_last = (L.Sg_last + 1) / 2;
        cmap.world = L.map.world;
        coarse_dist = true;
        for (int r = 0; r < cmap.world; ++r) {
          cmap.first[r] = L.map.first[r] / 2;
          cmap.count[r] = (L.map.count[r] + 1) / 2;
          const bool last = r == cmap.world - 1;
          if (cmap.first[r] % 2 != 0 || cmap.count[r] < 2 || (!last && cmap.count[r] % 2 != 0)) coarse_dist = false;
        }
        coarse_glob = Scg_last;
        for (int d = 0; d + 1 < N; ++d) coarse_glob *= Sc[d];
        if (coarse_glob <= replicate_below_) coarse_dist = false;
      }
      L.to_coarse = t;
      if (lev_.size() <= li + 1) lev_.push_back(std::make_unique<Level<N>>());
      Level<N>& Lc = *lev_[li + 1];
      Lc.dist = coarse_dist;
      Lc.map = cmap;
      Lc.Sg_last = Scg_last;
      Lc.open_lo = coarse_dist && L.open_lo;
      Lc.open_hi = coarse_dist && L.open_hi;
      Lc.l = 1;
      for (int d = 0; d < N; ++d) Lc.S[d] = Sc[d];
      if (L.dist && !coarse_dist) Lc.S[N - 1] = Scg_last;   // replicated: the whole coarse grid
      for (int d = 0; d < N; ++d) Lc.l *= Lc.S[d];
      Lc.plane = Lc.l / Lc.S[N - 1];
      Lc.box = true;
      Lc.nk = ipow3(N);
      Lc.Aown.alloc(static_cast<size_t>(Lc.nk) * Lc.l);
      Lc.A = Lc.Aown.p;
      // separable Galerkin product, one coarsened dimension per pass
      {
        long long Sin[N];
        for (int d = 0; d < N; ++d) Sin[d] = L.S[d];
        const mgreal* in = L.A;
        bool inbox = L.box;
        int npass = 0, pass = 0;
        for (int d = 0; d < N; ++d) npass += t.type[d] >= 0 ? 1 : 0;
        const bool regather = L.dist && !coarse_dist;   // local coarse slab first, then all-gather
        for (int d = 0; d < N; ++d) {
          if (t.type[d] < 0) continue;
          long long Sout[N];
          for (int e = 0; e < N; ++e) Sout[e] = Sin[e];
          Sout[d] = t.Sc[d];
          LevelGeom<N> gin = make_geom<N>(Sin), gout = make_geom<N>(Sout);
          mgreal* out;
          if (++pass == npass && !regather) {
            out = Lc.Aown.p;
          } else {
            rap_tmp_[pass & 1].alloc(static_cast<size_t>(ipow3(N)) * gout.l);
            out = rap_tmp_[pass & 1].p;
          }
          const mgreal* ghost = nullptr;
          if (L.dist && d == N - 1) {
            gin.open_lo = L.open_lo;
            gin.open_hi = L.open_hi;
            if (t.type[d] == 1) {   // rows of the lower neighbour's last plane, all operator planes
              const int nkin = inbox ? ipow3(N) : 1 + 2 * N;
              const long long pl = gin.stride[N - 1];
              rap_send_.alloc(static_cast<size_t>(nkin) * pl);
              rap_ghost_.alloc(static_cast<size_t>(nkin) * pl);
              EP_LAUNCH(k_pack_last_plane<mgreal>, blocks_for(nkin * pl), kBlock, 0, s, nkin, gin.l, pl, in, rap_send_.p);
              dd->send_hi_recv_lo_raw(rap_send_.p, rap_ghost_.p, static_cast<size_t>(nkin) * pl * sizeof(mgreal), s);
              ghost = rap_ghost_.p;
            }
          }
          const long long nthreads = gout.l * (ipow3(N) / 3);
          FastDiv divl;
          divl.set(static_cast<unsigned>(gout.l));
#define EP_RAP(DD)                                                                                                        \
  do {                                                                                                                \
    if (inbox) EP_LAUNCH((k_rap_1d<N, true, DD>), blocks_for(nthreads), kBlock, 0, s, t.type[d], gin, gout, in, out, ghost, divl); \
    else EP_LAUNCH((k_rap_1d<N, false, DD>), blocks_for(nthreads), kBlock, 0, s, t.type[d], gin, gout, in, out, ghost, divl);      \
  } while (0)
          if (d == 0) EP_RAP(0);
          else if (d == 1) EP_RAP((N > 1 ? 1 : 0));
          else EP_RAP((N > 2 ? 2 : 0));
#undef EP_RAP
          in = out;
          inbox = true;
          for (int e = 0; e < N; ++e) Sin[e] = Sout[e];
        }
        if (regather) gather_slabs(cmap, ipow3(N), Lc.plane, in, Lc.Aown.p);
      }
      ++li;
    }
    nlev_ = static_cast<int>(li) + 1;
    {
      Transfer host[kMaxLevels] = {};
      for (int q = 0; q + 1 < nlev_; ++q) host[q] = lev_[q]->to_coarse;
      transfers_dev_.alloc(kMaxLevels);
      EP_CUDA(cudaMemcpyAsync(transfers_dev_.p, host, sizeof(host), cudaMemcpyHostToDevice, s));
      EP_CUDA(cudaStreamSynchronize(s));
    }
    // dense LU of the coarsest operator
    Level<N>& Lc = *lev_[nlev_ - 1];
    dense_n_ = static_cast<int>(Lc.l);
    dense_.alloc(static_cast<size_t>(dense_n_) * dense_n_);
    piv_.alloc(static_cast<size_t>(dense_n_));
    dense_.zero(s);
    if (Lc.box) EP_LAUNCH((k_dense_build<N, true>), blocks_for(Lc.l), kBlock, 0, s, make_geom<N>(Lc.S), Lc.A, dense_.p);
    else EP_LAUNCH((k_dense_build<N, false>), blocks_for(Lc.l), kBlock, 0, s, make_geom<N>(Lc.S), Lc.A, dense_.p);
    dense_inv_.alloc(static_cast<size_t>(dense_n_) * dense_n_);
    static const bool small_dense = [] { const char* e = std::getenv("ECONPDES_DENSE_SMALL"); return !(e && e[0] == '0'); }();
    if (small_dense && dense_n_ <= kDenseSmall) {
      const size_t smem = 2 * static_cast<size_t>(dense_n_) * dense_n_ * sizeof(double);
      static bool cfg = false;
      if (!cfg) {
        EP_CUDA(cudaFuncSetAttribute(k_dense_invert_small, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     2 * kDenseSmall * kDenseSmall * static_cast<int>(sizeof(double))));
        cfg = true;
      }
      EP_LAUNCH(k_dense_invert_small, 1, 256, smem, s, dense_n_, dense_.p, dense_inv_.p);
    } else {
      EP_LAUNCH(k_dense_lu, 1, 256, 0, s, dense_n_, dense_.p, piv_.p);
      dense_work_.alloc(static_cast<size_t>(dense_n_) * dense_n_);
      EP_LAUNCH(k_dense_invert, (dense_n_ + 63) / 64, 64, 0, s, dense_n_, dense_.p, piv_.p, dense_inv_.p, dense_work_.p);
    }
    choose_coarse_start();
  }
  int nlev_ = 0;
  // first level of the sub-cycle that runs as one cluster kernel (k_coarse_cycle); 0 = none
  int coarse_li_ = 0;

  void choose_coarse_start() {
    // OFF by default: measured on B200 the one-cluster sub-cycle is SLOWER than the kernels it replaces (V-cycle at
    // 2048^2: 1.07 ms per-kernel, 1.09 / 1.20 / 1.32 / 1.61 ms with the levels of <= 1 200 / 5 000 / 17 000 / 70 000
    // points in the cluster kernel; 256^3: 5.49 against 5.50 / 5.80 / 5.80 / 6.72 ms) although it removes two thirds
    // of the launches: every phase is a cluster barrier plus a global-memory round trip (~2.7 us, ncu: 28 warps
    // stalled on barriers per issue, 10 % issue utilisation), no faster than a ~3 us kernel of a CUDA graph that
    // spreads the same phase over 148 SMs.  ECONPDES_COARSE_CLUSTER=1 turns it on (read at every set-up).
    const char* en = std::getenv("ECONPDES_COARSE_CLUSTER");
    const bool enabled = en != nullptr && en[0] == '1';
    const char* mp = std::getenv("ECONPDES_COARSE_MAX_POINTS");
    const long long max_pts = mp ? std::atoll(mp) : 5000LL;
    coarse_li_ = 0;
    if (!enabled) return;
    for (int li = 1; li + 1 < nlev_; ++li) {   // at least one smoothed level above the dense one
      bool ok = nlev_ - li <= kCoarseMaxLevels;
      for (int q = li; q < nlev_ && ok; ++q) {
        const Level<N>& L = *lev_[q];
        ok = !L.dist && L.l <= max_pts && (q == nlev_ - 1 || (L.box && L.mask == 0xffffffffu));
        for (int d = 0; d < N && ok; ++d) ok = L.S[d] <= kCoarseBatchPts;
      }
      if (ok) { coarse_li_ = li; return; }
    }
  }

  // lev_[coarse_li_]->x = V-cycle(lev_[coarse_li_]->b) over the levels coarse_li_ .. nlev_-1, one launch
  void launch_coarse_cycle() {
    static const int cluster = [] {
      const char* e = std::getenv("ECONPDES_COARSE_CLUSTER_SIZE");
      const int v = e ? std::atoi(e) : 8;
      return v == 1 || v == 2 || v == 4 || v == 8 ? v : 8;
    }();
    CoarseCycleArgs<N> a{};
    a.nlev = nlev_ - coarse_li_;
    for (int q = 0; q < a.nlev; ++q) {
      Level<N>& L = *lev_[coarse_li_ + q];
      a.lev[q].g = make_geom<N>(L.S);
      a.lev[q].A = L.A;
      a.lev[q].x = L.x.p;
      a.lev[q].b = L.b.p;
      a.lev[q].r = L.r.p;
    }
    a.transfers = transfers_dev_.p + coarse_li_;
    a.dense_n = dense_n_;
    a.dense_inv = dense_inv_.p;
    const size_t smem = 4 * static_cast<size_t>(kCoarseBatchPts) * sizeof(double);
    static bool cfg = false;
    if (!cfg) {
      EP_CUDA(cudaFuncSetAttribute(k_coarse_cycle<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      cfg = true;
    }
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(cluster);
    lc.blockDim = dim3(kCoarseThreads);
    lc.dynamicSmemBytes = smem;
    lc.stream = c_.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    EP_CUDA(cudaLaunchKernelEx(&lc, k_coarse_cycle<N>, a));
    g_kernel_launches.fetch_add(1);
  }

  void prepare_level(Level<N>& L) {
    cudaStream_t s = c_.stream;
    const size_t l = static_cast<size_t>(L.l);
    const size_t pad = L.dist ? static_cast<size_t>(L.plane) : 0;
    L.x.pad = L.b.pad = L.r.pad = L.tx.pad = L.tb.pad = pad;
    L.x.alloc(l); L.b.alloc(l); L.r.alloc(l); L.tx.alloc(l); L.tb.alloc(l);
    const int ndirs = L.dist ? N - 1 : N;   // lines along the slab dimension would cross ranks
    for (int d = 1; d < ndirs; ++d) {
      int perm[N];
      layout_dims<N>(L.S, d, L.Sd[d], perm);
      if (L.Aperm[d].cap < static_cast<size_t>(L.nk) * l) {   // planes outside the mask stay zero
        L.Aperm[d].alloc(static_cast<size_t>(L.nk) * l);
        L.Aperm[d].zero(s);
      }
      L.Aperm[d].alloc(static_cast<size_t>(L.nk) * l);
      const TransposeShape ts = transpose_shape<N>(L.S, d);
      L.mask_d[d] = L.box && L.mask != 0xffffffffu ? 0u : 0xffffffffu;
      for (int k = 0; k < L.nk; ++k) {
        if (!((L.mask >> k) & 1u)) continue;
        // plane k (offset vector o) lands at the plane whose offset vector is o permuted
        int kd = 0;
        if (L.box) {
          for (int e = 0; e < N; ++e) kd += (((k / p3(perm[e])) % 3)) * p3(e);
        } else {
          if (k == 0) kd = 0;
          else {
            const int dim = (k - 1) / 2, sgn = (k - 1) % 2;
            int slot = 0;
            for (int e = 0; e < N; ++e) if (perm[e] == dim) slot = e;
            kd = 1 + 2 * slot + sgn;
          }
        }
        if (L.mask_d[d] != 0xffffffffu) L.mask_d[d] |= 1u << kd;
        launch_transpose(s, L.A + static_cast<size_t>(k) * l, L.Aperm[d].p + static_cast<size_t>(kd) * l, ts, false);
      }
    }
    for (int e = 0; e < N; ++e) L.Sd[0][e] = L.S[e];
  }

  template <bool BOX>
  void line_sweep_layout(Level<N>& L, int d, const mgreal* A, const double* b, double* x) {
    if (BOX && masked(L)) line_sweep_layout_m<BOX, true>(L, d, A, b, x);
    else line_sweep_layout_m<BOX, false>(L, d, A, b, x);
  }
  template <bool BOX, bool MASKED>
  void line_sweep_layout_m(Level<N>& L, int d, const mgreal* A, const double* b, double* x) {
    const long long* Sd = L.Sd[d];
    const int n = static_cast<int>(Sd[0]);
    const LevelGeom<N> g = geom_layout(L, d);
    const int ncolors = 1 << (N - 1);
    int T = ((n + 31) / 32) * 32;
    int rpt = 1;
    if (T > 1024) { rpt = (n + 1023) / 1024; T = 1024; }
    const size_t smem = 4 * static_cast<size_t>(n) * sizeof(double);
    if (smem > 200 * 1024 || rpt > 8) throw std::runtime_error("line smoother: grid dimension too long (max 6400 points)");
    for (int color = 0; color < ncolors; ++color) {
      long long nlines = 1;
      for (int e = 1; e < N; ++e) {
        const int ce = (color >> (e - 1)) & 1;
        nlines *= (Sd[e] - ce + 1) / 2;
      }
      // measured on B200: V-cycle at 2048^2 1.46 -> 1.19 ms, C2 (64 x 4096) solve -19 %; ECONPDES_CHUNK_LINES=0
      // restores the block-wide PCR
      static const bool chunk_lines = [] { const char* e = std::getenv("ECONPDES_CHUNK_LINES"); return !(e && e[0] == '0'); }();
      // 3-D (27 FP32 planes, bandwidth-bound): fewer instructions and barriers still buy 8 % (V-cycle at 256^3
      // 9.2 -> 8.5 ms); ECONPDES_CHUNK_LINES_3D=0 restores the block-wide PCR
      static const bool chunk_lines_3d = [] { const char* e = std::getenv("ECONPDES_CHUNK_LINES_3D"); return !(e && e[0] == '0'); }();
      if (nlines > 0 && n <= 4096 &&   // 4096 = 1024 threads x 4 rows
          ((chunk_lines && N == 2 && n >= 256) || (chunk_lines_3d && N == 3 && n >= 128))) {
        // long lines of the 5-plane 2-D levels: four consecutive rows per thread, chunked solve
        constexpr int P = 4;
        const int Tc = (((n + P - 1) / P + 31) / 32) * 32;
        const size_t csmem = 12 * static_cast<size_t>(Tc) * sizeof(double);
        static bool cfg = false;
        if (!cfg) {
          EP_CUDA(cudaFuncSetAttribute(k_line_zebra_chunk<N, BOX, P, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024));
          cfg = true;
        }
        EP_LAUNCH((k_line_zebra_chunk<N, BOX, P, MASKED>), static_cast<unsigned>(nlines), Tc, csmem, c_.stream, g, A, b, x, color);
      } else if (nlines > 0 && n <= warp_lines_max_) {
        // one warp per line, 8 lines per block
        const int ppt = n <= 32 ? 1 : (n <= 64 ? 2 : (n <= 128 ? 4 : 8));
        const unsigned blocks = static_cast<unsigned>((nlines + 7) / 8);
        const size_t wsmem = 8 * 4 * static_cast<size_t>(32 * ppt + 32) * sizeof(double);
#define EP_WLINE(P)                                                                                          \
  do {                                                                                                       \
    static bool cfg = false;                                                                                 \
    if (!cfg) {                                                                                              \
      EP_CUDA(cudaFuncSetAttribute(k_line_zebra_warp<N, BOX, P, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   100 * 1024));                                                             \
      cfg = true;                                                                                            \
    }                                                                                                        \
    EP_LAUNCH((k_line_zebra_warp<N, BOX, P, MASKED>), blocks, 256, wsmem, c_.stream, g, A, b, x, color, nlines);     \
  } while (0)
        if (ppt == 1) EP_WLINE(1);
        else if (ppt == 2) EP_WLINE(2);
        else if (ppt == 4) EP_WLINE(4);
        else EP_WLINE(8);
#undef EP_WLINE
      } else if (nlines > 0) {
#define EP_LINE(R)                                                                                          \
  do {                                                                                                      \
    static bool cfg = false;                                                                                \
    if (!cfg) {                                                                                             \
      EP_CUDA(cudaFuncSetAttribute(k_line_zebra<N, BOX, R, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                   200 * 1024));                                                            \
      cfg = true;                                                                                           \
    }                                                                                                       \
    EP_LAUNCH((k_line_zebra<N, BOX, R, MASKED>), static_cast<unsigned>(nlines), T, smem, c_.stream, g, A, b, x, color); \
  } while (0)
        if (rpt == 1) EP_LINE(1);
        else if (rpt == 2) EP_LINE(2);
        else if (rpt <= 4) EP_LINE(4);
        else EP_LINE(8);
#undef EP_LINE
      }
      // Slab: the top color bit is the parity of the plane index along dimension N-1 (slabs start
      // on even planes, so local parity = global parity).  A rank's edge planes need the
      // neighbour's edge plane of the OTHER parity, which changed in the other half of the sweep:
      // one halo exchange after each half reproduces the single-device sweep exactly.
      if (L.dist && (color == ncolors / 2 - 1 || color == ncolors - 1)) halo(L, x);
    }
  }

  // one alternating-direction zebra line smoothing pass (dims ascending or descending)
  void smooth(Level<N>& L, const double* b, double* x, bool reverse) {
    cudaStream_t s = c_.stream;
    const int ndirs = L.dist ? N - 1 : N;
    for (int q = 0; q < ndirs; ++q) {
      const int d = reverse ? ndirs - 1 - q : q;
      if (L.S[d] < 2) continue;
      if (d == 0) {
        if (L.box) line_sweep_layout<true>(L, 0, L.A, b, x); else line_sweep_layout<false>(L, 0, L.A, b, x);
      } else {
        TransposeShape ts = transpose_shape<N>(L.S, d), tsx = ts;
        const long long g = L.dist ? L.plane : 0;   // ghost planes travel with x (N == 3, d == 1: per-plane batches)
        if (L.dist) tsx.batches += 2;
        if (!L.dist && 2 * ts.batches <= 65535) {
          launch_transpose2<double>(s, x, L.tx.p, b, L.tb.p, ts);
        } else {
          launch_transpose(s, x - g, L.tx.p - g, tsx, false);
          launch_transpose(s, b, L.tb.p, ts, false);
        }
        if (L.box) line_sweep_layout<true>(L, d, L.Aperm[d].p, L.tb.p, L.tx.p);
        else line_sweep_layout<false>(L, d, L.Aperm[d].p, L.tb.p, L.tx.p);
        launch_transpose(s, L.tx.p - g, x - g, tsx, true);
      }
    }
  }

  // x = V-cycle(b) at level li starting from x = 0
  void vcycle(int li, const double* b, double* x) {
    cudaStream_t s = c_.stream;
    Level<N>& L = *lev_[li];
    if (coarse_li_ > 0 && li == coarse_li_ && b == L.b.p && x == L.x.p) {
      launch_coarse_cycle();
      return;
    }
    if (li == nlev_ - 1) {
      EP_LAUNCH(k_dense_apply, (dense_n_ + 7) / 8, 256, 0, s, dense_n_, dense_inv_.p, b, x);
      return;
    }
    Level<N>& Lc = *lev_[li + 1];
    const LevelGeom<N> gf = geom(L);
    if (!L.dist) {
      EP_CUDA(cudaMemsetAsync(x, 0, L.l * sizeof(double), s));
      smooth(L, b, x, false);
      spmv(L, x, b, L.r.p, 1);
      EP_LAUNCH((k_restrict<N>), blocks_for(Lc.l), kBlock, 0, s, transfers_dev_.p + li, gf, make_geom<N>(Lc.S), L.r.p,
                Lc.b.p);
      vcycle(li + 1, Lc.b.p, Lc.x.p);
      EP_LAUNCH((k_prolong_add<N>), blocks_for(L.l), kBlock, 0, s, transfers_dev_.p + li, gf, make_geom<N>(Lc.S),
                Lc.x.p, x);
      smooth(L, b, x, true);
      return;
    }
    // ---- distributed level: x, r (and the coarse x) carry ghost planes ----
    Dist& dd = *c_.dist;
    const bool linear = L.to_coarse.type[N - 1] == 1;
    static const char* kSm[4] = {"vc L0 smooth", "vc L1 smooth", "vc L2 smooth", "vc L3+ smooth"};
    static const char* kTr[4] = {"vc L0 transfer", "vc L1 transfer", "vc L2 transfer", "vc L3+ transfer"};
    const bool prof = Prof::enabled() && !use_graph_;
    const int pl = li < 3 ? li : 3;
    EP_CUDA(cudaMemsetAsync(x - L.plane, 0, (L.l + 2 * L.plane) * sizeof(double), s));
    {
      ProfScope ps(prof ? kSm[pl] : "", prof ? s : nullptr);
      smooth(L, b, x, false);   // leaves the ghosts of x current
    }
    ProfScope* pt = prof ? new ProfScope(kTr[pl], s) : nullptr;
    if (masked(L)) EP_LAUNCH((k_spmv<N, true, mgreal, true>), blocks_for(L.l), kBlock, 0, s, gf, L.A, x, b, L.r.p, 1);
    else if (L.box) EP_LAUNCH((k_spmv<N, true, mgreal>), blocks_for(L.l), kBlock, 0, s, gf, L.A, x, b, L.r.p, 1);
    else EP_LAUNCH((k_spmv<N, false, mgreal>), blocks_for(L.l), kBlock, 0, s, gf, L.A, x, b, L.r.p, 1);
    // full weighting reaches one fine plane below the slab
    if (linear) dd.send_hi_recv_lo(L.r.p + (L.S[N - 1] - 1) * L.plane, L.r.p - L.plane, L.plane, s);
    // this rank's part of the coarse grid
    long long Scl[N];
    for (int d = 0; d < N; ++d) Scl[d] = Lc.S[d];
    const long long cfirst = L.map.first[dd.rank()] / 2;
    Scl[N - 1] = (L.S[N - 1] + 1) / 2;
    const LevelGeom<N> gcl = make_geom<N>(Scl);
    const double* ec;
    if (Lc.dist) {
      EP_LAUNCH((k_restrict<N>), blocks_for(gcl.l), kBlock, 0, s, transfers_dev_.p + li, gf, gcl, L.r.p, Lc.b.p);
      delete pt;
      vcycle(li + 1, Lc.b.p, Lc.x.p);
      pt = prof ? new ProfScope(kTr[pl], s) : nullptr;
      // interpolation reaches one coarse plane above the slab
      if (linear) dd.send_lo_recv_hi(Lc.x.p, Lc.x.p + Lc.l, Lc.plane, s);
      ec = Lc.x.p;
    } else {
      // replicated coarse level: all-gather the restricted residual, every rank runs the rest
      bloc_.alloc(static_cast<size_t>(gcl.l));
      EP_LAUNCH((k_restrict<N>), blocks_for(gcl.l), kBlock, 0, s, transfers_dev_.p + li, gf, gcl, L.r.p, bloc_.p);
      SlabMap cm{};
      cm.world = L.map.world;
      for (int r = 0; r < cm.world; ++r) { cm.first[r] = L.map.first[r] / 2; cm.count[r] = (L.map.count[r] + 1) / 2; }
      gather_slabs(cm, 1, Lc.plane, bloc_.p, Lc.b.p);
      delete pt;
      {
        ProfScope ps(prof ? "vc replicated levels" : "", prof ? s : nullptr);
        vcycle(li + 1, Lc.b.p, Lc.x.p);
      }
      pt = prof ? new ProfScope(kTr[pl], s) : nullptr;
      ec = Lc.x.p + cfirst * Lc.plane;   // the plane above the slab is simply the next one
    }
    EP_LAUNCH((k_prolong_add<N>), blocks_for(L.l), kBlock, 0, s, transfers_dev_.p + li, gf, gcl, ec, x);
    halo(L, x);
    delete pt;
    {
      ProfScope ps(prof ? kSm[pl] : "", prof ? s : nullptr);
      smooth(L, b, x, true);
    }
  }

  // z = V-cycle(v).  The cycle is ~25 small launches per level, launch-latency bound on all but
  // the finest levels, so it is captured once into a CUDA graph (buffers, grids and shared-memory
  // sizes depend only on the grid shape; per-set-up choices live in device memory) and replayed.
  void precondition(const double* v, double* z) {
    cudaStream_t s = c_.stream;
    ProfScope ps("lin vcycle", s);
    Level<N>& L0 = *lev_[0];
    if (!use_graph_) {
      if (L0.dist) {   // the cycle's x carries ghost planes: run it in the level's own buffer
        vcycle(0, v, L0.x.p);
        EP_CUDA(cudaMemcpyAsync(z, L0.x.p, L0.l * sizeof(double), cudaMemcpyDeviceToDevice, s));
      } else {
        vcycle(0, v, z);
      }
      return;
    }
    // host-side choices baked into a captured cycle: which levels are distributed and whether
    // their slab-dimension transfer needs ghost exchanges
    unsigned long long sig = 0;
    for (int q = 0; q < nlev_; ++q)
      sig = sig * 7 + (lev_[q]->dist ? 1 + (lev_[q]->to_coarse.type[N - 1] == 1 ? 1 : 0) : 0);
    sig = sig * 31 + static_cast<unsigned long long>(coarse_li_);   // ... and where the one-kernel sub-cycle starts
    if (graph_exec_ == nullptr || graph_A_ != L0.A || graph_nlev_ != nlev_ || graph_sig_ != sig) {
      if (graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
      vcycle(0, L0.b.p, L0.x.p);   // uncaptured once: function attributes, allocations
      EP_CUDA(cudaStreamSynchronize(s));
      const int64_t k0 = g_kernel_launches.load();
      cudaGraph_t graph = nullptr;
      EP_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      try {
        vcycle(0, L0.b.p, L0.x.p);
      } catch (...) {
        cudaStreamEndCapture(s, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      EP_CUDA(cudaStreamEndCapture(s, &graph));
      graph_kernels_ = g_kernel_launches.load() - k0;
      g_kernel_launches.fetch_sub(graph_kernels_);   // captured, not launched
      EP_CUDA(cudaGraphInstantiate(&graph_exec_, graph, 0));
      cudaGraphDestroy(graph);
      graph_A_ = L0.A;
      graph_nlev_ = nlev_;
      graph_sig_ = sig;
    }
    EP_CUDA(cudaMemcpyAsync(L0.b.p, v, L0.l * sizeof(double), cudaMemcpyDeviceToDevice, s));
    EP_CUDA(cudaGraphLaunch(graph_exec_, s));
    g_kernel_launches.fetch_add(graph_kernels_);
    EP_CUDA(cudaMemcpyAsync(z, L0.x.p, L0.l * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }

  // <V_q, w> (q < j) and <w, w> into device memory, no synchronisation (single device)
  void dots_dev(long long n, int j, const double* V, const double* w, double* dev_out) {
    cudaStream_t s = c_.stream;
    const int nb = 148 * 2;
    part_.alloc(static_cast<size_t>(nb) * (gw_.restart_alloc + 2));
    EP_LAUNCH(k_multi_dot, dim3(nb, j + 1), kBlock, 0, s, n, j, V, w, part_.p);
    EP_LAUNCH(k_finish_dots, (j + 1 + 7) / 8, 256, 0, s, j + 1, nb, part_.p, dev_out);
  }

  void dots(long long n, int j, const double* V, const double* w, double* host_out) {
    // host_out[0..j-1] = <V_q, w>, host_out[j] = <w, w>
    cudaStream_t s = c_.stream;
    ProfScope ps("lin dots", s);
    const int nb = 148 * 2;
    part_.alloc(static_cast<size_t>(nb) * (gw_.restart_alloc + 2));
    dots_.alloc(static_cast<size_t>(gw_.restart_alloc + 2));
    EP_LAUNCH(k_multi_dot, dim3(nb, j + 1), kBlock, 0, s, n, j, V, w, part_.p);
    EP_LAUNCH(k_finish_dots, (j + 1 + 7) / 8, 256, 0, s, j + 1, nb, part_.p, dots_.p);
    if (c_.dist != nullptr) {   // per-rank partial dot products, folded in rank order on every rank
      const double* g = c_.gather_to_host(dots_.p, j + 1);
      const int world = c_.dist->world();
      for (int q = 0; q <= j; ++q) {
        double t = 0.0;
        for (int r = 0; r < world; ++r) t += g[r * (j + 1) + q];
        host_out[q] = t;
      }
      return;
    }
    EP_CUDA(cudaMemcpyAsync(host_out, dots_.p, (j + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
    EP_CUDA(cudaStreamSynchronize(s));
  }

  // Flexible GMRES(m), right-preconditioned by the V-cycle.
  int fgmres(const double* b, double* x, const econpdes_options& opt, int* iters_out, double* relres_out) {
    if (opt.linear_solver == ECONPDES_LIN_BICGSTAB_MG)
      return bicgstab_run(c_.stream, lev_[0]->l, gw_, b, x, opt,
                          [&](const double* xx, const double* bb, double* yy, int mode) { spmv_top(xx, bb, yy, mode); },
                          [&](const double* v, double* z) { precondition(v, z); },
                          [&](long long n, int j, const double* V, const double* w, double* out) { dots(n, j, V, w, out); },
                          iters_out, relres_out);
    static const bool dev_gs = [] { const char* e = std::getenv("ECONPDES_DEVICE_GS"); return !(e && e[0] == '0'); }();
    if (dev_gs && (c_.dist == nullptr || c_.dist->world() == 1)) {   // several ranks fold their partial sums on the host
      if (!gw_.dots_dev)
        gw_.dots_dev = [this](long long n, int j, const double* V, const double* w, double* out) { dots_dev(n, j, V, w, out); };
    } else {
      gw_.dots_dev = nullptr;
    }
    return fgmres_run(c_.stream, lev_[0]->l, gw_, b, x, opt,
                      [&](const double* xx, const double* bb, double* yy, int mode) { spmv_top(xx, bb, yy, mode); },
                      [&](const double* v, double* z) { precondition(v, z); },
                      [&](long long n, int j, const double* V, const double* w, double* out) { dots(n, j, V, w, out); },
                      iters_out, relres_out);
  }
};

// ---------------------------------------------------------------------------------------
// Several unknowns per grid point (F > 1) on 2-D / 3-D grids
// ---------------------------------------------------------------------------------------
// The Jacobian couples the F unknowns of a point and of its stencil neighbours: planes hold F x F
// blocks, planes[((k F + fin) F + fout) l + pt] = dR_fout(pt) / dy_fin(pt + off_k).  The reference
// factorises it with sparse LU; here it is solved matrix-free by flexible GMRES preconditioned with
// alternating-direction BLOCK line relaxation (the "line-tridiagonal, ADI-style" preconditioner of
// BASELINE.json's north star): for every grid dimension in turn, the residual is updated and the
// block-tridiagonal systems along all lines of that dimension (centre blocks plus the +-1 couplings
// in that dimension, every other coupling dropped) are solved exactly by parallel cyclic reduction
// with F x F pivots (tridiag.cuh), the corrections accumulating multiplicatively.  One-level: the
// iteration count grows with the grid, which is adequate for the multi-unknown examples (grids of
// 10^3..10^5 points); the single-unknown configurations use the multigrid solver above.

// y = A x (mode 0) or y = b - A x (mode 1) with F x F blocks
template <int N, bool BOX, int F>
__global__ void __launch_bounds__(kBlock) k_spmv_block(LevelGeom<N> g, const double* __restrict__ A,
                                                       const double* __restrict__ x, const double* __restrict__ b,
                                                       double* __restrict__ y, int mode) {
  using St = Stencil<N, BOX>;
  const long long pt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pt >= g.l) return;
  long long idx[N];
  unravel_l<N>(pt, g, idx);
  double acc[F];
#pragma unroll
  for (int f = 0; f < F; ++f) acc[f] = 0.0;
#pragma unroll
  for (int k = 0; k < St::NK; ++k) {
    long long q = pt;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      const int o = St::off(k, d);
      if ((o < 0 && idx[d] == 0) || (o > 0 && idx[d] == g.S[d] - 1)) ok = false;
      q += o * g.stride[d];
    }
    const long long src = ok ? q : pt;   // entries pointing outside the grid hold 0
#pragma unroll
    for (int fin = 0; fin < F; ++fin) {
      const double xv = x[fin * g.l + src];
#pragma unroll
      for (int fout = 0; fout < F; ++fout) acc[fout] += A[(static_cast<long long>(k * F + fin) * F + fout) * g.l + pt] * xv;
    }
  }
#pragma unroll
  for (int f = 0; f < F; ++f) y[f * g.l + pt] = mode ? b[f * g.l + pt] - acc[f] : acc[f];
}

// One PCR step with stride s along the lines of one grid dimension (`stride` = distance in memory
// between neighbours of that dimension, Sd = its size): k_pcr_step (tridiag.cuh) with the row
// neighbours i -+ s*stride and the line ends as system boundaries.
template <int F>
__global__ void __launch_bounds__(kBlock) k_pcr_step_lines(long long l, long long s, long long stride, long long Sd,
                                                           PcrSystem in, PcrSystemOut out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= l) return;
  const long long id = (i / stride) % Sd;
  SmallMat<F> B = load_mat<F>(in.B, l, i);
  double d[F];
#pragma unroll
  for (int f = 0; f < F; ++f) d[f] = in.d[f * l + i];
  SmallMat<F> An, Cn;
#pragma unroll
  for (int e = 0; e < F * F; ++e) An.a[e] = Cn.a[e] = 0.0;
  if (id - s >= 0) {
    const long long im = i - s * stride;
    const SmallMat<F> A = load_mat<F>(in.A, l, i);
    const SmallMat<F> alpha = mat_mul<F>(A, mat_inv<F>(load_mat<F>(in.B, l, im)));
    const SmallMat<F> t1 = mat_mul<F>(alpha, load_mat<F>(in.A, l, im));
    const SmallMat<F> t2 = mat_mul<F>(alpha, load_mat<F>(in.C, l, im));
#pragma unroll
    for (int e = 0; e < F * F; ++e) { An.a[e] = -t1.a[e]; B.a[e] -= t2.a[e]; }
#pragma unroll
    for (int r = 0; r < F; ++r) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < F; ++k) t += alpha(r, k) * in.d[k * l + im];
      d[r] -= t;
    }
  }
  if (id + s < Sd) {
    const long long ip = i + s * stride;
    const SmallMat<F> C = load_mat<F>(in.C, l, i);
    const SmallMat<F> gamma = mat_mul<F>(C, mat_inv<F>(load_mat<F>(in.B, l, ip)));
    const SmallMat<F> t1 = mat_mul<F>(gamma, load_mat<F>(in.C, l, ip));
    const SmallMat<F> t2 = mat_mul<F>(gamma, load_mat<F>(in.A, l, ip));
#pragma unroll
    for (int e = 0; e < F * F; ++e) { Cn.a[e] = -t1.a[e]; B.a[e] -= t2.a[e]; }
#pragma unroll
    for (int r = 0; r < F; ++r) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < F; ++k) t += gamma(r, k) * in.d[k * l + ip];
      d[r] -= t;
    }
  }
  store_mat<F>(out.A, l, i, An);
  store_mat<F>(out.B, l, i, B);
  store_mat<F>(out.C, l, i, Cn);
#pragma unroll
  for (int f = 0; f < F; ++f) out.d[f * l + i] = d[f];
}

// The same step split in two, for a matrix that serves many right-hand sides (every Krylov iteration of a Newton
// step applies the line solves of the SAME Jacobian): the elimination multipliers alpha = A B_{i-s}^{-1},
// gamma = C B_{i+s}^{-1} and the next level's blocks depend on the matrix only, so they are formed ONCE per linear
// solve (k_pcr_factor_lines: two F x F inversions and four products per row and level) and a preconditioner
// application only updates the right-hand side, d <- d - alpha d_{i-s} - gamma d_{i+s} (k_pcr_apply_lines: two
// F x F mat-vecs).  Same operations on the same operands as k_pcr_step_lines, hence the same bits.
template <int F>
__global__ void __launch_bounds__(kBlock) k_pcr_factor_lines(long long l, long long s, long long stride, long long Sd,
                                                             PcrSystem in, PcrSystemOut out, double* __restrict__ alpha_out,
                                                             double* __restrict__ gamma_out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= l) return;
  const long long id = (i / stride) % Sd;
  SmallMat<F> B = load_mat<F>(in.B, l, i);
  SmallMat<F> An, Cn, alpha, gamma;
#pragma unroll
  for (int e = 0; e < F * F; ++e) An.a[e] = Cn.a[e] = alpha.a[e] = gamma.a[e] = 0.0;
  if (id - s >= 0) {
    const long long im = i - s * stride;
    const SmallMat<F> A = load_mat<F>(in.A, l, i);
    alpha = mat_mul<F>(A, mat_inv<F>(load_mat<F>(in.B, l, im)));
    const SmallMat<F> t1 = mat_mul<F>(alpha, load_mat<F>(in.A, l, im));
    const SmallMat<F> t2 = mat_mul<F>(alpha, load_mat<F>(in.C, l, im));
#pragma unroll
    for (int e = 0; e < F * F; ++e) { An.a[e] = -t1.a[e]; B.a[e] -= t2.a[e]; }
  }
  if (id + s < Sd) {
    const long long ip = i + s * stride;
    const SmallMat<F> C = load_mat<F>(in.C, l, i);
    gamma = mat_mul<F>(C, mat_inv<F>(load_mat<F>(in.B, l, ip)));
    const SmallMat<F> t1 = mat_mul<F>(gamma, load_mat<F>(in.C, l, ip));
    const SmallMat<F> t2 = mat_mul<F>(gamma, load_mat<F>(in.A, l, ip));
#pragma unroll
    for (int e = 0; e < F * F; ++e) { Cn.a[e] = -t1.a[e]; B.a[e] -= t2.a[e]; }
  }
  store_mat<F>(out.A, l, i, An);
  store_mat<F>(out.B, l, i, B);
  store_mat<F>(out.C, l, i, Cn);
  store_mat<F>(alpha_out, l, i, alpha);
  store_mat<F>(gamma_out, l, i, gamma);
}

template <int F>
__global__ void __launch_bounds__(kBlock) k_pcr_apply_lines(long long l, long long s, long long stride, long long Sd,
                                                            const double* __restrict__ alpha_in,
                                                            const double* __restrict__ gamma_in,
                                                            const double* __restrict__ din, double* __restrict__ dout) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= l) return;
  const long long id = (i / stride) % Sd;
  double d[F];
#pragma unroll
  for (int f = 0; f < F; ++f) d[f] = din[f * l + i];
  if (id - s >= 0) {
    const long long im = i - s * stride;
    const SmallMat<F> alpha = load_mat<F>(alpha_in, l, i);
#pragma unroll
    for (int r = 0; r < F; ++r) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < F; ++k) t += alpha(r, k) * din[k * l + im];
      d[r] -= t;
    }
  }
  if (id + s < Sd) {
    const long long ip = i + s * stride;
    const SmallMat<F> gamma = load_mat<F>(gamma_in, l, i);
#pragma unroll
    for (int r = 0; r < F; ++r) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < F; ++k) t += gamma(r, k) * din[k * l + ip];
      d[r] -= t;
    }
  }
#pragma unroll
  for (int f = 0; f < F; ++f) dout[f * l + i] = d[f];
}

// Binv = B^{-1} per row (end of the factorisation), and z (+)= Binv d (end of an application)
template <int F>
__global__ void __launch_bounds__(kBlock) k_block_invert(long long l, const double* __restrict__ B,
                                                         double* __restrict__ Binv) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= l) return;
  store_mat<F>(Binv, l, i, mat_inv<F>(load_mat<F>(B, l, i)));
}
template <int F>
__global__ void __launch_bounds__(kBlock) k_block_apply(long long l, const double* __restrict__ Binv,
                                                        const double* __restrict__ d, double* __restrict__ z,
                                                        int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= l) return;
  const SmallMat<F> M = load_mat<F>(Binv, l, i);
#pragma unroll
  for (int r = 0; r < F; ++r) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < F; ++k) t += M(r, k) * d[k * l + i];
    z[r * l + i] = accumulate ? z[r * l + i] + t : t;
  }
}

// z (+)= B^{-1} d
template <int F>
__global__ void __launch_bounds__(kBlock) k_pcr_final_lines(long long l, PcrSystem in, double* __restrict__ z,
                                                            int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= l) return;
  const SmallMat<F> Binv = mat_inv<F>(load_mat<F>(in.B, l, i));
#pragma unroll
  for (int r = 0; r < F; ++r) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < F; ++k) t += Binv(r, k) * in.d[k * l + i];
    z[r * l + i] = accumulate ? z[r * l + i] + t : t;
  }
}

template <int N, int F>
class BlockLineSolver final : public LinearSolver {
 public:
  explicit BlockLineSolver(Ctx& c) : c_(c) {
    if (c.dist != nullptr) throw std::runtime_error("slab-decomposed solves take single-unknown models");
    EP_CUDA(cudaMallocHost(&gw_.h, 256 * sizeof(double)));
    g_ = make_geom<N>(c.S);
  }
  ~BlockLineSolver() override {
    if (gw_.h) cudaFreeHost(gw_.h);
  }

  int solve(const double* planes, const double* b, double* x, const econpdes_options& opt, int* iters,
            double* relres) override {
    planes_ = planes;
    factored_ = false;   // a new Jacobian: the first preconditioner application factorises it
    const long long n = g_.l * F;
    if (opt.linear_solver == ECONPDES_LIN_BICGSTAB_MG)
      return bicgstab_run(c_.stream, n, gw_, b, x, opt,
                          [&](const double* xx, const double* bb, double* yy, int mode) { spmv(xx, bb, yy, mode); },
                          [&](const double* v, double* z) { precondition(v, z); },
                          [&](long long nn, int j, const double* V, const double* w, double* out) { dots(nn, j, V, w, out); },
                          iters, relres);
    static const bool dev_gs = [] { const char* e = std::getenv("ECONPDES_DEVICE_GS"); return !(e && e[0] == '0'); }();
    if (dev_gs && !gw_.dots_dev)
      gw_.dots_dev = [this](long long nn, int j, const double* V, const double* w, double* out) {
        const int nb = 148 * 2;
        part_.alloc(static_cast<size_t>(nb) * (gw_.restart_alloc + 2));
        EP_LAUNCH(k_multi_dot, dim3(nb, j + 1), kBlock, 0, c_.stream, nn, j, V, w, part_.p);
        EP_LAUNCH(k_finish_dots, (j + 1 + 7) / 8, 256, 0, c_.stream, j + 1, nb, part_.p, out);
      };
    return fgmres_run(c_.stream, n, gw_, b, x, opt,
                      [&](const double* xx, const double* bb, double* yy, int mode) { spmv(xx, bb, yy, mode); },
                      [&](const double* v, double* z) { precondition(v, z); },
                      [&](long long nn, int j, const double* V, const double* w, double* out) { dots(nn, j, V, w, out); },
                      iters, relres);
  }

  void bench(const double* planes, const double* v, int reps, double* flush, size_t flush_bytes,
             double* ms_out) override {
    (void)flush; (void)flush_bytes;
    planes_ = planes;
    factored_ = false;
    cudaStream_t s = c_.stream;
    const long long n = g_.l * F;
    DevBuf<double> y, z;
    y.alloc(static_cast<size_t>(n));
    z.alloc(static_cast<size_t>(n));
    cudaEvent_t e0, e1;
    EP_CUDA(cudaEventCreate(&e0));
    EP_CUDA(cudaEventCreate(&e1));
    auto timed = [&](auto&& fn) {
      fn();
      EP_CUDA(cudaEventRecord(e0, s));
      for (int r = 0; r < reps; ++r) fn();
      EP_CUDA(cudaEventRecord(e1, s));
      EP_CUDA(cudaEventSynchronize(e1));
      float ms = 0.f;
      EP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      return static_cast<double>(ms) / reps;
    };
    ms_out[0] = timed([&] { spmv(v, nullptr, y.p, 0); });
    ms_out[1] = timed([&] { precondition(v, z.p); });
    ms_out[2] = 0.0;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }

 private:
  Ctx& c_;
  LevelGeom<N> g_;
  const double* planes_ = nullptr;
  GmresWork gw_;
  DevBuf<double> r_, work_[2], dots_, part_;

  void spmv(const double* x, const double* b, double* y, int mode) {
    if (c_.box) EP_LAUNCH((k_spmv_block<N, true, F>), blocks_for(g_.l), kBlock, 0, c_.stream, g_, planes_, x, b, y, mode);
    else EP_LAUNCH((k_spmv_block<N, false, F>), blocks_for(g_.l), kBlock, 0, c_.stream, g_, planes_, x, b, y, mode);
  }

  void dots(long long n, int j, const double* V, const double* w, double* host_out) {
    cudaStream_t s = c_.stream;
    const int nb = 148 * 2;
    part_.alloc(static_cast<size_t>(nb) * (gw_.restart_alloc + 2));
    dots_.alloc(static_cast<size_t>(gw_.restart_alloc + 2));
    EP_LAUNCH(k_multi_dot, dim3(nb, j + 1), kBlock, 0, s, n, j, V, w, part_.p);
    EP_LAUNCH(k_finish_dots, (j + 1 + 7) / 8, 256, 0, s, j + 1, nb, part_.p, dots_.p);
    EP_CUDA(cudaMemcpyAsync(host_out, dots_.p, (j + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
    EP_CUDA(cudaStreamSynchronize(s));
  }

  // plane index of the centre and of the -1 / +1 neighbour along dimension d
  int k_of(int d, int o) const {
    if (c_.box) return (ipow3(N) - 1) / 2 + o * p3(d);
    return o == 0 ? 0 : (o < 0 ? 1 + 2 * d : 2 + 2 * d);
  }

  // Multipliers of the block cyclic reduction along every dimension, formed once per linear solve (factorize) and
  // re-used by every preconditioner application of that solve: per dimension, alpha and gamma of each reduction level
  // and the inverse of the final diagonal blocks.
  DevBuf<double> fac_[N], binv_[N], dbuf_[2];
  bool factored_ = false;
  static bool use_factors() {
    static const bool on = [] { const char* e = std::getenv("ECONPDES_BLOCK_FACTOR"); return !(e && e[0] == '0'); }();
    return on;
  }
  static int nsteps(long long Sd) {
    int k = 0;
    for (long long st = 1; st < Sd; st <<= 1) ++k;
    return k;
  }
  void factorize() {
    cudaStream_t s = c_.stream;
    const long long l = g_.l;
    const size_t blk = static_cast<size_t>(F) * F * l, sys = 3 * blk + static_cast<size_t>(F) * l;
    work_[0].alloc(sys);
    work_[1].alloc(sys);
    for (int d = 0; d < N; ++d) {
      const int ns = nsteps(g_.S[d]);
      fac_[d].alloc(static_cast<size_t>(2 * (ns > 0 ? ns : 1)) * blk);
      binv_[d].alloc(blk);
      PcrSystem in{planes_ + static_cast<size_t>(k_of(d, -1)) * blk, planes_ + static_cast<size_t>(k_of(d, 0)) * blk,
                   planes_ + static_cast<size_t>(k_of(d, +1)) * blk, nullptr};
      int cur = 0, q = 0;
      for (long long st = 1; st < g_.S[d]; st <<= 1, ++q) {
        PcrSystemOut out{work_[cur].p, work_[cur].p + blk, work_[cur].p + 2 * blk, nullptr};
        EP_LAUNCH(k_pcr_factor_lines<F>, blocks_for(l), kBlock, 0, s, l, st, g_.stride[d], g_.S[d], in, out,
                  fac_[d].p + static_cast<size_t>(2 * q) * blk, fac_[d].p + static_cast<size_t>(2 * q + 1) * blk);
        in = PcrSystem{out.A, out.B, out.C, nullptr};
        cur ^= 1;
      }
      EP_LAUNCH(k_block_invert<F>, blocks_for(l), kBlock, 0, s, l, in.B, binv_[d].p);
    }
    factored_ = true;
  }

  // z ~ A^{-1} v: one multiplicative sweep of exact block line solves over the grid dimensions
  void precondition(const double* v, double* z) {
    cudaStream_t s = c_.stream;
    const long long l = g_.l;
    const size_t blk = static_cast<size_t>(F) * F * l, sys = 3 * blk + static_cast<size_t>(F) * l;
    r_.alloc(static_cast<size_t>(F) * l);
    if (use_factors()) {
      if (!factored_) factorize();
      dbuf_[0].alloc(static_cast<size_t>(F) * l);
      dbuf_[1].alloc(static_cast<size_t>(F) * l);
      for (int d = 0; d < N; ++d) {
        const double* rhs = v;
        if (d > 0) {   // residual of the correction so far
          spmv(z, v, r_.p, 1);
          rhs = r_.p;
        }
        int cur = 0, q = 0;
        for (long long st = 1; st < g_.S[d]; st <<= 1, ++q) {
          EP_LAUNCH(k_pcr_apply_lines<F>, blocks_for(l), kBlock, 0, s, l, st, g_.stride[d], g_.S[d],
                    fac_[d].p + static_cast<size_t>(2 * q) * blk, fac_[d].p + static_cast<size_t>(2 * q + 1) * blk, rhs,
                    dbuf_[cur].p);
          rhs = dbuf_[cur].p;
          cur ^= 1;
        }
        EP_LAUNCH(k_block_apply<F>, blocks_for(l), kBlock, 0, s, l, binv_[d].p, rhs, z, d > 0 ? 1 : 0);
      }
      return;
    }
    work_[0].alloc(sys);
    work_[1].alloc(sys);
    for (int d = 0; d < N; ++d) {
      const double* rhs = v;
      if (d > 0) {   // residual of the correction so far
        spmv(z, v, r_.p, 1);
        rhs = r_.p;
      }
      PcrSystem in{planes_ + static_cast<size_t>(k_of(d, -1)) * blk, planes_ + static_cast<size_t>(k_of(d, 0)) * blk,
                   planes_ + static_cast<size_t>(k_of(d, +1)) * blk, rhs};
      int cur = 0;
      for (long long st = 1; st < g_.S[d]; st <<= 1) {
        PcrSystemOut out{work_[cur].p, work_[cur].p + blk, work_[cur].p + 2 * blk, work_[cur].p + 3 * blk};
        EP_LAUNCH(k_pcr_step_lines<F>, blocks_for(l), kBlock, 0, s, l, st, g_.stride[d], g_.S[d], in, out);
        in = PcrSystem{out.A, out.B, out.C, out.d};
        cur ^= 1;
      }
      EP_LAUNCH(k_pcr_final_lines<F>, blocks_for(l), kBlock, 0, s, l, in, z, d > 0 ? 1 : 0);
    }
  }
};

template <int N>
std::unique_ptr<LinearSolver> make_block_solver(Ctx& c) {
  switch (c.F) {
    case 2: return std::make_unique<BlockLineSolver<N, 2>>(c);
    case 3: return std::make_unique<BlockLineSolver<N, 3>>(c);
    case 4: return std::make_unique<BlockLineSolver<N, 4>>(c);
    case 5: return std::make_unique<BlockLineSolver<N, 5>>(c);
    case 6: return std::make_unique<BlockLineSolver<N, 6>>(c);
    case 7: return std::make_unique<BlockLineSolver<N, 7>>(c);
    case 8: return std::make_unique<BlockLineSolver<N, 8>>(c);
    default: throw std::invalid_argument("2-D/3-D systems take at most 8 unknowns per grid point");
  }
}

}  // namespace

std::unique_ptr<LinearSolver> make_krylov_solver(Ctx& c) {
  if (c.F != 1) {
    if (c.N == 2) return make_block_solver<2>(c);
    if (c.N == 3) return make_block_solver<3>(c);
    throw std::invalid_argument("Krylov solver handles 2-D and 3-D grids");
  }
  if (c.N == 2) return std::make_unique<KrylovSolver<2>>(c);
  if (c.N == 3) return std::make_unique<KrylovSolver<3>>(c);
  // four state variables (test/runtests.jl:294-320): FGMRES with one sweep of exact line solves per dimension
  // as the preconditioner (one level: 4-D grids are a few points per axis)
  if (c.N == 4) return std::make_unique<BlockLineSolver<4, 1>>(c);
  throw std::invalid_argument("Krylov solver handles 2-D, 3-D and 4-D grids");
}

}  // namespace econpdes
