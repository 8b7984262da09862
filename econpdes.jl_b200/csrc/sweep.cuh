// Marching sweeps for 2-D / 3-D grids: the residual `pde!` (src/pdesolve.jl:436-462, fused with the
// implicit-Euler / MCP wrapper of src/finiteschemesolve.jl:158-162, 197-203) and the fused colored
// finite-difference Jacobian (FiniteDiff, called at src/finiteschemesolve.jl:184), restructured so that the
// per-point overhead of the one-thread-per-point kernels (kernels.cuh) is paid once per thread:
//
//   * a thread owns one COLUMN (fixed indices in dimensions 0..N-2; consecutive threads = consecutive
//     dimension-0 indices, so every load and store is coalesced) and marches along the slowest dimension
//     N-1 over a tile of L planes, keeping the three planes of its +-1 box in a sliding register window:
//     3^(N-1) new loads per point instead of 3^N, through neighbour offsets computed once per tile;
//   * index arithmetic, face flags, the spacing records and reciprocals of dimensions 0..N-2 and the
//     cross-stencil divisors among them are formed once per tile; the records of dimension N-1 are the
//     same for the whole block and are staged in shared memory once per tile;
//   * TAME variants (1; 2 = also no boundary derivative values, the default) drop the NaN-select of every stencil quotient (mdiv_tame, stencil.cuh): the host has
//     checked that all grid reciprocals and boundary values are finite and of moderate magnitude, and the
//     residual sweep itself reports whether every unknown is (|y| < 2^199); if not, the host repeats the
//     sweep with the select-carrying variant.  Same bits either way;
//   * the row terms are compile-time variants: R = G, the scalar-step implicit residual, and the general
//     form (per-row signed steps, algebraic rows, bounds).
// Tiles of neighbouring columns and the same planes run concurrently, so the halo columns/planes a tile
// re-reads are L1/L2 hits: DRAM traffic stays near the algorithmic 16F (24F) bytes per point.
#pragma once
#include "kernels.cuh"

namespace econpdes {

constexpr int kMarchMaxL = 64;   // planes per tile (shared-memory staging of the last dimension's records)

struct MarchPlan {
  int L;            // planes per tile
  int nseg;         // tiles along dimension N-1
  int ncolblocks;   // tiles across the columns
  long long ncol;   // columns = prod S[0..N-2]
  long long SL;     // planes of dimension N-1 held on this device
};

inline MarchPlan make_march_plan(long long ncol, long long SL, int block, int resident_blocks, int force_L) {
  MarchPlan mp;
  mp.ncol = ncol;
  mp.SL = SL;
  mp.ncolblocks = static_cast<int>((ncol + block - 1) / block);
  long long L = force_L;
  if (L <= 0) {   // about six tiles per resident block, so the tail of the last wave stays short
    const long long want_seg = (6LL * resident_blocks + mp.ncolblocks - 1) / mp.ncolblocks;
    L = SL / (want_seg > 0 ? want_seg : 1);
    if (L < 8) L = 8;
  }
  if (L > kMarchMaxL) L = kMarchMaxL;
  if (L > SL) L = SL;
  if (L < 1) L = 1;
  mp.L = static_cast<int>(L);
  mp.nseg = static_cast<int>((SL + L - 1) / L);
  return mp;
}

// bit 1 of the flag word of a norm-producing sweep: some unknown is not tame
constexpr int kFlagNaN = 1, kFlagUntame = 2;
__device__ __forceinline__ bool is_tame(double v) {
  return (__double2hiint(v) & 0x7ff00000) < 0x4c600000;   // |v| < 2^199 (and neither Inf nor NaN)
}

// Thread-constant part of a column + the per-step part of the geometry.
template <class M>
struct MarchCol {
  static constexpr int N = M::N, F = M::F, KP = ipow3(M::N - 1), K = ipow3(M::N);
  long long idx[N];
  PointGeom<N, 2> g;
  double st[N];
  int poff[KP];   // offset of the in-plane neighbour kp (0 = the point itself when it lies outside the grid)

  __device__ __forceinline__ void setup(const GridView<N>& gv, long long col) {
    unsigned p = static_cast<unsigned>(col);
#pragma unroll
    for (int d = 0; d + 2 < N; ++d) {
      const unsigned q = gv.div[d].div(p);
      idx[d] = p - q * static_cast<unsigned>(gv.S[d]);
      p = q;
    }
    idx[N - 2] = p;
#pragma unroll
    for (int d = 0; d + 1 < N; ++d) load_axis(gv, d, idx[d], g, st);
    finish_geom<1, false>(g);
#pragma unroll
    for (int kp = 0; kp < KP; ++kp) {
      int q = 0;
      bool ok = true;
#pragma unroll
      for (int d = 0; d + 1 < N; ++d) {
        const int o = (kp / p3(d)) % 3 - 1;
        if ((o < 0 && g.lo[d]) || (o > 0 && g.hi[d])) ok = false;
        q += o * static_cast<int>(gv.stride[d]);
      }
      poff[kp] = ok ? q : 0;
    }
  }
  // geometry of plane J (global index) of dimension N-1 from the staged record
  __device__ __forceinline__ void step(const GridView<N>& gv, long long J, const double* rec, double x) {
    constexpr int d = N - 1;
    idx[d] = J;
    g.lo[d] = J == 0;
    g.hi[d] = J == gv.S[d] - 1;
    st[d] = x;
    const Pair* t = reinterpret_cast<const Pair*>(rec);
    const Pair t0 = t[0], t1 = t[1], t2 = t[2], t3 = t[3];
    g.dm[d] = t0.a; g.dp[d] = t0.b; g.d[d] = t1.a;
    g.rdp[d] = t1.b; g.rdm[d] = t2.a; g.rdpd[d] = t2.b; g.rdmd[d] = t3.a; g.rdpm[d] = t3.b;
    g.dpd[d] = g.dp[d] * g.d[d];
    g.dmd[d] = g.dm[d] * g.d[d];
    g.dpm[d] = g.dp[d] * g.dm[d];
    finish_geom<2, true>(g);
  }
};

// stage the records and coordinates of planes [j0, j1) of dimension N-1 in shared memory.  They are the same
// for every column; each WARP keeps its own copy so a tile starts without a block-wide barrier.
template <int N>
__device__ __forceinline__ void stage_last_axis(const GridView<N>& gv, int j0, int j1, double* s_tab, double* s_x) {
  const double* tab = gv.tab[N - 1] + (j0 + gv.off) * kGeomRows;
  const double* xs = gv.x[N - 1] + (j0 + gv.off);
  const int lane = threadIdx.x & 31;
  __syncwarp();
  for (int e = lane; e < (j1 - j0) * kGeomRows; e += 32) s_tab[e] = tab[e];
  for (int e = lane; e < j1 - j0; e += 32) s_x[e] = xs[e];
  __syncwarp();
}


// ROWKIND: 0 = R = G; 1 = G - (ypost - y)/delta with one finite scalar delta, no algebraic rows, no bounds;
//          2 = general (RowSpec as in k_residual)
// UNR: unroll factor of the marching loop (3 lets the compiler rename the window instead of moving it; the
//      27-value window of a 3-D box is better left rolled: register pressure)
// Tried on B200 and dropped (2048^2 / 256^3 sweeps): an explicit register prefetch of the plane two steps
// ahead and L2 prefetches further ahead (no gain: the loads of the next plane are already issued a step early);
// two columns per thread, to share the plane-only arithmetic and interleave two evaluations (slower: 168
// registers, three blocks per SM); 96- and 80-register builds for five / six blocks per SM (spills).
template <class M, int ROWKIND, int TAME, int UNR, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_residual_march(GridView<M::N> gv, ParamPack prm, const double* __restrict__ y, const double* __restrict__ bc,
                 RowSpec rs, double* __restrict__ out, NormWork nw, int want_norm, MarchPlan mp) {
  constexpr int N = M::N, F = M::F, KP = ipow3(N - 1), K = ipow3(N), C = (K - 1) / 2;
  static_assert(N >= 2, "marching sweeps are for 2-D / 3-D grids");
  __shared__ double scratch[96];
  __shared__ __align__(16) double s_tab_all[BLOCK / 32][kMarchMaxL * kGeomRows];
  __shared__ double s_x_all[BLOCK / 32][kMarchMaxL];
  double* s_tab = s_tab_all[threadIdx.x >> 5];
  double* s_x = s_x_all[threadIdx.x >> 5];
  double s = 0.0, m = 0.0;
  int flag = 0;
  const int ntiles = mp.ncolblocks * mp.nseg;
  const long long ncol = mp.ncol, l = gv.l;
  // last readable plane: the grid's last plane, or the upper ghost plane of an inner slab
  const int jmax = static_cast<int>(gv.S[N - 1] - 1 - gv.off < mp.SL ? gv.S[N - 1] - 1 - gv.off : mp.SL);
  const int jmin = gv.off > 0 ? -1 : 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int cb = tile % mp.ncolblocks, seg = tile / mp.ncolblocks;
    const long long col = static_cast<long long>(cb) * BLOCK + threadIdx.x;
    const int j0 = seg * mp.L;
    const int j1 = j0 + mp.L < static_cast<int>(mp.SL) ? j0 + mp.L : static_cast<int>(mp.SL);
    stage_last_axis<N>(gv, j0, j1, s_tab, s_x);
    if (col >= ncol) continue;
    MarchCol<M> mc;
    mc.setup(gv, col);
    const double* __restrict__ yb = y + col;
    double wa[F][KP], wb[F][KP], wc[F][KP];
    auto load_plane = [&](int j, double (*dst)[KP]) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const double* __restrict__ yp = yb + f * l + static_cast<long long>(j) * ncol;
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) dst[f][kp] = yp[mc.poff[kp]];
      }
    };
    load_plane(j0 - 1 > jmin ? j0 - 1 : jmin, wa);
    load_plane(j0, wb);
    load_plane(j0 + 1 < jmax ? j0 + 1 : jmax, wc);
#pragma unroll UNR
    for (int j = j0; j < j1; ++j) {
      mc.step(gv, j + gv.off, s_tab + (j - j0) * kGeomRows, s_x[j - j0]);
      typename M::Pre pre;
      M::prepare(prm.v, mc.st, pre);
      double w[F][K], bl[F][N], bh[F][N];
      const long long pt = col + static_cast<long long>(j) * ncol;
#pragma unroll
      for (int f = 0; f < F; ++f) {
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) {
          w[f][kp] = wa[f][kp];
          w[f][KP + kp] = wb[f][kp];
          w[f][2 * KP + kp] = wc[f][kp];
        }
        if (TAME != 2) load_bc<N>(bc != nullptr ? bc + f * l : nullptr, pt, mc.idx, gv, mc.g, bl[f], bh[f]);
      }
      double o[F];
      eval_point<M, false, TAME>(pre, mc.g, w, bl, bh, o, nullptr);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const long long row = f * l + pt;
        const double yc = w[f][C];
        double r;
        if (ROWKIND == 0) {
          r = o[f];
        } else if (ROWKIND == 1) {
          r = o[f] - mdiv(rs.ypost[row] - yc, rs.delta, rs.rdelta);
        } else {
          r = apply_row_terms(o[f], yc, make_row_terms(rs, row));
        }
        out[row] = r;
        if (!is_tame(yc)) flag |= kFlagUntame;
        if (r != r) {
          flag |= kFlagNaN;
        } else {
          s += r * r;
          const double a = fabs(r);
          m = a > m ? a : m;
        }
      }
      // slide the window
#pragma unroll
      for (int f = 0; f < F; ++f) {
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) {
          wa[f][kp] = wb[f][kp];
          wb[f][kp] = wc[f][kp];
        }
      }
      load_plane(j + 2 < jmax ? j + 2 : jmax, wc);
    }
  }
  if (want_norm) {
    // a NaN anywhere makes the 2-norm NaN, as in norm(ydot) (finiteschemesolve.jl:371,415)
    grid_norm_finish(s, m, flag, nw, scratch);
  }
}

// Fused colored FD Jacobian (see k_jacobian, kernels.cuh), marching.  POLICY as in jacobian_point.
template <class M, bool BOX, int POLICY, int TAME, int ROWKIND, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_jacobian_march(GridView<M::N> gv, ParamPack prm, const double* __restrict__ y, const double* __restrict__ bc,
                 RowSpec rs, const double* __restrict__ eps_table, double* __restrict__ planes,
                 double* __restrict__ r0out, MarchPlan mp) {
  constexpr int N = M::N, F = M::F, KP = ipow3(N - 1), K = ipow3(N);
  static_assert(N >= 2, "marching sweeps are for 2-D / 3-D grids");
  __shared__ __align__(16) double s_tab_all[BLOCK / 32][kMarchMaxL * kGeomRows];
  __shared__ double s_x_all[BLOCK / 32][kMarchMaxL];
  double* s_tab = s_tab_all[threadIdx.x >> 5];
  double* s_x = s_x_all[threadIdx.x >> 5];
  const int ntiles = mp.ncolblocks * mp.nseg;
  const long long ncol = mp.ncol, l = gv.l;
  const int jmax = static_cast<int>(gv.S[N - 1] - 1 - gv.off < mp.SL ? gv.S[N - 1] - 1 - gv.off : mp.SL);
  const int jmin = gv.off > 0 ? -1 : 0;
  const double* reps_table = POLICY == 0 ? eps_table + ncolors_grid<N>(gv) * F : nullptr;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int cb = tile % mp.ncolblocks, seg = tile / mp.ncolblocks;
    const long long col = static_cast<long long>(cb) * BLOCK + threadIdx.x;
    const int j0 = seg * mp.L;
    const int j1 = j0 + mp.L < static_cast<int>(mp.SL) ? j0 + mp.L : static_cast<int>(mp.SL);
    stage_last_axis<N>(gv, j0, j1, s_tab, s_x);
    if (col >= ncol) continue;
    MarchCol<M> mc;
    mc.setup(gv, col);
    const double* __restrict__ yb = y + col;
    double wa[F][KP], wb[F][KP], wc[F][KP];
    auto load_plane = [&](int j, double (*dst)[KP]) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const double* __restrict__ yp = yb + f * l + static_cast<long long>(j) * ncol;
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) dst[f][kp] = yp[mc.poff[kp]];
      }
    };
    load_plane(j0 - 1 > jmin ? j0 - 1 : jmin, wa);
    load_plane(j0, wb);
    load_plane(j0 + 1 < jmax ? j0 + 1 : jmax, wc);
#pragma unroll 1
    for (int j = j0; j < j1; ++j) {
      mc.step(gv, j + gv.off, s_tab + (j - j0) * kGeomRows, s_x[j - j0]);
      typename M::Pre pre;
      M::prepare(prm.v, mc.st, pre);
      double w[F][K], bl[F][N], bh[F][N];
      RowTerms rt[F];
      const long long pt = col + static_cast<long long>(j) * ncol;
#pragma unroll
      for (int f = 0; f < F; ++f) {
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) {
          w[f][kp] = wa[f][kp];
          w[f][KP + kp] = wb[f][kp];
          w[f][2 * KP + kp] = wc[f][kp];
        }
        if (TAME != 2) load_bc<N>(bc != nullptr ? bc + f * l : nullptr, pt, mc.idx, gv, mc.g, bl[f], bh[f]);
        if (ROWKIND == 2) {
          rt[f] = make_row_terms(rs, f * l + pt);
        } else if (ROWKIND == 1) {
          rt[f].ypost = rs.ypost[f * l + pt];
          rt[f].delta = rs.delta;
          rt[f].rdelta = rs.rdelta;
        }
      }
      double r0[F];
      jacobian_point<M, BOX, false, POLICY, TAME, ROWKIND, true>(pre, mc.g, mc.idx, gv, w, bl, bh, rt, eps_table, 0, r0,
                                                  [&](int k, int fin, int fout, double v) {
                                                    planes[(static_cast<long long>(k * F + fin) * F + fout) * l + pt] = v;
                                                  }, reps_table);
      if (r0out != nullptr) {
#pragma unroll
        for (int f = 0; f < F; ++f) r0out[f * l + pt] = r0[f];
      }
#pragma unroll
      for (int f = 0; f < F; ++f) {
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) {
          wa[f][kp] = wb[f][kp];
          wb[f][kp] = wc[f][kp];
        }
      }
      load_plane(j + 2 < jmax ? j + 2 : jmax, wc);
    }
  }
}

}  // namespace econpdes
