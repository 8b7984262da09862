// CUDA plumbing shared by the translation units: error checks, launch accounting,
// deterministic block/grid reductions.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <stdexcept>
#include <string>

#include "errors.h"

namespace econpdes {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

#define EP_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw ::econpdes::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                  std::to_string(__LINE__) + ")");                                 \
  } while (0)

// Every kernel launch of the library goes through EP_LAUNCH so `gpu_launches` in bench.py
// is a count, not a guess.
extern std::atomic<int64_t> g_kernel_launches;
#define EP_LAUNCH(kernel, grid, block, smem, stream, ...)        \
  do {                                                           \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);  \
    ::econpdes::g_kernel_launches.fetch_add(1, std::memory_order_relaxed); \
    EP_CUDA(cudaGetLastError());                                 \
  } while (0)

// Phase breakdown of a solve (ECONPDES_PROFILE=1): wall time between stream synchronisations,
// accumulated per phase name and printed to stderr at the end of Solver::solve.  Off by default
// (the synchronisations serialise host and device).
struct Prof {
  static bool enabled();
  static void add(const char* name, double ms);
  static void report(int rank);
};
struct ProfScope {
  const char* name;
  cudaStream_t s;
  double t0 = 0.0;
  ProfScope(const char* n, cudaStream_t st);
  ~ProfScope();
};

constexpr int kBlock = 256;
inline int blocks_for(long long n, int block = kBlock) { return static_cast<int>((n + block - 1) / block); }
// grid of a grid-stride sweep that ends in a grid-wide reduction: at most 16 blocks per SM of a B200
inline int sweep_blocks(long long n, int block = kBlock) {
  const int b = blocks_for(n, block);
  return b < 148 * 16 ? b : 148 * 16;
}

#if defined(__CUDACC__)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}
__device__ __forceinline__ int warp_or(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide (sum, max, or) in a fixed order: xor-butterfly inside each warp, then warp 0
// folds the per-warp values.  Result valid in every thread.  `scratch` needs 3*32 doubles.
__device__ __forceinline__ void block_reduce3(double& s, double& m, int& f, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  s = warp_sum(s);
  m = warp_max(m);
  f = warp_or(f);
  __syncthreads();
  if (lane == 0) {
    scratch[warp] = s;
    scratch[32 + warp] = m;
    scratch[64 + warp] = static_cast<double>(f);
  }
  __syncthreads();
  double s2 = lane < nw ? scratch[lane] : 0.0;
  double m2 = lane < nw ? scratch[32 + lane] : 0.0;
  int f2 = lane < nw ? static_cast<int>(scratch[64 + lane]) : 0;
  s = warp_sum(s2);
  m = warp_max(m2);
  f = warp_or(f2);
}
#endif

// Work area of a grid-wide norm: per-block partials + a ticket counter; the last block to
// finish folds the partials in block order, so the result does not depend on scheduling.
// out[0] = sum of squares, out[1] = max |.| over the non-NaN entries, out[2] = 1 if any NaN.
struct NormWork {
  double* part;       // [3][max_blocks]
  unsigned* counter;  // zero-initialised, reset by the finishing block
  double* out;        // [3]
  int max_blocks;
};

#if defined(__CUDACC__)
__device__ __forceinline__ void grid_norm_finish(double s, double m, int f, const NormWork& nw, double* scratch) {
  block_reduce3(s, m, f, scratch);
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    nw.part[blockIdx.x] = s;
    nw.part[nw.max_blocks + blockIdx.x] = m;
    nw.part[2 * nw.max_blocks + blockIdx.x] = static_cast<double>(f);
    __threadfence();
    const unsigned ticket = atomicAdd(nw.counter, 1u);
    is_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double s2 = 0.0, m2 = 0.0;
    int f2 = 0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
      s2 += nw.part[b];
      const double mm = nw.part[nw.max_blocks + b];
      m2 = mm > m2 ? mm : m2;
      f2 |= static_cast<int>(nw.part[2 * nw.max_blocks + b]);
    }
    block_reduce3(s2, m2, f2, scratch);
    if (threadIdx.x == 0) {
      nw.out[0] = s2;
      nw.out[1] = m2;
      nw.out[2] = static_cast<double>(f2);
      *nw.counter = 0u;
    }
  }
}
#endif

}  // namespace econpdes
