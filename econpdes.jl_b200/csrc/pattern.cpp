// Host-only set-up: Jacobian sparsity pattern and its closed-form distance-2 coloring,
// bit-exact (1-based Int64, rows sorted inside each column) with the reference's
// `local_stencil_jacobian` (src/utils/coloring.jl:18-35) and `stencil_colors` (:52-70).
// One-off O(nnz) work; needs no GPU.
#include <cstdint>
#include <vector>

#include "../../include/econpdes_cuda.h"
#include "errors.h"

namespace {

bool check_shape(int ndims, const int64_t* size, int nfuncs) {
  if (ndims < 1 || ndims > 8 || size == nullptr || nfuncs < 1) {
    econpdes::set_error("state grid must contain at least one state variable and one unknown");
    return false;
  }
  for (int d = 0; d < ndims; ++d) {
    if (size[d] < 2) {  // stategrid.jl:16
      econpdes::set_error("state grid entries must have at least 2 points");
      return false;
    }
  }
  return true;
}

}  // namespace

extern "C" {

int econpdes_stencil_colors(int ndims, const int64_t* size, int nfuncs, int64_t* colors_out) {
  if (!check_shape(ndims, size, nfuncs) || colors_out == nullptr) return ECONPDES_E_INVALID;
  int64_t l = 1;
  for (int d = 0; d < ndims; ++d) l *= size[d];
  std::vector<int64_t> idx(ndims, 0);
  int64_t ncg = 1;
  for (int d = 0; d < ndims; ++d) ncg *= size[d] < 3 ? size[d] : 3;
  for (int64_t i = 0; i < l; ++i) {
    int64_t base = 0, stride = 1;
    for (int d = 0; d < ndims; ++d) {
      base += (idx[d] % 3) * stride;
      stride *= size[d] < 3 ? size[d] : 3;
    }
    colors_out[i] = base + 1;
    for (int d = 0; d < ndims; ++d) {  // column-major increment
      if (++idx[d] < size[d]) break;
      idx[d] = 0;
    }
  }
  for (int f = 1; f < nfuncs; ++f)
    for (int64_t j = 0; j < l; ++j) colors_out[f * l + j] = colors_out[j] + f * ncg;
  return ECONPDES_OK;
}

int64_t econpdes_stencil_nnz(int ndims, const int64_t* size, int nfuncs) {
  if (!check_shape(ndims, size, nfuncs)) return ECONPDES_E_INVALID;
  int64_t nnz = static_cast<int64_t>(nfuncs) * nfuncs;
  for (int d = 0; d < ndims; ++d) nnz *= 3 * size[d] - 2;
  return nnz;
}

int econpdes_stencil_pattern(int ndims, const int64_t* size, int nfuncs, int64_t* colptr, int64_t* rowval) {
  if (!check_shape(ndims, size, nfuncs) || colptr == nullptr || rowval == nullptr) return ECONPDES_E_INVALID;
  int64_t l = 1;
  std::vector<int64_t> stride(ndims);
  for (int d = 0; d < ndims; ++d) {
    stride[d] = l;
    l *= size[d];
  }
  // column (c, fin) holds rows (ci, fout) for every ci in the clipped +-1 box of c: the
  // box relation is symmetric.  Sorted order = fout-major, then column-major grid point.
  std::vector<int64_t> idx(ndims, 0), lo(ndims), hi(ndims), cur(ndims);
  int64_t pos = 0;
  colptr[0] = 1;
  for (int fin = 0; fin < nfuncs; ++fin) {
    for (int d = 0; d < ndims; ++d) idx[d] = 0;
    for (int64_t c = 0; c < l; ++c) {
      for (int d = 0; d < ndims; ++d) {
        lo[d] = idx[d] > 0 ? idx[d] - 1 : 0;
        hi[d] = idx[d] < size[d] - 1 ? idx[d] + 1 : size[d] - 1;
      }
      for (int fout = 0; fout < nfuncs; ++fout) {
        for (int d = 0; d < ndims; ++d) cur[d] = lo[d];
        while (true) {
          int64_t lin = 0;
          for (int d = 0; d < ndims; ++d) lin += cur[d] * stride[d];
          rowval[pos++] = lin + fout * l + 1;
          int d = 0;
          for (; d < ndims; ++d) {
            if (++cur[d] <= hi[d]) break;
            cur[d] = lo[d];
          }
          if (d == ndims) break;
        }
      }
      colptr[fin * l + c + 1] = pos + 1;
      for (int d = 0; d < ndims; ++d) {
        if (++idx[d] < size[d]) break;
        idx[d] = 0;
      }
    }
  }
  return ECONPDES_OK;
}

}  // extern "C"
