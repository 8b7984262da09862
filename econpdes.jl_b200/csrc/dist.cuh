// Multi-GPU plumbing of the slab-decomposed single-grid solve (config C5):
// one process per GPU, the slowest grid dimension cut into contiguous slabs.  What crosses
// ranks on the hot path is
//   * one ghost plane per side of every field a stencil reads  (ncclSend/ncclRecv pairs over
//     NVLink, grouped so both directions are in flight together),
//   * a handful of doubles per reduction (norms, Krylov dot products, the per-color sums of
//     squares behind FiniteDiff's step): ncclAllGather of the per-rank partials, folded in
//     rank order so every rank computes bit-identical scalars and takes identical branches.
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded when the
// host is the Python package) so the single-GPU library has no link-time NCCL dependency.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>

namespace econpdes {

constexpr int kUniqueIdBytes = 128;  // NCCL_UNIQUE_ID_BYTES

struct NcclError : std::runtime_error {
  explicit NcclError(const std::string& m) : std::runtime_error(m) {}
};

class Dist {
 public:
  static void unique_id(uint8_t* id128);  // rank 0 creates, the host broadcasts (torch.distributed, MPI, a pipe)
  Dist(int rank, int world, const uint8_t* id128);  // collective: ncclCommInitRank on the current device
  ~Dist();
  Dist(const Dist&) = delete;
  Dist& operator=(const Dist&) = delete;

  int rank() const { return rank_; }
  int world() const { return world_; }
  bool has_lo() const { return rank_ > 0; }
  bool has_hi() const { return rank_ + 1 < world_; }

  // v points at the first OWNED plane of a field stored with one ghost plane on either side:
  // [ghost | plane_0 .. plane_{nplanes-1} | ghost], each `plane` doubles.  Fills both ghosts
  // from the neighbours' edge planes (nothing is written at the ends of the global grid).
  void exchange_halo(double* v, long long plane, long long nplanes, cudaStream_t s);
  // one-sided variants used by the multigrid transfers
  void send_lo_recv_hi(const double* send_first, double* recv_hi_ghost, long long count, cudaStream_t s);
  void send_hi_recv_lo(const double* send_last, double* recv_lo_ghost, long long count, cudaStream_t s);
  // recv[r*count .. (r+1)*count) = rank r's send[0..count)   (device pointers)
  void all_gather(const double* send, double* recv, long long count, cudaStream_t s);
  // untyped variants for the multigrid set-up (operator planes in the hierarchy's storage type)
  void all_gather_raw(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
  void send_hi_recv_lo_raw(const void* send_last, void* recv_lo_ghost, size_t bytes, cudaStream_t s);

  int64_t collectives() const { return ncoll_; }       // NCCL calls
  int64_t p2p_exchanges() const { return np2p_; }       // halo exchanges done by k_halo_p2p
  bool p2p_enabled() const { return box_ != nullptr; }

  // Direct NVLink halo path: every rank exposes a small mailbox (flags + one plane-sized slot per
  // neighbour) through CUDA IPC; ONE kernel per exchange stores this rank's edge planes into the
  // neighbours' slots, publishes them with a system-scope flag, waits for the neighbours' flags
  // and moves the received planes into the ghost storage (dist.cu: k_halo_p2p).  Collective;
  // `slot_doubles` = largest plane that will be exchanged.  Falls back to NCCL send/recv when
  // IPC is unavailable or ECONPDES_P2P_HALO=0.
  void enable_p2p(long long slot_doubles, cudaStream_t s);
  void check_p2p_error(cudaStream_t s);   // throws if a mailbox wait timed out

 private:
  struct PeerBox;
  void halo_any(const double* send_lo, const double* send_hi, double* recv_lo, double* recv_hi, long long count,
                cudaStream_t s);
  int rank_, world_;
  void* comm_ = nullptr;
  int64_t ncoll_ = 0, np2p_ = 0;
  PeerBox* box_ = nullptr;      // this rank's mailbox (device memory, IPC-exported)
  PeerBox* box_lo_ = nullptr;   // neighbours' mailboxes mapped into this process
  PeerBox* box_hi_ = nullptr;
  long long slot_ = 0;
};

}  // namespace econpdes
