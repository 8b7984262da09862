// NCCL binding (run-time, see dist.cuh) and the three communication patterns of the slab path.
#include "dist.cuh"

#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <vector>

namespace econpdes {
namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) {
      err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    auto sym = [&](const char* name) {
      void* p = dlsym(h, name);
      if (p == nullptr && err.empty()) err = std::string("libnccl has no symbol ") + name;
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  });
  if (!err.empty()) throw NcclError(err);
  return a;
}

#define EP_NCCL(expr)                                                                                     \
  do {                                                                                                    \
    ncclResult_t _r = (expr);                                                                             \
    if (_r != ncclSuccess)                                                                                \
      throw NcclError(std::string(#expr) + ": " + api().GetErrorString(_r) + " (" + __FILE__ + ":" +      \
                      std::to_string(__LINE__) + ")");                                                    \
  } while (0)

static_assert(sizeof(ncclUniqueId) == kUniqueIdBytes, "unique id size");

}  // namespace

void Dist::unique_id(uint8_t* id128) {
  ncclUniqueId id;
  EP_NCCL(api().GetUniqueId(&id));
  std::memcpy(id128, id.internal, kUniqueIdBytes);
}

Dist::Dist(int rank, int world, const uint8_t* id128) : rank_(rank), world_(world) {
  if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("rank/world out of range");
  if (world == 1) return;   // a single slab needs no communicator (every collective is a local copy)
  ncclUniqueId id;
  std::memcpy(id.internal, id128, kUniqueIdBytes);
  ncclComm_t comm = nullptr;
  EP_NCCL(api().CommInitRank(&comm, world, id, rank));
  comm_ = comm;
}

Dist::~Dist() {
  if (box_lo_) cudaIpcCloseMemHandle(box_lo_);
  if (box_hi_) cudaIpcCloseMemHandle(box_hi_);
  if (box_) cudaFree(box_);
  if (comm_) api().CommDestroy(static_cast<ncclComm_t>(comm_));
}

// ---- direct NVLink halo exchange ------------------------------------------------------------
// Mailbox of one rank.  Link "lo -> me" delivers into slot_from_lo, link "hi -> me" into
// slot_from_hi; each directed link numbers its transfers (epochs start at 1) and both ends count
// them identically because every rank issues the same sequence of exchanges.
struct Dist::PeerBox {
  unsigned long long data_from_lo, data_from_hi;  // written by the neighbour: epoch now in the slot
  unsigned long long ack_from_lo, ack_from_hi;    // written by the neighbour: last epoch of MINE it has consumed
  unsigned long long send_lo_epoch, send_hi_epoch, recv_lo_epoch, recv_hi_epoch;  // local counters (last used)
  unsigned int arrive[2];                         // block tickets of the two phases
  int error;                                      // 1: a wait timed out (peer gone)
  int pad_;
  unsigned long long reserved[6];                 // header = 128 bytes
  // followed by slot_from_lo[slot], slot_from_hi[slot]
};

namespace {

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// block-uniform wait until *flag >= want (thread 0 polls); false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long want, int* err) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = 1;
    const long long t0 = clock64();
    while (ld_flag(flag) < want) {
      if (*reinterpret_cast<volatile int*>(err) != 0 || clock64() - t0 > 20000000000LL) {   // ~10 s at 2 GHz
        ok = 0;
        *err = 1;
        break;
      }
      __nanosleep(20);
    }
  }
  __syncthreads();
  const bool r = ok != 0;
  __syncthreads();
  return r;
}

struct HaloArgs {
  unsigned long long* me;        // this rank's header (as u64 words, layout of PeerBox)
  unsigned long long* lo;        // neighbours' headers mapped here (nullptr at the ends)
  unsigned long long* hi;
  double* my_slot_from_lo;       // local slots
  double* my_slot_from_hi;
  double* lo_slot_from_hi;       // where my bottom plane lands on the lower neighbour
  double* hi_slot_from_lo;       // where my top plane lands on the upper neighbour
  const double* send_lo;         // planes to send (nullptr: direction inactive)
  const double* send_hi;
  double* recv_lo;               // ghost planes to fill (nullptr: direction inactive)
  double* recv_hi;
  long long count;
};
// word offsets inside the header
enum { W_DATA_LO = 0, W_DATA_HI, W_ACK_LO, W_ACK_HI, W_SEND_LO, W_SEND_HI, W_RECV_LO, W_RECV_HI, W_ARRIVE, W_ERR };

__global__ void __launch_bounds__(256) k_halo_p2p(HaloArgs a) {
  __shared__ bool last;
  unsigned int* arrive = reinterpret_cast<unsigned int*>(a.me + W_ARRIVE);
  int* err = reinterpret_cast<int*>(a.me + W_ERR);
  const unsigned long long es_lo = a.me[W_SEND_LO] + 1, es_hi = a.me[W_SEND_HI] + 1;
  const unsigned long long er_lo = a.me[W_RECV_LO] + 1, er_hi = a.me[W_RECV_HI] + 1;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  // ---- phase A: store my edge planes into the neighbours' slots ----
  if (a.send_lo != nullptr) {
    // the lower neighbour must have consumed my previous plane (it acks into MY header)
    if (wait_flag(a.me + W_ACK_LO, es_lo - 1, err))
      for (long long i = t; i < a.count; i += stride) a.lo_slot_from_hi[i] = a.send_lo[i];
  }
  if (a.send_hi != nullptr) {
    if (wait_flag(a.me + W_ACK_HI, es_hi - 1, err))
      for (long long i = t; i < a.count; i += stride) a.hi_slot_from_lo[i] = a.send_hi[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&arrive[0], 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence_system();
    if (a.send_lo != nullptr) st_flag(a.lo + W_DATA_HI, es_lo);   // I am the lower neighbour's "hi"
    if (a.send_hi != nullptr) st_flag(a.hi + W_DATA_LO, es_hi);
  }
  // ---- phase B: take what the neighbours stored into my slots ----
  if (a.recv_lo != nullptr) {
    if (wait_flag(a.me + W_DATA_LO, er_lo, err))
      for (long long i = t; i < a.count; i += stride) a.recv_lo[i] = __ldcg(a.my_slot_from_lo + i);
  }
  if (a.recv_hi != nullptr) {
    if (wait_flag(a.me + W_DATA_HI, er_hi, err))
      for (long long i = t; i < a.count; i += stride) a.recv_hi[i] = __ldcg(a.my_slot_from_hi + i);
  }
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&arrive[1], 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence_system();
    if (a.recv_lo != nullptr) { st_flag(a.lo + W_ACK_HI, er_lo); a.me[W_RECV_LO] = er_lo; }
    if (a.recv_hi != nullptr) { st_flag(a.hi + W_ACK_LO, er_hi); a.me[W_RECV_HI] = er_hi; }
    if (a.send_lo != nullptr) a.me[W_SEND_LO] = es_lo;
    if (a.send_hi != nullptr) a.me[W_SEND_HI] = es_hi;
    arrive[0] = 0u;
    arrive[1] = 0u;
    __threadfence();
  }
}

}  // namespace

void Dist::enable_p2p(long long slot_doubles, cudaStream_t s) {
  if (world_ == 1 || box_ != nullptr) return;
  if (const char* e = std::getenv("ECONPDES_P2P_HALO"))
    if (e[0] == '0') return;
  static_assert(sizeof(PeerBox) == 128, "mailbox header layout");
  const size_t bytes = sizeof(PeerBox) + 2 * static_cast<size_t>(slot_doubles) * sizeof(double);
  PeerBox* box = nullptr;
  if (cudaMalloc(&box, bytes) != cudaSuccess) { cudaGetLastError(); return; }
  cudaMemset(box, 0, bytes);
  cudaIpcMemHandle_t mine;
  int ok = cudaIpcGetMemHandle(&mine, box) == cudaSuccess ? 1 : 0;
  if (!ok) cudaGetLastError();
  // all-gather [ok flag | 64-byte handle] as 9 doubles per rank
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  double hsend[9], *dsend = nullptr, *drecv = nullptr;
  std::vector<double> hall(static_cast<size_t>(world_) * 9);
  hsend[0] = ok;
  std::memcpy(hsend + 1, &mine, 64);
  cudaMalloc(&dsend, sizeof(hsend));
  cudaMalloc(&drecv, hall.size() * sizeof(double));
  cudaMemcpyAsync(dsend, hsend, sizeof(hsend), cudaMemcpyHostToDevice, s);
  all_gather(dsend, drecv, 9, s);
  cudaMemcpyAsync(hall.data(), drecv, hall.size() * sizeof(double), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  bool all_ok = true;
  for (int r = 0; r < world_; ++r) all_ok = all_ok && hall[static_cast<size_t>(r) * 9] == 1.0;
  PeerBox *lo = nullptr, *hi = nullptr;
  int mapped = 1;
  if (all_ok) {
    auto open = [&](int r) -> PeerBox* {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, &hall[static_cast<size_t>(r) * 9 + 1], 64);
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        mapped = 0;
        return nullptr;
      }
      return static_cast<PeerBox*>(p);
    };
    if (has_lo()) lo = open(rank_ - 1);
    if (has_hi()) hi = open(rank_ + 1);
  } else {
    mapped = 0;
  }
  // every rank must take the same path: agree on success
  hsend[0] = mapped;
  cudaMemcpyAsync(dsend, hsend, sizeof(double), cudaMemcpyHostToDevice, s);
  all_gather(dsend, drecv, 1, s);
  cudaMemcpyAsync(hall.data(), drecv, static_cast<size_t>(world_) * sizeof(double), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  cudaFree(dsend);
  cudaFree(drecv);
  bool everyone = true;
  for (int r = 0; r < world_; ++r) everyone = everyone && hall[r] == 1.0;
  if (!everyone) {
    if (lo) cudaIpcCloseMemHandle(lo);
    if (hi) cudaIpcCloseMemHandle(hi);
    cudaFree(box);
    return;
  }
  box_ = box;
  box_lo_ = lo;
  box_hi_ = hi;
  slot_ = slot_doubles;
}

void Dist::halo_any(const double* send_lo, const double* send_hi, double* recv_lo, double* recv_hi, long long count,
                    cudaStream_t s) {
  if (box_ != nullptr && count <= slot_) {
    HaloArgs a{};
    auto slots = [&](PeerBox* b) { return reinterpret_cast<double*>(reinterpret_cast<char*>(b) + sizeof(PeerBox)); };
    a.me = reinterpret_cast<unsigned long long*>(box_);
    a.lo = reinterpret_cast<unsigned long long*>(box_lo_);
    a.hi = reinterpret_cast<unsigned long long*>(box_hi_);
    a.my_slot_from_lo = slots(box_);
    a.my_slot_from_hi = slots(box_) + slot_;
    a.lo_slot_from_hi = box_lo_ ? slots(box_lo_) + slot_ : nullptr;
    a.hi_slot_from_lo = box_hi_ ? slots(box_hi_) : nullptr;
    a.send_lo = has_lo() ? send_lo : nullptr;
    a.send_hi = has_hi() ? send_hi : nullptr;
    a.recv_lo = has_lo() ? recv_lo : nullptr;
    a.recv_hi = has_hi() ? recv_hi : nullptr;
    a.count = count;
    if (a.send_lo == nullptr && a.send_hi == nullptr && a.recv_lo == nullptr && a.recv_hi == nullptr) return;
    long long blocks = (count + 2047) / 2048;
    blocks = blocks < 1 ? 1 : (blocks > 32 ? 32 : blocks);
    k_halo_p2p<<<static_cast<unsigned>(blocks), 256, 0, s>>>(a);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    EP_CUDA(cudaGetLastError());
    ++np2p_;
    return;
  }
  const NcclApi& n = api();
  ncclComm_t comm = static_cast<ncclComm_t>(comm_);
  const size_t cnt = static_cast<size_t>(count);
  EP_NCCL(n.GroupStart());
  if (has_lo() && send_lo) EP_NCCL(n.Send(send_lo, cnt, ncclDouble, rank_ - 1, comm, s));
  if (has_lo() && recv_lo) EP_NCCL(n.Recv(recv_lo, cnt, ncclDouble, rank_ - 1, comm, s));
  if (has_hi() && send_hi) EP_NCCL(n.Send(send_hi, cnt, ncclDouble, rank_ + 1, comm, s));
  if (has_hi() && recv_hi) EP_NCCL(n.Recv(recv_hi, cnt, ncclDouble, rank_ + 1, comm, s));
  EP_NCCL(n.GroupEnd());
  ++ncoll_;
}

void Dist::check_p2p_error(cudaStream_t s) {
  if (box_ == nullptr) return;
  int e = 0;
  EP_CUDA(cudaMemcpyAsync(&e, &box_->error, sizeof(int), cudaMemcpyDeviceToHost, s));
  EP_CUDA(cudaStreamSynchronize(s));
  if (e != 0) throw NcclError("halo exchange timed out waiting for a neighbour rank");
}

void Dist::exchange_halo(double* v, long long plane, long long nplanes, cudaStream_t s) {
  if (world_ == 1) return;
  halo_any(v, v + (nplanes - 1) * plane, v - plane, v + nplanes * plane, plane, s);
}

void Dist::send_lo_recv_hi(const double* send_first, double* recv_hi_ghost, long long count, cudaStream_t s) {
  if (world_ == 1) return;
  halo_any(send_first, nullptr, nullptr, recv_hi_ghost, count, s);
}

void Dist::send_hi_recv_lo(const double* send_last, double* recv_lo_ghost, long long count, cudaStream_t s) {
  if (world_ == 1) return;
  halo_any(nullptr, send_last, recv_lo_ghost, nullptr, count, s);
}

void Dist::all_gather_raw(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s) {
  if (world_ == 1) {
    if (send != recv) cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, s);
    return;
  }
  EP_NCCL(api().AllGather(send, recv, bytes_per_rank, ncclChar, static_cast<ncclComm_t>(comm_), s));
  ++ncoll_;
}

void Dist::send_hi_recv_lo_raw(const void* send_last, void* recv_lo_ghost, size_t bytes, cudaStream_t s) {
  if (world_ == 1) return;
  const NcclApi& n = api();
  ncclComm_t comm = static_cast<ncclComm_t>(comm_);
  EP_NCCL(n.GroupStart());
  if (has_hi()) EP_NCCL(n.Send(send_last, bytes, ncclChar, rank_ + 1, comm, s));
  if (has_lo()) EP_NCCL(n.Recv(recv_lo_ghost, bytes, ncclChar, rank_ - 1, comm, s));
  EP_NCCL(n.GroupEnd());
  ++ncoll_;
}

void Dist::all_gather(const double* send, double* recv, long long count, cudaStream_t s) {
  if (world_ == 1) {
    if (send != recv)
      cudaMemcpyAsync(recv, send, static_cast<size_t>(count) * sizeof(double), cudaMemcpyDeviceToDevice, s);
    return;
  }
  EP_NCCL(api().AllGather(send, recv, static_cast<size_t>(count), ncclDouble, static_cast<ncclComm_t>(comm_), s));
  ++ncoll_;
}

}  // namespace econpdes

// ---- phase profile (common.cuh) ----
namespace econpdes {
namespace {
std::map<std::string, std::pair<double, long>>& prof_table() {
  static std::map<std::string, std::pair<double, long>> t;
  return t;
}
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace
bool Prof::enabled() {
  static const bool on = [] { const char* e = std::getenv("ECONPDES_PROFILE"); return e && e[0] == '1'; }();
  return on;
}
void Prof::add(const char* name, double ms) {
  auto& e = prof_table()[name];
  e.first += ms;
  e.second += 1;
}
void Prof::report(int rank) {
  if (!enabled()) return;
  for (const auto& kv : prof_table())
    std::fprintf(stderr, "[econpdes profile rank %d] %-22s %10.3f ms %8ld calls\n", rank, kv.first.c_str(),
                 kv.second.first, kv.second.second);
  prof_table().clear();
}
ProfScope::ProfScope(const char* n, cudaStream_t st) : name(n), s(st) {
  if (!Prof::enabled() || name[0] == 0) return;
  cudaStreamSynchronize(s);
  t0 = now_ms();
}
ProfScope::~ProfScope() {
  if (!Prof::enabled() || name[0] == 0) return;
  cudaStreamSynchronize(s);
  Prof::add(name, now_ms() - t0);
}
}  // namespace econpdes
